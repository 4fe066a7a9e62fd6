// Shared helpers of libmmsbm_b200 (sm_100a only).
#pragma once
#include <assert.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmsbm_b200.h"

// Device-side bounds checks of every indirect access, compiled in by
// `python -m mmsbm_b200.build --check` (libmmsbm_b200_check.so; compute-sanitizer is not
// available on the GPU pool).  A failed check traps the kernel and surfaces as a CUDA error.
#ifdef MMSBM_BOUNDS_CHECK
#define MMSBM_DEV_CHECK(cond) assert(cond)
#else
#define MMSBM_DEV_CHECK(cond) ((void)0)
#endif

namespace mmsbm {

constexpr double kEps = 2.220446049250313e-16;  // np.finfo(float).eps, src/kernels_numpy.py:51
constexpr unsigned kFull = 0xffffffffu;

// thread-local error text + launch counter (bench.py reports gpu_launches from it)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// device row stride of theta / eta: rows are whole 32-byte chunks (256-bit loads)
inline int row_stride(int x) { return (x + 3) & ~3; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

#define MMSBM_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::mmsbm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                         __FILE__, __LINE__);                                         \
      return (int)e__;                                                                \
    }                                                                                 \
  } while (0)

#define MMSBM_LAUNCH_CHECK(name)                                                      \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    ::mmsbm::count_launch();                                                          \
    if (e__ != cudaSuccess) {                                                         \
      ::mmsbm::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));   \
      return (int)e__;                                                                \
    }                                                                                 \
  } while (0)

#define MMSBM_REQUIRE(cond, code, ...)                                                \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::mmsbm::set_error(__VA_ARGS__);                                                \
      return (code);                                                                  \
    }                                                                                 \
  } while (0)

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (off + bytes > cap) return nullptr;
    T* out = reinterpret_cast<T*>(base + off);
    off += bytes;
    return out;
  }
};

// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }  // evict-first

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

}  // namespace mmsbm
