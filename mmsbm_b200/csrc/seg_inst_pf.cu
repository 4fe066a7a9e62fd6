// segment_pass_kernel instantiations with software prefetch (PF = 1) for the row widths of the
// BASELINE configurations: 12 doubles (K, L = 10: G = 3), 20 (G = 5), 32 (G = 8)
#include "segment_pass.cuh"
namespace mmsbm {
int launch_segment_pass_pf(const SegArgs& a, int G, int UN, int MINB, int RUNS, dim3 grid, size_t smem, cudaStream_t st) {
  MMSBM_SEG_LAUNCH_P(5, 1, 3, 3, 6, 1, RUNS == 6)
  MMSBM_SEG_LAUNCH_P(3, 1, 3, 3, 2, 1, RUNS == 2) MMSBM_SEG_LAUNCH_P(5, 1, 3, 3, 2, 1, RUNS == 2)
  MMSBM_SEG_LAUNCH_P(8, 1, 3, 3, 2, 1, RUNS == 2)
  MMSBM_SEG_LAUNCH_P(3, 1, 3, 3, 1, 1, RUNS == 1) MMSBM_SEG_LAUNCH_P(5, 1, 3, 3, 1, 1, RUNS == 1)
  MMSBM_SEG_LAUNCH_P(8, 1, 3, 3, 1, 1, RUNS == 1)
  return MMSBM_ERANGE;
}
}  // namespace mmsbm
