// segment_pass_kernel instantiations with 4 / 8 chunks per lane (rows of 65..256 doubles)
#include "segment_pass.cuh"
namespace mmsbm {
int launch_segment_pass_ch4(const SegArgs& a, int G, int UN, int MINB, dim3 grid, size_t smem, cudaStream_t st) {
  MMSBM_SEG_LAUNCH(5, 4, 1, 1) MMSBM_SEG_LAUNCH(6, 4, 1, 1) MMSBM_SEG_LAUNCH(7, 4, 1, 1) MMSBM_SEG_LAUNCH(8, 4, 1, 1)
  return MMSBM_ERANGE;
}
int launch_segment_pass_ch8(const SegArgs& a, int G, int UN, int MINB, dim3 grid, size_t smem, cudaStream_t st) {
  MMSBM_SEG_LAUNCH(5, 8, 1, 1) MMSBM_SEG_LAUNCH(6, 8, 1, 1) MMSBM_SEG_LAUNCH(7, 8, 1, 1) MMSBM_SEG_LAUNCH(8, 8, 1, 1)
  return MMSBM_ERANGE;
}
}  // namespace mmsbm
