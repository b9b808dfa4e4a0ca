// zorro (src/curve/zorro/g1.rs) instantiation of the MSM kernels.
#define BP_MSM_INSTANTIATE
#include "msm_kernels.cuh"
namespace bp {
template int msm_run<Zorro>(bp_ctx*, const affine*, const fe*, size_t, uint8_t*, int*);
template int msm_run_job<Zorro>(bp_ctx*, const MsmJob&, uint8_t (*)[64], int*);
template int synth_points_run<Zorro>(bp_ctx*, void*, size_t, uint64_t);
template int msm_run_streamed<Zorro>(bp_ctx*, const uint8_t*, const uint8_t*, size_t, const std::vector<size_t>&,
                                         const std::vector<size_t>&, uint8_t*, int*, const affine*);
}  // namespace bp
