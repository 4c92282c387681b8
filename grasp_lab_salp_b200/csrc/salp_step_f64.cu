// salp_step_f64.cu -- SALP_PRECISION_F64 instantiation of the step kernel ("reference mode").
// Built with -fmad=false: every float64 multiply and add is rounded separately, like numpy's.
#include "salp_step_kernel.cuh"

void salp_launch_step_f64(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags,
                          const int32_t* order, cudaStream_t stream) {
  const int block = 128;
  salp_step_kernel<SALP_PRECISION_F64><<<grid_for(v.n, block), block, 0, stream>>>(p, make_derived(p), v, io, flags, order);
}
