// SPD(8) instantiation of the acquisition kernels (one translation unit per matrix size: parallel nvcc jobs).
#include "acq_spd.cuh"

namespace gabo {
template int launch_acq_spd<8>(const gabo_gp_desc*, double*, int64_t, const gabo_rcg_opts*, double*, double*, int32_t*,
                               int32_t*, cudaStream_t);
}  // namespace gabo
