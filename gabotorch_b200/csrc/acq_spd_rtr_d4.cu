// SPD(4) instantiation of the trust-region kernel (one translation unit per matrix size: parallel nvcc jobs).
#include "acq_spd_rtr.cuh"

namespace gabo {
template int launch_rtr_spd<4>(const gabo_gp_desc*, double*, int64_t, const CtrParams&, double*, int32_t*, int32_t*,
                               cudaStream_t);
}  // namespace gabo
