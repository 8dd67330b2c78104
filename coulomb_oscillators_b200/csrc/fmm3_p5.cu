// order-5 instantiation of the FMM operator passes (see fmm3_order.cuh)
#include "fmm3_order.cuh"
namespace nbco { NBCO_INSTANTIATE_ORDER(5) }
