#include "vitb_registry.h"
namespace vitb {
// any generator polynomials, any rate up to GENERIC_MAX_R (acs_generic.cuh): one instantiation per constraint length
void register_generic(std::vector<KernelEntry>& v) {
    VITB_GENERIC_VARIANTS(v, 3, "K3,any")
    VITB_GENERIC_VARIANTS(v, 4, "K4,any")
    VITB_GENERIC_VARIANTS(v, 5, "K5,any")
    VITB_GENERIC_VARIANTS(v, 6, "K6,any")
    VITB_GENERIC_VARIANTS(v, 7, "K7,any")
}
}
