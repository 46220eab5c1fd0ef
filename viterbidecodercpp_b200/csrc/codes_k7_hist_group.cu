#include "vitb_registry.h"
namespace vitb {
using Voyager = Code<7, 2, 109, 79>;        // common_codes.h:23
using LTE = Code<7, 3, 91, 117, 121>;       // common_codes.h:24
using DAB = Code<7, 4, 109, 79, 83, 109>;   // common_codes.h:25
void register_k7_hist_group(std::vector<KernelEntry>& v) {
    VITB_HIST_GROUP_VARIANTS(v, Voyager, "K7,R2,voyager")
    VITB_HIST_GROUP_VARIANTS(v, LTE, "K7,R3,lte")
    VITB_HIST_GROUP_VARIANTS(v, DAB, "K7,R4,dab")
}
}
