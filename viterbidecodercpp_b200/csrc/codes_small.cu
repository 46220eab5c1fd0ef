#include "vitb_registry.h"
namespace vitb {
using K3R2 = Code<3, 2, 0b111, 0b101>;        // common_codes.h:21
using K5R2 = Code<5, 2, 0b10111, 0b11001>;    // common_codes.h:22
void register_small(std::vector<KernelEntry>& v) {
    VITB_VARIANTS(v, K3R2, 0, "K3,R2,basic,T1")
    VITB_VARIANTS(v, K5R2, 0, "K5,R2,basic,T1")
    VITB_VARIANTS(v, K5R2, 2, "K5,R2,basic,T4")
}
}
