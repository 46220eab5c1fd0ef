#include "vitb_registry.h"
namespace vitb {
using TheCode = Code<7, 3, 91, 117, 121>;   // LTE, common_codes.h:24
void register_k7r3_t1(std::vector<KernelEntry>& v) { VITB_VARIANTS(v, TheCode, 0, "K7,R3,lte,T1") }
}
