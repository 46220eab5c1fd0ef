#include "vitb_registry.h"
namespace vitb {
using LTE = Code<7, 3, 91, 117, 121>;   // common_codes.h:24
void register_k7r3(std::vector<KernelEntry>& v) { VITB_PAIR_VARIANTS(v, LTE, "K7,R3,lte") }
}
