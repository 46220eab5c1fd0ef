#include "vitb_registry.h"
namespace vitb {
using DAB = Code<7, 4, 109, 79, 83, 109>;   // common_codes.h:25
void register_k7r4(std::vector<KernelEntry>& v) { VITB_PAIR_VARIANTS(v, DAB, "K7,R4,dab") }
}
