#include "vitb_registry.h"
namespace vitb {
using TheCode = Code<7, 4, 109, 79, 83, 109>;   // DAB, common_codes.h:25
void register_k7r4_t1(std::vector<KernelEntry>& v) { VITB_VARIANTS(v, TheCode, 0, "K7,R4,dab,T1") }
}
