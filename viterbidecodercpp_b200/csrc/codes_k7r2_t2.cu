#include "vitb_registry.h"
namespace vitb {
using TheCode = Code<7, 2, 109, 79>;   // Voyager, common_codes.h:23
void register_k7r2_t2(std::vector<KernelEntry>& v) { VITB_VARIANTS(v, TheCode, 1, "K7,R2,voyager,T2") }
}
