#include "vitb_registry.h"
namespace vitb {
using Voyager = Code<7, 2, 109, 79>;   // common_codes.h:23
void register_k7r2(std::vector<KernelEntry>& v) { VITB_PAIR_VARIANTS(v, Voyager, "K7,R2,voyager") }
}
