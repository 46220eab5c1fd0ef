#include "vitb_registry.h"
namespace vitb {
using TheCode = Code<9, 4, 501, 441, 331, 315>;   // CDMA 2000, common_codes.h:27
void register_k9r4_t8(std::vector<KernelEntry>& v) { VITB_VARIANTS(v, TheCode, 3, "K9,R4,cdma2000,T8") }
}
