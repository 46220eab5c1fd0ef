#include "vitb_registry.h"
namespace vitb {
using TheCode = Code<9, 2, 491, 369>;   // CDMA IS-95A, common_codes.h:26
void register_k9r2_t16(std::vector<KernelEntry>& v) { VITB_VARIANTS(v, TheCode, 4, "K9,R2,is95a,T16") }
}
