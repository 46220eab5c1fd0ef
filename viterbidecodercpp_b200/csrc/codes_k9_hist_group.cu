#include "vitb_registry.h"
namespace vitb {
using IS95A = Code<9, 2, 491, 369>;                // CDMA IS-95A, common_codes.h:26
using CDMA2000 = Code<9, 4, 501, 441, 331, 315>;   // CDMA 2000, common_codes.h:27
void register_k9_hist_group(std::vector<KernelEntry>& v) {
    VITB_HIST_GROUP_VARIANTS(v, IS95A, "K9,R2,is95a")
    VITB_HIST_GROUP_VARIANTS(v, CDMA2000, "K9,R4,cdma2000")
}
}
