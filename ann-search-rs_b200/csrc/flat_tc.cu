// flat_tc.cu -- tensor-core flat search path (stub until the tcgen05 kernel lands).
#include "flat_tc.hpp"
#include "index.hpp"

namespace annb {
int tc_flat_prepare(annb_index*) { return ANNB_OK; }
bool tc_flat_supported(const annb_index*, int, uint32_t) { return false; }
int tc_flat_search(annb_index*, const uint8_t*, uint32_t, int, int, uint64_t, uint32_t, uint32_t, uint64_t*, float*, uint32_t*, cudaStream_t) {
    set_last_error("tensor path not built");
    return ANNB_ERR_UNSUPPORTED;
}
void tc_destroy(annb_index*) {}
}  // namespace annb
