// placeholder until the tcgen05 kernel lands
#include "rr_internal.h"
#include "dense_tc.h"
struct rr_tc_state { int unused; };
bool rr_tc_supported(int, int) { return false; }
int rr_tc_dense_topk(rr_tc_state**, const rr_index_desc*, int, const float*, int32_t, int32_t, int64_t*, float*,
                     int32_t*, rr_dense_stats*, rr_exact_fn, void*, cudaStream_t) {
    return rr_fail(RR_EUNSUPPORTED, "tensor path not built");
}
void rr_tc_destroy(rr_tc_state* s) { delete s; }
