// Tensor-core (tcgen05) shortlist path of rr_dense_topk; see dense_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/rr_b200.h"

struct rr_tc_state;

// exact-path callback used for queries the tensor path cannot certify
typedef int (*rr_exact_fn)(void* ctx, const float* d_q, int32_t B, int32_t pool, int64_t* d_idx, float* d_sims,
                           int32_t* d_count, cudaStream_t stream);

bool rr_tc_supported(int cc_major, int cc_minor);
// true if this build's tensor path can serve (dim_pad, pool); AUTO mode falls back to the exact path otherwise
bool rr_tc_can_handle(int dim_pad, int pool);
int rr_tc_dense_topk(rr_tc_state** state, const rr_index_desc* d, int sm_count, const float* d_q, int32_t B,
                     int32_t pool, int64_t* d_idx, float* d_sims, int32_t* d_count, rr_dense_stats* stats,
                     rr_exact_fn exact, void* exact_ctx, int32_t* d_uncertified, cudaStream_t stream, int kp_override = 0,
                     int depth = 0);
// d_uncertified == NULL: synchronous -- the stream is synchronised once and uncertified queries are redone through
// `exact` before returning.  d_uncertified != NULL (int32[B]): nothing is read back, the mask says which queries'
// results are not proven exact.
void rr_tc_destroy(rr_tc_state* state);

// debug: raw tensor-core scores of a row range (see dense_tc.cu)
int rr_tc_debug_scores(rr_tc_state** state, const rr_index_desc* d, int sm_count, const float* d_q, int32_t B,
                       int64_t row0, int32_t n_rows, float* d_out, cudaStream_t s);
