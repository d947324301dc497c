// Exact fp32 dense scoring and selection (sm_100a).
//
// Stands behind cosine_similarity_search utils.py:111-124 (= cosine_search app/test.py:125-132,
// _cosine_pool app/app_product_search.py:192-195):  sims = mat @ q ; top-k ; sort descending.
//
// CANONICAL DOT PRODUCT.  Every fp32 similarity this library returns is computed by
// `canonical_dot`: lane j of a warp accumulates, with fmaf in ascending element order, the
// float4 chunks c = j, j+32, j+64, ... of the row; the 32 partial sums are combined by the
// xor-butterfly 16,8,4,2,1.  The exact path (this file), the shortlist rescoring of the tensor
// path (K3) and every batch size therefore return bit-identical similarities, and differ from
// BLAS sgemv only by fp32 summation order (<= ~1e-7 abs on unit vectors; tolerance 1e-6).
//
//  dense_scores_f32_kernel   HBM-bound multi-query GEMV: each warp streams corpus rows with
//      16-byte loads (4 rows in flight), the 1 / 2 / 4 / 8 queries of a CTA sit in shared memory
//      (one instantiation per batch width).  Algorithmic bytes: 4*D per row per group of 8 queries.
//  row top-k on 64-bit composite keys (score desc, row asc):
//      topk_small_kernel / topk_chunk_kernel  rows of <= 16384 scores in ONE kernel, longer rows of small
//      batches in a 2-3 level tree: keys in shared memory, MSB-first 8-bit radix select, rank sort;
//      radix-select (rs_*)  everything else: 11/11/10-bit digit histograms over global memory, early
//      exit once the pivot bucket is exactly consumed, then collect + bitonic sort of the k survivors.
//  rescore_kernel (K3)       exact similarities of a shortlist (one warp per (query, row)).
#include <cstdlib>

#include "rr_internal.h"
#include "rr_kernels.h"

namespace {

constexpr int QB = 8;              // queries per CTA in the GEMV kernel
constexpr int GEMV_THREADS = 256;  // 8 warps
constexpr int GEMV_ROWS = 4;       // rows per warp iteration

__device__ __forceinline__ float warp_xor_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// warp_xor_sum of NQ values per lane (NQ = 1, 2, 4, 8) with 9 shuffles instead of 40 for NQ = 8: at the offsets 16, 8, 4
// a lane keeps one half of its values and trades the other half with its xor partner, so every value still receives
// exactly the additions of the butterfly (own + partner's, offsets 16, 8, 4, 2, 1 in that order) and the sums are
// bit-identical to warp_xor_sum.  On return v[0] is the sum of value `b` (the return value) on the lanes that own it;
// -1 on the other lanes.
template <int NQ>
__device__ __forceinline__ int warp_xor_sum_multi(float (&v)[NQ], int lane) {
    int b = 0, width = NQ;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if (width > 1) {
            const int h = width >> 1;
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NQ / 2; ++i) {
                if (i < h) {
                    const float send = up ? v[i] : v[i + h];
                    const float keep = up ? v[i + h] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            b = b * 2 + (up ? 1 : 0);
            width = h;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    // NQ = 8: offsets 16, 8, 4 chose the value, every lane of a group of 4 holds the same sum; one of them reports it
    constexpr int group = 32 / NQ;
    return (lane & (group - 1)) == 0 ? b : -1;
}

__device__ __forceinline__ float4 ldg_row16(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// lane-partial of the canonical dot for a row whose length is a multiple of 4 (16-byte aligned)
__device__ __forceinline__ float lane_partial_v4(const float4* __restrict__ row, const float4* __restrict__ q,
                                                 int n_chunks, int lane) {
    float acc = 0.f;
    for (int c = lane; c < n_chunks; c += 32) {
        const float4 m = __ldg(row + c);
        const float4 x = q[c];
        acc = fmaf(x.x, m.x, acc);
        acc = fmaf(x.y, m.y, acc);
        acc = fmaf(x.z, m.z, acc);
        acc = fmaf(x.w, m.w, acc);
    }
    return acc;
}
// generic (any D, any alignment): same chunking, scalar loads
__device__ __forceinline__ float lane_partial_scalar(const float* __restrict__ row, const float* __restrict__ q,
                                                     int D, int lane) {
    float acc = 0.f;
    for (int c = lane; c * 4 < D; c += 32) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = c * 4 + e;
            if (i < D) acc = fmaf(q[i], __ldg(row + i), acc);
        }
    }
    return acc;
}

// scores[b, row] for b in [0, nq), all rows.  Queries of the CTA's group live in smem.  NQ = queries per CTA (8 for
// batches; 1 / 2 / 4 for the single-query calls of the Streamlit app, where 8 accumulators per row would spend 8x the
// FMAs, shared-memory reads and shuffles and the registers that limit the loads in flight), ROWS = corpus rows a warp
// has in flight.  The per-(row, query) arithmetic is canonical_dot for every instantiation.
template <bool VEC4, int NQ, int ROWS>
__global__ void __launch_bounds__(GEMV_THREADS)
dense_scores_f32_kernel(const float* __restrict__ emb, long long n_rows, int D,
                        const float* __restrict__ queries, int n_queries,
                        float* __restrict__ scores, long long ld_scores) {
    extern __shared__ __align__(16) float s_q[];   // [NQ][Dp]
    const int Dp = (D + 3) & ~3;
    const int q0 = blockIdx.y * NQ;
    const int nq = min(NQ, n_queries - q0);
    for (int i = threadIdx.x; i < NQ * Dp; i += GEMV_THREADS) {
        const int b = i / Dp, d = i - b * Dp;
        s_q[i] = (b < nq && d < D) ? queries[(long long)(q0 + b) * D + d] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long warps_total = (long long)gridDim.x * (GEMV_THREADS / 32);
    const long long gw = (long long)blockIdx.x * (GEMV_THREADS / 32) + warp;

    if constexpr (VEC4) {
        const int n_chunks = D >> 2;
        for (long long r0 = gw * ROWS; r0 < n_rows; r0 += warps_total * ROWS) {
            float acc[ROWS][NQ];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                for (int b = 0; b < NQ; ++b) acc[rr][b] = 0.f;
#pragma unroll 3
            for (int c = lane; c < n_chunks; c += 32) {
                float4 m[ROWS];
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) {
                    const long long r = min(r0 + rr, n_rows - 1);
                    m[rr] = ldg_row16(reinterpret_cast<const float4*>(emb + r * D) + c);
                }
#pragma unroll
                for (int b = 0; b < NQ; ++b) {
                    const float4 x = reinterpret_cast<const float4*>(s_q + b * Dp)[c];
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr) {
                        float a = acc[rr][b];
                        a = fmaf(x.x, m[rr].x, a);
                        a = fmaf(x.y, m[rr].y, a);
                        a = fmaf(x.z, m[rr].z, a);
                        a = fmaf(x.w, m[rr].w, a);
                        acc[rr][b] = a;
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                const long long r = r0 + rr;
                const int b = warp_xor_sum_multi<NQ>(acc[rr], lane);
                if (b >= 0 && b < nq && r < n_rows) scores[(long long)(q0 + b) * ld_scores + r] = acc[rr][0];
            }
        }
    } else {
        for (long long r = gw; r < n_rows; r += warps_total) {
            for (int b = 0; b < nq; ++b) {
                const float s = warp_xor_sum(lane_partial_scalar(emb + r * D, s_q + b * Dp, D, lane));
                if (lane == 0) scores[(long long)(q0 + b) * ld_scores + r] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: exact similarity of shortlisted rows.  One warp per (query, slot).
// rows int64[B, n_slots] (<0 = empty -> -inf).  q float[B, D].
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ emb, long long n_rows, int D, const float* __restrict__ queries,
               const long long* __restrict__ rows, int n_slots, int B, float* __restrict__ out) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * n_slots) return;
    const int q = (int)(gw / n_slots);
    const long long r = rows[gw];
    float s;
    if (r < 0 || r >= n_rows) {
        s = -INFINITY;
    } else if constexpr (VEC4) {
        s = warp_xor_sum(lane_partial_v4(reinterpret_cast<const float4*>(emb + r * D),
                                         reinterpret_cast<const float4*>(queries + (long long)q * D), D >> 2, lane));
    } else {
        s = warp_xor_sum(lane_partial_scalar(emb + r * D, queries + (long long)q * D, D, lane));
    }
    if (lane == 0) out[gw] = s;
}

// Same similarities, RS shortlist slots of one query per warp: the RS rows' 16-byte loads of a chunk are issued
// together (RS x 1.5 KB in flight per warp instead of 1.5 KB; the r02 profile had the one-row kernel at 57 % of
// DRAM peak with 51 % occupancy) and the query chunk is read once.  Per row the lane-partial order is unchanged.
template <int RS>
__global__ void __launch_bounds__(256)
rescore_multi_kernel(const float* __restrict__ emb, long long n_rows, int D, const float* __restrict__ queries,
                     const long long* __restrict__ rows, int n_slots, int B, float* __restrict__ out) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int groups = (n_slots + RS - 1) / RS;
    if (gw >= (long long)B * groups) return;
    const int q = (int)(gw / groups);
    const int s0 = (int)(gw - (long long)q * groups) * RS;
    const float4* qv = reinterpret_cast<const float4*>(queries + (long long)q * D);
    const float4* rp[RS];
    bool ok[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const long long row = s0 + r < n_slots ? rows[(long long)q * n_slots + s0 + r] : -1;
        ok[r] = row >= 0 && row < n_rows;
        rp[r] = reinterpret_cast<const float4*>(emb + (ok[r] ? row : 0) * D);
    }
    float acc[RS];
    bool any = false;
#pragma unroll
    for (int r = 0; r < RS; ++r) { acc[r] = 0.f; any |= ok[r]; }
    const int n_chunks = any ? D >> 2 : 0;            // a group of empty / pruned slots has nothing to load
    for (int c = lane; c < n_chunks; c += 32) {
        float4 m[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) m[r] = ldg_row16(rp[r] + c);
        const float4 x = __ldg(qv + c);
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            float a = acc[r];
            a = fmaf(x.x, m[r].x, a);
            a = fmaf(x.y, m[r].y, a);
            a = fmaf(x.z, m[r].z, a);
            a = fmaf(x.w, m[r].w, a);
            acc[r] = a;
        }
    }
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const float s = warp_xor_sum(acc[r]);
        if (lane == 0 && s0 + r < n_slots) out[(long long)q * n_slots + s0 + r] = ok[r] ? s : -INFINITY;
    }
}

// ---------------------------------------------------------------------------------------------
// Best review per (query, candidate product): segmented arg-max of canonical dot products.
// One warp per (query, candidate).  rev_range[r] = {lo, hi}: the review slots of product row r
// (slots of one product are contiguous and in file order; rows that share a SKU share a range).
// slot_file (optional) = file position of every slot and limit[q] = first file position that the
// reference's `max_rows` cap drops for query q (app/app_product_search.py:343-346).
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256)
best_review_kernel(const float* __restrict__ rev, const long long* __restrict__ rev_range, long long n_products, int D,
                   const float* __restrict__ queries, const long long* __restrict__ cand, int pool, int B,
                   const long long* __restrict__ slot_file, const long long* __restrict__ limit,
                   float* __restrict__ out_score, long long* __restrict__ out_slot) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * pool) return;
    const int q = (int)(gw / pool);
    const long long r = cand[gw];
    float best = 0.f;
    long long best_i = -1;
    if (r >= 0 && r < n_products) {
        const long long lo = rev_range[2 * r], hi = rev_range[2 * r + 1];
        const long long lim = (slot_file != nullptr && limit != nullptr) ? limit[q] : 0x7fffffffffffffffLL;
        for (long long i = lo; i < hi; ++i) {
            if (slot_file != nullptr && slot_file[i] >= lim) break;       // file order inside a product
            float s;
            if constexpr (VEC4)
                s = warp_xor_sum(lane_partial_v4(reinterpret_cast<const float4*>(rev + i * D),
                                                 reinterpret_cast<const float4*>(queries + (long long)q * D), D >> 2, lane));
            else
                s = warp_xor_sum(lane_partial_scalar(rev + i * D, queries + (long long)q * D, D, lane));
            // first maximum, NaN counts as the maximum: np.argmax (:356)
            if (best_i < 0 || s > best || (s != s && best == best)) { best = s; best_i = i; }
        }
    }
    if (lane == 0) { out_score[gw] = best; out_slot[gw] = best_i; }
}

// ---------------------------------------------------------------------------------------------
// Radix select of the top-k composite keys of every row of a score matrix.
// ---------------------------------------------------------------------------------------------
constexpr int RS_BINS = 2048;
constexpr int RS_THREADS = 512;
constexpr int RS_LEVELS = 6;
__constant__ int c_rs_bits[RS_LEVELS] = {11, 11, 10, 11, 11, 10};

struct RsRow {
    unsigned long long prefix;   // digits chosen so far (top `bits_done` bits of the pivot)
    int bits_done;
    int k_rem;                   // how many of the pivot bucket's keys are still wanted
    int done;                    // 1: every key with (key >> (64-bits_done)) >= prefix is selected
    int out_count;
};

__global__ void rs_init_kernel(RsRow* st, unsigned* hist, int rows, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) { st[i].prefix = 0; st[i].bits_done = 0; st[i].k_rem = k; st[i].done = 0; st[i].out_count = 0; }
    for (long long j = i; j < (long long)rows * RS_BINS; j += (long long)gridDim.x * blockDim.x) hist[j] = 0;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const float* __restrict__ scores, long long ld, long long n, const RsRow* __restrict__ st,
               unsigned* __restrict__ hist, int level) {
    const int row = blockIdx.y;
    const RsRow s = st[row];
    if (s.done) return;
    __shared__ unsigned sh[RS_BINS];
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const int bits = c_rs_bits[level];
    const int shift = 64 - s.bits_done - bits;
    const float* src = scores + (long long)row * ld;
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(n, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += RS_THREADS) {
        const unsigned long long key = rr_make_key(src[i], (uint32_t)i);
        const bool in = s.bits_done == 0 ? true : ((key >> (64 - s.bits_done)) == s.prefix);
        if (in) atomicAdd(&sh[(unsigned)((key >> shift) & ((1u << bits) - 1u))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS)
        if (sh[i]) atomicAdd(&hist[(long long)row * RS_BINS + i], sh[i]);
}

__global__ void __launch_bounds__(RS_BINS / 2)
rs_scan_kernel(RsRow* st, unsigned* hist, int level) {
    const int row = blockIdx.x;
    RsRow s = st[row];
    unsigned* h = hist + (long long)row * RS_BINS;
    if (s.done) return;
    __shared__ unsigned sh[RS_BINS];
    __shared__ int s_bucket;
    const int bits = c_rs_bits[level];
    const int nb = 1 << bits;
    for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) { sh[i] = i < nb ? h[i] : 0u; h[i] = 0u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        // serial walk from the top bucket: nb <= 2048, once per row per level
        unsigned above = 0;
        int b = nb - 1;
        for (; b > 0; --b) {
            if (above + sh[b] >= (unsigned)s.k_rem) break;
            above += sh[b];
        }
        s.k_rem -= (int)above;
        s.prefix = (s.prefix << bits) | (unsigned long long)b;
        s.bits_done += bits;
        if ((int)sh[b] == s.k_rem || s.bits_done == 64) s.done = 1;
        st[row] = s;
        s_bucket = b;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_collect_kernel(const float* __restrict__ scores, long long ld, long long n, RsRow* st,
                  unsigned long long* __restrict__ out_keys, int k_cap) {
    const int row = blockIdx.y;
    const unsigned long long prefix = st[row].prefix;
    const int bits_done = st[row].bits_done;
    const float* src = scores + (long long)row * ld;
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(n, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += RS_THREADS) {
        const unsigned long long key = rr_make_key(src[i], (uint32_t)i);
        const bool sel = bits_done == 0 ? true : ((key >> (64 - bits_done)) >= prefix);
        if (sel) {
            const int slot = atomicAdd(&st[row].out_count, 1);
            if (slot < k_cap) out_keys[(long long)row * k_cap + slot] = key;
        }
    }
}

// Bitonic sort (descending) of up to n_pad (power of two) 64-bit keys in shared memory, one CTA
// per row; writes idx/score of the first k.  Missing entries are key 0 -> idx -1 / -inf.
__global__ void __launch_bounds__(1024)
sort_keys_kernel(const unsigned long long* __restrict__ keys, int k_cap, const RsRow* __restrict__ st,
                 int k, int n_pad, long long* __restrict__ out_idx, float* __restrict__ out_score,
                 int32_t* __restrict__ out_count, int out_ld) {
    extern __shared__ unsigned long long sk[];
    const int row = blockIdx.x;
    const int cnt = min(st[row].out_count, k_cap);
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) sk[i] = i < cnt ? keys[(long long)row * k_cap + i] : 0ull;
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n_pad / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a < b) == desc) { sk[lo] = b; sk[hi] = a; }
            }
            __syncthreads();
        }
    }
    const int kk = min(k, cnt);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        if (i < kk) {
            out_idx[(long long)row * out_ld + i] = (long long)rr_key_index(sk[i]);
            out_score[(long long)row * out_ld + i] = rr_key_score(sk[i]);
        } else {
            out_idx[(long long)row * out_ld + i] = -1;
            out_score[(long long)row * out_ld + i] = -INFINITY;
        }
    }
    if (threadIdx.x == 0 && out_count) out_count[row] = kk;
}

// ---------------------------------------------------------------------------------------------
// Short score rows and small batches: top-k inside shared memory instead of the 15-launch radix pipeline above.
//   n <= 16384 (the single-query shape of the Streamlit app): ONE kernel, one CTA per row (topk_small_kernel);
//   longer rows, at most TC_MAX_ROWS (32) queries, k <= 1024: a tree of the same step (topk_chunk_kernel) -- every CTA takes a
//   chunk of <= 4096 scores and keeps its k largest composite keys, the next level does the same over the surviving
//   keys, the last level (one CTA per row, <= 8192 keys) sorts: 2 launches up to ~220 k rows at k = 150, 3 up to ~12 M.
// The step: all keys of the chunk go to shared memory, an MSB-first radix select (8-bit digits, shared-memory histogram,
// early exit) finds the k-th key, the k survivors are compacted; the final level bitonic-sorts them and writes indices /
// scores in (score desc, row asc) order.  The top-k of the union of per-chunk top-k lists is the top-k of the row, and the
// composite keys are unique, so the result is the radix pipeline's bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int TS_MAX_N = 16384;
constexpr int TS_THREADS = 1024;         // 32 warps: the select is a chain of short dependent phases, more warps hide their latency
constexpr int TC_MAX_ROWS = 32;             // default of RR_TOPK_TREE_MAX_ROWS
constexpr int TC_MAX_K = 1024;

// The kk largest of the n keys in shared memory -> sel[0, kk) (unordered), sel[kk, k_pad) = 0.  Keys are unique except for
// 0 ("empty"), and kk <= the number of non-zero keys.  Called by all TS_THREADS threads; the keys may have been written
// just before the call (the first barrier orders them); sel is complete for every thread on return.
// EXACT = false (the caller sorts sel and takes the first kk): the digit passes stop as soon as the keys at or above the
// pivot bucket fit sel -- kk <= survivors <= k_pad, typically after 2 of the 4 score-digit passes.
// The first pass sees the sign and the high exponent bits, which nearly all keys share: its histogram is built with one
// shared-memory atomic per distinct digit of a warp (match.any) instead of 32 serialised ones on one address.
template <bool EXACT>
__device__ __forceinline__ void cta_select_keys(const unsigned long long* keys, int n, int kk, unsigned long long* sel,
                                                int k_pad) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_bits, s_krem, s_done, s_out;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { s_bits = 0; s_prefix = 0ull; s_krem = kk; s_done = kk >= n ? 1 : 0; s_out = 0; }
    for (int i = tid; i < k_pad; i += TS_THREADS) sel[i] = 0ull;
    __syncthreads();
    if (kk <= 0) return;
    for (int pass = 0; pass < 8; ++pass) {
        if (s_done) break;
        const int bits = s_bits;
        const unsigned long long prefix = s_prefix;
        if (tid < 256) s_hist[tid] = 0u;
        __syncthreads();
        const int shift = 64 - bits - 8;
        if (bits == 0) {
            for (int i0 = 0; i0 < n; i0 += TS_THREADS) {
                const int i = i0 + tid;
                const unsigned live = __ballot_sync(0xffffffffu, i < n);
                if (i < n) {
                    const unsigned digit = (unsigned)(keys[i] >> 56);
                    const unsigned peers = __match_any_sync(live, digit);
                    if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
                }
            }
        } else {
            for (int i = tid; i < n; i += TS_THREADS) {
                const unsigned long long key = keys[i];
                if ((key >> (64 - bits)) == prefix) atomicAdd(&s_hist[(unsigned)((key >> shift) & 0xFFu)], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {
            // walk the buckets from the top: lane l owns buckets 255-8l .. 248-8l
            unsigned loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = s_hist[255 - (lane * 8 + j)]; sum += loc[j]; }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const unsigned excl = incl - sum;
            const unsigned krem = (unsigned)s_krem;
            if (excl < krem && krem <= incl) {
                unsigned above = excl;
                int j = 0;
                for (; j < 7; ++j) { if (above + loc[j] >= krem) break; above += loc[j]; }
                s_krem = (int)(krem - above);
                s_prefix = (prefix << 8) | (unsigned long long)(255 - (lane * 8 + j));
                s_bits = bits + 8;
                const unsigned rem = krem - above;              // still wanted from the pivot bucket, which holds loc[j]
                if (loc[j] == rem || bits + 8 >= 64 || (!EXACT && (unsigned)kk - rem + loc[j] <= (unsigned)k_pad)) s_done = 1;
            }
        }
        __syncthreads();
    }
    // survivors: every key whose top `bits` bits are >= the pivot prefix (exactly kk of them; EXACT = false: kk .. k_pad)
    const int bits = s_bits;
    const unsigned long long prefix = s_prefix;
    for (int i0 = 0; i0 < n; i0 += TS_THREADS) {
        const int i = i0 + tid;
        bool take = false;
        unsigned long long key = 0ull;
        if (i < n) { key = keys[i]; take = bits == 0 || (key >> (64 - bits)) >= prefix; }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_out, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < k_pad) sel[slot] = key;
        }
    }
    __syncthreads();
}

// sel[0, k_pad) sorted descending (k_pad a power of two), the first kk written as (index, score); the rest of the k output
// slots are -1 / -inf
__device__ __forceinline__ void cta_sort_and_write(unsigned long long* sel, int k_pad, int k, int kk, int row,
                                                   long long* __restrict__ out_idx, float* __restrict__ out_score,
                                                   int32_t* __restrict__ out_count, int out_ld) {
    const int tid = threadIdx.x;
    if (k_pad <= 256) {
        // rank sort: a thread counts the keys above its own (the empty slots, all 0, by position) -- two barriers instead
        // of the 36 steps of the bitonic network
        unsigned long long mine = 0ull;
        int rank = 0;
        if (tid < k_pad) {
            mine = sel[tid];
#pragma unroll 8
            for (int j = 0; j < k_pad; ++j) {
                const unsigned long long o = sel[j];
                rank += (o > mine || (o == mine && j < tid)) ? 1 : 0;
            }
        }
        __syncthreads();
        if (tid < k_pad) sel[rank] = mine;
        __syncthreads();
    } else {
        for (int size = 2; size <= k_pad; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < k_pad / 2; i += TS_THREADS) {
                    const int lo = 2 * i - (i & (stride - 1));
                    const int hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const unsigned long long a = sel[lo], b = sel[hi];
                    if ((a < b) == desc) { sel[lo] = b; sel[hi] = a; }
                }
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < k; i += TS_THREADS) {
        if (i < kk) {
            out_idx[(long long)row * out_ld + i] = (long long)rr_key_index(sel[i]);
            out_score[(long long)row * out_ld + i] = rr_key_score(sel[i]);
        } else {
            out_idx[(long long)row * out_ld + i] = -1;
            out_score[(long long)row * out_ld + i] = -INFINITY;
        }
    }
    if (tid == 0 && out_count) out_count[row] = kk;
}

__global__ void __launch_bounds__(TS_THREADS)
topk_small_kernel(const float* __restrict__ scores, long long ld, int n, int k, long long* __restrict__ out_idx,
                  float* __restrict__ out_score, int32_t* __restrict__ out_count, int out_ld, int k_pad) {
    extern __shared__ unsigned long long ts_keys[];            // [n] keys, then [k_pad] survivors
    unsigned long long* sel = ts_keys + n;
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* src = scores + (long long)row * ld;
    for (int i = tid; i < n; i += TS_THREADS) {
        unsigned long long key = rr_make_key(src[i], (uint32_t)i);
        ts_keys[i] = key ? key : 1ull;
    }
    const int kk = min(k, n);
    cta_select_keys<false>(ts_keys, n, kk, sel, k_pad);
    cta_sort_and_write(sel, k_pad, k, kk, row, out_idx, out_score, out_count, out_ld);
}

// One level of the tree.  grid = (chunks of the row, rows).  FROM_KEYS: the source is the previous level's key array
// (0 = empty slot), else the score row itself.  FINAL (one chunk per row): sort and write the result; otherwise the chunk's
// survivors go to out_keys[row][chunk][k] (unordered, zero-padded).
template <bool FROM_KEYS, bool FINAL>
__global__ void __launch_bounds__(TS_THREADS)
topk_chunk_kernel(const float* __restrict__ scores, const unsigned long long* __restrict__ in_keys, long long ld, long long n,
                  int chunk, int k, unsigned long long* __restrict__ out_keys, long long out_ld,
                  long long* __restrict__ out_idx, float* __restrict__ out_score, int32_t* __restrict__ out_count,
                  int out_ld2, int k_pad) {
    extern __shared__ unsigned long long ts_keys[];            // [chunk] keys, then [k_pad] survivors
    unsigned long long* sel = ts_keys + chunk;
    __shared__ int s_valid;
    const int row = blockIdx.y, tid = threadIdx.x;
    const long long lo = (long long)blockIdx.x * chunk;
    const int n_loc = (int)min((long long)chunk, n - lo);
    int valid = n_loc;
    if (FROM_KEYS) {
        if (tid == 0) s_valid = 0;
        __syncthreads();
        const unsigned long long* src = in_keys + (long long)row * ld + lo;
        int nz = 0;
        for (int i = tid; i < n_loc; i += TS_THREADS) {
            const unsigned long long key = src[i];
            ts_keys[i] = key;
            nz += key != 0ull;
        }
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        if ((tid & 31) == 0 && nz) atomicAdd(&s_valid, nz);
        __syncthreads();
        valid = s_valid;
    } else {
        const float* src = scores + (long long)row * ld + lo;
        for (int i = tid; i < n_loc; i += TS_THREADS) {
            const unsigned long long key = rr_make_key(src[i], (uint32_t)(lo + i));
            ts_keys[i] = key ? key : 1ull;
        }
    }
    const int kk = min(k, valid);
    cta_select_keys<!FINAL>(ts_keys, n_loc, kk, sel, k_pad);
    if (FINAL) {
        cta_sort_and_write(sel, k_pad, k, kk, row, out_idx, out_score, out_count, out_ld2);
    } else {
        unsigned long long* dst = out_keys + (long long)row * out_ld + (long long)blockIdx.x * k;
        for (int i = tid; i < k; i += TS_THREADS) dst[i] = sel[i];
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

// Levels of the shared-memory top-k tree for rows of n scores (see topk_chunk_kernel): chunk[l] source elements per CTA,
// groups[l] CTAs per row; the last level has one group.  False when the tree does not apply.
struct ChunkPlan { int levels; int chunk[12]; int groups[12]; };

static int topk_env(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    if (!e) return dflt;
    const int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

static int topk_tree_max_rows() { return topk_env("RR_TOPK_TREE_MAX_ROWS", TC_MAX_ROWS, 1, 64); }

static bool plan_chunked_topk(int64_t n, int rows, int k, int sm_count, ChunkPlan* p) {
    // rows up to tree_min_n go through the one-kernel top-k; a level-0 CTA takes at most chunk_max scores, a middle level
    // mid_chunk keys, the last level (one CTA per row) at most final_max keys.  Defaults from the r02 sweep at B = 1
    // (profiles/r02_exact_path_gemv_topk.txt): 4096 / 8192 / 8192 -> 27 us at 100 k rows, 38 us at 1 M (16384 each: 31 / 46)
    const int tree_min_n = topk_env("RR_TOPK_TREE_MIN_N", TS_MAX_N, 1024, TS_MAX_N);
    if (n <= tree_min_n || rows > topk_tree_max_rows() || k > TC_MAX_K || k < 1 || k >= n || n > (int64_t)1 << 31) return false;
    if (getenv("RR_NO_CHUNKED_TOPK")) return false;
    const int two_k = (2 * k + 1023) / 1024 * 1024;
    const int chunk_max = std::max(two_k, topk_env("RR_TOPK_CHUNK_MAX", 4096, 1024, TS_MAX_N));
    const int mid_chunk = std::max(two_k, topk_env("RR_TOPK_MID_CHUNK", 8192, 1024, TS_MAX_N));
    const int final_max = std::max(two_k, topk_env("RR_TOPK_FINAL_MAX", 8192, 1024, TS_MAX_N));
    // level 0: enough chunks to fill the SMs, few enough that their survivors fit ONE final CTA when the row allows it
    const int64_t fit = std::max<int64_t>(1, final_max / k);                // chunks whose k survivors fit the last level
    int64_t c = std::max((n + std::max(sm_count, 1) - 1) / std::max(sm_count, 1), (n + fit - 1) / fit);
    c = std::min<int64_t>(chunk_max, std::max<int64_t>(1024, (c + 1023) / 1024 * 1024));
    int l = 0;
    int64_t cur = n;
    p->chunk[l] = (int)c; p->groups[l] = (int)((cur + c - 1) / c); cur = (int64_t)p->groups[l] * k; ++l;
    while (cur > final_max) {                       // every middle level at least halves the keys (mid_chunk >= 2 k)
        if (l >= 10) return false;
        p->chunk[l] = mid_chunk; p->groups[l] = (int)((cur + mid_chunk - 1) / mid_chunk); cur = (int64_t)p->groups[l] * k; ++l;
    }
    p->chunk[l] = (int)cur; p->groups[l] = 1; ++l;
    p->levels = l;
    return true;
}

size_t rr_exact_scratch_bytes(int rows, int k, int64_t n, int sm_count) {
    size_t bytes = sizeof(RsRow) * (size_t)rows + sizeof(unsigned) * (size_t)rows * RS_BINS +
                   sizeof(unsigned long long) * (size_t)rows * (size_t)k + 256;
    ChunkPlan p;
    const int tree_rows = std::min(rows, topk_tree_max_rows());   // the last slice of a larger batch may take the tree
    if (plan_chunked_topk(n, tree_rows, k, sm_count, &p)) {
        // two key arrays used alternately: level 0's survivors and level 1's
        const size_t a = (size_t)p.groups[0] * k, b = p.levels > 2 ? (size_t)p.groups[1] * k : 0;
        bytes = std::max(bytes, sizeof(unsigned long long) * (size_t)tree_rows * (a + b) + 256);
    }
    return bytes;
}

template <bool VEC4, int NQ, int ROWS>
static int launch_gemv(const float* d_emb, int64_t n_rows, int D, const float* d_q, int n_queries, float* d_scores,
                       int64_t ld_scores, int sm_count, cudaStream_t stream) {
    const int Dp = (D + 3) & ~3;
    const size_t smem = sizeof(float) * (size_t)NQ * Dp;
    const int warps_per_cta = GEMV_THREADS / 32;
    const long long want = (n_rows + (long long)warps_per_cta * ROWS - 1) / ((long long)warps_per_cta * ROWS);
    const int groups = (n_queries + NQ - 1) / NQ;
    const long long cap = (long long)sm_count * 8;
    dim3 grid((unsigned)max(1ll, min(want, cap)), (unsigned)groups);
    if (smem > 48 * 1024)
        RR_CUDA(cudaFuncSetAttribute(dense_scores_f32_kernel<VEC4, NQ, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RrProfScope prof(RR_PROF_DENSE_GEMV, stream);
    dense_scores_f32_kernel<VEC4, NQ, ROWS><<<grid, GEMV_THREADS, smem, stream>>>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores);
    RR_LAUNCH_CHECK();
    return RR_OK;
}

int rr_launch_dense_scores_f32(const float* d_emb, int64_t n_rows, int D, const float* d_q, int n_queries,
                               float* d_scores, int64_t ld_scores, int sm_count, cudaStream_t stream) {
    if (n_queries <= 0 || n_rows <= 0) return RR_OK;
    const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_emb) & 15) == 0);
    if (!vec4) return launch_gemv<false, QB, 1>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores, sm_count, stream);
    // RR_GEMV_WIDE=1: the 8-query kernel for every batch size (A/B switch of the narrow instantiations)
    if (n_queries > 4 || getenv("RR_GEMV_WIDE"))
        return launch_gemv<true, QB, GEMV_ROWS>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores, sm_count, stream);
    if (n_queries > 2) return launch_gemv<true, 4, 4>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores, sm_count, stream);
    if (n_queries > 1) return launch_gemv<true, 2, 4>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores, sm_count, stream);
    return launch_gemv<true, 1, 4>(d_emb, n_rows, D, d_q, n_queries, d_scores, ld_scores, sm_count, stream);
}

int rr_launch_rescore(const float* d_emb, int64_t n_rows, int D, const float* d_q, const int64_t* d_rows,
                      int n_slots, int B, float* d_out, cudaStream_t stream) {
    const long long warps = (long long)B * n_slots;
    if (warps <= 0) return RR_OK;
    const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_emb) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(d_q) & 15) == 0);
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    RrProfScope prof(RR_PROF_RESCORE, stream);
    if (vec4) {
        constexpr int RS = 4;
        const long long mwarps = (long long)B * ((n_slots + RS - 1) / RS);
        rescore_multi_kernel<RS><<<(unsigned)((mwarps * 32 + 255) / 256), 256, 0, stream>>>(
            d_emb, n_rows, D, d_q, reinterpret_cast<const long long*>(d_rows), n_slots, B, d_out);
    } else
        rescore_kernel<false><<<blocks, 256, 0, stream>>>(d_emb, n_rows, D, d_q, reinterpret_cast<const long long*>(d_rows), n_slots, B, d_out);
    RR_LAUNCH_CHECK();
    return RR_OK;
}

int rr_launch_best_review(const float* d_rev_emb, const int64_t* d_rev_range, int64_t n_products, int D,
                          const float* d_q, int B, const int64_t* d_cand, int pool, const int64_t* d_slot_file,
                          const int64_t* d_limit, float* d_score, int64_t* d_slot, cudaStream_t stream) {
    const long long warps = (long long)B * pool;
    if (warps <= 0) return RR_OK;
    const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_rev_emb) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(d_q) & 15) == 0);
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    RrProfScope prof(RR_PROF_MISC, stream);
    auto ll = [](const int64_t* p) { return reinterpret_cast<const long long*>(p); };
    if (vec4)
        best_review_kernel<true><<<blocks, 256, 0, stream>>>(d_rev_emb, ll(d_rev_range), n_products, D, d_q, ll(d_cand),
                                                             pool, B, ll(d_slot_file), ll(d_limit), d_score,
                                                             reinterpret_cast<long long*>(d_slot));
    else
        best_review_kernel<false><<<blocks, 256, 0, stream>>>(d_rev_emb, ll(d_rev_range), n_products, D, d_q, ll(d_cand),
                                                              pool, B, ll(d_slot_file), ll(d_limit), d_score,
                                                              reinterpret_cast<long long*>(d_slot));
    RR_LAUNCH_CHECK();
    return RR_OK;
}

// top-k (k <= 8192) of each of `rows` score rows of length n; results ordered (score desc, idx asc)
int rr_launch_topk_rows(const float* d_scores, int64_t ld, int64_t n, int rows, int k, void* d_scratch,
                        int64_t* d_idx, float* d_score, int32_t* d_count, int out_ld, int sm_count,
                        cudaStream_t stream) {
    if (rows <= 0) return RR_OK;
    if (k > 8192) return rr_fail(RR_EINVAL, "top-k larger than 8192 is not supported");
    const int kk = (int)(k < n ? (int64_t)k : n);
    ChunkPlan plan;
    const bool tree = plan_chunked_topk(n, rows, k, sm_count, &plan);
    if (!tree && n <= TS_MAX_N && n > 0 && !getenv("RR_NO_SMALL_TOPK")) {
        // single kernel per call for short rows (configs[0]: 10 k products)
        const int k_pad = max(2, next_pow2(max(kk, 1)));
        const size_t smem = sizeof(unsigned long long) * ((size_t)n + k_pad);
        static RrSmemOptIn optin;
        int dev = 0;
        if (optin.needed(smem, &dev)) {
            if (smem > 48 * 1024)
                RR_CUDA(cudaFuncSetAttribute(topk_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            optin.done(smem, dev);
        }
        RrProfScope prof(RR_PROF_SELECT_ROWS, stream);
        topk_small_kernel<<<rows, TS_THREADS, smem, stream>>>(d_scores, ld, (int)n, k, reinterpret_cast<long long*>(d_idx), d_score,
                                                            d_count, out_ld, k_pad);
        RR_LAUNCH_CHECK();
        return RR_OK;
    }
    if (tree) {
        const int k_pad = max(2, next_pow2(k));
        static RrSmemOptIn optin;
        int dev = 0;
        const size_t smem_max = sizeof(unsigned long long) * ((size_t)TS_MAX_N + TC_MAX_K);
        if (optin.needed(smem_max, &dev)) {
            RR_CUDA(cudaFuncSetAttribute(topk_chunk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
            RR_CUDA(cudaFuncSetAttribute(topk_chunk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
            RR_CUDA(cudaFuncSetAttribute(topk_chunk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
            optin.done(smem_max, dev);
        }
        unsigned long long* buf[2];
        buf[0] = static_cast<unsigned long long*>(d_scratch);
        buf[1] = buf[0] + (size_t)rows * plan.groups[0] * k;
        RrProfScope prof(RR_PROF_SELECT_ROWS, stream);
        long long cur_n = n;
        const unsigned long long* src = nullptr;
        for (int l = 0; l < plan.levels; ++l) {
            const int chunk = plan.chunk[l], groups = plan.groups[l];
            const size_t smem = sizeof(unsigned long long) * ((size_t)chunk + k_pad);
            const dim3 grid((unsigned)groups, (unsigned)rows);
            unsigned long long* dst = buf[l & 1];
            const long long keys_ld = (long long)groups * k;
            if (l == 0)
                topk_chunk_kernel<false, false><<<grid, TS_THREADS, smem, stream>>>(
                    d_scores, nullptr, ld, cur_n, chunk, k, dst, keys_ld, nullptr, nullptr, nullptr, 0, k_pad);
            else if (l + 1 < plan.levels)
                topk_chunk_kernel<true, false><<<grid, TS_THREADS, smem, stream>>>(
                    nullptr, src, cur_n, cur_n, chunk, k, dst, keys_ld, nullptr, nullptr, nullptr, 0, k_pad);
            else
                topk_chunk_kernel<true, true><<<grid, TS_THREADS, smem, stream>>>(
                    nullptr, src, cur_n, cur_n, chunk, k, nullptr, 0, reinterpret_cast<long long*>(d_idx), d_score, d_count,
                    out_ld, k_pad);
            RR_LAUNCH_CHECK();
            src = dst;
            cur_n = keys_ld;
        }
        return RR_OK;
    }
    char* p = static_cast<char*>(d_scratch);
    RsRow* st = reinterpret_cast<RsRow*>(p);
    p += (sizeof(RsRow) * (size_t)rows + 255) & ~(size_t)255;
    unsigned* hist = reinterpret_cast<unsigned*>(p);
    p += sizeof(unsigned) * (size_t)rows * RS_BINS;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(p);
    const int k_cap = max(kk, 1);
    RrProfScope prof(RR_PROF_SELECT_ROWS, stream);

    rs_init_kernel<<<max(1, min(1024, rows * 8)), 256, 0, stream>>>(st, hist, rows, kk);
    RR_LAUNCH_CHECK();
    if (kk > 0 && kk < n) {
        long long chunks = (n + (long long)RS_THREADS * 16 - 1) / ((long long)RS_THREADS * 16);
        unsigned gx = (unsigned)max(1ll, min(chunks, (long long)max(1, sm_count * 4 / max(1, min(rows, 4)))));
        for (int level = 0; level < RS_LEVELS; ++level) {
            rs_hist_kernel<<<dim3(gx, (unsigned)rows), RS_THREADS, 0, stream>>>(d_scores, ld, n, st, hist, level);
            RR_LAUNCH_CHECK();
            rs_scan_kernel<<<rows, RS_BINS / 2, 0, stream>>>(st, hist, level);
            RR_LAUNCH_CHECK();
        }
        rs_collect_kernel<<<dim3(gx, (unsigned)rows), RS_THREADS, 0, stream>>>(d_scores, ld, n, st, keys, k_cap);
        RR_LAUNCH_CHECK();
    } else if (kk > 0) {
        // k >= n: everything is selected (bits_done == 0 selects all)
        long long chunks = (n + (long long)RS_THREADS * 16 - 1) / ((long long)RS_THREADS * 16);
        unsigned gx = (unsigned)max(1ll, min(chunks, (long long)sm_count));
        rs_collect_kernel<<<dim3(gx, (unsigned)rows), RS_THREADS, 0, stream>>>(d_scores, ld, n, st, keys, k_cap);
        RR_LAUNCH_CHECK();
    }
    const int n_pad = max(2, next_pow2(k_cap));
    const size_t smem = sizeof(unsigned long long) * (size_t)n_pad;
    if (smem > 48 * 1024) RR_CUDA(cudaFuncSetAttribute(sort_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = max(32, min(1024, n_pad / 2));
    sort_keys_kernel<<<rows, threads, smem, stream>>>(keys, k_cap, st, k, n_pad, reinterpret_cast<long long*>(d_idx),
                                                      d_score, d_count, out_ld);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
