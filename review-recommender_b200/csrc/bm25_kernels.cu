// K1 -- BM25 scoring over the tile-blocked CSR postings (sm_100a).
//
// Stands behind rank_bm25.BM25Okapi.get_scores as the reference calls it
// (app/test.py:170, app/app_product_search.py:206) and behind the candidate gathers that follow
// it (app/test.py:171-173, app/app_product_search.py:207-208).
//
// Layout (built by bm25_build.cpp on the host or bm25_build_gpu.cu on the device): docs are cut into tiles of T docs.  Inside tile i the
// postings {u32 doc, f32 impact} are grouped by term (doc ascending inside a term);
// blk_off[i*(V+1)+t .. +t+1] bounds term t's segment relative to tile_base[i] (always even, so
// a tile's postings start on a 16-byte boundary).
//
//  bm25_tile_scores_kernel   one CTA per (tile, query): T fp32 accumulators live in shared
//      memory, the CTA streams the <= L segments of its query's terms with 16-byte loads
//      (2 postings per load, 4 loads in flight per thread; latency is hidden by 4 resident
//      CTAs per SM), adds impacts in QUERY-TERM ORDER (a barrier
//      separates consecutive terms, postings of one term hit distinct docs so plain
//      read-modify-write is race free), then writes the tile's scores coalesced.
//      HBM traffic = 8 B per posting of the query's terms + 4 B per doc: the algorithmic bytes.
//      Summation order = the reference's (`score += ...` per query token), in fp32.
//
//  bm25_candidates_kernel    one thread per (query, candidate): binary search of each query
//      term in the candidate's forward list (or of the candidate's doc id inside the term's
//      segment of the doc's tile when no forward index is loaded); same impacts, same order =>
//      bit-identical to the tile kernel at those docs.  Also gathers n_reviews / avg_stars /
//      global row so that the fusion kernel gets complete tuples, optionally straight into the
//      all-to-all send buffer of a sharded search.
#include <cstdlib>
#include <type_traits>

#include "rr_internal.h"

namespace {

constexpr int BM25_THREADS = 256;
constexpr int BM25_U = 4;       // 16-byte units (2 postings each) per thread per round
constexpr int BM25_MAXL = 64;   // query terms staged per pass over the accumulators
constexpr int BM25_MIN_CTAS = 4;

__device__ __forceinline__ uint4 ldg_stream16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// One CTA per (doc tile, query).  Latency is hidden by residency (4 CTAs of 256 threads per SM with 12288-doc
// tiles = 48 KB of accumulators each; every thread keeps 4 x 16 B loads in flight), not by register double-buffering:
// the r01 profile showed the double-buffered 512-thread version at 92 registers -> 1 CTA/SM -> 31 % of HBM.
__global__ void __launch_bounds__(BM25_THREADS, BM25_MIN_CTAS)
bm25_tile_scores_kernel(const uint4* __restrict__ postings, const uint64_t* __restrict__ tile_base,
                        const uint32_t* __restrict__ blk_off, int V, int T, long long n_docs,
                        const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_len, int l_max,
                        float* __restrict__ out, long long ld_out) {
    extern __shared__ __align__(16) float acc[];
    __shared__ uint32_t s_lo[BM25_MAXL], s_hi[BM25_MAXL];
    __shared__ uint32_t s_ustart[BM25_MAXL + 1];

    const int tile = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    const long long doc0 = (long long)tile * T;
    const uint32_t doc0u = (uint32_t)doc0;
    const int tile_n = (int)min((long long)T, n_docs - doc0);

    int L = q_len[q];
    if (L > l_max) L = l_max;
    const uint4* base = postings + (tile_base[tile] >> 1);
    const uint32_t* off = blk_off + (long long)tile * (V + 1);
    constexpr int R = BM25_THREADS * BM25_U;

    for (int i = tid; i < T / 4; i += BM25_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int l0 = 0; l0 < L; l0 += BM25_MAXL) {
        const int nl = min(BM25_MAXL, L - l0);
        __syncthreads();   // previous pass done with s_*, accumulators zeroed / settled
        if (tid < nl) {
            const int t = q_terms[(long long)q * l_max + l0 + tid];
            uint32_t lo = 0, hi = 0;
            if (t >= 0 && t < V) { lo = off[t]; hi = off[t + 1]; }
            s_lo[tid] = lo;
            s_hi[tid] = hi;
        }
        __syncthreads();
        if (tid < 32) {
            // exclusive scan of the per-term unit counts (nl <= 64: two per lane)
            const int i0 = tid, i1 = tid + 32;
            uint32_t u0 = 0, u1 = 0;
            if (i0 < nl) { const uint32_t lo = s_lo[i0], hi = s_hi[i0]; u0 = hi > lo ? ((hi + 1) >> 1) - (lo >> 1) : 0u; }
            if (i1 < nl) { const uint32_t lo = s_lo[i1], hi = s_hi[i1]; u1 = hi > lo ? ((hi + 1) >> 1) - (lo >> 1) : 0u; }
            uint32_t x0 = u0, x1 = u1;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
                if (tid >= o) { x0 += y0; x1 += y1; }
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, x0, 31);
            if (i0 < nl) s_ustart[i0] = x0 - u0;
            if (i1 < nl) s_ustart[i1] = tot0 + x1 - u1;
            if (tid == 31) s_ustart[nl] = tot0 + x1;      // lane 31's inclusive sums are the totals
        }
        __syncthreads();
        const uint32_t total = s_ustart[nl];
        if (total == 0) continue;
        const int n_rounds = (int)((total + R - 1) / R);

        int load_cursor = 0;     // per-thread, monotone
        int phase_first = 0;     // block-uniform, monotone
        int prev_seg = -1;       // block-uniform: last term accumulated

        // issue this thread's BM25_U 16-byte loads of round r (no wait)
        auto load_round = [&](int r, uint4 (&p)[BM25_U], uint32_t (&unit_of)[BM25_U], int (&seg)[BM25_U]) {
#pragma unroll
            for (int u = 0; u < BM25_U; ++u) {
                const uint32_t v = (uint32_t)(r * R + u * BM25_THREADS + tid);
                seg[u] = -1;
                if (v < total) {
                    while (v >= s_ustart[load_cursor + 1]) ++load_cursor;
                    const uint32_t unit = (s_lo[load_cursor] >> 1) + (v - s_ustart[load_cursor]);
                    p[u] = ldg_stream16(base + unit);
                    unit_of[u] = unit;
                    seg[u] = load_cursor;
                }
            }
        };
        // add round r's impacts, term by term in query order (a barrier separates consecutive terms)
        auto add_round = [&](int r, const uint4 (&p)[BM25_U], const uint32_t (&unit_of)[BM25_U], const int (&seg)[BM25_U]) {
            const uint32_t v_first = (uint32_t)r * R;
            const uint32_t v_last = min(total, v_first + (uint32_t)R) - 1u;
            while (v_first >= s_ustart[phase_first + 1]) ++phase_first;
            int phase_last = phase_first;
            while (v_last >= s_ustart[phase_last + 1]) ++phase_last;
            for (int s = phase_first; s <= phase_last; ++s) {
                if (s_ustart[s + 1] == s_ustart[s]) continue;
                if (s != prev_seg) {
                    __syncthreads();    // all adds of the previous term are done
                    prev_seg = s;
                }
                const uint32_t lo = s_lo[s], hi = s_hi[s];
#pragma unroll
                for (int u = 0; u < BM25_U; ++u) {
                    if (seg[u] == s) {
                        const uint32_t i0 = unit_of[u] * 2u;
                        if (i0 >= lo && i0 < hi) {
                            const uint32_t d = p[u].x - doc0u;
                            acc[d] = __fadd_rn(acc[d], __uint_as_float(p[u].y));
                        }
                        if (i0 + 1u >= lo && i0 + 1u < hi) {
                            const uint32_t d = p[u].z - doc0u;
                            acc[d] = __fadd_rn(acc[d], __uint_as_float(p[u].w));
                        }
                    }
                }
            }
        };

        // one round of loads in flight per thread.  Issuing round r+1 before accumulating round r (two rounds in
        // flight, 80 registers, 3 CTAs/SM) was measured SLOWER on B200 (r02: 49 % vs 62 % of HBM peak at 8 M docs),
        // as was the 512-thread register double buffer of r01: latency is hidden by residency instead.
        uint4 pa[BM25_U];
        uint32_t ua[BM25_U];
        int sa[BM25_U];
        for (int r = 0; r < n_rounds; ++r) {
            load_round(r, pa, ua, sa);
            add_round(r, pa, ua, sa);
        }
    }
    __syncthreads();
    float* dst = out + (long long)q * ld_out + doc0;
    const int n4 = tile_n >> 2;
    for (int i = tid; i < n4; i += BM25_THREADS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(acc)[i];
    for (int i = (n4 << 2) + tid; i < tile_n; i += BM25_THREADS) dst[i] = acc[i];
}

__global__ void __launch_bounds__(256)
bm25_candidates_kernel(const uint2* __restrict__ postings, const uint64_t* __restrict__ tile_base,
                       const uint32_t* __restrict__ blk_off, const unsigned long long* __restrict__ fwd_off,
                       const uint2* __restrict__ fwd_data, int V, int T, long long n_docs,
                       const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_len, int l_max,
                       const long long* __restrict__ cand, int pool, int B,
                       const double* __restrict__ n_reviews, const double* __restrict__ avg_stars, long long row_offset,
                       float* __restrict__ bm25_out, double* __restrict__ n_out, double* __restrict__ avg_out,
                       long long* __restrict__ grow_out, int pack_bg, long long pack_stride,
                       const float* __restrict__ dense_in, float* __restrict__ dense_out,
                       const int* __restrict__ uncertified) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * pool) return;
    const int q = (int)(gid / pool);
    // exchange layout (pack_bg > 0): the outputs are the field bases of destination rank 0's block; query q goes
    // to rank q / pack_bg, whose block of every field starts pack_stride bytes after the previous rank's
    long long oid = gid;
    long long shift = 0;
    if (pack_bg > 0) {
        const int g = q / pack_bg;
        oid = (long long)(q - g * pack_bg) * pool + (gid - (long long)q * pool);
        shift = (long long)g * pack_stride;
    }
    auto at = [shift](auto* base, long long i) {
        using T = std::remove_pointer_t<decltype(base)>;
        return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + shift) + i;
    };
    const long long doc = cand[gid];
    const bool valid = doc >= 0 && doc < n_docs;
    float sum = 0.f;
    if (valid && V > 0 && q_terms != nullptr && fwd_off != nullptr) {
        // forward index: the doc's own {term, impact} list (term-ascending, ~35 entries, contiguous)
        const unsigned long long f0 = fwd_off[doc], f1 = fwd_off[doc + 1];
        const uint2* list = fwd_data + f0;
        const int n = (int)(f1 - f0);
        int L = q_len[q];
        if (L > l_max) L = l_max;
        for (int l = 0; l < L; ++l) {
            const int t = q_terms[(long long)q * l_max + l];
            if (t < 0 || t >= V) continue;
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(&list[mid].x) < (uint32_t)t) lo = mid + 1; else hi = mid;
            }
            if (lo < n) {
                const uint2 e = __ldg(&list[lo]);
                if (e.x == (uint32_t)t) sum = __fadd_rn(sum, __uint_as_float(e.y));
            }
        }
    } else if (valid && V > 0 && q_terms != nullptr) {
        const int tile = (int)(doc / T);
        const uint2* base = postings + tile_base[tile];
        const uint32_t* off = blk_off + (long long)tile * (V + 1);
        const uint32_t d = (uint32_t)doc;
        int L = q_len[q];
        if (L > l_max) L = l_max;
        for (int l = 0; l < L; ++l) {
            const int t = q_terms[(long long)q * l_max + l];
            if (t < 0 || t >= V) continue;
            uint32_t lo = off[t], hi = off[t + 1];
            while (lo < hi) {                       // lower_bound on the doc ids of the segment
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&base[mid].x) < d) lo = mid + 1; else hi = mid;
            }
            if (lo < off[t + 1]) {
                const uint2 e = __ldg(&base[lo]);
                if (e.x == d) sum = __fadd_rn(sum, __uint_as_float(e.y));
            }
        }
    }
    if (bm25_out) *at(bm25_out, oid) = sum;
    if (n_out) *at(n_out, oid) = (valid && n_reviews) ? n_reviews[doc] : 0.0;
    if (avg_out) *at(avg_out, oid) = (valid && avg_stars) ? avg_stars[doc] : __longlong_as_double(0x7FF8000000000000ll);
    // -2 = "this shard could not certify its dense result for the query": the merging rank repeats the query
    const bool poisoned = uncertified != nullptr && uncertified[q] != 0;
    if (grow_out) *at(grow_out, oid) = poisoned ? -2 : (valid ? row_offset + doc : -1);
    if (dense_out) *at(dense_out, oid) = dense_in[gid];
}

}  // namespace

int rr_launch_bm25_tile_scores(const uint64_t* d_postings, const uint64_t* d_tile_base, const uint32_t* d_blk_off,
                               int V, int T, int n_tiles, int64_t n_docs, const int32_t* d_terms,
                               const int32_t* d_nterms, int B, int l_max, float* d_out, int64_t ld_out,
                               cudaStream_t stream) {
    if (B <= 0 || n_tiles <= 0) return RR_OK;
    const size_t smem = (size_t)T * sizeof(float);
    static RrSmemOptIn optin;
    int dev = 0;
    if (optin.needed(smem, &dev)) {
        RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // without this the driver sizes the shared-memory carveout for ONE resident CTA (ncu r01: occupancy limit 1)
        RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        optin.done(smem, dev);
    }
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = min(65535, B - b0);
        dim3 grid((unsigned)n_tiles, (unsigned)nb);
        RrProfScope prof(RR_PROF_BM25_TILE, stream);
        bm25_tile_scores_kernel<<<grid, BM25_THREADS, smem, stream>>>(
            reinterpret_cast<const uint4*>(d_postings), d_tile_base, d_blk_off, V, T, (long long)n_docs,
            d_terms + (int64_t)b0 * l_max, d_nterms + b0, l_max, d_out + (int64_t)b0 * ld_out, (long long)ld_out);
        RR_LAUNCH_CHECK();
    }
    return RR_OK;
}

int rr_launch_bm25_candidates(const uint64_t* d_postings, const uint64_t* d_tile_base, const uint32_t* d_blk_off,
                              const uint64_t* d_fwd_off, const uint64_t* d_fwd_data, int V, int T, int64_t n_docs, const int32_t* d_terms, const int32_t* d_nterms,
                              int B, int l_max, const int64_t* d_cand, int pool, const double* d_nrev,
                              const double* d_avg, int64_t row_offset, float* d_bm25, double* d_n_out,
                              double* d_avg_out, int64_t* d_grow_out, cudaStream_t stream, int pack_bg,
                              int64_t pack_stride, const float* d_dense_in, float* d_dense_out,
                              const int32_t* d_uncertified) {
    const int64_t total = (int64_t)B * pool;
    if (total <= 0) return RR_OK;
    const int threads = 256;
    RrProfScope prof(RR_PROF_BM25_CAND, stream);
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    bm25_candidates_kernel<<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const uint2*>(d_postings), d_tile_base, d_blk_off,
        reinterpret_cast<const unsigned long long*>(d_fwd_off), reinterpret_cast<const uint2*>(d_fwd_data), V, T,
        (long long)n_docs, d_terms, d_nterms,
        l_max, reinterpret_cast<const long long*>(d_cand), pool, B, d_nrev, d_avg, (long long)row_offset, d_bm25,
        d_n_out, d_avg_out, reinterpret_cast<long long*>(d_grow_out), pack_bg, (long long)pack_stride, d_dense_in,
        d_dense_out, d_uncertified);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
