// K1 -- BM25 scoring over the hybrid-blocked postings (sm_100a).
//
// Stands behind rank_bm25.BM25Okapi.get_scores as the reference calls it
// (app/test.py:170, app/app_product_search.py:206) and behind the candidate gathers that follow
// it (app/test.py:171-173, app/app_product_search.py:207-208).
//
// Layout (include/rr_b200.h "index layout"; built by bm25_build.cpp on the host or bm25_build_gpu.cu on the device):
// docs are cut into tiles of T docs.  FREQUENT terms (directory slot f = term_slot[t]) are tile-blocked: inside tile i
// the postings {u32 doc, f32 impact} are grouped by slot, doc-ascending, bounded by dir[i*(n_freq+1)+f .. +f+1]
// relative to tile_base[i] (tiles start on a 16-byte boundary).  RARE terms have one term-major, doc-ascending list
// behind the frequent region; bm25_rare_bounds_kernel finds, per batch, where every tile starts in the lists of the
// batch's rare query terms.
//
//  bm25_tile_scores_persistent_kernel (default, query term lists of <= 64 terms) / bm25_tile_scores_kernel (one CTA per
//      (query, tile), any length):  T fp32 accumulators live in shared memory.  A PRODUCER warp
//      streams the <= 64 term segments of an item with 1-D TMA bulk copies (cp.async.bulk, mbarrier complete_tx)
//      into a shared-memory ring, running up to NSTAGE chunks ahead of the eight CONSUMER warps, which add the
//      impacts in QUERY-TERM ORDER (a named barrier separates consecutive terms; postings of one term hit distinct
//      docs, so the plain read-modify-write is race free and the fp32 sum is deterministic), then write the tile's
//      scores with coalesced 16-byte stores.  Loads are decoupled from the per-term barriers: the r01 kernel issued
//      a round of register loads only after the previous round's adds (ncu: long-scoreboard + barrier stalls).
//      The persistent variant keeps 3 CTAs per SM resident, pulls (query, tile) items from a global counter and lets the
//      producer work ONE ITEM AHEAD (bounds double-buffered), which hides the directory look-ups, the zero-fill and the
//      write-out of an item behind the next item's stream: 0.95 of the measured HBM peak at configs[3] (r01: 0.69).
//      Items are query-fastest, so the CTAs that share a tile (and, for common terms, its segments) run together
//      and the shared segments are served by L2.
//      HBM traffic <= 8 B per posting of the query's terms + 4 B per doc: the algorithmic bytes.
//      Summation: fp32 adds in the order of the query's token list.  rank_bm25 adds float64 terms and the callers cast
//      once to float32; here every term is the correctly rounded fp32 of the float64 term and the adds are fp32, so
//      a score differs from the reference's by at most ~L * 2^-24 relative (tests pin 1e-5, north_star's bound).
//
//  bm25_candidates_kernel    one thread per (query, candidate): binary search of each query
//      term in the candidate's forward list (or of the candidate's doc id inside the term's
//      segment / rare list when no forward index is loaded); same impacts, same order =>
//      bit-identical to the tile kernel at those docs.  Also gathers n_reviews / avg_stars /
//      global row so that the fusion kernel gets complete tuples, optionally straight into the
//      all-to-all send buffer of a sharded search.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "rr_internal.h"

namespace {

constexpr int BM25_CONSUMERS = 256;                 // 8 consumer warps
constexpr int BM25_THREADS = BM25_CONSUMERS + 32;   // + 1 producer warp
constexpr int BM25_MAXL = 64;                       // query terms staged per pass over the accumulators
constexpr int BM25_MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BM25_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra BM25_DONE;\n\t"
        "bra BM25_WAIT;\n\t"
        "BM25_DONE:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 1-D TMA: `bytes` (multiple of 16) from global to shared, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(BM25_CONSUMERS) : "memory"); }

struct Bm25Args {
    const uint4* postings;            // 16-byte units (2 postings each)
    const uint64_t* tile_base;        // [n_tiles+1]
    const uint32_t* dir;              // [n_tiles, n_freq+1]
    const int32_t* term_slot;         // [V]
    const uint32_t* rtab;             // [B, l_max, n_tiles+1] rare-list tile bounds of this batch (relative to the rare
                                      // region), or NULL: every CTA searches its own bounds (small batches)
    const unsigned long long* rare_off;   // [V+1]
    int V, T, n_freq, n_tiles, tile0;
    long long n_docs;
    const int32_t* q_terms;
    const int32_t* q_len;
    int l_max;
    float* out;
    long long ld_out;
    int stage_units, n_stages;
};

// warp-cooperative 32-ary lower bound: first index in [0, n) whose doc id is >= target (3 round trips for 16 k entries)
__device__ __forceinline__ uint32_t warp_lower_bound(const uint2* __restrict__ list, uint32_t n, uint32_t target, int lane) {
    uint32_t lo = 0, hi = n;
    while (hi > lo) {
        const uint32_t span = hi - lo;
        if (span <= 32u) {
            const bool less = lo + (uint32_t)lane < hi && __ldg(&list[lo + lane].x) < target;
            return lo + (uint32_t)__popc(__ballot_sync(0xffffffffu, less));
        }
        const uint32_t step = (span + 31u) >> 5;
        const uint32_t p = min(hi - 1u, lo + (uint32_t)(lane + 1) * step - 1u);
        const bool less = __ldg(&list[p].x) < target;
        const uint32_t c = (uint32_t)__popc(__ballot_sync(0xffffffffu, less));
        const uint32_t nlo = min(hi, lo + c * step);
        hi = c >= 32u ? hi : min(hi, lo + (c + 1u) * step);
        lo = nlo;
    }
    return lo;
}

__global__ void __launch_bounds__(BM25_THREADS, 4)
bm25_tile_scores_kernel(const Bm25Args a) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    float* acc = reinterpret_cast<float*>(bm25_smem);
    uint4* ring = reinterpret_cast<uint4*>(bm25_smem + (size_t)a.T * sizeof(float));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.n_stages * a.stage_units);
    uint64_t* empty = full + BM25_MAX_STAGES;
    __shared__ unsigned long long s_lo[BM25_MAXL], s_hi[BM25_MAXL];

    const int q = blockIdx.x, tile = a.tile0 + blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == BM25_CONSUMERS / 32;
    const long long doc0 = (long long)tile * a.T;
    const uint32_t doc0u = (uint32_t)doc0;
    const int tile_n = (int)min((long long)a.T, a.n_docs - doc0);
    const int SU = a.stage_units, NS = a.n_stages;

    int L = a.q_len[q];
    if (L > a.l_max) L = a.l_max;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], BM25_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!producer)
        for (int i = tid; i < a.T / 4; i += BM25_CONSUMERS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    int stage = 0;
    uint32_t phase = 0;
    bool dirty = false;                 // consumers: some term has been accumulated since the last consumer barrier
    for (int l0 = 0; l0 < L; l0 += BM25_MAXL) {
        const int nl = min(BM25_MAXL, L - l0);
        __syncthreads();                // barriers initialised / previous pass done with s_lo, s_hi
        if (tid < nl) {
            const int l = l0 + tid;
            const int t = a.q_terms[(long long)q * a.l_max + l];
            unsigned long long lo = 0, hi = 0;
            if (t >= 0 && t < a.V) {
                const int slot = a.term_slot[t];
                if (slot >= 0) {
                    const unsigned long long base = a.tile_base[tile];
                    const uint32_t* d = a.dir + (long long)tile * (a.n_freq + 1) + slot;
                    lo = base + d[0];
                    hi = base + d[1];
                } else if (a.rtab != nullptr) {
                    const unsigned long long base = a.tile_base[a.n_tiles];
                    const uint32_t* r = a.rtab + ((long long)q * a.l_max + l) * (a.n_tiles + 1) + tile;
                    lo = base + r[0];
                    hi = base + r[1];
                } else {
                    lo = ~0ull;                       // rare term, bounds searched below by a warp
                    hi = (unsigned long long)(uint32_t)t;
                }
            }
            s_lo[tid] = lo;
            s_hi[tid] = hi;
        }
        __syncthreads();
        if (a.rtab == nullptr) {
            // small batches: no pre-pass; every warp resolves the rare terms l = warp, warp + 9, ... of this CTA
            for (int l = warp; l < nl; l += BM25_THREADS / 32) {
                if (s_lo[l] != ~0ull) continue;
                const int t = (int)s_hi[l];
                const unsigned long long b0 = a.rare_off[t], b1 = a.rare_off[t + 1];
                const unsigned long long base = a.tile_base[a.n_tiles] + b0;
                const uint2* list = reinterpret_cast<const uint2*>(a.postings) + base;
                const uint32_t n = (uint32_t)(b1 - b0);
                const uint32_t r0 = warp_lower_bound(list, n, doc0u, lane);
                const uint32_t r1 = tile + 1 < a.n_tiles ? warp_lower_bound(list, n, doc0u + (uint32_t)a.T, lane) : n;
                if (lane == 0) { s_lo[l] = base + r0; s_hi[l] = base + r1; }
            }
            __syncthreads();
        }
        if (producer) {
            // ===== producer warp (one lane issues): chunk c of term l -> ring stage, up to NS chunks ahead =====
            if (lane == 0) {
                for (int l = 0; l < nl; ++l) {
                    const unsigned long long lo = s_lo[l], hi = s_hi[l];
                    if (hi <= lo) continue;
                    const unsigned long long u0 = lo >> 1, u1 = (hi + 1) >> 1;
                    for (unsigned long long u = u0; u < u1; u += (unsigned long long)SU) {
                        const uint32_t n = (uint32_t)min((unsigned long long)SU, u1 - u);
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], n * 16u);
                        bulk_g2s(ring + (size_t)stage * SU, a.postings + u, n * 16u, &full[stage]);
                        if (++stage == NS) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        } else {
            // ===== consumers: impacts added term by term in query order =====
            for (int l = 0; l < nl; ++l) {
                const unsigned long long lo = s_lo[l], hi = s_hi[l];
                if (hi <= lo) continue;
                if (dirty) consumer_barrier();          // all adds of the previous term are done
                dirty = true;
                const uint32_t n_post = (uint32_t)(hi - lo);
                const int odd = (int)(lo & 1ull);
                const uint32_t n_units = (uint32_t)(((hi + 1) >> 1) - (lo >> 1));
                for (uint32_t c0 = 0; c0 < n_units; c0 += (uint32_t)SU) {
                    const uint32_t n = min((uint32_t)SU, n_units - c0);
                    mbar_wait(&full[stage], phase);
                    const uint4* src = ring + (size_t)stage * SU;
                    for (uint32_t u = (uint32_t)tid; u < n; u += BM25_CONSUMERS) {
                        const uint4 p = src[u];
                        const int r0 = (int)(2u * (c0 + u)) - odd;        // posting index relative to lo (-1: before it)
                        if ((uint32_t)r0 < n_post) {
                            const uint32_t d = p.x - doc0u;
                            acc[d] = __fadd_rn(acc[d], __uint_as_float(p.y));
                        }
                        if ((uint32_t)(r0 + 1) < n_post) {
                            const uint32_t d = p.z - doc0u;
                            acc[d] = __fadd_rn(acc[d], __uint_as_float(p.w));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            }
        }
    }
    if (producer) return;
    consumer_barrier();
    float* dst = a.out + (long long)q * a.ld_out + doc0;
    const int n4 = tile_n >> 2;
    for (int i = tid; i < n4; i += BM25_CONSUMERS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(acc)[i];
    for (int i = (n4 << 2) + tid; i < tile_n; i += BM25_CONSUMERS) dst[i] = acc[i];
}

// Persistent variant (query term lists of at most 64 terms): CTAs stay resident and pull (query, tile) items from a
// global counter.  The producer warp works one item AHEAD of the consumers: while they still accumulate / write out
// item k it has already looked up the segment bounds of item k+1 (double-buffered in shared memory) and keeps the ring
// full with its first chunks, so the directory round trips, the accumulator zero-fill and the write-out of one item
// overlap with the posting stream of the next -- the per-CTA prologue that the one-shot kernel above pays per item
// (4-5 dependent global loads, ~30 % of a 4-term item's lifetime) is hidden, and there are no waves to quantise.
__global__ void __launch_bounds__(BM25_THREADS, 3)
bm25_tile_scores_persistent_kernel(const Bm25Args a, int n_items, int B, unsigned* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    float* acc = reinterpret_cast<float*>(bm25_smem);
    uint4* ring = reinterpret_cast<uint4*>(bm25_smem + (size_t)a.T * sizeof(float));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.n_stages * a.stage_units);
    uint64_t* empty = full + BM25_MAX_STAGES;
    __shared__ unsigned long long s_lo[2][BM25_MAXL], s_hi[2][BM25_MAXL];
    __shared__ int s_item[2], s_nl[2];
    __shared__ uint64_t bfull[2], bempty[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == BM25_CONSUMERS / 32;
    const int SU = a.stage_units, NS = a.n_stages;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], BM25_CONSUMERS / 32); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], BM25_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;

    if (producer) {
        for (int k = 0;; ++k) {
            const int buf = k & 1;
            mbar_wait(&bempty[buf], (uint32_t)(((k >> 1) & 1) ^ 1));       // consumers are done with item k-2's bounds
            int item = 0;
            if (lane == 0) item = (int)atomicAdd(counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) {
                if (lane == 0) {
                    s_item[buf] = -1;
                    mbar_arrive(&bfull[buf]);
                    // the last CTA to run dry re-arms the two counters for the next launch (no memset per call)
                    if (atomicAdd(counter + 1, 1u) == gridDim.x - 1) { counter[0] = 0u; counter[1] = 0u; __threadfence(); }
                }
                break;
            }
            const int q = item % B, tile = a.tile0 + item / B;
            int L = a.q_len[q];
            if (L > a.l_max) L = a.l_max;
            if (L > BM25_MAXL) L = BM25_MAXL;
            const uint32_t doc0u = (uint32_t)((long long)tile * a.T);
            for (int l = lane; l < L; l += 32) {
                const int t = a.q_terms[(long long)q * a.l_max + l];
                unsigned long long lo = 0, hi = 0;
                if (t >= 0 && t < a.V) {
                    const int slot = a.term_slot[t];
                    if (slot >= 0) {
                        const unsigned long long base = a.tile_base[tile];
                        const uint32_t* d = a.dir + (long long)tile * (a.n_freq + 1) + slot;
                        lo = base + d[0];
                        hi = base + d[1];
                    } else if (a.rtab != nullptr) {
                        const unsigned long long base = a.tile_base[a.n_tiles];
                        const uint32_t* r = a.rtab + ((long long)q * a.l_max + l) * (a.n_tiles + 1) + tile;
                        lo = base + r[0];
                        hi = base + r[1];
                    } else {
                        lo = ~0ull;
                        hi = (unsigned long long)(uint32_t)t;
                    }
                }
                s_lo[buf][l] = lo;
                s_hi[buf][l] = hi;
            }
            __syncwarp();
            if (a.rtab == nullptr) {
                for (int l = 0; l < L; ++l) {
                    if (s_lo[buf][l] != ~0ull) continue;
                    const int t = (int)s_hi[buf][l];
                    const unsigned long long b0 = a.rare_off[t], b1 = a.rare_off[t + 1];
                    const unsigned long long base = a.tile_base[a.n_tiles] + b0;
                    const uint2* list = reinterpret_cast<const uint2*>(a.postings) + base;
                    const uint32_t n = (uint32_t)(b1 - b0);
                    const uint32_t r0 = warp_lower_bound(list, n, doc0u, lane);
                    const uint32_t r1 = tile + 1 < a.n_tiles ? warp_lower_bound(list, n, doc0u + (uint32_t)a.T, lane) : n;
                    __syncwarp();
                    if (lane == 0) { s_lo[buf][l] = base + r0; s_hi[buf][l] = base + r1; }
                    __syncwarp();
                }
            }
            if (lane == 0) {
                s_item[buf] = item;
                s_nl[buf] = L;
                mbar_arrive(&bfull[buf]);                                    // bounds of item k are published
                for (int l = 0; l < L; ++l) {
                    const unsigned long long lo = s_lo[buf][l], hi = s_hi[buf][l];
                    if (hi <= lo) continue;
                    const unsigned long long u0 = lo >> 1, u1 = (hi + 1) >> 1;
                    for (unsigned long long u = u0; u < u1; u += (unsigned long long)SU) {
                        const uint32_t n = (uint32_t)min((unsigned long long)SU, u1 - u);
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], n * 16u);
                        bulk_g2s(ring + (size_t)stage * SU, a.postings + u, n * 16u, &full[stage]);
                        if (++stage == NS) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            __syncwarp();
        }
        return;
    }

    // ===== consumers =====
    for (int k = 0;; ++k) {
        const int buf = k & 1;
        // own accumulator slots: written out by this thread at the end of the previous item, zeroed by it now
        for (int i = tid; i < a.T / 4; i += BM25_CONSUMERS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(&bfull[buf], (uint32_t)((k >> 1) & 1));
        const int item = s_item[buf];
        if (item < 0) break;
        const int nl = s_nl[buf];
        const int q = item % B, tile = a.tile0 + item / B;
        const long long doc0 = (long long)tile * a.T;
        const uint32_t doc0u = (uint32_t)doc0;
        const int tile_n = (int)min((long long)a.T, a.n_docs - doc0);
        for (int l = 0; l < nl; ++l) {
            const unsigned long long lo = s_lo[buf][l], hi = s_hi[buf][l];
            if (hi <= lo) continue;
            consumer_barrier();                 // accumulators zeroed / all adds of the previous term are done
            const uint32_t n_post = (uint32_t)(hi - lo);
            const int odd = (int)(lo & 1ull);
            const uint32_t n_units = (uint32_t)(((hi + 1) >> 1) - (lo >> 1));
            for (uint32_t c0 = 0; c0 < n_units; c0 += (uint32_t)SU) {
                const uint32_t n = min((uint32_t)SU, n_units - c0);
                mbar_wait(&full[stage], phase);
                const uint4* src = ring + (size_t)stage * SU;
                for (uint32_t u = (uint32_t)tid; u < n; u += BM25_CONSUMERS) {
                    const uint4 p = src[u];
                    const int r0 = (int)(2u * (c0 + u)) - odd;
                    if ((uint32_t)r0 < n_post) {
                        const uint32_t d = p.x - doc0u;
                        acc[d] = __fadd_rn(acc[d], __uint_as_float(p.y));
                    }
                    if ((uint32_t)(r0 + 1) < n_post) {
                        const uint32_t d = p.z - doc0u;
                        acc[d] = __fadd_rn(acc[d], __uint_as_float(p.w));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
        consumer_barrier();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bempty[buf]);            // the producer may reuse this bounds buffer
        float* dst = a.out + (long long)q * a.ld_out + doc0;
        const int n4 = tile_n >> 2;
        for (int i = tid; i < n4; i += BM25_CONSUMERS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(acc)[i];
        // a ragged tail (last tile only) is written by the owners of the covering float4 slots
        for (int i = (n4 << 2) + tid; i < tile_n; i += BM25_CONSUMERS) dst[i] = acc[i];
        if ((tile_n & 3) != 0) consumer_barrier();           // tail elements are read by other threads than their zero-fillers
    }
}

// Where tile i starts in the list of every RARE term of the batch: rtab[(q*l_max + l)*(n_tiles+1) + i] = number of
// postings of the term with doc < i*T (i = n_tiles: the list length).  One thread per (query, term slot, tile).
__global__ void __launch_bounds__(256)
bm25_rare_bounds_kernel(const uint2* __restrict__ postings, const uint64_t* __restrict__ tile_base,
                        const int32_t* __restrict__ term_slot, const unsigned long long* __restrict__ rare_off, int V, int T,
                        int n_tiles, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_len, int l_max, int B,
                        uint32_t* __restrict__ rtab) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_q = (long long)l_max * (n_tiles + 1);
    if (gid >= (long long)B * per_q) return;
    const int q = (int)(gid / per_q);
    const int l = (int)((gid - (long long)q * per_q) / (n_tiles + 1));
    const int i = (int)(gid - (long long)q * per_q - (long long)l * (n_tiles + 1));
    if (l >= q_len[q]) return;
    const int t = q_terms[(long long)q * l_max + l];
    if (t < 0 || t >= V || term_slot[t] >= 0) return;
    const unsigned long long b0 = rare_off[t], b1 = rare_off[t + 1];
    const uint2* list = postings + tile_base[n_tiles] + b0;
    uint32_t lo = 0, hi = (uint32_t)(b1 - b0);
    if (i < n_tiles) {
        const uint32_t target = (uint32_t)((long long)i * T);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&list[mid].x) < target) lo = mid + 1; else hi = mid;
        }
    } else {
        lo = hi;
    }
    rtab[gid] = (uint32_t)b0 + lo;
}

__global__ void __launch_bounds__(256)
bm25_candidates_kernel(const uint2* __restrict__ postings, const uint64_t* __restrict__ tile_base,
                       const uint32_t* __restrict__ dir, const int32_t* __restrict__ term_slot,
                       const unsigned long long* __restrict__ rare_off, int n_freq, int n_tiles,
                       const unsigned long long* __restrict__ fwd_off,
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
        // no forward index: look the doc up in the term's segment of the doc's tile (frequent) or in its list (rare)
        const int tile = (int)(doc / T);
        const uint32_t d = (uint32_t)doc;
        int L = q_len[q];
        if (L > l_max) L = l_max;
        for (int l = 0; l < L; ++l) {
            const int t = q_terms[(long long)q * l_max + l];
            if (t < 0 || t >= V) continue;
            const int slot = term_slot[t];
            const uint2* base;
            uint32_t lo, end;
            if (slot >= 0) {
                const uint32_t* off = dir + (long long)tile * (n_freq + 1) + slot;
                base = postings + tile_base[tile];
                lo = off[0]; end = off[1];
            } else {
                base = postings + tile_base[n_tiles] + rare_off[t];
                lo = 0; end = (uint32_t)(rare_off[t + 1] - rare_off[t]);
            }
            uint32_t hi = end;
            while (lo < hi) {                       // lower_bound on the doc ids
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&base[mid].x) < d) lo = mid + 1; else hi = mid;
            }
            if (lo < end) {
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


// The same gather for SMALL batches (a single query of the Streamlit app: 150 candidates), forward index only: one WARP per
// candidate.  The lanes read the document's whole {term, impact} list at once (one memory round trip instead of the ~6
// dependent probes per term of a binary search), every query term is matched by a ballot, and the impacts found are added
// in query-term order -- the additions of bm25_candidates_kernel, bit for bit (a term that is absent adds +0.0f to a sum
// that is never -0).
__global__ void __launch_bounds__(256)
bm25_candidates_warp_kernel(const unsigned long long* __restrict__ fwd_off, const uint2* __restrict__ fwd_data, int V,
                            long long n_docs, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_len,
                            int l_max, const long long* __restrict__ cand, int pool, int B,
                            const double* __restrict__ n_reviews, const double* __restrict__ avg_stars, long long row_offset,
                            float* __restrict__ bm25_out, double* __restrict__ n_out, double* __restrict__ avg_out,
                            long long* __restrict__ grow_out, int pack_bg, long long pack_stride,
                            const float* __restrict__ dense_in, float* __restrict__ dense_out,
                            const int* __restrict__ uncertified) {
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gid >= (long long)B * pool) return;                  // warp-uniform
    const int q = (int)(gid / pool);
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
    // metadata loads go out before the list is walked
    double nrev = 0.0, avg = __longlong_as_double(0x7FF8000000000000ll);
    if (lane == 0 && valid) {
        if (n_reviews) nrev = n_reviews[doc];
        if (avg_stars) avg = avg_stars[doc];
    }
    float sum = 0.f;
    if (valid && V > 0 && q_terms != nullptr) {
        const unsigned long long f0 = fwd_off[doc], f1 = fwd_off[doc + 1];
        const uint2* list = fwd_data + f0;
        const int n = (int)(f1 - f0);
        int L = q_len[q];
        if (L > l_max) L = l_max;
        for (int l0 = 0; l0 < L; l0 += 32) {                 // query terms in blocks of 32: lane j owns term l0 + j
            const int nl = min(32, L - l0);
            int my_term = -1;
            if (lane < nl) {
                my_term = q_terms[(long long)q * l_max + l0 + lane];
                if (my_term >= V) my_term = -1;
            }
            float my_impact = 0.f;
            for (int c0 = 0; c0 < n; c0 += 32) {
                uint2 e = make_uint2(0xFFFFFFFFu, 0u);
                if (c0 + lane < n) e = __ldg(&list[c0 + lane]);
                for (int l = 0; l < nl; ++l) {
                    const int t = __shfl_sync(0xffffffffu, my_term, l);
                    if (t < 0) continue;
                    const unsigned hit = __ballot_sync(0xffffffffu, e.x == (uint32_t)t);
                    if (hit) {
                        const float imp = __uint_as_float(__shfl_sync(0xffffffffu, e.y, __ffs(hit) - 1));
                        if (lane == l) my_impact = imp;
                    }
                }
            }
            for (int l = 0; l < nl; ++l) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, my_impact, l));
        }
    }
    if (lane != 0) return;
    if (bm25_out) *at(bm25_out, oid) = sum;
    if (n_out) *at(n_out, oid) = (valid && n_reviews) ? nrev : 0.0;
    if (avg_out) *at(avg_out, oid) = avg;
    const bool poisoned = uncertified != nullptr && uncertified[q] != 0;
    if (grow_out) *at(grow_out, oid) = poisoned ? -2 : (valid ? row_offset + doc : -1);
    if (dense_out) *at(dense_out, oid) = dense_in[gid];
}

}  // namespace

size_t rr_bm25_rtab_bytes(int B, int l_max, int n_tiles) {
    return sizeof(uint32_t) * (size_t)B * (size_t)l_max * ((size_t)n_tiles + 1);
}

static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    if (!e) return dflt;
    const int v = atoi(e);
    return v < lo || v > hi ? dflt : v;
}

int rr_launch_bm25_tile_scores(const rr_index_desc* d, const int32_t* d_terms, const int32_t* d_nterms, int B, int l_max,
                               float* d_out, int64_t ld_out, uint32_t* d_rtab, unsigned* d_counter, cudaStream_t stream) {
    if (B <= 0 || d->n_tiles <= 0) return RR_OK;
    const int T = d->tile_docs;
    // ring geometry: NSTAGE chunks of STAGE_UNITS 16-byte units in flight per CTA.  Default 6 x 4 KB: with 12288-doc
    // tiles that is 72 KB per CTA -> 3 resident CTAs/SM, 72 KB of loads in flight per SM (r02 sweep at 20 M docs, B = 64:
    // 256x4 93.1 %, 256x6 94.8 %, 512x4 87.3 % of the measured HBM peak)
    const int SU = env_int("RR_BM25_STAGE_UNITS", 256, 32, 4096) / 32 * 32;
    const int NS = env_int("RR_BM25_STAGES", 6, 2, BM25_MAX_STAGES);
    const size_t smem = (size_t)T * sizeof(float) + (size_t)NS * SU * 16 + 2 * BM25_MAX_STAGES * sizeof(uint64_t);
    static RrSmemOptIn optin;
    int dev = 0;
    if (optin.needed(smem, &dev)) {
        RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // without this the driver sizes the shared-memory carveout for ONE resident CTA (ncu r01: occupancy limit 1)
        RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        optin.done(smem, dev);
    }
    // tiny batches may let the CTAs search their own rare-list bounds (one launch less), RR_BM25_PREPASS_MIN_B sets from which
    // batch size the pre-pass runs.  Default 1 = always: with the persistent kernel the producer warp resolves the bounds one
    // item ahead, and ~5 rare terms x 2 searches x 3 dependent round trips per item made it the bottleneck at B = 1
    static const int prepass_min_b = env_int("RR_BM25_PREPASS_MIN_B", 1, 1, 1 << 20);
    if (B < prepass_min_b) d_rtab = nullptr;
    if (d_rtab != nullptr) {
        const long long total = (long long)B * l_max * (d->n_tiles + 1);
        RrProfScope prof(RR_PROF_MISC, stream);
        bm25_rare_bounds_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const uint2*>(d->d_postings), d->d_tile_base, d->d_term_slot,
            reinterpret_cast<const unsigned long long*>(d->d_rare_off), d->vocab_size, T, d->n_tiles, d_terms, d_nterms, l_max, B,
            d_rtab);
        RR_LAUNCH_CHECK();
    }
    Bm25Args a;
    a.postings = reinterpret_cast<const uint4*>(d->d_postings); a.tile_base = d->d_tile_base; a.dir = d->d_dir;
    a.term_slot = d->d_term_slot; a.rtab = d_rtab; a.rare_off = reinterpret_cast<const unsigned long long*>(d->d_rare_off); a.V = d->vocab_size; a.T = T; a.n_freq = d->n_freq; a.n_tiles = d->n_tiles;
    a.n_docs = d->n_docs; a.q_terms = d_terms; a.q_len = d_nterms; a.l_max = l_max; a.out = d_out; a.ld_out = ld_out;
    a.stage_units = SU; a.n_stages = NS;
    if (l_max <= BM25_MAXL && d_counter != nullptr && !getenv("RR_BM25_ONE_SHOT")) {
        static RrSmemOptIn optin_p;
        if (optin_p.needed(smem, &dev)) {
            RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            RR_CUDA(cudaFuncSetAttribute(bm25_tile_scores_persistent_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            optin_p.done(smem, dev);
        }
        int per_sm = 0, sms = 148;
        RR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bm25_tile_scores_persistent_kernel, BM25_THREADS, smem));
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        per_sm = std::max(per_sm, 1);
        // items: (query, tile), query fastest, so that the CTAs sharing a tile (and the segments of common terms) run together
        const long long tiles_per_launch = std::max<long long>(1, std::min<long long>(d->n_tiles, 0x7fffffffll / std::max(B, 1)));
        for (long long t0 = 0; t0 < d->n_tiles; t0 += tiles_per_launch) {
            const int nt = (int)std::min<long long>(tiles_per_launch, d->n_tiles - t0);
            const int n_items = nt * B;
            a.tile0 = (int)t0;
            RrProfScope prof(RR_PROF_BM25_TILE, stream);      // d_counter = {next item, CTAs finished}: zero on entry, re-armed by the kernel
            bm25_tile_scores_persistent_kernel<<<(unsigned)std::min<long long>(n_items, (long long)per_sm * sms), BM25_THREADS, smem, stream>>>(
                a, n_items, B, d_counter);
            RR_LAUNCH_CHECK();
        }
        return RR_OK;
    }
    // grid: query fastest, so that the CTAs sharing a tile (and the segments of common terms) are co-resident
    for (int t0 = 0; t0 < d->n_tiles; t0 += 65535) {
        a.tile0 = t0;
        dim3 grid((unsigned)B, (unsigned)min(65535, d->n_tiles - t0));
        RrProfScope prof(RR_PROF_BM25_TILE, stream);
        bm25_tile_scores_kernel<<<grid, BM25_THREADS, smem, stream>>>(a);
        RR_LAUNCH_CHECK();
    }
    return RR_OK;
}

int rr_launch_bm25_candidates(const rr_index_desc* d, int V, const int32_t* d_terms, const int32_t* d_nterms,
                              int B, int l_max, const int64_t* d_cand, int pool, float* d_bm25, double* d_n_out,
                              double* d_avg_out, int64_t* d_grow_out, cudaStream_t stream, int pack_bg,
                              int64_t pack_stride, const float* d_dense_in, float* d_dense_out,
                              const int32_t* d_uncertified) {
    const int64_t total = (int64_t)B * pool;
    if (total <= 0) return RR_OK;
    const int threads = 256;
    RrProfScope prof(RR_PROF_BM25_CAND, stream);
    // small batches (the latency path): one warp per candidate over the forward index; RR_BM25_CAND_WARP=0 disables
    const char* warp_env = getenv("RR_BM25_CAND_WARP");
    const bool warp_ok = !(warp_env && atoi(warp_env) == 0);
    if (warp_ok && total <= 8192 && d->d_fwd_off != nullptr && d->d_fwd_data != nullptr) {
        const unsigned blocks_w = (unsigned)((total * 32 + threads - 1) / threads);
        bm25_candidates_warp_kernel<<<blocks_w, threads, 0, stream>>>(
            reinterpret_cast<const unsigned long long*>(d->d_fwd_off), reinterpret_cast<const uint2*>(d->d_fwd_data), V,
            (long long)d->n_docs, d_terms, d_nterms, l_max, reinterpret_cast<const long long*>(d_cand), pool, B,
            d->d_n_reviews, d->d_avg_stars, (long long)d->row_offset, d_bm25, d_n_out, d_avg_out,
            reinterpret_cast<long long*>(d_grow_out), pack_bg, (long long)pack_stride, d_dense_in, d_dense_out, d_uncertified);
        RR_LAUNCH_CHECK();
        return RR_OK;
    }
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    bm25_candidates_kernel<<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const uint2*>(d->d_postings), d->d_tile_base, d->d_dir, d->d_term_slot,
        reinterpret_cast<const unsigned long long*>(d->d_rare_off), d->n_freq, d->n_tiles,
        reinterpret_cast<const unsigned long long*>(d->d_fwd_off), reinterpret_cast<const uint2*>(d->d_fwd_data), V,
        d->tile_docs > 0 ? d->tile_docs : 1, (long long)d->n_docs, d_terms, d_nterms,
        l_max, reinterpret_cast<const long long*>(d_cand), pool, B, d->d_n_reviews, d->d_avg_stars, (long long)d->row_offset, d_bm25,
        d_n_out, d_avg_out, reinterpret_cast<long long*>(d_grow_out), pack_bg, (long long)pack_stride, d_dense_in,
        d_dense_out, d_uncertified);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
