// K2 -- tensor-core shortlist path of rr_dense_topk (sm_100a: tcgen05.mma + TMEM + TMA).
//
// Stands behind the `sims = mat @ q` + top-pool of cosine_similarity_search (utils.py:111-124,
// = _cosine_pool app/app_product_search.py:192-195, cosine_search app/test.py:125-132) for a
// BATCH of queries, where the contraction is a genuine GEMM (B x N x D).
//
// Pipeline per batch
//   cvt_queries      fp32 queries -> bf16 [B_pad, dim_pad] (+ ||q||)
//   tc_filter_kernel bf16 x bf16 -> fp32 scores on the tensor cores, never materialised:
//                    CTA = 128 queries (UMMA M, one TMEM lane per query) x a strided set of
//                    256-doc tiles (UMMA N).  The 128 x D query tile is loaded once by TMA and
//                    stays in shared memory; 256 x 64 corpus k-blocks stream through a 4-stage
//                    TMA/mbarrier ring; one thread issues tcgen05.mma (K=16) into a double-
//                    buffered TMEM accumulator (2 x 256 columns); eight epilogue warps read the
//                    accumulator with tcgen05.ld (thread = query row) and keep only scores
//                    >= the query's running threshold tau, appended as 64-bit (score, doc) keys.
//   tc_select_kernel per query: merge kept list + new candidates, keep the best k', raise tau to
//                    the k'-th score.  The corpus is scanned in geometrically growing segments
//                    (filter launch + select launch per segment) so that the expected number of
//                    candidates per segment stays ~ (growth-1) * k'.
//   rescore (K3)     exact canonical fp32 similarity of the k' shortlisted rows (dense_exact.cu)
//   tc_finalize      order by (exact desc, row asc), emit the top `pool`, and CERTIFY: every row
//                    outside the shortlist has bf16 score <= c (the k'-th kept), hence exact score
//                    <= c + eps; the result is exact if the pool-th exact score > c + eps.
//                    Uncertified queries (near-ties denser than the margin, candidate overflow)
//                    are recomputed by the exact fp32 path -- results never depend on bf16.
//
// eps = (2^-7 + 2^-15) * ||q|| * max||row||  (bf16 round-to-nearest of both operands, unit roundoff
// 2^-8 each) + 1e-4 * ||q|| * max||row|| (fp32 accumulation slack for D <= 1024).
#include "rr_internal.h"
#include "rr_kernels.h"
#include "dense_tc.h"

#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

constexpr int TC_BM = 128;      // queries per CTA tile (UMMA M)
constexpr int TC_BN = 256;      // docs per tile (UMMA N)
constexpr int TC_BK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_MAX_KB = 6;    // dim_pad <= 384 keeps the query tile resident in shared memory ...
constexpr int TC_MAX_KB_STREAMED = 32;   // ... wider rows (<= 2048) stream the query k-blocks next to the corpus k-blocks
constexpr int TC_THREADS = 384;     // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue (2 x 4 warps)
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_Q_KB_BYTES = TC_BM * TC_BK * 2;    // 16 KB
constexpr int TC_B_KB_BYTES = TC_BN * TC_BK * 2;    // 32 KB
constexpr int TC_SMEM_Q = TC_MAX_KB * TC_Q_KB_BYTES;
constexpr int TC_SMEM_B = TC_STAGES * TC_B_KB_BYTES;
constexpr int TC_SMEM_BAR = 256;
constexpr int TC_SMEM_TOTAL = TC_SMEM_Q + TC_SMEM_B + TC_SMEM_BAR + 1024;   // + alignment slack
constexpr int TC_CAP_TOTAL = 16384;  // candidate slots per query per segment, split evenly over the CTAs
                                     // that scan for that query (each owns a private sub-list: no atomics)
constexpr int TC_SORT_MAX = 16384;   // largest per-query selection in tc_select_kernel (128 KB of keys)
constexpr int TC_GROWTH = 4;         // corpus prefix grows x4 per segment (x2 for very long shortlists)

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc_512(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M=128 N=256 K=16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFFu);   // start address
    d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}
constexpr uint32_t kInstrDesc = (1u << 4)      // D format: F32
                                | (1u << 7)    // A format: BF16
                                | (1u << 10)   // B format: BF16
                                | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);   // K-major A and B

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// predicated 8-byte global store (no branch): used by the epilogue's candidate appends
__device__ __forceinline__ void st_pred_v2(unsigned long long* p, uint32_t lo, uint32_t hi, bool pred) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %3, 0;\n\t"
        "@p st.global.v2.u32 [%0], {%1, %2};\n\t"
        "}" ::"l"(p), "r"(lo), "r"(hi), "r"((uint32_t)pred)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// query conversion
// ---------------------------------------------------------------------------------------------
__global__ void cvt_queries_kernel(const float* __restrict__ q, int B, int D, int dim_pad, int B_pad,
                                   __nv_bfloat16* __restrict__ out, float* __restrict__ qnorm) {
    const int row = blockIdx.x;
    float ss = 0.f;
    for (int d = threadIdx.x; d < dim_pad; d += blockDim.x) {
        float x = (row < B && d < D) ? q[(long long)row * D + d] : 0.f;
        out[(long long)row * dim_pad + d] = __float2bfloat16_rn(x);
        ss += x * x;
    }
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0 && row < B) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        qnorm[row] = sqrtf(t);
    }
}

// ---------------------------------------------------------------------------------------------
// the GEMM + threshold filter
// ---------------------------------------------------------------------------------------------
struct TcFilterArgs {
    long long n_docs;
    int n_kb;          // dim_pad / 64
    int B;
    int n_parts;                 // query parts covered by this launch (<= 8)
    int part_qt0[8], part_nq[8]; // first query tile / query tiles of every part
    int part_reps[8];            // CTAs per query tile in the part
    int part_ctas[8];            // = part_nq * part_reps
    int dt_lo, dt_hi;  // doc tiles of this segment
    int q_resident;    // 1: query tile loaded once (dim_pad <= 384); 0: its k-blocks are streamed with the corpus'
    const float* tau;  // [B]
    int pair_stages;   // tc_filter_pair_kernel: depth of the corpus ring
    int n_sub, cap_sub;              // sub-lists per query (2 per scanning CTA) and slots per sub-list
    unsigned long long* cand_keys;   // [B, n_sub, cap_sub]
    unsigned* cand_cnt;              // [B, n_sub]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                 const TcFilterArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // resident: [Q: 6 x 16 KB][B ring: 4 x 32 KB]      streamed: [ring: 4 x (A 16 KB | B 32 KB)]
    uint8_t* sQ = smem;
    uint8_t* sB = smem + TC_SMEM_Q;
    constexpr int kStreamStage = TC_Q_KB_BYTES + TC_B_KB_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_SMEM_Q + TC_SMEM_B);
    uint64_t* full = bars;                  // [TC_STAGES]  TMA -> MMA
    uint64_t* empty = bars + TC_STAGES;     // [TC_STAGES]  MMA -> TMA
    uint64_t* tfull = bars + 2 * TC_STAGES; // [2]          MMA -> epilogue
    uint64_t* tempty = tfull + 2;           // [2]          epilogue -> MMA
    uint64_t* qfull = tempty + 2;           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // one launch covers all query parts: the CTAs of part p (its nq query tiles x its reps) follow those of part p-1
    int cta = blockIdx.x, part = 0;
    while (part + 1 < a.n_parts && cta >= a.part_ctas[part]) { cta -= a.part_ctas[part]; ++part; }
    const int n_qt_p = a.part_nq[part], reps_p = a.part_reps[part];
    const int qt = a.part_qt0[part] + cta % n_qt_p;
    const int j0 = cta / n_qt_p;
    const int span = a.dt_hi - a.dt_lo;
    const int n_tiles = j0 < span ? (span - j0 + reps_p - 1) / reps_p : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_c);
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TC_EPI_THREADS); }
        mbar_init(qfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_512(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (one thread) =====
        if (lane == 0 && n_tiles > 0) {
            if (a.q_resident) {
                mbar_expect_tx(qfull, (uint32_t)(a.n_kb * TC_Q_KB_BYTES));
                for (int kb = 0; kb < a.n_kb; ++kb) tma_load_2d(&tmap_q, qfull, sQ + kb * TC_Q_KB_BYTES, kb * TC_BK, qt * TC_BM);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int dt = a.dt_lo + j0 + t * reps_p;
                for (int kb = 0; kb < a.n_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    if (a.q_resident) {
                        mbar_expect_tx(&full[stage], (uint32_t)TC_B_KB_BYTES);
                        tma_load_2d(&tmap_c, &full[stage], sB + stage * TC_B_KB_BYTES, kb * TC_BK, dt * TC_BN);
                    } else {
                        uint8_t* st = smem + stage * kStreamStage;
                        mbar_expect_tx(&full[stage], (uint32_t)kStreamStage);
                        tma_load_2d(&tmap_q, &full[stage], st, kb * TC_BK, qt * TC_BM);
                        tma_load_2d(&tmap_c, &full[stage], st + TC_Q_KB_BYTES, kb * TC_BK, dt * TC_BN);
                    }
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && n_tiles > 0) {
            if (a.q_resident) {
                mbar_wait(qfull, 0u);
                tc_fence_after();
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int as = t & 1;
                const uint32_t aphase = (uint32_t)((t >> 1) & 1);
                mbar_wait(&tempty[as], aphase ^ 1u);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TC_BN);
                for (int kb = 0; kb < a.n_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint8_t* pa = a.q_resident ? sQ + kb * TC_Q_KB_BYTES : smem + stage * kStreamStage;
                    const uint8_t* pb = a.q_resident ? sB + stage * TC_B_KB_BYTES : smem + stage * kStreamStage + TC_Q_KB_BYTES;
                    const uint64_t da = umma_desc_sw128(pa);
                    const uint64_t db = umma_desc_sw128(pb);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)       // +32 bytes (>>4 = 2) per K=16 step
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kInstrDesc, (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[stage]);               // smem slot reusable once these MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[as]);                       // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = (query row, column half), threshold filter =====
        // Warps 4..7 read accumulator columns 0..127, warps 8..11 columns 128..255 (a warp may only touch
        // TMEM lanes 32*(warp%4)..+31).  Every (CTA, query, half) owns a private candidate sub-list, so an
        // append is a plain store with a register counter: no atomics, nothing to wait for.
        const int ew = (warp - 4) & 3;
        const int half = (warp - 4) >> 2;
        const int q = qt * TC_BM + ew * 32 + lane;
        const bool active = q < a.B;
        const float tau = active ? a.tau[q] : INFINITY;
        const int sub = j0 * 2 + half;
        unsigned long long* my_keys = a.cand_keys + ((size_t)(active ? q : 0) * a.n_sub + sub) * a.cap_sub;
        unsigned cnt = 0;
        const unsigned cap = (unsigned)a.cap_sub;
        for (int t = 0; t < n_tiles; ++t) {
            const int as = t & 1;
            const uint32_t aphase = (uint32_t)((t >> 1) & 1);
            const int dt = a.dt_lo + j0 + t * reps_p;
            const long long doc_base = (long long)dt * TC_BN;
            const uint32_t doc_base_u = (uint32_t)doc_base;
            const int n_valid = (int)min((long long)TC_BN, a.n_docs - doc_base);   // rows past the corpus end are zero-filled
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * TC_BN + half * (TC_BN / 2));
#pragma unroll 1
            for (int chunk = 0; chunk < TC_BN / 64; chunk += 2) {
                uint32_t v[2][32];
                tmem_ld_x32(taddr0 + (uint32_t)(chunk * 32), v[0]);
                tmem_ld_x32(taddr0 + (uint32_t)(chunk * 32 + 32), v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // maxima of the four groups of 8, then descend only into groups that can hold a pass
                    float m8[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m = __uint_as_float(v[h][8 * g]);
#pragma unroll
                        for (int e = 1; e < 8; ++e) m = fmaxf(m, __uint_as_float(v[h][8 * g + e]));
                        m8[g] = m;
                    }
                    const float m32 = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
                    if (active && m32 >= tau) {
                        const int c0 = half * (TC_BN / 2) + (chunk + h) * 32;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (m8[g] >= tau) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    // branch-free: predicated 8-byte store of (score bits, doc), counter += pass
                                    const uint32_t sb = v[h][8 * g + e];
                                    const int col = c0 + 8 * g + e;
                                    const bool pass = __uint_as_float(sb) >= tau && col < n_valid;
                                    st_pred_v2(my_keys + cnt, sb, doc_base_u + (uint32_t)col, pass && cnt < cap);
                                    cnt += pass ? 1u : 0u;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[as]);
        }
        if (active) a.cand_cnt[(size_t)q * a.n_sub + sub] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// the same GEMM + filter on CTA PAIRS (tcgen05.mma.cta_group::2)
// ---------------------------------------------------------------------------------------------
// One-CTA tiles are co-limited by the SM's ingress from L2: a 256-doc x 384-d corpus tile is 196 KB per 3072 tensor
// clocks = the 64 B/clk an SM can take in, which is why the one-CTA kernel tops out at ~80-89 % tensor activity and
// loses efficiency as the clock rises (r02: 60 % of the tensor peak at 1965 MHz on a 1.25 M-row shard).  A CTA pair
// (two SMs of one TPC, cluster of 2) computes a 256-query x 256-doc tile per instruction: each CTA keeps ITS 128
// queries resident and streams only HALF of every corpus tile (128 docs); the pair's MMA reads both halves.  Per-SM
// ingress halves, the accumulator layout per CTA (128 lanes x 256 columns) and hence the whole epilogue are unchanged.
//   - both CTAs issue TMA for their half (cp.async.bulk.tensor .cta_group::2, completion counted on the LEADER's
//     mbarrier), the leader's single MMA thread issues tcgen05.mma.cta_group::2 and commits with a multicast arrive
//     to both CTAs' "stage empty" / "accumulator full" barriers; both CTAs' epilogue threads arrive on the leader's
//     "accumulator empty" barrier.
constexpr int TC2_MAX_STAGES = 8;
constexpr int TC2_B_KB_BYTES = (TC_BN / 2) * TC_BK * 2;      // 16 KB: this CTA's half of a corpus k-block
constexpr int tc2_smem_total(int stages) { return TC_SMEM_Q + 1024 + stages * TC2_B_KB_BYTES + 1024; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    // default semantics (.release at CTA scope), like a local arrive: r02 ncu showed `.release.cluster` fencing every
    // epilogue thread's candidate stores to global memory per tile (ERRBAR + arrive = the hottest epilogue lines), which made
    // the epilogue slower than the MMA.  What the MMA thread needs ordered is the TMEM read, which tcgen05.wait::ld +
    // tcgen05.fence::before_thread_sync have already completed.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_512_pair(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512_pair(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
// commit of the pair's MMAs: arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D[tmem] (+)= A * B^T over the pair: M = 256 (128 rows per CTA), N = 256 (128 docs per CTA), K = 16
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
constexpr uint32_t kInstrDescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                    ((uint32_t)((2 * TC_BM) >> 4) << 24);

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_filter_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c_half,
                      const TcFilterArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                  // [6 x 16 KB] this CTA's 128 queries, resident
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_SMEM_Q);      // 1 KB of barriers
    uint8_t* sB = smem + TC_SMEM_Q + 1024;               // [stages x 16 KB] this CTA's half of the corpus k-blocks
    const int TC2_STAGES = a.pair_stages;
    uint64_t* full = bars;                    // [TC2_STAGES]  used in the LEADER: both CTAs' TMA -> MMA
    uint64_t* empty = bars + 8;               // [TC2_STAGES]  per CTA: MMA (multicast commit) -> this CTA's TMA
    uint64_t* tfull = bars + 16;              // [2]           per CTA: MMA (multicast commit) -> this CTA's epilogue
    uint64_t* tempty = tfull + 2;             // [2]           used in the LEADER: both CTAs' epilogues -> MMA
    uint64_t* qfull = tempty + 2;             // [1]           used in the LEADER
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    uint64_t* qfree = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // [1] per CTA: MMA (multicast commit) -> this CTA's TMA: sQ reusable
    // PERSISTENT OVER THE QUERY PARTS: the grid has max_p(part_ctas[p]) CTAs and CTA c serves, one part after the other,
    // query tile part_qt0[p] + c % nq_p with doc tiles j0 = c / nq_p, j0 + reps_p, ...  (pairs (2c, 2c+1) share j0 and own
    // adjacent query tiles: every part has an even number of query tiles).  All CTAs of the launch are resident from the
    // start -- no second wave waiting for shared memory -- which is what lets the block scheduler place the small tail
    // kernels of another batch in flight next to them.
    const int cta = blockIdx.x;
    const int span = a.dt_hi - a.dt_lo;
    auto part_geometry = [&](int p, int& qt, int& j0, int& reps_p) -> int {      // -> doc tiles of this CTA in part p
        if (cta >= a.part_ctas[p]) return 0;
        reps_p = a.part_reps[p];
        qt = a.part_qt0[p] + cta % a.part_nq[p];
        j0 = cta / a.part_nq[p];
        return j0 < span ? (span - j0 + reps_p - 1) / reps_p : 0;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_c_half);
        for (int i = 0; i < TC2_STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        // accumulator-empty: one arrive per epilogue WARP of both CTAs (r02: one remote arrive per thread -- 512 per tile on one
        // barrier word across the cluster -- made the pair kernel 1.6x slower than the one-CTA kernel)
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * (TC_EPI_THREADS / 32)); }
        mbar_init(qfull, 2);
        mbar_init(qfree, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_512_pair(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of BOTH CTAs are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (one thread per CTA): per part its queries, then its half of every corpus k-block =====
        if (lane == 0) {
            const uint32_t qfull_leader = mapa_cluster(smem_u32(qfull), 0);
            int stage = 0, served = 0;
            uint32_t phase = 0;
            for (int p = 0; p < a.n_parts; ++p) {
                int qt = 0, j0 = 0, reps_p = 1;
                const int n_tiles = part_geometry(p, qt, j0, reps_p);
                if (n_tiles == 0) continue;
                if (served > 0) mbar_wait(qfree, (uint32_t)((served - 1) & 1));     // the previous part's MMAs are done with sQ
                if (leader) mbar_expect_tx(qfull, (uint32_t)(2 * a.n_kb * TC_Q_KB_BYTES));
                else mbar_arrive_remote(qfull_leader);
                for (int kb = 0; kb < a.n_kb; ++kb)
                    tma_load_2d_pair(&tmap_q, qfull_leader, sQ + kb * TC_Q_KB_BYTES, kb * TC_BK, qt * TC_BM);
                ++served;
                for (int t = 0; t < n_tiles; ++t) {
                    const int dt = a.dt_lo + j0 + t * reps_p;
                    for (int kb = 0; kb < a.n_kb; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        const uint32_t full_leader = mapa_cluster(smem_u32(&full[stage]), 0);
                        if (leader) mbar_expect_tx(&full[stage], (uint32_t)(2 * TC2_B_KB_BYTES));
                        else mbar_arrive_remote(full_leader);
                        tma_load_2d_pair(&tmap_c_half, full_leader, sB + stage * TC2_B_KB_BYTES, kb * TC_BK,
                                         dt * TC_BN + (int)rank * (TC_BN / 2));
                        if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER CTA issues for the pair =====
        if (leader && lane == 0) {
            int stage = 0, served = 0, tt = 0;
            uint32_t phase = 0;
            for (int p = 0; p < a.n_parts; ++p) {
                int qt = 0, j0 = 0, reps_p = 1;
                const int n_tiles = part_geometry(p, qt, j0, reps_p);
                if (n_tiles == 0) continue;
                mbar_wait(qfull, (uint32_t)(served & 1));
                tc_fence_after();
                ++served;
                for (int t = 0; t < n_tiles; ++t, ++tt) {
                    const int as = tt & 1;
                    const uint32_t aphase = (uint32_t)((tt >> 1) & 1);
                    mbar_wait(&tempty[as], aphase ^ 1u);       // both epilogues have drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * TC_BN);
                    for (int kb = 0; kb < a.n_kb; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint64_t da = umma_desc_sw128(sQ + kb * TC_Q_KB_BYTES);
                        const uint64_t db = umma_desc_sw128(sB + stage * TC2_B_KB_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            umma_bf16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kInstrDescPair, (uint32_t)((kb | k) != 0));
                        umma_commit_pair(&empty[stage]);
                        if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit_pair(&tfull[as]);
                }
                umma_commit_pair(qfree);                      // every MMA that reads this part's queries has retired
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): identical to tc_filter_kernel, the drain signal goes to the leader =====
        const int ew = (warp - 4) & 3;
        const int half = (warp - 4) >> 2;
        const unsigned cap = (unsigned)a.cap_sub;
        const uint32_t tempty_leader0 = mapa_cluster(smem_u32(&tempty[0]), 0);
        int tt = 0;
        for (int p = 0; p < a.n_parts; ++p) {
            int qt = 0, j0 = 0, reps_p = 1;
            const int n_tiles = part_geometry(p, qt, j0, reps_p);
            if (n_tiles == 0) continue;
            const int q = qt * TC_BM + ew * 32 + lane;
            const bool active = q < a.B;
            const float tau = active ? a.tau[q] : INFINITY;
            const int sub = j0 * 2 + half;
            unsigned long long* my_keys = a.cand_keys + ((size_t)(active ? q : 0) * a.n_sub + sub) * a.cap_sub;
            unsigned cnt = 0;
            for (int t = 0; t < n_tiles; ++t, ++tt) {
                const int as = tt & 1;
                const uint32_t aphase = (uint32_t)((tt >> 1) & 1);
                const int dt = a.dt_lo + j0 + t * reps_p;
                const long long doc_base = (long long)dt * TC_BN;
                const uint32_t doc_base_u = (uint32_t)doc_base;
                const int n_valid = (int)min((long long)TC_BN, a.n_docs - doc_base);
                mbar_wait(&tfull[as], aphase);
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * TC_BN + half * (TC_BN / 2));
#pragma unroll 1
                for (int chunk = 0; chunk < TC_BN / 64; chunk += 2) {
                    uint32_t v[2][32];
                    tmem_ld_x32(taddr0 + (uint32_t)(chunk * 32), v[0]);
                    tmem_ld_x32(taddr0 + (uint32_t)(chunk * 32 + 32), v[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float m8[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float m = __uint_as_float(v[h][8 * g]);
#pragma unroll
                            for (int e = 1; e < 8; ++e) m = fmaxf(m, __uint_as_float(v[h][8 * g + e]));
                            m8[g] = m;
                        }
                        const float m32 = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
                        if (active && m32 >= tau) {
                            const int c0 = half * (TC_BN / 2) + (chunk + h) * 32;
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                if (m8[g] >= tau) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const uint32_t sb = v[h][8 * g + e];
                                        const int col = c0 + 8 * g + e;
                                        const bool pass = __uint_as_float(sb) >= tau && col < n_valid;
                                        st_pred_v2(my_keys + cnt, sb, doc_base_u + (uint32_t)col, pass && cnt < cap);
                                        cnt += pass ? 1u : 0u;
                                    }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(tempty_leader0 + (uint32_t)(as * 8));
            }
            if (active) a.cand_cnt[(size_t)q * a.n_sub + sub] = cnt;
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // both CTAs are done with TMEM and with each other's barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_512_pair(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// per-query selection
// ---------------------------------------------------------------------------------------------
__device__ void bitonic_keys_desc(unsigned long long* key, int n_pad) {
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n_pad / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x = key[lo], y = key[hi];
                if ((x < y) == desc) { key[lo] = y; key[hi] = x; }
            }
            __syncthreads();
        }
    }
}


// Gather of one query's keys into shared memory with GU loads in flight per thread: positions [0, kc) come from
// the kept list, positions [kc, total) from the candidate sub-lists (sub-list r by binary search over s_off).
// The addresses of a batch are computed first, then all loads are issued, then the keys are converted and stored
// (r02 profile: the one-load-per-iteration loop left the selection waiting on a chain of global-load latencies).
template <int GU>
__device__ __forceinline__ void gather_query_keys(unsigned long long* sk, const unsigned long long* kept,
                                                  int kc, const unsigned long long* cand, const int* s_off,
                                                  int n_sub, int cap_sub, int total, int t, int stride) {
    for (int p0 = t; p0 < total; p0 += stride * GU) {
        const unsigned long long* src[GU];
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const int p = p0 + u * stride;
            src[u] = nullptr;
            if (p < kc) {
                src[u] = kept + p;
            } else if (p < total) {
                int lo_r = 0, hi_r = n_sub;                 // largest r with s_off[r] <= p
                while (hi_r - lo_r > 1) { const int mid = (lo_r + hi_r) >> 1; if (s_off[mid] <= p) lo_r = mid; else hi_r = mid; }
                src[u] = cand + (size_t)lo_r * cap_sub + (p - s_off[lo_r]);
            }
        }
        unsigned long long raw[GU];
#pragma unroll
        for (int u = 0; u < GU; ++u) raw[u] = src[u] ? *src[u] : 0ull;     // plain loads: the kept list is rewritten below
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const int p = p0 + u * stride;
            if (p < kc) {
                sk[p] = raw[u];
            } else if (p < total) {
                const unsigned long long k = rr_make_key(__uint_as_float((uint32_t)raw[u]), (uint32_t)(raw[u] >> 32));   // {score bits, doc}
                sk[p] = k ? k : 1ull;
            }
        }
    }
}

// Per query: pool = kept list + the candidate sub-lists of this segment.  Keeps the KP largest 64-bit keys
// (unordered) and raises tau to the KP-th score.  Selection is an MSB-first radix select over the key
// bits below the common prefix of (min, max): 8-bit digits, warp-aggregated shared-memory histogram,
// early exit as soon as the pivot bucket is consumed exactly.  O(n) per pass instead of a full sort.
constexpr int TC_MAX_SUB = 320;

__global__ void __launch_bounds__(256)
tc_select_kernel(const unsigned long long* __restrict__ cand_keys, unsigned* __restrict__ cand_cnt, int n_sub,
                 int cap_sub, unsigned long long* __restrict__ kept_keys, int* __restrict__ kept_cnt, int KP,
                 float* __restrict__ tau, int* __restrict__ overflow, int final_pass,
                 long long* __restrict__ rows_out, int sort_cap) {
    extern __shared__ unsigned long long sk[];      // sort_cap keys
    __shared__ int s_off[TC_MAX_SUB + 1];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_red[16];
    __shared__ unsigned long long s_prefix, s_minsel;
    __shared__ int s_bits, s_krem, s_done, s_over, s_out;
    __shared__ int s_wsum[8], s_wsum1[8];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kc = kept_cnt[q];

    // sub-list sizes -> offsets (block scan over n_sub <= 320 entries, up to two per thread)
    int c0 = 0, c1 = 0, over = 0;
    if (tid < n_sub) {
        const unsigned raw = cand_cnt[(size_t)q * n_sub + tid];
        c0 = (int)min(raw, (unsigned)cap_sub);
        over |= raw > (unsigned)cap_sub;
    }
    if (tid + 256 < n_sub) {
        const unsigned raw = cand_cnt[(size_t)q * n_sub + tid + 256];
        c1 = (int)min(raw, (unsigned)cap_sub);
        over |= raw > (unsigned)cap_sub;
    }
    int x0 = c0, x1 = c1;
    for (int o = 1; o < 32; o <<= 1) {
        const int y0 = __shfl_up_sync(0xffffffffu, x0, o);
        const int y1 = __shfl_up_sync(0xffffffffu, x1, o);
        if (lane >= o) { x0 += y0; x1 += y1; }
    }
    if (lane == 31) { s_wsum[wid] = x0; s_wsum1[wid] = x1; }
    if (tid == 0) { s_over = 0; s_done = 0; s_out = 0; s_minsel = ~0ull; }
    __syncthreads();
    int wbase0 = 0, wbase1 = 0, tot0 = 0, tot1 = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < wid) { wbase0 += s_wsum[w]; wbase1 += s_wsum1[w]; }
        tot0 += s_wsum[w];
        tot1 += s_wsum1[w];
    }
    if (tid < n_sub) s_off[tid] = kc + wbase0 + x0 - c0;
    if (tid + 256 < n_sub) s_off[tid + 256] = kc + tot0 + wbase1 + x1 - c1;
    const int total_all = kc + tot0 + tot1;
    if (tid == 0) s_off[n_sub] = total_all;
    if (over || (tid == 0 && total_all > sort_cap)) s_over = 1;
    __syncthreads();
    const int total = min(total_all, sort_cap);

    // gather into shared memory
    gather_query_keys<8>(sk, kept_keys + (size_t)q * KP, kc, cand_keys + (size_t)q * n_sub * cap_sub, s_off, n_sub, cap_sub,
                         total, tid, 256);
    __syncthreads();

    int newk = total;
    unsigned long long thr_prefix = 0ull;
    int thr_bits = 0;                      // select keys with (key >> (64-thr_bits)) >= thr_prefix; 0 = all
    if (total > KP) {
        // common prefix of all keys
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = tid; i < total; i += 256) { const unsigned long long k = sk[i]; kmin = min(kmin, k); kmax = max(kmax, k); }
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (lane == 0) { s_red[wid] = kmin; s_red[8 + wid] = kmax; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) { kmin = min(kmin, s_red[w]); kmax = max(kmax, s_red[8 + w]); }
            const int common = kmin == kmax ? 56 : __clzll((long long)(kmin ^ kmax));
            const int bits = min(common, 56);
            s_bits = bits;
            s_prefix = bits ? (kmax >> (64 - bits)) : 0ull;
            s_krem = KP;
        }
        __syncthreads();
        for (int pass = 0; pass < 9; ++pass) {
            const int bits = s_bits;
            const unsigned long long prefix = s_prefix;
            const int dig = min(8, 64 - bits);
            if (dig <= 0 || s_done) break;
            s_hist[tid] = 0u;
            __syncthreads();
            const int shift = 64 - bits - dig;
            for (int i0 = 0; i0 < total; i0 += 256) {
                const int i = i0 + tid;
                bool in = false;
                unsigned b = 0;
                if (i < total) {
                    const unsigned long long k = sk[i];
                    in = bits == 0 || (k >> (64 - bits)) == prefix;
                    b = (unsigned)((k >> shift) & ((1u << dig) - 1u));
                }
                if (in) atomicAdd(&s_hist[b], 1u);
            }
            __syncthreads();
            if (wid == 0) {
                // walk buckets from the top: lane l owns buckets 255-8l .. 248-8l
                unsigned loc[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { loc[j] = s_hist[255 - (lane * 8 + j)]; sum += loc[j]; }
                unsigned incl = sum;
                for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                const unsigned excl = incl - sum;
                const unsigned krem = (unsigned)s_krem;
                const bool mine = excl < krem && krem <= incl;
                if (mine) {
                    unsigned above = excl;
                    int j = 0;
                    for (; j < 7; ++j) { if (above + loc[j] >= krem) break; above += loc[j]; }
                    const int bucket = 255 - (lane * 8 + j);
                    s_krem = (int)(krem - above);
                    s_prefix = (prefix << dig) | (unsigned long long)bucket;
                    s_bits = bits + dig;
                    if (loc[j] == krem - above || bits + dig >= 64) s_done = 1;
                }
            }
            __syncthreads();
        }
        thr_prefix = s_prefix;
        thr_bits = s_bits;
        newk = KP;
    }
    // compact the selected keys to the kept list (unordered) and find the smallest selected key
    unsigned long long mymin = ~0ull;
    for (int i0 = 0; i0 < total; i0 += 256) {
        const int i = i0 + tid;
        bool sel = false;
        unsigned long long k = 0ull;
        if (i < total) { k = sk[i]; sel = thr_bits == 0 || (k >> (64 - thr_bits)) >= thr_prefix; }
        const unsigned m = __ballot_sync(0xffffffffu, sel);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_out, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sel) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < KP) {
                kept_keys[(size_t)q * KP + slot] = k;
                if (final_pass) rows_out[(size_t)q * KP + slot] = (long long)rr_key_index(k);
            }
            mymin = min(mymin, k);
        }
    }
    for (int o = 16; o > 0; o >>= 1) mymin = min(mymin, __shfl_xor_sync(0xffffffffu, mymin, o));
    if (lane == 0) atomicMin(&s_minsel, mymin);
    if (final_pass)
        for (int i = newk + tid; i < KP; i += 256) rows_out[(size_t)q * KP + i] = -1ll;
    for (int r = tid; r < n_sub; r += 256) cand_cnt[(size_t)q * n_sub + r] = 0u;
    __syncthreads();
    if (tid == 0) {
        kept_cnt[q] = newk;
        tau[q] = newk == KP ? rr_key_score(s_minsel) : -INFINITY;
        if (s_over) overflow[q] = 1;
    }
}

// Warp-per-query variant of the selection (used when a query's keys fit `cap` <= 2048 slots): no block
// barriers, 4..8 queries per CTA.  Same algorithm as tc_select_kernel: gather kept + candidate sub-lists into
// shared memory, MSB-first radix select below the common prefix, compact the k' survivors, tau = k'-th score.
__global__ void __launch_bounds__(256)
tc_select_warp_kernel(const unsigned long long* __restrict__ cand_keys, unsigned* __restrict__ cand_cnt, int n_sub,
                      int cap_sub, unsigned long long* __restrict__ kept_keys, int* __restrict__ kept_cnt, int KP,
                      float* __restrict__ tau, int* __restrict__ overflow, int final_pass,
                      long long* __restrict__ rows_out, int cap, int warps_per_cta, int B, int prune_m,
                      const float* __restrict__ qnorm, float eps_rel) {
    extern __shared__ unsigned long long smem_sel[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = blockIdx.x * warps_per_cta + wid;
    if (wid >= warps_per_cta || q >= B) return;
    // per-warp carve: keys[cap] | hist[256] (u32) | off[TC_MAX_SUB+1] (int)
    const size_t per_warp_words = (size_t)cap + 128 + (TC_MAX_SUB + 2) / 2 + 1;
    unsigned long long* sk = smem_sel + (size_t)wid * per_warp_words;
    unsigned* hist = reinterpret_cast<unsigned*>(sk + cap);
    int* s_off = reinterpret_cast<int*>(sk + cap + 128);
    const unsigned FULL = 0xffffffffu;

    const int kc = kept_cnt[q];
    // ---- sub-list sizes -> offsets ------------------------------------------------------------------
    int run = kc, over = 0;
    for (int r0 = 0; r0 < n_sub; r0 += 32) {
        const int r = r0 + lane;
        int c = 0;
        if (r < n_sub) {
            const unsigned raw = cand_cnt[(size_t)q * n_sub + r];
            c = (int)min(raw, (unsigned)cap_sub);
            over |= raw > (unsigned)cap_sub;
        }
        int x = c;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
        if (r < n_sub) s_off[r] = run + x - c;
        run += __shfl_sync(FULL, x, 31);
    }
    const int total_all = run;
    if (lane == 0) s_off[n_sub] = total_all;
    over = __any_sync(FULL, over) || total_all > cap;
    const int total = min(total_all, cap);
    __syncwarp();

    // ---- gather ----------------------------------------------------------------------------------
    gather_query_keys<8>(sk, kept_keys + (size_t)q * KP, kc, cand_keys + (size_t)q * n_sub * cap_sub, s_off, n_sub, cap_sub,
                         total, lane, 32);
    __syncwarp();

    // MSB-first radix select of the `want` largest keys of sk[0, total): returns (prefix, bits) such that the selection is
    // {k : (k >> (64 - bits)) >= prefix}; bits = 0 selects everything
    auto select_top = [&](int want, unsigned long long& out_prefix, int& out_bits) {
        out_prefix = 0ull;
        out_bits = 0;
        if (total <= want) return;
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = lane; i < total; i += 32) { const unsigned long long k = sk[i]; kmin = min(kmin, k); kmax = max(kmax, k); }
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(FULL, kmax, o));
        }
        int bits = kmin == kmax ? 56 : min(__clzll((long long)(kmin ^ kmax)), 56);
        unsigned long long prefix = bits ? (kmax >> (64 - bits)) : 0ull;
        int krem = want;
        bool done = false;
        for (int pass = 0; pass < 9 && !done; ++pass) {
            const int dig = min(8, 64 - bits);
            if (dig <= 0) break;
#pragma unroll
            for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0u;
            __syncwarp();
            const int shift = 64 - bits - dig;
            for (int i0 = 0; i0 < total; i0 += 32) {
                const int i = i0 + lane;
                bool in = false;
                unsigned b = 0;
                if (i < total) {
                    const unsigned long long k = sk[i];
                    in = bits == 0 || (k >> (64 - bits)) == prefix;
                    b = (unsigned)((k >> shift) & ((1u << dig) - 1u));
                }
                // plain shared-memory atomics: the r02 source profile had the warp-aggregated version
                // (match.any + ffs + popc) waiting on MATCH for most of the kernel
                if (in) atomicAdd(&hist[b], 1u);
            }
            __syncwarp();
            unsigned loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = hist[255 - (lane * 8 + j)]; sum += loc[j]; }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
            const unsigned excl = incl - sum;
            const bool mine = excl < (unsigned)krem && (unsigned)krem <= incl;
            int bucket = 0, nk = 0, fin = 0;
            if (mine) {
                unsigned above = excl;
                int j = 0;
#pragma unroll
                for (int jj = 0; jj < 7; ++jj) { if (j == jj && above + loc[jj] < (unsigned)krem) { above += loc[jj]; ++j; } }
                unsigned lj = 0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) if (jj == j) lj = loc[jj];
                bucket = 255 - (lane * 8 + j);
                nk = krem - (int)above;
                fin = (lj == (unsigned)nk) || (bits + dig >= 64);
            }
            const unsigned who = __ballot_sync(FULL, mine);
            const int src = __ffs(who) - 1;
            bucket = __shfl_sync(FULL, bucket, src);
            krem = __shfl_sync(FULL, nk, src);
            done = __shfl_sync(FULL, fin, src) != 0;
            prefix = (prefix << dig) | (unsigned long long)bucket;
            bits += dig;
            __syncwarp();
        }
        out_prefix = prefix;
        out_bits = bits;
    };
    unsigned long long thr_prefix = 0ull;
    int thr_bits = 0;
    select_top(KP, thr_prefix, thr_bits);
    const int newk = min(total, KP);
    // OPT-IN (RR_TC_PRUNE=1; r02 measured no wall-clock gain on a 1.25 M-row shard: ~30 % of the k' = 160 rows are pruned,
    // but the rescoring kernel of such a small shortlist is latency-bound and the extra selection costs what it saves).
    // Final pass of a sharded round (prune_m = the shard's pool m < k'): rows whose bf16 score is more than 2 eps below the
    // m-th best bf16 score cannot be among the exact top-m (at least m rows have exact >= t_m - eps, such a row has exact
    // < t_m - eps), so they need no exact rescoring: their row is reported as -1 (exact = -inf, sorted last).
    float prune_below = -INFINITY;
    if (final_pass && prune_m > 0 && prune_m < newk) {
        unsigned long long mp = 0ull;
        int mb = 0;
        select_top(prune_m, mp, mb);
        unsigned long long mmin = ~0ull;
        for (int i = lane; i < total; i += 32) {
            const unsigned long long k = sk[i];
            if (mb == 0 || (k >> (64 - mb)) >= mp) mmin = min(mmin, k);
        }
        for (int o = 16; o > 0; o >>= 1) mmin = min(mmin, __shfl_xor_sync(FULL, mmin, o));
        prune_below = rr_key_score(mmin) - 2.0f * eps_rel * qnorm[q];
    }
    // ---- compact survivors: rows to rescore from the front, pruned rows (final pass only) from the back, so that the
    // rescoring warps (4 consecutive slots each) of the pruned tail have nothing to load -----------------------------
    unsigned long long mymin = ~0ull;
    int front = 0, back = 0;
    for (int i0 = 0; i0 < total; i0 += 32) {
        const int i = i0 + lane;
        bool sel = false;
        unsigned long long k = 0ull;
        if (i < total) { k = sk[i]; sel = thr_bits == 0 || (k >> (64 - thr_bits)) >= thr_prefix; }
        const bool pruned = sel && rr_key_score(k) < prune_below;
        const unsigned mv = __ballot_sync(FULL, sel && !pruned);
        const unsigned mp = __ballot_sync(FULL, pruned);
        if (sel) {
            const unsigned below = (1u << lane) - 1u;
            const int slot = pruned ? newk - 1 - (back + __popc(mp & below)) : front + __popc(mv & below);
            if (slot >= 0 && slot < KP) {
                kept_keys[(size_t)q * KP + slot] = k;
                if (final_pass) rows_out[(size_t)q * KP + slot] = pruned ? -1ll : (long long)rr_key_index(k);
            }
            mymin = min(mymin, k);
        }
        front += __popc(mv);
        back += __popc(mp);
    }
    for (int o = 16; o > 0; o >>= 1) mymin = min(mymin, __shfl_xor_sync(FULL, mymin, o));
    if (final_pass)
        for (int i = newk + lane; i < KP; i += 32) rows_out[(size_t)q * KP + i] = -1ll;
    for (int r = lane; r < n_sub; r += 32) cand_cnt[(size_t)q * n_sub + r] = 0u;
    if (lane == 0) {
        kept_cnt[q] = newk;
        tau[q] = newk == KP ? rr_key_score(mymin) : -INFINITY;
        if (over) overflow[q] = 1;
    }
}

__global__ void __launch_bounds__(256)
tc_finalize_kernel(const unsigned long long* __restrict__ kept_keys, const int* __restrict__ kept_cnt, int KP,
                   const float* __restrict__ exact, const int* __restrict__ overflow, const float* __restrict__ qnorm,
                   const float* __restrict__ tau, float eps_rel, int pool, long long* __restrict__ out_idx, float* __restrict__ out_sims,
                   int32_t* __restrict__ out_count, int* __restrict__ n_flagged, int* __restrict__ flagged,
                   int* __restrict__ uncertified) {
    extern __shared__ unsigned long long sk[];
    const int q = blockIdx.x;
    const int kc = kept_cnt[q];
    int n_pad = 2;
    while (n_pad < KP) n_pad <<= 1;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        unsigned long long k = 0ull;
        if (i < kc) {
            k = rr_make_key(exact[(size_t)q * KP + i], rr_key_index(kept_keys[(size_t)q * KP + i]));
            if (k == 0ull) k = 1ull;
        }
        sk[i] = k;
    }
    __syncthreads();
    bitonic_keys_desc(sk, n_pad);
    const int P = min(pool, kc);
    for (int i = threadIdx.x; i < pool; i += blockDim.x) {
        out_idx[(size_t)q * pool + i] = i < P ? (long long)rr_key_index(sk[i]) : -1ll;
        out_sims[(size_t)q * pool + i] = i < P ? rr_key_score(sk[i]) : -INFINITY;
    }
    if (threadIdx.x == 0) {
        if (out_count) out_count[q] = P;
        bool ok = overflow[q] == 0;
        if (ok && kc == KP) {
            // every row outside the shortlist has bf16 score <= c, so exact score <= c + eps
            const float c = tau[q];                   // the KP-th best bf16 score
            const float eps = eps_rel * qnorm[q];
            const float pth = P > 0 ? rr_key_score(sk[P - 1]) : INFINITY;
            ok = pth > c + eps;
        }
        if (!ok) flagged[atomicAdd(n_flagged, 1)] = q;
        if (uncertified) uncertified[q] = ok ? 0 : 1;
    }
}

__global__ void tc_reset_kernel(unsigned* cand_cnt, int n_cnt, int* kept_cnt, float* tau, int* overflow,
                                int* n_flagged, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cnt) cand_cnt[i] = 0u;
    if (i < B) { kept_cnt[i] = 0; tau[i] = -INFINITY; overflow[i] = 0; }
    if (i == 0) *n_flagged = 0;
}
__global__ void tc_gather_queries_kernel(const float* __restrict__ q, const int* __restrict__ flagged, int D,
                                         float* __restrict__ out) {
    const int src = flagged[blockIdx.x];
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[(size_t)blockIdx.x * D + d] = q[(size_t)src * D + d];
}
__global__ void tc_scatter_results_kernel(const int* __restrict__ flagged, int pool, const long long* __restrict__ idx_in,
                                          const float* __restrict__ sims_in, const int32_t* __restrict__ cnt_in,
                                          long long* __restrict__ idx, float* __restrict__ sims, int32_t* __restrict__ cnt) {
    const int dst = flagged[blockIdx.x];
    for (int i = threadIdx.x; i < pool; i += blockDim.x) {
        idx[(size_t)dst * pool + i] = idx_in[(size_t)blockIdx.x * pool + i];
        sims[(size_t)dst * pool + i] = sims_in[(size_t)blockIdx.x * pool + i];
    }
    if (threadIdx.x == 0 && cnt) cnt[dst] = cnt_in[blockIdx.x];
}

// debug: scatter the (score bits, doc) candidates of an all-pass filter launch into a dense [B, n_rows] matrix
__global__ void tc_debug_scatter_kernel(const unsigned long long* __restrict__ cand_keys, const unsigned* __restrict__ cand_cnt,
                                        int n_sub, int cap_sub, long long row0, int n_rows, float* __restrict__ out) {
    const int q = blockIdx.x;
    for (int s = 0; s < n_sub; ++s) {
        const unsigned cnt = min(cand_cnt[(size_t)q * n_sub + s], (unsigned)cap_sub);
        const unsigned long long* src = cand_keys + ((size_t)q * n_sub + s) * cap_sub;
        for (unsigned e = threadIdx.x; e < cnt; e += blockDim.x) {
            const unsigned long long k = src[e];
            const long long r = (long long)(k >> 32) - row0;
            if (r >= 0 && r < n_rows) out[(size_t)q * n_rows + r] = __uint_as_float((uint32_t)k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap_bf16_rows(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return rr_fail(RR_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return rr_fail(RR_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RR_OK;
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RR_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return rr_fail(RR_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return RR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct rr_tc_state {
    CUtensorMap tmap_c, tmap_c_half;      // corpus rows in boxes of 256 (one-CTA tiles) / 128 (this CTA's half of a pair's tile)
    const void* tmap_c_base = nullptr;
    bool pair_attr_set = false;
    Buf q_bf16, qnorm, cand_keys, cand_cnt, kept_keys, kept_cnt, tau, overflow, rows, exact, flags;
    Buf fb_q[2], fb_idx[2], fb_sims[2], fb_cnt[2], fb_flag[2];     // per fallback depth
    int* h_nflag = nullptr;   // pinned: [0] read-back of the synchronous path, [1] of the last deferred call
    bool attr_set = false;
    // Shortlist feedback.  How many rows lie within the certification margin of the pool-th score depends on the data
    // (score spread ~ 1/sqrt(D) for unit rows: 768-d needs a longer shortlist than 384-d for the same pool).  When
    // more than 1/64 of a batch comes back uncertified, the shortlist of the following calls on this index grows.
    double boost = 1.0;
    cudaEvent_t deferred_done = nullptr;
    bool deferred_pending = false;
    int deferred_batch = 0;
    void feedback(int n_flagged, int B) {
        if (n_flagged > std::max(1, B / 64)) boost = std::min(boost * 1.3, 4.0);
    }
};

bool rr_tc_supported(int cc_major, int) { return cc_major == 10; }

void rr_tc_destroy(rr_tc_state* s) {
    if (!s) return;
    for (Buf* b : {&s->q_bf16, &s->qnorm, &s->cand_keys, &s->cand_cnt, &s->kept_keys, &s->kept_cnt, &s->tau, &s->overflow,
                   &s->rows, &s->exact, &s->flags, &s->fb_q[0], &s->fb_idx[0], &s->fb_sims[0], &s->fb_cnt[0], &s->fb_flag[0],
                   &s->fb_q[1], &s->fb_idx[1], &s->fb_sims[1], &s->fb_cnt[1], &s->fb_flag[1]})
        b->release();
    if (s->h_nflag) cudaFreeHost(s->h_nflag);
    if (s->deferred_done) cudaEventDestroy(s->deferred_done);
    delete s;
}

// Corpus prefix growth per segment.  A segment brings ~ (growth-1)*k' new candidates per query, so fewer, longer
// segments trade selection launches against the size of each selection.  Measured in r02 (10 M x 384, B = 4096):
// growth 8 (5 segments instead of 8 on one GPU, 5 instead of 7 on a 1.25 M-row shard) is no faster on one GPU and
// 3 % SLOWER on eight (3.67 vs 3.57 ms / step): the selections get bigger as fast as they get fewer.  Growth 16
// overflows the selection capacity.  Default 4, RR_TC_GROWTH overrides (2..16).
static int KP_growth(int kp) {
    const char* env = getenv("RR_TC_GROWTH");
    int g = env ? atoi(env) : 0;
    if (g < 2 || g > 16) g = TC_GROWTH;
    while (g > 2 && (long long)kp * (g + g / 2 + 2) > TC_SORT_MAX) --g;
    return g;
}

static int shortlist_size(int pool) {
    const char* env = getenv("RR_TC_SHORTLIST_FACTOR");
    double f = env ? atof(env) : 2.6;
    if (!(f >= 1.0)) f = 2.6;
    // rows within the certification margin of the pool-th score: ~0.9*pool on unit-norm 384-d data, plus five
    // standard deviations and a constant so that small pools (sharded round 1) certify as reliably as big ones.
    // This is only the starting point: the shortlist feedback (rr_tc_state::boost) lengthens it when the data
    // needs more (wider rows, clustered catalogues).
    const double base = 1.9 * pool;
    int kp = (int)std::ceil(std::max(pool * f, base + 5.0 * std::sqrt(base) + 16.0));
    kp = (kp + 31) / 32 * 32;
    return std::max(kp, 64);
}

bool rr_tc_can_handle(int dim_pad, int pool) {
    return dim_pad > 0 && dim_pad <= TC_MAX_KB_STREAMED * TC_BK && shortlist_size(pool) <= TC_SORT_MAX / 4;
}

int rr_tc_dense_topk(rr_tc_state** state, const rr_index_desc* d, int sm_count, const float* d_q, int32_t B,
                     int32_t pool, int64_t* d_idx, float* d_sims, int32_t* d_count, rr_dense_stats* stats,
                     rr_exact_fn exact_fn, void* exact_ctx, int32_t* d_uncertified, cudaStream_t s, int kp_override,
                     int depth) {
    if (d->dim_pad > TC_MAX_KB_STREAMED * TC_BK)
        return rr_fail(RR_EUNSUPPORTED, "tensor path supports dim <= %d in this build", TC_MAX_KB_STREAMED * TC_BK);
    if (shortlist_size(pool) > TC_SORT_MAX / 4)
        return rr_fail(RR_EUNSUPPORTED, "tensor path supports pool <= %d", (int)(TC_SORT_MAX / 4 / 2.6));
    if (!*state) {
        *state = new (std::nothrow) rr_tc_state();
        if (!*state) return rr_fail(RR_ENOMEM, "out of host memory");
    }
    rr_tc_state* st = *state;
    if (!st->h_nflag) RR_CUDA(cudaMallocHost(&st->h_nflag, 2 * sizeof(int)));
    if (!st->deferred_done) RR_CUDA(cudaEventCreateWithFlags(&st->deferred_done, cudaEventDisableTiming));
    if (st->deferred_pending && cudaEventQuery(st->deferred_done) == cudaSuccess) {
        st->feedback(st->h_nflag[1], st->deferred_batch);          // outcome of the previous sync-free call
        st->deferred_pending = false;
    }
    const int KP = kp_override > 0 ? kp_override
                                   : std::min(TC_SORT_MAX / 4, (int)((std::ceil(shortlist_size(pool) * st->boost) + 31) / 32) * 32);
    const int growth = KP_growth(KP);
    if (!st->attr_set) {
        RR_CUDA(cudaFuncSetAttribute(tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_TOTAL));
        RR_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SORT_MAX * 8));
        RR_CUDA(cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SORT_MAX * 8));
        RR_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RR_CUDA(cudaFuncSetAttribute(tc_select_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        RR_CUDA(cudaFuncSetAttribute(tc_select_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RR_CUDA(cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        st->attr_set = true;
    }
    if (st->tmap_c_base != d->d_emb_bf16) {
        RR_TRY(make_tmap_bf16_rows(&st->tmap_c, d->d_emb_bf16, (uint64_t)d->n_docs, (uint64_t)d->dim_pad, TC_BN));
        RR_TRY(make_tmap_bf16_rows(&st->tmap_c_half, d->d_emb_bf16, (uint64_t)d->n_docs, (uint64_t)d->dim_pad, TC_BN / 2));
        st->tmap_c_base = d->d_emb_bf16;
    }
    // CTA pairs (tcgen05 cta_group::2): resident query tiles only (dim_pad <= 384), and enough query tiles that padding
    // their number to an even one does not waste much (an odd tile out of >= 8 costs < 13 %)
    const int n_qt_real = (B + TC_BM - 1) / TC_BM;
    const char* pair_env = getenv("RR_TC_PAIR");
    const bool use_pair = !(pair_env && pair_env[0] == '0') && d->dim_pad <= TC_MAX_KB * TC_BK &&
                          (n_qt_real >= 8 || (n_qt_real >= 2 && n_qt_real % 2 == 0));
    const int n_qt = use_pair ? (n_qt_real + 1) / 2 * 2 : n_qt_real;
    const int B_pad = n_qt * TC_BM;
    if (use_pair && !st->pair_attr_set) {
        RR_CUDA(cudaFuncSetAttribute(tc_filter_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_total(TC2_MAX_STAGES)));
        st->pair_attr_set = true;
    }
    const char* ps_env = getenv("RR_TC_PAIR_STAGES");
    // 6 stages (96 KB of queries + 96 KB of ring): as fast as 8 on its own (r02: 21.5 ms / step either way), and it leaves
    // ~33 KB of shared memory per SM, so that the small tail kernels of ANOTHER batch in flight (selection with 2 warps per
    // CTA, rescoring, candidate BM25, finalize) can co-reside with the GEMM instead of waiting for it
    // (1.25 M-row shard, 2 batches in flight: 4.24 -> 4.18 ms / step)
    const int pair_stages = std::max(2, std::min(TC2_MAX_STAGES, ps_env ? atoi(ps_env) : 6));
    const int sm_units = use_pair ? sm_count / 2 * 2 : sm_count;      // pairs occupy whole TPCs

    // Split the query tiles into parts so that (tiles per part) x (CTAs per tile) fills the SMs; an extra
    // part re-reads the corpus from HBM, so it has to buy at least 5 % more SM occupancy.
    int best_parts = 0;
    double best_cost = 1e30;
    const int unit = use_pair ? 2 : 1;                 // query tiles are dealt to the parts in units of `unit`
    const int n_units = n_qt / unit;
    auto part_tiles = [&](int parts, int p) { return unit * (n_units / parts + (p < n_units % parts ? 1 : 0)); };
    for (int parts = 1; parts <= 8 && parts <= n_units; ++parts) {
        double cost = 0;
        bool ok = true;
        for (int p = 0; p < parts; ++p) {
            const int nq = part_tiles(parts, p);
            if (nq > sm_units) { ok = false; break; }
            cost += 1.0 / (double)(sm_units / nq);
        }
        if (ok && cost < best_cost * 0.95) { best_cost = cost; best_parts = parts; }
    }
    if (best_parts == 0) return rr_fail(RR_EUNSUPPORTED, "batch too large for the tensor path (%d query tiles)", n_qt);
    int reps_max = 1, reps_min = sm_units;
    for (int p = 0; p < best_parts; ++p) {
        const int nq = part_tiles(best_parts, p);
        reps_max = std::max(reps_max, sm_units / nq);
        reps_min = std::min(reps_min, sm_units / nq);
    }
    const int n_sub = 2 * reps_max;                                              // one sub-list per (CTA, column half)
    if (n_sub > TC_MAX_SUB) return rr_fail(RR_EUNSUPPORTED, "too many SMs for the candidate sub-list table");
    const int cap_sub = std::max(TC_BN / 2, TC_CAP_TOTAL / n_sub / 32 * 32);     // >= half of an all-pass tile

    RR_TRY(st->q_bf16.ensure(sizeof(__nv_bfloat16) * (size_t)B_pad * d->dim_pad));
    RR_TRY(st->qnorm.ensure(sizeof(float) * (size_t)B_pad));
    RR_TRY(st->cand_keys.ensure(sizeof(unsigned long long) * (size_t)B * n_sub * cap_sub));
    RR_TRY(st->cand_cnt.ensure(sizeof(unsigned) * (size_t)B * n_sub));
    RR_TRY(st->kept_keys.ensure(sizeof(unsigned long long) * (size_t)B * KP));
    RR_TRY(st->kept_cnt.ensure(sizeof(int) * (size_t)B));
    RR_TRY(st->tau.ensure(sizeof(float) * (size_t)B));
    RR_TRY(st->overflow.ensure(sizeof(int) * (size_t)B));
    RR_TRY(st->rows.ensure(sizeof(long long) * (size_t)B * KP));
    RR_TRY(st->exact.ensure(sizeof(float) * (size_t)B * KP));
    RR_TRY(st->flags.ensure(sizeof(int) * ((size_t)B + 1)));
    int* n_flagged = static_cast<int*>(st->flags.p);
    int* flagged = n_flagged + 1;

    CUtensorMap tmap_q;
    RR_TRY(make_tmap_bf16_rows(&tmap_q, st->q_bf16.p, (uint64_t)B_pad, (uint64_t)d->dim_pad, TC_BM));

    {
        RrProfScope prof(RR_PROF_MISC, s);
        cvt_queries_kernel<<<B_pad, 128, 0, s>>>(d_q, B, d->dim, d->dim_pad, B_pad,
                                                 static_cast<__nv_bfloat16*>(st->q_bf16.p), static_cast<float*>(st->qnorm.p));
        RR_LAUNCH_CHECK();
        const int n_cnt = B * n_sub;
        tc_reset_kernel<<<(std::max(n_cnt, B) + 255) / 256, 256, 0, s>>>(
            static_cast<unsigned*>(st->cand_cnt.p), n_cnt, static_cast<int*>(st->kept_cnt.p),
            static_cast<float*>(st->tau.p), static_cast<int*>(st->overflow.p), n_flagged, B);
        RR_LAUNCH_CHECK();
    }

    const int n_dt = (int)((d->n_docs + TC_BN - 1) / TC_BN);
    // first segment: tau = -inf, everything passes (256 keys per tile and query go through the selection), so it is
    // sized to bring about as many keys as a later segment is expected to, (growth-1)*k'.  Limits: every (CTA, query,
    // half) sub-list must hold all the tiles its CTA scans (cap_sub / 128 of them), the total must stay sortable.
    const char* seg0_env = getenv("RR_TC_SEG0_TILES");
    const int seg0_want = seg0_env && atoi(seg0_env) > 0 ? atoi(seg0_env) : ((growth - 1) * KP + TC_BN - 1) / TC_BN;
    const int seg0_tiles = std::max(1, std::min(seg0_want, std::min(reps_min * std::max(1, cap_sub / (TC_BN / 2)),
                                                                    (TC_SORT_MAX - KP) / TC_BN)));
    int n_segments = 0;
    int dt_lo = 0;
    while (dt_lo < n_dt) {
        const int dt_hi = dt_lo == 0 ? std::min(n_dt, seg0_tiles)
                                     : (int)std::min<long long>(n_dt, (long long)dt_lo * growth);
        {
            // ONE launch per segment for all query parts (r02: every filter launch has ~30 us of fixed cost -- TMEM
            // allocation, barrier set-up, the resident query tile, pipeline fill and drain -- which the small early
            // segments paid once per part); the parts' CTAs follow each other in the grid, so part p+1 starts on the SMs
            // part p frees.
            TcFilterArgs a;
            a.n_docs = d->n_docs; a.n_kb = d->dim_pad / TC_BK; a.B = B;
            a.dt_lo = dt_lo; a.dt_hi = dt_hi; a.tau = static_cast<const float*>(st->tau.p);
            a.q_resident = d->dim_pad <= TC_MAX_KB * TC_BK ? 1 : 0;
            a.n_sub = n_sub; a.cap_sub = cap_sub; a.pair_stages = pair_stages;
            a.cand_keys = static_cast<unsigned long long*>(st->cand_keys.p);
            a.cand_cnt = static_cast<unsigned*>(st->cand_cnt.p);
            a.n_parts = best_parts;
            int qt0 = 0, grid = 0;
            for (int p = 0; p < 8; ++p) { a.part_qt0[p] = a.part_nq[p] = a.part_reps[p] = a.part_ctas[p] = 0; }
            for (int p = 0; p < best_parts; ++p) {
                const int nq = part_tiles(best_parts, p);
                a.part_qt0[p] = qt0; a.part_nq[p] = nq;
                a.part_reps[p] = std::max(1, std::min(sm_units / nq, dt_hi - dt_lo));
                a.part_ctas[p] = nq * a.part_reps[p];
                grid += a.part_ctas[p];
                qt0 += nq;
            }
            {
                RrProfScope prof(RR_PROF_TC_FILTER, s);
                if (use_pair) {
                    int grid_pair = 0;             // persistent over the parts: CTA c serves part 0, then part 1, ...
                    for (int p = 0; p < best_parts; ++p) grid_pair = std::max(grid_pair, a.part_ctas[p]);
                    tc_filter_pair_kernel<<<grid_pair, TC_THREADS, tc2_smem_total(pair_stages), s>>>(tmap_q, st->tmap_c_half, a);
                } else
                    tc_filter_kernel<<<grid, TC_THREADS, TC_SMEM_TOTAL, s>>>(tmap_q, st->tmap_c, a);
            }
            RR_LAUNCH_CHECK();
        }
        const int final_pass = dt_hi >= n_dt;
        {
            // keys a query can bring to this selection: everything of the all-pass first segment, else the
            // kept list plus ~ (growth-1)*k' expected passes (2x head-room; more is flagged as overflow
            // and that query is redone exactly)
            // (kept k' + (growth-1) k' expected passes) with 50 % head-room
            // (a prefix that under-represents the rest of the corpus -- duplicates, rows sorted by category -- lets more
            // rows pass than (growth-1)*k': the head-room is 50 % of that plus k', r02 probe: 25 % was not enough)
            const int expect = dt_lo == 0 ? (dt_hi - dt_lo) * TC_BN + KP : KP * (growth + growth / 2 + 2);
            const int sort_cap = std::min(TC_SORT_MAX, std::max(1024, (expect + 255) / 256 * 256));
            RrProfScope prof(RR_PROF_TC_SELECT, s);
            static const int warp_max = getenv("RR_TC_SELECT_WARP_MAX") ? atoi(getenv("RR_TC_SELECT_WARP_MAX")) : 2048;
            static const int wpc_max = getenv("RR_TC_SELECT_WPC") ? std::max(1, std::min(8, atoi(getenv("RR_TC_SELECT_WPC")))) : 2;   // see pair_stages
            if (sort_cap <= warp_max) {
                // warp per query: 4..8 queries per CTA, no block barriers
                const size_t per_warp_words = (size_t)sort_cap + 128 + (TC_MAX_SUB + 2) / 2 + 1;
                const int wpc = (int)std::max<size_t>(1, std::min<size_t>((size_t)wpc_max, (size_t)(96 * 1024) / (per_warp_words * 8)));
                tc_select_warp_kernel<<<(B + wpc - 1) / wpc, wpc * 32, per_warp_words * 8 * wpc, s>>>(
                    static_cast<const unsigned long long*>(st->cand_keys.p), static_cast<unsigned*>(st->cand_cnt.p), n_sub,
                    cap_sub, static_cast<unsigned long long*>(st->kept_keys.p), static_cast<int*>(st->kept_cnt.p), KP,
                    static_cast<float*>(st->tau.p), static_cast<int*>(st->overflow.p), final_pass,
                    static_cast<long long*>(st->rows.p), sort_cap, wpc, B, getenv("RR_TC_PRUNE") ? pool : 0,
                    static_cast<const float*>(st->qnorm.p), (0.0078125f + 0.000030517578125f + 1e-4f) * d->max_row_norm);
            } else {
                tc_select_kernel<<<B, 256, (size_t)sort_cap * 8, s>>>(
                    static_cast<const unsigned long long*>(st->cand_keys.p), static_cast<unsigned*>(st->cand_cnt.p), n_sub,
                    cap_sub, static_cast<unsigned long long*>(st->kept_keys.p), static_cast<int*>(st->kept_cnt.p), KP,
                    static_cast<float*>(st->tau.p), static_cast<int*>(st->overflow.p), final_pass,
                    static_cast<long long*>(st->rows.p), sort_cap);
            }
        }
        RR_LAUNCH_CHECK();
        dt_lo = dt_hi;
        ++n_segments;
    }
    RR_TRY(rr_launch_rescore(d->d_emb_f32, d->n_docs, d->dim, d_q, static_cast<const int64_t*>(st->rows.p), KP, B,
                             static_cast<float*>(st->exact.p), s));
    const float eps_rel = (0.0078125f + 0.000030517578125f + 1e-4f) * d->max_row_norm;
    {
        RrProfScope prof(RR_PROF_TC_FINALIZE, s);
        int fin_pad = 2;
        while (fin_pad < KP) fin_pad <<= 1;
        tc_finalize_kernel<<<B, 256, (size_t)fin_pad * 8, s>>>(static_cast<const unsigned long long*>(st->kept_keys.p),
                                                    static_cast<const int*>(st->kept_cnt.p), KP,
                                                    static_cast<const float*>(st->exact.p),
                                                    static_cast<const int*>(st->overflow.p),
                                                    static_cast<const float*>(st->qnorm.p),
                                                    static_cast<const float*>(st->tau.p), eps_rel, pool,
                                                    reinterpret_cast<long long*>(d_idx), d_sims, d_count, n_flagged, flagged,
                                                    d_uncertified);
    }
    RR_LAUNCH_CHECK();
    if (d_uncertified != nullptr) {
        // deferred mode: nothing is read back; the caller acts on the per-query mask (uncertified results
        // are best-effort and must be redone by a synchronous call)
        if (stats) {
            stats->path = 2; stats->n_uncertified = -1; stats->n_overflow = 0; stats->shortlist = KP;
            stats->n_segments = n_segments; stats->eps = eps_rel;
        }
        if (!st->deferred_pending) {       // feedback for the next call, read without ever blocking
            RR_CUDA(cudaMemcpyAsync(st->h_nflag + 1, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, s));
            RR_CUDA(cudaEventRecord(st->deferred_done, s));
            st->deferred_pending = true;
            st->deferred_batch = B;
        }
        return RR_OK;
    }
    RR_CUDA(cudaMemcpyAsync(st->h_nflag, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, s));
    RR_CUDA(cudaStreamSynchronize(s));
    const int nf = *st->h_nflag;
    if (depth == 0) st->feedback(nf, B);
    if (stats && depth == 0) {
        stats->path = 2; stats->n_uncertified = nf; stats->n_overflow = 0; stats->shortlist = KP;
        stats->n_segments = n_segments; stats->eps = eps_rel;
    }
    if (nf > 0) {
        // Uncertified queries (rows denser than the margin around the pool-th score, candidate overflow): first a SECOND
        // TENSOR PASS over the corpus for those queries only, with a 4x longer shortlist -- one bf16 sweep for up to 128
        // queries costs less than the fp32 GEMV does for 8 -- and only what still cannot be proven goes to the exact
        // fp32 path.  Results never depend on bf16 either way.
        Buf &fq = st->fb_q[depth], &fi = st->fb_idx[depth], &fs = st->fb_sims[depth], &fc = st->fb_cnt[depth], &ff = st->fb_flag[depth];
        RR_TRY(fq.ensure(sizeof(float) * (size_t)nf * d->dim));
        RR_TRY(fi.ensure(sizeof(long long) * (size_t)nf * pool));
        RR_TRY(fs.ensure(sizeof(float) * (size_t)nf * pool));
        RR_TRY(fc.ensure(sizeof(int32_t) * (size_t)nf));
        RR_TRY(ff.ensure(sizeof(int) * (size_t)nf));
        RR_CUDA(cudaMemcpyAsync(ff.p, flagged, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToDevice, s));
        const int* my_flags = static_cast<const int*>(ff.p);
        tc_gather_queries_kernel<<<nf, 128, 0, s>>>(d_q, my_flags, d->dim, static_cast<float*>(fq.p));
        RR_LAUNCH_CHECK();
        const int KP2 = std::min(TC_SORT_MAX / 4, KP * 4);
        if (depth == 0 && KP2 > KP && !getenv("RR_TC_NO_SECOND_PASS")) {
            RR_TRY(rr_tc_dense_topk(state, d, sm_count, static_cast<const float*>(fq.p), nf, pool, static_cast<int64_t*>(fi.p),
                                    static_cast<float*>(fs.p), static_cast<int32_t*>(fc.p), nullptr, exact_fn, exact_ctx, nullptr,
                                    s, KP2, 1));
            if (stats) stats->n_overflow = *st->h_nflag;          // queries the second pass still had to hand to the exact path
        } else {
            RR_TRY(exact_fn(exact_ctx, static_cast<const float*>(fq.p), nf, pool, static_cast<int64_t*>(fi.p),
                            static_cast<float*>(fs.p), static_cast<int32_t*>(fc.p), s));
        }
        tc_scatter_results_kernel<<<nf, 128, 0, s>>>(my_flags, pool, static_cast<const long long*>(fi.p),
                                                     static_cast<const float*>(fs.p), static_cast<const int32_t*>(fc.p),
                                                     reinterpret_cast<long long*>(d_idx), d_sims, d_count);
        RR_LAUNCH_CHECK();
    }
    return RR_OK;
}
// Debug / test entry: the raw bf16 x bf16 -> fp32 tensor-core scores of rows [row0, row0 + n_rows) for up to 128
// queries, exactly as tc_filter_kernel sees them (tau = -inf, every score is kept), so that the stated tolerance of
// the shortlist stage ("within 1e-3 absolute before rescoring") can be asserted against the exact similarities.
int rr_tc_debug_scores(rr_tc_state** state, const rr_index_desc* d, int sm_count, const float* d_q, int32_t B,
                       int64_t row0, int32_t n_rows, float* d_out, cudaStream_t s) {
    if (B <= 0 || B > TC_BM) return rr_fail(RR_EINVAL, "rr_dense_debug_bf16_scores: 1..%d queries", TC_BM);
    if (row0 < 0 || (row0 % TC_BN) || n_rows <= 0 || row0 + n_rows > d->n_docs)
        return rr_fail(RR_EINVAL, "rr_dense_debug_bf16_scores: row0 must be a multiple of %d and the range inside the corpus", TC_BN);
    const int n_dt = (n_rows + TC_BN - 1) / TC_BN;
    if (n_dt > sm_count) return rr_fail(RR_EINVAL, "rr_dense_debug_bf16_scores: at most %d rows per call", sm_count * TC_BN);
    if (d->dim_pad > TC_MAX_KB_STREAMED * TC_BK) return rr_fail(RR_EUNSUPPORTED, "tensor path supports dim <= %d", TC_MAX_KB_STREAMED * TC_BK);
    if (!*state) {
        *state = new (std::nothrow) rr_tc_state();
        if (!*state) return rr_fail(RR_ENOMEM, "out of host memory");
    }
    rr_tc_state* st = *state;
    if (!st->attr_set) {
        RR_CUDA(cudaFuncSetAttribute(tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_TOTAL));
        RR_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SORT_MAX * 8));
        RR_CUDA(cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SORT_MAX * 8));
        RR_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RR_CUDA(cudaFuncSetAttribute(tc_select_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        RR_CUDA(cudaFuncSetAttribute(tc_select_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        RR_CUDA(cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        st->attr_set = true;
    }
    if (st->tmap_c_base != d->d_emb_bf16) {
        RR_TRY(make_tmap_bf16_rows(&st->tmap_c, d->d_emb_bf16, (uint64_t)d->n_docs, (uint64_t)d->dim_pad, TC_BN));
        st->tmap_c_base = d->d_emb_bf16;
    }
    const int B_pad = TC_BM, reps = n_dt, n_sub = 2 * reps, cap_sub = TC_BN / 2;
    RR_TRY(st->q_bf16.ensure(sizeof(__nv_bfloat16) * (size_t)B_pad * d->dim_pad));
    RR_TRY(st->qnorm.ensure(sizeof(float) * (size_t)B_pad));
    RR_TRY(st->cand_keys.ensure(sizeof(unsigned long long) * (size_t)B * n_sub * cap_sub));
    RR_TRY(st->cand_cnt.ensure(sizeof(unsigned) * (size_t)B * n_sub));
    RR_TRY(st->kept_cnt.ensure(sizeof(int) * (size_t)B));
    RR_TRY(st->tau.ensure(sizeof(float) * (size_t)B));
    RR_TRY(st->overflow.ensure(sizeof(int) * (size_t)B));
    RR_TRY(st->flags.ensure(sizeof(int) * ((size_t)B + 1)));
    CUtensorMap tmap_q;
    RR_TRY(make_tmap_bf16_rows(&tmap_q, st->q_bf16.p, (uint64_t)B_pad, (uint64_t)d->dim_pad, TC_BM));
    cvt_queries_kernel<<<B_pad, 128, 0, s>>>(d_q, B, d->dim, d->dim_pad, B_pad, static_cast<__nv_bfloat16*>(st->q_bf16.p),
                                             static_cast<float*>(st->qnorm.p));
    RR_LAUNCH_CHECK();
    const int n_cnt = B * n_sub;
    tc_reset_kernel<<<(std::max(n_cnt, B) + 255) / 256, 256, 0, s>>>(static_cast<unsigned*>(st->cand_cnt.p), n_cnt,
                                                                     static_cast<int*>(st->kept_cnt.p), static_cast<float*>(st->tau.p),
                                                                     static_cast<int*>(st->overflow.p), static_cast<int*>(st->flags.p), B);
    RR_LAUNCH_CHECK();
    TcFilterArgs a;
    a.n_docs = d->n_docs; a.n_kb = d->dim_pad / TC_BK; a.B = B;
    a.n_parts = 1;
    for (int p = 0; p < 8; ++p) { a.part_qt0[p] = a.part_nq[p] = a.part_reps[p] = a.part_ctas[p] = 0; }
    a.part_qt0[0] = 0; a.part_nq[0] = 1; a.part_reps[0] = reps; a.part_ctas[0] = reps;
    a.dt_lo = (int)(row0 / TC_BN); a.dt_hi = a.dt_lo + n_dt; a.tau = static_cast<const float*>(st->tau.p);
    a.q_resident = d->dim_pad <= TC_MAX_KB * TC_BK ? 1 : 0;
    a.n_sub = n_sub; a.cap_sub = cap_sub; a.pair_stages = 0;
    a.cand_keys = static_cast<unsigned long long*>(st->cand_keys.p);
    a.cand_cnt = static_cast<unsigned*>(st->cand_cnt.p);
    tc_filter_kernel<<<reps, TC_THREADS, TC_SMEM_TOTAL, s>>>(tmap_q, st->tmap_c, a);
    RR_LAUNCH_CHECK();
    tc_debug_scatter_kernel<<<B, 256, 0, s>>>(static_cast<const unsigned long long*>(st->cand_keys.p),
                                              static_cast<const unsigned*>(st->cand_cnt.p), n_sub, cap_sub, (long long)row0,
                                              n_rows, d_out);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
