// K4 -- score normalisation, Bayesian prior, trust, weighted fusion and final top-k (sm_100a).
//
// Stands behind the numeric tail of run_search (app/app_product_search.py:256-312, Streamlit
// driver, use_trust=1) and of search (app/test.py:252-309, CLI driver, use_trust=0):
//
//   _dense = minmax(dense)            utils.py:46-55   f32 arithmetic, divisor rounded to f32
//   _bm25  = minmax(bm25_raw)
//   prior  = bayes(avg, n, C) with g = nanmean(avg) over the POOL   utils.py:103-109 (float64)
//   vol    = log1p(n) / (max log1p(n) + 1e-9)                       :267            (float64)
//   _prior = minmax(prior) * 0.7 + 0.3 * vol                        :268   f32*f32 -> f64 add
//   _trust = (0.6*clip(n/max(min_reviews,1),0,1) + 0.4*min(1, log1p(n)/log1p(max(sat,1)))).f32   :238-242
//   final  = (w_d*_dense + w_b*_bm25 + w_r*_rerank + w_p*_prior + w_best*_best).astype(f32)     :306-308
//            with NumPy-2 promotion: f32 terms are multiplied by the weight rounded to f32 and
//            added in f32 until the first float64 term appears, from then on in float64
//   final  = final * _trust * _gate   (CLI: final * _gate)                                     :309-310
//   sort_values(_final, descending).head(k)                                                     :312
//
// One CTA per query.  The candidate tuples are first ordered by (dense desc, global row asc) and cut
// to `pool` -- for a single shard this is the identity, for a row-sharded corpus it is the
// cross-shard merge (K5) -- then fused, then ordered by (final desc, pool position asc).
// nanmean follows NumPy's pairwise summation order so that g is bit-identical.
#include "rr_internal.h"
#include "rr_kernels.h"

namespace {

constexpr int FUSE_THREADS = 256;
constexpr int FUSE_MAX_POOL = 2048;
constexpr int FUSE_MAX_IN = 8192;
constexpr int FUSE_MAX_LEAVES = 256;

struct FuseArgs {
    rr_fusion_params p;
    int B, n_in;
    int n_shards;            // tuples come from n_shards blocks of per_shard = n_in / n_shards entries
    long long shard_stride;  // bytes between consecutive shard blocks of one field
    const int32_t* count;
    const float* dense;
    const float* bm25;
    const double* n;
    const double* avg;
    const long long* grow;
    const float* rerank;
    const float* best;
    const float* gate;
    int extras_by_slot;      // 1: best / gate are fields of the input tuples (sharded layout, indexed like dense);
                             // 0: given in pool order [B, pool]
    long long* top_row;
    float* top_final;
    int32_t* top_pos;
    float* components;
    int32_t* incomplete;
};

__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7FF8000000000000ll); }

// block-wide bitonic sort, descending by key, carrying a 32-bit value
__device__ void bitonic_desc(unsigned long long* key, unsigned* val, int n_pad) {
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n_pad / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = key[lo], b = key[hi];
                if ((a < b) == desc) {
                    key[lo] = b; key[hi] = a;
                    const unsigned t = val[lo]; val[lo] = val[hi]; val[hi] = t;
                }
            }
            __syncthreads();
        }
    }
}

// n_pad <= blockDim.x (the default pools of 100 / 150 candidates): rank sort -- every thread counts the keys above its
// own (equal keys, i.e. the empty slots, by position) and moves its element there: two barriers instead of the 36 steps
// of the bitonic network, same order for distinct keys
__device__ void rank_sort_desc(unsigned long long* key, unsigned* val, int n_pad) {
    const int i = threadIdx.x;
    unsigned long long mine = 0ull;
    unsigned v = 0u;
    int rank = 0;
    if (i < n_pad) {
        mine = key[i]; v = val[i];
#pragma unroll 8
        for (int j = 0; j < n_pad; ++j) {
            const unsigned long long o = key[j];
            rank += (o > mine || (o == mine && j < i)) ? 1 : 0;
        }
    }
    __syncthreads();
    if (i < n_pad) { key[rank] = mine; val[rank] = v; }
    __syncthreads();
}

__device__ __forceinline__ void sort_desc(unsigned long long* key, unsigned* val, int n_pad) {
    if (n_pad <= (int)blockDim.x) rank_sort_desc(key, val, n_pad);
    else bitonic_desc(key, val, n_pad);
}

// Block-wide reduction, result on every thread.  op must be associative and commutative (min / max / or / integer add):
// the warps' partial results are combined by a second butterfly over aligned groups of 8 lanes (FUSE_THREADS / 32 = 8
// warps), not by a serial walk.  scratch: >= 8 elements of T in shared memory.
static_assert(FUSE_THREADS == 256, "block_reduce assumes 8 warps");
template <class T, class Op>
__device__ T block_reduce(T v, Op op, T* scratch) {
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = scratch[threadIdx.x & 7];
    for (int o = 4; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}

// min, max and "holds a NaN" of NV float columns over the first P pool entries: one pass, one pair of barriers.
// scratch_f >= 16 NV floats, scratch_i >= 8 ints.  bad: bit v set when column v holds a NaN.
template <int NV>
__device__ void block_minmax(const float* const (&col)[NV], int P, float (&lo)[NV], float (&hi)[NV], int& bad,
                             float* scratch_f, int* scratch_i) {
#pragma unroll
    for (int v = 0; v < NV; ++v) { lo[v] = INFINITY; hi[v] = -INFINITY; }
    bad = 0;
    for (int i = threadIdx.x; i < P; i += FUSE_THREADS) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float x = col[v][i];
            if (x != x) bad |= 1 << v;
            lo[v] = fminf(lo[v], x); hi[v] = fmaxf(hi[v], x);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            lo[v] = fminf(lo[v], __shfl_xor_sync(0xffffffffu, lo[v], o));
            hi[v] = fmaxf(hi[v], __shfl_xor_sync(0xffffffffu, hi[v], o));
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
#pragma unroll
        for (int v = 0; v < NV; ++v) { scratch_f[(2 * v) * 8 + w] = lo[v]; scratch_f[(2 * v + 1) * 8 + w] = hi[v]; }
        scratch_i[w] = bad;
    }
    __syncthreads();
    const int g = threadIdx.x & 7;
#pragma unroll
    for (int v = 0; v < NV; ++v) { lo[v] = scratch_f[(2 * v) * 8 + g]; hi[v] = scratch_f[(2 * v + 1) * 8 + g]; }
    bad = scratch_i[g];
    for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            lo[v] = fminf(lo[v], __shfl_xor_sync(0xffffffffu, lo[v], o));
            hi[v] = fmaxf(hi[v], __shfl_xor_sync(0xffffffffu, hi[v], o));
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
}

// NumPy pairwise-sum leaf enumeration (numpy/core/src/umath/loops_utils.h.src semantics):
// n < 8 or n <= 128 are leaves; otherwise split at n2 = (n/2) - (n/2)%8.
__device__ void pw_leaves(int lo, int n, int* leaf_lo, int* leaf_n, int& count) {
    if (n <= 128) {
        if (count < FUSE_MAX_LEAVES) { leaf_lo[count] = lo; leaf_n[count] = n; }
        ++count;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    pw_leaves(lo, n2, leaf_lo, leaf_n, count);
    pw_leaves(lo + n2, n - n2, leaf_lo, leaf_n, count);
}
__device__ double pw_combine(int n, const double* leaf_sum, int& next) {
    if (n <= 128) return leaf_sum[next++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double a = pw_combine(n2, leaf_sum, next);
    const double b = pw_combine(n - n2, leaf_sum, next);
    return a + b;
}

__global__ void __launch_bounds__(FUSE_THREADS)
fuse_topk_kernel(const FuseArgs a, int n_pad_in, int n_pad_pool) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // carve
    unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);
    const int n_pad = max(n_pad_in, n_pad_pool);
    unsigned* val = reinterpret_cast<unsigned*>(key + n_pad);
    double* s_n = reinterpret_cast<double*>(val + n_pad + (n_pad & 1));
    double* s_avg = s_n + a.p.pool;
    double* s_prior = s_avg + a.p.pool;
    float* s_dense = reinterpret_cast<float*>(s_prior + a.p.pool);
    float* s_bm25 = s_dense + a.p.pool;
    float* s_final = s_bm25 + a.p.pool;
    float* s_best = s_final + a.p.pool;
    float* s_gate = s_best + a.p.pool;
    __shared__ double red_d[32];
    __shared__ float red_f[64];
    __shared__ int red_i[32];
    __shared__ int s_leaf_lo[FUSE_MAX_LEAVES], s_leaf_n[FUSE_MAX_LEAVES];
    __shared__ double s_leaf_r[FUSE_MAX_LEAVES * 8];
    __shared__ double s_leaf_sum[FUSE_MAX_LEAVES];
    __shared__ int s_nleaf;
    __shared__ double s_g;

    const int b = blockIdx.x, tid = threadIdx.x;
    const int per_shard = a.n_in / a.n_shards;
    const int cnt_in = a.count ? min(a.count[b], a.n_in) : a.n_in;
    // element index of input slot i for a field whose elements are `esz` bytes wide
    auto at = [&](int i, int esz) -> long long {
        const int s = i / per_shard, j = i - s * per_shard;
        return (long long)s * (a.shard_stride / esz) + (long long)b * per_shard + j;
    };

    // ---- 1. order candidates by (dense desc, global row asc), cut to pool -----------------------
    int n_valid_local = 0;
    for (int i = tid; i < n_pad_in; i += FUSE_THREADS) {
        unsigned long long k = 0ull;
        if (i < a.n_in) {
            const long long g = a.grow[at(i, 8)];
            const bool ok = g >= 0 && (a.count == nullptr || a.n_in != a.p.pool || i < cnt_in);
            if (ok) { k = rr_make_key(a.dense[at(i, 4)], (uint32_t)g); if (k == 0ull) k = 1ull; ++n_valid_local; }
        }
        key[i] = k;
        val[i] = (unsigned)i;
    }
    __syncthreads();
    const int n_valid = block_reduce<int>(n_valid_local, [](int x, int y) { return x + y; }, red_i);
    sort_desc(key, val, n_pad_in);
    const int P = min(a.p.pool, n_valid);

    // ---- 1b. sharded input with local top-m < pool: is the merged pool provably the global pool? -------
    if (a.incomplete != nullptr) {
        __shared__ int s_incomplete;
        if (tid == 0) s_incomplete = 0;
        __syncthreads();
        // a shard that could not certify its dense result marks its tuples with global row -2
        for (int s = tid; s < a.n_shards; s += FUSE_THREADS)
            if (a.grow[at(s * per_shard, 8)] == -2) s_incomplete = 1;
        if (per_shard < a.p.pool) {
            const float cut = n_valid >= a.p.pool ? rr_key_score(key[P - 1]) : -INFINITY;
            for (int s = tid; s < a.n_shards; s += FUSE_THREADS) {
                int cnt = 0;
                float weakest = INFINITY;
                for (int j = 0; j < per_shard; ++j) {
                    const int i = s * per_shard + j;
                    if (a.grow[at(i, 8)] >= 0) { ++cnt; weakest = fminf(weakest, a.dense[at(i, 4)]); }
                }
                // a shard that sent everything it was allowed to may still hold rows scoring >= cut
                if (cnt == per_shard && weakest >= cut) s_incomplete = 1;
            }
        }
        __syncthreads();
        if (tid == 0) a.incomplete[b] = s_incomplete;
    }

    // ---- 2. load the pool -----------------------------------------------------------------------
    for (int i = tid; i < P; i += FUSE_THREADS) {
        const int s = (int)val[i];
        s_dense[i] = a.dense[at(s, 4)];
        s_bm25[i] = a.bm25 ? a.bm25[at(s, 4)] : 0.f;
        s_n[i] = a.n ? a.n[at(s, 8)] : 0.0;
        s_avg[i] = a.avg ? a.avg[at(s, 8)] : nan64();
        // optional per-candidate columns: tuple fields (sharded) or pool-order arrays
        const long long e = a.extras_by_slot ? at(s, 4) : (long long)b * a.p.pool + i;
        s_best[i] = a.best ? a.best[e] : 0.f;
        s_gate[i] = a.gate ? a.gate[e] : 1.0f;
        // the sorted keys are consumed (1b read key[P - 1] before the barrier above): key[] keeps the global row of every
        // pool position from here on (val is reused by the second sort)
        key[i] = (unsigned long long)a.grow[at(s, 8)];
    }
    __syncthreads();

    // ---- 3. min-max of dense and bm25 (float32 semantics) ----------------------------------------
    // raw best-review similarities (pool order) get the same float32 min-max (:294, app/test.py:288)
    float dmm_lo, dmm_div, bmm_lo, bmm_div, emm_lo = 0.f, emm_div = 1.f;
    bool d_zero, b_zero, e_zero = false;
    const bool best_raw = a.best != nullptr && a.p.best_is_raw;
    {
        const float* const cols[3] = {s_dense, s_bm25, s_best};
        float lo[3], hi[3];
        int bad;
        block_minmax<3>(cols, P, lo, hi, bad, red_f, red_i);
        auto finish = [&](int v, float& mm_lo, float& mm_div) -> bool {
            const double dl = (double)lo[v], dh = (double)hi[v];
            mm_lo = lo[v]; mm_div = (float)(dh - dl + 1e-12);
            return ((bad >> v) & 1) || isinf(lo[v]) || isinf(hi[v]) || (dh - dl < 1e-12) || P == 0;
        };
        d_zero = finish(0, dmm_lo, dmm_div);
        b_zero = finish(1, bmm_lo, bmm_div);
        if (best_raw) e_zero = finish(2, emm_lo, emm_div);
    }

    // ---- 4. g = nanmean(avg) over the pool, NumPy pairwise order -----------------------------------
    if (tid == 0) { int c = 0; pw_leaves(0, P, s_leaf_lo, s_leaf_n, c); s_nleaf = c; }
    __syncthreads();
    const int nleaf = min(s_nleaf, FUSE_MAX_LEAVES);
    for (int w = tid; w < nleaf * 8; w += FUSE_THREADS) {
        const int leaf = w >> 3, j = w & 7;
        const int lo = s_leaf_lo[leaf], n = s_leaf_n[leaf];
        double r = 0.0;
        if (n >= 8) {
            const double* src = s_avg + lo;
            double x = src[j]; r = (x != x) ? 0.0 : x;
            for (int i = 8; i < n - (n % 8); i += 8) { x = src[i + j]; r += (x != x) ? 0.0 : x; }
        }
        s_leaf_r[w] = r;
    }
    __syncthreads();
    for (int leaf = tid; leaf < nleaf; leaf += FUSE_THREADS) {
        const int lo = s_leaf_lo[leaf], n = s_leaf_n[leaf];
        const double* src = s_avg + lo;
        double res;
        if (n < 8) {
            res = 0.0;
            for (int i = 0; i < n; ++i) { const double x = src[i]; res += (x != x) ? 0.0 : x; }
        } else {
            const double* r = s_leaf_r + leaf * 8;
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (int i = n - (n % 8); i < n; ++i) { const double x = src[i]; res += (x != x) ? 0.0 : x; }
        }
        s_leaf_sum[leaf] = res;
    }
    int nn_local = 0;
    for (int i = tid; i < P; i += FUSE_THREADS) nn_local += (s_avg[i] == s_avg[i]) ? 1 : 0;
    const int n_notnan = block_reduce<int>(nn_local, [](int x, int y) { return x + y; }, red_i);
    if (tid == 0) {
        int next = 0;
        const double tot = P > 0 ? pw_combine(P, s_leaf_sum, next) : 0.0;
        s_g = (n_notnan > 0) ? tot / (double)n_notnan : nan64();
    }
    __syncthreads();
    const double g = s_g;

    // ---- 5. prior rating, its min-max (float64), volume -------------------------------------------
    double pl = INFINITY, ph = -INFINITY, lmax = -INFINITY; int pbad = 0, lbad = 0;
    for (int i = tid; i < P; i += FUSE_THREADS) {
        const double n = s_n[i], av = s_avg[i];
        const double pr = __ddiv_rn(__dadd_rn(__dmul_rn(av, n), __dmul_rn(g, a.p.prior_C)),
                                    __dadd_rn(__dadd_rn(n, a.p.prior_C), 1e-9));
        s_prior[i] = pr;
        if (pr != pr) pbad = 1;
        pl = fmin(pl, pr); ph = fmax(ph, pr);
        const double l1 = log1p(n);
        s_avg[i] = l1;                   // avg is not read again: the blend (6) takes log1p(n) from here
        if (l1 != l1) lbad = 1;
        lmax = fmax(lmax, l1);
    }
    {
        // one combined reduction: min(pl), max(ph), max(lmax), or(pbad | lbad << 1)
        int flags = pbad | (lbad << 1);
        for (int o = 16; o > 0; o >>= 1) {
            pl = fmin(pl, __shfl_xor_sync(0xffffffffu, pl, o));
            ph = fmax(ph, __shfl_xor_sync(0xffffffffu, ph, o));
            lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        }
        __syncthreads();
        if ((tid & 31) == 0) { const int w = tid >> 5; red_d[w] = pl; red_d[8 + w] = ph; red_d[16 + w] = lmax; red_i[w] = flags; }
        __syncthreads();
        const int g8 = tid & 7;
        pl = red_d[g8]; ph = red_d[8 + g8]; lmax = red_d[16 + g8]; flags = red_i[g8];
        for (int o = 4; o > 0; o >>= 1) {
            pl = fmin(pl, __shfl_xor_sync(0xffffffffu, pl, o));
            ph = fmax(ph, __shfl_xor_sync(0xffffffffu, ph, o));
            lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        }
        pbad = flags & 1; lbad = (flags >> 1) & 1;
    }
    if (lbad) lmax = nan64();
    const bool p_zero = pbad || isinf(pl) || isinf(ph) || (ph - pl < 1e-12) || P == 0;
    const double p_div = ph - pl + 1e-12;
    const double vol_div = lmax + 1e-9;

    // ---- 6. blend ------------------------------------------------------------------------------
    const float wd32 = (float)a.p.w_dense, wb32 = (float)a.p.w_bm25, wr32 = (float)a.p.w_rerank,
                wbest32 = (float)a.p.w_best;
    const double trust_den = (double)max(a.p.min_reviews, 1);
    const double sat_den = log1p((double)max(a.p.saturation, 1));
    const long long ex0 = (long long)b * a.p.pool;      // extras are given in pool order
    for (int i = tid; i < P; i += FUSE_THREADS) {
        const float dm = d_zero ? 0.f : __fdiv_rn(__fsub_rn(s_dense[i], dmm_lo), dmm_div);
        const float bm = b_zero ? 0.f : __fdiv_rn(__fsub_rn(s_bm25[i], bmm_lo), bmm_div);
        const float pm = p_zero ? 0.f : (float)__ddiv_rn(__dadd_rn(s_prior[i], -pl), p_div);
        const double n = s_n[i];
        const double l1 = s_avg[i];      // log1p(n), computed in (5)
        const double vol = __ddiv_rn(l1, vol_div);
        const double prior = __dadd_rn((double)__fmul_rn(pm, 0.7f), __dmul_rn(0.3, vol));

        float acc32 = __fmul_rn(wd32, dm);
        double acc64 = 0.0;
        bool is64 = false;
        // bm25 term
        if (a.p.bm25_is_f64_zero) { acc64 = __dadd_rn((double)acc32, __dmul_rn(a.p.w_bm25, 0.0)); is64 = true; }
        else acc32 = __fadd_rn(acc32, __fmul_rn(wb32, bm));
        // rerank term
        if (a.p.rerank_is_f32) {
            const float rr = a.rerank ? a.rerank[ex0 + i] : 0.f;
            const float t = __fmul_rn(wr32, rr);
            if (is64) acc64 = __dadd_rn(acc64, (double)t); else acc32 = __fadd_rn(acc32, t);
        } else {
            if (!is64) { acc64 = (double)acc32; is64 = true; }
            acc64 = __dadd_rn(acc64, __dmul_rn(a.p.w_rerank, 0.0));
        }
        // prior term (always float64)
        if (!is64) { acc64 = (double)acc32; is64 = true; }
        acc64 = __dadd_rn(acc64, __dmul_rn(a.p.w_prior, prior));
        // best-review term (float32 column)
        float be;
        {
            be = s_best[i];
            if (best_raw) be = e_zero ? 0.f : __fdiv_rn(__fsub_rn(be, emm_lo), emm_div);
            acc64 = __dadd_rn(acc64, (double)__fmul_rn(wbest32, be));
        }
        float fin = (float)acc64;
        const double ramp = fmin(fmax(__ddiv_rn(n, trust_den), 0.0), 1.0);
        const double satv = fmin(1.0, __ddiv_rn(l1, sat_den));
        const float trust = (float)__dadd_rn(__dmul_rn(0.6, ramp), __dmul_rn(0.4, satv));
        const float gate = s_gate[i];
        if (a.p.use_trust) fin = __fmul_rn(fin, trust);
        fin = __fmul_rn(fin, gate);
        s_final[i] = fin;
        if (a.components) {
            float* c = a.components + ((long long)b * a.p.pool + i) * 8;
            c[0] = dm; c[1] = bm; c[2] = (float)prior; c[3] = trust; c[4] = fin;
            c[5] = s_dense[i]; c[6] = s_bm25[i]; c[7] = be;
        }
    }
    if (a.components) {
        for (int i = P + tid; i < a.p.pool; i += FUSE_THREADS) {
            float* c = a.components + ((long long)b * a.p.pool + i) * 8;
            for (int j = 0; j < 8; ++j) c[j] = 0.f;
        }
    }
    __syncthreads();

    // ---- 7. order by (final desc, pool position asc); NaN last --------------------------------------
    // global rows were stashed in key[0..P); move them behind the sort area first
    long long* s_rows = reinterpret_cast<long long*>(s_prior);   // prior no longer needed (8 B each)
    for (int i = tid; i < P; i += FUSE_THREADS) s_rows[i] = (long long)key[i];
    __syncthreads();
    for (int i = tid; i < n_pad_pool; i += FUSE_THREADS) {
        unsigned long long k = 0ull;
        if (i < P) { k = rr_make_key(s_final[i], (uint32_t)i); if (k == 0ull) k = 1ull; }
        key[i] = k;
        val[i] = (unsigned)i;
    }
    __syncthreads();
    sort_desc(key, val, n_pad_pool);
    for (int i = tid; i < a.p.k; i += FUSE_THREADS) {
        const long long o = (long long)b * a.p.k + i;
        if (i < P) {
            const int pos = (int)val[i];
            a.top_row[o] = s_rows[pos];
            a.top_final[o] = s_final[pos];
            if (a.top_pos) a.top_pos[o] = pos;
        } else {
            a.top_row[o] = -1;
            a.top_final[o] = __uint_as_float(0x7FC00000u);
            if (a.top_pos) a.top_pos[o] = -1;
        }
    }
}

int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

}  // namespace

int rr_launch_fuse(const rr_fusion_params* p, int B, int n_in, int n_shards, int64_t shard_stride_bytes,
                   const int32_t* d_count, const float* d_dense,
                   const float* d_bm25, const double* d_n, const double* d_avg, const int64_t* d_grow,
                   const float* d_rerank, const float* d_best, const float* d_gate, int64_t* d_top_row,
                   float* d_top_final, int32_t* d_top_pos, float* d_components, int32_t* d_incomplete,
                   cudaStream_t stream, int extras_by_slot) {
    if (!p || B < 0 || !d_dense || !d_grow || !d_top_row || !d_top_final)
        return rr_fail(RR_EINVAL, "rr_fuse_topk: null argument");
    if (p->pool <= 0 || p->pool > FUSE_MAX_POOL) return rr_fail(RR_EINVAL, "rr_fuse_topk: pool must be in 1..%d", FUSE_MAX_POOL);
    if (n_in <= 0 || n_in > FUSE_MAX_IN) return rr_fail(RR_EINVAL, "rr_fuse_topk: n_in must be in 1..%d", FUSE_MAX_IN);
    if (p->k <= 0) return rr_fail(RR_EINVAL, "rr_fuse_topk: k must be positive");
    if ((d_rerank || ((d_best || d_gate) && !extras_by_slot)) && n_in != p->pool)
        return rr_fail(RR_EINVAL, "rr_fuse_topk: rerank/best/gate are given in pool order and need n_in == pool");
    if (B == 0) return RR_OK;
    FuseArgs a;
    a.p = *p; a.B = B; a.n_in = n_in; a.n_shards = n_shards > 0 ? n_shards : 1; a.shard_stride = shard_stride_bytes;
    if (n_in % a.n_shards) return rr_fail(RR_EINVAL, "rr_fuse_topk: n_in must be a multiple of n_shards");
    a.count = d_count; a.dense = d_dense; a.bm25 = d_bm25; a.n = d_n; a.avg = d_avg;
    a.grow = reinterpret_cast<const long long*>(d_grow); a.rerank = d_rerank; a.best = d_best; a.gate = d_gate;
    a.top_row = reinterpret_cast<long long*>(d_top_row); a.top_final = d_top_final; a.top_pos = d_top_pos;
    a.components = d_components;
    a.incomplete = d_incomplete;
    a.extras_by_slot = extras_by_slot;
    const int n_pad_in = next_pow2(n_in), n_pad_pool = next_pow2(p->pool);
    const int n_pad = n_pad_in > n_pad_pool ? n_pad_in : n_pad_pool;
    const size_t smem = (size_t)n_pad * 8 + (size_t)(n_pad + (n_pad & 1)) * 4 + (size_t)p->pool * (8 * 3 + 4 * 5) + 16;
    // ~21 KB of static shared memory on top: opt in for the largest request seen so far on this device (set once per
    // size class, never while a stream of this thread is being captured into a graph by a warm caller)
    static RrSmemOptIn optin;
    int dev = 0;
    const size_t want = smem > 24 * 1024 ? smem : 1;
    if (optin.needed(want, &dev)) {
        if (smem > 24 * 1024)
            RR_CUDA(cudaFuncSetAttribute(fuse_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RR_CUDA(cudaFuncSetAttribute(fuse_topk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        optin.done(want, dev);
    }
    {
        RrProfScope prof(RR_PROF_FUSE, stream);
        fuse_topk_kernel<<<B, FUSE_THREADS, smem, stream>>>(a, n_pad_in, n_pad_pool);
    }
    RR_LAUNCH_CHECK();
    return RR_OK;
}
