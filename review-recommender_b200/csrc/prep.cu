// Index preparation on the GPU: row normalisation of product_emb.npy and the bf16 copy K2 reads.
//
// Reference: l2_normalize utils.py:40-44 (= _l2norm app/app_product_search.py:179-180, app/test.py:109-112),
// applied once at load (app/app_product_search.py:110, app/test.py:145):
//
//     n = np.linalg.norm(x, axis=1, keepdims=True)        # sqrt(add.reduce(x*x, axis=1)) in float32
//     return x / np.maximum(n, 1e-12)
//
// Bit-exact restatement: the squares are summed in NumPy's float32 PAIRWISE order (blocks of <= 128 elements
// with 8 strided accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), halves split at a multiple of 8),
// then sqrt, max and the division are the correctly rounded float32 operations.  One warp per row: the row is
// staged in shared memory with coalesced loads, the (block, accumulator) chains run one per lane, lane 0 combines
// the block sums in NumPy's recursion order.  HBM-bound: 4*D bytes read, 4*D (+ 2*dim_pad) bytes written per row.
#include "rr_internal.h"
#include "rr_kernels.h"

#include <cuda_bf16.h>

namespace {

constexpr int PREP_WARPS = 8;
constexpr int PREP_MAX_LEAVES = 64;          // rows up to 8192 elements

struct PwPlan {
    int n_leaves;
    int lo[PREP_MAX_LEAVES];
    int n[PREP_MAX_LEAVES];
};

void plan_leaves(int lo, int n, PwPlan& p) {
    if (n <= 128) {
        if (p.n_leaves < PREP_MAX_LEAVES) { p.lo[p.n_leaves] = lo; p.n[p.n_leaves] = n; }
        ++p.n_leaves;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    plan_leaves(lo, n2, p);
    plan_leaves(lo + n2, n - n2, p);
}

__device__ float pw_combine_f32(int n, const float* leaf_sum, int& next) {
    if (n <= 128) return leaf_sum[next++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const float a = pw_combine_f32(n2, leaf_sum, next);
    const float b = pw_combine_f32(n - n2, leaf_sum, next);
    return __fadd_rn(a, b);
}

__global__ void __launch_bounds__(PREP_WARPS * 32)
normalize_rows_kernel(const float* in, long long n_rows, int D, const PwPlan plan, float eps,
                      float* out_f32, __nv_bfloat16* __restrict__ out_bf16, int dim_pad,
                      float* __restrict__ out_norm) {
    extern __shared__ float smem_prep[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * PREP_WARPS + wib;
    if (row >= n_rows) return;
    const int n_chain = plan.n_leaves * 8;
    float* x = smem_prep + (size_t)wib * (D + n_chain + plan.n_leaves);
    float* chain = x + D;
    float* leaf_sum = chain + n_chain;
    const float* src = in + row * D;
    for (int i = lane; i < D; i += 32) x[i] = src[i];
    __syncwarp();
    // one (block, accumulator) chain per lane: r_j = sq[j] + sq[8+j] + sq[16+j] + ... (ascending)
    for (int c = lane; c < n_chain; c += 32) {
        const int leaf = c >> 3, j = c & 7;
        const int lo = plan.lo[leaf], n = plan.n[leaf];
        float r = 0.f;
        if (n >= 8) {
            const float* a = x + lo;
            r = __fmul_rn(a[j], a[j]);
            for (int i = 8; i < n - (n % 8); i += 8) r = __fadd_rn(r, __fmul_rn(a[i + j], a[i + j]));
        }
        chain[c] = r;
    }
    __syncwarp();
    for (int leaf = lane; leaf < plan.n_leaves; leaf += 32) {
        const int lo = plan.lo[leaf], n = plan.n[leaf];
        const float* a = x + lo;
        float res;
        int i;
        if (n < 8) {
            res = 0.f;
            i = 0;
        } else {
            const float* r = chain + leaf * 8;
            res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                            __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
            i = n - (n % 8);
        }
        for (; i < n; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
        leaf_sum[leaf] = res;
    }
    __syncwarp();
    float denom = 0.f;
    if (lane == 0) {
        int next = 0;
        const float total = __fadd_rn(0.f, pw_combine_f32(D, leaf_sum, next));
        const float nrm = __fsqrt_rn(total);
        denom = fmaxf(nrm, eps);                       // np.maximum(n, 1e-12): eps is a weak scalar -> float32
        if (out_norm) out_norm[row] = nrm;
    }
    denom = __shfl_sync(0xffffffffu, denom, 0);
    float* dst = out_f32 ? out_f32 + row * D : nullptr;
    __nv_bfloat16* dst16 = out_bf16 ? out_bf16 + row * dim_pad : nullptr;
    for (int i = lane; i < (dst16 ? dim_pad : D); i += 32) {
        const float v = i < D ? __fdiv_rn(x[i], denom) : 0.f;
        if (dst && i < D) dst[i] = v;
        if (dst16) dst16[i] = __float2bfloat16_rn(v);
    }
}

// round-to-nearest-even bf16 copy of rows that are already normalised, zero-padded to dim_pad columns
__global__ void __launch_bounds__(256)
bf16_rows_kernel(const float* __restrict__ in, long long n_rows, int D, __nv_bfloat16* __restrict__ out, int dim_pad) {
    const long long total = n_rows * dim_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim_pad;
        const int c = (int)(i - r * dim_pad);
        out[i] = __float2bfloat16_rn(c < D ? in[r * D + c] : 0.f);
    }
}

}  // namespace

// max over the rows of ||row||_2 (the bf16 error bound of the tensor path scales with it): one warp per row, float
// accumulation, atomicMax on the (non-negative) float bits.  *d_out must be zeroed by the caller.
__global__ void __launch_bounds__(256)
max_row_norm_kernel(const float* __restrict__ x, long long n_rows, int D, unsigned* __restrict__ out_bits) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    float best = 0.f;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = x[r * D + d]; ss = fmaf(v, v, ss); }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        best = fmaxf(best, sqrtf(ss));
    }
    if (lane == 0 && best > 0.f) atomicMax(out_bits, __float_as_uint(best));
}

int rr_launch_max_row_norm(const float* d_in, int64_t n_rows, int D, float* d_out, int sm_count, cudaStream_t stream) {
    RR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), stream));
    if (n_rows <= 0) return RR_OK;
    RrProfScope prof(RR_PROF_MISC, stream);
    max_row_norm_kernel<<<sm_count * 8, 256, 0, stream>>>(d_in, (long long)n_rows, D, reinterpret_cast<unsigned*>(d_out));
    RR_LAUNCH_CHECK();
    return RR_OK;
}

int rr_launch_bf16_rows(const float* d_in, int64_t n_rows, int D, uint16_t* d_out, int dim_pad, int sm_count,
                        cudaStream_t stream) {
    if (n_rows <= 0) return RR_OK;
    RrProfScope prof(RR_PROF_MISC, stream);
    bf16_rows_kernel<<<sm_count * 8, 256, 0, stream>>>(d_in, (long long)n_rows, D, reinterpret_cast<__nv_bfloat16*>(d_out), dim_pad);
    RR_LAUNCH_CHECK();
    return RR_OK;
}

int rr_launch_normalize_rows(const float* d_in, int64_t n_rows, int D, float* d_out_f32, uint16_t* d_out_bf16,
                             int dim_pad, float* d_norms, cudaStream_t stream) {
    if (n_rows <= 0) return RR_OK;
    PwPlan plan{};
    plan_leaves(0, D, plan);
    if (plan.n_leaves > PREP_MAX_LEAVES) return rr_fail(RR_EUNSUPPORTED, "rr_normalize_rows: dim %d too large", D);
    const size_t smem = sizeof(float) * (size_t)PREP_WARPS * (D + plan.n_leaves * 9);
    if (smem > 200 * 1024) return rr_fail(RR_EUNSUPPORTED, "rr_normalize_rows: dim %d too large", D);
    static RrSmemOptIn optin;
    int dev = 0;
    if (smem > 48 * 1024 && optin.needed(smem, &dev)) {
        RR_CUDA(cudaFuncSetAttribute(normalize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin.done(smem, dev);
    }
    RrProfScope prof(RR_PROF_MISC, stream);
    normalize_rows_kernel<<<(unsigned)((n_rows + PREP_WARPS - 1) / PREP_WARPS), PREP_WARPS * 32, smem, stream>>>(
        d_in, (long long)n_rows, D, plan, 1e-12f, d_out_f32, reinterpret_cast<__nv_bfloat16*>(d_out_bf16), dim_pad, d_norms);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
