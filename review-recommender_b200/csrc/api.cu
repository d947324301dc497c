// C ABI of librr_b200.so: handle management, scratch, and the call sequences of the hot path.
// See include/rr_b200.h for the contract of every entry point.
#include "rr_internal.h"
#include "rr_kernels.h"
#include "dense_tc.h"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

// ---------------------------------------------------------------------------------------------
// errors / counters
// ---------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "";
std::atomic<int64_t> g_rr_launches{0};

int rr_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* rr_last_error(void) { return t_err; }
extern "C" int rr_abi_version(void) { return 3; }
// sizeof of the three structs that cross the boundary, so that a binding can verify its own layout at load time
extern "C" void rr_struct_sizes(int32_t* out3) {
    if (!out3) return;
    out3[0] = (int32_t)sizeof(rr_index_desc);
    out3[1] = (int32_t)sizeof(rr_fusion_params);
    out3[2] = (int32_t)sizeof(rr_dense_stats);
}
extern "C" int64_t rr_launch_count(int reset) {
    return reset ? g_rr_launches.exchange(0) : g_rr_launches.load();
}

// ---------------------------------------------------------------------------------------------
// per-class kernel timing (bench.py): CUDA events on the launching stream
// ---------------------------------------------------------------------------------------------
namespace {
struct ProfRec { int cls; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::atomic<int> g_prof_on{0};
std::vector<ProfRec*> g_prof_live, g_prof_free;
}  // namespace

RrProfScope::RrProfScope(int cls_, cudaStream_t s) : cls(cls_), stream(s), rec(nullptr) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec* r = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        if (!g_prof_free.empty()) { r = g_prof_free.back(); g_prof_free.pop_back(); }
    }
    if (!r) {
        r = new (std::nothrow) ProfRec();
        if (!r) return;
        if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    }
    r->cls = cls;
    cudaEventRecord(r->a, stream);
    rec = r;
}
RrProfScope::~RrProfScope() {
    if (!rec) return;
    ProfRec* r = static_cast<ProfRec*>(rec);
    cudaEventRecord(r->b, stream);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_live.push_back(r);
}

extern "C" int rr_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); return RR_OK; }

extern "C" int rr_profile_collect(double* h_ms, int64_t* h_launches, int32_t n_classes) {
    if (!h_ms || !h_launches || n_classes < RR_PROF_CLASSES) return rr_fail(RR_EINVAL, "rr_profile_collect: need %d classes", RR_PROF_CLASSES);
    for (int i = 0; i < n_classes; ++i) { h_ms[i] = 0.0; h_launches[i] = 0; }
    std::vector<ProfRec*> live;
    {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        live.swap(g_prof_live);
    }
    for (ProfRec* r : live) {
        float ms = 0.f;
        if (cudaEventSynchronize(r->b) == cudaSuccess && cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) {
            h_ms[r->cls] += ms;
            h_launches[r->cls] += 1;
        }
    }
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (ProfRec* r : live) g_prof_free.push_back(r);
    return RR_OK;
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_buf_generation{1};     // bumped whenever any scratch buffer moves (captured graphs hold pointers)

struct DeviceBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RR_OK;
        g_buf_generation.fetch_add(1);
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return rr_fail(RR_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return RR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct rr_index {
    rr_index_desc d;
    int device = 0;
    int sm_count = 148;
    int cc_major = 0, cc_minor = 0;
    std::mutex mu;
    DeviceBuf scores;     // exact path: [Bchunk, ld] fp32
    DeviceBuf select;     // radix-select state
    DeviceBuf tuples;     // hybrid: candidate tuples
    DeviceBuf staging;    // *_host entry points
    DeviceBuf rtab;       // rr_bm25_get_scores: tile bounds of the batch's rare terms
    rr_tc_state* tc = nullptr;
    rr_dense_stats stats{};
    // the scratch buffers above are shared by all callers of the handle: a call that uses them on another stream
    // than the previous one first waits for the previous call's work (the mutex only serialises the host side)
    cudaEvent_t fence = nullptr;
    cudaStream_t fence_stream = nullptr;
    bool fence_armed = false;
    // small-batch host searches replay a captured CUDA graph (one launch instead of ~20 latency-bound ones)
    struct GraphEntry { uint64_t key; uint64_t generation; int seen; cudaGraphExec_t exec; };
    std::vector<GraphEntry> graphs;
    cudaStream_t capture_stream = nullptr;
    void* h_pinned = nullptr;          // small-batch host searches: one packed H2D and one packed D2H through pinned memory
    size_t h_pinned_cap = 0;
};

namespace {
struct ScratchFence {
    rr_index* ix;
    cudaStream_t s;
    ScratchFence(rr_index* ix_, cudaStream_t s_) : ix(ix_), s(s_) {
        if (!ix->fence) cudaEventCreateWithFlags(&ix->fence, cudaEventDisableTiming);
        if (ix->fence && ix->fence_armed && ix->fence_stream != s) cudaStreamWaitEvent(s, ix->fence, 0);
    }
    ~ScratchFence() {
        if (ix->fence && cudaEventRecord(ix->fence, s) == cudaSuccess) { ix->fence_stream = s; ix->fence_armed = true; }
    }
};
}  // namespace

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
    char* p;
    template <class T> T* take(size_t n) {
        T* r = reinterpret_cast<T*>(p);
        p += align_up(sizeof(T) * n, 256);
        return r;
    }
};
}  // namespace

extern "C" int rr_index_create(rr_index** out, const rr_index_desc* desc, int device) {
    if (!out || !desc) return rr_fail(RR_EINVAL, "rr_index_create: null argument");
    if (desc->n_docs <= 0 || desc->dim <= 0 || !desc->d_emb_f32)
        return rr_fail(RR_EINVAL, "rr_index_create: n_docs, dim and d_emb_f32 are required");
    if (desc->n_docs > 0xFFFFFFF0ll) return rr_fail(RR_EINVAL, "rr_index_create: more than 2^32 rows per shard");
    if (desc->d_emb_bf16 && (desc->dim_pad < desc->dim || desc->dim_pad % 64))
        return rr_fail(RR_EINVAL, "rr_index_create: dim_pad must be a multiple of 64 and >= dim");
    if (desc->vocab_size > 0) {
        if (!desc->d_postings || !desc->d_tile_base || !desc->d_dir || !desc->d_term_slot || !desc->d_rare_off ||
            desc->n_freq < 0 || desc->tile_docs <= 0 || (desc->tile_docs & 3) ||
            desc->n_tiles != (int32_t)((desc->n_docs + desc->tile_docs - 1) / desc->tile_docs))
            return rr_fail(RR_EINVAL, "rr_index_create: inconsistent BM25 postings description");
        if ((size_t)desc->tile_docs * 4 > 200 * 1024)
            return rr_fail(RR_EINVAL, "rr_index_create: tile_docs too large for shared-memory accumulators");
    }
    RR_CUDA(cudaSetDevice(device));
    rr_index* ix = new (std::nothrow) rr_index();
    if (!ix) return rr_fail(RR_ENOMEM, "rr_index_create: out of host memory");
    ix->d = *desc;
    ix->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete ix; return rr_fail(RR_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    ix->sm_count = prop.multiProcessorCount;
    ix->cc_major = prop.major;
    ix->cc_minor = prop.minor;
    *out = ix;
    return RR_OK;
}

extern "C" void rr_index_destroy(rr_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    ix->scores.release();
    ix->select.release();
    ix->tuples.release();
    ix->staging.release();
    ix->rtab.release();
    rr_tc_destroy(ix->tc);
    for (auto& g : ix->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ix->capture_stream) cudaStreamDestroy(ix->capture_stream);
    if (ix->h_pinned) cudaFreeHost(ix->h_pinned);
    if (ix->fence) cudaEventDestroy(ix->fence);
    delete ix;
}

extern "C" int rr_dense_last_stats(rr_index* ix, rr_dense_stats* out) {
    if (!ix || !out) return rr_fail(RR_EINVAL, "rr_dense_last_stats: null argument");
    *out = ix->stats;
    return RR_OK;
}

// ---------------------------------------------------------------------------------------------
// BM25
// ---------------------------------------------------------------------------------------------
extern "C" int rr_bm25_get_scores(rr_index* ix, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                                  int32_t l_max, float* d_out, int64_t ld_out, rr_stream stream) {
    if (!ix || B < 0 || l_max < 0) return rr_fail(RR_EINVAL, "rr_bm25_get_scores: bad argument");
    if (B == 0) return RR_OK;                           // an empty batch has no output buffer to check
    if (!d_out) return rr_fail(RR_EINVAL, "rr_bm25_get_scores: bad argument");
    if (ld_out < ix->d.n_docs || (ld_out & 3) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return rr_fail(RR_EINVAL, "rr_bm25_get_scores: ld_out must be >= n_docs and a multiple of 4, d_out 16-byte aligned");
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ix->d.vocab_size <= 0 || l_max == 0 || !d_term_ids || !d_n_terms) {
        // "BM25 absent" / no tokens -> zeros (app/app_product_search.py:202-204)
        RR_CUDA(cudaMemset2DAsync(d_out, sizeof(float) * (size_t)ld_out, 0, sizeof(float) * (size_t)ix->d.n_docs, (size_t)B, s));
        return RR_OK;
    }
    // per batch: where every tile starts in the lists of the batch's rare terms (scratch of the handle, <= 256 MB a go)
    ScratchFence fence(ix, s);
    const size_t per_q = rr_bm25_rtab_bytes(1, l_max, ix->d.n_tiles);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)256 << 20) / std::max<size_t>(per_q, 1)));
    {
        const void* before = ix->rtab.p;
        RR_TRY(ix->rtab.ensure(256 + per_q * (size_t)chunk));     // [work counters | rare-bound table]
        if (ix->rtab.p != before) RR_CUDA(cudaMemsetAsync(ix->rtab.p, 0, 256, s));   // the kernel re-arms them itself afterwards
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        RR_TRY(rr_launch_bm25_tile_scores(&ix->d, d_term_ids + (int64_t)b0 * l_max, d_n_terms + b0, nb, l_max,
                                          d_out + (int64_t)b0 * ld_out, ld_out,
                                          reinterpret_cast<uint32_t*>(static_cast<char*>(ix->rtab.p) + 256),
                                          static_cast<unsigned*>(ix->rtab.p), s));
    }
    return RR_OK;
}

static int candidates_locked(rr_index* ix, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                             int32_t l_max, const int64_t* d_cand, int32_t pool, float* d_bm25, double* d_n,
                             double* d_avg, int64_t* d_grow, cudaStream_t s) {
    const bool have = ix->d.vocab_size > 0 && l_max > 0 && d_term_ids && d_n_terms;
    return rr_launch_bm25_candidates(&ix->d, have ? ix->d.vocab_size : 0, have ? d_term_ids : nullptr, d_n_terms, B, l_max,
                                     d_cand, pool, d_bm25, d_n, d_avg, d_grow, s);
}

extern "C" int rr_bm25_candidates(rr_index* ix, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                                  int32_t l_max, const int64_t* d_cand, int32_t pool, float* d_out, rr_stream stream) {
    if (!ix || !d_cand || !d_out || B < 0 || pool <= 0) return rr_fail(RR_EINVAL, "rr_bm25_candidates: bad argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    return candidates_locked(ix, d_term_ids, d_n_terms, B, l_max, d_cand, pool, d_out, nullptr, nullptr, nullptr,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int rr_candidate_tuples(rr_index* ix, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                                   int32_t l_max, const int64_t* d_cand, int32_t pool, float* d_bm25,
                                   double* d_n_reviews, double* d_avg_stars, int64_t* d_global_row, rr_stream stream) {
    if (!ix || !d_cand || B < 0 || pool <= 0) return rr_fail(RR_EINVAL, "rr_candidate_tuples: bad argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    return candidates_locked(ix, d_term_ids, d_n_terms, B, l_max, d_cand, pool, d_bm25, d_n_reviews, d_avg_stars,
                             d_global_row, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// dense
// ---------------------------------------------------------------------------------------------
static int dense_exact_locked(rr_index* ix, const float* d_q, int32_t B, int32_t pool, int64_t* d_idx,
                              float* d_sims, int32_t* d_count, cudaStream_t s) {
    const int64_t n = ix->d.n_docs;
    const int64_t ld = (int64_t)align_up((size_t)n, 4);
    const size_t budget = (size_t)1 << 30;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>(64, budget / (sizeof(float) * (size_t)ld)));
    chunk = std::min(chunk, B);
    if (chunk > 8) chunk = chunk / 8 * 8;
    RR_TRY(ix->scores.ensure(sizeof(float) * (size_t)ld * chunk));
    RR_TRY(ix->select.ensure(rr_exact_scratch_bytes(chunk, (int)std::min<int64_t>(pool, n), n, ix->sm_count)));
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        RR_TRY(rr_launch_dense_scores_f32(ix->d.d_emb_f32, n, ix->d.dim, d_q + (int64_t)b0 * ix->d.dim, nb,
                                          static_cast<float*>(ix->scores.p), ld, ix->sm_count, s));
        RR_TRY(rr_launch_topk_rows(static_cast<const float*>(ix->scores.p), ld, n, nb, pool, ix->select.p,
                                   d_idx + (int64_t)b0 * pool, d_sims + (int64_t)b0 * pool,
                                   d_count ? d_count + b0 : nullptr, pool, ix->sm_count, s));
    }
    return RR_OK;
}

static int dense_topk_locked(rr_index* ix, const float* d_q, int32_t B, int32_t pool, int32_t mode,
                             int64_t* d_idx, float* d_sims, int32_t* d_count, cudaStream_t s,
                             int32_t* d_uncertified = nullptr) {
    if (mode == RR_DENSE_AUTO) {
        const bool tc_ok = rr_tc_supported(ix->cc_major, ix->cc_minor) && ix->d.d_emb_bf16 != nullptr &&
                           rr_tc_can_handle(ix->d.dim_pad, pool);
        mode = (tc_ok && B >= 32 && ix->d.n_docs >= 65536) ? RR_DENSE_TENSOR : RR_DENSE_EXACT;
    }
    if (mode == RR_DENSE_TENSOR) {
        if (!ix->d.d_emb_bf16) return rr_fail(RR_EUNSUPPORTED, "tensor path needs the bf16 corpus copy (d_emb_bf16)");
        if (!rr_tc_supported(ix->cc_major, ix->cc_minor))
            return rr_fail(RR_EUNSUPPORTED, "tensor path needs an sm_100 device (found sm_%d%d)", ix->cc_major, ix->cc_minor);
        // the tensor path takes at most 8 x SM query tiles per call: very large batches go through in slices
        static const int32_t slice = getenv("RR_TC_MAX_BATCH") && atoi(getenv("RR_TC_MAX_BATCH")) >= 128 ? atoi(getenv("RR_TC_MAX_BATCH")) : 65536;
        for (int32_t b0 = 0; b0 < B; b0 += slice) {
            const int32_t nb = std::min(slice, B - b0);
            RR_TRY(rr_tc_dense_topk(&ix->tc, &ix->d, ix->sm_count, d_q + (int64_t)b0 * ix->d.dim, nb, pool,
                                    d_idx + (int64_t)b0 * pool, d_sims + (int64_t)b0 * pool, d_count ? d_count + b0 : nullptr,
                                    &ix->stats,
                                    [](void* ctx, const float* q, int32_t b, int32_t p, int64_t* idx, float* sims, int32_t* cnt,
                                       cudaStream_t st) {
                                        return dense_exact_locked(static_cast<rr_index*>(ctx), q, b, p, idx, sims, cnt, st);
                                    },
                                    ix, d_uncertified ? d_uncertified + b0 : nullptr, s));
        }
        return RR_OK;
    }
    if (mode != RR_DENSE_EXACT) return rr_fail(RR_EINVAL, "rr_dense_topk: unknown mode %d", mode);
    ix->stats = rr_dense_stats{};
    ix->stats.path = 1;
    if (d_uncertified) RR_CUDA(cudaMemsetAsync(d_uncertified, 0, sizeof(int32_t) * (size_t)B, s));   // exact path: all proven
    return dense_exact_locked(ix, d_q, B, pool, d_idx, d_sims, d_count, s);
}

extern "C" int rr_dense_topk(rr_index* ix, const float* d_q, int32_t B, int32_t pool, int32_t mode,
                             int64_t* d_idx, float* d_sims, int32_t* d_count, rr_stream stream) {
    if (!ix || !d_q || !d_idx || !d_sims || B < 0 || pool <= 0) return rr_fail(RR_EINVAL, "rr_dense_topk: bad argument");
    if (pool > 8192) return rr_fail(RR_EINVAL, "rr_dense_topk: pool larger than 8192 is not supported");
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    return dense_topk_locked(ix, d_q, B, pool, mode, d_idx, d_sims, d_count, static_cast<cudaStream_t>(stream));
}

extern "C" int rr_dense_topk_deferred(rr_index* ix, const float* d_q, int32_t B, int32_t pool, int32_t mode,
                                      int64_t* d_idx, float* d_sims, int32_t* d_count, int32_t* d_uncertified,
                                      rr_stream stream) {
    if (!ix || !d_q || !d_idx || !d_sims || !d_uncertified || B < 0 || pool <= 0)
        return rr_fail(RR_EINVAL, "rr_dense_topk_deferred: bad argument");
    if (pool > 8192) return rr_fail(RR_EINVAL, "rr_dense_topk_deferred: pool larger than 8192 is not supported");
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    return dense_topk_locked(ix, d_q, B, pool, mode, d_idx, d_sims, d_count, static_cast<cudaStream_t>(stream), d_uncertified);
}

extern "C" int rr_dense_debug_bf16_scores(rr_index* ix, const float* d_q, int32_t B, int64_t row0, int32_t n_rows,
                                          float* d_out, rr_stream stream) {
    if (!ix || !d_q || !d_out) return rr_fail(RR_EINVAL, "rr_dense_debug_bf16_scores: null argument");
    if (!ix->d.d_emb_bf16) return rr_fail(RR_EUNSUPPORTED, "tensor path needs the bf16 corpus copy (d_emb_bf16)");
    if (!rr_tc_supported(ix->cc_major, ix->cc_minor))
        return rr_fail(RR_EUNSUPPORTED, "tensor path needs an sm_100 device (found sm_%d%d)", ix->cc_major, ix->cc_minor);
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    return rr_tc_debug_scores(&ix->tc, &ix->d, ix->sm_count, d_q, B, row0, n_rows, d_out, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// fusion
// ---------------------------------------------------------------------------------------------
extern "C" int rr_fuse_topk(const rr_fusion_params* p, int32_t B, int32_t n_in, const int32_t* d_count,
                            const float* d_dense, const float* d_bm25, const double* d_n_reviews,
                            const double* d_avg_stars, const int64_t* d_global_row, const float* d_rerank,
                            const float* d_best, const float* d_gate, int64_t* d_top_row, float* d_top_final,
                            int32_t* d_top_pos, float* d_components, int device, rr_stream stream) {
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_fuse(p, B, n_in, 1, 0, d_count, d_dense, d_bm25, d_n_reviews, d_avg_stars, d_global_row, d_rerank,
                          d_best, d_gate, d_top_row, d_top_final, d_top_pos, d_components, nullptr,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int rr_fuse_topk_sharded(const rr_fusion_params* p, int32_t B, int32_t n_shards, int32_t per_shard,
                                    int64_t shard_stride_bytes, const float* d_dense, const float* d_bm25,
                                    const double* d_n_reviews, const double* d_avg_stars, const int64_t* d_global_row,
                                    const float* d_gate, const float* d_best,
                                    int64_t* d_top_row, float* d_top_final, int32_t* d_incomplete, int device,
                                    rr_stream stream) {
    if (n_shards <= 0 || per_shard <= 0 || (shard_stride_bytes & 7))
        return rr_fail(RR_EINVAL, "rr_fuse_topk_sharded: bad shard geometry");
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_fuse(p, B, n_shards * per_shard, n_shards, shard_stride_bytes, nullptr, d_dense, d_bm25,
                          d_n_reviews, d_avg_stars, d_global_row, nullptr, d_best, d_gate, d_top_row, d_top_final,
                          nullptr, nullptr, d_incomplete, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int rr_shard_tuples(rr_index* ix, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                               int32_t B, int32_t l_max, int32_t m, int32_t dense_mode, int32_t n_ranks, void* d_send,
                               rr_stream stream) {
    if (!ix || !d_q || !d_send || B <= 0 || m <= 0 || n_ranks <= 0 || B % n_ranks)
        return rr_fail(RR_EINVAL, "rr_shard_tuples: bad argument (B must be a multiple of n_ranks)");
    if (m > 8192) return rr_fail(RR_EINVAL, "rr_shard_tuples: m larger than 8192 is not supported");
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t bm = (size_t)B * m;
    RR_TRY(ix->tuples.ensure(align_up(bm * 8, 256) + align_up(bm * 4, 256) + 2 * align_up((size_t)B * 4, 256)));
    Carver c{static_cast<char*>(ix->tuples.p)};
    int64_t* idx = c.take<int64_t>(bm);
    float* sims = c.take<float>(bm);
    int32_t* cnt = c.take<int32_t>(B);
    int32_t* uncert = c.take<int32_t>(B);
    RR_TRY(dense_topk_locked(ix, d_q, B, m, dense_mode, idx, sims, cnt, s, uncert));
    // packed exchange layout: per destination rank [rows i64 | n f64 | avg f64 | dense f32 | bm25 f32] x (B/n_ranks * m)
    const int bg = B / n_ranks;
    const size_t bp = (size_t)bg * m;
    char* base = static_cast<char*>(d_send);
    const bool have = ix->d.vocab_size > 0 && l_max > 0 && d_term_ids && d_n_terms;
    return rr_launch_bm25_candidates(&ix->d, have ? ix->d.vocab_size : 0, have ? d_term_ids : nullptr, d_n_terms, B, l_max,
                                     idx, m, reinterpret_cast<float*>(base + bp * 28),
                                     reinterpret_cast<double*>(base + bp * 8), reinterpret_cast<double*>(base + bp * 16),
                                     reinterpret_cast<int64_t*>(base), s, bg, (int64_t)(bp * 32), sims,
                                     reinterpret_cast<float*>(base + bp * 24), uncert);
}

extern "C" int rr_best_review_scores(const float* d_rev_emb, const int64_t* d_rev_range, int64_t n_products, int32_t dim,
                                     const float* d_q, int32_t B, const int64_t* d_cand, int32_t pool,
                                     const int64_t* d_slot_file, const int64_t* d_limit,
                                     float* d_best_score, int64_t* d_best_slot, int device, rr_stream stream) {
    if (!d_rev_emb || !d_rev_range || !d_q || !d_cand || !d_best_score || !d_best_slot || dim <= 0 || B < 0 || pool <= 0)
        return rr_fail(RR_EINVAL, "rr_best_review_scores: bad argument");
    if ((d_slot_file == nullptr) != (d_limit == nullptr))
        return rr_fail(RR_EINVAL, "rr_best_review_scores: d_slot_file and d_limit go together");
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_best_review(d_rev_emb, d_rev_range, n_products, dim, d_q, B, d_cand, pool, d_slot_file, d_limit,
                                 d_best_score, d_best_slot, static_cast<cudaStream_t>(stream));
}

extern "C" int rr_normalize_rows(const float* d_in, int64_t n_rows, int32_t dim, float* d_out_f32, uint16_t* d_out_bf16,
                                 int32_t dim_pad, float* d_norms, int device, rr_stream stream) {
    if (!d_in || n_rows < 0 || dim <= 0 || (!d_out_f32 && !d_out_bf16 && !d_norms))
        return rr_fail(RR_EINVAL, "rr_normalize_rows: bad argument");
    if (d_out_bf16 && (dim_pad < dim || dim_pad % 64))
        return rr_fail(RR_EINVAL, "rr_normalize_rows: dim_pad must be a multiple of 64 and >= dim");
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_normalize_rows(d_in, n_rows, dim, d_out_f32, d_out_bf16, dim_pad, d_norms,
                                    static_cast<cudaStream_t>(stream));
}

extern "C" int rr_max_row_norm(const float* d_in, int64_t n_rows, int32_t dim, float* d_out, int device, rr_stream stream) {
    if (!d_in || !d_out || n_rows < 0 || dim <= 0) return rr_fail(RR_EINVAL, "rr_max_row_norm: bad argument");
    RR_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return rr_launch_max_row_norm(d_in, n_rows, dim, d_out, sms, static_cast<cudaStream_t>(stream));
}

extern "C" int rr_bf16_rows(const float* d_in, int64_t n_rows, int32_t dim, uint16_t* d_out_bf16, int32_t dim_pad,
                            int device, rr_stream stream) {
    if (!d_in || !d_out_bf16 || n_rows < 0 || dim <= 0 || dim_pad < dim || dim_pad % 64)
        return rr_fail(RR_EINVAL, "rr_bf16_rows: bad argument");
    RR_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return rr_launch_bf16_rows(d_in, n_rows, dim, d_out_bf16, dim_pad, sms, static_cast<cudaStream_t>(stream));
}

extern "C" int rr_gate_factors(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                               const uint32_t* d_fixed_bits, const uint8_t* d_pat, const int32_t* d_pat_off,
                               const int32_t* d_group_pat_off, const int32_t* d_group_fixed,
                               const int32_t* d_query_group_off, int32_t B, const int64_t* d_cand, int32_t pool,
                               double penalty, float* d_gate, int32_t* d_hits, int device, rr_stream stream) {
    if (!d_text || !d_text_off || !d_text_len || !d_pat_off || !d_group_pat_off || !d_query_group_off || !d_cand ||
        !d_gate || B < 0 || pool <= 0)
        return rr_fail(RR_EINVAL, "rr_gate_factors: bad argument");
    if (d_fixed_bits != nullptr && d_group_fixed == nullptr)
        return rr_fail(RR_EINVAL, "rr_gate_factors: d_fixed_bits needs d_group_fixed");
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_gate_query(d_text, d_text_off, d_text_len, n_docs, d_fixed_bits, d_pat, d_pat_off, d_group_pat_off,
                                d_group_fixed, d_query_group_off, B, d_cand, pool, penalty, d_gate, d_hits,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int rr_gate_fixed_bitmaps(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len,
                                     int64_t n_docs, const uint8_t* d_pat, const int32_t* d_pat_off,
                                     const int32_t* d_group_pat_off, int32_t n_groups, uint32_t* d_bits, int device,
                                     rr_stream stream) {
    if (!d_text || !d_text_off || !d_text_len || !d_pat || !d_pat_off || !d_group_pat_off || !d_bits || n_groups < 0 ||
        n_groups > 32)
        return rr_fail(RR_EINVAL, "rr_gate_fixed_bitmaps: bad argument (at most 32 groups)");
    RR_CUDA(cudaSetDevice(device));
    return rr_launch_gate_bitmaps(d_text, d_text_off, d_text_len, n_docs, d_pat, d_pat_off, d_group_pat_off, n_groups,
                                  d_bits, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// one-shot hybrid search
// ---------------------------------------------------------------------------------------------
static int hybrid_locked(rr_index* ix, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                         int32_t B, int32_t l_max, const rr_fusion_params* fp, int32_t dense_mode,
                         int64_t* d_top_row, float* d_top_final, cudaStream_t s, int32_t* d_uncertified = nullptr) {
    const int pool = fp->pool;
    const size_t bp = (size_t)B * pool;
    size_t bytes = 0;
    bytes += align_up(sizeof(int64_t) * bp, 256) * 2;   // cand, grow
    bytes += align_up(sizeof(float) * bp, 256) * 2;     // dense, bm25
    bytes += align_up(sizeof(double) * bp, 256) * 2;    // n, avg
    bytes += align_up(sizeof(int32_t) * (size_t)B, 256);
    RR_TRY(ix->tuples.ensure(bytes));
    Carver c{static_cast<char*>(ix->tuples.p)};
    int64_t* cand = c.take<int64_t>(bp);
    int64_t* grow = c.take<int64_t>(bp);
    float* dense = c.take<float>(bp);
    float* bm25 = c.take<float>(bp);
    double* nrev = c.take<double>(bp);
    double* avg = c.take<double>(bp);
    int32_t* count = c.take<int32_t>((size_t)B);
    RR_TRY(dense_topk_locked(ix, d_q, B, pool, dense_mode, cand, dense, count, s, d_uncertified));
    RR_TRY(candidates_locked(ix, d_term_ids, d_n_terms, B, l_max, cand, pool, bm25, nrev, avg, grow, s));
    return rr_launch_fuse(fp, B, pool, 1, 0, count, dense, bm25, nrev, avg, grow, nullptr, nullptr, nullptr, d_top_row,
                          d_top_final, nullptr, nullptr, nullptr, s);
}

static int check_hybrid_args(rr_index* ix, const void* q, const rr_fusion_params* fp, const void* top_row,
                             const void* top_final, int32_t B) {
    if (!ix || !q || !fp || !top_row || !top_final || B < 0) return rr_fail(RR_EINVAL, "rr_hybrid_search: bad argument");
    if (fp->pool < fp->k) return rr_fail(RR_EINVAL, "rr_hybrid_search: pool must be >= k (pool = max(k, rerank_k, floor))");
    return RR_OK;
}

extern "C" int rr_hybrid_search(rr_index* ix, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                                int32_t B, int32_t l_max, const rr_fusion_params* fp, int32_t dense_mode,
                                int64_t* d_top_row, float* d_top_final, rr_stream stream) {
    RR_TRY(check_hybrid_args(ix, d_q, fp, d_top_row, d_top_final, B));
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    return hybrid_locked(ix, d_q, d_term_ids, d_n_terms, B, l_max, fp, dense_mode, d_top_row, d_top_final,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int rr_hybrid_search_deferred(rr_index* ix, const float* d_q, const int32_t* d_term_ids,
                                         const int32_t* d_n_terms, int32_t B, int32_t l_max, const rr_fusion_params* fp,
                                         int32_t dense_mode, int64_t* d_top_row, float* d_top_final,
                                         int32_t* d_uncertified, rr_stream stream) {
    RR_TRY(check_hybrid_args(ix, d_q, fp, d_top_row, d_top_final, B));
    if (!d_uncertified) return rr_fail(RR_EINVAL, "rr_hybrid_search_deferred: d_uncertified is required");
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    return hybrid_locked(ix, d_q, d_term_ids, d_n_terms, B, l_max, fp, dense_mode, d_top_row, d_top_final,
                         static_cast<cudaStream_t>(stream), d_uncertified);
}

extern "C" int rr_hybrid_search_host(rr_index* ix, const float* h_q, const int32_t* h_term_ids,
                                     const int32_t* h_n_terms, int32_t B, int32_t l_max, const rr_fusion_params* fp,
                                     int32_t dense_mode, int64_t* h_top_row, float* h_top_final, rr_stream stream) {
    RR_TRY(check_hybrid_args(ix, h_q, fp, h_top_row, h_top_final, B));
    if (B == 0) return RR_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    RR_CUDA(cudaSetDevice(ix->device));
    ScratchFence fence(ix, static_cast<cudaStream_t>(stream));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t q_bytes = sizeof(float) * (size_t)B * ix->d.dim;
    const bool have_terms = h_term_ids && h_n_terms && l_max > 0;
    const size_t t_bytes = have_terms ? sizeof(int32_t) * (size_t)B * l_max : 0;
    const size_t n_bytes = have_terms ? sizeof(int32_t) * (size_t)B : 0;
    const size_t r_bytes = sizeof(int64_t) * (size_t)B * fp->k;
    const size_t f_bytes = sizeof(float) * (size_t)B * fp->k;
    RR_TRY(ix->staging.ensure(align_up(q_bytes, 256) + align_up(t_bytes, 256) + align_up(n_bytes, 256) +
                              align_up(r_bytes, 256) + align_up(f_bytes, 256) + 256));
    char* p = static_cast<char*>(ix->staging.p);
    float* d_q = reinterpret_cast<float*>(p); p += align_up(q_bytes, 256);
    int32_t* d_t = reinterpret_cast<int32_t*>(p); p += align_up(t_bytes, 256);
    int32_t* d_n = reinterpret_cast<int32_t*>(p); p += align_up(n_bytes, 256);
    int64_t* d_r = reinterpret_cast<int64_t*>(p); p += align_up(r_bytes, 256);
    float* d_f = reinterpret_cast<float*>(p);
    // Inputs and outputs are carved contiguously ([q | terms | n_terms] and [rows | final]).  Small batches travel as ONE
    // copy each way through a pinned bounce buffer of the handle: five separate cudaMemcpyAsync calls on pageable memory
    // cost more than the whole search of a single query (r02: 200 us per query at 10 k docs, most of it copies).
    const size_t in_span = (size_t)(reinterpret_cast<char*>(d_r) - reinterpret_cast<char*>(d_q));
    const size_t out_span = align_up(r_bytes, 256) + f_bytes;
    const bool packed = in_span + out_span <= ((size_t)256 << 10);
    if (packed) {
        if (ix->h_pinned_cap < in_span + out_span) {
            if (ix->h_pinned) cudaFreeHost(ix->h_pinned);
            ix->h_pinned = nullptr; ix->h_pinned_cap = 0;
            RR_CUDA(cudaMallocHost(&ix->h_pinned, (size_t)256 << 10));
            ix->h_pinned_cap = (size_t)256 << 10;
        }
        char* hp = static_cast<char*>(ix->h_pinned);
        memcpy(hp, h_q, q_bytes);
        if (have_terms) {
            memcpy(hp + (reinterpret_cast<char*>(d_t) - reinterpret_cast<char*>(d_q)), h_term_ids, t_bytes);
            memcpy(hp + (reinterpret_cast<char*>(d_n) - reinterpret_cast<char*>(d_q)), h_n_terms, n_bytes);
        }
        RR_CUDA(cudaMemcpyAsync(d_q, hp, in_span, cudaMemcpyHostToDevice, s));
    } else {
        RR_CUDA(cudaMemcpyAsync(d_q, h_q, q_bytes, cudaMemcpyHostToDevice, s));
        if (have_terms) {
            RR_CUDA(cudaMemcpyAsync(d_t, h_term_ids, t_bytes, cudaMemcpyHostToDevice, s));
            RR_CUDA(cudaMemcpyAsync(d_n, h_n_terms, n_bytes, cudaMemcpyHostToDevice, s));
        }
    }
    // Small batches on the exact path (the single-query shape Streamlit issues, app/app_product_search.py:245-261)
    // are launch-latency bound: ~20 dependent launches of a few microseconds each.  The second call with the same
    // shape captures the device work into a CUDA graph (on a private stream: the caller's may be the legacy default
    // stream, which cannot be captured); later calls replay it with one launch.  Buffers the graph points into are
    // the handle's staging / scratch buffers; any reallocation invalidates the graphs (generation counter).
    bool done = false;
    const bool exact = dense_mode == RR_DENSE_EXACT ||
                       (dense_mode == RR_DENSE_AUTO && !(rr_tc_supported(ix->cc_major, ix->cc_minor) && ix->d.d_emb_bf16 != nullptr &&
                                                         rr_tc_can_handle(ix->d.dim_pad, fp->pool) && B >= 32 && ix->d.n_docs >= 65536));
    if (B <= 8 && exact && !g_prof_on.load(std::memory_order_relaxed) && !getenv("RR_NO_GRAPHS")) {
        uint64_t key = 1469598103934665603ull;
        auto mix = [&key](const void* p, size_t n) {
            const unsigned char* c = static_cast<const unsigned char*>(p);
            for (size_t i = 0; i < n; ++i) { key ^= c[i]; key *= 1099511628211ull; }
        };
        const int32_t shape[4] = {B, l_max, have_terms ? 1 : 0, dense_mode};
        mix(shape, sizeof(shape));
        mix(fp, sizeof(*fp));
        rr_index::GraphEntry* ent = nullptr;
        for (auto& g : ix->graphs) if (g.key == key) { ent = &g; break; }
        if (!ent) {
            if (ix->graphs.size() >= 16) {
                for (auto& g : ix->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
                ix->graphs.clear();
            }
            ix->graphs.push_back({key, 0, 0, nullptr});
            ent = &ix->graphs.back();
        }
        const uint64_t gen = g_buf_generation.load();
        if (ent->exec && ent->generation != gen) { cudaGraphExecDestroy(ent->exec); ent->exec = nullptr; ent->seen = 1; }
        if (!ent->exec && ent->seen == 1) {
            // warm (buffers allocated, attributes set by the previous call): capture
            if (!ix->capture_stream) cudaStreamCreateWithFlags(&ix->capture_stream, cudaStreamNonBlocking);
            cudaGraph_t graph = nullptr;
            if (ix->capture_stream && cudaStreamBeginCapture(ix->capture_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = hybrid_locked(ix, d_q, have_terms ? d_t : nullptr, have_terms ? d_n : nullptr, B, l_max, fp,
                                             dense_mode, d_r, d_f, ix->capture_stream);
                const cudaError_t e = cudaStreamEndCapture(ix->capture_stream, &graph);
                if (rc == RR_OK && e == cudaSuccess && graph && g_buf_generation.load() == gen &&
                    cudaGraphInstantiate(&ent->exec, graph, 0) == cudaSuccess) {
                    ent->generation = gen;
                } else {
                    ent->exec = nullptr;
                    ent->seen = -1;                 // this shape cannot be captured: stay on the plain path
                    cudaGetLastError();
                }
                if (graph) cudaGraphDestroy(graph);
            } else {
                ent->seen = -1;
                cudaGetLastError();
            }
        }
        if (ent->exec) {
            RR_CUDA(cudaGraphLaunch(ent->exec, s));
            rr_count_launch();
            done = true;
        } else if (ent->seen == 0) {
            ent->seen = 1;
        }
    }
    if (!done)
        RR_TRY(hybrid_locked(ix, d_q, have_terms ? d_t : nullptr, have_terms ? d_n : nullptr, B, l_max, fp, dense_mode,
                             d_r, d_f, s));
    if (packed) {
        char* hp = static_cast<char*>(ix->h_pinned) + in_span;
        RR_CUDA(cudaMemcpyAsync(hp, d_r, out_span, cudaMemcpyDeviceToHost, s));
        RR_CUDA(cudaStreamSynchronize(s));
        memcpy(h_top_row, hp, r_bytes);
        memcpy(h_top_final, hp + align_up(r_bytes, 256), f_bytes);
        return RR_OK;
    }
    RR_CUDA(cudaMemcpyAsync(h_top_row, d_r, r_bytes, cudaMemcpyDeviceToHost, s));
    RR_CUDA(cudaMemcpyAsync(h_top_final, d_f, f_bytes, cudaMemcpyDeviceToHost, s));
    RR_CUDA(cudaStreamSynchronize(s));
    return RR_OK;
}
