// BM25 index construction on the GPU: flat tokenised corpus (device) -> document statistics, forward index and
// tile-blocked CSR postings, bit-identical to the host builder (bm25_build.cpp) and therefore to the impacts
// rank_bm25.BM25Okapi would add in get_scores (app/test.py:156, app/app_product_search.py:142).
//
//   begin   key(token) = doc << 32 | term  (warp per doc)  ->  radix sort  ->  run heads = unique (doc, term) pairs
//           with tf = run length, in doc-major / term-ascending order (= forward-index order)
//   stats   df[t] += 1 per pair; first_pos[t] = min flat position of t (the dict order rank_bm25 sums idf in)
//           -- the caller all-reduces them over the row shards and computes idf on the host (V values, libm log)
//   begin   (cont.) local df -> term classes (frequent: directory slot, rare: term-major list; include/rr_b200.h
//           "index layout"), frequent postings per tile -> tile bases (16-byte aligned tiles)
//   finish  impact = (float)(idf * (tf*(k1+1) / (tf + k1*(1 - b + b*len/avgdl))))  in float64, the host builder's
//           operation order; forward entries; key2 = frequent: tile | slot | doc-in-tile, rare: 1<<63 | term | doc
//           -> radix sort -> frequent postings in tile-blocked order followed by the rare term-major lists;
//           dir by a lower-bound per (tile, slot).
//
// Sorting and prefix sums use CUB (cub::DeviceRadixSort / DeviceScan, part of the CUDA toolkit); everything else
// is hand-written.  This is the index-build path, not the query path.
#include "rr_internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <new>

namespace {

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RR_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e != cudaSuccess) return rr_fail(RR_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return RR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

constexpr unsigned long long INVALID_KEY = ~0ull;

// one warp per document: key = doc << 32 | term; out-of-vocabulary ids sort to the very end and are dropped
__global__ void __launch_bounds__(256)
make_token_keys_kernel(const long long* __restrict__ doc_off, const int* __restrict__ tok, long long n_docs, int V,
                       unsigned long long* __restrict__ keys, unsigned long long* __restrict__ n_valid) {
    const long long doc = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (doc >= n_docs) return;
    const long long base = doc_off[0], lo = doc_off[doc] - base, hi = doc_off[doc + 1] - base;
    unsigned long long good = 0;
    for (long long i = lo + lane; i < hi; i += 32) {
        const int t = tok[i];
        const bool ok = t >= 0 && t < V;
        keys[i] = ok ? (((unsigned long long)doc << 32) | (unsigned)t) : INVALID_KEY;
        good += ok;
    }
    for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
    if (lane == 0 && good) atomicAdd(n_valid, good);
}

__global__ void mark_heads_kernel(const unsigned long long* __restrict__ keys, long long n_valid, int* __restrict__ head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_valid) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// unique pair u starts at sorted position i: remember the key and the start (tf = next start - this start)
__global__ void collect_pairs_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ head,
                                     const int* __restrict__ uidx, long long n_valid, unsigned long long* __restrict__ u_key,
                                     long long* __restrict__ u_start, long long n_unique) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_valid && head[i]) { u_key[uidx[i]] = keys[i]; u_start[uidx[i]] = i; }
    if (i == 0) u_start[n_unique] = n_valid;
}

// first[d] = index of the first element whose group id is >= d, for d in [0, n_groups]; `gid(u)` is non-decreasing
template <class GroupOf>
__device__ void fill_group_starts(long long u, long long n, long long n_groups, unsigned long long* first, GroupOf gid) {
    if (u >= n) return;
    const long long g = gid(u);
    const long long gp = u == 0 ? -1 : gid(u - 1);
    for (long long d = gp + 1; d <= g; ++d) first[d] = (unsigned long long)u;
    if (u == n - 1) for (long long d = g + 1; d <= n_groups; ++d) first[d] = (unsigned long long)n;
}
__global__ void doc_starts_kernel(const unsigned long long* __restrict__ u_key, long long n_unique, long long n_docs,
                                  unsigned long long* __restrict__ fwd_off) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_unique == 0) { if (u <= n_docs) fwd_off[u] = 0ull; return; }
    fill_group_starts(u, n_unique, n_docs, fwd_off, [&](long long i) { return (long long)(u_key[i] >> 32); });
}

// local document frequency of every term (one unique pair = one document containing the term)
__global__ void pair_df_kernel(const unsigned long long* __restrict__ u_key, long long n_unique, unsigned* __restrict__ df) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_unique) atomicAdd(&df[(unsigned)(u_key[u] & 0xffffffffull)], 1u);
}
// class flags: frequent (df >= theta) -> 1 in is_freq; rare -> its posting count in rare_cnt
__global__ void classify_kernel(const unsigned* __restrict__ df, int V, unsigned theta, int* __restrict__ is_freq,
                                unsigned long long* __restrict__ rare_cnt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V) return;
    const bool f = df[t] >= theta;
    is_freq[t] = f ? 1 : 0;
    rare_cnt[t] = f ? 0ull : (unsigned long long)df[t];
}
// term_slot[t] = exclusive rank among the frequent terms, or -1; totals[0] = n_freq, totals[1] = rare postings
__global__ void slots_kernel(const int* __restrict__ is_freq, const int* __restrict__ rank, const unsigned long long* __restrict__ rare_off,
                             const unsigned long long* __restrict__ rare_cnt, int V, int* __restrict__ term_slot,
                             unsigned long long* __restrict__ rare_off_out, unsigned long long* __restrict__ totals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V) return;
    term_slot[t] = is_freq[t] ? rank[t] : -1;
    rare_off_out[t] = rare_off[t];
    if (t == V - 1) {
        totals[0] = (unsigned long long)(rank[t] + is_freq[t]);
        totals[1] = rare_off[t] + rare_cnt[t];
        rare_off_out[V] = rare_off[t] + rare_cnt[t];
    }
}
// frequent postings per tile (pairs are doc-major: neighbouring lanes mostly share the tile -> warp-aggregated add)
__global__ void tile_count_kernel(const unsigned long long* __restrict__ u_key, long long n_unique, int T,
                                  const int* __restrict__ term_slot, unsigned long long* __restrict__ tile_cnt) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long long tile = -1;
    if (u < n_unique) {
        const unsigned long long k = u_key[u];
        if (term_slot[(unsigned)(k & 0xffffffffull)] >= 0) tile = (long long)(k >> 32) / T;
    }
    const long long lead = __shfl_sync(0xffffffffu, tile, 0);
    const unsigned same = __ballot_sync(0xffffffffu, tile == lead && tile >= 0);
    if (tile >= 0 && tile != lead) atomicAdd(&tile_cnt[tile], 1ull);
    if (lane == 0 && same) atomicAdd(&tile_cnt[lead], (unsigned long long)__popc(same));
}

__global__ void pair_stats_kernel(const unsigned long long* __restrict__ u_key, long long n_unique,
                                  unsigned long long* __restrict__ df) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_unique) atomicAdd(&df[(unsigned)(u_key[u] & 0xffffffffull)], 1ull);
}
__global__ void first_pos_kernel(const int* __restrict__ tok, long long n_tokens, int V, long long pos0,
                                 long long* __restrict__ first_pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tokens) return;
    const int t = tok[i];
    if (t < 0 || t >= V) return;
    const long long pos = pos0 + i;
    if (pos < first_pos[t]) atomicMin(&first_pos[t], pos);          // the unsynchronised read only filters
}

__device__ __forceinline__ unsigned long long pack_u32_f32(unsigned a, float f) {
    return (unsigned long long)a | ((unsigned long long)__float_as_uint(f) << 32);
}

// impact of every unique (doc, term) pair; forward entry; key / value of the tile-blocked sort
__global__ void impacts_kernel(const unsigned long long* __restrict__ u_key, const long long* __restrict__ u_start,
                               long long n_unique, const long long* __restrict__ doc_off, const double* __restrict__ idf,
                               double avgdl, double k1, double b, int T, const int* __restrict__ term_slot,
                               unsigned long long* __restrict__ fwd_data,
                               unsigned long long* __restrict__ key2, unsigned long long* __restrict__ val2) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const unsigned long long k = u_key[u];
    const long long doc = (long long)(k >> 32);
    const unsigned term = (unsigned)(k & 0xffffffffull);
    const double f = (double)(u_start[u + 1] - u_start[u]);
    const double len = (double)(doc_off[doc + 1] - doc_off[doc]);
    // norm = k1 * (1.0 - b + b * len / avgdl);  v = idf * (f * (k1 + 1.0) / (f + norm))     (bm25_build.cpp)
    const double norm = __dmul_rn(k1, __dadd_rn(__dsub_rn(1.0, b), __ddiv_rn(__dmul_rn(b, len), avgdl)));
    const double v = __dmul_rn(idf[term], __ddiv_rn(__dmul_rn(f, __dadd_rn(k1, 1.0)), __dadd_rn(f, norm)));
    const float imp = (float)v;
    fwd_data[u] = pack_u32_f32(term, imp);
    const long long tile = doc / T;
    const int slot = term_slot[term];
    key2[u] = slot >= 0 ? (((unsigned long long)tile << 40) | ((unsigned long long)slot << 16) | (unsigned long long)(doc - tile * T))
                        : ((1ull << 63) | ((unsigned long long)term << 32) | (unsigned long long)doc);
    val2[u] = pack_u32_f32((unsigned)doc, imp);
}

// tile_base[i] = sum over earlier tiles of their frequent posting counts rounded up to an even number (16-byte
// aligned tiles); tile_start[i] = the same sum without the rounding (position in the sorted frequent pairs)
__global__ void tile_bases_kernel(const unsigned long long* __restrict__ tile_cnt, long long n_tiles,
                                  unsigned long long* __restrict__ tile_start, unsigned long long* __restrict__ tile_base,
                                  int* __restrict__ overflow) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long total = 0, run = 0;
    for (long long t = 0; t < n_tiles; ++t) {
        const unsigned long long nnz = tile_cnt[t];
        if (nnz > 0xFFFFFFFFull) *overflow = 1;
        tile_start[t] = run;
        tile_base[t] = total;
        run += nnz;
        total += (nnz + 1ull) & ~1ull;
    }
    tile_start[n_tiles] = run;
    tile_base[n_tiles] = total;
}

// dir[tile][f] = number of frequent postings of the tile whose slot is < f  (f in [0, n_freq]), by lower bound
__global__ void dir_kernel(const unsigned long long* __restrict__ key2, const unsigned long long* __restrict__ tile_start,
                           long long n_tiles, int n_freq, unsigned* __restrict__ dir) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)n_freq + 1;
    if (i >= n_tiles * stride) return;
    const long long tile = i / stride;
    const unsigned long long f = (unsigned long long)(i - tile * stride);
    long long lo = (long long)tile_start[tile], hi = (long long)tile_start[tile + 1];
    const long long s0 = lo;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (((key2[mid] >> 16) & 0xFFFFFFull) < f) lo = mid + 1; else hi = mid;
    }
    dir[i] = (unsigned)(lo - s0);
}

// sorted pairs -> postings: [0, n_freq_pairs) tile-blocked, [n_freq_pairs, n_unique) the rare term-major lists
__global__ void scatter_postings_kernel(const unsigned long long* __restrict__ key2, const unsigned long long* __restrict__ val2,
                                        long long n_unique, long long n_freq_pairs, const unsigned long long* __restrict__ tile_start,
                                        const unsigned long long* __restrict__ tile_base, long long n_tiles,
                                        unsigned long long* __restrict__ postings) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long rare_base = tile_base[n_tiles];
    if (u < n_freq_pairs) {
        const long long tile = (long long)(key2[u] >> 40);
        postings[tile_base[tile] + ((unsigned long long)u - tile_start[tile])] = val2[u];
    } else if (u < n_unique) {
        postings[rare_base + (unsigned long long)(u - n_freq_pairs)] = val2[u];
    }
    if (u < n_tiles) {          // the padding slot of a tile with an odd number of postings
        const unsigned long long nnz = tile_start[u + 1] - tile_start[u];
        if (nnz & 1ull) postings[tile_base[u] + nnz] = pack_u32_f32(0xFFFFFFFFu, 0.0f);
    }
    if (u == 0 && ((n_unique - n_freq_pairs) & 1ll))
        postings[rare_base + (unsigned long long)(n_unique - n_freq_pairs)] = pack_u32_f32(0xFFFFFFFFu, 0.0f);
}

inline unsigned grid_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

struct rr_bm25_gpu_builder {
    int device = 0;
    long long n_docs = 0, n_tokens = 0, n_valid = 0, n_unique = 0, n_tiles = 0, n_postings = 0;
    long long n_freq_pairs = 0, n_rare = 0;
    int V = 0, T = 0, n_freq = 0;
    const long long* doc_off = nullptr;
    const int* tok = nullptr;
    DBuf keys, keys_alt, tmp, head, uidx, u_key, u_start, counters, key2, key2_alt, val2, val2_alt, tile_start, tile_base;
    DBuf df, is_freq, rank, rare_cnt, rare_scan, term_slot, rare_off, tile_cnt;
    void release_all() {
        for (DBuf* b : {&keys, &keys_alt, &tmp, &head, &uidx, &u_key, &u_start, &counters, &key2, &key2_alt, &val2, &val2_alt,
                        &tile_start, &tile_base, &df, &is_freq, &rank, &rare_cnt, &rare_scan, &term_slot, &rare_off, &tile_cnt})
            b->release();
    }
};

extern "C" int32_t rr_bm25_dir_threshold(int32_t n_tiles);

extern "C" int rr_bm25_gpu_build_begin(rr_bm25_gpu_builder** out, const int64_t* d_doc_offsets, const int32_t* d_token_ids,
                                       int64_t n_docs, int64_t n_tokens, int32_t vocab_size, int32_t tile_docs,
                                       int64_t* n_unique_out, int64_t* n_postings_out, int32_t* n_tiles_out,
                                       int32_t* n_freq_out, int device, rr_stream stream) {
    if (!out || !d_doc_offsets || n_docs <= 0 || n_tokens < 0 || vocab_size <= 0 || vocab_size > (1 << 24) || tile_docs <= 0 ||
        (tile_docs & 3) || tile_docs > 65536 || n_docs > 0xFFFFFFF0ll || (!d_token_ids && n_tokens > 0))
        return rr_fail(RR_EINVAL, "rr_bm25_gpu_build_begin: bad argument (vocab <= 2^24, tile_docs <= 65536 and a multiple of 4)");
    RR_CUDA(cudaSetDevice(device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rr_bm25_gpu_builder* h = new (std::nothrow) rr_bm25_gpu_builder();
    if (!h) return rr_fail(RR_ENOMEM, "out of host memory");
    auto fail = [&](int rc) { h->release_all(); delete h; return rc; };
    h->device = device; h->n_docs = n_docs; h->n_tokens = n_tokens; h->V = vocab_size; h->T = tile_docs;
    h->doc_off = reinterpret_cast<const long long*>(d_doc_offsets); h->tok = d_token_ids;
    h->n_tiles = (n_docs + tile_docs - 1) / tile_docs;
    if (h->n_tiles >= (1ll << 23)) return fail(rr_fail(RR_EINVAL, "too many tiles"));
    int rc;
    const size_t nt = (size_t)std::max<long long>(n_tokens, 1);
    if ((rc = h->keys.ensure(nt * 8)) || (rc = h->keys_alt.ensure(nt * 8)) || (rc = h->counters.ensure(64))) return fail(rc);
    cudaMemsetAsync(h->counters.p, 0, 64, s);
    if (n_tokens > 0) {
        make_token_keys_kernel<<<grid_for(n_docs * 32, 256), 256, 0, s>>>(h->doc_off, h->tok, n_docs, vocab_size,
                                                                           h->keys.as<unsigned long long>(),
                                                                           h->counters.as<unsigned long long>());
        rr_count_launch();
        cub::DoubleBuffer<unsigned long long> db(h->keys.as<unsigned long long>(), h->keys_alt.as<unsigned long long>());
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, (long long)n_tokens, 0, 64, s);
        if ((rc = h->tmp.ensure(tmp_bytes))) return fail(rc);
        // invalid keys are all-ones, so every bit has to take part in the sort
        if (cub::DeviceRadixSort::SortKeys(h->tmp.p, tmp_bytes, db, (long long)n_tokens, 0, 64, s) != cudaSuccess)
            return fail(rr_fail(RR_ECUDA, "radix sort of the token keys failed: %s", cudaGetErrorString(cudaGetLastError())));
        if (db.Current() != h->keys.as<unsigned long long>()) std::swap(h->keys, h->keys_alt);
    }
    unsigned long long n_valid = 0;
    if (cudaMemcpyAsync(&n_valid, h->counters.p, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        return fail(rr_fail(RR_ECUDA, "token key pass failed: %s", cudaGetErrorString(cudaGetLastError())));
    h->n_valid = (long long)n_valid;
    if (h->n_valid >= (1ll << 31)) return fail(rr_fail(RR_EOVERFLOW, "more than 2^31 tokens per shard"));
    // run heads -> unique (doc, term) pairs
    const size_t nv = (size_t)std::max<long long>(h->n_valid, 1);
    if ((rc = h->head.ensure(nv * 4)) || (rc = h->uidx.ensure(nv * 4))) return fail(rc);
    long long n_unique = 0;
    if (h->n_valid > 0) {
        mark_heads_kernel<<<grid_for(h->n_valid, 256), 256, 0, s>>>(h->keys.as<unsigned long long>(), h->n_valid, h->head.as<int>());
        rr_count_launch();
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->head.as<int>(), h->uidx.as<int>(), (int)h->n_valid, s);
        if ((rc = h->tmp.ensure(tmp_bytes))) return fail(rc);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tmp_bytes, h->head.as<int>(), h->uidx.as<int>(), (int)h->n_valid, s);
        int last_idx = 0, last_head = 0;
        cudaMemcpyAsync(&last_idx, h->uidx.as<int>() + (h->n_valid - 1), 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(&last_head, h->head.as<int>() + (h->n_valid - 1), 4, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess)
            return fail(rr_fail(RR_ECUDA, "unique-pair scan failed: %s", cudaGetErrorString(cudaGetLastError())));
        n_unique = (long long)last_idx + last_head;
    }
    h->n_unique = n_unique;
    const size_t nu = (size_t)std::max<long long>(n_unique, 1);
    if ((rc = h->u_key.ensure(nu * 8)) || (rc = h->u_start.ensure((nu + 1) * 8))) return fail(rc);
    if (h->n_valid > 0) {
        collect_pairs_kernel<<<grid_for(h->n_valid, 256), 256, 0, s>>>(h->keys.as<unsigned long long>(), h->head.as<int>(),
                                                                       h->uidx.as<int>(), h->n_valid,
                                                                       h->u_key.as<unsigned long long>(), h->u_start.as<long long>(), n_unique);
        rr_count_launch();
    }
    // term classes by local document frequency, then the tile geometry of the frequent region
    const size_t V = (size_t)vocab_size;
    if ((rc = h->df.ensure(V * 4)) || (rc = h->is_freq.ensure(V * 4)) || (rc = h->rank.ensure(V * 4)) ||
        (rc = h->rare_cnt.ensure(V * 8)) || (rc = h->rare_scan.ensure(V * 8)) || (rc = h->term_slot.ensure(V * 4)) ||
        (rc = h->rare_off.ensure((V + 1) * 8)) || (rc = h->tile_cnt.ensure(((size_t)h->n_tiles + 1) * 8)) ||
        (rc = h->tile_start.ensure(((size_t)h->n_tiles + 1) * 8)) || (rc = h->tile_base.ensure(((size_t)h->n_tiles + 1) * 8)))
        return fail(rc);
    cudaMemsetAsync(h->df.p, 0, V * 4, s);
    cudaMemsetAsync(h->tile_cnt.p, 0, ((size_t)h->n_tiles + 1) * 8, s);
    if (n_unique > 0) {
        pair_df_kernel<<<grid_for(n_unique, 256), 256, 0, s>>>(h->u_key.as<unsigned long long>(), n_unique, h->df.as<unsigned>());
        rr_count_launch();
    }
    const unsigned theta = (unsigned)rr_bm25_dir_threshold((int32_t)h->n_tiles);
    classify_kernel<<<grid_for((long long)V, 256), 256, 0, s>>>(h->df.as<unsigned>(), vocab_size, theta, h->is_freq.as<int>(),
                                                                 h->rare_cnt.as<unsigned long long>());
    rr_count_launch();
    {
        size_t t1 = 0, t2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, t1, h->is_freq.as<int>(), h->rank.as<int>(), vocab_size, s);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, h->rare_cnt.as<unsigned long long>(), h->rare_scan.as<unsigned long long>(), vocab_size, s);
        if ((rc = h->tmp.ensure(std::max(t1, t2)))) return fail(rc);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, t1, h->is_freq.as<int>(), h->rank.as<int>(), vocab_size, s);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, t2, h->rare_cnt.as<unsigned long long>(), h->rare_scan.as<unsigned long long>(), vocab_size, s);
    }
    slots_kernel<<<grid_for((long long)V, 256), 256, 0, s>>>(h->is_freq.as<int>(), h->rank.as<int>(), h->rare_scan.as<unsigned long long>(),
                                                              h->rare_cnt.as<unsigned long long>(), vocab_size, h->term_slot.as<int>(),
                                                              h->rare_off.as<unsigned long long>(), h->counters.as<unsigned long long>() + 4);
    rr_count_launch();
    if (n_unique > 0) {
        tile_count_kernel<<<grid_for(n_unique, 256), 256, 0, s>>>(h->u_key.as<unsigned long long>(), n_unique, tile_docs,
                                                                   h->term_slot.as<int>(), h->tile_cnt.as<unsigned long long>());
        rr_count_launch();
    }
    tile_bases_kernel<<<1, 32, 0, s>>>(h->tile_cnt.as<unsigned long long>(), h->n_tiles, h->tile_start.as<unsigned long long>(),
                                       h->tile_base.as<unsigned long long>(), h->counters.as<int>() + 4);
    rr_count_launch();
    unsigned long long total = 0, freq_pairs = 0, totals[2] = {0, 0};
    int overflow = 0;
    cudaMemcpyAsync(&total, h->tile_base.as<unsigned long long>() + h->n_tiles, 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&freq_pairs, h->tile_start.as<unsigned long long>() + h->n_tiles, 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(totals, h->counters.as<unsigned long long>() + 4, 16, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&overflow, h->counters.as<int>() + 4, 4, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess)
        return fail(rr_fail(RR_ECUDA, "tile geometry failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (overflow) return fail(rr_fail(RR_EOVERFLOW, "tile has more than 2^32 postings"));
    h->n_freq = (int)totals[0];
    h->n_rare = (long long)totals[1];
    h->n_freq_pairs = (long long)freq_pairs;
    if (h->n_rare > 0xFFFFFFFFll) return fail(rr_fail(RR_EOVERFLOW, "more than 2^32 rare postings"));
    h->n_postings = (long long)total + ((h->n_rare + 1) & ~1ll);
    if (n_unique_out) *n_unique_out = n_unique;
    if (n_postings_out) *n_postings_out = h->n_postings;
    if (n_tiles_out) *n_tiles_out = (int32_t)h->n_tiles;
    if (n_freq_out) *n_freq_out = (int32_t)h->n_freq;
    *out = h;
    return RR_OK;
}

extern "C" int rr_bm25_gpu_build_stats(rr_bm25_gpu_builder* h, int64_t token_pos0, int64_t* d_df, int64_t* d_first_pos,
                                       rr_stream stream) {
    if (!h || !d_df || !d_first_pos) return rr_fail(RR_EINVAL, "rr_bm25_gpu_build_stats: bad argument");
    RR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (h->n_unique > 0) {
        pair_stats_kernel<<<grid_for(h->n_unique, 256), 256, 0, s>>>(h->u_key.as<unsigned long long>(), h->n_unique,
                                                                     reinterpret_cast<unsigned long long*>(d_df));
        RR_LAUNCH_CHECK();
    }
    if (h->n_tokens > 0) {
        first_pos_kernel<<<grid_for(h->n_tokens, 256), 256, 0, s>>>(h->tok, h->n_tokens, h->V, (long long)token_pos0,
                                                                    reinterpret_cast<long long*>(d_first_pos));
        RR_LAUNCH_CHECK();
    }
    return RR_OK;
}

extern "C" int rr_bm25_gpu_build_finish(rr_bm25_gpu_builder* h, const double* d_idf, double avgdl, double k1, double b,
                                        uint64_t* d_postings, uint64_t* d_tile_base, uint32_t* d_dir, int32_t* d_term_slot,
                                        uint64_t* d_rare_off, uint64_t* d_fwd_off, uint64_t* d_fwd_data, rr_stream stream) {
    if (!h || !d_idf || !(avgdl > 0.0) || !d_postings || !d_tile_base || !d_dir || !d_term_slot || !d_rare_off || !d_fwd_off ||
        !d_fwd_data)
        return rr_fail(RR_EINVAL, "rr_bm25_gpu_build_finish: bad argument");
    RR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long nu = h->n_unique;
    const size_t nu1 = (size_t)std::max<long long>(nu, 1);
    RR_TRY(h->key2.ensure(nu1 * 8));
    RR_TRY(h->key2_alt.ensure(nu1 * 8));
    RR_TRY(h->val2.ensure(nu1 * 8));
    RR_TRY(h->val2_alt.ensure(nu1 * 8));
    doc_starts_kernel<<<grid_for(std::max<long long>(nu, h->n_docs + 1), 256), 256, 0, s>>>(
        h->u_key.as<unsigned long long>(), nu, h->n_docs, reinterpret_cast<unsigned long long*>(d_fwd_off));
    RR_LAUNCH_CHECK();
    if (nu > 0) {
        impacts_kernel<<<grid_for(nu, 256), 256, 0, s>>>(h->u_key.as<unsigned long long>(), h->u_start.as<long long>(), nu,
                                                          h->doc_off, d_idf, avgdl, k1, b, h->T, h->term_slot.as<int>(),
                                                          reinterpret_cast<unsigned long long*>(d_fwd_data),
                                                          h->key2.as<unsigned long long>(), h->val2.as<unsigned long long>());
        RR_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned long long> kb(h->key2.as<unsigned long long>(), h->key2_alt.as<unsigned long long>());
        cub::DoubleBuffer<unsigned long long> vb(h->val2.as<unsigned long long>(), h->val2_alt.as<unsigned long long>());
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, nu, 0, 64, s);
        RR_TRY(h->tmp.ensure(tmp_bytes));
        if (cub::DeviceRadixSort::SortPairs(h->tmp.p, tmp_bytes, kb, vb, nu, 0, 64, s) != cudaSuccess)
            return rr_fail(RR_ECUDA, "radix sort of the postings failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (kb.Current() != h->key2.as<unsigned long long>()) std::swap(h->key2, h->key2_alt);
        if (vb.Current() != h->val2.as<unsigned long long>()) std::swap(h->val2, h->val2_alt);
    }
    RR_CUDA(cudaMemcpyAsync(d_tile_base, h->tile_base.p, ((size_t)h->n_tiles + 1) * 8, cudaMemcpyDeviceToDevice, s));
    RR_CUDA(cudaMemcpyAsync(d_term_slot, h->term_slot.p, (size_t)h->V * 4, cudaMemcpyDeviceToDevice, s));
    RR_CUDA(cudaMemcpyAsync(d_rare_off, h->rare_off.p, ((size_t)h->V + 1) * 8, cudaMemcpyDeviceToDevice, s));
    const long long n_dir = h->n_tiles * ((long long)h->n_freq + 1);
    dir_kernel<<<grid_for(n_dir, 256), 256, 0, s>>>(h->key2.as<unsigned long long>(), h->tile_start.as<unsigned long long>(),
                                                     h->n_tiles, h->n_freq, d_dir);
    RR_LAUNCH_CHECK();
    scatter_postings_kernel<<<grid_for(std::max<long long>(std::max<long long>(nu, h->n_tiles), 1), 256), 256, 0, s>>>(
        h->key2.as<unsigned long long>(), h->val2.as<unsigned long long>(), nu, h->n_freq_pairs,
        h->tile_start.as<unsigned long long>(), h->tile_base.as<unsigned long long>(), h->n_tiles,
        reinterpret_cast<unsigned long long*>(d_postings));
    RR_LAUNCH_CHECK();
    return RR_OK;
}

extern "C" void rr_bm25_gpu_build_free(rr_bm25_gpu_builder* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->release_all();
    delete h;
}
