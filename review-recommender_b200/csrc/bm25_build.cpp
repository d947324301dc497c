// Host-side BM25 index construction (C++17, std::thread).
//
// Stands behind rank_bm25.BM25Okapi.__init__ as the reference calls it
// (app/test.py:156, app/app_product_search.py:142): per-document term frequencies, document
// frequencies, avgdl, idf with the epsilon floor -- and lays the result out for the GPU as a
// tile-blocked CSR of {doc, fp32 impact} postings (DESIGN.md "sparse index layout").
//
// All statistics are float64 and follow the library's operation order so that the impacts are the
// correctly rounded fp32 of the values the reference sums in get_scores.
#include "rr_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

struct rr_postings {
    int64_t nnz = 0;
    int32_t n_tiles = 0;
    int32_t vocab = 0;
    int32_t n_freq = 0;
    std::vector<uint64_t> data;
    std::vector<uint64_t> tile_base;
    std::vector<uint32_t> dir;        // [n_tiles][n_freq + 1]
    std::vector<int32_t> term_slot;   // [vocab]: slot of a frequent term, -1 = rare
    std::vector<uint64_t> rare_off;   // [vocab + 1]
    std::vector<uint64_t> fwd_off;
    std::vector<uint64_t> fwd_data;
};

namespace {

int pick_threads(int requested, int64_t work_items) {
    int n = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    if ((int64_t)n > work_items) n = (int)std::max<int64_t>(1, work_items);
    return n;
}

template <class F>
void parallel_for(int n_threads, int64_t n_items, F&& body) {
    if (n_threads <= 1) { body(0, (int64_t)0, n_items); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
        int64_t lo = n_items * t / n_threads, hi = n_items * (t + 1) / n_threads;
        pool.emplace_back([&body, t, lo, hi] { body(t, lo, hi); });
    }
    for (auto& th : pool) th.join();
}

inline uint64_t pack_posting(uint32_t doc, float impact) {
    uint32_t bits;
    std::memcpy(&bits, &impact, 4);
    return (uint64_t)doc | ((uint64_t)bits << 32);
}

// unique (term, tf) of one document, term-ascending
inline void doc_term_freqs(const int32_t* tok, int64_t len, int32_t vocab, std::vector<int32_t>& scratch,
                           std::vector<std::pair<int32_t, int32_t>>& out) {
    scratch.assign(tok, tok + len);
    std::sort(scratch.begin(), scratch.end());
    out.clear();
    for (size_t i = 0; i < scratch.size();) {
        size_t j = i + 1;
        while (j < scratch.size() && scratch[j] == scratch[i]) ++j;
        if (scratch[i] >= 0 && scratch[i] < vocab) out.emplace_back(scratch[i], (int32_t)(j - i));
        i = j;
    }
}

}  // namespace

extern "C" int rr_bm25_local_stats(const int64_t* h_doc_offsets, const int32_t* h_token_ids, int64_t n_docs,
                                   int32_t vocab_size, int64_t token_pos0,
                                   int64_t* h_df, int64_t* h_first_pos, int64_t* h_total_tokens) {
    if (!h_doc_offsets || (!h_token_ids && n_docs > 0 && h_doc_offsets[n_docs] > h_doc_offsets[0]) || n_docs < 0 ||
        vocab_size <= 0 || !h_df || !h_first_pos || !h_total_tokens)
        return rr_fail(RR_EINVAL, "rr_bm25_local_stats: bad argument");
    const int nt = pick_threads(0, n_docs / 4096 + 1);
    std::vector<std::vector<int64_t>> df(nt), fp(nt);
    const int64_t base = h_doc_offsets[0];
    parallel_for(nt, n_docs, [&](int t, int64_t lo, int64_t hi) {
        auto& d = df[t];
        auto& f = fp[t];
        d.assign(vocab_size, 0);
        f.assign(vocab_size, std::numeric_limits<int64_t>::max());
        std::vector<int64_t> stamp(vocab_size, -1);
        for (int64_t doc = lo; doc < hi; ++doc) {
            for (int64_t i = h_doc_offsets[doc]; i < h_doc_offsets[doc + 1]; ++i) {
                int32_t w = h_token_ids[i - base];
                if (w < 0 || w >= vocab_size) continue;
                if (stamp[w] != doc) {
                    stamp[w] = doc;
                    d[w] += 1;
                    if (f[w] == std::numeric_limits<int64_t>::max()) f[w] = token_pos0 + (i - base);
                }
            }
        }
    });
    for (int t = 0; t < nt; ++t)
        for (int32_t w = 0; w < vocab_size; ++w) {
            h_df[w] += df[t][w];
            h_first_pos[w] = std::min(h_first_pos[w], fp[t][w]);
        }
    *h_total_tokens += h_doc_offsets[n_docs] - base;
    return RR_OK;
}

extern "C" int rr_bm25_idf(const int64_t* h_df, const int64_t* h_first_pos, int32_t vocab_size,
                           int64_t corpus_size, double epsilon, double* h_idf, double* h_average_idf) {
    if (!h_df || !h_first_pos || vocab_size <= 0 || corpus_size <= 0 || !h_idf)
        return rr_fail(RR_EINVAL, "rr_bm25_idf: bad argument");
    std::vector<int32_t> order;
    order.reserve(vocab_size);
    for (int32_t w = 0; w < vocab_size; ++w) {
        h_idf[w] = 0.0;
        if (h_df[w] > 0) order.push_back(w);
    }
    // rank_bm25 sums idf while walking its {word: df} dict, i.e. in first-appearance order
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return h_first_pos[a] < h_first_pos[b]; });
    double idf_sum = 0.0;
    std::vector<int32_t> negative;
    for (int32_t w : order) {
        const double freq = (double)h_df[w];
        const double idf = std::log((double)corpus_size - freq + 0.5) - std::log(freq + 0.5);
        h_idf[w] = idf;
        idf_sum += idf;
        if (idf < 0) negative.push_back(w);
    }
    const double avg = order.empty() ? 0.0 : idf_sum / (double)order.size();
    const double eps = epsilon * avg;
    for (int32_t w : negative) h_idf[w] = eps;
    if (h_average_idf) *h_average_idf = avg;
    return RR_OK;
}

// Term classes (see include/rr_b200.h "index layout"): a term with at least RR_DIR_MIN_PER_TILE postings per tile
// on average is FREQUENT and gets a directory slot; every other term is RARE and lives in one term-major list.
extern "C" int32_t rr_bm25_dir_threshold(int32_t n_tiles) {
    const int64_t th = (int64_t)RR_DIR_MIN_PER_TILE * std::max(n_tiles, 1);
    return (int32_t)std::min<int64_t>(th, std::numeric_limits<int32_t>::max());
}

extern "C" int rr_bm25_build_postings(const int64_t* h_doc_offsets, const int32_t* h_token_ids, int64_t n_docs,
                                      int32_t vocab_size, const double* h_idf, double avgdl, double k1, double b,
                                      int32_t tile_docs, int32_t n_threads, rr_postings** out) {
    if (!h_doc_offsets || n_docs < 0 || vocab_size <= 0 || !h_idf || !out || tile_docs <= 0 || (tile_docs & 3) ||
        !(avgdl > 0.0) || n_docs > 0xFFFFFFF0ll)
        return rr_fail(RR_EINVAL, "rr_bm25_build_postings: bad argument");
    rr_postings* p = nullptr;
    try {
        p = new rr_postings();
        const int64_t n_tiles = (n_docs + tile_docs - 1) / tile_docs;
        if (n_tiles > std::numeric_limits<int32_t>::max()) { delete p; return rr_fail(RR_EINVAL, "too many tiles"); }
        p->n_tiles = (int32_t)n_tiles;
        p->vocab = vocab_size;
        const int64_t base = h_doc_offsets[0];
        const int nt = pick_threads(n_threads, n_tiles);
        // threads own contiguous tile ranges [tlo[t], tlo[t+1]) -- i.e. contiguous, ascending document ranges
        std::vector<int64_t> tlo((size_t)nt + 1);
        for (int t = 0; t <= nt; ++t) tlo[t] = n_tiles * t / nt;

        // pass 0: per-thread document frequencies; unique terms per doc for the forward index
        std::vector<std::vector<uint32_t>> df_t((size_t)nt);   // released after the classification
        p->fwd_off.assign((size_t)n_docs + 1, 0ull);
        parallel_for(nt, nt, [&](int, int64_t a, int64_t z) {
            std::vector<int32_t> scratch;
            std::vector<std::pair<int32_t, int32_t>> tf;
            for (int64_t t = a; t < z; ++t) {
                auto& df = df_t[(size_t)t];
                df.assign((size_t)vocab_size, 0u);
                const int64_t d0 = tlo[t] * tile_docs, d1 = std::min<int64_t>(n_docs, tlo[t + 1] * tile_docs);
                for (int64_t doc = d0; doc < d1; ++doc) {
                    doc_term_freqs(h_token_ids + (h_doc_offsets[doc] - base),
                                   h_doc_offsets[doc + 1] - h_doc_offsets[doc], vocab_size, scratch, tf);
                    for (auto& e : tf) df[(size_t)e.first] += 1;
                    p->fwd_off[doc + 1] = tf.size();
                }
            }
        });
        // classification by LOCAL document frequency; rare lists: per-term totals and per-thread start cursors
        const uint32_t theta = (uint32_t)rr_bm25_dir_threshold((int32_t)n_tiles);
        p->term_slot.assign((size_t)vocab_size, -1);
        p->rare_off.assign((size_t)vocab_size + 1, 0ull);
        int32_t n_freq = 0;
        uint64_t rare_total = 0;
        for (int32_t w = 0; w < vocab_size; ++w) {
            uint64_t df = 0;
            for (int t = 0; t < nt; ++t) df += df_t[(size_t)t][(size_t)w];
            p->rare_off[(size_t)w] = rare_total;
            if (df >= theta) p->term_slot[(size_t)w] = n_freq++;
            else rare_total += df;
        }
        p->rare_off[(size_t)vocab_size] = rare_total;
        p->n_freq = n_freq;
        df_t.clear();
        df_t.shrink_to_fit();
        const int64_t stride = (int64_t)n_freq + 1;
        p->dir.assign((size_t)(n_tiles * stride), 0u);
        p->tile_base.assign((size_t)n_tiles + 1, 0ull);

        // pass 1: postings per (tile, frequent slot) -> per-tile exclusive scan; rare postings per (thread, term)
        std::vector<uint64_t> tile_nnz((size_t)n_tiles, 0);
        std::vector<std::vector<uint32_t>> rare_cnt((size_t)nt);
        parallel_for(nt, nt, [&](int, int64_t a, int64_t z) {
            std::vector<int32_t> scratch;
            std::vector<std::pair<int32_t, int32_t>> tf;
            for (int64_t t = a; t < z; ++t) {
                auto& rc = rare_cnt[(size_t)t];
                rc.assign((size_t)vocab_size, 0u);
                for (int64_t tile = tlo[t]; tile < tlo[t + 1]; ++tile) {
                    uint32_t* cnt = p->dir.data() + tile * stride;
                    const int64_t d0 = tile * tile_docs, d1 = std::min<int64_t>(n_docs, d0 + tile_docs);
                    for (int64_t doc = d0; doc < d1; ++doc) {
                        doc_term_freqs(h_token_ids + (h_doc_offsets[doc] - base),
                                       h_doc_offsets[doc + 1] - h_doc_offsets[doc], vocab_size, scratch, tf);
                        for (auto& e : tf) {
                            const int32_t slot = p->term_slot[(size_t)e.first];
                            if (slot >= 0) cnt[slot] += 1; else rc[(size_t)e.first] += 1;
                        }
                    }
                    uint64_t run = 0;
                    for (int64_t f = 0; f < n_freq; ++f) {
                        const uint32_t c = cnt[f];
                        cnt[f] = (uint32_t)run;
                        run += c;
                    }
                    cnt[n_freq] = (uint32_t)run;
                    tile_nnz[(size_t)tile] = run;
                }
            }
        });
        uint64_t total = 0;
        for (int64_t tile = 0; tile < n_tiles; ++tile) {
            if (tile_nnz[tile] > 0xFFFFFFFFull) { delete p; return rr_fail(RR_EOVERFLOW, "tile has more than 2^32 postings"); }
            p->tile_base[tile] = total;
            total += (tile_nnz[tile] + 1ull) & ~1ull;      // every tile starts on a 16-byte boundary
        }
        p->tile_base[n_tiles] = total;                     // = first posting of the rare region
        const uint64_t rare_base = total;
        total += (rare_total + 1ull) & ~1ull;
        p->nnz = (int64_t)total;
        p->data.assign((size_t)total, pack_posting(0xFFFFFFFFu, 0.0f));
        for (int64_t doc = 0; doc < n_docs; ++doc) p->fwd_off[doc + 1] += p->fwd_off[doc];
        p->fwd_data.assign((size_t)p->fwd_off[n_docs], 0ull);
        // rare cursors: thread t starts term w at rare_off[w] + sum of rare_cnt[t' < t][w]
        for (int32_t w = 0; w < vocab_size; ++w) {
            if (p->term_slot[(size_t)w] >= 0) continue;
            uint64_t run = p->rare_off[(size_t)w];
            for (int t = 0; t < nt; ++t) {
                const uint32_t c = rare_cnt[(size_t)t][(size_t)w];
                if (run > 0xFFFFFFFFull) { delete p; return rr_fail(RR_EOVERFLOW, "more than 2^32 rare postings"); }
                rare_cnt[(size_t)t][(size_t)w] = (uint32_t)run;
                run += c;
            }
        }

        // pass 2: scatter, docs ascending inside every (tile, slot) segment and inside every rare list
        parallel_for(nt, nt, [&](int, int64_t a, int64_t z) {
            std::vector<int32_t> scratch;
            std::vector<std::pair<int32_t, int32_t>> tf;
            std::vector<uint32_t> cursor((size_t)std::max<int32_t>(n_freq, 1));
            for (int64_t t = a; t < z; ++t) {
                auto& rcur = rare_cnt[(size_t)t];
                for (int64_t tile = tlo[t]; tile < tlo[t + 1]; ++tile) {
                    const uint32_t* off = p->dir.data() + tile * stride;
                    if (n_freq > 0) std::memcpy(cursor.data(), off, sizeof(uint32_t) * (size_t)n_freq);
                    uint64_t* dst = p->data.data() + p->tile_base[tile];
                    const int64_t d0 = tile * tile_docs, d1 = std::min<int64_t>(n_docs, d0 + tile_docs);
                    for (int64_t doc = d0; doc < d1; ++doc) {
                        const int64_t len = h_doc_offsets[doc + 1] - h_doc_offsets[doc];
                        doc_term_freqs(h_token_ids + (h_doc_offsets[doc] - base), len, vocab_size, scratch, tf);
                        // rank_bm25 get_scores: idf * (f*(k1+1) / (f + k1*(1 - b + b*doc_len/avgdl)))
                        const double norm = k1 * (1.0 - b + b * (double)len / avgdl);
                        for (auto& e : tf) {
                            const double f = (double)e.second;
                            const double idf = h_idf[e.first];
                            const double v = idf * (f * (k1 + 1.0) / (f + norm));
                            const int32_t slot = p->term_slot[(size_t)e.first];
                            if (slot >= 0) dst[cursor[(size_t)slot]++] = pack_posting((uint32_t)doc, (float)v);
                            else p->data[rare_base + rcur[(size_t)e.first]++] = pack_posting((uint32_t)doc, (float)v);
                            p->fwd_data[p->fwd_off[doc] + (size_t)(&e - tf.data())] = pack_posting((uint32_t)e.first, (float)v);
                        }
                    }
                }
            }
        });
    } catch (const std::bad_alloc&) {
        delete p;
        return rr_fail(RR_ENOMEM, "rr_bm25_build_postings: out of host memory");
    }
    *out = p;
    return RR_OK;
}

extern "C" int64_t rr_postings_nnz(const rr_postings* p) { return p ? p->nnz : 0; }
extern "C" int32_t rr_postings_n_tiles(const rr_postings* p) { return p ? p->n_tiles : 0; }
extern "C" const uint64_t* rr_postings_data(const rr_postings* p) { return p ? p->data.data() : nullptr; }
extern "C" const uint64_t* rr_postings_tile_base(const rr_postings* p) { return p ? p->tile_base.data() : nullptr; }
extern "C" int32_t rr_postings_n_freq(const rr_postings* p) { return p ? p->n_freq : 0; }
extern "C" const uint32_t* rr_postings_dir(const rr_postings* p) { return p ? p->dir.data() : nullptr; }
extern "C" const int32_t* rr_postings_term_slot(const rr_postings* p) { return p ? p->term_slot.data() : nullptr; }
extern "C" const uint64_t* rr_postings_rare_off(const rr_postings* p) { return p ? p->rare_off.data() : nullptr; }
extern "C" const uint64_t* rr_postings_fwd_off(const rr_postings* p) { return p ? p->fwd_off.data() : nullptr; }
extern "C" const uint64_t* rr_postings_fwd_data(const rr_postings* p) { return p ? p->fwd_data.data() : nullptr; }
extern "C" void rr_postings_free(rr_postings* p) { delete p; }
