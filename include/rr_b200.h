/* rr_b200.h -- C ABI of the B200-native hybrid-retrieval hot path.
 *
 * The reference (Ntropy86/review-recommender) is pure Python and has no FFI of its own; the
 * seam this library replaces is the set of Python functions listed below.  Each entry point
 * names the reference interface it stands behind (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; every pointer is either a HOST pointer (prefix h_) or a DEVICE
 *     pointer (prefix d_) and is documented as such;
 *   - every call returns 0 on success or a negative RR_E* code and never throws, aborts or
 *     falls back to a CPU implementation; rr_last_error() gives the thread-local message;
 *   - the index handle BORROWS the device buffers named in rr_index_desc (the caller -- torch
 *     tensors in the shipped host code -- owns them and must keep them alive); the handle owns
 *     only its internal scratch;
 *   - calls on one handle are serialised by an internal mutex (Streamlit runs sessions on
 *     several threads against one cached index, app/app_product_search.py:53,71,119);
 *   - work is enqueued on the caller's stream; *_host entry points synchronise that stream
 *     before returning, device-pointer entry points do not.
 */
#ifndef RR_B200_H
#define RR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_OK              0
#define RR_EINVAL         -1   /* bad argument */
#define RR_ECUDA          -2   /* CUDA runtime error (message has the CUDA string) */
#define RR_ENOMEM         -3
#define RR_EUNSUPPORTED   -4   /* e.g. tensor path asked for but no bf16 copy / not sm_100 */
#define RR_EOVERFLOW      -5

typedef struct rr_index rr_index;
typedef struct rr_postings rr_postings;
typedef void* rr_stream;       /* cudaStream_t */

const char* rr_last_error(void);
int rr_abi_version(void);
/* out3 = { sizeof(rr_index_desc), sizeof(rr_fusion_params), sizeof(rr_dense_stats) }: lets a binding check its
 * struct layouts against the library it loaded (the Python binding does so in _lib.load()). */
void rr_struct_sizes(int32_t* out3);

/* ------------------------------------------------------------------------------------------
 * Host-side BM25 index construction.
 * Replaces rank_bm25.BM25Okapi.__init__ as called at app/test.py:156 and
 * app/app_product_search.py:142 (corpus = blob["corpus"] of product_bm25.pkl,
 * nlp/12_product_prep.py:85-88), k1=1.5 b=0.75 epsilon=0.25.
 * The corpus is passed flat: h_doc_offsets int64[n_docs+1], h_token_ids int32[total], term ids
 * in [0, vocab_size).  Three steps so that a row-sharded build can all-reduce the statistics:
 * ---------------------------------------------------------------------------------------- */

/* df[t] += number of local docs containing t; first_pos[t] = min(global flat position of the
 * first occurrence of t) -- the dict insertion order rank_bm25 sums idf in.  The caller
 * initialises h_df to 0 and h_first_pos to INT64_MAX.  token_pos0 = global flat index of this
 * shard's first token. */
int rr_bm25_local_stats(const int64_t* h_doc_offsets, const int32_t* h_token_ids, int64_t n_docs,
                        int32_t vocab_size, int64_t token_pos0,
                        int64_t* h_df, int64_t* h_first_pos, int64_t* h_total_tokens);

/* idf[t] = ln(N-df+.5) - ln(df+.5); negatives floored to epsilon*mean(idf) (mean over all
 * present terms, summed in first_pos order); absent terms get 0. */
int rr_bm25_idf(const int64_t* h_df, const int64_t* h_first_pos, int32_t vocab_size,
                int64_t corpus_size, double epsilon, double* h_idf, double* h_average_idf);

/* Index layout ("hybrid-blocked postings").  posting = {uint32 local doc, float impact} with
 *   impact = (float)( idf[t] * ( tf*(k1+1) / (tf + k1*(1 - b + b*len/avgdl)) ) )   (float64 math)
 * Documents are cut into tiles of tile_docs.  Terms fall in two classes by their LOCAL document frequency:
 *   FREQUENT  df >= rr_bm25_dir_threshold(n_tiles) = RR_DIR_MIN_PER_TILE * n_tiles (at least 8 postings per tile on
 *             average): tile-blocked -- inside tile i the postings are grouped by frequent slot (= term-ascending),
 *             doc-ascending inside a group; dir[i*(n_freq+1) + f .. + f+1] bounds slot f's group relative to
 *             tile_base[i] (every tile starts on a 16-byte boundary).  term_slot[t] = f.
 *   RARE      everything else (term_slot[t] = -1): ONE term-major list per term, doc-ascending, at
 *             tile_base[n_tiles] + rare_off[t] .. + rare_off[t+1]; the part of a list that falls into a tile is
 *             found by a lower-bound on the doc id (per batch, rr_bm25_get_scores does it for all tiles at once).
 * Metadata is n_tiles*(n_freq+1)*4 + vocab*12 bytes: the dense [n_tiles, vocab+1] table of ABI 2 (1.3 GB at 20 M
 * documents x 200 k terms) is gone; a directory entry is only spent where it indexes >= 64 bytes of postings. */
#define RR_DIR_MIN_PER_TILE 8
int32_t rr_bm25_dir_threshold(int32_t n_tiles);
int rr_bm25_build_postings(const int64_t* h_doc_offsets, const int32_t* h_token_ids, int64_t n_docs,
                           int32_t vocab_size, const double* h_idf, double avgdl, double k1, double b,
                           int32_t tile_docs, int32_t n_threads, rr_postings** out);
int64_t         rr_postings_nnz(const rr_postings*);        /* entries incl. alignment padding (both regions) */
int32_t         rr_postings_n_tiles(const rr_postings*);
int32_t         rr_postings_n_freq(const rr_postings*);
const uint64_t* rr_postings_data(const rr_postings*);       /* host, nnz x {u32 doc, f32 impact} */
const uint64_t* rr_postings_tile_base(const rr_postings*);  /* host, n_tiles+1 */
const uint32_t* rr_postings_dir(const rr_postings*);        /* host, n_tiles*(n_freq+1) */
const int32_t*  rr_postings_term_slot(const rr_postings*);  /* host, vocab_size */
const uint64_t* rr_postings_rare_off(const rr_postings*);   /* host, vocab_size+1 */
/* Forward (doc-major) copy of the same impacts for candidate-mode scoring: doc d's entries
 * {u32 term, f32 impact}, term-ascending, are fwd_data[fwd_off[d] .. fwd_off[d+1]). */
const uint64_t* rr_postings_fwd_off(const rr_postings*);    /* host, n_docs+1 */
const uint64_t* rr_postings_fwd_data(const rr_postings*);   /* host, fwd_off[n_docs] entries */
void            rr_postings_free(rr_postings*);

/* The same construction on the GPU, for a tokenised corpus that is already in device memory (bit-identical output).
 * begin:  sorts the (doc, term) token keys, finds the unique pairs, classifies the terms by local df and derives the
 *         tile geometry; reports how many forward entries (n_unique), postings incl. alignment padding (n_postings)
 *         and frequent terms (n_freq) the caller has to allocate for;
 * stats:  adds this shard's df to d_df int64[V] (caller-zeroed) and lowers d_first_pos int64[V] (caller-initialised to
 *         INT64_MAX) -- all-reduce them over the shards, then rr_bm25_idf on the host;
 * finish: impacts, forward index, postings of both regions, tile_base[n_tiles+1], dir[n_tiles*(n_freq+1)],
 *         term_slot[V], rare_off[V+1], fwd_off[n_docs+1] into caller-owned device buffers.
 *         d_doc_offsets / d_token_ids must stay valid until finish.
 * vocab_size <= 2^24, tile_docs <= 65536 (multiple of 4), fewer than 2^31 tokens per shard. */
typedef struct rr_bm25_gpu_builder rr_bm25_gpu_builder;
int rr_bm25_gpu_build_begin(rr_bm25_gpu_builder** out, const int64_t* d_doc_offsets, const int32_t* d_token_ids,
                            int64_t n_docs, int64_t n_tokens, int32_t vocab_size, int32_t tile_docs,
                            int64_t* n_unique_out, int64_t* n_postings_out, int32_t* n_tiles_out, int32_t* n_freq_out,
                            int device, rr_stream);
int rr_bm25_gpu_build_stats(rr_bm25_gpu_builder*, int64_t token_pos0, int64_t* d_df, int64_t* d_first_pos, rr_stream);
int rr_bm25_gpu_build_finish(rr_bm25_gpu_builder*, const double* d_idf, double avgdl, double k1, double b,
                             uint64_t* d_postings, uint64_t* d_tile_base, uint32_t* d_dir, int32_t* d_term_slot,
                             uint64_t* d_rare_off, uint64_t* d_fwd_off, uint64_t* d_fwd_data, rr_stream);
void rr_bm25_gpu_build_free(rr_bm25_gpu_builder*);

/* ------------------------------------------------------------------------------------------
 * Device index
 * ---------------------------------------------------------------------------------------- */
typedef struct rr_index_desc {
    int64_t n_docs;              /* rows of this shard */
    int64_t row_offset;          /* global row of local row 0 (row-sharded corpus) */
    int32_t dim;                 /* D */
    int32_t dim_pad;             /* row length of d_emb_bf16 (multiple of 64) or 0 */
    const float*    d_emb_f32;   /* [n_docs, dim] row-major; the reference's Vn (app/app_product_search.py:110) */
    const uint16_t* d_emb_bf16;  /* [n_docs, dim_pad] round-to-nearest bf16 copy, or NULL (exact path only) */
    float   max_row_norm;        /* max ||row||_2, used for the bf16 error bound */
    int32_t vocab_size;          /* 0 = no BM25 index ("BM25 absent": zeros, app/app_product_search.py:202) */
    int32_t tile_docs;
    int32_t n_tiles;
    int32_t n_freq;              /* frequent terms (directory slots per tile) */
    const uint64_t* d_postings;  /* as rr_postings_data */
    const uint64_t* d_tile_base; /* [n_tiles+1]; tile_base[n_tiles] = first posting of the rare region */
    const uint32_t* d_dir;       /* [n_tiles, n_freq+1] */
    const int32_t*  d_term_slot; /* [vocab_size] */
    const uint64_t* d_rare_off;  /* [vocab_size+1] */
    const uint64_t* d_fwd_off;   /* [n_docs+1] forward index (optional: NULL = candidate mode searches the postings) */
    const uint64_t* d_fwd_data;  /* {u32 term, f32 impact} per (doc, term), term-ascending inside a doc */
    const double*   d_n_reviews; /* [n_docs] n_reviews with NaN already mapped to 0 (:264), or NULL */
    const double*   d_avg_stars; /* [n_docs] avg_stars, NaN allowed (:265), or NULL */
} rr_index_desc;

int  rr_index_create(rr_index** out, const rr_index_desc* desc, int device);
void rr_index_destroy(rr_index*);

/* ------------------------------------------------------------------------------------------
 * Hot path, device pointers.  B = queries in the batch.
 * ---------------------------------------------------------------------------------------- */

/* rank_bm25.BM25Okapi.get_scores (app/test.py:170, app/app_product_search.py:206), batched:
 * d_out[b, doc] = sum over the query's term list (duplicates counted) of impact(term, doc).
 * d_term_ids int32[B, l_max] (ids <0 or >=V are "unknown": contribute 0), d_n_terms int32[B].
 * d_out float[B, ld_out], ld_out >= n_docs and a multiple of 4. */
int rr_bm25_get_scores(rr_index*, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                       int32_t l_max, float* d_out, int64_t ld_out, rr_stream);

/* The gather half of bm25_scores (app/test.py:168-173) / _bm25_for_candidates
 * (app/app_product_search.py:201-208) without materialising N scores: BM25 of the given
 * candidate rows only, bit-identical to rr_bm25_get_scores at those rows.
 * d_cand int64[B, pool] local rows (<0 = no candidate -> 0). */
int rr_bm25_candidates(rr_index*, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                       int32_t l_max, const int64_t* d_cand, int32_t pool, float* d_out, rr_stream);

/* cosine_similarity_search utils.py:111-124 (= cosine_search app/test.py:125-132, _cosine_pool
 * app/app_product_search.py:192-195), batched: for each query the `pool` rows of largest
 * fp32 dot product, ordered (similarity desc, row asc).  d_q float[B, dim].
 * mode: RR_DENSE_AUTO | RR_DENSE_EXACT (fp32 everywhere) | RR_DENSE_TENSOR (bf16 tcgen05
 * shortlist + exact fp32 rescoring + certification; uncertified queries are redone exactly).
 * d_idx int64[B, pool] local rows (-1 past the end when pool > n_docs), d_sims float[B, pool],
 * d_count int32[B] = min(pool, n_docs). */
#define RR_DENSE_AUTO   0
#define RR_DENSE_EXACT  1
#define RR_DENSE_TENSOR 2
int rr_dense_topk(rr_index*, const float* d_q, int32_t B, int32_t pool, int32_t mode,
                  int64_t* d_idx, float* d_sims, int32_t* d_count, rr_stream);
/* Same without any host synchronisation: queries whose pool the tensor path could not prove exact are NOT redone;
 * d_uncertified int32[B] says which (1 = best-effort, repeat through rr_dense_topk; 0 = final). */
int rr_dense_topk_deferred(rr_index*, const float* d_q, int32_t B, int32_t pool, int32_t mode,
                           int64_t* d_idx, float* d_sims, int32_t* d_count, int32_t* d_uncertified, rr_stream);

/* Test / debug entry of the shortlist stage: the raw bf16 x bf16 -> fp32 tensor-core scores (tcgen05.mma, exactly what
 * the threshold filter compares) of rows [row0, row0+n_rows) for B <= 128 queries, d_out float[B, n_rows].  row0 must be
 * a multiple of 256 and n_rows <= 256 * (number of SMs).  north_star states "dense cosine within 1e-3 absolute before
 * rescoring": tests compare these with the exact fp32 similarities. */
int rr_dense_debug_bf16_scores(rr_index*, const float* d_q, int32_t B, int64_t row0, int32_t n_rows, float* d_out,
                               rr_stream);

/* Candidate tuples for fusion: BM25 at the candidates plus their metadata and global rows.
 * Output arrays are [B, pool]. */
int rr_candidate_tuples(rr_index*, const int32_t* d_term_ids, const int32_t* d_n_terms, int32_t B,
                        int32_t l_max, const int64_t* d_cand, int32_t pool,
                        float* d_bm25, double* d_n_reviews, double* d_avg_stars, int64_t* d_global_row,
                        rr_stream);

/* Fusion + selection: app/app_product_search.py:256-312 (use_trust=1) and app/test.py:252-309
 * (use_trust=0) on candidate tuples.  Inputs are [B, n_in]; per query the first d_count[b]
 * entries are valid.  They are first ordered by (dense desc, global row asc) and cut to `pool`
 * (this is the cross-shard merge when n_in = n_shards*pool), then normalised, blended and
 * sorted by (final desc, pool position asc).
 * Optional per-candidate inputs (NULL = absent): d_rerank (already min-max normalised, in pool
 * order -- only meaningful when n_in == pool), d_best, d_gate.
 * Outputs: d_top_row int64[B, k] global rows (-1 padding), d_top_final float[B, k],
 * d_top_pos int32[B, k] pool positions; d_components float[B, pool, 8] or NULL:
 * {dense_mm, bm25_mm, prior, trust, final, dense_raw, bm25_raw, best_mm}. */
typedef struct rr_fusion_params {
    double w_dense, w_bm25, w_rerank, w_prior, w_best;
    double prior_C;
    int32_t min_reviews;
    int32_t saturation;      /* 80 in run_search (:303) */
    int32_t use_trust;       /* 1 = Streamlit driver, 0 = CLI driver */
    int32_t rerank_is_f32;   /* 1 when rerank_k > 0 (column is float32), 0 = the float64 `0.0` column */
    int32_t bm25_is_f64_zero;/* 1 = CLI without BM25 (`cand["_bm25"] = 0.0`, app/test.py:252) */
    int32_t k;
    int32_t pool;
    int32_t best_is_raw;     /* 1 = d_best holds raw best-review similarities; K4 min-max normalises them
                                (`best_contrib = _minmax(best_contrib)`, :294); 0 = already normalised */
} rr_fusion_params;

int rr_fuse_topk(const rr_fusion_params*, int32_t B, int32_t n_in, const int32_t* d_count,
                 const float* d_dense, const float* d_bm25, const double* d_n_reviews,
                 const double* d_avg_stars, const int64_t* d_global_row,
                 const float* d_rerank, const float* d_best, const float* d_gate,
                 int64_t* d_top_row, float* d_top_final, int32_t* d_top_pos, float* d_components,
                 int device, rr_stream);

/* Same, for tuples received from n_shards row shards (the cross-shard merge of a row-sharded
 * corpus): shard s holds its [B, per_shard] block of every field at byte offset
 * s*shard_stride_bytes from the field's base pointer (n_in = n_shards*per_shard).
 * per_shard may be SMALLER than pool (each shard sends only its local top-m): d_incomplete[b]
 * (optional) is then set to 1 when some shard sent all m of its tuples and its weakest one still
 * reaches the merged pool's cut-off, i.e. that shard may hold further pool members and the query has
 * to be repeated with per_shard = pool; 0 means the merged pool is provably the exact global pool.
 * A shard block whose first global row is -2 (rr_shard_tuples: dense result not certified) also sets it.
 * d_gate / d_best (optional, NULL = 1.0 / 0.0): the per-candidate gate factor and best-review similarity
 * (raw when best_is_raw) as two more float fields of the tuples, laid out like d_dense -- run_search's `_gate`
 * and `_best` columns (app/app_product_search.py:285-310) in a row-sharded search. */
int rr_fuse_topk_sharded(const rr_fusion_params*, int32_t B, int32_t n_shards, int32_t per_shard,
                         int64_t shard_stride_bytes,
                         const float* d_dense, const float* d_bm25, const double* d_n_reviews,
                         const double* d_avg_stars, const int64_t* d_global_row,
                         const float* d_gate, const float* d_best,
                         int64_t* d_top_row, float* d_top_final, int32_t* d_incomplete, int device, rr_stream);

/* Row-shard half of a distributed search with NO host synchronisation: the shard's exact top-m by dense
 * similarity for all B queries and the candidate tuples, written straight into the all-to-all send buffer.
 * d_send: n_ranks blocks of (B/n_ranks)*m*32 bytes; block g holds, for the queries g*B/n_ranks .. owned by rank
 * g, the five fields back to back: global row int64 | n_reviews f64 | avg_stars f64 | dense f32 | bm25 f32, each
 * [B/n_ranks, m].  A query whose dense result could not be certified by the tensor path (rr_dense_topk would
 * redo it exactly, which needs a read-back) gets global row -2 in all its m tuples: rr_fuse_topk_sharded then
 * reports it in d_incomplete and the caller repeats it through the synchronous entry points. */
int rr_shard_tuples(rr_index*, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                    int32_t B, int32_t l_max, int32_t m, int32_t dense_mode, int32_t n_ranks, void* d_send, rr_stream);

/* One-shot single-shard search (rerank/best/gate absent): dense top-pool -> tuples -> fuse. */
int rr_hybrid_search(rr_index*, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                     int32_t B, int32_t l_max, const rr_fusion_params*, int32_t dense_mode,
                     int64_t* d_top_row, float* d_top_final, rr_stream);

/* Same without ANY host synchronisation, for callers that keep several batches in flight: the tensor path's
 * read-back of the uncertified count is replaced by d_uncertified int32[B] (1 = this query's dense pool could not be
 * proven exact from the bf16 shortlist; its outputs are best-effort and the caller must repeat the query through
 * rr_hybrid_search, which redoes such queries on the exact fp32 path; 0 = final). */
int rr_hybrid_search_deferred(rr_index*, const float* d_q, const int32_t* d_term_ids, const int32_t* d_n_terms,
                              int32_t B, int32_t l_max, const rr_fusion_params*, int32_t dense_mode,
                              int64_t* d_top_row, float* d_top_final, int32_t* d_uncertified, rr_stream);

/* Same, HOST buffers in and out (pinned or pageable); copies are issued on the stream inside
 * the call and the stream is synchronised before returning.  This is the call the reference's
 * search functions make through the Python binding. */
int rr_hybrid_search_host(rr_index*, const float* h_q, const int32_t* h_term_ids, const int32_t* h_n_terms,
                          int32_t B, int32_t l_max, const rr_fusion_params*, int32_t dense_mode,
                          int64_t* h_top_row, float* h_top_final, rr_stream);

/* Best-review scoring: the dense contraction of _best_snippets (app/app_product_search.py:320-370) and
 * best_review_snippets (app/test.py:181-215).  d_rev_emb float[n_slots, dim] holds the review embeddings
 * grouped by product, file order inside a product, already L2-normalised (the reference normalises the
 * selected rows on every call, :349).  d_rev_range int64[n_products, 2] = {first slot, end slot} of the
 * reviews of product row r (rows with the same SKU share a range; no reviews: lo == hi).
 * For every (query, candidate row) the maximum similarity over the product's reviews and the slot of the
 * FIRST review attaining it (np.argmax, :356); products without reviews and invalid candidates give
 * score 0 and slot -1 (run_search leaves best_contrib at 0 for them, :290-293).
 * Optional `max_rows` cap (:343-346): d_slot_file int64[n_slots] = file position of every slot and
 * d_limit int64[B] = first file position dropped for the query (INT64_MAX = no cap); both or neither.
 * d_cand int64[B, pool] product rows; outputs are [B, pool]. */
int rr_best_review_scores(const float* d_rev_emb, const int64_t* d_rev_range, int64_t n_products, int32_t dim,
                          const float* d_q, int32_t B, const int64_t* d_cand, int32_t pool,
                          const int64_t* d_slot_file, const int64_t* d_limit,
                          float* d_best_score, int64_t* d_best_slot, int device, rr_stream);

/* Index preparation: l2_normalize utils.py:40-44 (= _l2norm app/app_product_search.py:179-180) over the rows of
 * product_emb.npy as done once at load (app/app_product_search.py:110, app/test.py:145), bit-identical to NumPy
 * (float32 pairwise sum of squares, sqrt, max(., 1e-12), divide), fused with the round-to-nearest bf16 copy the
 * tensor path reads.  d_in float[n_rows, dim]; d_out_f32 float[n_rows, dim] (may alias d_in) or NULL;
 * d_out_bf16 [n_rows, dim_pad] (dim_pad a multiple of 64, padding zero-filled) or NULL; d_norms float[n_rows]
 * (the norms before normalisation) or NULL. */
int rr_normalize_rows(const float* d_in, int64_t n_rows, int32_t dim, float* d_out_f32, uint16_t* d_out_bf16,
                      int32_t dim_pad, float* d_norms, int device, rr_stream);

/* max over the rows of ||row||_2 into *d_out (device float): rr_index_desc.max_row_norm, the scale of the tensor path's
 * bf16 error bound.  (~1 for the reference's unit-norm Vn, app/app_product_search.py:110.) */
int rr_max_row_norm(const float* d_in, int64_t n_rows, int32_t dim, float* d_out, int device, rr_stream);

/* The bf16 copy alone (rows already normalised by the caller): round-to-nearest-even, zero padding. */
int rr_bf16_rows(const float* d_in, int64_t n_rows, int32_t dim, uint16_t* d_out_bf16, int32_t dim_pad,
                 int device, rr_stream);

/* Attribute gates: calculate_gate_factor utils.py:88-101 (= _gate_factor app/app_product_search.py:228-236)
 * over `agg_text[:6000]` of every pool member (app/app_product_search.py:297-302, app/test.py:291-297).
 * Text store: d_text = UTF-8 bytes of str(agg_text)[:6000].lower() of every product row, each text starting
 * on a 16-byte boundary (blob length a multiple of 16); d_text_off int64[n_docs] start offsets,
 * d_text_len int32[n_docs] byte lengths.
 * Patterns (lower-case group members): d_pat bytes, d_pat_off int32[n_pat+1]; group g owns patterns
 * d_group_pat_off[g] .. d_group_pat_off[g+1]; query b owns groups d_query_group_off[b] ..
 * d_query_group_off[b+1] (at most 32; the reference keeps 6, utils.py:86).
 * d_group_fixed int32[n_groups] (optional) = bit of d_fixed_bits uint32[n_docs] that already answers the
 * group (the fixed COLORS / SYNONYMS sets, utils.py:15-38, precomputed by rr_gate_fixed_bitmaps) or -1.
 * Output d_gate float[B, pool] = (float32) penalty^(groups without a match), 1.0 for invalid candidates;
 * d_hits int32[B, pool] (optional) = groups with a match. */
int rr_gate_factors(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                    const uint32_t* d_fixed_bits, const uint8_t* d_pat, const int32_t* d_pat_off,
                    const int32_t* d_group_pat_off, const int32_t* d_group_fixed,
                    const int32_t* d_query_group_off, int32_t B, const int64_t* d_cand, int32_t pool,
                    double penalty, float* d_gate, int32_t* d_hits, int device, rr_stream);

/* Per-row bitmap of up to 32 query-independent groups (bit g = some member of group g occurs in the row's
 * text), computed once at load. */
int rr_gate_fixed_bitmaps(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len,
                          int64_t n_docs, const uint8_t* d_pat, const int32_t* d_pat_off,
                          const int32_t* d_group_pat_off, int32_t n_groups, uint32_t* d_bits, int device, rr_stream);

/* Counters for bench.py: number of kernels this library launched since the last reset, and
 * diagnostics of the last rr_dense_topk on this handle. */
int64_t rr_launch_count(int reset);
typedef struct rr_dense_stats {
    int32_t path;            /* 1 exact, 2 tensor */
    int32_t n_uncertified;   /* queries the first tensor pass could not certify (redone: second pass, then exact path) */
    int32_t n_overflow;      /* of those, queries the second tensor pass (4x shortlist) handed to the exact fp32 path */
    int32_t shortlist;       /* k' */
    int32_t n_segments;
    float   eps;             /* bf16 score error bound used for certification */
} rr_dense_stats;
int rr_dense_last_stats(rr_index*, rr_dense_stats* out);

/* Per-kernel-class device timing for bench.py: while enabled, every launch of the classes below is
 * bracketed by CUDA events on its own stream.  rr_profile_collect synchronises those events and
 * returns, per class, the summed milliseconds and the number of launches since the last collect. */
#define RR_PROF_BM25_TILE   0
#define RR_PROF_BM25_CAND   1
#define RR_PROF_DENSE_GEMV  2
#define RR_PROF_SELECT_ROWS 3
#define RR_PROF_TC_FILTER   4
#define RR_PROF_TC_SELECT   5
#define RR_PROF_RESCORE     6
#define RR_PROF_TC_FINALIZE 7
#define RR_PROF_FUSE        8
#define RR_PROF_MISC        9
#define RR_PROF_CLASSES    10
int rr_profile_enable(int on);
int rr_profile_collect(double* h_ms, int64_t* h_launches, int32_t n_classes);

#ifdef __cplusplus
}
#endif
#endif /* RR_B200_H */
