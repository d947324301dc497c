// Launcher prototypes shared between the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/rr_b200.h"

// bm25_kernels.cu
size_t rr_bm25_rtab_bytes(int B, int l_max, int n_tiles);
int rr_launch_bm25_tile_scores(const rr_index_desc* d, const int32_t* d_terms, const int32_t* d_nterms, int B, int l_max,
                               float* d_out, int64_t ld_out, uint32_t* d_rtab, unsigned* d_counter, cudaStream_t stream);
// V = 0: no BM25 terms (scores are zero), the metadata gather still runs
int rr_launch_bm25_candidates(const rr_index_desc* d, int V, const int32_t* d_terms, const int32_t* d_nterms,
                              int B, int l_max, const int64_t* d_cand, int pool, float* d_bm25, double* d_n_out,
                              double* d_avg_out, int64_t* d_grow_out, cudaStream_t stream, int pack_bg = 0,
                              int64_t pack_stride = 0, const float* d_dense_in = nullptr, float* d_dense_out = nullptr,
                              const int32_t* d_uncertified = nullptr);

// dense_exact.cu
size_t rr_exact_scratch_bytes(int rows, int k, int64_t n, int sm_count);
int rr_launch_dense_scores_f32(const float* d_emb, int64_t n_rows, int D, const float* d_q, int n_queries,
                               float* d_scores, int64_t ld_scores, int sm_count, cudaStream_t stream);
int rr_launch_rescore(const float* d_emb, int64_t n_rows, int D, const float* d_q, const int64_t* d_rows,
                      int n_slots, int B, float* d_out, cudaStream_t stream);
int rr_launch_topk_rows(const float* d_scores, int64_t ld, int64_t n, int rows, int k, void* d_scratch,
                        int64_t* d_idx, float* d_score, int32_t* d_count, int out_ld, int sm_count,
                        cudaStream_t stream);

int rr_launch_best_review(const float* d_rev_emb, const int64_t* d_rev_range, int64_t n_products, int D,
                          const float* d_q, int B, const int64_t* d_cand, int pool, const int64_t* d_slot_file,
                          const int64_t* d_limit, float* d_score, int64_t* d_slot, cudaStream_t stream);

// prep.cu
int rr_launch_normalize_rows(const float* d_in, int64_t n_rows, int D, float* d_out_f32, uint16_t* d_out_bf16,
                             int dim_pad, float* d_norms, cudaStream_t stream);

int rr_launch_max_row_norm(const float* d_in, int64_t n_rows, int D, float* d_out, int sm_count, cudaStream_t stream);
int rr_launch_bf16_rows(const float* d_in, int64_t n_rows, int D, uint16_t* d_out, int dim_pad, int sm_count,
                        cudaStream_t stream);

// gate.cu
int rr_launch_gate_query(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                         const uint32_t* d_fixed_bits, const uint8_t* d_pat, const int32_t* d_pat_off,
                         const int32_t* d_group_pat_off, const int32_t* d_group_fixed, const int32_t* d_query_group_off,
                         int B, const int64_t* d_cand, int pool, double penalty, float* d_gate, int32_t* d_hits,
                         cudaStream_t stream);
int rr_launch_gate_bitmaps(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                           const uint8_t* d_pat, const int32_t* d_pat_off, const int32_t* d_group_pat_off, int n_groups,
                           uint32_t* d_bits, cudaStream_t stream);

// fuse.cu
int rr_launch_fuse(const rr_fusion_params* p, int B, int n_in, int n_shards, int64_t shard_stride_bytes,
                   const int32_t* d_count, const float* d_dense,
                   const float* d_bm25, const double* d_n, const double* d_avg, const int64_t* d_grow,
                   const float* d_rerank, const float* d_best, const float* d_gate, int64_t* d_top_row,
                   float* d_top_final, int32_t* d_top_pos, float* d_components, int32_t* d_incomplete,
                   cudaStream_t stream, int extras_by_slot = 0);

// dense_tc.cu (tcgen05 shortlist path)
struct rr_tc_plan;
