"""BM25 index construction on the GPU (rr_bm25_gpu_build_*) against the host C++ builder: statistics, forward index,
postings of both regions, tile bases, term classes, tile directory and rare-list offsets must be IDENTICAL (bit for bit), including out-of-vocabulary
token ids, empty documents, odd posting counts per tile (alignment padding) and trailing empty tiles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _corpora():
    import review_recommender_b200 as rr
    rng = np.random.default_rng(11)
    out = []
    offs, toks = rr.synth.corpus_tokens(5000, 700)
    out.append(("zipf-5000", offs, toks.astype(np.int32), 700, 512))
    offs, toks = rr.synth.corpus_tokens(40_000, 3000)
    out.append(("zipf-40000", offs, toks.astype(np.int32), 3000, 12288))
    # ragged: empty docs (also leading / trailing), invalid ids, one long doc, tile size that leaves an empty last tile region
    lens = rng.integers(0, 9, size=333)
    lens[:3] = 0
    lens[-40:] = 0
    lens[100] = 700
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    toks = rng.integers(0, 50, size=int(offs[-1])).astype(np.int32)
    toks[rng.random(toks.size) < 0.05] = -1
    toks[rng.random(toks.size) < 0.05] = 77          # >= vocab
    out.append(("ragged", offs, toks, 50, 64))
    out.append(("one-doc", np.array([0, 3], dtype=np.int64), np.array([2, 2, 1], dtype=np.int32), 4, 16))
    return out


@pytest.mark.parametrize("case", range(4))
def test_gpu_builder_equals_host_builder(case):
    import torch
    import review_recommender_b200 as rr
    eng = rr.engine
    name, offs, toks, V, tile = _corpora()[case]
    host_stats = eng.BM25Stats.local(offs, toks, V, token_pos0=1000)
    gb = eng.GpuIndexBuilder(torch.from_numpy(offs).cuda(), torch.from_numpy(toks).cuda(), V, tile)
    dev_stats = gb.local_stats(token_pos0=1000)
    np.testing.assert_array_equal(dev_stats.df, host_stats.df)
    np.testing.assert_array_equal(dev_stats.first_pos, host_stats.first_pos)
    valid = int(((toks >= 0) & (toks < V)).sum())
    assert dev_stats.n_docs == host_stats.n_docs and dev_stats.total_tokens == host_stats.total_tokens == toks.size
    assert valid <= toks.size
    host_stats.finalize()
    dev_stats.finalize()
    np.testing.assert_array_equal(dev_stats.idf, host_stats.idf)
    hp = eng.build_postings(offs, toks, host_stats, tile_docs=tile)
    dp = gb.finish(dev_stats)
    assert dp.n_tiles == hp.n_tiles
    np.testing.assert_array_equal(dp.tile_base.cpu().numpy().view(np.uint64), hp.tile_base)
    assert dp.n_freq == hp.n_freq
    np.testing.assert_array_equal(dp.term_slot.cpu().numpy(), hp.term_slot)
    np.testing.assert_array_equal(dp.rare_off.cpu().numpy().view(np.uint64), hp.rare_off)
    np.testing.assert_array_equal(dp.dir.cpu().numpy().view(np.uint32)[:hp.dir.size], hp.dir)
    np.testing.assert_array_equal(dp.fwd_off.cpu().numpy().view(np.uint64), hp.fwd_off)
    n_f = int(hp.fwd_off[-1])
    np.testing.assert_array_equal(dp.fwd_data.cpu().numpy().view(np.uint64)[:n_f], hp.fwd_data)
    np.testing.assert_array_equal(dp.data.cpu().numpy().view(np.uint64)[:hp.data.size], hp.data)


def test_index_from_device_corpus_scores_like_the_host_built_index():
    import torch
    import review_recommender_b200 as rr
    c = rr.synth.make_corpus(30_000, 64, 2000)
    a = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, 2000, c.n_reviews, c.avg_stars, make_bf16=False)
    b = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), 2000,
                              c.n_reviews, c.avg_stars, make_bf16=False)
    qt = rr.synth.query_terms(16, 4, c.doc_offsets, c.token_ids, 2000).astype(np.int32)
    nt = np.full(16, 4, dtype=np.int32)
    assert torch.equal(a.bm25_get_scores(qt, nt), b.bm25_get_scores(qt, nt))
    cand = torch.randint(0, 30_000, (16, 150), device="cuda", dtype=torch.int64)
    assert torch.equal(a.bm25_candidates(qt, nt, cand), b.bm25_candidates(qt, nt, cand))
    a.close()
    b.close()
