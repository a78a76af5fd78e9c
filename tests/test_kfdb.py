"""keyframe database scoring. CPU: the oracle's L1 score equals the vendored DBoW2
(oracle/_ref/libref_dbow.so). GPU: orbdb_query (common words, first shared word, L1 score as bit-identical doubles) and the
DetectLoopCandidates replay against the scalar restatement of src/KeyFrameDatabase.cc:76-197."""
import numpy as np
import pytest

import oracle_lib as O


def bow_vectors(rng, n_kf, n_words, places=40, words_per_kf=180):
    """Keyframes of `places` places: each place has a word distribution; a keyframe draws its words from its place's
    favourites plus background words, with L1-normalised tf-idf-like weights (what TemplatedVocabulary::transform emits)."""
    fav = [rng.choice(n_words, 400, replace=False) for _ in range(places)]
    out, place = [], rng.integers(0, places, n_kf)
    for k in range(n_kf):
        ids = np.unique(np.concatenate([rng.choice(fav[place[k]], words_per_kf * 2 // 3), rng.integers(0, n_words, words_per_kf // 3)]))
        w = rng.uniform(0.2, 3.0, len(ids))
        out.append((ids.astype(np.int32), (w / w.sum()).astype(np.float64)))
    return out, place


@pytest.mark.skipif(not O.dbow_ref_available(), reason="libref_dbow.so absent")
def test_l1_score_equals_vendored_dbow2():
    rng = np.random.default_rng(0)
    vecs, _ = bow_vectors(rng, 60, 5000)
    for i in range(0, 60, 2):
        a, b = vecs[i], vecs[i + 1]
        assert O.l1_score(a[0], a[1], b[0], b[1]) == O.ref_l1_score(a[0], a[1], b[0], b[1])
    assert O.l1_score(vecs[0][0], vecs[0][1], vecs[0][0], vecs[0][1]) == O.ref_l1_score(vecs[0][0], vecs[0][1], vecs[0][0], vecs[0][1])
    empty = (np.zeros(0, np.int32), np.zeros(0))
    assert O.l1_score(empty[0], empty[1], vecs[0][0], vecs[0][1]) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_kf", [(0, 1500), (1, 300)])
def test_database_scores_and_loop_candidates_match_oracle(seed, n_kf):
    from multiagent_orb_slam2_b200.kfdb import KeyFrameDatabase
    rng = np.random.default_rng(seed)
    n_words = 20000
    vecs, place = bow_vectors(rng, n_kf + 8, n_words)
    db = KeyFrameDatabase(n_words)
    for ids, w in vecs[:n_kf]:
        db.add(ids, w)
    alive = np.ones(n_kf, bool)
    for slot in rng.choice(n_kf, n_kf // 20, replace=False):  # KeyFrameDatabase::erase
        db.erase(int(slot)); alive[slot] = False
    neigh = {s: [int(x) for x in rng.choice(n_kf, 10, replace=False)] for s in range(n_kf)}
    covis_state = rng.uniform(0.0, 0.05, n_kf).astype(np.float32)
    reloc_oracle = rng.uniform(0.0, 0.05, n_kf).astype(np.float32)
    reloc_device = reloc_oracle.copy()
    for qi in range(n_kf, n_kf + 8):
        q_ids, q_w = vecs[qi]
        common, first, score = db.score_all(q_ids, q_w)
        for s in range(0, n_kf, 7):
            ids, w = vecs[s]
            shared = np.intersect1d(ids, q_ids)
            if not alive[s]:
                assert common[s] == 0 and first[s] == -1
                continue
            assert common[s] == len(shared)
            assert first[s] == (shared[0] if len(shared) else -1)
            assert score[s] == O.l1_score(q_ids, q_w, ids, w)  # bit-identical double
        connected = set(int(x) for x in rng.choice(n_kf, 15, replace=False))
        min_score = 0.5 * float(np.sort(score)[-max(3, n_kf // 50)])
        want = O.detect_loop_candidates(vecs[:n_kf], alive, q_ids, q_w, min_score, connected, lambda s: neigh[s])
        got = db.DetectLoopCandidates(q_ids, q_w, min_score, connected, lambda s: neigh[s])
        assert got == want
        assert len(want) >= 1
        # the fork's covisibility query and the relocalisation query (restatements pinned to the reference in
        # tests/test_oracle_vs_reference_matcher.py), with their stale per-keyframe scores as explicit state
        ignore = [int(x) for x in rng.choice(n_kf, 12, replace=False)]
        want_c = O.detect_covisibility_candidates(vecs[:n_kf], alive, q_ids, q_w, min_score, ignore, lambda s: neigh[s], covis_state)
        assert db.DetectCovisibilityCandidates(q_ids, q_w, min_score, ignore, lambda s: neigh[s], covis_state) == want_c and len(want_c) >= 1
        want_r = O.detect_relocalization_candidates(vecs[:n_kf], alive, q_ids, q_w, lambda s: neigh[s], reloc_oracle)
        assert db.DetectRelocalizationCandidates(q_ids, q_w, lambda s: neigh[s], reloc_device) == want_r and len(want_r) >= 1
        assert np.array_equal(reloc_oracle, reloc_device)
    # a query that shares no word with anything
    none_ids = np.array([n_words - 1], np.int32)
    if not any(n_words - 1 in v[0] for v in vecs[:n_kf]):
        assert db.DetectLoopCandidates(none_ids, np.array([1.0]), 0.01) == []
