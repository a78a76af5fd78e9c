"""DBoW2 vocabulary transform (Frame::ComputeBoW): oracle restatement vs the reference's vendored DBoW2 compiled
unmodified (CPU), CUDA descent + host assembly vs the oracle (GPU). Integer outputs bit-exact; BowVector values
compared as float64 bit patterns (same summation order)."""
import numpy as np
import pytest

import oracle_lib as O
from multiagent_orb_slam2_b200 import synth


def features(voc, n, seed):
    """Noisy copies of random leaf descriptors (so that descents are non-trivial) plus pure noise."""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(~np.isin(np.arange(len(voc["parent"])), voc["parent"][1:]))
    src = voc["desc"][rng.choice(leaves, n)]
    f = synth.descriptors_fast(n, seed + 1, src, 30)
    f[: n // 10] = rng.integers(0, 256, (n // 10, 32), dtype=np.uint8)
    return f


@pytest.mark.skipif(not O.dbow_ref_available(), reason="oracle/_ref/libref_dbow.so not built (reference tree absent)")
@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 2), (6, 4, 4), (10, 2, 4), (3, 5, 1)])
def test_oracle_equals_vendored_dbow2(tmp_path, k, L, levelsup):
    voc = synth.vocabulary(k, L, seed=k * 10 + L)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, voc)
    ref = O.RefVocabulary(path)
    assert ref.size() == k ** L
    orc = O.OracleVocabulary(voc)
    f = features(voc, 700, 5)
    for a, b in zip(orc.descend(f, levelsup), ref.descend(f, levelsup)):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    ob, ov, ofv = orc.transform(f, levelsup)
    rb, rv, rfv = ref.transform(f, levelsup)
    assert np.array_equal(ob, rb) and np.array_equal(ov.view(np.uint64), rv.view(np.uint64)) and ofv == rfv
    assert abs(ov.sum() - 1.0) < 1e-12 and len(ob) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 2), (10, 4, 4), (16, 2, 1), (20, 2, 0), (5, 4, 3)])
def test_cuda_transform_equals_oracle(k, L, levelsup):
    from multiagent_orb_slam2_b200.vocabulary import ORBVocabulary
    voc = synth.vocabulary(k, L, seed=k + L)
    orc = O.OracleVocabulary(voc)
    g = ORBVocabulary(voc)
    for n in (1, 33, 1007):
        f = features(voc, n, n)
        for a, b in zip(g.descend(f, levelsup), orc.descend(f, levelsup)):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        gb, gv, gfv = g.transform(f, levelsup)
        ob, ov, ofv = orc.transform(f, levelsup)
        assert np.array_equal(gb, ob) and np.array_equal(gv.view(np.uint64), ov.view(np.uint64)) and gfv == ofv


@pytest.mark.gpu
def test_cuda_loads_orbvoc_text_format(tmp_path):
    from multiagent_orb_slam2_b200.vocabulary import ORBVocabulary
    voc = synth.vocabulary(10, 3, seed=1)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, voc)
    g = ORBVocabulary()
    assert g.loadFromTextFile(path)
    f = features(voc, 500, 9)
    a = g.descend(f, 2)
    b = O.OracleVocabulary(voc).descend(f, 2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
