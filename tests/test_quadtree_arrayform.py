"""The sorted-key / array-list formulation of DistributeOctTree that the CUDA kernel uses must
select the same keypoints, in the same output order, as the direct std::list restatement."""
import numpy as np
import pytest

import oracle_lib as O


def random_case(rng, W, H, n, clustered):
    if clustered:
        k = rng.integers(1, 6)
        cx, cy = rng.integers(0, W, k), rng.integers(0, H, k)
        s = rng.integers(2, 40)
        which = rng.integers(0, k, n)
        x = np.clip(cx[which] + rng.normal(0, s, n), 0, W - 1).astype(np.int32)
        y = np.clip(cy[which] + rng.normal(0, s, n), 0, H - 1).astype(np.int32)
    else:
        x = rng.integers(0, W, n).astype(np.int32)
        y = rng.integers(0, H, n).astype(np.int32)
    # distinct pixels, cell-row-major-ish arbitrary candidate order
    _, first = np.unique(x.astype(np.int64) * 8192 + y, return_index=True)
    first = np.sort(first)
    x, y = x[first], y[first]
    sc = rng.integers(7, 60 if rng.random() < 0.5 else 255, len(x)).astype(np.int32)
    return x, y, sc


@pytest.mark.parametrize("seed", range(12))
def test_arrayform_equals_list_form(seed):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        W, H = [(608, 448), (1209, 344), (720, 448), (147, 102), (314, 73), (1000, 333)][rng.integers(0, 6)]
        n = int(rng.choice([0, 1, 2, 3, 7, 50, 300, 1500, 6000]))
        N = int(rng.choice([1, 5, 60, 217, 434, 1000]))
        x, y, sc = random_case(rng, W - 6, H - 6, n, rng.random() < 0.5)
        a = O.quadtree(x, y, sc, 16, 16 + W, 16, 16 + H, N)
        b = O.quadtree_arrayform(x, y, sc, 16, 16 + W, 16, 16 + H, N)
        assert np.array_equal(a, b), (W, H, n, N)


def test_arrayform_on_real_candidates():
    from multiagent_orb_slam2_b200 import synth
    for kind in ("blocks", "noise"):
        img = synth.image(kind, 752, 480, 5)
        o = O.OracleExtractor(1200, 1.2, 8)
        o(img)
        q = o.tables()["quota"]
        for l in range(8):
            L = o.level(l)
            h, w = L["img"].shape
            c = L["cand"]
            b = O.quadtree_arrayform(c[:, 0] - 16, c[:, 1] - 16, c[:, 2], 16, w - 16, 16, h - 16, int(q[l]))
            assert np.array_equal(L["sel"], b), (kind, l)
