"""Seeded synthetic inputs of the BASELINE.json shapes: grayscale images
for extraction and 256-bit descriptor sets (with planted near-duplicates) for matching.
numpy only - usable on the GPU box and here."""
import numpy as np


def _sep_blur(img, sigma):
    r = max(1, int(3 * sigma + 0.5))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img.astype(np.float64), r, mode="reflect")
    t = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(2 * r + 1))
    return sum(k[i] * t[i:i + img.shape[0], :] for i in range(2 * r + 1))


def image(kind, w, h, seed):
    """kind: 'blocks' (rectangles + noise, ~6-10k FAST candidates at 640x480, exercises the
    minThFAST fallback), 'blurnoise' (smoothed noise, under-filled top levels) or 'noise'
    (raw uniform noise: stress, ~66k candidates)."""
    rng = np.random.default_rng(seed)
    if kind == "blocks":
        img = np.full((h, w), 128.0)
        for _ in range(600 * w * h // (640 * 480)):
            bw, bh = rng.integers(8, 81, 2)
            x, y = rng.integers(-20, w), rng.integers(-20, h)
            img[max(y, 0):y + bh, max(x, 0):x + bw] = rng.integers(0, 256)
        img += rng.normal(0, 3, (h, w))
    elif kind == "blurnoise":
        img = _sep_blur(rng.integers(0, 256, (h, w)).astype(np.float64), 1.5)
        img = (img - img.min()) / (img.max() - img.min()) * 255.0
    elif kind == "noise":
        img = rng.integers(0, 256, (h, w)).astype(np.float64)
    elif kind == "flat":
        img = np.full((h, w), 77.0)
    else:
        raise ValueError(kind)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shifted_pair(kind, w, h, seed, max_shift=8):
    """Frame pair for frame-to-frame matching: image and a translated, re-noised copy."""
    rng = np.random.default_rng(seed + 7919)
    pad = max_shift
    big = image(kind, w + 2 * pad, h + 2 * pad, seed)
    dx, dy = rng.integers(-max_shift, max_shift + 1, 2)
    a = big[pad:pad + h, pad:pad + w]
    b = big[pad + dy:pad + dy + h, pad + dx:pad + dx + w].astype(np.float64)
    b = np.clip(np.rint(b + rng.normal(0, 2, b.shape)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a), b, (int(dx), int(dy))


def stereo_pair(kind, w, h, seed):
    """Right image = left shifted by a per-row-band disparity of 5..60 px."""
    rng = np.random.default_rng(seed + 104729)
    left = image(kind, w + 64, h, seed)
    right = np.empty((h, w), np.uint8)
    band = 47
    for y0 in range(0, h, band):
        d = int(rng.integers(5, 61))
        right[y0:y0 + band] = left[y0:y0 + band, d:d + w]
    return np.ascontiguousarray(left[:, :w]), right


def descriptors(n, seed, dup_from=None, max_flips=60):
    """n random 256-bit descriptors. With dup_from (another set), each row is a copy of a random
    row of dup_from with k in [0,max_flips] random bits flipped (so ratio tests accept some)."""
    rng = np.random.default_rng(seed)
    if dup_from is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    src = dup_from[rng.integers(0, len(dup_from), n)].copy()
    bits = np.unpackbits(src, axis=1)
    k = rng.integers(0, max_flips + 1, n)
    for i in range(n):
        pos = rng.choice(256, k[i], replace=False)
        bits[i, pos] ^= 1
    return np.packbits(bits, axis=1)


def descriptors_fast(n, seed, dup_from, max_flips=60):
    """Vectorised variant of descriptors(dup_from=...) for 200k-row sets: flips are drawn with
    replacement (an even number of hits on a bit cancels), which keeps the distance <= k."""
    rng = np.random.default_rng(seed)
    src = dup_from[rng.integers(0, len(dup_from), n)].copy()
    k = rng.integers(0, max_flips + 1, n)
    pos = rng.integers(0, 256, (n, max_flips))
    mask = np.arange(max_flips)[None, :] < k[:, None]
    rows = np.repeat(np.arange(n), max_flips)[mask.ravel()]
    p = pos.ravel()[mask.ravel()]
    np.bitwise_xor.at(src, (rows, p >> 3), (1 << (7 - (p & 7))).astype(np.uint8))
    return src


def vocabulary(k=10, L=3, seed=0, stop_fraction=0.02):
    """Synthetic DBoW2 vocabulary tree (the real ORBvoc.txt, k=10 L=6, is not shipped with the reference):
    nodes in breadth-first id order (root 0, children of a node contiguous), every child a noisy copy of
    its parent so that descents are meaningful. Returns dict(parent, desc, weight, k, L); leaves carry a
    positive idf-like weight, a few are 'stopped' (weight 0)."""
    rng = np.random.default_rng(seed)
    parent, desc, level = [-1], [np.zeros(32, np.uint8)], [0]
    frontier = [0]
    for lvl in range(1, L + 1):
        nxt = []
        flips = max(4, 96 >> (lvl - 1))
        for p in frontier:
            for _ in range(k):
                if lvl == 1:
                    d = rng.integers(0, 256, 32, dtype=np.uint8)
                else:
                    bits = np.unpackbits(desc[p])
                    bits[rng.choice(256, flips, replace=False)] ^= 1
                    d = np.packbits(bits)
                parent.append(p); desc.append(d); level.append(lvl)
                nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    level = np.asarray(level)
    weight = np.zeros(n, np.float64)
    leaves = level == L
    weight[leaves] = rng.uniform(0.5, 12.0, leaves.sum())
    weight[leaves & (rng.random(n) < stop_fraction)] = 0.0
    return dict(parent=np.asarray(parent, np.int32), desc=np.stack(desc).astype(np.uint8), weight=weight, k=k, L=L)


def vocabulary_fast(k=10, L=6, seed=0):
    """Vectorised variant for the full-size tree (k=10, L=6: 1 111 111 nodes) used by the benchmark."""
    rng = np.random.default_rng(seed)
    descs = [np.zeros((1, 32), np.uint8)]
    parents = [np.full(1, -1, np.int32)]
    first = 0
    for lvl in range(1, L + 1):
        prev = descs[-1]
        npar = len(prev)
        if lvl == 1:
            d = rng.integers(0, 256, (npar * k, 32), dtype=np.uint8)
        else:
            d = np.repeat(prev, k, axis=0)
            flips = max(4, 96 >> (lvl - 1))
            pos = rng.integers(0, 256, (len(d), flips))
            rows = np.repeat(np.arange(len(d)), flips)
            np.bitwise_xor.at(d, (rows, pos.ravel() >> 3), (1 << (7 - (pos.ravel() & 7))).astype(np.uint8))
        parents.append(np.repeat(np.arange(first, first + npar, dtype=np.int32), k))
        descs.append(d)
        first += npar
    desc = np.concatenate(descs)
    parent = np.concatenate(parents)
    weight = np.zeros(len(desc), np.float64)
    weight[first:] = rng.uniform(0.5, 12.0, len(desc) - first)
    return dict(parent=parent, desc=desc, weight=weight, k=k, L=L)


def write_vocabulary_text(path, voc):
    """ORBvoc text format of TemplatedVocabulary::saveToTextFile / loadFromTextFile
    (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1450): 'k L scoring weighting' then one line per node
    'parent isLeaf d0..d31 weight' in node id order."""
    parent, desc, weight = voc["parent"], voc["desc"], voc["weight"]
    n = len(parent)
    has_child = np.zeros(n, bool)
    has_child[parent[1:]] = True
    with open(path, "w") as f:
        f.write("%d %d 0 0\n" % (voc["k"], voc["L"]))  # L1_NORM, TF_IDF
        lines = ["%d %d %s %s" % (parent[i], 0 if has_child[i] else 1, " ".join(str(int(b)) for b in desc[i]), repr(float(weight[i])))
                 for i in range(1, n)]
        # no trailing newline: loadFromTextFile's `while(!f.eof())` loop would otherwise parse one empty line into a
        # bogus extra child of the root (its failed `>> pid` yields 0)
        f.write("\n".join(lines))
