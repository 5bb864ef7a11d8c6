"""Synthetic genome pairs for the benchmark configurations of BASELINE.json.

genome  : i.i.d. bases, contigs of pairwise-distinct lengths, 200-N gaps between contigs of a
          scaffold.
copy    : per-base substitutions / short insertions / short deletions at total rate ``div``
          (80 % / 10 % / 10 %, indel length geometric with mean 2), structural indels of
          0.5-5 kbp about every ``sv_every`` bases (they stop a wave, align.c:546, giving
          alignments of realistic length) and a few reverse-complemented segments >= 50 kbp so
          the C strand is exercised.  Everything is a pure function of the seed (PCG64).
"""
import numpy as np


def random_contigs(rng, total_bp, ncontig):
    """ncontig pairwise-distinct lengths summing to ~total_bp"""
    w = rng.uniform(0.5, 1.5, ncontig)
    lens = np.maximum((w / w.sum() * total_bp).astype(np.int64), 1000)
    lens = lens + np.arange(ncontig)          # make ties impossible
    assert len(set(lens.tolist())) == ncontig
    return [rng.integers(0, 4, int(n), dtype=np.uint8) for n in lens]


def _small_mutations(rng, a, div):
    n = len(a)
    if n == 0 or div <= 0:
        return a.copy()
    sub = rng.random(n) < 0.8 * div
    a = a.copy()
    a[sub] = (a[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
    # deletion events: start prob 0.05*div, geometric length (mean 2)
    keep = np.ones(n, dtype=bool)
    ds = np.flatnonzero(rng.random(n) < 0.05 * div)
    dl = rng.geometric(0.5, len(ds))
    for k in range(1, int(dl.max()) + 1 if len(ds) else 0):
        idx = ds[dl >= k] + (k - 1)
        keep[idx[idx < n]] = False
    nins = np.zeros(n, dtype=np.int64)
    isel = np.flatnonzero(rng.random(n) < 0.05 * div)
    nins[isel] = rng.geometric(0.5, len(isel))
    reps = keep.astype(np.int64) + nins
    src = np.repeat(np.arange(n), reps)
    starts = np.cumsum(reps) - reps
    first = (np.arange(len(src)) - starts[src] == 0) & keep[src]
    out = rng.integers(0, 4, len(src), dtype=np.uint8)
    out[first] = a[src[first]]
    return out


def _revcomp(a):
    return (3 - a[::-1]).astype(np.uint8)


def diverged_copy(rng, contig, div, sv_every=200000, inversions=True):
    n = len(contig)
    pieces = []
    pos = 0
    while pos < n:
        seg = int(rng.integers(sv_every // 2, sv_every * 3 // 2)) if sv_every > 0 else n
        end = min(n, pos + seg)
        piece = _small_mutations(rng, contig[pos:end], div)
        if inversions and end - pos >= 60000 and rng.random() < 0.05:
            piece = _revcomp(piece)
        pieces.append(piece)
        pos = end
        if pos < n and sv_every > 0:
            svl = int(rng.integers(500, 5000))
            if rng.random() < 0.5:
                pos = min(n, pos + svl)                                  # structural deletion
            else:
                pieces.append(rng.integers(0, 4, svl, dtype=np.uint8))   # structural insertion
    return np.concatenate(pieces) if pieces else contig[:0].copy()


def make_pair(seed, total_bp, ncontig, div, sv_every=200000, inversions=True):
    """Returns (contigsA, contigsB): lists of uint8 base arrays (B contig i is the diverged copy
    of A contig i, emitted in a shuffled order)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = random_contigs(rng, total_bp, ncontig)
    B = [diverged_copy(rng, a, div, sv_every, inversions) for a in A]
    order = rng.permutation(ncontig)
    B = [B[i] for i in order]
    # pairwise-distinct lengths in B as well (libc qsort tie order would otherwise matter)
    seen = set()
    for i, b in enumerate(B):
        while len(b) in seen:
            b = b[:-1]
        seen.add(len(b))
        B[i] = b
    return A, B


def scaffolds_of(contigs, prefix, per_scaffold=1, gap=200):
    """Groups contigs into scaffolds with N gaps; returns [(name, codes-with-4-for-N)]"""
    out = []
    for s in range(0, len(contigs), per_scaffold):
        grp = contigs[s:s + per_scaffold]
        parts = []
        for i, c in enumerate(grp):
            if i:
                parts.append(np.full(gap, 4, dtype=np.uint8))
            parts.append(c)
        out.append(("%s_%d" % (prefix, s // per_scaffold + 1), np.concatenate(parts)))
    return out
