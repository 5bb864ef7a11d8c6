"""Edge-case inputs shared by the GPU parity tests and the CPU oracle-pin tests: ragged and tiny
contigs, unrelated genomes (empty result), identical and reverse-complemented copies, tandem
repeats (frequency cutoff, wide bands), many small contigs.  Each case returns
(A contigs, B contigs, reference threads, check(alignments, hits))."""
import numpy as np

from fastga_b200 import synth


def _distinct(contigs):
    seen, out = set(), []
    for c in contigs:
        while len(c) in seen:
            c = c[:-1]
        seen.add(len(c))
        out.append(c)
    return out


def ragged():
    rng = np.random.default_rng(51)
    core = rng.integers(0, 4, 150_003, dtype=np.uint8)
    A = _distinct([core[:70_001], core[70_001:], rng.integers(0, 4, 39, dtype=np.uint8),
                   rng.integers(0, 4, 11, dtype=np.uint8), rng.integers(0, 4, 41, dtype=np.uint8)])
    mut = synth._small_mutations(rng, core, 0.04)
    B = _distinct([mut[:33_333], mut[33_333:], rng.integers(0, 4, 13, dtype=np.uint8)])
    return A, B, 4, lambda al, hits: len(al) > 0


def unrelated():
    rng = np.random.default_rng(52)
    A = _distinct([rng.integers(0, 4, 300_000, dtype=np.uint8), rng.integers(0, 4, 200_000, dtype=np.uint8)])
    B = _distinct([rng.integers(0, 4, 400_000, dtype=np.uint8)])
    def check(al, hits):
        assert len(al) == 0 and hits == 0
    return A, B, 4, check


def identical_and_revcomp():
    rng = np.random.default_rng(53)
    a1 = rng.integers(0, 4, 400_000, dtype=np.uint8)
    a2 = rng.integers(0, 4, 250_001, dtype=np.uint8)
    B = _distinct([a1.copy()[:399_990], (3 - a2[::-1]).astype(np.uint8)])
    def check(al, hits):
        assert {int(x) for x in al.fields[:, 0]} == {0, 1}
    return [a1, a2], B, 4, check


def tandem_repeats():
    rng = np.random.default_rng(54)
    parts = []
    while sum(len(p) for p in parts) < 600_000:
        parts.append(rng.integers(0, 4, int(rng.integers(20_000, 60_000)), dtype=np.uint8))
        unit = rng.integers(0, 4, int(rng.integers(20, 400)), dtype=np.uint8)
        parts.append(np.concatenate([synth._small_mutations(rng, unit, 0.02)
                                     for _ in range(int(rng.integers(20, 120)))]))
    a = np.concatenate(parts)[:600_000]
    B = [synth.diverged_copy(rng, a, 0.04, sv_every=80_000)]
    return [a], B, 4, lambda al, hits: len(al) > 0


def many_small_contigs():
    rng = np.random.default_rng(55)
    base = rng.integers(0, 4, 1_200_000, dtype=np.uint8)
    cuts = np.sort(rng.choice(np.arange(2_000, 1_198_000), 119, replace=False))
    A = _distinct(list(np.split(base, cuts)))
    mut = synth._small_mutations(rng, base, 0.08)
    cuts2 = np.sort(rng.choice(np.arange(2_000, len(mut) - 2_000), 60, replace=False))
    B = _distinct(list(np.split(mut, cuts2)))
    return A, B, 8, lambda al, hits: len(al) > 50


CASES = {"ragged": ragged, "unrelated": unrelated, "identical_and_revcomp": identical_and_revcomp,
         "tandem_repeats": tandem_repeats, "many_small_contigs": many_small_contigs}
