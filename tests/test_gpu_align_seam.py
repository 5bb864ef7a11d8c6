"""-m gpu: the align.h seam -- fgb_local_alignments (batched Local_Alignment) against the UNMODIFIED
reference's Local_Alignment (oracle/_ref/libfastga_ref.so) on random call tuples, borders included."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from fastga_b200 import formats, lib, synth

pytestmark = pytest.mark.gpu


class Path(C.Structure):
    _fields_ = [("trace", C.c_void_p), ("tlen", C.c_int), ("diffs", C.c_int), ("abpos", C.c_int),
                ("bbpos", C.c_int), ("aepos", C.c_int), ("bepos", C.c_int)]


class Alignment(C.Structure):
    _fields_ = [("path", C.POINTER(Path)), ("flags", C.c_uint32), ("aseq", C.c_void_p), ("bseq", C.c_void_p),
                ("alen", C.c_int), ("blen", C.c_int)]


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("borders", [False, True])
def test_batched_local_alignment_matches_reference(borders):
    rng = np.random.default_rng(41 + int(borders))
    ncont = 6
    A = [rng.integers(0, 4, int(rng.integers(3000, 40000)), dtype=np.uint8) for _ in range(ncont)]
    B, pad = [], []
    for a in A:
        rate = float(rng.choice([0.02, 0.05, 0.1, 0.15]))
        b = synth.diverged_copy(rng, a, rate, sv_every=0, inversions=False)      # small mutations only
        pad.append(int(rng.integers(0, 300)))
        B.append(np.concatenate([rng.integers(0, 4, pad[-1], dtype=np.uint8), b]))
    gA, gB = formats.genome_from_arrays(A), formats.genome_from_arrays(B)
    dA, dB = lib.DeviceGenome(gA, want_revcomp=True), lib.DeviceGenome(gB)
    jobs = []
    for _ in range(300):
        i = int(rng.integers(0, ncont))
        comp = int(rng.random() < 0.4)
        la, lb = len(A[i]), len(B[i])
        x = int(rng.integers(100, la - 100))
        y = int(np.clip(x + (lb - la) + int(rng.integers(-60, 60)), 50, lb - 50))
        if comp:                       # the strand-C call sees reverse-complemented A: any diagonal will do
            y = int(rng.integers(50, lb - 50))
        d, anti = x - y, x + y
        low, hgh = d - int(rng.integers(0, 70)), d + int(rng.integers(0, 70))
        lbd = hbd = -1
        if borders:
            lbd = int(rng.integers(0, 40)) if rng.random() < 0.7 else -1
            hbd = int(rng.integers(0, 40)) if rng.random() < 0.7 else -1
        jobs.append((i, i, comp, low, hgh, anti, lbd, hbd))
    jobs = np.array(jobs, dtype=np.int32)
    paths, toff, traces = lib.local_alignments(dA, dB, jobs, gA.freq)

    ref = C.CDLL(ol.REF_SO)
    ref.New_Work_Data.restype = C.c_void_p
    ref.New_Align_Spec.restype = C.c_void_p
    ref.New_Align_Spec.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_float), C.c_int]
    ref.Local_Alignment.argtypes = [C.POINTER(Alignment), C.c_void_p, C.c_void_p] + [C.c_int] * 5
    freq = (C.c_float * 4)(*[float(v) for v in gA.freq])
    work = ref.New_Work_Data()
    spec = ref.New_Align_Spec(0.7, 100, freq, 0)
    fA = [ol._framed(a) for a in A]
    fAC = [ol._framed(3 - a[::-1]) for a in A]
    fB = [ol._framed(b) for b in B]
    nonempty = 0
    for q, (i, j, comp, low, hgh, anti, lbd, hbd) in enumerate(jobs.tolist()):
        a, b = (fAC[i] if comp else fA[i]), fB[j]
        p = Path()
        al = Alignment(C.pointer(p), 2 if comp else 0, a.ctypes.data + 1, b.ctypes.data + 1, len(a) - 2, len(b) - 2)
        assert ref.Local_Alignment(C.byref(al), work, spec, low, hgh, anti, lbd, hbd) == 0
        rt = np.ctypeslib.as_array(C.cast(p.trace, C.POINTER(C.c_uint16)), shape=(max(p.tlen, 1),))[:p.tlen]
        got = paths[q]
        assert got[6] == 0, (q, got)
        assert (p.abpos, p.bbpos, p.aepos, p.bepos, p.diffs, p.tlen) == tuple(int(v) for v in got[:6]), (q, jobs[q])
        assert np.array_equal(rt.astype(np.uint8), traces[int(toff[q]):int(toff[q]) + p.tlen]), q
        nonempty += int(p.aepos > p.abpos)
    assert nonempty > 100
