"""__graft_entry__.smoke(): one small pass of the whole hot path on cuda:0, checked against the
oracle (the only place outside tests/ and bench.py that may touch oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from . import formats, lib, synth

    if not lib.device_ready():
        raise RuntimeError("fastga_b200 smoke: no CUDA device (there is no CPU fallback)")
    A, B = synth.make_pair(3, 400_000, 2, 0.05, sv_every=50_000)
    gA, gB = formats.genome_from_arrays(A), formats.genome_from_arrays(B)
    alns, stats = lib.fastga(gA, gB)

    # oracle: same path on the CPU
    pa, ra = ol.contig_rank(gA.clen)
    pb, rb = ol.contig_rank(gB.clen)
    tA, _ = ol.gix_build(gA, ra)
    tB, sB = ol.gix_build(gB, rb)
    seeds, _ = ol.merge(tA, tB, sB)
    amx, bmx = int(gA.clen.max()), int(gB.clen.max())
    ab = int(amx + bmx).bit_length()
    layout = (ab, max(ab - 6, 1), int(max(gB.ncontig - 1, 1)).bit_length(),
              int(max(gA.ncontig - 1, 1)).bit_length(), amx, bmx)
    recs = ol.seed_records(seeds, layout)
    ov, tp, nhit = ol.search(recs, layout, gA, gB, pa, pb, gA.freq)
    assert stats["nkmers1"] == len(tA) and stats["nkmers2"] == len(tB), "GIX size mismatch"
    assert stats["nseeds"] == len(seeds), "seed count mismatch"
    assert stats["nhits"] == nhit, "chain hit count mismatch"
    packed = ol.pack_overlaps(ov, tp, ra, rb, layout[2], layout[3])     # keep the handle alive
    raw = lib.filter_overlaps(packed.h, pa, pb, layout[2], layout[3])
    assert raw.canonical_lines() == alns.canonical_lines(), "alignment records differ from the oracle"
    print("smoke ok: %d k-mers, %d seeds, %d hits, %d alignments (bit-exact vs oracle)"
          % (stats["nkmers1"] + stats["nkmers2"], stats["nseeds"], stats["nhits"], len(alns)))
