"""Regenerates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref) -- only
possible where /root/reference was available to build oracle/_ref.  Committed with its output.

For each case: a synthetic pair from fastga_b200.synth (pure function of the seed), the reference's
own -v counters (seeds / hits / aln's / non-redundant), the md5 of the canonical-sorted ONEview dump
of its .1aln, and the md5s of its GIX files' content (entries + index), so that the oracle -- and
through it the CUDA path -- can be pinned on a box where /root/reference does not exist.
"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np          # noqa: E402
import oracle_lib as ol     # noqa: E402
from fastga_b200 import formats, synth   # noqa: E402

CASES = {
    "pair_a": dict(seed=21, total=300_000, ncontig=2, div=0.05, sv=40_000, per_scaffold=1),
    "pair_b": dict(seed=22, total=800_000, ncontig=4, div=0.10, sv=50_000, per_scaffold=2),
    "pair_c": dict(seed=23, total=500_000, ncontig=3, div=0.02, sv=0, per_scaffold=1),
}


def main():
    out = {}
    for name, c in CASES.items():
        A, B = synth.make_pair(c["seed"], c["total"], c["ncontig"], c["div"], sv_every=c["sv"])
        with tempfile.TemporaryDirectory() as wd:
            formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(A, "sa", c["per_scaffold"]))
            formats.write_fasta(os.path.join(wd, "B.fasta"), synth.scaffolds_of(B, "sb", c["per_scaffold"]))
            log = ol.ref_fastga(wd, "A", "B", threads=4)
            st = ol.parse_fastga_log(log)
            recs = ol.oneview_records(os.path.join(wd, "ref.1aln"))
            g = {}
            for nm in ("A", "B"):
                gx = formats.read_gix(os.path.join(wd, nm + ".gix"))
                g[nm] = {"n": gx.n, "post_bytes": gx.post_bytes, "cont_bytes": gx.cont_bytes,
                         "nparts": gx.nparts, "part_n": [int(x) for x in gx.part_n],
                         "entries_md5": hashlib.md5(formats.canonical_ktab(gx.entries, gx.esize, gx.index).tobytes()).hexdigest(),
                         "index_md5": hashlib.md5(gx.index.tobytes()).hexdigest(),
                         "bps_md5": hashlib.md5(open(os.path.join(wd, "." + nm + ".bps"), "rb").read()).hexdigest()}
        out[name] = {"case": c, "counters": st, "aln_md5": ol.md5_lines(recs), "aln_records": len(recs),
                     "first_records": recs[:3], "gix": g}
        print(name, st, len(recs))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
