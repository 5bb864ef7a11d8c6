"""CPU tests that pin the oracle (oracle/fastga_oracle.c) to the UNMODIFIED reference:
  * committed golden vectors (tests/golden/reference_golden.json, made by tests/golden/make_golden.py
    from runs of oracle/_ref) -- work without the reference;
  * live comparisons against oracle/_ref (libfastga_ref.so Local_Alignment, FastGA binaries) when
    that directory has been built (skipped otherwise)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from fastga_b200 import formats, synth

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.json")))


def _pair(c):
    A, B = synth.make_pair(c["seed"], c["total"], c["ncontig"], c["div"], sv_every=c["sv"])
    return formats.genome_from_arrays(A), formats.genome_from_arrays(B)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_reference_golden(name):
    g = GOLD[name]
    gA, gB = _pair(g["case"])
    r = ol.oracle_pipeline(gA, gB)
    cnt = g["counters"]
    assert r["nseeds"] == cnt["seeds"]
    assert abs(r["sumlen"] / r["nseeds"] - cnt["avelen"]) < 0.051
    assert r["nhit"] == cnt["hits"]
    assert r["nraw"] == cnt["alns"]
    assert len(r["lines"]) == cnt["kept"] == g["aln_records"]
    assert r["lines"][:3] == g["first_records"]
    assert ol.md5_lines(r["lines"]) == g["aln_md5"]          # bit-exact .1aln content
    for nm, genome, tab, pstart in (("A", gA, r["tabA"], r["pstartA"]), ("B", gB, r["tabB"], r["pstartB"])):
        gg = g["gix"][nm]
        assert hashlib.md5(genome.bps.tobytes()).hexdigest() == gg["bps_md5"]
        assert len(tab) == gg["n"]
        pb, cb = formats.gix_bytes(genome)
        assert (pb, cb) == (gg["post_bytes"], gg["cont_bytes"])
        index = pstart[1:].astype(np.int64)                  # stub index = cumulative counts
        assert hashlib.md5(index.tobytes()).hexdigest() == gg["index_md5"]
        part_first = np.cumsum([0] + gg["part_n"][:-1])
        ent = formats.ktab_entries_from_table(tab, pb, cb, part_first)
        E = 9 + pb + cb                                      # .ktab bytes, equal k-mers canonicalised
        assert hashlib.md5(formats.canonical_ktab(ent, E, index).tobytes()).hexdigest() == gg["entries_md5"]


# ---------------------------------------------------------------------------------------------
#  live pins against oracle/_ref
# ---------------------------------------------------------------------------------------------

class Path(C.Structure):
    _fields_ = [("trace", C.c_void_p), ("tlen", C.c_int), ("diffs", C.c_int), ("abpos", C.c_int),
                ("bbpos", C.c_int), ("aepos", C.c_int), ("bepos", C.c_int)]


class Alignment(C.Structure):
    _fields_ = [("path", C.POINTER(Path)), ("flags", C.c_uint32), ("aseq", C.c_void_p), ("bseq", C.c_void_p),
                ("alen", C.c_int), ("blen", C.c_int)]


class OPath(C.Structure):
    _fields_ = [("abpos", C.c_int), ("bbpos", C.c_int), ("aepos", C.c_int), ("bepos", C.c_int),
                ("diffs", C.c_int), ("tlen", C.c_int), ("trace", C.POINTER(C.c_uint8)), ("tmax", C.c_int)]


def _mutate(rng, a, rate):
    out, i = [], 0
    r = rng.random(len(a) * 2 + 8)
    q = 0
    while i < len(a):
        u = r[q]
        q += 1
        if u < rate * 0.8:
            out.append((a[i] + rng.integers(1, 4)) % 4)
            i += 1
        elif u < rate * 0.9:
            out.append(rng.integers(0, 4))
        elif u < rate:
            i += 1
        else:
            out.append(a[i])
            i += 1
    return np.array(out, dtype=np.int8)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_local_alignment_bit_exact_vs_reference_library(seed):
    _la_compare(seed, borders=False)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [11, 12])
def test_local_alignment_with_band_borders_vs_reference_library(seed):
    """lbord / hbord >= 0 confine the band to [low-lbord, hgh+hbord] (align.c:1466-1481): what
    align_contigs passes for a contig against itself (FastGA.c:3247-3262); the CUDA waves take the
    same minp/maxp arguments but FastGA's non-self mode never sets them"""
    _la_compare(seed, borders=True)


def _la_compare(seed, borders):
    ref = C.CDLL(ol.REF_SO)
    orc = ol.orc()
    ref.New_Work_Data.restype = C.c_void_p
    ref.New_Align_Spec.restype = C.c_void_p
    ref.New_Align_Spec.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_float), C.c_int]
    ref.Local_Alignment.argtypes = [C.POINTER(Alignment), C.c_void_p, C.c_void_p] + [C.c_int] * 5
    orc.orc_local_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + \
        [C.c_int] * 6 + [C.POINTER(OPath)]
    freq = (C.c_float * 4)(.25, .25, .25, .25)
    work = ref.New_Work_Data()
    spec = ref.New_Align_Spec(0.7, 100, freq, 0)
    ospec, _tabs, _ = ol.make_spec(np.array([.25] * 4, np.float32), 0.7)
    owork = C.c_void_p(orc.orc_new_work())
    rng = np.random.default_rng(seed)
    for it in range(400):
        L = int(rng.integers(300, 5000))
        rate = float(rng.choice([0.0, 0.02, 0.05, 0.1, 0.15, 0.3]))
        core = rng.integers(0, 4, L).astype(np.int8)
        fa, fb, ta, tb = (rng.integers(0, 4, int(rng.integers(0, 400))).astype(np.int8) for _ in range(4))
        if rng.random() < 0.3:
            fa = fa[:0]
        if rng.random() < 0.3:
            tb = tb[:0]
        a = np.concatenate([fa, core, ta])
        b = np.concatenate([fb, _mutate(rng, core, rate), tb])
        acomp = int(rng.random() < 0.5)
        ab, bb = ol._framed(a), ol._framed(b)
        xa, xb = len(fa) + L // 2, len(fb) + L // 2
        d, anti = xa - xb, xa + xb + int(rng.integers(-100, 100))
        low, hgh = d - int(rng.integers(0, 80)), d + int(rng.integers(0, 80))
        lb = hb = -1
        if borders:
            lb = int(rng.integers(0, 40)) if rng.random() < 0.7 else -1
            hb = int(rng.integers(0, 40)) if rng.random() < 0.7 else -1
        p = Path()
        al = Alignment(C.pointer(p), 2 if acomp else 0, ab.ctypes.data + 1, bb.ctypes.data + 1, len(a), len(b))
        assert ref.Local_Alignment(C.byref(al), work, spec, low, hgh, anti, lb, hb) == 0
        rt = np.ctypeslib.as_array(C.cast(p.trace, C.POINTER(C.c_uint16)), shape=(max(p.tlen, 1),))[:p.tlen]
        rt = rt.astype(np.uint8)
        op = OPath()
        orc.orc_local_alignment(owork, C.byref(ospec), ab.ctypes.data + 1, len(a), bb.ctypes.data + 1, len(b),
                                acomp, low, hgh, anti, lb, hb, C.byref(op))
        ot = np.ctypeslib.as_array(op.trace, shape=(max(op.tlen, 1),))[:op.tlen] if op.tlen else np.zeros(0, np.uint8)
        assert (p.abpos, p.bbpos, p.aepos, p.bepos, p.diffs, p.tlen) == \
               (op.abpos, op.bbpos, op.aepos, op.bepos, op.diffs, op.tlen), (it, L, rate, acomp, lb, hb)
        assert np.array_equal(rt, ot), (it, L, rate, acomp, lb, hb)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_oracle_vs_live_reference_run(tmp_path):
    A, B = synth.make_pair(31, 600_000, 3, 0.07, sv_every=30_000)
    wd = str(tmp_path)
    formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(A, "sa", 1))
    formats.write_fasta(os.path.join(wd, "B.fasta"), synth.scaffolds_of(B, "sb", 3))
    st = ol.parse_fastga_log(ol.ref_fastga(wd, "A", "B", threads=4))
    ref = ol.oneview_records(os.path.join(wd, "ref.1aln"))
    gA = formats.genome_from_fasta(os.path.join(wd, "A.fasta"))
    gB = formats.genome_from_fasta(os.path.join(wd, "B.fasta"))
    clen, names, scaf, sbeg = ol.read_gdb_ascii(os.path.join(wd, "B.1gdb"))
    assert np.array_equal(clen, gB.clen) and np.array_equal(sbeg, gB.sbeg) and np.array_equal(scaf, gB.scaf)
    r = ol.oracle_pipeline(gA, gB)
    assert (r["nseeds"], r["nhit"], r["nraw"], len(r["lines"])) == (st["seeds"], st["hits"], st["alns"], st["kept"])
    assert r["lines"] == ref


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_written_1aln_is_read_by_reference_tools(tmp_path):
    """a .1aln written by formats.write_1aln_ascii is accepted by the reference's
    ONEview (same records back) and -- after ONEview -b adds the binary index the threaded readers
    need -- by its ALNtoPAF next to the reference-made GDBs (same PAF as from the reference's own
    .1aln)."""
    A, B = synth.make_pair(41, 500_000, 3, 0.06, sv_every=40_000)
    wd = str(tmp_path)
    formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(A, "sa", 2))
    formats.write_fasta(os.path.join(wd, "B.fasta"), synth.scaffolds_of(B, "sb", 1))
    ol.ref_fastga(wd, "A", "B", threads=4)
    gA = formats.genome_from_fasta(os.path.join(wd, "A.fasta"))
    gB = formats.genome_from_fasta(os.path.join(wd, "B.fasta"))
    r = ol.oracle_pipeline(gA, gB)
    formats.write_1aln_ascii(os.path.join(wd, "mine.1aln"), r["alns"], gA, gB, "./A.1gdb", "./B.1gdb", wd)
    assert ol.oneview_records(os.path.join(wd, "mine.1aln")) == ol.oneview_records(os.path.join(wd, "ref.1aln"))
    paf_ref = sorted(ol.run_ref(["ALNtoPAF", "-T2", "ref"], cwd=wd).split("\n"))
    # the threaded converters need ONEcode's binary index: ONEview -b turns the ASCII file into it
    ol.run_ref(["ONEview", "-b", "-o", "mineb.1aln", "mine.1aln"], cwd=wd)
    paf_mine = sorted(ol.run_ref(["ALNtoPAF", "-T2", "mineb"], cwd=wd).split("\n"))
    assert len(paf_ref) > 3 and paf_mine == paf_ref


import edge_cases  # noqa: E402


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(edge_cases.CASES))
def test_oracle_edge_cases_vs_live_reference(name, tmp_path):
    """the oracle on the edge-case inputs of tests/edge_cases.py, against a live reference run"""
    A, B, threads, check = edge_cases.CASES[name]()
    wd = str(tmp_path)
    formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(A, "sa", 1))
    formats.write_fasta(os.path.join(wd, "B.fasta"), synth.scaffolds_of(B, "sb", 1))
    st = ol.parse_fastga_log(ol.ref_fastga(wd, "A", "B", threads=threads))
    ref = ol.oneview_records(os.path.join(wd, "ref.1aln"))
    r = ol.oracle_pipeline(formats.genome_from_arrays(A), formats.genome_from_arrays(B))
    assert r["nseeds"] == st.get("seeds", 0)
    assert r["lines"] == ref
    check(r["alns"], r["nhit"])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_oracle_on_example_regions_vs_live_reference(tmp_path):
    """the EXAMPLE regions fixture of the GPU regression test, oracle against a live reference run"""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_regions.npz"))
    A = [z["a%d" % i] for i in range(10)]
    B = [z["b%d" % i] for i in range(10)]
    wd = str(tmp_path)
    formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(A, "sa", 1))
    formats.write_fasta(os.path.join(wd, "B.fasta"), synth.scaffolds_of(B, "sb", 1))
    st = ol.parse_fastga_log(ol.ref_fastga(wd, "A", "B", threads=4))
    ref = ol.oneview_records(os.path.join(wd, "ref.1aln"))
    r = ol.oracle_pipeline(formats.genome_from_arrays(A), formats.genome_from_arrays(B))
    assert r["nseeds"] == st.get("seeds", 0)
    assert r["lines"] == ref


def _self_genomes():
    """genomes with internal homology for SELF mode (FastGA A): a diverged duplication inside and
    across contigs plus an inverted copy, and two tandem-repeat genomes (near-diagonal chains: the
    'nothing across the main diagonal' branch and the band borders of align_contigs)"""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 4, 300_000, dtype=np.uint8)
    dup = synth.diverged_copy(rng, a[50_000:150_000], 0.05, sv_every=40_000)
    c0 = np.concatenate([a, dup, rng.integers(0, 4, 20_000, dtype=np.uint8)])
    c1 = np.concatenate([rng.integers(0, 4, 100_000, dtype=np.uint8),
                         synth.diverged_copy(rng, a[200_000:280_000], 0.04, sv_every=30_000),
                         (3 - a[10_000:60_000][::-1]).astype(np.uint8)])
    out = {"dup": [c0, c1[:len(c1) - 3]]}
    for seed in (61, 62):
        rng = np.random.default_rng(seed)
        parts = []
        while sum(len(p) for p in parts) < 500_000:
            parts.append(rng.integers(0, 4, int(rng.integers(20_000, 60_000)), dtype=np.uint8))
            unit = rng.integers(0, 4, int(rng.integers(150, 1200)), dtype=np.uint8)
            parts.append(np.concatenate([synth._small_mutations(rng, unit, 0.03)
                                         for _ in range(int(rng.integers(8, 60)))]))
        g = np.concatenate(parts)[:500_000]
        cut = int(rng.integers(200_000, 300_000))
        out["tandem%d" % seed] = [g[:cut], g[cut:]]
    # C-strand self pairs, identical contigs, a palindrome, a long near-diagonal repeat array
    rng = np.random.default_rng(101)
    a = rng.integers(0, 4, 200_000, dtype=np.uint8)
    inv = (3 - a[30_000:90_000][::-1]).astype(np.uint8)
    out["inverted_dup"] = [np.concatenate([a, rng.integers(0, 4, 5000, dtype=np.uint8),
                                           synth._small_mutations(rng, inv, 0.03)])]
    out["identical_contigs"] = [a.copy(), a.copy()[:199_990], rng.integers(0, 4, 50_000, dtype=np.uint8)]
    out["palindrome"] = [np.concatenate([a[:60_000], (3 - a[:60_000][::-1]).astype(np.uint8)]),
                         rng.integers(0, 4, 30_001, dtype=np.uint8)]
    rng = np.random.default_rng(102)
    unit = rng.integers(0, 4, 5000, dtype=np.uint8)
    out["near_diagonal_repeats"] = [np.concatenate(
        [rng.integers(0, 4, 50_000, dtype=np.uint8)] +
        [synth._small_mutations(rng, unit, 0.04) for _ in range(30)] +
        [rng.integers(0, 4, 50_000, dtype=np.uint8)])]
    return out


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["dup", "tandem61", "tandem62", "inverted_dup", "identical_contigs",
                                  "palindrome", "near_diagonal_repeats"])
def test_oracle_self_mode_vs_live_reference(name, tmp_path):
    """SURVEY row a-7 groundwork: the oracle's SELF mode (self block rule, band borders for a contig
    against itself) against `FastGA A` of the reference"""
    G = _self_genomes()[name]
    wd = str(tmp_path)
    formats.write_fasta(os.path.join(wd, "A.fasta"), synth.scaffolds_of(G, "sa", 1))
    st = ol.parse_fastga_log(ol.ref_fastga(wd, "A", None, threads=4))
    ref = ol.oneview_records(os.path.join(wd, "ref.1aln"))
    r = ol.oracle_pipeline_self(formats.genome_from_arrays(G))
    # every thread of the reference halves its own pair count (FastGA.c:1907): off by < #threads
    assert abs(r["nseeds"] // 2 - st["seeds"]) < 4
    assert r["nhit"] == st["hits"] and r["nraw"] == st["alns"]
    assert r["lines"] == ref
