"""CPU tests of the host side: formats, C-ABI surface, sharding + gather (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from fastga_b200 import formats, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import fastga_b200
    lib = fastga_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "fastga_b200.h")).read()
    names = set(re.findall(r"\b(fgb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 35
    for n in sorted(names):
        assert hasattr(lib, n), "libfastga_b200.so does not export %s" % n


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import fastga_b200
    monkeypatch.setattr(fastga_b200, "_lib", None)
    monkeypatch.setattr(fastga_b200, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(fastga_b200.LibraryMissing):
        fastga_b200.load_library()


def test_no_product_module_touches_the_oracle():
    pk = os.path.join(ROOT, "fastga_b200")
    for f in os.listdir(pk):
        if f.endswith(".py") and f != "smoke.py":
            src = open(os.path.join(pk, f)).read()
            assert "oracle" not in src.replace("oracle/_ref", ""), f


def test_fasta_roundtrip_and_contig_split(tmp_path):
    rng = np.random.default_rng(1)
    c1, c2, c3 = (rng.integers(0, 4, n, dtype=np.uint8) for n in (1001, 37, 4000))
    scaf = [("s1 first", np.concatenate([c1, np.full(7, 4, np.uint8), c2])), ("s2", c3)]
    p = str(tmp_path / "x.fasta.gz")
    formats.write_fasta(p, scaf, width=60)
    g = formats.genome_from_fasta(p)
    assert list(g.clen) == [1001, 37, 4000]
    assert list(g.scaf) == [0, 0, 1] and list(g.sbeg) == [0, 1008, 0]
    assert g.names == ["s1 first", "s2"]
    for i, c in enumerate((c1, c2, c3)):
        assert np.array_equal(g.contig(i), c)
    assert list(g.boff) == [0, 251, 261]
    assert abs(float(g.freq.sum()) - 1.0) < 1e-6


def test_gix_file_roundtrip(tmp_path):
    rng = np.random.default_rng(2)
    n, pb, cb = 1000, 3, 1
    E = 9 + pb + cb
    ent = rng.integers(0, 256, n * E, dtype=np.uint8)
    index = np.sort(rng.integers(0, n + 1, 1 << 24)).astype(np.int64)
    p = str(tmp_path / "g.gix")
    formats.write_gix(p, 40, index, pb, cb, 123, np.arange(5, dtype=np.int32), ent, [400, 600])
    g = formats.read_gix(p)
    assert (g.kmer, g.nparts, g.post_bytes, g.cont_bytes, g.maxpre, g.marker) == (40, 2, pb, cb, 123, -1)
    assert np.array_equal(g.entries, ent) and np.array_equal(g.index, index) and g.part_n == [400, 600]


def test_ktab_lcp_and_canonical_form():
    tab = np.array([[0x0001 << 48 | 5, 0x10], [0x0001 << 48 | 9, 0x10], [0x0002 << 48 | 1, 0x10],
                    [3, 0x4000000000000011]], dtype=np.uint64)
    lcp = formats.table_lcp(tab, [0])
    assert list(lcp) == [0, 40, 39, 0]


def test_shard_contigs_partition():
    lens = [50, 10, 40, 30, 20, 60, 5]
    for world in (1, 2, 3, 8):
        parts = [shard.shard_contigs(lens, r, world) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lens)))
        loads = [sum(lens[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lens)


def test_kmer_space_ownership_covers_everything_once():
    """the sharded path's two ownership maps: k-mer prefix ranges (by the first four bases) tile the
    prefix space, contigs are spread by length"""
    for world in (1, 2, 3, 4, 8):
        cuts = shard.top_byte_cuts(world)
        assert cuts[0] == 0 and cuts[-1] == 256 and all(a <= b for a, b in zip(cuts, cuts[1:]))
        assert max(b - a for a, b in zip(cuts, cuts[1:])) - min(b - a for a, b in zip(cuts, cuts[1:])) <= 1
        lens = [50, 10, 40, 30, 20, 60, 5, 33]
        own = shard.owner_of_contigs(lens, world)
        assert set(own.tolist()) <= set(range(world))
        loads = [sum(l for l, o in zip(lens, own) if o == r) for r in range(world)]
        assert max(loads) - min(loads) <= max(lens)


_XWORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from fastga_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)                       # same stream on every rank: everybody knows every block
send = rng.integers(0, 50, (world, world))           # send[src][dst] rows
send[0, 1] = 0                                       # an empty block
blocks = [[rng.integers(-2**62, 2**62, (int(send[s, d]), 2)) for d in range(world)] for s in range(world)]
mine = np.concatenate(blocks[rank]) if send[rank].sum() else np.zeros((0, 2), np.int64)
got = shard.exchange_rows(dist, torch.from_numpy(mine.astype(np.int64)), [int(v) for v in send[rank]])
want = np.concatenate([blocks[s][rank] for s in range(world)])
assert got.shape == want.shape and np.array_equal(got.numpy(), want), (rank, got.shape, want.shape)
print("XCHG_OK", rank)
dist.destroy_process_group()
'''


def test_record_exchange_gloo_world2(tmp_path):
    """the all-to-all of 16-byte records that moves k-mer records and seeds between ranks (N > 1 path)"""
    script = tmp_path / "xworker.py"
    script.write_text(_XWORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29583", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("XCHG_OK") == 2, r.stdout[-3000:]


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from fastga_b200 import lib, shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(100 + rank)
n = 5 + 3 * rank
fields = rng.integers(0, 1000, (n, 9)).astype(np.int32)
fields[:, 1] = rng.integers(0, 2, n)          # local contig numbers
fields[:, 8] = 2 * rng.integers(1, 6, n)      # tlen
toff = np.concatenate([[0], np.cumsum(fields[:-1, 8])]).astype(np.int64)
pool = rng.integers(0, 256, int(fields[:, 8].sum()), dtype=np.uint8)
al = lib.Alignments(fields, toff, pool, n)
cmap = np.array([10 * rank, 10 * rank + 1], dtype=np.int32)
merged = shard.gather_alignments(al, cmap, dist, torch.device("cpu"))
np.save(os.path.join(%(out)r, "lines_%%d.npy" %% rank), np.array(
    [l.replace("A %%d " %% int(f[1]), "A %%d " %% int(cmap[f[1]]), 1) for l, f in zip(
        lib.Alignments(fields, toff, pool, n).canonical_lines_unsorted(), fields)], dtype=object), allow_pickle=True)
if rank == 0:
    assert len(merged) == sum(5 + 3 * r for r in range(world))
    key = merged.fields[:, [1, 3, 2, 0]]
    assert all(tuple(key[i]) <= tuple(key[i + 1]) for i in range(len(key) - 1))
    np.save(os.path.join(%(out)r, "merged.npy"), np.array(merged.canonical_lines(), dtype=object), allow_pickle=True)
dist.destroy_process_group()
'''


def test_gather_alignments_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "out": str(tmp_path)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29581")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29581", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    merged = list(np.load(tmp_path / "merged.npy", allow_pickle=True))
    want = sorted(list(np.load(tmp_path / "lines_0.npy", allow_pickle=True)) +
                  list(np.load(tmp_path / "lines_1.npy", allow_pickle=True)))
    assert merged == want
