"""-m gpu: every CUDA stage against the oracle on the same seeded input, through the C-ABI."""
import numpy as np
import pytest

import oracle_lib as ol
from fastga_b200 import lib

pytestmark = pytest.mark.gpu


def test_sort128_matches_stable_numpy_sort():
    rng = np.random.default_rng(5)
    for n in (0, 1, 31, 4096, 4097, 100_003, 1_500_000):
        recs = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64)
        recs[:, 1] &= np.uint64(0xffff)            # many equal high keys -> exercises stability
        recs[:, 0] &= np.uint64(0xffff0000ffffffff)
        want = recs.copy()
        key_hi = want[:, 1]
        key_lo = want[:, 0] >> np.uint64(48)
        order = np.lexsort((key_lo, key_hi))        # lexsort is stable
        want = want[order]
        got = lib.sort128_host(recs.copy(), 6, 16)
        assert np.array_equal(got, want), n


def test_sort128_full_key_random():
    rng = np.random.default_rng(6)
    recs = rng.integers(0, 1 << 63, size=(300_000, 2), dtype=np.uint64)
    got = lib.sort128_host(recs.copy(), 0, 16)
    order = np.lexsort((recs[:, 0], recs[:, 1]))
    assert np.array_equal(got, recs[order])


def test_staged_genome_and_revcomp(small_pair):
    gA, _ = small_pair
    dg = lib.DeviceGenome(gA, want_revcomp=True)
    words, woff = dg.download(False)
    rwords, _ = dg.download(True)
    for c in range(gA.ncontig):
        a = gA.contig(c)
        n = len(a)
        w = words[woff[c]:woff[c + 1]]
        bits = ((w[:, None] >> (np.arange(32, dtype=np.uint64) * np.uint64(2))) & np.uint64(3)).reshape(-1)
        assert np.array_equal(bits[:n].astype(np.uint8), a)
        assert not bits[n:].any()
        rw = rwords[woff[c]:woff[c + 1]]
        rbits = ((rw[:, None] >> (np.arange(32, dtype=np.uint64) * np.uint64(2))) & np.uint64(3)).reshape(-1)
        assert np.array_equal(rbits[:n].astype(np.uint8), (3 - a[::-1]))
        assert not rbits[n:].any()


def test_gix_build_matches_oracle(small_pair):
    for g in small_pair:
        dg = lib.DeviceGenome(g)
        perm, rank = ol.contig_rank(g.clen)
        assert np.array_equal(dg.perm, perm)
        want, wstart = ol.gix_build(g, rank)
        gx = lib.DeviceGix.build(dg)
        tab, pstart, _ = gx.download()
        assert gx.n == len(want)
        assert np.array_equal(tab, want)
        assert np.array_equal(pstart, wstart)


def test_merge_and_seed_sort_match_oracle(small_pair):
    gA, gB = small_pair
    dA, dB = lib.DeviceGenome(gA), lib.DeviceGenome(gB)
    xA, xB = lib.DeviceGix.build(dA), lib.DeviceGix.build(dB)
    amx, bmx = int(gA.clen.max()), int(gB.clen.max())
    ds = lib.DeviceSeeds.find(xA, xB, amx, bmx, 10)
    tA, _, _ = xA.download(False)
    tB, pB, _ = xB.download()
    seeds, sumlen = ol.merge(tA, tB, pB, 10)
    assert ds.n == len(seeds) and ds.sumlen == sumlen
    want = ol.seed_records(seeds, ds.layout + (amx, bmx), sort=True)
    got = ds.download()
    assert np.array_equal(got, want)


def _canon(recs, pool, aread=None, bread=None, comp=None):
    out = []
    for i, r in enumerate(recs):
        tr = bytes(pool[int(r["toff"]):int(r["toff"]) + int(r["tlen"])])
        out.append((int(comp[i]) if comp is not None else int(r["comp"]),
                    int(aread[i]) if aread is not None else int(r["aread"]),
                    int(bread[i]) if bread is not None else int(r["bread"]),
                    int(r["abpos"]), int(r["bbpos"]), int(r["aepos"]), int(r["bepos"]), int(r["diffs"]), tr))
    return out


def test_extend_matches_oracle(small_pair):
    gA, gB = small_pair
    dA, dB = lib.DeviceGenome(gA, want_revcomp=True), lib.DeviceGenome(gB)
    xA, xB = lib.DeviceGix.build(dA), lib.DeviceGix.build(dB)
    amx, bmx = int(gA.clen.max()), int(gB.clen.max())
    ds = lib.DeviceSeeds.find(xA, xB, amx, bmx, 10)
    seeds = ds.download()
    layout = ds.layout + (amx, bmx)
    want, wpool, whits = ol.search(seeds, layout, gA, gB, dA.perm, dB.perm, gA.freq)
    ov = lib.DeviceOverlaps.extend(ds, dA, dB, gA.freq)
    got, gpool = ov.records()
    cnt = ov.counters()
    assert cnt["hits"] == whits
    # pairkey = (comp, icont rank, jcont rank) packed as in the seed record
    jb, ib = layout[2], layout[3]
    pk = got["pairkey"].astype(np.int64)
    jc = pk & ((1 << jb) - 1)
    ic = (pk >> jb) & ((1 << ib) - 1)
    comp = pk >> (jb + ib)
    g = _canon(got, gpool, dA.perm[ic], dB.perm[jc], comp)
    w = _canon(want, wpool)
    assert len(g) == len(w)
    assert g == w          # same records in the same (reference discovery) order
