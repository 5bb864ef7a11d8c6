"""Host-side views of the reference's data formats on either side of the hot path.

* ``Genome``            -- what the path needs of a GDB (GDB.h:28-34): contig lengths, byte
                           offsets and the 2-bit ``.bps`` image (gene_core.c:349-400: base i of a
                           contig in bits 2*(i&3) of byte i>>2, every contig on a byte boundary).
* ``genome_from_fasta`` -- FASTA(.gz) -> Genome with Create_GDB's rules (GDB.c:350-366, :840-1030):
                           a/c/g/t -> 0..3, n/N splits contigs (NCUT = 0), any other letter is 'a'.
* ``read_gix`` / ``write_gix`` -- the ``<root>.gix`` stub + ``.<root>.ktab.<p>`` parts
                           (GIXmake.c:1503-1580, libfastk.c:815-840).
"""
import gzip
import os

import numpy as np

_LUT = np.zeros(256, dtype=np.uint8)
_LUT[ord("c")] = _LUT[ord("C")] = 1
_LUT[ord("g")] = _LUT[ord("G")] = 2
_LUT[ord("t")] = _LUT[ord("T")] = 3
_LUT[ord("n")] = _LUT[ord("N")] = 4


class Genome:
    def __init__(self, clen, bps, names=None, scaf=None, sbeg=None, slen=None):
        self.clen = np.ascontiguousarray(clen, dtype=np.int64)
        nb = (self.clen + 3) >> 2
        self.boff = np.zeros(len(clen), dtype=np.int64)
        if len(clen) > 1:
            self.boff[1:] = np.cumsum(nb)[:-1]
        self.bps = np.ascontiguousarray(bps, dtype=np.uint8)
        assert self.bps.size == int(nb.sum()), (self.bps.size, int(nb.sum()))
        self.names = names          # scaffold header strings
        self.scaf = scaf            # contig -> scaffold index
        self.sbeg = sbeg            # contig start inside its scaffold
        self.slen = slen            # scaffold lengths (with gaps)
        self._freq = None

    @property
    def ncontig(self):
        return len(self.clen)

    @property
    def seqtot(self):
        return int(self.clen.sum())

    def contig(self, c):
        """bases of contig c, one per byte (values 0..3)"""
        n = int(self.clen[c])
        b = self.bps[int(self.boff[c]): int(self.boff[c]) + ((n + 3) >> 2)]
        out = np.empty(((n + 3) >> 2) * 4, dtype=np.uint8)
        out[0::4] = b & 3
        out[1::4] = (b >> 2) & 3
        out[2::4] = (b >> 4) & 3
        out[3::4] = (b >> 6) & 3
        return out[:n]

    @property
    def freq(self):
        """base frequencies exactly as Create_GDB stores them: float((1.*count)/seqtot)"""
        if self._freq is None:
            cnt = np.zeros(4, dtype=np.int64)
            for c in range(self.ncontig):
                cnt += np.bincount(self.contig(c), minlength=4)[:4]
            tot = int(cnt.sum())
            self._freq = np.array([np.float32((1.0 * int(x)) / tot) for x in cnt], dtype=np.float32)
        return self._freq


def pack_bases(a):
    """uint8 bases 0..3 -> .bps bytes of one contig"""
    n = len(a)
    p = np.zeros(((n + 3) >> 2) * 4, dtype=np.uint8)
    p[:n] = a
    return (p[0::4] | (p[1::4] << 2) | (p[2::4] << 4) | (p[3::4] << 6)).astype(np.uint8)


def genome_from_arrays(contigs, names=None, scaf=None, sbeg=None, slen=None):
    clen = [len(c) for c in contigs]
    bps = np.concatenate([pack_bases(c) for c in contigs]) if contigs else np.zeros(0, np.uint8)
    return Genome(clen, bps, names, scaf, sbeg, slen)


def genome_from_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        data = f.read()
    raw = np.frombuffer(data, dtype=np.uint8)
    contigs, names, scaf, sbeg, slen = [], [], [], [], []
    # record boundaries
    starts = np.flatnonzero(raw == ord(">"))
    starts = starts[(starts == 0) | (raw[np.maximum(starts - 1, 0)] == ord("\n"))]
    for i, s in enumerate(starts):
        e = starts[i + 1] if i + 1 < len(starts) else len(raw)
        rec = raw[s:e]
        nl = int(np.flatnonzero(rec == ord("\n"))[0]) if (rec == ord("\n")).any() else len(rec)
        names.append(bytes(rec[1:nl]).decode())
        body = rec[nl + 1:]
        body = body[(body != ord("\n")) & (body != ord("\r"))]
        codes = _LUT[body]
        isn = codes == 4
        # maximal runs of non-N
        d = np.diff(np.concatenate([[1], isn.astype(np.int8), [1]]))
        rb = np.flatnonzero(d == -1)
        re = np.flatnonzero(d == 1)
        for b0, e0 in zip(rb, re):
            contigs.append(codes[b0:e0])
            scaf.append(i)
            sbeg.append(int(b0))
        slen.append(len(body))
    return genome_from_arrays(contigs, names, np.array(scaf, np.int32), np.array(sbeg, np.int64),
                              np.array(slen, np.int64))


def write_fasta(path, scaffolds, width=100):
    """scaffolds: list of (name, uint8 code array with 4 = N)"""
    alpha = np.frombuffer(b"acgtn", dtype=np.uint8)
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        for name, codes in scaffolds:
            f.write(b">" + name.encode() + b"\n")
            txt = alpha[codes]
            n = len(txt)
            full = (n // width) * width
            if full:
                lines = np.empty((n // width, width + 1), dtype=np.uint8)
                lines[:, :width] = txt[:full].reshape(-1, width)
                lines[:, width] = ord("\n")
                f.write(lines.tobytes())
            if n > full:
                f.write(txt[full:].tobytes() + b"\n")


# ------------------------------------------------------------------------------------------
#  GIX files
# ------------------------------------------------------------------------------------------

class GixFile:
    pass


def _hidden(path, suffix):
    d, b = os.path.split(path)
    return os.path.join(d, "." + b + suffix)


def read_gix(path):
    """path: '<dir>/<root>.gix'.  Returns GixFile with the stub fields and all entries."""
    g = GixFile()
    root = path[:-4]
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(16), dtype=np.int32)
        g.kmer, g.nparts, g.minval, g.ibyte = (int(x) for x in hdr)
        g.index = np.frombuffer(f.read(8 << 24), dtype=np.int64).copy()
        t = np.frombuffer(f.read(12), dtype=np.int32)
        g.post_bytes, g.cont_bytes, nparts2 = (int(x) for x in t)
        g.maxpre = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
        t = np.frombuffer(f.read(8), dtype=np.int32)
        g.freq, g.ncontig = int(t[0]), int(t[1])
        g.perm = np.frombuffer(f.read(4 * g.ncontig), dtype=np.int32).copy()
        g.marker = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
    g.esize = (g.kmer // 4 - 3) + 2 + g.post_bytes + g.cont_bytes
    parts, g.part_n = [], []
    for p in range(1, g.nparts + 1):
        with open(_hidden(root, ".ktab.%d" % p), "rb") as f:
            k = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
            n = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
            assert k == g.kmer
            parts.append(np.frombuffer(f.read(n * g.esize), dtype=np.uint8))
            g.part_n.append(n)
    g.entries = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    g.n = int(sum(g.part_n))
    return g


def write_gix(path, kmer, index, post_bytes, cont_bytes, maxpre, perm, entries, part_n):
    """Writes '<root>.gix' + '.<root>.ktab.<p>' in the layout GIXmake.c:1503-1580 produces."""
    root = path[:-4]
    nparts = len(part_n)
    esize = (kmer // 4 - 3) + 2 + post_bytes + cont_bytes
    with open(path, "wb") as f:
        f.write(np.array([kmer, nparts, 1, 3], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(index, dtype=np.int64).tobytes())
        f.write(np.array([post_bytes, cont_bytes, nparts], dtype=np.int32).tobytes())
        f.write(np.array([maxpre], dtype=np.int64).tobytes())
        f.write(np.array([0, len(perm)], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(perm, dtype=np.int32).tobytes())
        f.write(np.array([-1], dtype=np.int64).tobytes())
    off = 0
    for p, n in enumerate(part_n):
        with open(_hidden(root, ".ktab.%d" % (p + 1)), "wb") as f:
            f.write(np.array([kmer], dtype=np.int32).tobytes())
            f.write(np.array([n], dtype=np.int64).tobytes())
            f.write(entries[off * esize:(off + n) * esize].tobytes())
        off += n


def ksplit_from_buckets(buck1024, nparts):
    """Part boundaries over the 1024 first-5-base buckets, as distribute() picks them
    (GIXmake.c:655-691) from the sampler histogram."""
    buck = np.cumsum(np.asarray(buck1024, dtype=np.int64))
    ksplit = [0] * (nparts + 1)
    n = 1
    t = int(buck[-1]) // nparts
    for i in range(1024):
        if buck[i] >= t and n < nparts + 1:
            prev = int(buck[i - 1]) if i > 0 else 0
            if int(buck[i]) - t > t - prev:
                ksplit[n] = i
            else:
                ksplit[n] = i + 1
            n += 1
            t = (n * int(buck[-1])) // nparts
    ksplit[nparts] = 1024
    return ksplit


def gix_nparts(seqtot, ncontig, post_bytes, cont_bytes, nthreads=8, kmer=40):
    """NPARTS of GIXmake.c:1907-1917"""
    nels = 0x100000000 // (cont_bytes + post_bytes + kmer // 4 + 2)
    nbit = int((.81 * (seqtot - (kmer - 1) * ncontig)) / nels)
    nparts = ((nbit - 1) // nthreads + 1) * nthreads
    return min(max(nparts, 8), 64)


def _clz64(v):
    r = np.zeros(len(v), dtype=np.int64)
    vv = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        m = (vv >> np.uint64(64 - s)) == 0
        r[m] += s
        vv[m] = vv[m] << np.uint64(s)
    return r


def table_lcp(tab, part_first):
    """LCP byte of every entry of a sorted device-layout table (n,2 uint64 [lo,hi]): true LCP in
    bases with the predecessor, 40 for a duplicate, 0 for the first entry of a part."""
    n = len(tab)
    lcp = np.zeros(n, dtype=np.int64)
    if n > 1:
        hi, lo = tab[:, 1], tab[:, 0]
        x = hi[1:] ^ hi[:-1]
        y = (lo[1:] ^ lo[:-1]) >> np.uint64(48)
        lcp[1:] = np.where(x != 0, _clz64(x) >> 1, np.where(y != 0, 32 + ((_clz64(y) - 48) >> 1), 40))
    lcp[np.asarray(part_first, dtype=np.int64)] = 0
    return lcp


def ktab_entries_from_table(tab, post_bytes, cont_bytes, part_first):
    """numpy restatement of the .ktab entry layout (GIXmake.c:1235-1261) from device-layout records"""
    n = len(tab)
    E = 9 + post_bytes + cont_bytes
    out = np.zeros((n, E), dtype=np.uint8)
    hi, lo = tab[:, 1], tab[:, 0]
    suf = ((hi & np.uint64(0xffffffffff)) << np.uint64(16)) | (lo >> np.uint64(48))
    for k in range(7):
        out[:, k] = ((suf >> np.uint64(8 * (6 - k))) & np.uint64(0xff)).astype(np.uint8)
    out[:, 8] = table_lcp(tab, part_first).astype(np.uint8)
    post = lo & np.uint64(0xffffffff)
    for k in range(post_bytes):
        out[:, 9 + k] = ((post >> np.uint64(8 * k)) & np.uint64(0xff)).astype(np.uint8)
    cs = (lo >> np.uint64(32)) & np.uint64(0xffff)
    cv = (cs & np.uint64(0x7fff)) | ((cs >> np.uint64(15)) << np.uint64(8 * cont_bytes - 1))
    for k in range(cont_bytes):
        out[:, 9 + post_bytes + k] = ((cv >> np.uint64(8 * k)) & np.uint64(0xff)).astype(np.uint8)
    return out.reshape(-1)


def gix_bytes(genome, nthreads=8):
    """PostBytes / ContBytes of GIXmake.c:1888-1901 (ncontig padded to NTHREADS by short_GDB_fix,
    GIXmake.c:1605-1624)"""
    pb, cum = 0, 1
    while cum < int(genome.clen.max()):
        cum *= 256
        pb += 1
    nc = max(genome.ncontig, nthreads)
    cb, cum = 0, 1
    while cum < 2 * nc:
        cum *= 256
        cb += 1
    return pb, cb


def canonical_ktab(entries, esize, index):
    """.ktab entries in a run-to-run reproducible form.  Two things in the reference's output are
    not reproducible: the order of equal k-mers (its in-place MSD sort is not stable,
    MSDsort.c:285-323) -- payloads of every run of equal k-mers are put in sorted order here --
    and the LCP byte of the three entries at which the FIRST base changes: compress_thread
    temporarily stores 12 into the next panel's LCP slot (GIXmake.c:1229-1231) while another
    thread may be reading it, so those bytes come out 0 or 12 depending on timing; zeroed here.
    index = the stub's cumulative 2^24 table."""
    ent = np.ascontiguousarray(entries, dtype=np.uint8).reshape(-1, esize).copy()
    if len(ent) == 0:
        return ent.reshape(-1)
    for x in (0x3fffff, 0x7fffff, 0xbfffff):
        i = int(index[x])
        if i < len(ent):
            ent[i, 8] = 0
    run = np.cumsum(ent[:, 8] != 40)
    pay = np.zeros(len(ent), dtype=np.uint64)
    for k in range(9, esize):
        pay |= ent[:, k].astype(np.uint64) << np.uint64(8 * (k - 9))
    order = np.lexsort((pay, run))
    ent[:, 9:] = ent[order, 9:]
    return ent.reshape(-1)


# ------------------------------------------------------------------------------------------
#  .1aln output (ONEcode, ASCII flavour -- the reference's tools read both flavours)
# ------------------------------------------------------------------------------------------

def _skeleton_lines(genome, prefix="s"):
    """'g' group with S/G/C lines, as Write_Skeleton emits (GDB.c:2065-2092)"""
    out = ["g"]
    nscaf = int(genome.scaf.max()) + 1 if genome.scaf is not None and len(genome.scaf) else genome.ncontig
    for s in range(nscaf):
        if genome.scaf is not None:
            ctgs = np.flatnonzero(genome.scaf == s)
            name = genome.names[s]
            slen = int(genome.slen[s])
        else:
            ctgs, name, slen = np.array([s]), "%s_%d" % (prefix, s + 1), int(genome.clen[s])
        out.append("S %d %s" % (len(name), name))
        pos = 0
        for c in ctgs:
            b = int(genome.sbeg[c]) if genome.sbeg is not None else 0
            if b > pos:
                out.append("G %d" % (b - pos))
            out.append("C %d" % int(genome.clen[c]))
            pos = b + int(genome.clen[c])
        if slen > pos:
            out.append("G %d" % (slen - pos))
    return out


# the .1aln schema as the file states it about itself (alncode.c:19-52); ONEcode ASCII files carry
# their schema in '~' header lines, which is what lets generic readers (ONEview) parse them
_ALN_SCHEMA = """~ D t 1 3 INT
~ O g 0
~ G S 0
~ O S 1 6 STRING
~ D G 1 3 INT
~ D C 1 3 INT
~ O a 0
~ G A 0
~ D p 2 3 INT 3 INT
~ O A 6 3 INT 3 INT 3 INT 3 INT 3 INT 3 INT
~ D L 2 3 INT 3 INT
~ D R 0
~ D D 1 3 INT
~ D T 1 8 INT_LIST
~ D X 1 8 INT_LIST
~ D Q 1 3 INT
~ D E 1 3 INT
~ D Z 1 6 STRING
~ D U 1 3 INT
"""


def write_1aln_ascii(path, alns, gA, gB, gdb1="./A.1gdb", gdb2="./B.1gdb", cwd=".", tspace=100,
                     command="fastga_b200"):
    """Writes the alignments as an ASCII ONEcode '.1aln' with the schema of alncode.c:19-52:
    provenance, the two GDB references + cwd, 't' trace spacing, the GDB skeleton(s), then per
    alignment  A aread abpos aepos bread bbpos bepos / R / D diffs / T b-advances / X diffs
    (Write_Aln_Overlap, Write_Aln_Trace, alncode.c:272-305)."""
    import datetime
    stamp = datetime.datetime.now().strftime("%Y-%m-%d_%H:%M:%S")
    with open(path, "w") as f:
        f.write("1 3 aln 2 1\n")
        f.write("! 4 %d %s 3 0.1 %d %s %d %s\n" % (len("fastga_b200"), "fastga_b200", len(command), command,
                                                  len(stamp), stamp))
        f.write("< %d %s 1\n" % (len(gdb1), gdb1))
        if gdb2 is not None:
            f.write("< %d %s 2\n" % (len(gdb2), gdb2))
        f.write("< %d %s 3\n" % (len(cwd), cwd))
        f.write(_ALN_SCHEMA)
        f.write("t %d\n" % tspace)
        f.write("\n".join(_skeleton_lines(gA, "a")) + "\n")
        if gB is not None:
            f.write("\n".join(_skeleton_lines(gB, "b")) + "\n")
        for i in range(len(alns)):
            comp, ar, br, ab, bb, ae, be, df, tl = (int(x) for x in alns.fields[i])
            t = alns.trace(i)
            f.write("A %d %d %d %d %d %d\n" % (ar, ab, ae, br, bb, be))
            if comp:
                f.write("R\n")
            f.write("D %d\n" % df)
            f.write("T %d%s\n" % (tl // 2, "".join(" %d" % v for v in t[1::2])))
            f.write("X %d%s\n" % (tl // 2, "".join(" %d" % v for v in t[0::2])))
