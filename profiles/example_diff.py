"""diagnostic: EXAMPLE (HAP1 x HAP2) through the CUDA path vs the reference binary run on the same
box; prints the canonical records that differ."""
import os, sys, tempfile, shutil
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle_lib as ol
from fastga_b200 import formats, lib
EX = '/root/repo/tests/data/EXAMPLE'
gA = formats.genome_from_fasta(os.path.join(EX, 'HAP1.fasta.gz'))
gB = formats.genome_from_fasta(os.path.join(EX, 'HAP2.fasta.gz'))
alns, stats = lib.fastga(gA, gB)
mine = alns.canonical_lines()
print('cuda', len(mine), ol.md5_lines(mine), stats['nseeds'], stats['nhits'], alns.nraw)
if len(sys.argv) > 1:
    sys.exit(0)
wd = tempfile.mkdtemp()
for n in ('HAP1', 'HAP2'):
    shutil.copy(os.path.join(EX, n + '.fasta.gz'), os.path.join(wd, n + '.fasta.gz'))
ol.ref_fastga(wd, 'HAP1', 'HAP2', threads=32)
ref = ol.oneview_records(os.path.join(wd, 'ref.1aln'))
print('ref', len(ref), ol.md5_lines(ref))
a, b = set(mine), set(ref)
for l in sorted(a - b)[:10]:
    print('ONLY CUDA', l[:400])
for l in sorted(b - a)[:10]:
    print('ONLY REF ', l[:400])
