import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle_lib as ol
from fastga_b200 import formats, lib
g = '/root/repo/tests/golden/'
a = np.load(g + 'example_mini_a.npy'); b = np.load(g + 'example_mini_b.npy')
gA = formats.genome_from_arrays([a]); gB = formats.genome_from_arrays([b])
r = ol.oracle_pipeline(gA, gB)
alns, stats = lib.fastga(gA, gB)
mine = alns.canonical_lines()
print(len(mine), len(r['lines']), stats['nseeds'], r['nseeds'])
A, B = set(mine), set(r['lines'])
for l in sorted(A - B): print('ONLY CUDA', l[:300])
for l in sorted(B - A): print('ONLY ORAC', l[:300])
