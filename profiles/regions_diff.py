import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle_lib as ol
from fastga_b200 import formats, lib
z = np.load('/root/repo/tests/golden/example_regions.npz')
A = [z['a%d' % i] for i in range(10)]; B = [z['b%d' % i] for i in range(10)]
gA = formats.genome_from_arrays(A); gB = formats.genome_from_arrays(B)
r = ol.oracle_pipeline(gA, gB)
alns, stats = lib.fastga(gA, gB)
mine = alns.canonical_lines()
a, b = set(mine), set(r['lines'])
print('cuda', len(mine), 'oracle', len(r['lines']), 'only cuda', len(a - b), 'only oracle', len(b - a))
