import sys,time; sys.path.insert(0,'/root/repo')
import numpy as np
from fastga_b200 import formats, lib, synth
import bench
A,B=bench.workload(1)
gA=formats.genome_from_arrays(A); gB=formats.genome_from_arrays(B)
dA=lib.DeviceGenome(gA,want_revcomp=True); dB=lib.DeviceGenome(gB)
xA=lib.DeviceGix.build(dA); xB=lib.DeviceGix.build(dB)
for i in range(3):
    lib.timings_reset()
    ds=lib.DeviceSeeds.find(xA,xB,int(gA.clen.max()),int(gB.clen.max()),10)
    t=lib.timings_get()
    print(t['merge_ms'], t['ssort_ms'], ds.n)
