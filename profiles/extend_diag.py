"""diagnostic (not a bench): the extension stage alone on the bench pair, printing the kernel time and
the pair counters (front/back wait cycles need the EX_DIAG=1 variant: FGB_LIB=.../libfastga_b200_diag.so)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fastga_b200 import formats, lib
A, B = bench.workload(1)
gA, gB = formats.genome_from_arrays(A), formats.genome_from_arrays(B)
dA, dB = lib.DeviceGenome(gA, want_revcomp=True), lib.DeviceGenome(gB)
xA, xB = lib.DeviceGix.build_forward(dA), lib.DeviceGix.build(dB)
ds = lib.DeviceSeeds.find(xA, xB, int(gA.clen.max()), int(gB.clen.max()), 10)
for it in range(3):
    lib.timings_reset()
    ov = lib.DeviceOverlaps.extend(ds, dA, dB, gA.freq)
    tm = lib.timings_get()
    c = ov.counters()
    ov.close()
c["extend_ms"] = tm["extend_ms"]; c["triples_ms"] = tm["triples_ms"]
v = c["pairings"]
c["diag_slowest"] = {"cycles": (v >> 40) << 12, "wave_cycles": ((v >> 20) & 0xfffff) << 12, "extract_cycles": (v & 0xfffff) << 12}
c["cycles_per_wave_slowest"] = c["slowest_warp"]["cycles"] / max(1, c["slowest_warp"]["waves"])
print(json.dumps(c))
