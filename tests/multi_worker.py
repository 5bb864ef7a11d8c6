"""torchrun worker for tests/test_gpu_multi.py: every rank aligns its shard of genome-1 contigs
against all of genome 2 on its own GPU; rank 0 gathers the record streams over NCCL and checks the
union against a single-GPU run of the whole pair."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np          # noqa: E402
import torch                # noqa: E402
import torch.distributed as dist   # noqa: E402
from fastga_b200 import formats, lib, shard, synth   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    A, B = synth.make_pair(77, 6_000_000, 6, 0.05, sv_every=80_000)
    gA_full, gB = formats.genome_from_arrays(A), formats.genome_from_arrays(B)
    mine = shard.shard_contigs([len(a) for a in A], rank, world)
    gA = formats.genome_from_arrays([A[i] for i in mine])
    dA, dB = lib.DeviceGenome(gA, want_revcomp=True), lib.DeviceGenome(gB)
    dev = torch.device("cuda", local)
    xA = lib.DeviceGix.build(dA)
    xB = shard.build_table_cooperatively(dB, dist, dev)          # shares all-gathered over NCCL
    alns, _ = lib.align_tables(dA, dB, xA, xB, gA_full.freq)
    if rank == 0:                                                # the assembled table == a local build
        want, wps, _ = lib.DeviceGix.build(dB).download()
        got, gps, _ = xB.download()
        assert np.array_equal(got, want) and np.array_equal(gps, wps)
    merged = shard.gather_alignments(alns, np.array(mine, dtype=np.int32), dist, torch.device("cuda", local))
    if rank == 0:
        dF = lib.DeviceGenome(gA_full, want_revcomp=True)
        whole, _ = lib.align_resident(dF, dB, gA_full.freq)
        a, b = merged.canonical_lines(), whole.canonical_lines()
        assert len(a) == len(b) and a == b, (len(a), len(b))
        print("MULTI_OK world=%d records=%d" % (world, len(a)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
