"""torchrun worker for tests/test_gpu_multi.py: the k-mer-space sharded path (shard.align_sharded:
every rank scans its contigs, k-mer records and seeds are exchanged with NCCL all-to-alls, every rank
extends the seeds of its A-contigs); rank 0 gathers the record streams over NCCL and checks the union
against a single-GPU run of the whole pair."""
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
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    A, B = synth.make_pair(77, 6_000_000, 6, 0.05, sv_every=80_000)
    gA, gB = formats.genome_from_arrays(A), formats.genome_from_arrays(B)
    dA, dB = lib.DeviceGenome(gA, want_revcomp=True), lib.DeviceGenome(gB)
    alns, st = shard.align_sharded(dA, dB, gA.freq, dist, dev)
    tot = torch.tensor([st["nkmers1_fwd"], st["nkmers2"], st["nseeds_merged"], st["nseeds"], st["nhits"]],
                       dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    merged = shard.gather_alignments(alns, None, dist, dev)
    if rank == 0:
        whole, ws = lib.align_resident(dA, dB, gA.freq)
        t = [int(v) for v in tot.tolist()]
        assert t[0] == ws["nkmers1_fwd"] and t[1] == ws["nkmers2"], (t, ws["nkmers1_fwd"], ws["nkmers2"])
        assert t[2] == t[3] == ws["nseeds"], (t, ws["nseeds"])
        assert t[4] == ws["nhits"], (t, ws["nhits"])
        a, b = merged.canonical_lines(), whole.canonical_lines()
        assert len(a) == len(b) and a == b, (len(a), len(b))
        assert merged.nraw == whole.nraw
        print("MULTI_OK world=%d records=%d" % (world, len(a)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
