"""Multi-GPU plumbing of the path: contig sharding (no data-path collective) and the gather of the
per-rank alignment record streams (the only collective)."""
import numpy as np


def shard_contigs(lengths, rank, world):
    """Greedy length balance of genome-1 contigs over ranks, longest first (the reference balances
    its A-contig panels by bp the same way, FastGA.c:5067-5085).  Deterministic on every rank."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    load = [0] * world
    owner = np.zeros(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[int(i)] = r
        load[r] += int(lengths[int(i)])
    return [i for i in range(len(lengths)) if owner[i] == rank]


def pack_alignments(alns, contig_map=None):
    """Alignments -> one uint8 buffer (fields | toff | pool); contig_map renumbers the shard's
    local A-contig numbers to global ones."""
    fields = alns.fields.copy()
    if contig_map is not None and len(fields):
        fields[:, 1] = np.asarray(contig_map, dtype=np.int32)[fields[:, 1]]
    n = np.array([len(alns), alns.pool.size if len(alns) else 0], dtype=np.int64)
    pool = alns.pool if len(alns) else np.zeros(0, np.uint8)
    return np.concatenate([n.view(np.uint8), fields.reshape(-1).view(np.uint8),
                           alns.toff.view(np.uint8), pool.view(np.uint8)])


def unpack_alignments(buf):
    from .lib import Alignments
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n, pb = (int(x) for x in buf[:16].view(np.int64))
    o = 16
    fields = buf[o:o + n * 36].view(np.int32).reshape(n, 9).copy()
    o += n * 36
    toff = buf[o:o + n * 8].view(np.int64).copy()
    o += n * 8
    pool = buf[o:o + pb].copy()
    return Alignments(fields, toff, pool if pb else np.zeros(1, np.uint8), n)


def gather_alignments(alns, contig_map, dist, device):
    """Variable-length gather of every rank's records on rank 0 (NCCL on GPUs, gloo in tests).
    Returns the merged Alignments in (aread, abpos, bread, comp) order on rank 0, None elsewhere."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = torch.from_numpy(pack_alignments(alns, contig_map)).to(device)
    size = torch.tensor([payload.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, size)
    mx = max(int(s.item()) for s in sizes)
    pad = torch.zeros(mx, dtype=torch.uint8, device=device)
    pad[:payload.numel()] = payload
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    parts = [unpack_alignments(b[:int(s.item())].cpu().numpy()) for b, s in zip(bufs, sizes)]
    return merge_alignments(parts)


def merge_alignments(parts):
    """Concatenates per-shard results and restores the global SORT_MAP order (FastGA.c:3800-3836)."""
    from .lib import Alignments
    fields = np.concatenate([p.fields for p in parts]) if parts else np.zeros((0, 9), np.int32)
    pools, toffs, base = [], [], 0
    for p in parts:
        toffs.append(p.toff + base)
        pl = p.pool[:int((p.toff + p.fields[:, 8]).max())] if len(p) else np.zeros(0, np.uint8)
        pools.append(pl)
        base += len(pl)
    toff = np.concatenate(toffs) if toffs else np.zeros(0, np.int64)
    pool = np.concatenate(pools) if pools else np.zeros(1, np.uint8)
    order = np.lexsort((fields[:, 0], fields[:, 2], fields[:, 3], fields[:, 1]))
    return Alignments(fields[order], toff[order], pool if pool.size else np.zeros(1, np.uint8),
                      sum(p.nraw for p in parts))


def build_table_cooperatively(dgenome, dist, device):
    """Every rank scans the whole genome but keeps, sorts and indexes only the k-mers whose
    12-base prefix falls in its 1/N slice of the prefix space; the sorted shares concatenate in
    rank order into the complete table, exchanged with one NCCL all-gather over NVLink.  This is
    the one real exchange step of the path when genome 2 is needed by every rank."""
    import torch
    from .lib import DeviceGix
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return DeviceGix.build(dgenome)
    plo, phi = (rank << 24) // world, ((rank + 1) << 24) // world
    share = DeviceGix.build_range(dgenome, plo, phi)
    n_local = share.n
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=device))
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    local = torch.empty(mx * 16, dtype=torch.uint8, device=device)
    share.copy_table_to(local.data_ptr())
    pb, cb = share.post_bytes, share.cont_bytes
    share.close()
    gathered = torch.empty(world * mx * 16, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, local)
    total = sum(sizes)
    if all(s == mx for s in sizes):
        full = gathered
    else:
        full = torch.empty(total * 16, dtype=torch.uint8, device=device)
        o = 0
        for r, s in enumerate(sizes):
            full[o:o + s * 16] = gathered[r * mx * 16: r * mx * 16 + s * 16]
            o += s * 16
    torch.cuda.synchronize()
    gx = DeviceGix.from_device(full.data_ptr(), total, pb, cb, dgenome.genome.ncontig)
    del gathered, full, local
    return gx
