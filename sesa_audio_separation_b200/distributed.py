"""Chunk-range sharding of one track across ranks (BASELINE config 3).

Rank r runs the model on the contiguous chunk range ``shard_chunks(n_chunks, world, r)`` and owns the padded
positions from its first chunk's start up to the next rank's first chunk's start.  Chunk k covers
``[k*step, k*step + L)``, so the owned positions of rank r+1 also receive contributions from the last
``num_overlap - 1`` chunks of rank r: rank r sends those raw partial sums (``border = L - step`` samples per
stem/channel) to rank r+1 with ONE send/recv pair per boundary — the only data-path exchange.  The window kind of
every chunk and the ``counter`` divisor are functions of the global schedule and are recomputed locally (never
exchanged).  Because the receiver SEEDS its accumulator with the sender's sums and then adds its own chunks in
ascending order, the floating-point addition order equals the reference's single loop (utils.py:439-442) and the
sharded result is bit-identical to the unsharded one.

``ops`` abstracts the two device operations so that the choreography is testable on CPU with gloo
(tests/test_distributed.py provides a torch-CPU ``ops``; the product path passes CUDA closures over
``sesa_overlap_add_range``).
"""
import torch
import torch.distributed as dist

from .plan import shard_chunks


def shard_layout(plan, world):
    """[(lo, hi, own_begin, own_end)] per rank in padded coordinates; inactive ranks get lo == hi."""
    n, L = plan.n_chunks, plan.chunk_size
    ranges = [shard_chunks(n, world, r) for r in range(world)]
    active = [r for r in range(world) if ranges[r][1] > ranges[r][0]]
    ov_reach = -(-L // plan.step) - 1            # chunks of the previous range that reach into ours
    for r in active[:-1]:
        lo, hi = ranges[r]
        if hi - lo < ov_reach:
            raise ValueError(f'chunk-range sharding needs at least {ov_reach} chunks per rank '
                             f'(rank {r} has {hi - lo}); use fewer ranks for this track')
    out = []
    for r in range(world):
        lo, hi = ranges[r]
        if hi <= lo:
            out.append((lo, hi, 0, 0))
            continue
        nxt = [a for a in active if a > r]
        begin = 0 if r == active[0] else plan.starts[lo]
        end = plan.starts[ranges[nxt[0]][0]] if nxt else plan.padded
        out.append((lo, hi, begin, end))
    return out


def sharded_overlap_add(plan, world, rank, n_rows, ops, device, group=None, gather_root=0):
    """Run the halo exchange and the owned-range finish; returns the full cropped result [n_rows, out_len] on
    ``gather_root`` (None elsewhere).  ops.raw(p0, p1) -> tensor [n_rows, p1-p0] of this rank's raw sums;
    ops.final(p0, p1, init, init_p0) -> tensor [n_rows, q1-q0] of finished samples for the cropped range
    [q0, q1) = [max(p0,crop), min(p1, crop+out_len)) - crop."""
    layout = shard_layout(plan, world)
    lo, hi, begin, end = layout[rank]
    active = [r for r in range(world) if layout[r][1] > layout[r][0]]
    crop = plan.border if plan.pad else 0
    out_len = plan.length
    L = plan.chunk_size
    reqs = []
    mine = None
    if hi > lo:
        idx = active.index(rank)
        halo = None
        halo_p0 = 0
        if idx > 0:                                   # receive the previous rank's tail sums
            prev = active[idx - 1]
            plo, phi = layout[prev][0], layout[prev][1]
            halo_p0 = begin
            halo_len = min(plan.padded, plan.starts[phi - 1] + L) - begin
            halo = torch.empty(n_rows, max(halo_len, 0), device=device, dtype=torch.float32)
            if halo_len > 0:
                reqs.append(dist.irecv(halo, src=prev, group=group))
        if idx + 1 < len(active):                     # send our tail sums to the next rank
            nxt = active[idx + 1]
            p0 = end
            p1 = min(plan.padded, plan.starts[hi - 1] + L)
            if p1 > p0:
                tail = ops.raw(p0, p1).contiguous()
                reqs.append(dist.isend(tail, dst=nxt, group=group))
        for q in reqs:
            q.wait()
        mine = ops.final(begin, end, halo if (halo is not None and halo.shape[1] > 0) else None, halo_p0)
    # gather the disjoint owned ranges on the root (send/recv; no reduction collective on the data path)
    def cropped(b, e):
        return max(b, crop) - crop, max(min(e, crop + out_len) - crop, max(b, crop) - crop)
    if rank == gather_root:
        result = torch.zeros(n_rows, out_len, device=device, dtype=torch.float32)
        for r in active:
            q0, q1 = cropped(layout[r][2], layout[r][3])
            if q1 <= q0:
                continue
            if r == rank:
                result[:, q0:q1] = mine
            else:
                buf = torch.empty(n_rows, q1 - q0, device=device, dtype=torch.float32)
                dist.recv(buf, src=r, group=group)
                result[:, q0:q1] = buf
        return result
    if hi > lo:
        q0, q1 = cropped(begin, end)
        if q1 > q0:
            dist.send(mine.contiguous(), dst=gather_root, group=group)
    return None
