"""Chunk-range sharding of one track across ranks (BASELINE config 3).

Rank r runs the model on the contiguous chunk range ``shard_chunks(n_chunks, world, r)`` and owns the padded
positions from its first chunk's start up to the next rank's first chunk's start.  Chunk k covers
``[k*step, k*step + L)``, so the first ``span - 1`` step-long regions owned by rank r+1 (``span = ceil(L / step)``) also
receive contributions from the last ``span - 1`` chunks of rank r: rank r sends those raw partial sums
(``~(L - step)`` samples per stem/channel) to rank r+1 with ONE send/recv pair per boundary — the only data-path
exchange.  The window kind of every chunk and the window-sum divisor are functions of the global schedule and are
recomputed locally (never exchanged).  Because the receiver CONTINUES from the sender's sums and then adds its own chunks
in ascending order, the floating-point addition order equals the reference's single loop (utils.py:439-442) and the
sharded result is bit-identical to the unsharded one.

Schedule of one rank (``run_sharded_track``), built so that no rank ever waits for a neighbour's model passes:

1. TAIL FIRST: run the model on the rank's LAST chunks (at least ``span - 1`` of them) and fold them into the regions the
   next rank owns (those depend on nothing else); the tail outputs are kept;
2. ONE grouped exchange (ncclGroupStart / ncclSend to the next rank / ncclRecv from the previous rank / ncclGroupEnd):
   the halos cross NVLink while every rank is still busy with the rest of its range;
3. walk the remaining chunks in ascending engine batches, each folded into the running sums as soon as its forward is
   done (``sesa_overlap_accumulate``); the first fold continues from the received halo;
4. fold the kept tail outputs into the rank's own regions (ascending order is preserved: they are its last chunks);
5. gather the disjoint finished ranges on the root: one grouped set of point-to-point receives straight into the rows
   of the result (no staging buffer, no reduction collective).

``ops`` abstracts the device work so that the choreography is testable on CPU with gloo (tests/test_distributed.py
provides a torch-CPU ``ops``; the product path passes CUDA closures over the C-ABI, demix.py).
"""
import torch
import torch.distributed as dist

from .plan import shard_chunks


def shard_layout(plan, world, strict=False):
    """[(lo, hi, own_begin, own_end)] per rank in padded coordinates; inactive ranks get lo == hi.

    Every rank but the last must hold at least ``span - 1`` chunks (what the next rank's first regions depend on).  A track
    too short for ``world`` ranks is sharded over the largest number of ranks that satisfies this and the remaining ranks
    stay idle (every rank derives the same layout from the schedule alone); ``strict=True`` raises instead."""
    n, L = plan.n_chunks, plan.chunk_size
    ov_reach = -(-L // plan.step) - 1            # chunks of the previous range that reach into ours
    for w in range(world, 0, -1):
        ranges = [shard_chunks(n, w, r) if r < w else (n, n) for r in range(world)]
        active = [r for r in range(world) if ranges[r][1] > ranges[r][0]]
        short = [r for r in active[:-1] if ranges[r][1] - ranges[r][0] < ov_reach]
        if not short:
            break
        if strict:
            r = short[0]
            raise ValueError(f'chunk-range sharding needs at least {ov_reach} chunks per rank '
                             f'(rank {r} has {ranges[r][1] - ranges[r][0]}); use fewer ranks for this track')
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


def cropped_range(plan, begin, end):
    """Owned padded range [begin, end) -> the range [q0, q1) of result samples it finishes (after the border crop)."""
    crop = plan.border if plan.pad else 0
    q0 = min(max(begin - crop, 0), plan.length)
    q1 = min(max(end - crop, q0), plan.length)
    return q0, q1


def _peer(group, r):
    """Rank r OF THE GROUP as the global rank torch.distributed's point-to-point calls expect."""
    return dist.get_global_rank(group, r) if group is not None else r


def tail_size(n_own, span, engine_batch):
    """Chunks a non-last rank runs first: at least span-1 (everything the next rank's regions depend on), and sized so
    that the remaining chunks split into full engine batches."""
    rem = n_own % engine_batch or engine_batch
    return min(n_own, max(span - 1, rem))


def run_sharded_track(plan, world, rank, ops, engine_batch, group=None, stats=None):
    """One rank's share of a chunk-range-sharded track.  ``ops`` provides
      forward(k0, nb, keep) -> y            model outputs of chunks [k0, k0+nb) ([nb, rows, L]); keep=True: own storage
      accumulate(y, k0, nb, r0, r1)         fold into regions [r0, r1) (sesa_overlap_accumulate)
      read_partial(p0, p1) -> tensor        contiguous copy [rows, p1-p0] of the running sums
      seed_partial(p0, tensor)              store received sums at padded position p0
      empty(rows, cols) -> tensor           scratch on the ops' device
      mark(name)                            optional timing hook
    and leaves the finished owned range in ops' output buffer.  Returns (outstanding send request, its buffer) or
    (None, None); the caller waits on the request before releasing the buffer."""
    layout = shard_layout(plan, world)
    lo, hi, begin, end = layout[rank]
    if hi <= lo:
        return None, None
    active = [r for r in range(world) if layout[r][1] > layout[r][0]]
    idx = active.index(rank)
    prev = active[idx - 1] if idx > 0 else None
    nxt = active[idx + 1] if idx + 1 < len(active) else None
    step, L = plan.step, plan.chunk_size
    span = -(-L // step)
    n_regions = -(-plan.padded // step)
    mark = getattr(ops, 'mark', lambda name: None)

    halo, recv_op, reqs_in = None, None, []
    if prev is not None:
        halo_p1 = min(plan.padded, (lo + span - 1) * step)
        if halo_p1 > begin:
            halo = ops.empty(ops.rows, halo_p1 - begin)
            recv_op = dist.P2POp(dist.irecv, halo, _peer(group, prev), group)

    def seed():
        if reqs_in:
            mark('halo_wait_begin')
            for q in reqs_in:
                q.wait()
            reqs_in.clear()
            ops.seed_partial(begin, halo)
            mark('halo_wait_end')

    n_tail, y_tail, req_out, sent = 0, None, None, None
    if nxt is not None:
        n_tail = tail_size(hi - lo, span, engine_batch)
        y_tail = ops.forward(hi - n_tail, n_tail, True)
        r1 = min(hi + span - 1, n_regions)
        ops.accumulate(y_tail, hi - n_tail, n_tail, hi, r1)
        p1 = min(plan.padded, r1 * step)
        if p1 > end:
            sent = ops.read_partial(end, p1)
            if stats is not None:
                stats['halo_bytes'] = sent.numel() * sent.element_size()
    # ONE grouped exchange per rank (ncclGroupStart/End): receive the previous rank's halo, send ours to the next.  Every
    # rank reaches this point after its first (tail) batch, so nothing waits on a neighbour's remaining model passes.
    p2p = ([recv_op] if recv_op is not None else []) + \
          ([dist.P2POp(dist.isend, sent, _peer(group, nxt), group)] if sent is not None else [])
    if p2p:
        reqs = dist.batch_isend_irecv(p2p)
        if recv_op is not None and sent is not None and len(reqs) == 2:
            reqs_in, req_out = [reqs[0]], reqs[1]
        elif recv_op is not None:
            reqs_in = list(reqs)              # backends that coalesce a group hand back one request for all of it
            req_out = reqs[-1] if sent is not None else None
        else:
            req_out = reqs[-1]
    k = lo
    while k < hi - n_tail:
        nb = min(engine_batch, hi - n_tail - k)
        y = ops.forward(k, nb, False)
        seed()
        ops.accumulate(y, k, nb, k, min(k + nb + span - 1, n_regions if nxt is None else hi))
        k += nb
    if n_tail:
        seed()
        ops.accumulate(y_tail, hi - n_tail, n_tail, hi - n_tail, hi)
    return req_out, sent


def gather_owned(plan, world, rank, out, out_q0, make_result, group=None, gather_root=0):
    """Collect every rank's finished range on ``gather_root``.  ``out`` is this rank's [rows, >= q1-q0] buffer whose
    column 0 is result sample ``out_q0``; ``make_result()`` allocates the full [rows, length] result on the root (it may
    alias ``out``).  Receives go row by row straight into the result (each row slice is contiguous) as one grouped set of
    point-to-point operations; there is no reduction collective on the data path."""
    layout = shard_layout(plan, world)
    active = [r for r in range(world) if layout[r][1] > layout[r][0]]
    ops = []
    result = None
    if rank == gather_root:
        result = make_result()
        for r in active:
            q0, q1 = cropped_range(plan, layout[r][2], layout[r][3])
            if q1 <= q0:
                continue
            if r == rank:
                if out is not None and not (out.data_ptr() == result.data_ptr() and out_q0 == 0):
                    result[:, q0:q1] = out[:, q0 - out_q0:q1 - out_q0]
                continue
            for row in range(result.shape[0]):
                ops.append(dist.P2POp(dist.irecv, result[row, q0:q1], _peer(group, r), group))
    elif layout[rank][1] > layout[rank][0]:
        q0, q1 = cropped_range(plan, layout[rank][2], layout[rank][3])
        if q1 > q0:
            for row in range(out.shape[0]):
                ops.append(dist.P2POp(dist.isend, out[row, q0 - out_q0:q1 - out_q0], _peer(group, gather_root), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return result
