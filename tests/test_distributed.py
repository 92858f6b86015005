"""CPU, world_size 2 and 3 over gloo: the chunk-range sharding choreography of distributed.py (tail-first halo
send/recv, streamed overlap-add, owned ranges, grouped gather on the root, sub-groups) against the single-process
oracle overlap-add.  ``CpuOps`` below is test infrastructure: a torch-CPU statement of what the CUDA closures over
sesa_overlap_accumulate do (demix.py), region by region with the same seeded / complete rules."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import demix as odemix


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class CpuOps:
    """forward / accumulate / partial-sum access of one rank, with the kernel's ascending-chunk accumulation."""

    def __init__(self, plan, y_all, nrows, window, part_p0, part_len, out, out_q0):
        self.plan, self.y_all, self.rows, self.window = plan, y_all, nrows, window
        self.part_p0, self.partial = part_p0, torch.full((nrows, part_len), float('nan'))   # never read before written
        self.out, self.out_q0 = out, out_q0
        self.forwards = []

    def _w(self, k):
        from sesa_audio_separation_b200.plan import KIND_NO_FADEIN, KIND_NO_FADEOUT
        w = self.window.clone()
        fade = self.plan.fade
        if self.plan.kinds[k] == KIND_NO_FADEIN:
            w[:fade] = 1
        elif self.plan.kinds[k] == KIND_NO_FADEOUT:
            w[-fade:] = 1
        return w

    def forward(self, k0, nb, keep):
        self.forwards.append((k0, nb))
        return self.y_all[k0:k0 + nb].clone()

    def empty(self, r, c):
        return torch.empty(r, c)

    def read_partial(self, p0, p1):
        return self.partial[:, p0 - self.part_p0:p1 - self.part_p0].contiguous()

    def seed_partial(self, p0, t):
        self.partial[:, p0 - self.part_p0:p0 - self.part_p0 + t.shape[1]] = t

    def accumulate(self, y, k0, nb, r0, r1):
        plan = self.plan
        step, L, n = plan.step, plan.chunk_size, plan.n_chunks
        span = -(-L // step)
        crop = plan.border if plan.pad else 0
        for r in range(r0, min(r1, -(-plan.padded // step))):
            p0, p1 = r * step, min((r + 1) * step, plan.padded)
            kf, kl = max(0, r - span + 1), min(r, n - 1)
            sl = slice(p0 - self.part_p0, p1 - self.part_p0)
            acc = self.partial[:, sl].clone() if kf < k0 else torch.zeros(self.rows, p1 - p0)
            for k in range(max(kf, k0), min(kl, k0 + nb - 1) + 1):
                s, ln = plan.starts[k], plan.lens[k]
                a, b = max(s, p0), min(s + ln, p1)
                if b > a:
                    acc[:, a - p0:b - p0] += y[k - k0][:, a - s:b - s] * self._w(k)[a - s:b - s]
            if kl >= k0 + nb:
                self.partial[:, sl] = acc
                continue
            cnt = torch.zeros(p1 - p0)
            for k in range(kf, kl + 1):
                s, ln = plan.starts[k], plan.lens[k]
                a, b = max(s, p0), min(s + ln, p1)
                if b > a:
                    cnt[a - p0:b - p0] += self._w(k)[a - s:b - s]
            res = torch.nan_to_num(acc / cnt, nan=0.0)
            a, b = max(p0, crop), min(p1, crop + plan.length)
            if b > a:
                self.out[:, a - crop - self.out_q0:b - crop - self.out_q0] = res[:, a - p0:b - p0]


def _run_rank(plan, world, rank, y_all, nrows, engine_batch, group):
    from sesa_audio_separation_b200.distributed import cropped_range, gather_owned, run_sharded_track, shard_layout
    from sesa_audio_separation_b200.plan import windowing_array
    L = plan.chunk_size
    lo, hi, begin, end = shard_layout(plan, world)[rank]
    span = -(-L // plan.step)
    q0, q1 = cropped_range(plan, begin, end) if hi > lo else (0, 0)
    full = torch.full((nrows, plan.length), float('nan')) if rank == 0 else None
    out, out_q0 = (full, 0) if rank == 0 else (torch.full((nrows, max(q1 - q0, 1)), float('nan')), q0)
    req = None
    if hi > lo:
        part_p0 = plan.starts[lo]
        part_p1 = min(plan.padded, (hi + span - 1) * plan.step)
        ops = CpuOps(plan, y_all, nrows, windowing_array(L, plan.fade), part_p0, part_p1 - part_p0, out, out_q0)
        req, sent = run_sharded_track(plan, world, rank, ops, engine_batch, group=group)
        # every chunk of the range went through the model exactly once
        done = sorted(k for k0, nb in ops.forwards for k in range(k0, k0 + nb))
        assert done == list(range(lo, hi)), (rank, ops.forwards)
    res = gather_owned(plan, world, rank, out, out_q0, lambda: full, group=group)
    if req is not None:
        req.wait()
    return res


def _worker(rank, world, port, length, L, ov, bs, eb, out_path, sub):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sesa_audio_separation_b200.plan import make_plan
    plan = make_plan(length, L, ov, bs)
    g = torch.Generator().manual_seed(7)
    nrows = 4
    y_all = torch.randn(plan.n_chunks, nrows, L, generator=g)          # every rank derives the same "model outputs"
    if sub:
        # a sub-group that is NOT ranks 0..W-1 of the world: group rank r is global rank sub[r]
        group = dist.new_group(ranks=sub)
        if rank in sub:
            grank = dist.get_group_rank(group, rank)
            res = _run_rank(plan, len(sub), grank, y_all, nrows, eb, group)
            if grank == 0:
                np.save(out_path, res.numpy())
            else:
                assert res is None
    else:
        res = _run_rank(plan, world, rank, y_all, nrows, eb, None)
        if rank == 0:
            np.save(out_path, res.numpy())
        else:
            assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _reference(length, L, ov, bs):
    """The single-process statement: the oracle's demix bookkeeping over the same per-chunk 'model outputs'."""
    from sesa_audio_separation_b200.plan import make_plan
    plan = make_plan(length, L, ov, bs)
    g = torch.Generator().manual_seed(7)
    y_all = torch.randn(plan.n_chunks, 4, L, generator=g)
    k = [0]

    def model(a):
        o = y_all[k[0]:k[0] + a.shape[0]].reshape(a.shape[0], 2, 2, L)
        k[0] += a.shape[0]
        return o
    mix = np.zeros((2, length), dtype=np.float32)
    return np.asarray(odemix.demix(mix, model, L, ov, bs, 2)).reshape(4, length)


@pytest.mark.parametrize('world,length,L,ov,bs,eb', [(3, 1000, 1000, 4, 1, 4), (3, 2500, 1000, 4, 1, 4), (2, 23456, 1000, 4, 2, 4), (3, 30011, 1000, 4, 1, 3), (2, 9000, 1000, 2, 3, 4),
                                                      (3, 40000, 1000, 8, 4, 4), (2, 4100, 1000, 1, 1, 2), (3, 26000, 1001, 3, 2, 2)])
def test_sharded_overlap_add_equals_single_process(tmp_path, world, length, L, ov, bs, eb):
    out_path = str(tmp_path / 'res.npy')
    mp.spawn(_worker, args=(world, _free_port(), length, L, ov, bs, eb, out_path, None), nprocs=world, join=True)
    got = np.load(out_path)
    ref = _reference(length, L, ov, bs)
    assert got.shape == ref.shape == (4, length)
    # continuing from the sender's partial sums keeps the addition order => bit-identical
    assert np.array_equal(got, ref)


def test_sharding_over_a_sub_group(tmp_path):
    """Group ranks are translated to global ranks for every point-to-point call: a group made of global ranks [1, 2] of
    a 3-process world (so group rank 0 = global rank 1, and global rank 0 takes no part) must give the same bits."""
    out_path = str(tmp_path / 'res.npy')
    mp.spawn(_worker, args=(3, _free_port(), 23456, 1000, 4, 2, 4, out_path, [1, 2]), nprocs=3, join=True)
    assert np.array_equal(np.load(out_path), _reference(23456, 1000, 4, 2))


def test_streamed_accumulate_equals_one_shot_for_any_batching():
    """Single process: folding engine batches of any size one after the other (what DemixEngine does on one GPU) gives
    the oracle's bits."""
    from sesa_audio_separation_b200.plan import make_plan, windowing_array
    for length, L, ov, bs in [(23456, 1000, 4, 2), (5003, 1000, 4, 1), (300, 1000, 4, 2), (9999, 1000, 8, 4), (2600, 1001, 3, 2)]:
        plan = make_plan(length, L, ov, bs)
        g = torch.Generator().manual_seed(7)
        y_all = torch.randn(plan.n_chunks, 4, L, generator=g)
        ref = _reference(length, L, ov, bs)
        span = -(-L // plan.step)
        for eb in (1, 3, 5):
            out = torch.full((4, length), float('nan'))
            ops = CpuOps(plan, y_all, 4, windowing_array(L, plan.fade), 0, plan.padded, out, 0)
            for k in range(0, plan.n_chunks, eb):
                nb = min(eb, plan.n_chunks - k)
                ops.accumulate(y_all[k:k + nb], k, nb, k, k + nb + span - 1)
            assert np.array_equal(out.numpy(), ref), (length, L, ov, bs, eb)


def test_shard_layout_rejects_too_many_ranks():
    from sesa_audio_separation_b200.distributed import shard_layout
    from sesa_audio_separation_b200.plan import make_plan
    plan = make_plan(6000, 1000, 4, 1)
    with pytest.raises(ValueError):
        shard_layout(plan, 16, strict=True)
    lay = shard_layout(plan, 2)
    assert lay[0][2] == 0 and lay[0][3] == lay[1][2] and lay[1][3] == plan.padded
    # a track too short for the ranks offered is sharded over fewer of them; the others stay idle
    lay = shard_layout(plan, 16)
    active = [r for r in range(16) if lay[r][1] > lay[r][0]]
    assert active == list(range(len(active))) and 1 <= len(active) < 16
    assert all(lay[r][1] - lay[r][0] >= 3 for r in active[:-1])
    assert lay[active[-1]][3] == plan.padded and sum(l[1] - l[0] for l in lay) == plan.n_chunks
