"""CPU, world_size 2 and 3 over gloo: the chunk-range sharding choreography of distributed.py (halo send/recv,
owned ranges, gather on the root) against the single-process oracle overlap-add.  The torch-CPU ``ops`` below is
test infrastructure standing in for the CUDA closures over sesa_overlap_add_range (demix.py)."""
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
    """raw / final of one rank, written with the same ascending-chunk accumulation as the kernel."""

    def __init__(self, plan, y, lo, hi, window, nrows):
        self.plan, self.y, self.lo, self.hi, self.window, self.nrows = plan, y, lo, hi, window, nrows

    def _w(self, k):
        from sesa_audio_separation_b200.plan import KIND_NO_FADEIN, KIND_NO_FADEOUT
        w = self.window.clone()
        fade = self.plan.fade
        if self.plan.kinds[k] == KIND_NO_FADEIN:
            w[:fade] = 1
        elif self.plan.kinds[k] == KIND_NO_FADEOUT:
            w[-fade:] = 1
        return w

    def _sum(self, p0, p1, init, init_p0):
        acc = torch.zeros(self.nrows, p1 - p0)
        if init is not None:
            acc[:, init_p0 - p0:init_p0 - p0 + init.shape[1]] = init
        for k in range(self.lo, self.hi):
            s, n = self.plan.starts[k], self.plan.lens[k]
            a, b = max(s, p0), min(s + n, p1)
            if b > a:
                acc[:, a - p0:b - p0] += self.y[k - self.lo][:, a - s:b - s] * self._w(k)[a - s:b - s]
        return acc

    def raw(self, p0, p1):
        return self._sum(p0, p1, None, 0)

    def final(self, p0, p1, init, init_p0):
        plan = self.plan
        acc = self._sum(p0, p1, init, init_p0)
        cnt = torch.zeros(p1 - p0)
        for k in range(plan.n_chunks):
            s, n = plan.starts[k], plan.lens[k]
            a, b = max(s, p0), min(s + n, p1)
            if b > a:
                save = (self.lo, self.hi)
                cnt[a - p0:b - p0] += self._w(k)[a - s:b - s]
        crop = plan.border if plan.pad else 0
        q0 = max(p0, crop)
        q1 = max(min(p1, crop + plan.length), q0)
        res = torch.nan_to_num(acc / cnt, nan=0.0)
        return res[:, q0 - p0:q1 - p0]


def _worker(rank, world, port, length, L, ov, bs, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sesa_audio_separation_b200.distributed import sharded_overlap_add
    from sesa_audio_separation_b200.plan import make_plan, shard_chunks, windowing_array
    plan = make_plan(length, L, ov, bs)
    g = torch.Generator().manual_seed(7)
    nrows = 4
    y_all = torch.randn(plan.n_chunks, nrows, L, generator=g)          # every rank derives the same "model outputs"
    lo, hi = shard_chunks(plan.n_chunks, world, rank)
    ops = CpuOps(plan, y_all[lo:hi], lo, hi, windowing_array(L, plan.fade), nrows)
    res = sharded_overlap_add(plan, world, rank, nrows, ops, torch.device('cpu'))
    if rank == 0:
        np.save(out_path, res.numpy())
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,length,L,ov,bs', [(2, 23456, 1000, 4, 2), (3, 30011, 1000, 4, 1), (2, 9000, 1000, 2, 3),
                                                   (3, 40000, 1000, 8, 4), (2, 4100, 1000, 1, 1)])
def test_sharded_overlap_add_equals_single_process(tmp_path, world, length, L, ov, bs):
    from sesa_audio_separation_b200.plan import make_plan, windowing_array
    out_path = str(tmp_path / 'res.npy')
    port = _free_port()
    mp.spawn(_worker, args=(world, port, length, L, ov, bs, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    plan = make_plan(length, L, ov, bs)
    g = torch.Generator().manual_seed(7)
    y_all = torch.randn(plan.n_chunks, 4, L, generator=g)
    ops = CpuOps(plan, y_all, 0, plan.n_chunks, windowing_array(L, plan.fade), 4)
    ref = ops.final(0, plan.padded, None, 0).numpy()
    assert got.shape == ref.shape == (4, length)
    # seeding the receiver with the sender's partial sums keeps the addition order => bit-identical
    assert np.array_equal(got, ref)
    # and the single-process statement equals the oracle's demix bookkeeping
    k = [0]

    def model(a):
        o = y_all[k[0]:k[0] + a.shape[0]].reshape(a.shape[0], 2, 2, L)
        k[0] += a.shape[0]
        return o
    mix = np.zeros((2, length), dtype=np.float32)
    oref = odemix.demix(mix, model, L, ov, bs, 2)
    assert np.array_equal(np.asarray(oref).reshape(4, length), ref)


def test_shard_layout_rejects_too_many_ranks():
    from sesa_audio_separation_b200.distributed import shard_layout
    from sesa_audio_separation_b200.plan import make_plan
    plan = make_plan(6000, 1000, 4, 1)
    with pytest.raises(ValueError):
        shard_layout(plan, 16)
    lay = shard_layout(plan, 2)
    assert lay[0][2] == 0 and lay[0][3] == lay[1][2] and lay[1][3] == plan.padded
