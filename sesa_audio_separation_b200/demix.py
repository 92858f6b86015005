"""Chunked separation inference: the B200-native ``demix`` (reference: utils.py:330-477 and its
copy inference_pytorch.py:55-186).

What changes relative to the reference loop: the mix is uploaded ONCE; border padding, chunk framing, the model
forward, the windowed overlap-add (streamed: every engine batch is folded into the running sums as soon as its
forward is done), the divide and the crop all run on the device; finished regions of the result leave for page-locked
host memory while later batches are still computing (the reference does an H2D and a D2H + sync per chunk and
accumulates on the CPU).  What does not change: the chunk schedule, pad modes, per-flush window rule and the
ascending-order accumulation, which are reproduced bit-exactly (plan.py, sesa_overlap_accumulate).  One track can also be
sharded over several GPUs in contiguous chunk ranges (distributed.py), and the test-time-augmentation variants of a mix
run as extra chunks of the same engine run.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call
from .config import prefer_target_instrument
from .module import KernelModule
from .plan import make_plan, windowing_array


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _unwrap(model):
    """Accept the engine model or a PyTorchBackend-style wrapper holding one."""
    inner = getattr(model, 'model', None)
    if isinstance(inner, KernelModule):
        return inner
    if isinstance(model, KernelModule):
        return model
    raise TypeError('demix() drives the sm_100a engine: pass a model built by '
                    'sesa_audio_separation_b200.get_model_from_config (got %r). There is no PyTorch fallback.'
                    % type(model).__name__)


def _round4(n):
    return (int(n) + 3) // 4 * 4


class _Schedule:
    """Device copies of one DemixPlan (chunk starts / lengths / pad modes / window kinds, fade window)."""

    def __init__(self, plan, dev):
        self.plan = plan
        self.starts = torch.tensor(plan.starts, dtype=torch.int64).to(dev)
        self.lens = torch.tensor(plan.lens, dtype=torch.int64).to(dev)
        self.modes = torch.tensor(plan.modes, dtype=torch.int32).to(dev)
        self.kinds = torch.tensor(plan.kinds, dtype=torch.int32).to(dev)
        self.window = windowing_array(plan.chunk_size, plan.fade).to(dev)
        self.rel = {}

    def rel_starts(self, p0, dev):
        """Chunk starts relative to a slab of the padded mix that begins at padded position p0."""
        if p0 not in self.rel:
            self.rel[p0] = torch.tensor([s - p0 for s in self.plan.starts], dtype=torch.int64).to(dev)
        return self.rel[p0]


class DemixEngine:
    """Device-resident demix of one track.  ``engine_batch`` chunks go through the model per launch
    group; it is a throughput knob only and never changes the result's bookkeeping.

    The windowed overlap-add is STREAMED: every engine batch is folded into the running sums as soon as its forward
    is done (sesa_overlap_accumulate), so device memory is the result plus one batch, whatever the track length (the
    reference accumulates on the host and has no bound either).  With ``world > 1`` the track's chunks are sharded in
    contiguous ranges over the ranks of ``group`` (distributed.py): each rank uploads only the slice of the mix its
    chunks touch, and rank ``gather_root`` returns the result (the other ranks return None)."""

    def __init__(self, config, model, device, engine_batch=None, world=1, rank=0, progress=None, group=None):
        _lib.require_cuda()
        self.model = _unwrap(model)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.SesaError(f'demix needs a CUDA device, got {device!r}; there is no CPU path')
        self.model.to(self.device)
        self.chunk_size = int(config.audio.chunk_size)
        self.num_overlap = int(config.inference.num_overlap)
        self.batch_size = int(config.inference.batch_size)
        self.instruments = list(prefer_target_instrument(config))
        # chunks per launch group: a throughput knob only (results are batch-invariant); 4 fills the 148 SMs' tile
        # waves of every GEMM at the BASELINE model sizes and keeps the workspace under 40 GB
        self.engine_batch = int(engine_batch or 4)
        self.world, self.rank, self.group = world, rank, group
        self.progress = progress
        self._sched = {}
        self.stats = {}

    # ------------------------------------------------------------------ helpers
    def _schedule(self, length):
        key = int(length)
        if key not in self._sched:
            if len(self._sched) > 8:
                self._sched.clear()
            self._sched[key] = _Schedule(make_plan(key, self.chunk_size, self.num_overlap, self.batch_size), self.device)
        return self._sched[key]

    def _forward(self, chunks, nb, out=None):
        n_inst, L = len(self.instruments), self.chunk_size
        y = self.model.forward(chunks[:nb]) if out is None else self.model.forward(chunks[:nb], out=out)
        y = y.reshape(nb, n_inst, chunks.shape[1], -1)
        if y.shape[-1] != L:
            raise RuntimeError(f'model returned {y.shape[-1]} samples for a {L}-sample chunk '
                               '(chunk_size must be a multiple of the hop for this model)')
        return y

    def _frame(self, sch, slab, slab_p0, k0, nb, chunks, slot=0):
        """chunks[slot : slot+nb] = framed chunks k0.. of the padded slab that starts at padded position slab_p0."""
        C, L = slab.shape[0], self.chunk_size
        starts = sch.starts if slab_p0 == 0 else sch.rel_starts(slab_p0, self.device)
        call('sesa_frame_chunks', _ptr(slab), slab.shape[1], C, ctypes.c_void_p(starts.data_ptr() + 8 * k0),
             ctypes.c_void_p(sch.lens.data_ptr() + 8 * k0), ctypes.c_void_p(sch.modes.data_ptr() + 4 * k0),
             nb, L, ctypes.c_void_p(chunks.data_ptr() + 4 * slot * C * L), _stream())

    def _accumulate(self, sch, y, k0, nb, r0, r1, partial, part_p0, out, out_q0, out_cols):
        plan = sch.plan
        n_inst = len(self.instruments)
        C = y.shape[2]
        crop = plan.border if plan.pad else 0
        call('sesa_overlap_accumulate', _ptr(y), k0, nb, _ptr(sch.starts), _ptr(sch.lens), _ptr(sch.kinds), plan.n_chunks,
             plan.step, self.chunk_size, plan.fade, _ptr(sch.window), n_inst, C, plan.padded, r0, r1, _ptr(partial),
             partial.shape[1], part_p0, crop, plan.length, _ptr(out), out.stride(0), out_q0, out_cols, _stream())

    def _download(self, t):
        """Device tensor -> numpy through page-locked memory (torch's caching host allocator recycles the block once the
        caller drops the previous result, so steady-state runs pay no pinning cost) in ONE copy."""
        t = t.contiguous()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    def _report(self, plan, k_done, state):
        if self.progress is None:
            return
        pct = int(min(1.0, (plan.starts[k_done - 1] + plan.step) / plan.padded) * 100)
        if pct > state[0]:
            state[0] = pct
            self.progress(pct)

    # ------------------------------------------------------------------ entry
    def run(self, mix, return_counter=False, to_host=True, tta=False, variants=None):
        """mix: (channels, time) host array or CUDA tensor.  Returns the (instruments, channels, time) estimates (numpy when
        ``to_host``).  ``tta=True`` additionally separates the channel-swapped and the polarity-inverted mix IN THE SAME
        engine run (their chunks share launch groups with the original's; one upload, one download) and returns the
        utils.apply_tta average.  ``variants='tta_pair'`` returns only the two augmented estimates [2, n, C, len]."""
        with torch.cuda.device(self.device):
            if self.world > 1:
                if tta or variants:
                    raise NotImplementedError('test-time augmentation runs on one GPU per track (shard tracks instead)')
                return self._run_sharded(mix, to_host)
            return self._run(mix, return_counter, to_host, tta, variants)

    def _upload(self, mix, cols=None):
        dev = self.device
        if isinstance(mix, torch.Tensor) and mix.device.type == 'cuda':
            m = mix.to(device=dev, dtype=torch.float32)
            if m.ndim != 2:
                raise ValueError('mix must have shape (channels, time)')
            return (m if cols is None else m[:, cols[0]:cols[1]]).contiguous()
        mix_h = torch.as_tensor(np.asarray(mix), dtype=torch.float32)
        if mix_h.ndim != 2:
            raise ValueError('mix must have shape (channels, time)')
        if cols is not None:
            mix_h = mix_h[:, cols[0]:cols[1]]
        return mix_h.contiguous().to(dev, non_blocking=True)

    def _run(self, mix, return_counter, to_host, tta, variants):
        dev = self.device
        mix_d = self._upload(mix)
        C, length = mix_d.shape
        L, EB = self.chunk_size, self.engine_batch
        sch = self._schedule(length)
        plan = sch.plan
        self.plan = plan
        st = _stream()
        n_inst = len(self.instruments)
        rows = n_inst * C
        if tta or variants == 'tta_pair':
            sw, ng = torch.empty_like(mix_d), torch.empty_like(mix_d)
            call('sesa_tta_variants', _ptr(mix_d), _ptr(sw), _ptr(ng), C, length, st)
            sources = [sw, ng] if variants == 'tta_pair' else [mix_d, sw, ng]
        else:
            sources = [mix_d]
        V = len(sources)
        slabs = []
        for src in sources:
            if plan.pad:
                padded = torch.empty(C, plan.padded, device=dev, dtype=torch.float32)
                call('sesa_pad_reflect', _ptr(src), _ptr(padded), C, length, plan.border, plan.border, st)
                slabs.append(padded)
            else:
                slabs.append(src)
        ld = _round4(length)
        results = torch.empty(V, rows, ld, device=dev, dtype=torch.float32)
        partial = torch.empty(V, rows, _round4(plan.padded) + 4, device=dev, dtype=torch.float32)
        chunks = torch.empty(EB, C, L, device=dev, dtype=torch.float32)
        span = -(-L // plan.step)
        tasks = [(v, k) for v in range(V) for k in range(plan.n_chunks)]
        state = [-1]
        t = 0
        # to_host: finished regions of the result leave for (page-locked) host memory on a side stream while later
        # batches are still being computed, so the device->host read costs no time after the last batch
        sink = None
        if to_host and V == 1:
            sink = dict(host=torch.empty(rows, length, dtype=torch.float32, pin_memory=True), done=0,
                        stream=self._side_stream(), crop=plan.border if plan.pad else 0)
        while t < len(tasks):
            batch = tasks[t:t + EB]
            nb = len(batch)
            segs = []                                  # runs of consecutive chunks of one variant: (v, k0, n, slot)
            for slot, (v, k) in enumerate(batch):
                if segs and segs[-1][0] == v:
                    segs[-1][2] += 1
                else:
                    segs.append([v, k, 1, slot])
            for v, k0, n, slot in segs:
                self._frame(sch, slabs[v], 0, k0, n, chunks, slot)
            y = self._forward(chunks, nb)
            for v, k0, n, slot in segs:
                self._accumulate(sch, y[slot:slot + n], k0, n, k0, k0 + n + span - 1, partial[v], 0, results[v], 0, length)
            t += nb
            if sink is not None:
                self._drain(sink, results[0], length if t >= len(tasks) else min(max(t * plan.step - sink['crop'], 0), length))
            self._report(plan, (t - 1) % plan.n_chunks + 1 if V == 1 else max(1, t * plan.n_chunks // len(tasks)), state)
        results = results[:, :, :length]
        if tta:
            out = torch.empty(n_inst, C, length, device=dev, dtype=torch.float32)
            r = results.contiguous() if ld != length else results
            call('sesa_tta_combine', _ptr(r[0]), _ptr(r[1]), _ptr(r[2]), _ptr(out), n_inst, C, length, st)
            result = out
        elif variants == 'tta_pair':
            result = results.reshape(V, n_inst, C, length)
        else:
            result = results[0].reshape(n_inst, C, length)
        counter = None
        if return_counter:
            counter = torch.empty(plan.padded, device=dev, dtype=torch.float32)
            crop = plan.border if plan.pad else 0
            call('sesa_overlap_add', None, _ptr(sch.starts), _ptr(sch.lens), _ptr(sch.kinds), plan.n_chunks, plan.step,
                 L, plan.fade, _ptr(sch.window), n_inst, C, plan.padded, crop, 0, None, _ptr(counter), st)
        if not to_host:
            return (result, counter) if return_counter else result
        if sink is not None:
            sink['stream'].synchronize()
            est = sink['host'].numpy().reshape(n_inst, C, length)
        else:
            est = self._download(result)
        return (est, counter.cpu().numpy()) if return_counter else est

    def _side_stream(self):
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def _drain(self, sink, result_rows, upto):
        """Queue the device->host copy of result samples [sink.done, upto) of every row behind the work issued so far."""
        a, b = sink['done'], int(upto)
        if b <= a:
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(sink['stream']):
            sink['stream'].wait_event(ev)
            for row in range(result_rows.shape[0]):
                sink['host'][row, a:b].copy_(result_rows[row, a:b], non_blocking=True)
        sink['done'] = b

    # ------------------------------------------------------------------ chunk-range sharded
    def _run_sharded(self, mix, to_host):
        from .distributed import cropped_range, gather_owned, run_sharded_track, shard_layout
        dev = self.device
        if isinstance(mix, torch.Tensor):
            C, length = mix.shape
        else:
            mix = np.asarray(mix)
            C, length = mix.shape
        L, EB = self.chunk_size, self.engine_batch
        sch = self._schedule(length)
        plan = sch.plan
        self.plan = plan
        n_inst = len(self.instruments)
        rows = n_inst * C
        st = _stream()
        layout = shard_layout(plan, self.world)
        lo, hi, begin, end = layout[self.rank]
        span = -(-L // plan.step)
        crop = plan.border if plan.pad else 0
        root = 0
        stats = self.stats = {'halo_bytes': 0}
        ev = {}

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev[name] = e

        q0, q1 = cropped_range(plan, begin, end) if hi > lo else (0, 0)
        ld = _round4(length)
        full = None
        if self.rank == root:
            full = torch.empty(rows, ld, device=dev, dtype=torch.float32)
            out, out_q0 = full, 0                         # the root's range is finished in place
        else:
            out, out_q0 = torch.empty(rows, _round4(max(q1 - q0, 1)), device=dev, dtype=torch.float32), q0
        req, sent = None, None
        if hi > lo:
            # the slice of the padded mix this rank's chunks read, and the mix samples behind it
            s0 = plan.starts[lo]
            s1 = min(plan.padded, plan.starts[hi - 1] + L)
            if plan.pad:
                idx = [abs(j) if j < length else 2 * (length - 1) - j
                       for j in (s0 - plan.border, s1 - 1 - plan.border)]
                m0, m1 = min(idx), max(idx)
                if s0 - plan.border < 0:
                    m0 = 0
                if s1 - 1 - plan.border >= length:
                    m1 = length - 1
                win = self._upload(mix, (m0, m1 + 1))
                slab = torch.empty(C, s1 - s0, device=dev, dtype=torch.float32)
                call('sesa_pad_reflect_slice', _ptr(win), win.shape[1], m0, _ptr(slab), C, length, plan.border, s0, s1 - s0, st)
            else:
                slab = self._upload(mix, (s0, s1))
            stats['h2d_bytes'] = int(slab.numel() * 4 if not plan.pad else win.numel() * 4)
            n_regions = -(-plan.padded // plan.step)
            part_p0 = s0
            part_p1 = min(plan.padded, min(hi + span - 1, n_regions) * plan.step)
            partial = torch.empty(rows, _round4(part_p1 - part_p0) + 4, device=dev, dtype=torch.float32)
            chunks = torch.empty(EB, C, L, device=dev, dtype=torch.float32)
            state = [-1]
            engine = self

            class _Ops:
                pass
            ops = _Ops()
            ops.rows = rows
            ops.mark = mark

            def forward(k0, nb, keep):
                if not keep:
                    engine._frame(sch, slab, s0, k0, nb, chunks)
                    y = engine._forward(chunks, nb)
                    engine._report(plan, k0 + nb, state)
                    return y
                kept = torch.empty(nb, n_inst, C, L, device=dev, dtype=torch.float32)
                j = 0
                while j < nb:
                    n = min(EB, nb - j)
                    engine._frame(sch, slab, s0, k0 + j, n, chunks)
                    engine._forward(chunks, n, out=kept[j:j + n])
                    j += n
                return kept

            def accumulate(y, k0, nb, r0, r1):
                engine._accumulate(sch, y, k0, nb, r0, r1, partial, part_p0, out, out_q0, out.shape[1])

            ops.forward, ops.accumulate = forward, accumulate
            ops.read_partial = lambda p0, p1: partial[:, p0 - part_p0:p1 - part_p0].contiguous()
            ops.seed_partial = lambda p0, t: partial[:, p0 - part_p0:p0 - part_p0 + t.shape[1]].copy_(t)
            ops.empty = lambda r, c: torch.empty(r, c, device=dev, dtype=torch.float32)
            mark('begin')
            req, sent = run_sharded_track(plan, self.world, self.rank, ops, EB, group=self.group, stats=stats)
            mark('compute_end')
        else:
            mark('begin')
            mark('compute_end')
        result = gather_owned(plan, self.world, self.rank, out, out_q0, lambda: full, group=self.group, gather_root=root)
        mark('gather_end')
        if req is not None:
            req.wait()
        self._events = ev
        if result is None:
            return None
        result = result[:, :length].reshape(n_inst, C, length)
        return self._download(result) if to_host else result

    def timings(self):
        """Device-side milliseconds of the last sharded run on this rank (call after a synchronize)."""
        ev = getattr(self, '_events', None)
        if not ev:
            return {}
        t = {'compute_ms': ev['begin'].elapsed_time(ev['compute_end']),
             'gather_ms': ev['compute_end'].elapsed_time(ev['gather_end'])}
        if 'halo_wait_begin' in ev:
            t['halo_ms'] = ev['halo_wait_begin'].elapsed_time(ev['halo_wait_end'])
        else:
            t['halo_ms'] = 0.0
        t.update(self.stats)
        return t


def demix(config, model, mix, device, model_type, pbar=False, engine_batch=None):
    """Drop-in for utils.demix (utils.py:330-477), generic mode.  Returns {instrument: ndarray(C, len)}."""
    if model_type == 'htdemucs':
        raise NotImplementedError('htdemucs (demucs mode of demix) is out of scope of the B200 hot path')
    eng = DemixEngine(config, model, device, engine_batch=engine_batch)
    est = eng.run(mix)
    return {k: v for k, v in zip(eng.instruments, est)}


def demix_pytorch_optimized(config, backend, mix, device, pbar=False, engine_batch=None):
    """Drop-in for inference_pytorch.demix_pytorch_optimized (:55-186): same result as demix() and the
    ``[SESA_PROGRESS]<int>`` stdout protocol the GUI parses (processing.py:345-359)."""
    eng = DemixEngine(config, backend, device, engine_batch=engine_batch,
                      progress=lambda p: print(f"[SESA_PROGRESS]{p}", flush=True))
    est = eng.run(mix)
    print("[SESA_PROGRESS]100", flush=True)
    return {k: v for k, v in zip(eng.instruments, est)}


def normalize_audio(audio):
    """utils.py:199-217."""
    mono = audio.mean(0)
    mean, std = mono.mean(), mono.std()
    return (audio - mean) / std, {"mean": mean, "std": std}


def denormalize_audio(audio, norm_params):
    """utils.py:220-238."""
    return audio * norm_params["std"] + norm_params["mean"]


def apply_tta(config, model, mix, waveforms_orig, device, model_type):
    """Drop-in for utils.apply_tta (utils.py:241-292): averages ``waveforms_orig`` (updated in place, like the reference)
    with the estimates of the channel-swapped and the polarity-inverted mix.  Both augmented mixes are built on the
    device from ONE upload and separated in ONE engine run (their chunks share launch groups), and come back in one
    download; the combination keeps the reference's order of operations (+= swapped[::-1]; -= inverted; /= 3)."""
    if model_type == 'htdemucs':
        raise NotImplementedError('htdemucs (demucs mode of demix) is out of scope of the B200 hot path')
    eng = DemixEngine(config, model, device)
    pair = eng.run(mix, variants='tta_pair')                  # [2, instruments, C, len]
    for i, name in enumerate(eng.instruments):
        waveforms_orig[name] += pair[0][i][::-1].copy()
        waveforms_orig[name] -= pair[1][i]
    for name in waveforms_orig:
        waveforms_orig[name] /= 3
    return waveforms_orig


def demix_tta(config, model, mix, device, model_type, engine_batch=None, progress=None):
    """demix() followed by apply_tta() as ONE engine run: the original and the two augmented mixes are separated
    together and combined on the device (sesa_tta_combine); bit-identical to the two-call sequence."""
    if model_type == 'htdemucs':
        raise NotImplementedError('htdemucs (demucs mode of demix) is out of scope of the B200 hot path')
    eng = DemixEngine(config, model, device, engine_batch=engine_batch, progress=progress)
    est = eng.run(mix, tta=True)
    return {k: v for k, v in zip(eng.instruments, est)}
