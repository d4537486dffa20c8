"""Host-side logic of a read set sharded across GPUs, one process per GPU.

Reads are independent until normalisation, so each rank scores a contiguous shard (file order is
kept: rank r holds reads [lo_r, hi_r)) and only the small reductions of main.cpp:169-261 cross
ranks. `sharded_finalize` drives the split-phase C-ABI calls of include/filtlong_b200.h and puts
an all-reduce between them; the compute backend is pluggable so the protocol itself is tested on
CPU with gloo (tests/test_sharded_select.py) while bench.py runs it over NCCL / NVLink.
"""
import numpy as np


def shard_by_bases(lengths, world):
    """Contiguous read ranges [lo, hi) per rank, balanced by bases (not by read count)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    csum = np.cumsum(lengths)
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cuts.append(int(np.searchsorted(csum, target, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


class Buffers:
    """The small reduction buffers of the split-phase API (torch tensors on the backend's device)."""

    def __init__(self, torch, device, world):
        f64 = torch.zeros(8, dtype=torch.float64, device=device)
        self.sums, self.mn, self.mx, self.sq = f64[0:4], f64[4:5], f64[5:6], f64[6:7]
        self.hist = torch.zeros(256, dtype=torch.int64, device=device)
        self.tie = torch.zeros(max(world, 1), dtype=torch.int64, device=device)
        self.keeping = torch.zeros(1, dtype=torch.int64, device=device)
        self._keep = f64


def sharded_finalize(backend, dist, buf, rank, world, total_bases_global):
    """main.cpp:169-261 over all ranks. `backend` exposes the split-phase calls on tensors;
    `dist` is torch.distributed (or None when world == 1). Returns the backend's summary."""
    def ar(t, op=None):
        if world > 1:
            dist.all_reduce(t) if op is None else dist.all_reduce(t, op=op)

    backend.norm_partial1(buf.sums, buf.mn, buf.mx)
    ar(buf.sums)
    if world > 1:
        ar(buf.mn, dist.ReduceOp.MIN)
        ar(buf.mx, dist.ReduceOp.MAX)
    backend.norm_partial2(buf.sums, buf.mn, buf.mx, buf.sq)
    ar(buf.sq)
    backend.norm_apply(buf.sums, buf.mn, buf.mx, buf.sq)
    backend.select_begin(total_bases_global, buf.sums)
    for level in range(8):
        backend.select_hist(level, buf.hist)
        ar(buf.hist)                    # per-shard histogram of bases per score digit
        backend.select_pick(level, buf.hist)
    backend.select_tie_local(buf.tie, rank, world)
    ar(buf.tie)
    backend.select_apply(buf.tie, rank, buf.keeping)
    ar(buf.keeping)
    return backend.select_summary(buf.sums, buf.mn, buf.mx, buf.sq, buf.keeping, total_bases_global)


class CabiBackend:
    """The split-phase calls of the CUDA library on one context."""

    def __init__(self, ctx):
        import ctypes as C
        from . import capi
        self.ctx, self.L, self.C, self.capi = ctx, capi.lib(), C, capi

    def _ck(self, rc, what):
        self.capi.check(self.ctx.h, rc, what)

    def norm_partial1(self, sums, mn, mx):
        self._ck(self.L.fl_norm_partial1(self.ctx.h, sums.data_ptr(), mn.data_ptr(), mx.data_ptr()), "fl_norm_partial1")

    def norm_partial2(self, sums, mn, mx, sq):
        self._ck(self.L.fl_norm_partial2(self.ctx.h, sums.data_ptr(), mn.data_ptr(), mx.data_ptr(), sq.data_ptr()), "fl_norm_partial2")

    def norm_apply(self, sums, mn, mx, sq):
        self._ck(self.L.fl_norm_apply(self.ctx.h, sums.data_ptr(), mn.data_ptr(), mx.data_ptr(), sq.data_ptr()), "fl_norm_apply")

    def select_begin(self, total, sums):
        self._ck(self.L.fl_select_begin(self.ctx.h, total, sums.data_ptr()), "fl_select_begin")

    def select_hist(self, level, hist):
        self._ck(self.L.fl_select_hist(self.ctx.h, level, hist.data_ptr()), "fl_select_hist")

    def select_pick(self, level, hist):
        self._ck(self.L.fl_select_pick(self.ctx.h, level, hist.data_ptr()), "fl_select_pick")

    def select_tie_local(self, tie, rank, world):
        self._ck(self.L.fl_select_tie_local(self.ctx.h, tie.data_ptr(), rank, world), "fl_select_tie_local")

    def select_apply(self, tie, rank, keeping):
        self._ck(self.L.fl_select_apply(self.ctx.h, tie.data_ptr(), rank, keeping.data_ptr()), "fl_select_apply")

    def select_summary(self, sums, mn, mx, sq, keeping, total):
        s = self.capi.Summary()
        self._ck(self.L.fl_select_summary(self.ctx.h, sums.data_ptr(), mn.data_ptr(), mx.data_ptr(), sq.data_ptr(),
                                          keeping.data_ptr(), total, self.C.byref(s)), "fl_select_summary")
        return s
