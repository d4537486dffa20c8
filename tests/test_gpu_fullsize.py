"""GPU tests at BASELINE.json's FULL sizes (config 2: 20 Gbases Phred, config 3: 20 Gbases vs a
10 Mbp assembly), where the oracle cannot follow. Parity is carried by size-independent properties:

* two independent exact implementations of the same scores must agree bit-for-bit on every one of
  the 2 M reads (Phred: lattice warp kernels vs one-thread-per-chain work items; k-mer: the position-anchored
  table with and without the L2 pre-filter, and the plain bitmap);
* results must not depend on how the read set is cut into batches;
* the first few thousand reads, regenerated on the host with the same counter-based generator,
  must match the oracle bit-for-bit;
* the selection must satisfy the reference's prefix-walk invariants (main.cpp:251-257): every kept
  row scores at least as high as every passed row that was dropped, kept bases reach the target and
  would not reach it without the lowest-scoring kept row.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

SAMPLE = 3000


def _env(**kw):
    class E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return E()


def _selection_invariants(rows, summ, target):
    passed = rows["passed"].astype(bool)
    kept = rows["passed_final"].astype(bool)
    length = (rows["end"] - rows["start"]).astype(np.int64)
    assert not np.any(kept & ~passed)                      # a failed read never comes back (main.cpp:251)
    assert summ.status == 3
    kept_bases = int(length[kept].sum())
    assert kept_bases == summ.keeping
    assert kept_bases >= target                            # the crossing read is kept (main.cpp:252-254)
    fs = rows["final_score"]
    lo_kept = fs[kept].min()
    dropped = passed & ~kept
    assert fs[dropped].max() <= lo_kept                    # descending-score prefix
    # without the rows at the lowest kept score the target is not reached (bases_so_far < target there)
    assert kept_bases - int(length[kept & (fs == lo_kept)].sum()) < target


def _phred_setup(torch, bench, api, capi, dev):
    w = bench.phred_workload(0, 2000000, 20 * 10 ** 9)
    t_len = torch.from_numpy(w["len"]).to(dev)
    t_off = torch.from_numpy(w["off"].view(np.int64)).to(dev)
    t_qbar = torch.from_numpy(w["qbar"]).to(dev)
    d_qual = torch.empty(w["padded"] + 64, dtype=torch.uint8, device=dev)
    return w, t_len, t_off, t_qbar, d_qual


def test_config2_phred_full_size():
    import torch
    import bench
    from filtlong_b200 import api, capi, sharding
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    L = capi.lib()
    w, t_len, t_off, t_qbar, d_qual = _phred_setup(torch, bench, api, capi, dev)
    target = 5 * 10 ** 9
    params = api.make_params(target_bases=target)
    results = {}
    for tag, mode, nbatch in (("lattice", 1, 1), ("items", 0, 1), ("lattice_batched", 1, 5)):
        with _env(FL_PHRED_MODE=mode):
            ctx = api.Context(params)
        if tag == "lattice":
            capi.check(ctx.h, L.fl_synth_qual_device(ctx.h, w["seed"], w["n"], t_off.data_ptr(), t_len.data_ptr(),
                                                     t_qbar.data_ptr(), w["read_base"], d_qual.data_ptr()), "synth_qual")
            ctx.sync()
        keep = []
        for lo, hi in sharding.shard_by_bases(w["len"], nbatch):
            base = int(w["off"][lo])
            end = int(w["off"][hi]) if hi < w["n"] else w["padded"]
            rel = torch.from_numpy((w["off"][lo:hi] - np.uint64(base)).view(np.int64)).to(dev)
            keep.append(rel)
            b = api.device_batch(hi - lo, end - base, rel, t_len[lo:hi], qual=d_qual[base:])
            ctx.push_device(b)
        summ = ctx.finalize(w["bases"])
        results[tag] = (ctx.read_results(), ctx.row_results(), summ)
        ctx.close()
    rr, rows, summ = results["lattice"]
    assert len(rr["mean_q"]) == w["n"]
    for other in ("items", "lattice_batched"):
        r2, rows2, s2 = results[other]
        for k in ("mean_q", "window_q"):
            assert np.array_equal(rr[k].view(np.uint64), r2[k].view(np.uint64)), (other, k)
        assert np.array_equal(rr["passed"], r2["passed"])
        assert np.array_equal(rows["final_score"].view(np.uint64), rows2["final_score"].view(np.uint64)), other
        assert np.array_equal(rows["passed_final"], rows2["passed_final"]), other
        assert (s2.status, s2.target, s2.keeping) == (summ.status, summ.target, summ.keeping)
    _selection_invariants(rows, summ, target)
    # the first reads against the oracle, regenerated on the host
    lens = np.ascontiguousarray(w["len"][:SAMPLE])
    off, padded = bench.layout(lens)
    qbar = np.ascontiguousarray(w["qbar"][:SAMPLE])
    qual = np.zeros(padded + 64, dtype=np.uint8)
    L.fl_synth_qual_host(w["seed"], SAMPLE, capi.ptr(off), capi.ptr(lens), capi.ptr(qbar), w["read_base"], capi.ptr(qual))
    reads = [(b"A" * int(n), qual[int(o):int(o) + int(n)].tobytes()) for o, n in zip(off, lens)]
    sc = orc.score(reads, orc.make_params(target_bases=target), None)
    parity.check_reads_vs_oracle({k: v[:SAMPLE] for k, v in rr.items()}, sc)


def test_config3_kmer_full_size():
    import torch
    import bench
    from filtlong_b200 import api, capi
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    L = capi.lib()
    w = bench.kmer_workload(0, 2000000, 20 * 10 ** 9, 10 ** 7)
    gb = w["genome_bases"]
    t_len = torch.from_numpy(w["len"]).to(dev)
    t_off = torch.from_numpy(w["off"].view(np.int64)).to(dev)
    target = 5 * 10 ** 9
    params = api.make_params(target_bases=target, keep_percent=90.0, trim=True, split=500)
    d_genome = torch.zeros(gb // 16 + 8, dtype=torch.int32, device=dev)
    d_seq = torch.zeros(w["padded"] // 16 + 8, dtype=torch.int32, device=dev)

    def as_torch(x):
        x = x.view(np.int64) if x.dtype == np.uint64 else (x.view(np.int32) if x.dtype == np.uint32 else x)
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    keep = [as_torch(w[k]) for k in ("start", "strand", "err", "junk_pos", "junk_len")]
    results = {}
    for tag, filt, anch in (("filter", 1, 1), ("direct", 0, 1), ("bitmap", 1, 0)):
        with _env(FL_FILTER=filt, FL_ANCHOR=anch):
            ctx = api.Context(params)
        if tag == "filter":
            capi.check(ctx.h, L.fl_synth_genome_device(ctx.h, w["genome_seed"], gb, d_genome.data_ptr()), "synth_genome")
            desc = capi.SynthReads()
            desc.n, desc.genome_bases = w["n"], gb
            desc.off, desc.len = t_off.data_ptr(), t_len.data_ptr()
            desc.start, desc.strand, desc.err_ppm, desc.junk_pos, desc.junk_len = [t.data_ptr() for t in keep]
            capi.check(ctx.h, L.fl_synth_reads_device(ctx.h, w["seed"], d_genome.data_ptr(), C.byref(desc), w["read_base"],
                                                      d_seq.data_ptr()), "synth_reads")
            ctx.sync()
        g_off = torch.zeros(1, dtype=torch.int64, device=dev)
        g_len = torch.tensor([gb], dtype=torch.int32, device=dev)
        ctx.kmers_add_device(api.device_batch(1, (gb + 63) & ~63, g_off, g_len, seq2b=d_genome), False)
        n_kmers = ctx.kmers_count()
        ctx.push_device(api.device_batch(w["n"], w["padded"], t_off, t_len, seq2b=d_seq))
        summ = ctx.finalize(w["bases"])
        results[tag] = (ctx.read_results(), ctx.row_results(), summ, n_kmers)
        ctx.close()
    rr, rows, summ, n_kmers = results["filter"]
    assert 19 * 10 ** 6 < n_kmers <= 2 * (gb - 15)
    for other in ("direct", "bitmap"):
        r2, rows2, s2, n2 = results[other]
        assert n_kmers == n2
        for k in ("mean_q", "window_q"):
            assert np.array_equal(rr[k].view(np.uint64), r2[k].view(np.uint64)), (other, k)
        for k in ("first_base_in_kmer", "last_base_in_kmer", "n_bad", "n_child", "passed"):
            assert np.array_equal(rr[k], r2[k]), (other, k)
        for k in ("start", "end", "passed_final"):
            assert np.array_equal(rows[k], rows2[k]), (other, k)
        assert (s2.status, s2.target, s2.keeping) == (summ.status, summ.target, summ.keeping)
    # children tile their parent without overlap, in coordinate order (read.cpp:119-130)
    assert np.all(rows["end"] >= rows["start"])
    same_parent = rows["parent"][1:] == rows["parent"][:-1]
    assert np.all(rows["start"][1:][same_parent] >= rows["end"][:-1][same_parent])
    assert int((rows["end"] - rows["start"]).astype(np.int64).sum()) == summ.rows_bases
    _selection_invariants(rows, summ, summ.target)
    # the first reads against the oracle (host-regenerated genome and reads, same generator)
    n_s = 400
    g = np.zeros(gb // 16 + 8, dtype=np.uint32)
    L.fl_synth_genome_host(w["genome_seed"], gb, capi.ptr(g))
    lens = np.ascontiguousarray(w["len"][:n_s])
    off, padded = bench.layout(lens)
    d = capi.SynthReads()
    arrs = [np.ascontiguousarray(w[k][:n_s]) for k in ("start", "strand", "err", "junk_pos", "junk_len")]
    d.n, d.genome_bases = n_s, gb
    d.off, d.len = capi.ptr(off), capi.ptr(lens)
    d.start, d.strand, d.err_ppm, d.junk_pos, d.junk_len = [capi.ptr(a) for a in arrs]
    arena = np.zeros(padded // 16 + 8, dtype=np.uint32)
    L.fl_synth_reads_host(w["seed"], capi.ptr(g), C.byref(d), w["read_base"], capi.ptr(arena))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)

    def unpack(words, n):
        return lut[((words[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))) & 3).reshape(-1)[:n]].tobytes()

    genome = unpack(g[:gb // 16 + 1], gb)
    reads = [(unpack(arena[int(o) // 16:(int(o) + int(n) + 15) // 16], int(n)), None) for o, n in zip(off, lens)]
    ok = orc.Kmers()
    ok.add_assembly([genome])
    sc = orc.score(reads, orc.make_params(target_bases=target, keep_percent=90.0, trim=True, split=500), ok)
    parity.check_reads_vs_oracle({k: v[:n_s] for k, v in rr.items()}, sc)
    # ... and their child ranges (rows are in file order: the first reads own the first rows)
    row = 0
    for p, kids in zip(sc.parents, sc.children):
        for r in (kids if kids else [p]):
            assert (rows["start"][row], rows["end"][row]) == (r.start, r.end), row
            assert rows["mean_q"][row] == r.mean_q and rows["window_q"][row] == r.window_q, row
            row += 1
