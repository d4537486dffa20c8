#!/usr/bin/env python
"""Throughput of the reference 16-mer set build on one B200 (SURVEY 8a K2-K5, 8d "K-build"):
  assembly   : every forward / reverse 16-mer of N contigs x 3 Mbp (BASELINE config 5: 1000 contigs = 3 Gbp)
  short reads: the >= 4 sightings rule (3 on a Bloom false positive) in closed form over 2 x 150 bp
               reads at 150x of a 10 Mbp genome (BASELINE config 3: 10 M reads = 1.5 Gbp)
Inputs are generated on the device; prints one JSON line per case. Measurement tool (no oracle).

    python tools/kbuild_bench.py [--contigs 1000] [--short-reads 10000000]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from filtlong_b200 import api, capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=1000)
    ap.add_argument("--contig-bases", type=int, default=3000000)
    ap.add_argument("--short-reads", type=int, default=10000000)
    ap.add_argument("--genome-bases", type=int, default=10000000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    L = capi.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def timed(fn):
        torch.cuda.synchronize(dev)
        t = time.time()
        r = fn()
        torch.cuda.synchronize(dev)
        return r, time.time() - t

    # ---- assembly: N contigs x contig_bases ----
    ctx = api.Context(api.make_params())
    cb = (a.contig_bases + 63) & ~63
    gb = a.contigs * cb
    d_g = torch.zeros(gb // 16 + 8, dtype=torch.int32, device=dev)
    capi.check(ctx.h, L.fl_synth_genome_device(ctx.h, 4, gb, d_g.data_ptr()), "synth_genome")
    off = torch.arange(a.contigs, dtype=torch.int64, device=dev) * cb
    ln = torch.full((a.contigs,), a.contig_bases, dtype=torch.int32, device=dev)
    batch = api.device_batch(a.contigs, gb, off, ln, seq2b=d_g)
    ctx.sync()
    ctx.enable_timing(True)
    _, wall = timed(lambda: ctx.kmers_add_device(batch, False))
    n_kmers, wall_count = timed(ctx.kmers_count)
    k_ms, k_n = ctx.kernel_time("kmers_add")
    bases = a.contigs * a.contig_bases
    adds = 2 * a.contigs * (a.contig_bases - 15)
    alg = bases / 4 + 2 * 32 * (bases - 15 * a.contigs)       # 2-bit stream + one 32-byte sector read-modify-write per add
    print(json.dumps({"case": "assembly (-a)", "contigs": a.contigs, "bases": bases, "adds": adds, "n_kmers": int(n_kmers),
                      "kernel": "k_kmers_add<0>", "kernel_ms": k_ms, "launches": int(k_n), "wall_s": wall, "count_s": wall_count,
                      "Gadds_per_s": adds / (k_ms * 1e-3) / 1e9, "Gbases_per_s": bases / (k_ms * 1e-3) / 1e9,
                      "algorithmic_GBps": alg / (k_ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
                      "frac_of_formula_roofline": alg / (k_ms * 1e-3) / 1e9 / peak}))
    ctx.close()
    del d_g, batch
    torch.cuda.empty_cache()

    # ---- short reads: >= 4 copies ----
    ctx = api.Context(api.make_params())
    g = a.genome_bases
    d_genome = torch.zeros(g // 16 + 8, dtype=torch.int32, device=dev)
    capi.check(ctx.h, L.fl_synth_genome_device(ctx.h, 2, g, d_genome.data_ptr()), "synth_genome")
    n = a.short_reads
    rng = np.random.default_rng(11)
    lens = np.full(n, 150, dtype=np.int32)
    offs = (np.arange(n, dtype=np.uint64) * 192)
    padded = n * 192
    start = (rng.random(n) * (g - 150)).astype(np.uint64)
    strand = (rng.random(n) < 0.5).astype(np.uint8)
    err = np.full(n, 2000, dtype=np.uint32)                    # 0.2 %
    zeros = np.zeros(n, dtype=np.int32)

    def t(x):
        x = x.view(np.int64) if x.dtype == np.uint64 else (x.view(np.int32) if x.dtype == np.uint32 else x)
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    t_off, t_len = t(offs), t(lens)
    keep = [t(start), t(strand), t(err), t(zeros), t(zeros)]
    d_seq = torch.zeros(padded // 16 + 8, dtype=torch.int32, device=dev)
    desc = capi.SynthReads()
    desc.n, desc.genome_bases = n, g
    desc.off, desc.len = t_off.data_ptr(), t_len.data_ptr()
    desc.start, desc.strand, desc.err_ppm, desc.junk_pos, desc.junk_len = [x.data_ptr() for x in keep]
    capi.check(ctx.h, L.fl_synth_reads_device(ctx.h, 5, d_genome.data_ptr(), C.byref(desc), 0, d_seq.data_ptr()), "synth_reads")
    batch = api.device_batch(n, padded, t_off, t_len, seq2b=d_seq)
    ctx.sync()
    ctx.enable_timing(True)
    _, wall = timed(lambda: ctx.kmers_add_device(batch, True))
    n_kmers, wall_final = timed(ctx.kmers_count)                # closed-form promotion (Bloom bit times, count >= 4 / == 3 rule)
    k_ms, k_n = ctx.kernel_time("kmers_add")
    bases = n * 150
    adds = 2 * n * (150 - 15)
    print(json.dumps({"case": "short reads (-1/-2), >= 4 copies", "reads": n, "bases": bases, "adds": adds, "n_kmers": int(n_kmers),
                      "genome_16mers_both_strands": 2 * (g - 15),
                      "kernel": "k_kmers_add<1>", "kernel_ms": k_ms, "launches": int(k_n), "wall_add_s": wall,
                      "finalize_s (k_bloom_times + k_promote + count)": wall_final,
                      "Gadds_per_s": adds / (k_ms * 1e-3) / 1e9, "Gadds_per_s_incl_finalize": adds / (wall + wall_final) / 1e9,
                      "reference_cpu_ns_per_add": 105}))
    ctx.close()


if __name__ == "__main__":
    main()
