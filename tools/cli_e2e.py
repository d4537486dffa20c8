#!/usr/bin/env python
"""Text in -> text out: the `filtlong` command line of this repo (GPU) beside the unmodified reference
binary (oracle/_ref/filtlong_ref, CPU) on the same synthetic FASTQ, same arguments, on this box.
Checks that stdout is byte-identical and reports wall-clock seconds and Gbases/s. Test / measurement
infrastructure (it executes oracle/_ref); writes one JSON line.

    python tools/cli_e2e.py [--phred-reads 100000] [--kmer-reads 20000] > gpurun_out/cli_e2e.json
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload generators)

OURS = os.path.join(ROOT, "filtlong_b200", "bin", "filtlong")
REF = os.path.join(ROOT, "oracle", "_ref", "filtlong_ref")


def run(cmd, out_path, timing=False):
    env = dict(os.environ, LC_ALL="C", LANG="C")
    if timing:
        env["FL_CLI_TIMING"] = "1"
    t = time.time()
    with open(out_path, "wb") as f:
        r = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, env=env)
    dt = time.time() - t
    h = hashlib.md5()
    with open(out_path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    err = r.stderr.decode(errors="replace").strip().splitlines()
    return dict(seconds=dt, rc=r.returncode, md5=h.hexdigest(), bytes=os.path.getsize(out_path),
                stderr_tail=[x for x in err if not x.startswith("[timing]")][-3:],
                phases=[x[9:].strip() for x in err if x.startswith("[timing]")])


def case(name, fq, bases, extra, td):
    res = {"case": name, "bases": bases, "args": extra}
    for tag, exe in (("reference_cpu", REF), ("ours_gpu", OURS)):
        if not os.path.exists(exe):
            res[tag] = {"error": "missing " + exe}
            continue
        # second run of ours: CUDA context creation and module load are paid once per process
        r = run([exe] + extra + [fq], os.path.join(td, name + "." + tag + ".out"), timing=(tag == "ours_gpu"))
        r["gbases_per_s"] = bases / r["seconds"] / 1e9
        res[tag] = r
    a, b = res.get("reference_cpu", {}), res.get("ours_gpu", {})
    res["stdout_identical"] = bool(a.get("md5") and a.get("md5") == b.get("md5"))
    if a.get("seconds") and b.get("seconds"):
        res["speedup_wall"] = a["seconds"] / b["seconds"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phred-reads", type=int, default=100000)
    ap.add_argument("--kmer-reads", type=int, default=20000)
    a = ap.parse_args()
    out = {"what": "filtlong CLI, FASTQ text in -> FASTQ text out, wall clock of the whole process", "cases": []}
    with tempfile.TemporaryDirectory(prefix="flcli_") as td:
        w = bench.phred_workload(0, 2000000, 20 * 10 ** 9)
        fq = os.path.join(td, "phred.fastq")
        bases = bench.write_sample_fastq(fq, w, a.phred_reads)
        out["cases"].append(case("phred_target_25pct", fq, bases, ["--target_bases", str(bases // 4)], td))
        wk = bench.kmer_workload(0, 2000000, 20 * 10 ** 9, 10 ** 7)
        fa = os.path.join(td, "genome.fasta")
        g = bench.genome_fasta(fa, wk)
        fqk = os.path.join(td, "kmer.fastq")
        kb = bench.write_sample_fastq(fqk, wk, a.kmer_reads, genome=g)
        out["cases"].append(case("kmer_assembly_trim_split", fqk, kb, ["-a", fa, "--trim", "--split", "500", "-p", "90"], td))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
