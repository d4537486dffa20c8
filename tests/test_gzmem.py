"""The feeder's gzip front end (filtlong_b200/csrc/host/gzmem.cpp; reference: gzread under kseq, src/main.cpp:70-75):
a .gz input is inflated once into memory and then parsed on the device like a plain file. CPU-only: the inflated bytes
must be exactly what zlib's own reader returns -- one member, concatenated members, trailing bytes that are not a gzip
header (gzread ignores them), BGZF (bgzip) blocks inflated in parallel -- and anything truncated, corrupt, not gzip or
over the memory budget must be declined so that the host reader reports it the way the reference does."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "filtlong_b200", "csrc", "host")


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gz") / "gzmem_dump")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-I", HOST, os.path.join(ROOT, "tests", "gzmem_dump.cpp"),
                        os.path.join(HOST, "gzmem.cpp"), "-lz", "-lpthread", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def fastq_text(rng, n_bytes):
    out = []
    size = 0
    i = 0
    while size < n_bytes:
        L = int(rng.integers(50, 3000))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L))
        qual = bytes(rng.integers(35, 75, size=L).astype(np.uint8))
        rec = b"@read%d some comment\n" % i + seq + b"\n+\n" + qual + b"\n"
        out.append(rec)
        size += len(rec)
        i += 1
    return b"".join(out)


def bgzf(data, block=0xff00, level=6):
    """bgzip's container (SAM specification 4.1): independent gzip members with a 'BC' extra field + the empty EOF block"""
    out = []
    for lo in list(range(0, len(data), block)) + [None]:
        chunk = b"" if lo is None else data[lo:lo + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize)
                   + body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk) & 0xffffffff))
    return b"".join(out)


def run(dumper, path, threads=0, budget=0):
    r = subprocess.run([dumper, path, str(threads), str(budget)], capture_output=True)
    return r.returncode, r.stdout, r.stderr.decode().strip()


def test_single_and_concatenated_members(dumper, tmp_path):
    rng = np.random.default_rng(11)
    a, b = fastq_text(rng, 3_000_000), fastq_text(rng, 700_000)
    p = str(tmp_path / "one.gz")
    open(p, "wb").write(gzip.compress(a, 6))
    rc, out, info = run(dumper, p)
    assert rc == 0 and out == a and info.split() == ["1", "1", "0"]
    p2 = str(tmp_path / "two.gz")
    open(p2, "wb").write(gzip.compress(a, 1) + gzip.compress(b, 9) + gzip.compress(b"", 6))
    rc, out, info = run(dumper, p2)
    assert rc == 0 and out == a + b and info.split() == ["3", "1", "0"]
    assert gzip.open(p2, "rb").read() == out                     # what zlib's own reader returns


def test_trailing_bytes_after_the_last_member_are_ignored_like_gzread(dumper, tmp_path):
    rng = np.random.default_rng(12)
    a = fastq_text(rng, 200_000)
    p = str(tmp_path / "garbage.gz")
    open(p, "wb").write(gzip.compress(a) + b"\0" * 512)          # tar-style zero padding
    rc, out, _ = run(dumper, p)
    assert rc == 0 and out == a


def test_bgzf_blocks_are_inflated_in_parallel(dumper, tmp_path):
    rng = np.random.default_rng(13)
    a = fastq_text(rng, 9_000_000)
    p = str(tmp_path / "blocks.gz")
    open(p, "wb").write(bgzf(a))
    assert gzip.open(p, "rb").read() == a                        # it IS a valid multi-member gzip file
    for threads in (1, 3, 8):
        rc, out, info = run(dumper, p, threads)
        members, used, is_bgzf = info.split()
        assert rc == 0 and out == a and is_bgzf == "1" and int(used) == threads
        assert int(members) == (len(a) + 0xff00 - 1) // 0xff00 + 1
    # a BGZF file followed by a plain member is not pure BGZF: the sequential path must give the same bytes
    p2 = str(tmp_path / "mixed.gz")
    open(p2, "wb").write(bgzf(a[:300_000]) + gzip.compress(a[300_000:400_000]))
    rc, out, info = run(dumper, p2, 4)
    assert rc == 0 and out == a[:400_000] and info.split()[2] == "0"


def test_declined_inputs(dumper, tmp_path):
    rng = np.random.default_rng(14)
    a = fastq_text(rng, 500_000)
    z = gzip.compress(a)
    cases = {
        "plain.fastq": a,                                         # not gzip
        "truncated.gz": z[:len(z) // 2],
        "corrupt.gz": z[:1000] + bytes(64) + z[1064:],
        "badcrc.gz": z[:-8] + struct.pack("<I", (zlib.crc32(a) ^ 1) & 0xffffffff) + z[-4:],
        "bgzf_corrupt.gz": (lambda b: b[:5000] + bytes(x ^ 0x55 for x in b[5000:5040]) + b[5040:])(bgzf(a)),
        "empty.gz": gzip.compress(b""),
    }
    for name, data in cases.items():
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        rc, out, why = run(dumper, p)
        assert rc == 2 and out == b"" and why, (name, rc, why)
    # over the memory budget: declined (the host reader streams it instead)
    p = str(tmp_path / "big.gz")
    open(p, "wb").write(z)
    rc, _, why = run(dumper, p, 0, 100_000)
    assert rc == 2 and "budget" in why
    open(p, "wb").write(bgzf(a))
    rc, _, why = run(dumper, p, 0, 100_000)
    assert rc == 2 and "budget" in why
