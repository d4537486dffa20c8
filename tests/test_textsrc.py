"""The host side of the device-text path, on the CPU (filtlong_b200/csrc/host/textsrc.cpp + fastx.cpp): a plain or gzip
file becomes one byte range, the chunk plan cuts it only at record starts (FASTQ: an '@' line whose third line starts
with '+' and whose fourth is as long as its second; FASTA: a '>' line, wrapped records included), and the memory-backed
FastxReader -- what Kmers::add_reference continues with where the device hands a chunk back -- parses the bytes from any
record start exactly like the file-backed one parses the file (kseq semantics: reference src/kseq.h:161-224)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "filtlong_b200", "csrc", "host")


@pytest.fixture(scope="module")
def tools(tmp_path_factory):
    d = tmp_path_factory.mktemp("ts")
    out = {}
    for name, srcs in (("textsrc_dump", ["textsrc.cpp", "gzmem.cpp", "fastx.cpp"]), ("fastx_offsets_dump", ["fastx.cpp"])):
        exe = str(d / name)
        r = subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", name + ".cpp")] + [os.path.join(HOST, s) for s in srcs]
                           + ["-lz", "-lpthread", "-o", exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out[name] = exe
    return out


def dump(exe, path, target, parse_from, env=None):
    r = subprocess.run([exe, str(path), str(target), str(parse_from)], capture_output=True, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split(b"\n")
    head = lines[0].split()
    chunks = [tuple(int(x) for x in l.split()[1:]) for l in lines if l.startswith(b"CHUNK")]
    recs = [tuple(l[4:].split(b"\x01")) for l in lines if l.startswith(b"REC ")]
    end = [l for l in lines if l.startswith(b"END")]
    return head, chunks, recs, (int(end[0].split()[1]) if end else None), any(l == b"NOPLAN" for l in lines)


def fastq(rng, n, tricky=True):
    recs = []
    for i in range(n):
        L = int(rng.integers(1, 600))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L))
        qual = bytes(rng.integers(33, 127, size=L).astype(np.uint8))
        if tricky and i % 5 == 0:
            qual = b"@" + qual[1:]                      # a quality line that looks like a header
        if tricky and i % 7 == 0 and L > 1:
            qual = qual[:1] + b"+" + qual[2:] if L > 2 else qual
        recs.append((b"r%d" % i, [b"", b"some comment", b"x\ty"][i % 3], seq, qual))
    return recs


def fastq_bytes(recs):
    return b"".join(b"@" + n + ((b" " + c) if c else b"") + b"\n" + s + b"\n+\n" + q + b"\n" for n, c, s, q in recs)


def test_fastq_chunks_cut_only_at_record_starts_and_memory_reader_matches(tools, tmp_path):
    rng = np.random.default_rng(5)
    recs = fastq(rng, 4000)
    text = fastq_bytes(recs)
    starts = set(np.cumsum([0] + [len(fastq_bytes([r])) for r in recs]).tolist())
    for name, data in (("a.fastq", text), ("a.fastq.gz", gzip.compress(text))):
        p = tmp_path / name
        p.write_bytes(data)
        head, chunks, parsed, end, noplan = dump(tools["textsrc_dump"], p, 100000, 0)
        assert head == [b"OPEN", str(len(text)).encode(), b"1", b"1" if name.endswith(".gz") else b"0"] and not noplan
        assert chunks[0][0] == 0 and chunks[-1][1] == len(text) and len(chunks) > 5
        assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
        assert all(c[0] in starts and 0 < c[1] - c[0] <= 100000 for c in chunks)
        assert end == -1 and parsed == recs
        # from the start of any chunk on: exactly the records from there on
        k = sorted(starts).index(chunks[3][0])
        _, _, tail_recs, end, _ = dump(tools["textsrc_dump"], p, 100000, chunks[3][0])
        assert end == -1 and tail_recs == recs[k:]
    # FL_GZ_HOST leaves gzip to the streaming reader
    assert dump(tools["textsrc_dump"], tmp_path / "a.fastq.gz", 100000, 0, {"FL_GZ_HOST": "1"})[0] == [b"DECLINED"]


def test_wrapped_fasta_chunks_and_memory_reader(tools, tmp_path):
    rng = np.random.default_rng(6)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=int(L))) for L in rng.integers(1, 30000, size=60)]
    wrap = lambda q: b"".join(q[i:i + 60] + b"\n" for i in range(0, len(q), 60))
    text = b"".join(b">c%d description here\n" % i + wrap(q) for i, q in enumerate(seqs))
    p = tmp_path / "asm.fasta"
    p.write_bytes(text)
    head, chunks, parsed, end, noplan = dump(tools["textsrc_dump"], p, 200000, 0)
    assert head[2] == b"2" and not noplan and len(chunks) >= 4
    assert all(text[c[0]:c[0] + 1] == b">" for c in chunks) and chunks[-1][1] == len(text)
    assert end == -1 and [r[2] for r in parsed] == seqs and [r[0] for r in parsed] == [b"c%d" % i for i in range(60)]
    assert all(r[1] == b"description here" and r[3] == b"" for r in parsed)
    # a record larger than the chunk target: no plan (the host reader streams the file)
    assert dump(tools["textsrc_dump"], p, 20000, 0)[4]


def test_memory_reader_equals_file_reader_on_untidy_input(tools, tmp_path):
    """CR LF, blank lines, multi-line records, a '+' line repeating the name, a truncated last record: both readers are one
    parser over two sources, so record for record and error code for error code they must agree."""
    rng = np.random.default_rng(7)
    recs = fastq(rng, 300, tricky=False)
    parts = []
    for i, (n, c, s, q) in enumerate(recs):
        hdr = b"@" + n + ((b" " + c) if c else b"")
        if i % 4 == 1 and len(s) > 10:
            h = len(s) // 2
            parts.append(hdr + b"\n" + s[:h] + b"\n" + s[h:] + b"\n+" + n + b"\n" + q[:h] + b"\n" + q[h:] + b"\n")
        elif i % 4 == 2:
            parts.append(hdr + b"\r\n" + s + b"\r\n+\r\n" + q + b"\r\n")
        elif i % 4 == 3:
            parts.append(b"\n" + hdr + b"\n" + s + b"\n+\n" + q + b"\n")
        else:
            parts.append(hdr + b"\n" + s + b"\n+\n" + q + b"\n")
    for name, data in (("tidy_end.fastq", b"".join(parts)), ("cut.fastq", b"".join(parts)[:-37])):
        p = tmp_path / name
        p.write_bytes(data)
        _, _, mem, end_mem, _ = dump(tools["textsrc_dump"], p, 1 << 30, 0)
        r = subprocess.run([tools["fastx_offsets_dump"], str(p)], capture_output=True)
        file_lines = r.stdout.split(b"\n")
        file_recs = [l.split(b"\t") for l in file_lines if l and not l.startswith(b"END")]
        end_file = int([l for l in file_lines if l.startswith(b"END")][0].split()[1])
        assert end_mem == end_file
        assert [(m[0], len(m[1]), len(m[2]), len(m[3])) for m in mem] == [(f[0], int(f[1]), int(f[2]), int(f[3])) for f in file_recs]
    assert end_mem in (-1, -2)
