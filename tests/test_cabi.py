"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
refuses to run without a GPU (no fallback), and its host-only helpers (packer, Phred tables,
synthetic generators) behave. No compute call needs a device here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from filtlong_b200 import api, capi
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "filtlong_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = header_symbols()
    assert len(declared) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for s in declared:
        assert s in exported, "header declares %s but the library does not export it" % s
        assert hasattr(L, s)
    assert sorted(n for n, _, _ in capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with the header"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.FLError) as e:
        api.Context()
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "filtlong_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
                assert "oracle/" not in txt.replace("see oracle for derivation", ""), f


def test_packer_layout():
    seq = b"ACGTNacgtRYGGTTAACCGGTTACGTACGTAAGGCCTTA"    # 40 bases incl. non-ACGT
    qual = bytes(range(40, 80))
    hb = api.HostBatch([seq, b"TTTT"], [qual, b"IIII"], want_seq=True, want_nmask=True)
    assert list(hb.off) == [0, 64] and hb.padded_bases == 128
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for i, ch in enumerate(seq.upper()):
        got = (int(hb.seq2b[i >> 4]) >> (30 - 2 * (i & 15))) & 3
        assert got == code.get(ch, 0)
        assert ((int(hb.nmask[i >> 5]) >> (i & 31)) & 1) == (0 if ch in code else 1)
    assert bytes(hb.qual[:40]) == qual and bytes(hb.qual[64:68]) == b"IIII"
    # the first word is directly the reference's forward 16-mer of the first 16 bases (kmers.cpp:222-229)
    L = orc.lib()
    k = 0
    for ch in seq[:16]:
        k = ((k << 2) | L.orc_base_fwd(bytes([ch]))) & 0xFFFFFFFF
    assert int(hb.seq2b[0]) == k


def test_packer_random_sequences_and_offsets():
    """fl_pack_sequence against a plain restatement of kmers.cpp:176-196 (any character, any length,
    any offset -- the arena always uses multiples of 64, the packer itself takes any)."""
    rng = np.random.default_rng(5)
    L = capi.lib()
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-*", dtype=np.uint8)
    code = np.zeros(256, dtype=np.uint32)
    other = np.ones(256, dtype=np.uint32)
    for i, ch in enumerate(b"AaCcGgTt"):
        code[ch], other[ch] = i >> 1, 0
    for trial in range(200):
        n = int(rng.integers(0, 300))
        off = int(rng.integers(0, 200)) if trial % 2 else 64 * int(rng.integers(0, 4))
        seq = alphabet[rng.integers(0, len(alphabet), size=n)] if trial % 3 else rng.integers(1, 256, size=n).astype(np.uint8)
        qual = rng.integers(33, 127, size=n).astype(np.uint8)
        words = (off + n + 63) // 16 + 2
        seq2b = np.zeros(words, dtype=np.uint32)
        nmask = np.zeros(words, dtype=np.uint32)
        qout = np.zeros(off + n + 64, dtype=np.uint8)
        L.fl_pack_sequence(seq.tobytes(), qual.tobytes(), n, off, capi.ptr(seq2b), capi.ptr(qout), capi.ptr(nmask))
        want2, wantm = np.zeros_like(seq2b), np.zeros_like(nmask)
        for i in range(n):
            b = off + i
            want2[b >> 4] |= code[seq[i]] << np.uint32(30 - 2 * (b & 15))
            wantm[b >> 5] |= other[seq[i]] << np.uint32(b & 31)
        assert np.array_equal(seq2b, want2), (trial, n, off)
        assert np.array_equal(nmask, wantm), (trial, n, off)
        assert bytes(qout[off:off + n]) == qual.tobytes() and not qout[:off].any()


def test_anchored_table_mapping():
    """The position-anchored probe table: the four 16-mers starting at read positions
    4g .. 4g+3 fall into ONE 32-byte sector (8 words), into four different 64-bit quarters, and the map
    (16-mer, alignment) -> bit is injective (no two members share a bit, so membership stays exact)."""
    L = capi.lib()
    rng = np.random.default_rng(9)

    def slot(kmer, pos):
        w, b = C.c_uint32(), C.c_uint32()
        L.fl_anchor_slot_host(int(kmer), pos & 3, C.byref(w), C.byref(b))
        return w.value, b.value

    code = rng.integers(0, 4, size=4000)
    kmers = []
    for i in range(len(code) - 15):
        k = 0
        for c in code[i:i + 16]:
            k = (k << 2) | int(c)
        kmers.append(k)
    for g in range(0, len(kmers) - 3, 4):
        slots = [slot(kmers[g + j], g + j) for j in range(4)]
        assert len({w >> 3 for w, _ in slots}) == 1                       # one sector
        assert sorted((w & 7) >> 1 for w, _ in slots) == [0, 1, 2, 3]    # a quarter per alignment
        assert all(w < (1 << 29) and b < 32 for w, b in slots)
    # injective per alignment: invert the slot back to the 16-mer
    for _ in range(2000):
        k = int(rng.integers(0, 1 << 32))
        for pos in range(4):
            w, b = slot(k, pos)
            r = 3 - pos
            key, q = w >> 3, (w & 7) >> 1
            rest = ((w & 1) << 5) | b
            assert q == r
            sh = 6 - 2 * r
            top, bottom = rest >> sh, rest & ((1 << sh) - 1)
            back = ((top << (32 - 2 * r)) if r else 0) | (key << sh) | bottom
            assert back == k


def test_phred_tables_match_reference_formula():
    q = np.zeros(256)
    a = np.zeros(256)
    capi.lib().fl_phred_luts(250, capi.ptr(q), capi.ptr(a))
    L = orc.lib()
    for b in range(256):
        ref = L.orc_qscore_to_quality(bytes([b]))
        assert (q[b] == ref) or (np.isnan(q[b]) and np.isnan(ref))
        assert a[b] == ref / 250 or (np.isinf(ref))


def test_synth_generators_are_deterministic_and_sane():
    L = capi.lib()
    n = 50
    length = np.full(n, 3000, dtype=np.int32)
    off = (np.arange(n, dtype=np.uint64) * 3008)
    qbar = np.full(n, 14, dtype=np.uint8)
    q1 = np.zeros(n * 3008, dtype=np.uint8)
    q2 = np.zeros_like(q1)
    L.fl_synth_qual_host(1, n, capi.ptr(off), capi.ptr(length), capi.ptr(qbar), 0, capi.ptr(q1))
    L.fl_synth_qual_host(1, n, capi.ptr(off), capi.ptr(length), capi.ptr(qbar), 0, capi.ptr(q2))
    assert np.array_equal(q1, q2)
    vals = q1.reshape(n, 3008)[:, :3000].astype(int) - 33
    assert vals.min() >= 1 and vals.max() <= 50
    assert abs(vals.mean() - 14) < 0.2 and 3.5 < vals.std() < 4.5
    g = np.zeros(1000 // 16 + 1, dtype=np.uint32)
    L.fl_synth_genome_host(2, 1000, capi.ptr(g))
    codes = [(int(g[i >> 4]) >> (30 - 2 * (i & 15))) & 3 for i in range(1000)]
    assert sorted(set(codes)) == [0, 1, 2, 3]
    # a read with no errors on the forward strand reproduces the genome slice
    d = capi.SynthReads()
    rl = np.array([200], dtype=np.int32); ro = np.array([0], dtype=np.uint64)
    st = np.array([100], dtype=np.uint64); sd = np.array([0], dtype=np.uint8); er = np.array([0], dtype=np.uint32)
    jp = np.array([0], dtype=np.int32); jl = np.array([0], dtype=np.int32)
    d.n = 1; d.genome_bases = 1000
    d.off, d.len, d.start, d.strand, d.err_ppm, d.junk_pos, d.junk_len = map(capi.ptr, (ro, rl, st, sd, er, jp, jl))
    out = np.zeros(256 // 16, dtype=np.uint32)
    L.fl_synth_reads_host(3, capi.ptr(g), C.byref(d), 0, capi.ptr(out))
    rc = [(int(out[i >> 4]) >> (30 - 2 * (i & 15))) & 3 for i in range(200)]
    assert rc == codes[100:300]
    sd[0] = 1
    out[:] = 0
    L.fl_synth_reads_host(3, capi.ptr(g), C.byref(d), 0, capi.ptr(out))
    rc = [(int(out[i >> 4]) >> (30 - 2 * (i & 15))) & 3 for i in range(200)]
    span = 200 + 200 // 8 + 64                        # fl_synth_span: the template a read may consume
    assert rc == [3 - c for c in codes[100:100 + span]][::-1][:200]
    # ONT model: 50/25/25 substitutions / insertions / deletions at the per-read rate, adapters at both ends
    g2 = np.zeros(20000 // 16 + 1, dtype=np.uint32)
    L.fl_synth_genome_host(2, 20000, capi.ptr(g2))
    gc = np.array([(int(g2[i >> 4]) >> (30 - 2 * (i & 15))) & 3 for i in range(20000)])
    rl[0] = 8000; st[0] = 500; sd[0] = 0; er[0] = 100000
    a5 = np.array([40], dtype=np.int32); a3 = np.array([30], dtype=np.int32)
    d.flags = capi.SYNTH_INDELS; d.genome_bases = 20000
    d.adap5, d.adap3 = capi.ptr(a5), capi.ptr(a3)
    out = np.zeros(8064 // 16, dtype=np.uint32)
    L.fl_synth_reads_host(7, capi.ptr(g2), C.byref(d), 5, capi.ptr(out))
    out2 = np.zeros_like(out)
    L.fl_synth_reads_host(7, capi.ptr(g2), C.byref(d), 5, capi.ptr(out2))
    assert np.array_equal(out, out2)
    rd = np.array([(int(out[i >> 4]) >> (30 - 2 * (i & 15))) & 3 for i in range(8000)])
    # 16-mers of the read found in the template: far fewer than for an error-free copy, far more than chance
    tmpl = {tuple(gc[i:i + 16]) for i in range(500, 500 + 9100)}
    hits = sum(tuple(rd[i:i + 16]) in tmpl for i in range(40, 8000 - 30 - 16))
    assert 0.05 * 7900 < hits < 0.45 * 7900, hits        # (1 - 0.1)^16 = 0.185 of the 16-mers survive a 10 % error rate
    # assembly with runs of N, and its text form
    nc, cb = 3, 5000
    pad = int(L.fl_padded_len(cb))
    seq = np.zeros(nc * pad // 16, dtype=np.uint32); nm = np.zeros(nc * pad // 32, dtype=np.uint32)
    L.fl_synth_assembly_host(4, nc, cb, 200000, capi.ptr(seq), capi.ptr(nm))
    aoff = (np.arange(nc, dtype=np.uint64) * pad); alen = np.full(nc, cb, dtype=np.int32)
    txt = np.zeros(nc * pad, dtype=np.uint8)
    L.fl_synth_ascii_host(nc, capi.ptr(aoff), capi.ptr(alen), capi.ptr(seq), capi.ptr(nm), capi.ptr(txt))
    contig0 = txt[:cb].tobytes()
    assert set(contig0) <= set(b"ACGTN") and txt[cb:pad].max() == 0
    n_frac = sum(txt[c * pad:c * pad + cb].tobytes().count(b"N") for c in range(nc)) / (nc * cb)
    assert 0.03 < n_frac < 0.5, n_frac
    # N bases are whole 1024-base blocks (of the arena) and carry code 0
    for c in range(nc):
        t = txt[c * pad:c * pad + cb]
        for b0 in range(0, cb, 1024):
            blk = t[b0:min(b0 + 1024, cb)]
            assert (blk == ord("N")).all() or not (blk == ord("N")).any() or (c * pad + b0) % 1024 != 0
    # packing the text back through the library's host packer reproduces codes and mask
    s2 = np.zeros_like(seq); n2 = np.zeros_like(nm)
    for c in range(nc):
        L.fl_pack_sequence(txt[c * pad:c * pad + cb].tobytes(), None, cb, c * pad, capi.ptr(s2), None, capi.ptr(n2))
    assert np.array_equal(s2, seq) and np.array_equal(n2, nm)
