"""CPU model of the arithmetic behind k_phred_sum / k_phred_win (fl_phred.cu), against the oracle.

The kernels never walk a read base by base; they rely on one identity: while a running double stays
inside a binade [C, 2C) every add / subtract of a table value moves it by that value rounded to the
binade's grid (IEEE round-to-nearest, no tie), so partial sums may be formed in any order, lane-local
anchors (C for the sum, 0.75 for the window) reproduce the rounding, and a binade crossing has to be
replayed with true adds. This file restates both kernels' schedules in plain Python floats (IEEE
doubles) -- 32 "lanes", 512-base steps for the sum with the room budget and the crossing replay,
window-length steps with values carried in "registers" -- and checks the results bit for bit against
the oracle's sequential loops. It pins the algorithm on the CPU; the GPU tests pin the implementation."""
import math
import random

import pytest

from oracle import oracle as orc


def tables(ws):
    L = orc.lib()
    q = [L.orc_qscore_to_quality(bytes([b])) for b in range(256)]
    return q, [v / ws for v in q]


def expo(x):
    return math.frexp(x)[1] - 1 if x > 0 else -1023


def tie_info(q):
    any_, chars = 0, {}
    for e in range(64):
        for c, v in enumerate(q):
            if 0 < v < 1:
                s = math.ldexp(v, 52 - e)
                if s - math.floor(s) == 0.5:
                    any_ |= 1 << e
                    chars.setdefault(e, []).append(c)
    return any_, chars


def model_sum(qs, ws, q, tie_any, tie_chars, stats):
    """k_phred_first (first min(L, H) bases serially) + k_phred_sum."""
    L, H, T = len(qs), (ws + 15) & ~15, 512
    s = 0.0
    for c in qs[:min(L, H)]:
        s += q[c]
    j = H
    e = expo(s)
    C = math.ldexp(1.0, e) if e > -1000 else 0.0
    acc, budget = [C] * 32, 0
    while j < L:
        n = min(T, L - j)
        chunk = [[qs[j + 16 * l + k] for k in range(16) if 16 * l + k < n] for l in range(32)]
        in_range = 5 <= e < 52
        mode = 0
        if budget == 0 or (tie_any >> max(e, 0)) & 1:
            if budget == 0:
                s += sum(a - C for a in acc)            # exact: grid multiples, below 2C
                acc = [C] * 32
                if in_range:
                    room = math.floor((C + C) - s)
                    budget = (room - 1) // T if room >= 1 else 0
            if not in_range or s <= 0:
                mode = 2
            else:
                if (tie_any >> e) & 1 and (len(tie_chars[e]) > 1 or any(tie_chars[e][0] in ch for ch in chunk)):
                    mode = 2
                if mode == 2 and budget > 0:
                    s += sum(a - C for a in acc)
                    acc, budget = [C] * 32, 0
                if mode == 0 and budget == 0:
                    mode = 1
        for l in range(32):
            for c in chunk[l]:
                acc[l] = acc[l] + q[c]
        if mode == 0:
            budget -= 1
            stats["fast"] += 1
            j += n
            continue
        serial = mode == 2
        if not serial:
            cur, ce, lo, Cc, part = s, e, 0, C, acc[:]
            for rnd in range(6):
                P, t = [], 0.0
                for l in range(32):
                    t += (part[l] - Cc) if l >= lo else 0.0
                    P.append(t)
                cross = [l for l in range(32) if cur + P[l] >= Cc + Cc]
                if not cross:
                    s, e, C = cur + P[31], ce, Cc
                    break
                lx = cross[0]
                t0 = cur + (P[lx] - ((part[lx] - Cc) if lx >= lo else 0.0))
                Cc, ce = Cc + Cc, ce + 1
                v = [t0 if l == lx else Cc for l in range(32)]
                for l in range(lx, 32):
                    for c in chunk[l]:
                        v[l] = v[l] + q[c]
                hit = (tie_any >> ce) & 1 and (len(tie_chars[ce]) > 1 or any(tie_chars[ce][0] in ch for ch in chunk))
                if hit or expo(v[lx]) != ce or rnd == 5:
                    serial = True
                    break
                cur, lo, part = v[lx], lx + 1, v
                stats["cross"] += 1
        if serial:
            stats["serial"] += 1
            for p in range(n):
                s = s + q[qs[j + p]]
            e = expo(s)
            C = math.ldexp(1.0, e) if e > -1000 else 0.0
        acc, budget = [C] * 32, 0
        j += n
    return s + sum(a - C for a in acc)


def model_window(qs, ws, q, a):
    """k_phred_win; returns None when the kernel would hand the read to k_phred_fallback."""
    L = len(qs)
    K = 2 if ws <= 64 else (4 if ws <= 128 else 8)
    ra = []
    for v in a:
        ok = v >= 0 and v * ws <= 1 - 1e-10
        if ok and v > 0:
            sc = math.ldexp(v, 53)
            ok = sc - math.floor(sc) != 0.5
        ra.append((0.5 + v) - 0.5 if ok else float("nan"))
    amax = max(x for x in ra if x == x)
    thr = 0.5 + 2 * amax
    s0 = 0.0
    for c in qs[:ws]:
        s0 += q[c]
    W = s0 / ws
    if not (thr <= W < 1.0):
        return None
    mn = W
    nb = [max(0, min(K, ws - K * l)) for l in range(32)]
    was = [[ra[qs[K * l + k]] for k in range(nb[l])] for l in range(32)]
    j = ws
    while j < L:
        n = min(ws, L - j)
        off = 0.0
        for l in range(32):
            x = m = 0.75
            for k in range(nb[l]):
                p = K * l + k
                if p < n:
                    now = ra[qs[j + p]]
                    x = x + (now - was[l][k])
                    m = min(m, x) if x == x else m
                    was[l][k] = now
            if x != x:
                return None
            cand = (W + off) + (m - 0.75)
            mn = min(mn, cand)
            off += x - 0.75
        W += off
        j += n
    return mn if mn >= thr else None


@pytest.mark.parametrize("ws", [250, 100, 64, 33, 200, 256])
def test_lattice_schedules_reproduce_the_sequential_chains(ws):
    random.seed(ws)
    q, a = tables(ws)
    tie_any, tie_chars = tie_info(q)
    params = orc.make_params(window_size=ws)
    stats = {"fast": 0, "cross": 0, "serial": 0}
    reads, fell_back = [], 0
    for t in range(36):
        L = random.randint(ws + 1, random.choice([1500, 6000, 14000]))
        kind = t % 6
        if kind == 0:
            qs = [random.randint(33 + 40, 33 + 50) for _ in range(L)]            # Q44 ties while the sum is in [512, 1024)
        elif kind == 1:
            qs = [random.choice([33, 34, 126, 112, 122]) for _ in range(L)]
        elif kind == 2:
            qs = [min(126, max(33, int(round(random.gauss(3, 1.5))) + 33)) for _ in range(L)]   # window near / below 0.5
        else:
            mq = random.uniform(5, 40)
            qs = [min(126, max(34, int(round(random.gauss(mq, 4))) + 33)) for _ in range(L)]
        reads.append(qs)
    sc = orc.score([(b"A" * len(qs), bytes(qs)) for qs in reads], params, None)
    for qs, row in zip(reads, sc.parents):
        s = model_sum(qs, ws, q, tie_any, tie_chars, stats)
        assert 100.0 * s / len(qs) == row.mean_q
        mn = model_window(qs, ws, q, a)
        if mn is None:
            fell_back += 1
            continue
        if mn < 0.5 / ws:
            mn = 0.0
        assert 100.0 * mn == row.window_q
    assert stats["fast"] > 0 and stats["cross"] > 0
    if ws in (250, 100, 64, 256):
        assert fell_back < len(reads) // 2          # the fast path is the rule for tie-free window sizes
