"""The reference's own black-box suite (test/test_sort.py, test_trim.py, test_split.py,
test_error_messages.py, test_unit_suffixes.py: 93 unittest cases) run UNMODIFIED against
filtlong_b200/bin/filtlong -- the suite finds its binary at <root>/bin/filtlong, so the staged copy under
oracle/_ref/reftests/ (made by oracle/Makefile; expected numbers have their thousands separators stripped
because the image has no en_US locale, nothing else) gets a bin/ that points at ours. Plus BASELINE
config 1 and the reference's known answers on the committed fixture bytes, through the C ABI."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref", "reftests")
OURS = os.path.join(ROOT, "filtlong_b200", "bin", "filtlong")
FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")


@pytest.mark.gpu
def test_reference_unittest_suite_passes_against_our_binary():
    assert os.path.isdir(os.path.join(STAGE, "test")), "oracle/_ref/reftests missing: run `make -C oracle` where /root/reference exists"
    assert os.path.exists(OURS), "filtlong_b200/bin/filtlong not built"
    link = os.path.join(STAGE, "bin", "filtlong")
    os.makedirs(os.path.dirname(link), exist_ok=True)
    if os.path.lexists(link):
        os.unlink(link)
    os.symlink(OURS, link)
    env = dict(os.environ, LC_ALL="C")
    env.pop("LANG", None)
    r = subprocess.run([sys.executable, "-m", "unittest", "discover", "-s", "test", "-p", "test_*.py"], cwd=STAGE, env=env,
                       capture_output=True, text=True, timeout=3000)
    tail = r.stderr[-3000:]
    m = re.search(r"Ran (\d+) tests", r.stderr)
    assert m and int(m.group(1)) == 93, tail
    assert r.returncode == 0 and "OK" in r.stderr.splitlines()[-1], tail


def _fixture_reads(name):
    return [(s, q) for _, s, q in util.read_fastx(os.path.join(FIX, name))]


@pytest.mark.gpu
def test_baseline_config1_on_the_reference_fixture():
    """BASELINE configs[0]: test/test_sort.fastq --min_length 1 --keep_percent 90 -> target 13 500 bp,
    keeping 15 000 bp, all three reads out; raw values are the survey's hex doubles."""
    from filtlong_b200 import api
    reads = _fixture_reads("test_sort.fastq")
    assert [len(s) for s, _ in reads] == [5000, 5000, 5000]
    ctx, summ = api.score_and_filter(reads, api.make_params(min_length=1, keep_percent=90.0))
    assert (summ.status, summ.target, summ.keeping) == (3, 13500, 15000)
    rows = ctx.row_results()
    assert [int(x) for x in rows["passed_final"]] == [1, 1, 1]
    want = [("0x1.6765056776ee5p+6", "0x1.656d069f0b576p+6"), ("0x1.8bebf07f8e0a2p+6", "0x1.8bd0b23c524bfp+6"),
            ("0x1.831476491630dp+6", "0x1.82b0ce9fc8fd7p+6")]
    for i, (m, w) in enumerate(want):
        assert rows["mean_q"][i] == float.fromhex(m) and rows["window_q"][i] == float.fromhex(w), i
    assert [round(float(x), 2) for x in rows["final_score"]] == [0.00, 70.70, 61.54]
    ctx.close()


@pytest.mark.gpu
def test_reference_known_answers_with_the_assembly_fixture():
    """test_sort / test_trim / test_split against test_reference.fasta: 199 964 16-mers, the score order
    1 > 3 > 2, the drifted window qualities 79.999999999999986 / 59.99.. / 19.99.., and the child names
    test_trim.py:74-112 / test_split.py:74-223 pin."""
    from filtlong_b200 import api
    asm = [s for _, s, _ in util.read_fastx(os.path.join(FIX, "test_reference.fasta"))]
    ctx, summ = api.score_and_filter(_fixture_reads("test_sort.fastq"), api.make_params(min_length=1, keep_percent=90.0), assembly=asm)
    assert ctx.kmers_count() == 199964
    rows = ctx.row_results()
    assert [round(float(x), 2) for x in rows["final_score"]] == [70.71, 0.00, 63.23]
    ctx.close()
    ctx, _ = api.score_and_filter(_fixture_reads("test_split.fastq"), api.make_params(min_length=1), assembly=asm)
    assert [repr(float(x)) for x in ctx.read_results()["window_q"][1:]] == ["79.99999999999999", "59.999999999999964", "19.99999999999993"]
    ctx.close()
    ctx, _ = api.score_and_filter(_fixture_reads("test_trim.fastq"), api.make_params(min_length=1, trim=True), assembly=asm)
    rows = ctx.row_results()
    spans = [(int(p), int(s) + 1, int(e)) for p, s, e in zip(rows["parent"], rows["start"], rows["end"])]
    assert spans == [(0, 1, 1300), (1, 21, 701), (2, 1, 970), (3, 13, 1885)]       # test_trim_2_21-701, _3_1-970, _4_13-1885
    ctx.close()
    ctx, _ = api.score_and_filter(_fixture_reads("test_split.fastq"), api.make_params(min_length=1, split=200), assembly=asm)
    rows = ctx.row_results()
    assert len(rows["parent"]) == 5 and int((rows["end"] - rows["start"]).sum()) == 11400   # test_split.py: --split 200 -> 5 reads, 11 400 bp
    ctx.close()
