"""CPU tests of the host side: the C ABI library loads and exports every declared symbol, parameter derivation,
error behaviour without a GPU, the sort emulation, the synthetic BWT builder, the CLI's option handling and the
two-rank sharding logic (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read_fasta


@pytest.fixture(scope="module")
def api():
    from longreadselfcorrect_b200 import api as a
    a.lib()
    return a


def test_library_exports_every_declared_symbol(api):
    hdr = open(os.path.join(ROOT, "include", "pbsc.h")).read()
    declared = set(re.findall(r"\b(pbsc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = api.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(api.EXPORTED)


@pytest.mark.parametrize("name,kw", [("tiny", dict(coverage=30, genome=5)), ("tiny100", dict(coverage=100, genome=10))])
def test_params_and_threshold_table_match_reference(api, name, kw):
    p = api.Params.make(no_dp=True, **kw)
    assert p.threshold_table_text() == open(os.path.join(GOLDEN, f"{name}.threshold-table")).read()
    if name == "tiny":
        assert p.pool == [5, 9, 15, 17, 19] and p.c.start_kmer == 17 and list(p.c.offset) == [0, 0, -2]
    else:
        assert p.pool == [5, 9, 15, 19, 23] and p.c.start_kmer == 19 and list(p.c.offset) == [0, 4, -4]


def test_option_validation_messages(api):
    for kw, msg in ((dict(coverage=0), "invalid number of coverage"), (dict(genome=7), "invalid genome size"),
                    (dict(mode=3), "invalid mode"), (dict(error_rate=1.5), "invalid error rate"),
                    (dict(max_leaves=64), "max leaves")):
        with pytest.raises(api.PbscError) as e:
            api.Params.make(**kw)
        assert msg in str(e.value)


def test_no_cpu_fallback(api):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.PbscError) as e:
        api.Index.load(os.path.join(GOLDEN, "tiny"))
    assert e.value.code == -4   # PBSC_ERR_CUDA: the product path fails loudly, it never computes on the CPU


def test_sort_emulation_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_sort_emul.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok")


@pytest.mark.parametrize("flags", [[], ["-DPBSC_FUSED_UPDATE"]])
def test_rank_primitives_match_naive_counting(tmp_path, flags):
    """occ / occ_pair / count_dollars / update_interval of csrc/fm_table.cuh — the code every kernel calls — compiled for the
    host, against naive counting over random BWTs with '$' symbols (default build and the experimental fused update)."""
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", cuda_inc] + flags + [os.path.join(ROOT, "tests", "cpp", "test_fm_occ.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout


@pytest.mark.parametrize("flags", [[], ["-DPBSC_FUSED_UPDATE"]])
def test_backward_search_primitives_match_reference_findinterval(tmp_path, flags):
    """init_interval / update_interval of csrc/fm_table.cuh compiled for the host, on the reference-built tiny.bwt / tiny.rbwt,
    against the (lower, upper) pairs the reference's own BWTAlgorithms::findInterval printed for 2000 queries (fm_dump),
    including its raw values at the early break."""
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", cuda_inc] + flags + [os.path.join(ROOT, "tests", "cpp", "test_fm_search.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, GOLDEN], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok: 4000 intervals"), r.stdout


def test_lf_step_spells_every_read_of_the_reference_index(tmp_path):
    """lf_step of csrc/fm_table.cuh (BWT symbol + LF-mapping from one sector: the step of dp_retrieve_kernel) compiled for the
    host: LF-walking from the '$' rows of the reference-built tiny.bwt / tiny.rbwt spells each of the 396 reads exactly once."""
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", cuda_inc, os.path.join(ROOT, "tests", "cpp", "test_fm_lf.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, GOLDEN], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok: 396 reads"), r.stdout


def test_integer_ratio_rule_equals_the_reference_double_comparison(tmp_path):
    """eval4 (csrc/pbsc_walk_thread.cuh) replaces `(double)kmerFreq/(double)maxfreq >= cutoff` by an integer cross-multiplication;
    the two agree on 142 M (a, b) pairs including every cutoff boundary for b up to 2^31 - 1."""
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "test_ratio_rule.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout


def test_thread_per_alignment_dp_matches_oracle_extendmatch(tmp_path):
    """The fill / traceback templates that dp_align_thread_kernel instantiates (csrc/pbsc_dp_thread.cuh), compiled for the host
    over accessors that mimic the device storage, against the oracle's Overlapper::extendMatch: forward and reverse-complement
    placements, arbitrary band origins, clipped and unclipped bands, tie-rich two-letter and homopolymer strings."""
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", os.path.join(ROOT, "tests", "cpp", "test_dp_thread.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe, "12000"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and " failed 0" in r.stdout, r.stdout
    # ... and against the reference's own Overlapper: the 360 extendMatch records answered by oracle/_ref/dp_dump
    r = subprocess.run([exe, "--vectors", os.path.join(GOLDEN, "dp_units.txt"), os.path.join(GOLDEN, "dp_units.ref.txt")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "tested 360" in r.stdout and " failed 0" in r.stdout, r.stdout


def test_multiple_alignment_column_model_matches_reference_vectors(tmp_path):
    """msa::consensus (csrc/pbsc_dp_msa.cuh), the per-job code of dp_msa_kernel, compiled for the host: the 60 pile-ups answered
    by the reference's own MultipleAlignment (oracle/_ref/dp_dump) and random pile-ups against the oracle's padded rows."""
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpp", "test_dp_msa.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "--vectors", os.path.join(GOLDEN, "dp_units.txt"), os.path.join(GOLDEN, "dp_units.ref.txt")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "tested 60" in r.stdout and " failed 0" in r.stdout, r.stdout
    r = subprocess.run([exe, "1500"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and " failed 0" in r.stdout, r.stdout


def test_bwt_builder_equals_reference_index(oracle_bin, tmp_path):
    """The torch suffix sorter used for synthetic inputs yields the intervals of the reference's own (ropebwt2) index."""
    from longreadselfcorrect_b200 import bwt_build
    recs = read_fasta(os.path.join(GOLDEN, "tiny.reads.fa"))
    lut = np.zeros(256, dtype=np.uint8)
    lut[ord("C")], lut[ord("G")], lut[ord("T")] = 1, 2, 3
    codes = np.concatenate([lut[np.frombuffer(s.encode(), dtype=np.uint8)] for _, s in recs])
    off = np.zeros(len(recs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for _, s in recs])
    bwt_build.build_index_files(str(tmp_path / "mine"), codes, off, device="cpu")
    for ext in ("bwt", "rbwt"):
        got = subprocess.run([oracle_bin, "findinterval", str(tmp_path / f"mine.{ext}"), os.path.join(GOLDEN, "tiny.fm_queries.txt")],
                             check=True, stdout=subprocess.PIPE, text=True).stdout
        assert got == open(os.path.join(GOLDEN, f"tiny.fm_{ext}.txt")).read()


def test_cli_option_errors():
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    if not os.path.exists(exe):
        pytest.skip("pbcorrect not built")
    r = subprocess.run([exe, "-o", "/tmp/pbsc_cli_test", "-g", "7", "reads.fa"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no prefix" in r.stderr and "invalid genome size: 7, must be (5/10/100)[m]" in r.stderr
    assert "Usage: StriDe PacBioSelfCorrection" in r.stderr
    r = subprocess.run([exe, "-p", "x", "-o", "/tmp/pbsc_cli_test"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "missing arguments" in r.stderr
    # default options (DP fallback on) are accepted; without a GPU or an index the run fails loudly, never on a CPU path
    r = subprocess.run([exe, "-p", "x", "-o", "/tmp/pbsc_cli_test", "reads.fa"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and ("no CUDA device" in r.stderr or "x.bwt" in r.stderr) and "--nodp" not in r.stderr


def test_balanced_ranges():
    from longreadselfcorrect_b200 import sharding
    rng = np.random.default_rng(3)
    lens = rng.integers(500, 50000, size=1000)
    for n in (1, 2, 3, 8):
        rg = sharding.balanced_ranges(lens, n)
        assert rg[0][0] == 0 and rg[-1][1] == 1000 and all(rg[i][1] == rg[i + 1][0] for i in range(n - 1))
        sums = [int(lens[b:e].sum()) for b, e in rg]
        assert max(sums) - min(sums) <= 2 * int(lens.max())
    assert sharding.balanced_ranges([], 2) == [(0, 0), (0, 0)]
    assert sharding.balanced_ranges([5], 4)[-1][1] == 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from longreadselfcorrect_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(11)
lens = rng.integers(10, 200, size=57)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
buf = rng.integers(0, 4, size=int(off[-1])).astype(np.uint8)
sbuf, soff, (b, e) = sharding.shard(buf, off, rank, 2)
# stand-in for the per-read hot path: any pure per-read function; here a checksum per read
mine = [int(sbuf[int(soff[i]):int(soff[i + 1])].astype(np.int64).dot(np.arange(1, int(soff[i + 1] - soff[i]) + 1))) for i in range(e - b)]
gathered = [None, None]
dist.all_gather_object(gathered, (b, e, mine))
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
if rank == 0:
    out = []
    for b_, e_, m in sorted(gathered):
        out += m
    ref = [int(buf[int(off[i]):int(off[i + 1])].astype(np.int64).dot(np.arange(1, int(lens[i]) + 1))) for i in range(57)]
    assert out == ref and float(t) == 2.0
    print("ok")
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ok" in outs[0][0]
