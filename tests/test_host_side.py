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


@pytest.mark.parametrize("name,fasta", [("tiny", "tiny.reads.fa"), ("index_edge", "index_edge.fa")])
def test_index_builder_model_reproduces_reference_files(name, fasta, tmp_path):
    """oracle/index_model.py (the GPU builder's algorithm, step by step, in numpy) writes the reference's four index files byte
    for byte: tests/golden/<name>.{bwt,rbwt,sai,rsai} were written by `oracle/_ref/stride index`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import index_model
    codes, off = index_model.read_fasta_codes(os.path.join(GOLDEN, fasta))
    index_model.build(str(tmp_path / "m"), codes, off)
    for ext in ("bwt", "rbwt", "sai", "rsai"):
        assert open(tmp_path / f"m.{ext}", "rb").read() == open(os.path.join(GOLDEN, f"{name}.{ext}"), "rb").read(), ext


def test_onlyseed_barcode_check_matches_reference(tmp_path):
    """csrc/pbsc_bcode.h (what `pbcorrect --onlyseed -b FILE` runs on the host) on the reference's own seeds of tiny.reads.fa and a
    synthetic barcode file: DIR/total.seed and the TOTAL line equal what the unmodified reference wrote
    (tests/golden/make_onlyseed_golden.py)."""
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_bcode.cpp"), "-o", exe, "-lz"], check=True)
    r = subprocess.run([exe, os.path.join(GOLDEN, "tiny.reads.fa"), os.path.join(GOLDEN, "tiny.seeds.tsv"), os.path.join(GOLDEN, "tiny.barcode.txt"),
                        str(tmp_path / "total.seed")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0
    assert r.stdout == open(os.path.join(GOLDEN, "tiny.onlyseed.stdout")).read()
    assert open(tmp_path / "total.seed").read() == open(os.path.join(GOLDEN, "tiny.onlyseed.total.seed")).read()


def test_kmercheck_summaries_match_reference(oracle_bin, tmp_path):
    """`kmercheck` on the CPU: k-mer frequencies from the oracle's backward search (both strands of the reference-built tiny index),
    barcode arithmetic and five-number summaries from csrc/pbsc_bcode.h; DIR/total.box and DIR/value.box equal what the unmodified
    reference wrote (tests/golden/tiny.kmercheck.*, `stride kmercheck -c 30 -l 15 -u 23 -s 4`)."""
    exe = str(tmp_path / "t")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_bcode.cpp"), "-o", exe, "-lz"], check=True)
    reads = dict(read_fasta(os.path.join(GOLDEN, "tiny.reads.fa")))
    comp = str.maketrans("ACGT", "TGCA")
    recs, fwd_q, rvc_q = [], [], []
    blocks_seen = {}
    for line in open(os.path.join(GOLDEN, "tiny.barcode.txt")):
        f = line.split()
        if len(f) != 9:
            continue
        rid, a, b = f[0], int(f[1]), int(f[2])
        bi = blocks_seen.get(rid, 0)
        blocks_seen[rid] = bi + 1
        seq = reads[rid]
        for k in range(15, 24, 4):
            for pos in range(a, b - k + 1):
                w = seq[pos:pos + k]
                recs.append((rid, bi, k, pos))
                fwd_q.append(w[::-1])                      # findInterval(RBWT, reverse(w))
                rvc_q.append(w[::-1].translate(comp))      # findInterval(BWT, reverse-complement(w))
    def sizes(ext, qs):
        qf = tmp_path / f"q.{ext}"
        qf.write_text("\n".join(qs) + "\n")
        out = subprocess.run([oracle_bin, "findinterval", os.path.join(GOLDEN, f"tiny.{ext}"), str(qf)], check=True, stdout=subprocess.PIPE, text=True).stdout
        res = []
        for l in out.splitlines():
            t = l.split()
            lo, hi = int(t[-2]), int(t[-1])
            res.append(hi - lo + 1 if hi >= lo else 0)
        return res
    fs, rs = sizes("rbwt", fwd_q), sizes("bwt", rvc_q)
    assert len(fs) == len(recs) == len(rs)
    with open(tmp_path / "freqs.tsv", "w") as f:
        for (rid, bi, k, pos), a, b in zip(recs, fs, rs):
            assert a + b >= 1
            f.write(f"{rid}\t{bi}\t{k}\t{pos}\t{a + b}\n")
    r = subprocess.run([exe, "--kmercheck", os.path.join(GOLDEN, "tiny.reads.fa"), os.path.join(GOLDEN, "tiny.barcode.txt"), str(tmp_path / "freqs.tsv"), "30", "15", "23", "4",
                        str(tmp_path / "total.box"), str(tmp_path / "value.box")])
    assert r.returncode == 0
    assert open(tmp_path / "total.box").read() == open(os.path.join(GOLDEN, "tiny.kmercheck.total.box")).read()
    assert open(tmp_path / "value.box").read() == open(os.path.join(GOLDEN, "tiny.kmercheck.value.box")).read()


def test_cli_option_errors():
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    if not os.path.exists(exe):
        pytest.skip("pbcorrect not built")
    r = subprocess.run([exe, "-o", "/tmp/pbsc_cli_test", "-g", "7", "reads.fa"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no prefix" in r.stderr and "invalid genome size: 7, must be (5/10/100)[m]" in r.stderr
    assert "Usage: StriDe PacBioSelfCorrection" in r.stderr
    r = subprocess.run([exe, "-p", "x", "-o", "/tmp/pbsc_cli_test"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "missing arguments" in r.stderr
    r = subprocess.run([exe, "-p", "x", "-o", "/tmp/pbsc_cli_test", "--onlyseed", "reads.fa"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no barcode" in r.stderr
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
import os, subprocess, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import read_fasta
from longreadselfcorrect_b200 import parity, sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
golden = os.path.join({root!r}, "tests", "golden")
reads = read_fasta(os.path.join(golden, "tiny.reads.fa"))[:120]
buf = np.frombuffer("".join(s for _, s in reads).encode(), dtype=np.uint8)
off = np.concatenate(([0], np.cumsum([len(s) for _, s in reads]))).astype(np.uint64)
sbuf, soff, (b, e) = sharding.shard(buf, off, rank, 2)
# the per-read hot path of this rank's shard: here the CPU oracle stands where a GPU rank calls pbsc_correct_batch
d = {tmp!r} + "/rank%d" % rank
os.makedirs(d)
fa = d + "/shard.fa"
with open(fa, "w") as f:
    for (rid, _), i in zip(reads[b:e], range(e - b)):
        f.write(">%s\n%s\n" % (rid, sbuf[int(soff[i]):int(soff[i + 1])].tobytes().decode()))
subprocess.run([{oracle!r}, "pbcorrect", "--threads", "2", "-p", os.path.join(golden, "tiny"), "-o", d + "/o", "-c", "30", "-g", "5", fa], check=True,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
mine = (b, e, open(d + "/o/correct.fa").read(), open(d + "/o/discard.fa").read())
gathered = [None, None]
dist.all_gather_object(gathered, mine)
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
if rank == 0:
    parts = sorted(gathered)
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == len(reads)      # contiguous ranges that tile the read set
    correct = "".join(p[2] for p in parts); discard = "".join(p[3] for p in parts)
    ids = set(r for r, _ in reads)
    def first_records(path):
        out, keep = [], False
        for line in open(path):
            if line.startswith(">"):
                keep = line[1:].strip() in ids
            if keep:
                out.append(line)
        return "".join(out)
    assert correct == first_records(os.path.join(golden, "tiny.dp.correct.fa"))      # the reference's records, in input order
    assert discard == first_records(os.path.join(golden, "tiny.dp.discard.fa"))
    assert float(t) == 2.0
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
    from conftest import ORACLE, _ensure_oracle
    _ensure_oracle()
    script.write_text(_WORKER.format(root=ROOT, port=port, tmp=str(tmp_path), oracle=ORACLE))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ok" in outs[0][0]


def test_parity_digests_agree_between_results_files_and_goldens(tmp_path):
    """parity.py: the digest of a read computed from a batch result (pieces + counters), from output files, and by the golden
    generator are the same 8 bytes; the committed whole-output goldens load and describe every read of their workload."""
    import importlib.util
    import numpy as np
    from longreadselfcorrect_b200 import api, parity
    reads = [b"ACGTACGTAC", b"GGGTTTAAAC", b"TTTT"]
    corrected = [b"ACGTTACGTAC", None, b"TTTTA"]           # read 1 is discarded
    letters = np.frombuffer(b"".join(reads), dtype=np.uint8)
    off = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.uint64)
    out = np.frombuffer(b"".join(c for c in corrected if c), dtype=np.uint8)
    poff = np.array([0, 11, 16], dtype=np.uint64)
    first = np.array([0, 1, 1, 2], dtype=np.uint64)
    stats = np.zeros(3, dtype=api.STATS_DTYPE)
    stats["merge"] = [1, 0, 1]
    dig = parity.result_digests(out, poff, first, stats, letters, off, first_read_id=5)
    with open(tmp_path / "correct.fa", "wb") as f:
        f.write(b">r5\n" + corrected[0] + b"\n>r7\n" + corrected[2] + b"\n")
    with open(tmp_path / "discard.fa", "wb") as f:
        f.write(b">r6\n" + reads[1] + b"\n")
    files = parity.fasta_digests(str(tmp_path / "correct.fa"), str(tmp_path / "discard.fa"))
    spec = importlib.util.spec_from_file_location("mfg", os.path.join(ROOT, "tests", "golden", "make_full_golden.py"))
    mfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mfg)
    for i, rid in enumerate((5, 6, 7)):
        assert dig[i].tobytes() == files[rid]
        seq = corrected[i] if corrected[i] else reads[i]
        assert files[rid] == mfg.record_digest(b"r%d" % rid, seq, corrected[i] is not None)
    assert dig[1, 7] == 0 and dig[0, 7] == 1
    assert len(parity.output_sha256(dig)) == 64
    for wl, n in (("cfg2", 28935), ("mini", 1015)):
        ids, d, meta = parity.load_golden(wl)
        assert ids.size == n == meta["reads_total"] and d.shape == (n, 8) and meta["sample"] == 1.0
        assert "stride pbcorrect" in meta["command"]
    # a perfect result compares clean, one flipped byte is found
    ids, d, meta = parity.load_golden("mini")
    assert parity.compare_with_golden("mini", d.copy())["identical"]
    bad = d.copy(); bad[17, 3] ^= 1
    r = parity.compare_with_golden("mini", bad)
    assert not r["identical"] and r["mismatches"] == 1 and r["first_mismatch"] == 17
