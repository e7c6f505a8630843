"""GPU parity tests: every call goes through the C ABI of libpbsc.so (ctypes) and is compared with
(a) golden vectors the reference binary produced and (b) the CPU oracle on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import read_fasta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from longreadselfcorrect_b200 import api as a
    assert a.device_count() > 0, "no CUDA device: the hot path has no CPU fallback"
    return a


@pytest.fixture(scope="module")
def tiny_index(api, golden):
    idx = api.Index.load(os.path.join(golden, "tiny"))
    yield idx
    idx.close()


@pytest.fixture(scope="module")
def tiny_reads(golden):
    return read_fasta(os.path.join(golden, "tiny.reads.fa"))


def _golden_intervals(golden, which):
    return np.loadtxt(os.path.join(golden, f"tiny.fm_{which}.txt"), dtype=np.int64).reshape(-1, 2)


def test_index_symbols_roundtrip(api, tiny_index, golden):
    # decode the on-disk runs independently and compare with what the device table holds
    for which, ext in ((api.PBSC_BWT, "bwt"), (api.PBSC_RBWT, "rbwt")):
        raw = open(os.path.join(golden, f"tiny.{ext}"), "rb").read()
        runs = np.frombuffer(raw[30:], dtype=np.uint8)
        sym = np.repeat(np.frombuffer(b"$ACGT", dtype=np.uint8)[runs >> 5], runs & 31)
        n = tiny_index.num_symbols(which)
        assert n == sym.size
        assert tiny_index.symbols(which, 0, n) == sym.tobytes()


@pytest.mark.parametrize("k0", [0, 4, 11])
def test_findinterval_matches_reference(api, tiny_index, golden, k0):
    tiny_index.build_prefix_table(k0)
    qs = [l.strip() for l in open(os.path.join(golden, "tiny.fm_queries.txt")) if l.strip()]
    for which, name in ((api.PBSC_BWT, "bwt"), (api.PBSC_RBWT, "rbwt")):
        lo, hi, _ = tiny_index.find_interval(which, qs)
        ref = _golden_intervals(golden, name)
        valid = ref[:, 0] <= ref[:, 1]
        # valid intervals must agree exactly; empty ones must be empty (the reference's raw values at the
        # early break are never consumed on this path)
        assert np.array_equal(lo[valid], ref[valid, 0]) and np.array_equal(hi[valid], ref[valid, 1])
        assert np.all(lo[~valid] > hi[~valid])
        # without the prefix table even the raw values at the early break are the reference's
        if k0 == 0:
            assert np.array_equal(lo, ref[:, 0]) and np.array_equal(hi, ref[:, 1])
    tiny_index.build_prefix_table(0)


def _params(api, name):
    if name == "tiny":
        return api.Params.make(coverage=30, genome=5, no_dp=True)
    return api.Params.make(coverage=100, genome=10, no_dp=True)


@pytest.mark.parametrize("name", ["tiny", "tiny100"])
def test_threshold_table(api, golden, name):
    assert _params(api, name).threshold_table_text() == open(os.path.join(golden, f"{name}.threshold-table")).read()


@pytest.mark.parametrize("name,fixture", [("tiny", "oracle_tiny"), ("tiny100", "oracle_tiny100")])
def test_seeds_match_reference_and_oracle(api, tiny_index, tiny_reads, golden, name, fixture, request):
    p = _params(api, name)
    seeds, off = tiny_index.search_seeds(p, [s for _, s in tiny_reads])
    # (a) golden seed dumps of the reference: seedStr, maxFixedMerFreq, seedStartPos, isRepeat
    lines = []
    for r, (rid, seq) in enumerate(tiny_reads):
        lines.append(f"#{rid}\n")
        if len(seq) >= p.c.start_kmer:
            for s in seeds[int(off[r]):int(off[r + 1])]:
                lines.append(f"{seq[s['start']:s['start'] + s['len']]}\t{s['max_fixed_freq']}\t{s['start']}\t{'Yes' if s['is_repeat'] else 'No'}\n")
    assert "".join(lines) == open(os.path.join(golden, f"{name}.seeds.tsv")).read()
    # (b) oracle: also the per-seed best k-mer sizes
    o = request.getfixturevalue(fixture)
    for r, rec in enumerate(o["dump"]):
        got = [[int(s["start"]), int(s["len"]), int(s["max_fixed_freq"]), int(s["is_repeat"]), int(s["start_best_k"]), int(s["end_best_k"])]
               for s in seeds[int(off[r]):int(off[r + 1])]]
        assert got == rec["seeds"], rec["id"]


@pytest.mark.parametrize("name,fixture", [("tiny", "oracle_tiny"), ("tiny100", "oracle_tiny100")])
@pytest.mark.parametrize("k0", [0, 10])
def test_extend_pairs_match_oracle(api, tiny_index, name, fixture, k0, request):
    p = _params(api, name)
    tiny_index.build_prefix_table(k0)
    o = request.getfixturevalue(fixture)
    pairs = [pr for rec in o["dump"] for pr in rec["pairs"]]
    assert len(pairs) > 100
    min_sa = (p.c.pb_coverage // 60) * 3 if p.c.pb_coverage > 60 else 3
    status, merged = tiny_index.extend_overlap(p, [x["src"] for x in pairs], [x["path"] for x in pairs], [x["trg"] for x in pairs],
                                               [x["dis"] for x in pairs], [x["k"] for x in pairs], [min_sa] * len(pairs))
    bad = [(i, int(status[i]), pairs[i]["status"]) for i in range(len(pairs)) if int(status[i]) != pairs[i]["status"]]
    assert not bad, bad[:10]
    wrong = [i for i in range(len(pairs)) if pairs[i]["status"] > 0 and merged[i] != pairs[i]["out"]]
    assert not wrong, wrong[:10]
    tiny_index.build_prefix_table(0)


@pytest.mark.parametrize("name", ["tiny", "tiny100"])
@pytest.mark.parametrize("k0", [0, 12])
def test_correct_reads_byte_identical(api, tiny_index, tiny_reads, golden, name, k0):
    p = _params(api, name)
    tiny_index.build_prefix_table(k0)
    out, poff, first, stats = tiny_index.correct_reads(p, [s for _, s in tiny_reads])
    pieces = api.Index.pieces_as_strings(out, poff, first)
    correct, discard = [], []
    for (rid, seq), pc, st in zip(tiny_reads, pieces, stats):
        if st["merge"]:
            for s in pc:
                correct.append(f">{rid}\n{s}\n")
        else:
            discard.append(f">{rid}\n{seq}\n")
    assert "".join(correct) == open(os.path.join(golden, f"{name}.correct.fa")).read()
    assert "".join(discard) == open(os.path.join(golden, f"{name}.discard.fa")).read()
    # the summary counters of the reference's stdout block
    m = stats[stats["merge"] == 1]
    summ = dict(l.split(":")[0:2] for l in open(os.path.join(golden, f"{name}.summary.txt")) if ":" in l)
    assert int(m["total_reads_len"].sum()) == int(summ["TotalReadsLen"])
    assert int(m["corrected_len"].sum()) == int(summ["CorrectedLen"].split(",")[0])
    assert int(m["total_seed_num"].sum()) == int(summ["TotalSeedNum"])
    assert int(m["total_walk_num"].sum()) == int(summ["TotalWalkNum"])
    assert int(m["fm_num"].sum()) == int(summ["FMNum"].split(",")[0])
    assert int(m["high_error_num"].sum()) == int(summ["HighErrorNum"].split(",")[0])
    assert int(m["exceed_depth_num"].sum()) == int(summ["ExceedDepthNum"].split(",")[0])
    assert int(m["exceed_leave_num"].sum()) == int(summ["ExceedLeaveNum"].split(",")[0])
    tiny_index.build_prefix_table(0)


def test_edge_cases(api, tiny_index, tiny_reads):
    p = _params(api, "tiny")
    # empty batch, reads shorter than the static k-mer, a read with a single seed-able stretch
    out, poff, first, stats = tiny_index.correct_reads(p, [])
    assert first.tolist() == [0]
    reads = ["ACGT", "A" * 16, tiny_reads[0][1][:40], tiny_reads[1][1]]
    out, poff, first, stats = tiny_index.correct_reads(p, reads)
    assert [int(s["merge"]) for s in stats][:2] == [0, 0]
    assert int(stats[3]["total_reads_len"]) == len(reads[3])
    with pytest.raises(api.PbscError):
        tiny_index.correct_reads(p, ["ACGTN" * 10])


def test_synthetic_index_matches_oracle(api, oracle_bin, tmp_path):
    """The device-generated synthetic BWT (microbenchmark input) searched by the oracle through the on-disk format."""
    idx = api.Index.synthetic(200_000, 300, 5)
    rng = np.random.Generator(np.random.PCG64(9))
    qs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(rng.integers(2, 14)))) for _ in range(2000)]
    (tmp_path / "q.txt").write_text("\n".join(qs) + "\n")
    rank = {ord("$"): 0, ord("A"): 1, ord("C"): 2, ord("G"): 3, ord("T"): 4}
    for which, ext in ((api.PBSC_BWT, "bwt"), (api.PBSC_RBWT, "rbwt")):
        n = idx.num_symbols(which)
        sym = np.frombuffer(idx.symbols(which, 0, n), dtype=np.uint8)
        assert int((sym == ord("$")).sum()) == idx.num_strings(which)
        r = np.vectorize(rank.get)(sym).astype(np.uint8)
        # run-length encode (runs of at most 31, RLUnit.h)
        change = np.flatnonzero(np.diff(r)) + 1
        starts = np.concatenate(([0], change))
        lens = np.diff(np.concatenate((starts, [n])))
        runs = bytearray()
        for s, l in zip(starts, lens):
            while l > 0:
                c = min(int(l), 31)
                runs.append((int(r[s]) << 5) | c)
                l -= c
        import struct
        path = tmp_path / f"syn.{ext}"
        path.write_bytes(struct.pack("<HQQQi", 0xCACA, idx.num_strings(which), n, len(runs), 0) + bytes(runs))
        ref = subprocess.run([oracle_bin, "findinterval", str(path), str(tmp_path / "q.txt")], check=True, stdout=subprocess.PIPE, text=True).stdout
        ref = np.array([list(map(int, l.split())) for l in ref.splitlines()], dtype=np.int64)
        lo, hi, _ = idx.find_interval(which, qs)
        assert np.array_equal(lo, ref[:, 0]) and np.array_equal(hi, ref[:, 1])
    idx.close()


def _write_records(pieces, stats, reads, split=False):
    correct, discard = [], []
    for (rid, seq), pc, st in zip(reads, pieces, stats):
        if st["merge"]:
            for j, s in enumerate(pc):
                correct.append(f">{rid}{'_%d' % j if split else ''}\n{s}\n")
        else:
            discard.append(f">{rid}\n{seq}\n")
    return "".join(correct), "".join(discard)


@pytest.mark.parametrize("engine", ["thread", "warp"])
def test_both_engines_byte_identical(api, tiny_index, tiny_reads, golden, engine, monkeypatch):
    monkeypatch.setenv("PBSC_ENGINE", engine)
    p = _params(api, "tiny")
    tiny_index.build_prefix_table(13)
    out, poff, first, stats = tiny_index.correct_reads(p, [s for _, s in tiny_reads])
    c, d = _write_records(api.Index.pieces_as_strings(out, poff, first), stats, tiny_reads)
    assert c == open(os.path.join(golden, "tiny.correct.fa")).read()
    assert d == open(os.path.join(golden, "tiny.discard.fa")).read()
    tiny_index.build_prefix_table(0)


def test_cli_drop_in(golden, tmp_path):
    """The host binary end to end: same files in, byte-identical files out."""
    from conftest import ROOT
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    out = tmp_path / "out"
    r = subprocess.run([exe, "pbcorrect", "-t", "4", "-p", os.path.join(golden, "tiny"), "-o", str(out), "-c", "30", "-g", "5", "--nodp",
                        "--batch-mbp", "0.2", os.path.join(golden, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for f in ("correct.fa", "discard.fa", "threshold-table"):
        assert open(out / f, "rb").read() == open(os.path.join(golden, f"tiny.{f}"), "rb").read(), f
    want = [l for l in open(os.path.join(golden, "tiny.summary.txt")).read().strip().splitlines() if l]
    got = [l for l in r.stdout.strip().splitlines() if l and not l.startswith("Time of")]
    assert got == want
    # gz input and a FASTA whose last line lacks the trailing newline (the reference drops that read's last line)
    import gzip
    raw = open(os.path.join(golden, "tiny.reads.fa"), "rb").read()
    with gzip.open(tmp_path / "r.fa.gz", "wb") as f:
        f.write(raw)
    out2 = tmp_path / "out2"
    r = subprocess.run([exe, "-p", os.path.join(golden, "tiny"), "-o", str(out2), "-c", "30", "-g", "5", "--nodp", str(tmp_path / "r.fa.gz")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out2 / "correct.fa", "rb").read() == open(os.path.join(golden, "tiny.correct.fa"), "rb").read()


@pytest.mark.parametrize("opts,kw", [
    (["-c", "45", "-g", "5", "--nodp"], dict(coverage=45, genome=5)),
    (["-c", "70", "-g", "10", "--nodp", "--split"], dict(coverage=70, genome=10, split=True)),
    (["-c", "45", "-g", "5", "--nodp", "-n", "3", "-m", "2"], dict(coverage=45, genome=5, next_target=3, mode=2)),
    (["-c", "45", "-g", "100", "--nodp", "-l", "8", "-e", "0.2"], dict(coverage=45, genome=100, max_leaves=8, error_rate=0.2)),
])
@pytest.mark.parametrize("engine", ["thread", "warp"])
def test_repeat_rich_dataset_vs_oracle(api, oracle_bin, tmp_path, opts, kw, engine, monkeypatch):
    """Fresh data with repeat families and tandem arrays (the cases where equal idmers, strand swaps and look-aheads
    occur), index built by the torch suffix sorter, oracle as the checker."""
    from conftest import run_oracle
    from longreadselfcorrect_b200 import bwt_build, synth
    monkeypatch.setenv("PBSC_ENGINE", engine)
    g = synth.make_genome(15000, 21, repeat_families=1, tandem_arrays=3)
    codes, off = synth.simulate_reads(g, 45, 1200, 211, min_len=300)
    reads = synth.read_strings(codes, off)
    fa = str(tmp_path / "reads.fa")
    synth.write_fasta(fa, codes, off)
    prefix = str(tmp_path / "idx")
    bwt_build.build_index_files(prefix, codes, off)
    run_oracle(oracle_bin, prefix, fa, str(tmp_path / "or"), opts)
    idx = api.Index.load(prefix)
    idx.build_prefix_table(13)
    p = api.Params.make(no_dp=True, **kw)
    out, poff, first, stats = idx.correct_reads(p, reads)
    named = [(f"r{i}", s) for i, s in enumerate(reads)]
    c, d = _write_records(api.Index.pieces_as_strings(out, poff, first), stats, named, split=kw.get("split", False))
    assert c == open(tmp_path / "or" / "correct.fa").read()
    assert d == open(tmp_path / "or" / "discard.fa").read()
    idx.close()


# ---- DP / multiple-alignment fallback (default options, correctByMSAlignment) ------------------------------------------

def _select_dp_kernel(monkeypatch, dp_kernel):
    # the multiple alignment has two kernels as well: thread per pile-up (default for small ones) and warp per pile-up;
    # "thread" runs forces every pile-up through the former, the others through the latter
    monkeypatch.setenv("PBSC_MSA_WARP_MIN", "-1" if dp_kernel == "thread" else "0")
    if dp_kernel == "thread":
        monkeypatch.setenv("PBSC_DP_THREAD", "1")
        monkeypatch.setenv("PBSC_DPT_MIN_ROWS", "1")
    elif dp_kernel == "warp":
        monkeypatch.setenv("PBSC_DP_THREAD", "0")


def _summary_counters(stats):
    m = stats[stats["merge"] == 1]
    return {k: int(m[k].sum()) for k in ("total_reads_len", "corrected_len", "total_seed_num", "total_walk_num", "fm_num", "dp_num",
                                         "high_error_num", "exceed_depth_num", "exceed_leave_num", "seed_dis")}


@pytest.mark.parametrize("name,kw", [("tiny", dict(coverage=30, genome=5)), ("tiny100", dict(coverage=100, genome=10))])
@pytest.mark.parametrize("k0", [0, 13])
@pytest.mark.parametrize("dp_kernel", ["thread", "warp", "auto"])
def test_dp_fallback_byte_identical_to_reference(api, tiny_index, tiny_reads, golden, name, kw, k0, dp_kernel, monkeypatch):
    """Default options: failed walks go through retrieveStr / extendMatch / MultipleAlignment on the GPU; fixtures written by
    the reference binary (tests/golden/make_golden.py dp).  Both alignment kernels: one alignment per thread (forced for every
    pass, however few rows it has), warp per row alone, and the default mix (a pass with few rows is left to the warp kernel)."""
    _select_dp_kernel(monkeypatch, dp_kernel)
    p = api.Params.make(no_dp=False, **kw)
    tiny_index.build_prefix_table(k0)
    out, poff, first, stats = tiny_index.correct_reads(p, [s for _, s in tiny_reads])
    c, d = _write_records(api.Index.pieces_as_strings(out, poff, first), stats, tiny_reads)
    assert c == open(os.path.join(golden, f"{name}.dp.correct.fa")).read()
    assert d == open(os.path.join(golden, f"{name}.dp.discard.fa")).read()
    summ = dict(l.split(":")[0:2] for l in open(os.path.join(golden, f"{name}.dp.summary.txt")) if ":" in l)
    got = _summary_counters(stats)
    assert got["dp_num"] == int(summ["DPNum"].split(",")[0]) and got["dp_num"] > 0
    assert got["fm_num"] == int(summ["FMNum"].split(",")[0])
    assert got["corrected_len"] == int(summ["CorrectedLen"].split(",")[0])
    assert got["total_walk_num"] == int(summ["TotalWalkNum"])
    assert got["seed_dis"] // got["total_walk_num"] == int(summ["DisBetweenSeeds"])
    tm = api.last_timing()
    assert tm["dp_jobs"] > 0
    if dp_kernel != "auto":
        assert (tm["dp_thread_rows"] > 0) == (dp_kernel == "thread")
    assert tm["dp_thread_rows"] <= tm["dp_rows"]
    tiny_index.build_prefix_table(0)


@pytest.mark.parametrize("cov,opts,kw", [
    (45, ["-c", "45", "-g", "5"], dict(coverage=45, genome=5)),
    (45, ["-c", "70", "-g", "10", "--split", "-n", "2"], dict(coverage=70, genome=10, split=True, next_target=2)),
    (9, ["-c", "30", "-g", "5"], dict(coverage=30, genome=5)),
    (9, ["-c", "30", "-g", "5", "--split"], dict(coverage=30, genome=5, split=True)),
])
@pytest.mark.parametrize("dp_kernel", ["thread", "warp"])
def test_dp_fallback_repeat_rich_vs_oracle(api, oracle_bin, tmp_path, cov, opts, kw, dp_kernel, monkeypatch):
    """Repeat-rich data; the 9x set makes the fallback itself fail (three or fewer rows) so --split matters; a tiny chunk
    budget forces the chunked path."""
    from conftest import run_oracle
    from longreadselfcorrect_b200 import bwt_build, synth
    monkeypatch.setenv("PBSC_DP_CHUNK_MB", "8")
    _select_dp_kernel(monkeypatch, dp_kernel)
    g = synth.make_genome(15000, 23, repeat_families=1, tandem_arrays=3)
    codes, off = synth.simulate_reads(g, cov, 1200, 213, min_len=300)
    reads = synth.read_strings(codes, off)
    fa = str(tmp_path / "reads.fa")
    synth.write_fasta(fa, codes, off)
    prefix = str(tmp_path / "idx")
    bwt_build.build_index_files(prefix, codes, off)
    o = run_oracle(oracle_bin, prefix, fa, str(tmp_path / "or"), opts, threads=8)
    idx = api.Index.load(prefix)
    idx.build_prefix_table(13)
    p = api.Params.make(no_dp=False, **kw)
    out, poff, first, stats = idx.correct_reads(p, reads)
    named = [(f"r{i}", s) for i, s in enumerate(reads)]
    c, d = _write_records(api.Index.pieces_as_strings(out, poff, first), stats, named, split=kw.get("split", False))
    assert c == open(tmp_path / "or" / "correct.fa").read()
    assert d == open(tmp_path / "or" / "discard.fa").read()
    summ = dict(l.split(":")[0:2] for l in o.stdout.splitlines() if ":" in l)
    got = _summary_counters(stats)
    assert got["dp_num"] == int(summ["DPNum"].split(",")[0])
    assert got["total_walk_num"] - got["fm_num"] - got["dp_num"] == int(summ["OutcastNum"].split(",")[0])
    idx.close()


def test_cli_drop_in_default_options(golden, tmp_path):
    """`pbcorrect` without --nodp: byte-identical correct.fa / discard.fa and the same stdout summary as the reference."""
    from conftest import ROOT
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    out = tmp_path / "out"
    r = subprocess.run([exe, "pbcorrect", "-t", "4", "-p", os.path.join(golden, "tiny"), "-o", str(out), "-c", "30", "-g", "5",
                        "--batch-mbp", "0.2", os.path.join(golden, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for f in ("correct.fa", "discard.fa"):
        assert open(out / f, "rb").read() == open(os.path.join(golden, f"tiny.dp.{f}"), "rb").read(), f
    want = [l for l in open(os.path.join(golden, "tiny.dp.summary.txt")).read().strip().splitlines() if l]
    got = [l for l in r.stdout.strip().splitlines() if l and not l.startswith("Time of")]
    assert got == want


def test_pinned_buffers_and_cached_blocks_give_the_same_bytes(api, tiny_index, tiny_reads, golden):
    """End-to-end entry from page-locked buffers (pbsc_host_alloc), twice, so the second call runs on cached device blocks
    and reused pinned outputs: same bytes as the pageable path and as the reference."""
    p = api.Params.make(coverage=30, genome=5)
    tiny_index.build_prefix_table(13)
    buf, off = api._concat([s for _, s in tiny_reads])
    pin = (api.pinned_copy(buf), api.pinned_copy(off))
    want = open(os.path.join(golden, "tiny.dp.correct.fa")).read()
    for _ in range(2):
        out, poff, first, stats = tiny_index.correct_reads(p, packed=pin, pinned_out=True)
        c, _d = _write_records(api.Index.pieces_as_strings(out, poff, first), stats, tiny_reads)
        assert c == want
    t = api.last_timing()
    assert t["walk_launches"] > 0 and t["walk_ms"] > 0 and t["dp_rows"] > 0
    tiny_index.build_prefix_table(0)


def test_random_sector_roofline_probe(api):
    g = api.random_sector_peak(256 << 20)
    assert 50.0 < g < 8000.0   # GB/s of independent 32-byte reads; a sanity bound, not a performance assertion
