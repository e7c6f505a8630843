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
