"""GPU tests of the persisted / relocatable flat index (PREFIX.fmg, clone, blob export/import) and of the lanes (several
batches of one index in flight on separate streams): every variant must give the reference's bytes (tests/golden, written
by the unmodified reference binary)."""
import os
import threading

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
def tiny_reads(golden):
    return read_fasta(os.path.join(golden, "tiny.reads.fa"))


def _records(api, out, poff, first, stats, reads):
    pieces = api.Index.pieces_as_strings(out, poff, first)
    correct, discard = [], []
    for (rid, seq), pc, st in zip(reads, pieces, stats):
        if st["merge"]:
            correct += [f">{rid}\n{s}\n" for s in pc]
        else:
            discard.append(f">{rid}\n{seq}\n")
    return "".join(correct), "".join(discard)


def _check_against_reference(api, idx, reads, golden):
    p = api.Params.make(coverage=30, genome=5)
    out, poff, first, stats = idx.correct_reads(p, [s for _, s in reads])
    c, d = _records(api, out, poff, first, stats, reads)
    assert c == open(os.path.join(golden, "tiny.dp.correct.fa")).read()
    assert d == open(os.path.join(golden, "tiny.dp.discard.fa")).read()


def _intervals(api, idx, golden):
    qs = [l.strip() for l in open(os.path.join(golden, "tiny.fm_queries.txt")) if l.strip()][:400]
    return [idx.find_interval(w, qs)[:2] for w in (api.PBSC_BWT, api.PBSC_RBWT)]


def test_fmg_roundtrip_and_open(api, tiny_reads, golden, tmp_path):
    """save -> load_fmg gives the same tables (symbols, intervals, corrected bytes); pbsc_index_open prefers a matching
    PREFIX.fmg, rebuilds when the prefix-table length differs and rejects damaged files."""
    import shutil
    for ext in ("bwt", "rbwt", "sai"):
        shutil.copy(os.path.join(golden, f"tiny.{ext}"), tmp_path / f"t.{ext}")
    prefix = str(tmp_path / "t")
    idx, used = api.Index.open(prefix, k0=11, write_fmg=True)
    assert not used and os.path.exists(prefix + ".fmg")
    want_iv = _intervals(api, idx, golden)
    n = idx.num_symbols(api.PBSC_BWT)
    want_sym = idx.symbols(api.PBSC_BWT, 0, n)
    idx.close()
    idx2, used = api.Index.open(prefix, k0=11, write_fmg=False)
    assert used, "a matching PREFIX.fmg must be used"
    assert idx2.symbols(api.PBSC_BWT, 0, n) == want_sym
    for (lo, hi), (lo2, hi2) in zip(want_iv, _intervals(api, idx2, golden)):
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    _check_against_reference(api, idx2, tiny_reads, golden)
    # replacing the prefix table of a blob-backed index works and keeps results
    idx2.build_prefix_table(9)
    _check_against_reference(api, idx2, tiny_reads, golden)
    idx2.close()
    # another k0: the file does not match, the run-length files are used
    idx3, used = api.Index.open(prefix, k0=10, write_fmg=False)
    assert not used
    idx3.close()
    # damage: flip a header byte -> checksum fails -> falls back to the .bwt files; a direct load reports the format error
    raw = bytearray(open(prefix + ".fmg", "rb").read())
    raw[40] ^= 0xFF
    open(prefix + ".fmg", "wb").write(raw)
    with pytest.raises(api.PbscError):
        api.Index.load_fmg(prefix + ".fmg")
    idx4, used = api.Index.open(prefix, k0=11, write_fmg=False)
    assert not used
    idx4.close()
    # truncated file
    open(prefix + ".fmg", "wb").write(bytes(raw[: len(raw) // 2]))
    with pytest.raises(api.PbscError):
        api.Index.load_fmg(prefix + ".fmg")


def test_truncated_bwt_is_a_format_error_not_a_crash(api, golden, tmp_path):
    raw = open(os.path.join(golden, "tiny.bwt"), "rb").read()
    for ext in ("rbwt", "sai"):
        open(tmp_path / f"t.{ext}", "wb").write(open(os.path.join(golden, f"tiny.{ext}"), "rb").read())
    open(tmp_path / "t.bwt", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(api.PbscError) as e:
        api.Index.load(str(tmp_path / "t"))
    assert e.value.code == -3
    # a header that promises an absurd number of runs must not be trusted with an allocation
    bad = bytearray(raw)
    bad[18:26] = (1 << 60).to_bytes(8, "little")
    open(tmp_path / "t.bwt", "wb").write(bad)
    with pytest.raises(api.PbscError) as e:
        api.Index.load(str(tmp_path / "t"))
    assert e.value.code == -3


def test_clone_and_blob_import(api, tiny_reads, golden):
    """pbsc_index_clone (peer copies; here onto the same GPU, and onto GPU 1 when the box has one) and export/import of
    the blob through a caller-owned device buffer (what bench.py broadcasts with NCCL)."""
    import torch
    idx = api.Index.load(os.path.join(golden, "tiny"))
    idx.build_prefix_table(10)
    want_iv = _intervals(api, idx, golden)
    targets = [0] + ([1] if api.device_count() > 1 else [])
    for dev in targets:
        c = idx.clone(dev)
        for (lo, hi), (lo2, hi2) in zip(want_iv, _intervals(api, c, golden)):
            assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
        _check_against_reference(api, c, tiny_reads, golden)
        c.close()
    nb = idx.blob_size()
    buf = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
    idx.export_blob(buf.data_ptr(), nb)
    torch.cuda.synchronize()
    imp = api.Index.import_blob(buf.data_ptr(), nb, 0, 0)
    del buf
    _check_against_reference(api, imp, tiny_reads, golden)
    host = torch.empty(nb, dtype=torch.uint8)
    tmp = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
    idx.export_blob(tmp.data_ptr(), nb)
    host.copy_(tmp)
    imp2 = api.Index.import_blob(host.data_ptr(), nb, -1, 0)
    _check_against_reference(api, imp2, tiny_reads, golden)
    imp.close(); imp2.close(); idx.close()


@pytest.mark.parametrize("lanes", [2, 3])
def test_lanes_concurrent_batches_byte_identical(api, tiny_reads, golden, lanes):
    """Batches of one index run at the same time on separate streams and arenas, driven by host threads (the stream pipeline
    that replaces SequenceProcessFramework's worker threads); pieces concatenated in input order equal the reference's."""
    idx = api.Index.load(os.path.join(golden, "tiny"))
    idx.build_prefix_table(12)
    idx.set_lanes(lanes)
    assert idx.lanes() == lanes
    p = api.Params.make(coverage=30, genome=5)
    parts = 6
    bounds = [len(tiny_reads) * i // parts for i in range(parts + 1)]
    results = [None] * parts
    errors = []

    def work(i):
        try:
            sub = tiny_reads[bounds[i]:bounds[i + 1]]
            b = api.Batch(idx, p, reads=[s for _, s in sub])
            b.run()
            results[i] = (sub, b.fetch())
            b.close()
        except Exception as e:   # surfaces in the main thread
            errors.append(e)

    for rep in range(2):
        th = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors
        c, d = "", ""
        for sub, (out, poff, first, stats) in results:
            cc, dd = _records(api, out, poff, first, stats, sub)
            c += cc
            d += dd
        assert c == open(os.path.join(golden, "tiny.dp.correct.fa")).read()
        assert d == open(os.path.join(golden, "tiny.dp.discard.fa")).read()
    # a different -i on a multi-lane index rebuilds the shared idmer table with no batch running
    p7 = api.Params.make(coverage=30, genome=5, idmer_len=7)
    out, poff, first, stats = idx.correct_reads(p7, [s for _, s in tiny_reads[:20]])
    idx.set_lanes(1)
    out1, poff1, first1, stats1 = idx.correct_reads(p7, [s for _, s in tiny_reads[:20]])
    assert out[: int(poff[int(first[-1])])].tobytes() == out1[: int(poff1[int(first1[-1])])].tobytes()
    idx.close()


def _option_goldens(golden):
    import json
    return json.load(open(os.path.join(golden, "tiny.options.json")))


@pytest.mark.parametrize("variant", ["adjust_k17_u2_r-2", "adjust_k21", "idmer7", "minkmer15", "error0.2", "leaves16", "mode2_nodp", "fastq"])
def test_cli_options_byte_identical_to_reference(golden, tmp_path, variant):
    """`pbcorrect` with the options the other fixtures leave at their defaults (-k/-u/-r, -i, -s, -e, -l, -m) and with FASTQ
    input (lower-case bases, quality lines that start with '@' or '>'): the sha256 of correct.fa and discard.fa equal the
    reference's (tests/golden/make_option_golden.py)."""
    import hashlib
    import subprocess
    from conftest import ROOT
    sys_path = os.path.join(ROOT, "tests", "golden")
    g = _option_goldens(golden)[variant]
    reads = os.path.join(golden, "tiny.reads.fa")
    if g["input"] == "fastq":
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_option_golden", os.path.join(sys_path, "make_option_golden.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        reads = str(tmp_path / "tiny.reads.fq")
        m.write_fastq(reads, read_fasta(os.path.join(golden, "tiny.reads.fa")))
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    out = tmp_path / "out"
    r = subprocess.run([exe, "pbcorrect", "-t", "3", "-p", os.path.join(golden, "tiny"), "-o", str(out), "--batch-mbp", "0.15", "--lanes", "2"] + g["options"] + [reads],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert hashlib.sha256(open(out / "correct.fa", "rb").read()).hexdigest() == g["correct_sha256"]
    assert hashlib.sha256(open(out / "discard.fa", "rb").read()).hexdigest() == g["discard_sha256"]
    summ = {l.split(":")[0]: l.split(":")[1].split(",")[0].strip() for l in r.stdout.splitlines() if ":" in l and not l.startswith("Time of")}
    assert summ == g["summary"]


def test_cli_fmg_and_multi_gpu_give_the_same_files(golden, tmp_path):
    """The binary with --write-fmg (cold), then again from PREFIX.fmg (warm), then over every GPU of the box (index cloned by
    peer copies, batches dealt to the GPUs, records written in input order): identical files, equal to the reference's."""
    import shutil
    import subprocess
    from conftest import ROOT
    from longreadselfcorrect_b200 import api as a
    for ext in ("bwt", "rbwt", "sai"):
        shutil.copy(os.path.join(golden, f"tiny.{ext}"), tmp_path / f"t.{ext}")
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    want_c = open(os.path.join(golden, "tiny.dp.correct.fa"), "rb").read()
    want_d = open(os.path.join(golden, "tiny.dp.discard.fa"), "rb").read()
    runs = [("cold", ["--write-fmg", "--gpus", "1"], "PREFIX.bwt"), ("warm", ["--gpus", "1"], "PREFIX.fmg"), ("all", ["--gpus", str(a.device_count())], "PREFIX.fmg")]
    for name, extra, source in runs:
        out = tmp_path / f"out_{name}"
        r = subprocess.run([exe, "-t", "2", "-p", str(tmp_path / "t"), "-o", str(out), "-c", "30", "-g", "5", "--batch-mbp", "0.1"] + extra + [os.path.join(golden, "tiny.reads.fa")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr[-800:]
        assert f"from {source}" in r.stderr, r.stderr[-800:]
        assert open(out / "correct.fa", "rb").read() == want_c and open(out / "discard.fa", "rb").read() == want_d, name
    assert os.path.exists(tmp_path / "t.fmg")


def test_device_and_host_run_length_decoders_agree(api, golden, tmp_path, monkeypatch):
    """The index is decoded from the reference's run bytes on the device (default) or by the sequential host loop
    (PBSC_HOST_DECODE=1): same symbols, same '$' handling, same intervals; malformed run bytes are a format error."""
    res = []
    for host in ("0", "1"):
        monkeypatch.setenv("PBSC_HOST_DECODE", host)
        idx = api.Index.load(os.path.join(golden, "tiny"))
        n = idx.num_symbols(api.PBSC_RBWT)
        res.append((idx.symbols(api.PBSC_BWT, 0, idx.num_symbols(api.PBSC_BWT)), idx.symbols(api.PBSC_RBWT, 0, n), _intervals(api, idx, golden), idx.device_bytes()))
        idx.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] and res[0][3] == res[1][3]
    assert b"$" in res[0][0]
    for (lo, hi), (lo2, hi2) in zip(res[0][2], res[1][2]):
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    monkeypatch.setenv("PBSC_HOST_DECODE", "0")
    raw = bytearray(open(os.path.join(golden, "tiny.bwt"), "rb").read())
    for ext in ("rbwt", "sai"):
        open(tmp_path / f"t.{ext}", "wb").write(open(os.path.join(golden, f"tiny.{ext}"), "rb").read())
    for bad_byte in (0xE3, 0x40):     # symbol rank 7; run of length 0
        b = bytearray(raw)
        b[30 + 1000] = bad_byte
        open(tmp_path / "t.bwt", "wb").write(b)
        with pytest.raises(api.PbscError) as e:
            api.Index.load(str(tmp_path / "t"))
        assert e.value.code == -3


@pytest.mark.parametrize("mode", ["nodp", "default"])
def test_cli_debugseed_dumps_equal_the_reference(golden, tmp_path, mode):
    """`pbcorrect --debugseed`: seed/<id>.seed, seed/error/<id>.seed, extend/<id>.log, .ext and .dp of every read, byte for byte
    what the reference writes (sha256 per file class over all reads, tests/golden/tiny.debugseed.json); the committed
    tiny.seeds.tsv / tiny.ext.tsv fixtures are the same dumps in clear text."""
    import importlib.util
    import json
    import subprocess
    from conftest import ROOT
    g = json.load(open(os.path.join(golden, "tiny.debugseed.json")))[mode]
    spec = importlib.util.spec_from_file_location("mog", os.path.join(ROOT, "tests", "golden", "make_option_golden.py"))
    mog = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mog)
    exe = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")
    out = tmp_path / "out"
    r = subprocess.run([exe, "pbcorrect", "-t", "2", "-p", os.path.join(golden, "tiny"), "-o", str(out), "--batch-mbp", "0.2"] + g["options"]
                       + [os.path.join(golden, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    ids = [rid for rid, _ in read_fasta(os.path.join(golden, "tiny.reads.fa"))]
    got = mog.debug_dump_digest(str(out) + "/", ids)
    for cls in ("seed", "seed_error", "ext", "dp", "log"):
        assert got[cls] == g["dumps"][cls], cls
    if mode == "nodp":
        # clear-text fixtures of round 1
        seeds = "".join(f"#{i}\n" + open(out / "seed" / f"{i}.seed").read() for i in ids)
        assert seeds == open(os.path.join(golden, "tiny.seeds.tsv")).read()
        ext = "".join(f"#{i}\n" + (open(out / "extend" / f"{i}.ext").read() if os.path.exists(out / "extend" / f"{i}.ext") else "") for i in ids)
        assert ext == open(os.path.join(golden, "tiny.ext.tsv")).read()
    name = "tiny" if mode == "nodp" else "tiny.dp"
    assert open(out / "correct.fa").read() == open(os.path.join(golden, f"{name}.correct.fa")).read()
