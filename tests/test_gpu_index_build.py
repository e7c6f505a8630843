"""GPU tests of index construction (`stride index`): the suffix sorter of pbsc_build.cu writes the four files
the unmodified reference writes for the same reads, byte for byte (tests/golden/tiny.* and index_edge.* come from
oracle/_ref/stride index), through the C ABI and through `pbcorrect index`."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read_fasta

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "longreadselfcorrect_b200", "pbcorrect")


@pytest.fixture(scope="module")
def api():
    from longreadselfcorrect_b200 import api as a
    assert a.device_count() > 0, "no CUDA device: the builder has no CPU fallback"
    return a


def _packed(seqs):
    bases = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return bases, off


def _same_files(prefix, name, exts=("bwt", "rbwt", "sai", "rsai")):
    for ext in exts:
        got, want = open(f"{prefix}.{ext}", "rb").read(), open(os.path.join(GOLDEN, f"{name}.{ext}"), "rb").read()
        assert got == want, f"{name}.{ext}: {len(got)} bytes against {len(want)}"


@pytest.mark.parametrize("name,fasta", [("tiny", "tiny.reads.fa"), ("index_edge", "index_edge.fa")])
def test_index_files_equal_the_reference(api, name, fasta, tmp_path):
    """index_edge: exact duplicates, one-base reads, reads that are prefixes of other reads, homopolymer and AT-repeat reads."""
    seqs = [s for _, s in read_fasta(os.path.join(GOLDEN, fasta))]
    api.build_index_files(_packed(seqs), str(tmp_path / "x"))
    _same_files(str(tmp_path / "x"), name)


def test_run_chunks(api, tmp_path, monkeypatch):
    """Inputs of billions of runs are written (pbsc_build.cu) and decoded (pbsc_index.cu) 2^28 runs at a time; PBSC_RUN_CHUNK makes
    the chunks small enough for the fixture to need hundreds of them: same files, same index, same corrected reads."""
    monkeypatch.setenv("PBSC_RUN_CHUNK", "997")
    recs = read_fasta(os.path.join(GOLDEN, "tiny.reads.fa"))
    api.build_index_files(_packed([s for _, s in recs]), str(tmp_path / "c"))
    _same_files(str(tmp_path / "c"), "tiny")
    idx, _ = api.Index.open(str(tmp_path / "c"), k0=11)
    p = api.Params.make(coverage=30, genome=5)
    out, poff, first, stats = idx.correct_reads(p, [s for _, s in recs])
    pieces = api.Index.pieces_as_strings(out, poff, first)
    correct = "".join(f">{rid}\n{s}\n" for (rid, _), pc, st in zip(recs, pieces, stats) if st["merge"] for s in pc)
    assert correct == open(os.path.join(GOLDEN, "tiny.dp.correct.fa")).read()
    idx.close()


def test_lower_case_and_strand_flags(api, tmp_path):
    seqs = [s.lower() for _, s in read_fasta(os.path.join(GOLDEN, "index_edge.fa"))]
    api.build_index_files(_packed(seqs), str(tmp_path / "f"), reverse=False)
    _same_files(str(tmp_path / "f"), "index_edge", ("bwt", "sai"))
    assert not os.path.exists(tmp_path / "f.rbwt") and not os.path.exists(tmp_path / "f.rsai")
    api.build_index_files(_packed(seqs), str(tmp_path / "r"), forward=False)
    _same_files(str(tmp_path / "r"), "index_edge", ("rbwt", "rsai"))
    assert not os.path.exists(tmp_path / "r.bwt")


def test_letters_other_than_acgt_are_rejected(api, tmp_path):
    with pytest.raises(api.PbscError) as e:
        api.build_index_files(_packed(["ACGTNACGT", "ACGT"]), str(tmp_path / "n"))
    assert "ACGT" in str(e.value)
    with pytest.raises(api.PbscError):
        api.build_index_files((np.zeros(0, np.uint8), np.zeros(1, np.uint64)), str(tmp_path / "e"))


def test_in_memory_strand_feeds_the_index(api):
    """pbsc_build_bwt -> pbsc_index_create: the corrected reads are the reference's (tests/golden/tiny.dp.*)."""
    recs = read_fasta(os.path.join(GOLDEN, "tiny.reads.fa"))
    packed = _packed([s for _, s in recs])
    fr, fn, flex = api.build_bwt(packed, reverse=False)
    rr, rn, rlex = api.build_bwt(packed, reverse=True)
    assert fr.tobytes() == open(os.path.join(GOLDEN, "tiny.bwt"), "rb").read()[30:]
    assert rr.tobytes() == open(os.path.join(GOLDEN, "tiny.rbwt"), "rb").read()[30:]
    want = [int(l.split()[0]) for l in open(os.path.join(GOLDEN, "tiny.sai")).read().split("\n")[3:] if l]
    assert flex.tolist() == want
    idx = api.Index.from_runs(fr, fn, len(recs), rr, rn, len(recs))
    idx.build_prefix_table(11)
    p = api.Params.make(coverage=30, genome=5)
    out, poff, first, stats = idx.correct_reads(p, [s for _, s in recs])
    pieces = api.Index.pieces_as_strings(out, poff, first)
    correct = "".join(f">{rid}\n{s}\n" for (rid, _), pc, st in zip(recs, pieces, stats) if st["merge"] for s in pc)
    assert correct == open(os.path.join(GOLDEN, "tiny.dp.correct.fa")).read()
    idx.close()


def test_larger_input_equals_the_torch_builder(api):
    """3 Mbp of simulated reads (the `mini` workload): several buckets of a few hundred thousand suffixes, four rounds deep;
    the run-length bytes equal the ones of bwt_build (itself checked against the reference's index on the CPU)."""
    import bench
    from longreadselfcorrect_b200 import bwt_build
    wl = bench.WORKLOADS["mini"]
    codes, off = bench.make_data(wl)
    packed = bench.packed_ascii(codes, off)
    for rev in (False, True):
        runs, nsym, lex = api.build_bwt(packed, reverse=rev)
        want = bwt_build.run_length_bytes(bwt_build.bwt_symbols(codes, off, reverse=rev, device="cuda:0"))
        assert nsym == codes.size + off.size - 1
        assert np.array_equal(runs, np.asarray(want)), f"reverse={rev}"
        assert sorted(lex.tolist()) == list(range(off.size - 1))


def test_cli_index_subcommand(tmp_path):
    """`pbcorrect index` = `stride index`: default prefix from the reads file's name, -p, --no-reverse, gzip input."""
    import gzip
    import shutil
    fa = tmp_path / "reads.fa"
    shutil.copy(os.path.join(GOLDEN, "tiny.reads.fa"), fa)
    r = subprocess.run([EXE, "index", "-t", "4", "reads.fa"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "done bwt construction" in r.stdout and "Build FM index" in r.stderr
    _same_files(str(tmp_path / "reads"), "tiny")
    with open(fa, "rb") as f, gzip.open(tmp_path / "z.fa.gz", "wb") as g:
        g.write(f.read())
    r = subprocess.run([EXE, "index", "-p", str(tmp_path / "sub" / "pre"), "--no-reverse", str(tmp_path / "z.fa.gz")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    _same_files(str(tmp_path / "sub" / "pre"), "tiny", ("bwt", "sai"))
    assert not os.path.exists(tmp_path / "sub" / "pre.rbwt")
    r = subprocess.run([EXE, "index"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "missing arguments" in r.stderr
    # the files feed `pbcorrect` itself: corrected reads = the reference's
    out = tmp_path / "out"
    r = subprocess.run([EXE, "-p", str(tmp_path / "reads"), "-o", str(out), "-c", "30", "-g", "5", str(fa)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out / "correct.fa").read() == open(os.path.join(GOLDEN, "tiny.dp.correct.fa")).read()


def test_cli_onlyseed(tmp_path):
    """`pbcorrect --onlyseed -b BARCODE` (PacBioSelfCorrectionProcess.cpp:58-62,265-287,315-335; PacBio/BCode.cpp): seeds from the GPU,
    validated on the host; DIR/total.seed and stdout equal the reference's, the seed dumps are written, nothing is corrected."""
    out = tmp_path / "o"
    r = subprocess.run([EXE, "-p", os.path.join(GOLDEN, "tiny"), "-o", str(out), "-c", "30", "-g", "5", "--onlyseed", "-b", os.path.join(GOLDEN, "tiny.barcode.txt"),
                        os.path.join(GOLDEN, "tiny.reads.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "Loading BARCODE" in r.stderr
    assert r.stdout == open(os.path.join(GOLDEN, "tiny.onlyseed.stdout")).read()
    assert open(out / "total.seed").read() == open(os.path.join(GOLDEN, "tiny.onlyseed.total.seed")).read()
    assert sorted(os.listdir(out)) == ["extend", "seed", "threshold-table", "total.seed"]
    assert len([f for f in os.listdir(out / "seed") if f.endswith(".seed")]) == 396
    assert not [f for f in os.listdir(out / "extend") if not f.endswith(".log")]


def test_cli_kmercheck(tmp_path):
    """`pbcorrect kmercheck` = `stride kmercheck` (StriDe/kmercheck.cpp, PacBio/KmerCheckProcess.cpp): k-mer frequencies by batched
    backward search on the GPU, barcode arithmetic and five-number summaries on the host; DIR/total.box and DIR/value.box equal the
    reference's, and a second run appends like the reference does."""
    out = tmp_path / "kc"
    cmd = [EXE, "kmercheck", "-t", "2", "-c", "30", "-p", os.path.join(GOLDEN, "tiny"), "-o", str(out), "-b", os.path.join(GOLDEN, "tiny.barcode.txt"), "-l", "15", "-u", "23",
           "-s", "4", os.path.join(GOLDEN, "tiny.reads.fa")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "Using kmer size : 15 - 23 (4)" in r.stderr
    want_t, want_v = (open(os.path.join(GOLDEN, f"tiny.kmercheck.{n}.box")).read() for n in ("total", "value"))
    assert open(out / "total.box").read() == want_t and open(out / "value.box").read() == want_v
    assert subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode == 0
    assert open(out / "total.box").read() == want_t * 2
    r = subprocess.run([EXE, "kmercheck", "-p", "x", "-o", str(tmp_path / "e"), "-l", "5", "reads.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no barcode" in r.stderr and "invalid range of kmer size:5 - 35" in r.stderr
