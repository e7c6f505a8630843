import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE = os.path.join(ROOT, "oracle", "pbsc_oracle")
REF_STRIDE = os.path.join(ROOT, "oracle", "_ref", "stride")
REF_FMDUMP = os.path.join(ROOT, "oracle", "_ref", "fm_dump")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle():
    if not os.path.exists(ORACLE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "pbsc_oracle"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ORACLE


@pytest.fixture(scope="session")
def oracle_bin():
    return _ensure_oracle()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def read_fasta(path):
    recs, name, seq = [], None, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(seq)))
                name, seq = line[1:].split()[0], []
            elif line:
                seq.append(line)
    if name is not None:
        recs.append((name, "".join(seq)))
    return recs


def run_oracle(oracle_bin, prefix, reads, outdir, opts, dump=None, threads=4):
    cmd = [oracle_bin, "pbcorrect", "--threads", str(threads), "-p", prefix, "-o", outdir] + opts
    if dump:
        cmd += ["--dump", dump]
    cmd.append(reads)
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.fixture(scope="session")
def oracle_tiny(oracle_bin, golden, tmp_path_factory):
    """Oracle run on the tiny fixture with -c 30 -g 5 --nodp --debugseed; returns dict with dir and parsed dump."""
    import json
    d = tmp_path_factory.mktemp("oracle_tiny")
    out = str(d / "out")
    dump = str(d / "dump.jsonl")
    r = run_oracle(oracle_bin, os.path.join(golden, "tiny"), os.path.join(golden, "tiny.reads.fa"), out,
                   ["-c", "30", "-g", "5", "--nodp", "--debugseed"], dump=dump)
    recs = [json.loads(l) for l in open(dump)]
    return {"dir": out, "dump": recs, "stdout": r.stdout}


@pytest.fixture(scope="session")
def oracle_tiny100(oracle_bin, golden, tmp_path_factory):
    import json
    d = tmp_path_factory.mktemp("oracle_tiny100")
    out = str(d / "out")
    dump = str(d / "dump.jsonl")
    r = run_oracle(oracle_bin, os.path.join(golden, "tiny"), os.path.join(golden, "tiny.reads.fa"), out,
                   ["-c", "100", "-g", "10", "--nodp", "--debugseed"], dump=dump)
    recs = [json.loads(l) for l in open(dump)]
    return {"dir": out, "dump": recs, "stdout": r.stdout}


def _oracle_dp_run(oracle_bin, golden, tmp_path_factory, name, opts):
    d = tmp_path_factory.mktemp("oracle_dp_" + name)
    out = str(d / "out")
    r = run_oracle(oracle_bin, os.path.join(golden, "tiny"), os.path.join(golden, "tiny.reads.fa"), out, opts + ["--debugseed"], threads=8)
    return {"dir": out, "stdout": r.stdout}


@pytest.fixture(scope="session")
def oracle_tiny_dp(oracle_bin, golden, tmp_path_factory):
    """Oracle run with DEFAULT options (DP / multiple-alignment fallback on), -c 30 -g 5."""
    return _oracle_dp_run(oracle_bin, golden, tmp_path_factory, "tiny", ["-c", "30", "-g", "5"])


@pytest.fixture(scope="session")
def oracle_tiny100_dp(oracle_bin, golden, tmp_path_factory):
    return _oracle_dp_run(oracle_bin, golden, tmp_path_factory, "tiny100", ["-c", "100", "-g", "10"])
