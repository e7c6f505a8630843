"""Per-read digests of a correction result, for whole-output parity checks at BASELINE sizes.

The reference writes one FASTA record per read: `>id\\nSEQ\\n` into correct.fa (corrected) or discard.fa (fewer than two
seeds; PacBio/PacBioSelfCorrectionProcess.cpp:313-370, Util/Util.h:57-61).  `tests/golden/make_full_golden.py` stores, for
every read of a workload, the first 7 bytes of the sha256 of that record plus one byte that says which file it went to,
computed from the output of the UNMODIFIED reference.  The functions here compute the same 8 bytes from a GPU result, so a
whole timed run can be compared read by read, and `output_sha256` (sha256 over the digests of all reads in input order)
identifies a whole output independently of how the reads were sharded over GPUs.

This module only hashes bytes; it never computes a correction (test / bench infrastructure)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def record_digest(rid: bytes, seq: bytes, in_correct: bool) -> bytes:
    return hashlib.sha256(b">" + rid + b"\n" + seq + b"\n").digest()[:7] + (b"\x01" if in_correct else b"\x00")


def result_digests(out, poff, first, stats, reads_ascii, read_off, first_read_id: int = 0, prefix: bytes = b"r") -> np.ndarray:
    """uint8 [n, 8]: digest of the record the reference would write for each read of a batch result.

    out/poff/first/stats: what pbsc_correct_batch / pbsc_batch_fetch returned (no --split: one piece per corrected read);
    reads_ascii/read_off: the batch's input reads (ASCII, offsets) -- a discarded read is written as it came in;
    first_read_id: id of the batch's first read in the whole read set (ids are `r<index>` as bench.py names them)."""
    n = len(stats)
    res = np.empty((n, 8), dtype=np.uint8)
    o = memoryview(np.ascontiguousarray(out))
    rd = memoryview(np.ascontiguousarray(reads_ascii))
    merge = stats["merge"]
    for r in range(n):
        rid = prefix + str(first_read_id + r).encode()
        if merge[r]:
            j = int(first[r])
            seq = o[int(poff[j]):int(poff[j + 1])]
        else:
            seq = rd[int(read_off[r]):int(read_off[r + 1])]
        h = hashlib.sha256()
        h.update(b">" + rid + b"\n")
        h.update(seq)
        h.update(b"\n")
        d = h.digest()
        res[r, :7] = np.frombuffer(d[:7], dtype=np.uint8)
        res[r, 7] = 1 if merge[r] else 0
    return res


def fasta_digests(correct_path: str, discard_path: str) -> dict:
    """{read index: 8-byte digest} from a pair of output files whose ids are `r<index>`."""
    dig = {}
    for fn, flag in ((correct_path, True), (discard_path, False)):
        name = None
        with open(fn, "rb") as f:
            for line in f:
                if line.startswith(b">"):
                    name = line[1:].strip()
                else:
                    dig[int(name[1:])] = record_digest(name, line.rstrip(b"\n"), flag)
    return dig


def output_sha256(digests: np.ndarray) -> str:
    """sha256 over the 8-byte digests of all reads in input order: identical for identical outputs, whatever the sharding."""
    return hashlib.sha256(np.ascontiguousarray(digests, dtype=np.uint8).tobytes()).hexdigest()


def load_golden(workload: str):
    """(ids int32[m], digest uint8[m, 8], meta dict) of tests/golden/<workload>.read_sha.npz, or None when absent."""
    path = os.path.join(GOLDEN_DIR, f"{workload}.read_sha.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    return z["ids"], z["digest"], json.loads(str(z["meta"]))


def compare_with_golden(workload: str, digests: np.ndarray, first_read_id: int = 0) -> dict | None:
    """Compare the digests of reads [first_read_id, first_read_id + len) with the committed reference digests."""
    g = load_golden(workload)
    if g is None:
        return None
    ids, want, meta = g
    sel = (ids >= first_read_id) & (ids < first_read_id + len(digests))
    got = digests[ids[sel] - first_read_id]
    bad = np.any(got != want[sel], axis=1)
    return {"golden": f"tests/golden/{workload}.read_sha.npz", "reads_in_golden": int(ids.size), "reads_compared": int(sel.sum()),
            "mismatches": int(bad.sum()), "identical": bool(sel.sum() > 0 and not bad.any()),
            "first_mismatch": (int(ids[sel][np.argmax(bad)]) if bad.any() else None),
            "golden_is": ("every read" if meta.get("sample", 1.0) >= 1.0 else f"seeded {meta['sample']:.0%} sample spread over the whole read set")
                         + " of the workload, written by the unmodified reference"}
