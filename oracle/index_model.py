#!/usr/bin/env python3
"""CPU model (numpy) of the suffix sorter in longreadselfcorrect_b200/csrc/pbsc_build.cu — TEST INFRASTRUCTURE, like the rest of
oracle/: only tests/ may import it.  It follows the CUDA code step by step (same buckets, same 21-symbol keys, same round
logic), so a disagreement between the GPU builder and the reference's files can be narrowed down on the CPU.

What is being restated is `stride index` (StriDe/index.cpp:86-214): BWTCA::runRopebwt2 (SuffixTools/BWTCARopebwt.cpp:160-247)
inserts the reads in input order (MR_SO_IO), i.e. the BWT of the collection with one sentinel per read, sentinels ordered by
read index; BWTWriterBinary (SuffixTools/BWTWriterBinary.cpp:28-94) writes it as run-length units of at most 31 symbols;
SampledSuffixArray::buildLexicoIndex / writeLexicoIndex (SuffixTools/SampledSuffixArray.cpp:158-190,248-258) writes, for the
r-th '$' of the BWT, the index of the read whose first base that row precedes.

    python oracle/index_model.py reads.fa prefix      # writes prefix.bwt .rbwt .sai .rsai
"""
from __future__ import annotations

import struct
import sys

import numpy as np

SYM_PER_KEY = 21          # 3 bits per symbol: $ = 0, A..T = 1..4
BUCKET_SYMS = 2


def make_text(codes: np.ndarray, offsets: np.ndarray, reverse: bool):
    """text[N]: the reads (each reversed for the .rbwt) with a 0 after each; dollar[n]: positions of the zeros."""
    n = offsets.size - 1
    lens = np.diff(offsets).astype(np.int64)
    total = int(offsets[-1])
    rid = np.repeat(np.arange(n, dtype=np.int64), lens)
    pos = np.arange(total, dtype=np.int64)
    src = (offsets[:-1][rid] + (offsets[1:][rid] - 1 - pos)) if reverse else pos
    text = np.zeros(total + n, dtype=np.uint8)
    text[pos + rid] = codes[src].astype(np.uint8) + 1
    dollar = offsets[1:].astype(np.int64) + np.arange(n, dtype=np.int64)
    return text, dollar


def window_key(text: np.ndarray, p: np.ndarray, off: int) -> np.ndarray:
    """21 symbols from p + off on, 3 bits each, first symbol in the top bits; everything after the first 0 reads as 0."""
    N = text.size
    key = np.zeros(p.size, dtype=np.uint64)
    alive = np.ones(p.size, dtype=bool)
    for j in range(SYM_PER_KEY):
        q = p + off + j
        s = np.where(alive & (q < N), text[np.minimum(q, N - 1)], 0).astype(np.uint64)
        alive &= s != 0
        key = (key << np.uint64(3)) | s
    return key


def suffix_array(text: np.ndarray, dollar: np.ndarray) -> np.ndarray:
    N = text.size
    sa = np.zeros(N, dtype=np.int64)
    allp = np.arange(N, dtype=np.int64)
    nxt = np.where(text != 0, np.concatenate([text[1:], [0]]), 0)
    code = (text.astype(np.int64) << 3) | nxt          # bucket = first two symbols (the second reads 0 after a sentinel)
    base = 0
    for b in range(40):
        P = allp[code == b]
        m = P.size
        if m == 0:
            continue
        R = np.full(m, base, dtype=np.int64)            # SA row of the head of each element's group
        T = np.zeros(m, dtype=bool)                     # the element's earlier windows held its sentinel
        base += m
        off = 0
        while P.size:
            rid = np.searchsorted(dollar, P, side="left")   # the read of position p = number of sentinels before it
            K = np.where(T, rid.astype(np.uint64), window_key(text, P, off))
            o = np.argsort(K, kind="stable")
            o = o[np.argsort(R[o], kind="stable")]
            P, R, K, T = P[o], R[o], K[o], T[o]
            m = P.size
            idx = np.arange(m)
            oldhead = np.ones(m, dtype=bool)
            oldhead[1:] = R[1:] != R[:-1]
            ogs = np.maximum.accumulate(np.where(oldhead, idx, 0))
            row = R + idx - ogs                             # absolute SA row of every element
            head = oldhead.copy()
            head[1:] |= K[1:] != K[:-1]
            nxt_head = np.ones(m, dtype=bool)
            nxt_head[:-1] = head[1:]
            single = head & nxt_head
            sa[row[single]] = P[single]
            ngs = np.maximum.accumulate(np.where(head, idx, 0))
            newR = row[ngs]
            Tn = T | ((K & np.uint64(7)) == 0)
            if np.any(T & ~single):
                raise AssertionError("two suffixes of one read index")
            keep = ~single
            P, R, T = P[keep], newR[keep], Tn[keep]
            off += SYM_PER_KEY
    assert base == N
    return sa


def bwt_from_sa(text: np.ndarray, sa: np.ndarray) -> np.ndarray:
    return np.where(sa == 0, 0, text[np.maximum(sa - 1, 0)]).astype(np.uint8)


def run_length_bytes(bwt: np.ndarray) -> np.ndarray:
    """RLUnit bytes (SuffixTools/RLUnit.h:13-16): symbol rank in the top 3 bits, count (1..31) below."""
    N = bwt.size
    head = np.ones(N, dtype=bool)
    head[1:] = bwt[1:] != bwt[:-1]
    starts = np.nonzero(head)[0]
    lens = np.diff(np.concatenate([starts, [N]]))
    syms = bwt[starts]
    units = (lens + 30) // 31
    out_sym = np.repeat(syms, units)
    first = np.cumsum(units) - units
    k = np.arange(int(units.sum())) - np.repeat(first, units)
    rl = np.repeat(lens, units)
    cnt = np.minimum(31, rl - 31 * k)
    return ((out_sym.astype(np.uint8) << 5) | cnt.astype(np.uint8)).astype(np.uint8)


def write_bwt(path: str, runs: np.ndarray, n_strings: int, n_symbols: int) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<HQQQI", 0xCACA, n_strings, n_symbols, runs.size, 0))
        f.write(runs.tobytes())


def write_sai(path: str, lex: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write(f"51914\n{lex.size}\n{lex.size}\n")
        f.write("".join(f"{int(r)} 0\n" for r in lex))


def build(prefix: str, codes: np.ndarray, offsets: np.ndarray) -> None:
    n = offsets.size - 1
    for ext, sext, rev in ((".bwt", ".sai", False), (".rbwt", ".rsai", True)):
        text, dollar = make_text(codes, offsets, rev)
        sa = suffix_array(text, dollar)
        bwt = bwt_from_sa(text, sa)
        write_bwt(prefix + ext, run_length_bytes(bwt), n, text.size)
        write_sai(prefix + sext, np.searchsorted(dollar, sa[bwt == 0], side="left"))


def read_fasta_codes(path: str):
    lut = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
        lut[ord(chr(c).lower())] = i
    seqs = [line.strip() for line in open(path, "rb") if not line.startswith(b">")]
    codes = np.concatenate([lut[np.frombuffer(s, dtype=np.uint8)] for s in seqs])
    if (codes == 255).any():
        raise SystemExit("only ACGT reads")
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    return codes, offsets


if __name__ == "__main__":
    c, o = read_fasta_codes(sys.argv[1])
    build(sys.argv[2], c, o)
