"""Test / measurement helper, NOT the product's index builder (that is csrc/pbsc_build.cu, `pbcorrect index`, `api.build_bwt`).

The two BWTs `stride index` produces (PREFIX.bwt = BWT of the reads, PREFIX.rbwt = BWT of the reversed reads;
SuffixTools/BWTCARopebwt.cpp:160-247) by prefix-doubling suffix sorting with torch.sort, on the GPU or, for small inputs, on the
CPU.  It is what the CPU tests use to make index files without a GPU, what the reference arm of bench.py uses to give the
reference its index (none of the product's kernels on that arm), and an independent check of the product builder
(`test_larger_input_equals_the_torch_builder`).  Each read gets its own sentinel, ordered by read index, which is the order
ropebwt2 produces with MR_SO_IO: the .bwt / .rbwt bytes equal the reference's (checked on tests/golden/tiny.*).  `write_sai_file`
writes a placeholder read order (identity); the real .sai / .rsai come from the product builder.
"""
from __future__ import annotations

import struct

import numpy as np
import torch


def _text(codes: torch.Tensor, offsets: torch.Tensor, reverse: bool):
    """Concatenate reads (optionally each reversed) with one '$' after each.  Returns (sym[N] uint8 in 0..4,
    rid[N] int64 read index, dollar_pos[n] int64)."""
    n = offsets.numel() - 1
    lens = offsets[1:] - offsets[:-1]
    total = int(offsets[-1])
    dev = codes.device
    rid_base = torch.repeat_interleave(torch.arange(n, device=dev), lens)
    pos = torch.arange(total, device=dev)
    if reverse:
        src = offsets[:-1][rid_base] + (offsets[1:][rid_base] - 1 - pos)
        bases = codes[src]
    else:
        bases = codes
    N = total + n
    sym = torch.zeros(N, dtype=torch.uint8, device=dev)
    tpos = pos + rid_base                       # read i is shifted by its i preceding sentinels
    sym[tpos] = bases.to(torch.uint8) + 1
    dollar = offsets[1:] + torch.arange(n, device=dev)
    rid = torch.zeros(N, dtype=torch.int64, device=dev)
    rid[tpos] = rid_base
    rid[dollar] = torch.arange(n, device=dev)
    return sym, rid, dollar


def suffix_array(sym: torch.Tensor, rid: torch.Tensor, dollar: torch.Tensor) -> torch.Tensor:
    N = sym.numel()
    dev = sym.device
    p = torch.arange(N, device=dev)
    # exact rank for h = 8 symbols: base-5 code truncated after the first '$', ties among windows that contain a
    # sentinel broken by its read index
    d = dollar[rid] - p                          # distance to this read's sentinel
    W = 8
    code = torch.zeros(N, dtype=torch.int64, device=dev)
    padded = torch.cat([sym, torch.zeros(W, dtype=torch.uint8, device=dev)]).to(torch.int64)
    for j in range(W):
        c = padded[j:j + N]
        c = torch.where(d >= j, c, torch.zeros_like(c))
        code = code * 5 + c
    tie = torch.where(d < W, rid + 1, torch.zeros_like(rid))
    key = (code << 32) | tie
    del code, tie, padded, d
    h = W
    rank = None
    while True:
        skey, perm = torch.sort(key)
        flag = torch.ones(N, dtype=torch.int64, device=dev)
        flag[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        flag[0] = 0
        r_sorted = torch.cumsum(flag, 0)
        n_unique = int(r_sorted[-1]) + 1
        if n_unique == N:
            return perm
        rank = torch.empty(N, dtype=torch.int64, device=dev)
        rank[perm] = r_sorted
        del skey, perm, flag, r_sorted
        r2 = torch.zeros(N, dtype=torch.int64, device=dev)
        if h < N:
            r2[: N - h] = rank[h:] + 1
        key = (rank << 32) | r2
        del r2
        h *= 2


def bwt_symbols(codes, offsets, reverse: bool = False, device: str | None = None) -> torch.Tensor:
    """BWT of the read collection as ranks 0..4 ($ACGT), uint8 tensor of length bases + reads."""
    dev = torch.device(device) if device else (torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))
    c = torch.as_tensor(np.asarray(codes), device=dev).to(torch.int64)
    o = torch.as_tensor(np.asarray(offsets, dtype=np.int64), device=dev)
    sym, rid, dollar = _text(c, o, reverse)
    sa = suffix_array(sym, rid, dollar)
    prev = sa - 1
    prev[prev < 0] = sym.numel() - 1
    return sym[prev]


def run_length_bytes(bwt: torch.Tensor) -> np.ndarray:
    """Encode BWT ranks as the reference's run bytes: symbol rank << 5 | run length 1..31 (RLUnit.h:13-16)."""
    n = bwt.numel()
    dev = bwt.device
    b = bwt.to(torch.int64)
    change = torch.ones(n, dtype=torch.bool, device=dev)
    change[1:] = b[1:] != b[:-1]
    starts = torch.nonzero(change).flatten()
    lens = torch.diff(torch.cat([starts, torch.tensor([n], device=dev)]))
    sym = b[starts]
    chunks = (lens + 30) // 31
    run_id = torch.repeat_interleave(torch.arange(starts.numel(), device=dev), chunks)
    first_chunk = torch.cumsum(chunks, 0) - chunks
    idx_in_run = torch.arange(run_id.numel(), device=dev) - first_chunk[run_id]
    is_last = idx_in_run == (chunks[run_id] - 1)
    clen = torch.where(is_last, lens[run_id] - 31 * (chunks[run_id] - 1), torch.full_like(run_id, 31))
    out = ((sym[run_id] << 5) | clen).to(torch.uint8)
    return out.cpu().numpy()


def write_bwt_file(path: str, runs: np.ndarray, n_strings: int, n_symbols: int) -> None:
    """On-disk format of BWTWriterBinary (SuffixTools/BWTWriterBinary.cpp:28-94)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<HQQQi", 0xCACA, n_strings, n_symbols, runs.size, 0))
        f.write(runs.tobytes())


def write_sai_file(path: str, n_strings: int) -> None:
    """A well-formed .sai (SampledSuffixArray::readSAI); pbcorrect loads it and never consults it."""
    with open(path, "w") as f:
        f.write(f"51914\n{n_strings}\n{n_strings}\n")
        f.write("".join(f"{i} 0\n" for i in range(n_strings)))


def build_index_files(prefix: str, codes, offsets, device: str | None = None):
    """Write PREFIX.bwt / .rbwt / .sai / .rsai; returns dict with runs and sizes for pbsc_index_create."""
    n = len(offsets) - 1
    res = {}
    for ext, rev in (("bwt", False), ("rbwt", True)):
        b = bwt_symbols(codes, offsets, reverse=rev, device=device)
        runs = run_length_bytes(b)
        write_bwt_file(f"{prefix}.{ext}", runs, n, int(b.numel()))
        res[ext] = (runs, int(b.numel()), n)
        del b
    write_sai_file(prefix + ".sai", n)
    write_sai_file(prefix + ".rsai", n)
    return res
