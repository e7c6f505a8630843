"""Seeded synthetic inputs for the pbcorrect hot path.

Genomes are uniform random ACGT, optionally with injected repeat families and tandem
arrays; reads are CLR-like: log-normal lengths, both strands, ~13 % error split
55/30/15 % insertion/deletion/substitution.  Everything is numpy-vectorised so the
230 Mbp of config 2 is generated in seconds, and fully determined by the seeds.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(length: int, seed: int, repeat_families: int = 0, tandem_arrays: int = 0) -> np.ndarray:
    """Return a genome as a uint8 array of codes 0..3 (A,C,G,T)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.integers(0, 4, size=length, dtype=np.uint8)
    for _ in range(repeat_families):
        flen = int(rng.integers(1000, 6001))
        src = int(rng.integers(0, max(1, length - flen)))
        unit = g[src:src + flen].copy()
        for _ in range(20):
            ident = rng.uniform(0.97, 0.99)
            cp = unit.copy()
            mut = rng.random(cp.size) > ident
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            dst = int(rng.integers(0, max(1, length - flen)))
            g[dst:dst + cp.size] = cp[: max(0, min(cp.size, length - dst))]
    for _ in range(tandem_arrays):
        period = int(rng.integers(2, 201))
        total = int(rng.integers(500, 5001))
        unit = rng.integers(0, 4, size=period, dtype=np.uint8)
        arr = np.tile(unit, total // period + 1)[:total]
        dst = int(rng.integers(0, max(1, length - total)))
        g[dst:dst + total] = arr[: max(0, min(total, length - dst))]
    return g


def simulate_reads(genome: np.ndarray, coverage: float, mean_len: int, seed: int,
                   error: float = 0.13, sigma: float = 0.5, min_len: int = 500,
                   ins: float = 0.55, dele: float = 0.30):
    """Return (codes, offsets): concatenated read codes (uint8 0..3) and int64 offsets[n+1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    glen = genome.size
    n_est = int(coverage * glen / mean_len * 1.3) + 16
    mu = np.log(mean_len) - 0.5 * sigma * sigma
    lens = np.maximum(rng.lognormal(mu, sigma, size=n_est).astype(np.int64), min_len)
    lens = np.minimum(lens, glen)
    cum = np.cumsum(lens)
    n = int(np.searchsorted(cum, coverage * glen)) + 1
    lens = lens[:n]
    starts = (rng.random(n) * (glen - lens + 1)).astype(np.int64)
    strand = rng.random(n) < 0.5
    # template bases, concatenated
    toff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=toff[1:])
    total = int(toff[-1])
    rid = np.repeat(np.arange(n, dtype=np.int64), lens)
    within = np.arange(total, dtype=np.int64) - toff[rid]
    rev = strand[rid]
    gpos = np.where(rev, starts[rid] + lens[rid] - 1 - within, starts[rid] + within)
    base = genome[gpos]
    base = np.where(rev, 3 - base, base).astype(np.uint8)
    # per-template-base edit ops
    u = rng.random(total)
    p_del = error * dele
    p_sub = error * (1.0 - ins - dele)
    is_del = u < p_del
    is_sub = (u >= p_del) & (u < p_del + p_sub)
    base = np.where(is_sub, (base + rng.integers(1, 4, size=total, dtype=np.uint8)) & 3, base).astype(np.uint8)
    n_ins = (rng.random(total) < error * ins).astype(np.int64)
    out_cnt = n_ins + (~is_del).astype(np.int64)
    ooff = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(out_cnt, out=ooff[1:])
    out = np.empty(int(ooff[-1]), dtype=np.uint8)
    ins_pos = ooff[:-1][n_ins > 0]
    out[ins_pos] = rng.integers(0, 4, size=ins_pos.size, dtype=np.uint8)
    keep = ~is_del
    out[(ooff[:-1] + n_ins)[keep]] = base[keep]
    offsets = ooff[toff]
    return out, offsets


def simulate_reads_device(genome: np.ndarray, coverage: float, mean_len: int, seed: int, device: str = "cuda:0",
                          chunk_bases: int = 200_000_000, error: float = 0.13, sigma: float = 0.5, min_len: int = 500,
                          ins: float = 0.55, dele: float = 0.30):
    """The read model of simulate_reads with the per-base work on the GPU (torch), a few hundred Mbp of templates at a time: for
    read sets numpy cannot hold in one piece (config 4: 4 Gbp, whose temporaries would take > 150 GB of host memory).  Read
    lengths, positions and strands come from the same numpy generator; the per-base draws come from torch's generator, so a
    seed does not give the reads simulate_reads gives."""
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    glen = genome.size
    n_est = int(coverage * glen / mean_len * 1.3) + 16
    mu = np.log(mean_len) - 0.5 * sigma * sigma
    lens = np.maximum(rng.lognormal(mu, sigma, size=n_est).astype(np.int64), min_len)
    lens = np.minimum(lens, glen)
    cum = np.cumsum(lens)
    n = int(np.searchsorted(cum, coverage * glen)) + 1
    lens = lens[:n]
    starts = (rng.random(n) * (glen - lens + 1)).astype(np.int64)
    strand = rng.random(n) < 0.5
    dev = torch.device(device)
    g = torch.from_numpy(np.ascontiguousarray(genome)).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    p_del = error * dele
    p_sub = error * (1.0 - ins - dele)
    chunks, read_lens = [], []
    toff_all = np.concatenate([[0], np.cumsum(lens)])
    r0 = 0
    while r0 < n:
        r1 = int(np.searchsorted(toff_all, toff_all[r0] + chunk_bases, side="right")) - 1
        r1 = min(n, max(r1, r0 + 1))
        l = torch.from_numpy(lens[r0:r1]).to(dev)
        st = torch.from_numpy(starts[r0:r1]).to(dev)
        sd = torch.from_numpy(strand[r0:r1]).to(dev)
        m = r1 - r0
        toff = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        toff[1:] = torch.cumsum(l, 0)
        total = int(toff[-1])
        rid = torch.repeat_interleave(torch.arange(m, device=dev), l)
        within = torch.arange(total, device=dev) - toff[rid]
        rev = sd[rid]
        gpos = torch.where(rev, st[rid] + l[rid] - 1 - within, st[rid] + within)
        del within
        base = g[gpos]
        del gpos
        base = torch.where(rev, 3 - base, base)
        del rev, rid
        u = torch.rand(total, device=dev, generator=gen)
        is_del = u < p_del
        is_sub = (u >= p_del) & (u < p_del + p_sub)
        del u
        sub = torch.randint(1, 4, (total,), device=dev, generator=gen, dtype=torch.uint8)
        base = torch.where(is_sub, (base + sub) & 3, base)
        del sub, is_sub
        n_ins = (torch.rand(total, device=dev, generator=gen) < error * ins).to(torch.int64)
        out_cnt = n_ins + (~is_del).to(torch.int64)
        ooff = torch.zeros(total + 1, dtype=torch.int64, device=dev)
        ooff[1:] = torch.cumsum(out_cnt, 0)
        del out_cnt
        out = torch.empty(int(ooff[-1]), dtype=torch.uint8, device=dev)
        ins_pos = ooff[:-1][n_ins > 0]
        out[ins_pos] = torch.randint(0, 4, (int(ins_pos.numel()),), device=dev, generator=gen, dtype=torch.uint8)
        keep = ~is_del
        out[(ooff[:-1] + n_ins)[keep]] = base[keep]
        ends = ooff[toff]
        read_lens.append((ends[1:] - ends[:-1]).cpu().numpy())
        chunks.append(out.cpu().numpy())
        del out, ooff, n_ins, keep, is_del, base, ins_pos
        r0 = r1
    del g
    torch.cuda.empty_cache()
    codes = np.concatenate(chunks)
    offsets = np.concatenate([[0], np.cumsum(np.concatenate(read_lens))]).astype(np.int64)
    return codes, offsets


def write_fasta(path: str, codes: np.ndarray, offsets: np.ndarray, prefix: str = "r") -> None:
    letters = _ACGT[codes]
    with open(path, "wb") as f:
        for i in range(offsets.size - 1):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            f.write(letters[offsets[i]:offsets[i + 1]].tobytes())
            f.write(b"\n")


def read_strings(codes: np.ndarray, offsets: np.ndarray):
    letters = _ACGT[codes].tobytes()
    return [letters[offsets[i]:offsets[i + 1]].decode() for i in range(offsets.size - 1)]
