"""Deterministic synthetic valid pairs: positions uniform over the
mappable part of each chromosome, separation log-uniform in [1 kb, chromosome length] with a
random sign, acceptance weighted by a per-bin log-normal bias so ICE has structure to remove,
and an unmappable hole (0-9.5 Mb on acrocentric chromosomes, a 3 Mb centromere elsewhere) to
exercise the min_nnz / MAD-max / gap logic.  NumPy version for tests and small runs; torch
version for generating bench-sized inputs directly on the GPU.
"""
from __future__ import annotations

import numpy as np

HG19 = {
    "1": 249250621, "2": 243199373, "3": 198022430, "4": 191154276, "5": 180915260, "6": 171115067,
    "7": 159138663, "8": 146364022, "9": 141213431, "10": 135534747, "11": 135006516, "12": 133851895,
    "13": 115169878, "14": 107349540, "15": 102531392, "16": 90354753, "17": 81195210, "18": 78077248,
    "19": 59128983, "20": 63025520, "21": 48129895, "22": 51304566, "X": 155270560, "Y": 59373566,
    "M": 16571,
}
ACROCENTRIC = {"13", "14", "15", "21", "22"}
BIAS_RES = 40000


def write_genome_size(path, genome=None):
    genome = HG19 if genome is None else genome
    with open(path, "w") as fh:
        for c, l in genome.items():
            fh.write("chr%s\t%d\n" % (c, l))
    return path


def hole(chrom: str, length: int):
    """(start, end) of the unmappable region."""
    if chrom in ACROCENTRIC:
        return 0, min(9_500_000, length // 3)
    mid = int(length * 0.4)
    return mid, min(mid + 3_000_000, length)


def _bias_track(rng, length):
    return np.exp(rng.normal(0.0, 0.3, size=length // BIAS_RES + 1))


def cis_pairs(chrom: str, length: int, n: int, seed: int):
    """n accepted cis pairs (pos1, pos2) on one chromosome."""
    rng = np.random.default_rng(seed)
    bias = _bias_track(rng, length)
    bmax = bias.max()
    h0, h1 = hole(chrom, length)
    out1, out2, have = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], 0
    while have < n:
        m = int((n - have) * 2.2) + 1024
        p1 = rng.integers(0, length, size=m)
        sep = np.exp(rng.uniform(np.log(1e3), np.log(length), size=m)).astype(np.int64)
        p2 = p1 + np.where(rng.random(m) < 0.5, -sep, sep)
        ok = (p2 >= 0) & (p2 < length)
        ok &= ~((p1 >= h0) & (p1 < h1)) & ~((p2 >= h0) & (p2 < h1))
        p2c = np.clip(p2, 0, length - 1)
        acc = bias[p1 // BIAS_RES] * bias[p2c // BIAS_RES] / (bmax * bmax)
        ok &= rng.random(m) < acc
        out1.append(p1[ok]); out2.append(p2[ok]); have += int(ok.sum())
    return np.concatenate(out1)[:n].astype(np.int32), np.concatenate(out2)[:n].astype(np.int32)


def genome_pairs(genome: dict, order, n_total: int, seed: int, trans_frac=0.0):
    """Columnar pairs over several chromosomes: cis pairs proportional to chromosome length,
    plus ``trans_frac`` uniform trans pairs.  Chromosome ids index ``order``.  Shuffled."""
    rng = np.random.default_rng(seed)
    lens = np.array([genome[c] for c in order], dtype=np.float64)
    n_trans = int(n_total * trans_frac)
    share = np.floor((n_total - n_trans) * lens / lens.sum()).astype(np.int64)
    share[0] += (n_total - n_trans) - share.sum()
    c1, p1, c2, p2 = [], [], [], []
    for i, c in enumerate(order):
        a, b = cis_pairs(c, genome[c], int(share[i]), seed * 1000 + i)
        c1.append(np.full(a.size, i, np.int32)); c2.append(np.full(a.size, i, np.int32))
        p1.append(a); p2.append(b)
    if n_trans:
        w = lens / lens.sum()
        ca = rng.choice(len(order), size=n_trans, p=w).astype(np.int32)
        cb = rng.choice(len(order), size=n_trans, p=w).astype(np.int32)
        same = ca == cb
        cb[same] = (cb[same] + 1) % len(order)
        pa = (rng.random(n_trans) * lens[ca]).astype(np.int32)
        pb = (rng.random(n_trans) * lens[cb]).astype(np.int32)
        c1.append(ca); c2.append(cb); p1.append(pa); p2.append(pb)
    c1, p1, c2, p2 = (np.concatenate(x) for x in (c1, p1, c2, p2))
    perm = rng.permutation(c1.size)
    return c1[perm], p1[perm], c2[perm], p2[perm]


def valid23_lines(order, c1, p1, c2, p2):
    """23-column ``*_Valid.bed`` text (filtering.py:16-47); only columns 1, 6, 8, 13 carry data."""
    for a, x, b, y in zip(c1, p1, c2, p2):
        yield ("read\tchr%s\t+\t%d\t50\t0\t%d\t0\tchr%s\t-\t%d\t50\t0\t%d\t0\t.\t.\t.\t.\t.\t.\t.\t.\n"
               % (order[a], x, x, order[b], y, y))


def allelic_lines(order, c1, p1, c2, p2, mark=None):
    names = {0: "Both", 1: "R1", 2: "R2"}
    for i, (a, x, b, y) in enumerate(zip(c1, p1, c2, p2)):
        if mark is None:
            yield "chr%s\t%d\tchr%s\t%d\n" % (order[a], x, order[b], y)
        else:
            yield "chr%s\t%d\tchr%s\t%d\t%s\n" % (order[a], x, order[b], y, names[int(mark[i])])


# ---- device-side generator for bench-sized inputs ---------------------------------------
def genome_pairs_torch(genome: dict, order, n_total: int, seed: int, device, trans_frac=0.0):
    """Same pair model generated with torch on ``device`` (not bit-identical to the NumPy
    generator).  Returns int32 tensors (c1, p1, c2, p2), shuffled."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.tensor([genome[c] for c in order], dtype=torch.float64)
    n_trans = int(n_total * trans_frac)
    share = torch.floor((n_total - n_trans) * lens / lens.sum()).to(torch.int64)
    share[0] += (n_total - n_trans) - int(share.sum())
    cs, p1s, p2s = [], [], []
    for i, c in enumerate(order):
        L, n = genome[c], int(share[i])
        nb = L // BIAS_RES + 1
        bias = torch.exp(0.3 * torch.randn(nb, generator=g, device=device, dtype=torch.float32))
        bmax2 = float(bias.max()) ** 2
        h0, h1 = hole(c, L)
        got1, got2, have = [], [], 0
        while have < n:
            m = int((n - have) * 2.2) + 4096
            a = (torch.rand(m, generator=g, device=device, dtype=torch.float64) * L).to(torch.int64)
            lo, hi = np.log(1e3), np.log(L)
            sep = torch.exp(torch.rand(m, generator=g, device=device, dtype=torch.float64) * (hi - lo) + lo).to(torch.int64)
            sign = torch.rand(m, generator=g, device=device) < 0.5
            b = a + torch.where(sign, -sep, sep)
            ok = (b >= 0) & (b < L) & ~((a >= h0) & (a < h1)) & ~((b >= h0) & (b < h1))
            bc = b.clamp(0, L - 1)
            acc = bias[a // BIAS_RES] * bias[bc // BIAS_RES] / bmax2
            ok &= torch.rand(m, generator=g, device=device) < acc
            got1.append(a[ok]); got2.append(b[ok]); have += int(ok.sum())
        p1s.append(torch.cat(got1)[:n].to(torch.int32)); p2s.append(torch.cat(got2)[:n].to(torch.int32))
        cs.append(torch.full((n,), i, dtype=torch.int32, device=device))
    c1 = torch.cat(cs); c2 = c1.clone(); p1 = torch.cat(p1s); p2 = torch.cat(p2s)
    if n_trans:
        w = (lens / lens.sum()).to(device)
        ca = torch.multinomial(w, n_trans, replacement=True, generator=g).to(torch.int32)
        cb = torch.multinomial(w, n_trans, replacement=True, generator=g).to(torch.int32)
        same = ca == cb
        cb[same] = (cb[same] + 1) % len(order)
        ld = lens.to(device)
        pa = (torch.rand(n_trans, generator=g, device=device, dtype=torch.float64) * ld[ca.long()]).to(torch.int32)
        pb = (torch.rand(n_trans, generator=g, device=device, dtype=torch.float64) * ld[cb.long()]).to(torch.int32)
        c1 = torch.cat([c1, ca]); c2 = torch.cat([c2, cb]); p1 = torch.cat([p1, pa]); p2 = torch.cat([p2, pb])
    n = int(c1.numel())
    if n <= (1 << 30):
        perm = torch.randperm(n, generator=g, device=device)
    else:
        # torch.randperm does not finish in minutes at 2 G elements on a B200 (measured: 0.4 s at 0.5 G, > 240 s at 2 G); a
        # multiplicative bijection i -> (i * A + B) mod n with gcd(A, n) = 1 interleaves the chromosome blocks just as well
        A = 1_000_000_007
        while np.gcd(A, n) != 1:
            A += 2
        perm = (torch.arange(n, dtype=torch.int64, device=device) * A + 12345) % n
    out = []
    for t in (c1, p1, c2, p2):          # one column at a time: the gathered copy replaces its source
        out.append(t[perm].contiguous())
    return tuple(out)
