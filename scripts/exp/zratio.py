"""Experiment: bits per call of zlib levels by MAF bin on SnpFactory-shaped autosome rows (iid alleles)."""
import sys, zlib, numpy as np
sys.path.insert(0, "/root/repo")
from dna_factory_b200.maf_cdf import MAF_CDF
N = 20000
rng = np.random.default_rng(1)
maf = np.array([m for m, _ in MAF_CDF]); cdf = np.array([c for _, c in MAF_CDF])
pdf = np.diff(np.concatenate([[0], cdf]))
start = 1  # -f 0.01 -> bins with maf >= 0.01
w = pdf[start:] / pdf[start:].sum()
def row(p, nrows=3):
    out = []
    for r in range(nrows):
        a = (rng.random(2 * N) < p).astype(np.uint8) + 48
        t = np.empty(4 * N, np.uint8); t[0::4] = a[0::2]; t[1::4] = 47; t[2::4] = a[1::2]; t[3::4] = 9; t[-1] = 10
        out.append(b"1\t12345678\trs1234567\tA\tC\t40\tPASS\t.\tGT\t" + t.tobytes())
    return b"".join(out)
def bits(text, level):
    tot = 0
    for i in range(0, len(text), 65536):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, 0)
        tot += len(c.compress(text[i:i + 65536]) + c.flush())
    return tot * 8
levels = (1, 2, 4, 6, 9)
acc = {l: 0.0 for l in levels}; ent = 0.0
print("maf   w     H/call " + " ".join("z%d" % l for l in levels))
for p, wi in zip(maf[start:], w):
    t = row(p); calls = 3 * N
    h = 2 * (-(p * np.log2(p) + (1 - p) * np.log2(1 - p)))
    b = {l: bits(t, l) / calls for l in levels}
    for l in levels: acc[l] += wi * b[l]
    ent += wi * h
    if round(p * 1000) % 25 == 0 or p < 0.03:
        print("%.3f %.4f %.3f  " % (p, wi, h) + " ".join("%.2f" % b[l] for l in levels))
print("mix: entropy %.3f bits/call; " % ent + " ".join("z%d %.3f (%.1fx)" % (l, acc[l], 3.952 * 8 / acc[l]) for l in levels))
