"""Experiment: bit cost of allele-domain LZ parses (deflate tokens, static per-MAF Huffman) vs zlib on autosome rows.
usage: lzproto.py [variant...]"""
import sys, heapq, zlib, numpy as np
sys.path.insert(0, "/root/repo")
from dna_factory_b200.maf_cdf import MAF_CDF
N = 20000
maf = np.array([m for m, _ in MAF_CDF]); cdf = np.array([c for _, c in MAF_CDF])
pdf = np.diff(np.concatenate([[0], cdf])); w_all = pdf[1:] / pdf[1:].sum(); maf_all = maf[1:]
# representative bins: aggregate weights of neighbouring bins onto a subset
REP = [0.01, 0.015, 0.02, 0.03, 0.045, 0.065, 0.09, 0.12, 0.16, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.495]
wrep = np.zeros(len(REP))
for p, wi in zip(maf_all, w_all):
    wrep[int(np.argmin([abs(p - r) for r in REP]))] += wi

LEN_BASE = [3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LEN_EXTRA = [0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]
DIST_BASE = [1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577]
DIST_EXTRA = [0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13]
import bisect
def len_sym(l):
    i = bisect.bisect_right(LEN_BASE, l) - 1
    if l == 258: i = 28
    return i, LEN_EXTRA[i]
def dist_sym(d):
    i = bisect.bisect_right(DIST_BASE, d) - 1
    return i, DIST_EXTRA[i]

def huff_cost(counts, maxlen=15):
    items = [c for c in counts if c > 0]
    if len(items) <= 1: return sum(items)  # 1 bit each
    h = [(c, i, None) for i, c in enumerate(items)]
    heapq.heapify(h); uid = len(items); total = 0
    while len(h) > 1:
        a = heapq.heappop(h); b = heapq.heappop(h)
        total += a[0] + b[0]
        heapq.heappush(h, (a[0] + b[0], uid, None)); uid += 1
    return total  # = sum count*depth (ignores the 15-bit cap; fine for estimates)

class Stats:
    def __init__(self):
        self.ll = {}; self.dd = {}; self.extra = 0; self.calls = 0
    def lit(self, c): self.ll[c] = self.ll.get(c, 0) + 1
    def match(self, l, d):
        assert 3 <= l <= 258 and 1 <= d <= 32768, (l, d)
        s, e = len_sym(l); self.ll[257 + s] = self.ll.get(257 + s, 0) + 1; self.extra += e
        s, e = dist_sym(d); self.dd[s] = self.dd.get(s, 0) + 1; self.extra += e
    def bits(self):
        return huff_cost(self.ll.values()) + huff_cost(self.dd.values()) + self.extra

def parse_block(z, nall, st, cfg):
    """z: python int, bit s = allele s of the block (nall alleles); emits tokens into st.
    byte coordinates: allele s at byte 2s, its separator at 2s+1; the block owns separator -1 (before allele 0)... we
    treat byte -1 as part of the block start (cost ignored ~ constant)."""
    span = cfg.get("span", 128)        # alleles per span (token boundary); 0 = none
    L0 = cfg.get("keybits", 12)        # alleles hashed
    cap = cfg.get("chain", 64)
    ds = cfg.get("near", (1,))         # always-tried cell distances
    use_hash = cfg.get("hash", True)
    only_events = cfg.get("only_events", False)
    maxd = 8192
    # position lists per key
    table = {}
    keymask = (1 << L0) - 1
    keys = None
    if use_hash:
        keys = [0] * nall
        for s in range(nall):
            keys[s] = ((z >> s) & keymask) << 1 | (s & 1)
    inserted_upto = 0
    segmax = cfg.get("segmax", 0)      # number of segments per block (0: chains)
    seglen = (nall + segmax - 1) // segmax if segmax else 0
    segtab = {}
    if segmax:
        for s_ in range(nall):
            if s_ + L0 <= nall:
                segtab[(s_ // seglen, keys[s_])] = s_     # increasing s_: the last one stays = max
    def insert_upto(s_end):
        nonlocal inserted_upto
        for s in range(inserted_upto, s_end):
            if s + L0 <= nall and (not only_events or (z >> s) & 1):
                table.setdefault(keys[s], []).append(s)
        inserted_upto = max(inserted_upto, s_end)
    def common(s, D, limit):
        # number of alleles from s on equal to those 2D alleles back, at most limit
        x = ((z >> s) ^ (z >> (s - 2 * D))) & ((1 << limit) - 1)
        if x == 0: return limit
        return (x & -x).bit_length() - 1
    p = 0  # byte position (allele 0 at byte 0); start with allele 0 literal handled as generic
    nbytes = 2 * nall
    while p < nbytes - 1:   # last separator belongs to the next block / newline: ignore
        s = (p + 1) // 2    # first allele at or after p
        span_end = nall if not span else min(nall, (s // span + 1) * span)
        # bytes available in this span: up to byte 2*span_end - 2 (the separator before the next span's first allele is owned by next span)
        lim_bytes = min(258, 2 * span_end - 1 - p)
        best_len, best_d, best_sc = 0, 0, 0
        odd = p & 1
        if s < span_end:
            limit = min(span_end - s, 130)
            cands = [D for D in ds if s - 2 * D >= 0]
            if segmax and s + L0 <= nall:
                sg = s // seglen
                for g in range(sg, -1, -1):
                    j = segtab.get((g, keys[s]))
                    if j is not None and j < s and (s - j) % 2 == 0 and (s - j) // 2 <= maxd:
                        cands.append((s - j) // 2)
            elif use_hash and s + L0 <= nall:
                insert_upto(s)   # everything strictly before s (sources may overlap the target)
                lst = table.get(keys[s])
                if lst:
                    n = 0
                    for j in reversed(lst):
                        D2 = s - j
                        if D2 // 2 > maxd: break
                        cands.append(D2 // 2); n += 1
                        if n >= cap: break
            for D in cands:
                k = common(s, D, limit)
                l = min(2 * k + odd, lim_bytes)
                if cfg.get("rate"):
                    if l >= 3:
                        sc = l * cfg["rate"] - (1 if D == 1 else 3 if D == 2 else 2 + (4 * D).bit_length())
                        if best_len == 0 or sc > best_sc:
                            best_len, best_d, best_sc = l, D, sc
                # tie-break: prefer longer; equal -> nearer
                elif l > best_len or (l == best_len and D < best_d):
                    best_len, best_d = l, D
        elif odd and lim_bytes >= 1:
            best_len = 0
        if cfg.get("lazy") and best_len >= 3 and best_len < cfg["lazy"] and not odd and s + 1 < span_end:
            # candidate: literal allele, then best match from p+1 (allele s+1 at odd byte)
            s2 = s + 1; limit2 = min(span_end - s2, 130); lim2 = min(258, 2 * span_end - 1 - (p + 1))
            c2 = [D for D in ds if s2 - 2 * D >= 0]
            if use_hash and s2 + L0 <= nall:
                insert_upto(s2)
                lst = table.get(keys[s2])
                if lst:
                    n = 0
                    for j in reversed(lst):
                        if (s2 - j) // 2 > maxd: break
                        c2.append((s2 - j) // 2); n += 1
                        if n >= cap: break
            b2 = 0
            for D in c2:
                k = common(s2, D, limit2); l = min(2 * k + 1, lim2)
                if l > b2: b2 = l
            if b2 > best_len + 1:
                best_len = 0   # take the literal now; the next iteration finds the longer match
        if best_len >= 3 and not (cfg.get("minfar", 0) and best_d > 1 and best_len < cfg["minfar"]):
            st.match(best_len, 4 * best_d); p += best_len
        else:
            if best_len >= 3:  # far match too short: fall back to the near candidates only
                bl, bd = 0, 0
                for D in ds:
                    if s - 2 * D >= 0:
                        k = common(s, D, min(span_end - s, 130)); l = min(2 * k + odd, lim_bytes)
                        if l > bl: bl, bd = l, D
                if bl >= 3:
                    st.match(bl, 4 * bd); p += bl; continue
            st.lit(('/' if (p // 2) % 2 == 0 else 't') if odd else ((z >> (p // 2)) & 1)); p += 1

def run(cfg, reps=REP, rows=2, seed=1):
    rng = np.random.default_rng(seed)
    tot = 0.0; out = []
    for p, wi in zip(reps, wrep):
        st = Stats()
        if cfg.get("rate") == "auto" or cfg.get("_auto"):
            cfg = dict(cfg); cfg["_auto"] = True
            h = -(p * np.log2(p) + (1 - p) * np.log2(1 - p))
            cfg["rate"] = cfg.get("ratemul", 1.6) * h / 2.0     # bits per text byte actually paid (about 1.6 x entropy)
        for r in range(rows):
            a = (rng.random(2 * N) < p)
            nseg = cfg.get("nseg", 2)
            per = (N + nseg - 1) // nseg
            for g in range(nseg):
                bits = a[2 * g * per: 2 * min(N, (g + 1) * per)]
                z = int.from_bytes(np.packbits(bits, bitorder="little").tobytes(), "little")
                parse_block(z, len(bits), st, cfg)
        b = st.bits() / (rows * N)
        out.append(b); tot += wi * b
    return tot, out

def zl(level, reps=REP, rows=2, seed=1):
    rng = np.random.default_rng(seed); tot = 0; out = []
    for p, wi in zip(reps, wrep):
        bits = 0
        for r in range(rows):
            a = (rng.random(2 * N) < p).astype(np.uint8) + 48
            t = np.empty(4 * N, np.uint8); t[0::4] = a[0::2]; t[1::4] = 47; t[2::4] = a[1::2]; t[3::4] = 9
            tb = t.tobytes()
            for i in range(0, len(tb), 65536):
                c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, 0)
                bits += 8 * len(c.compress(tb[i:i + 65536]) + c.flush())
        out.append(bits / (rows * N)); tot += wi * out[-1]
    return tot, out

VARIANTS = {
    "p4": dict(hash=False, near=(1,)),
    "p48": dict(hash=False, near=(1, 2)),
    "p4_nospan": dict(hash=False, near=(1,), span=0),
    "lz12": dict(keybits=12, chain=64, near=(1, 2)),
    "lz12_nospan": dict(keybits=12, chain=64, near=(1, 2), span=0),
    "lz12_c8": dict(keybits=12, chain=8, near=(1, 2)),
    "lz12_c2": dict(keybits=12, chain=2, near=(1, 2)),
    "lz8": dict(keybits=8, chain=64, near=(1, 2)),
    "lz16": dict(keybits=16, chain=64, near=(1, 2)),
    "lz12_ev": dict(keybits=12, chain=64, near=(1, 2), only_events=True),
    "sm16_12": dict(keybits=12, segmax=16, near=(1, 2)),
    "sm32_12": dict(keybits=12, segmax=32, near=(1, 2)),
    "sm8_12": dict(keybits=12, segmax=8, near=(1, 2)),
    "sm16_10": dict(keybits=10, segmax=16, near=(1, 2)),
    "sm16_8": dict(keybits=8, segmax=16, near=(1, 2)),
    "sm32_8": dict(keybits=8, segmax=32, near=(1, 2)),
    "sm16_12n4": dict(keybits=12, segmax=16, near=(1, 2, 3, 4)),
    "sm16_12r": dict(keybits=12, segmax=16, near=(1, 2), rate="auto"),
    "sm32_12r": dict(keybits=12, segmax=32, near=(1, 2), rate="auto"),
    "lz12r": dict(keybits=12, chain=64, near=(1, 2), rate="auto"),
    "lz12r_c8": dict(keybits=12, chain=8, near=(1, 2), rate="auto"),
    "lz8r": dict(keybits=8, chain=64, near=(1, 2), rate="auto"),
    "lz10_c8": dict(keybits=10, chain=8, near=(1, 2)),
    "lz10_c16": dict(keybits=10, chain=16, near=(1, 2)),
    "lz10_c16_lazy": dict(keybits=10, chain=16, near=(1, 2), lazy=64),
    "lz10_c16_s256": dict(keybits=10, chain=16, near=(1, 2), span=256),
    "lz10_c64_lazy": dict(keybits=10, chain=64, near=(1, 2), lazy=258),
    "lz10_c4": dict(keybits=10, chain=4, near=(1, 2)),
    "lz10_c2": dict(keybits=10, chain=2, near=(1, 2)),
    "lz10_c1": dict(keybits=10, chain=1, near=(1, 2)),
    "lz12_1seg": dict(keybits=12, chain=64, near=(1, 2), nseg=1),
}
if __name__ == "__main__":
    names = sys.argv[1:] or ["p4", "lz12"]
    print("maf      " + " ".join("%6.3f" % r for r in REP))
    print("weight   " + " ".join("%6.3f" % x for x in wrep))
    for lv in (2, 6, 9):
        t, o = zl(lv)
        print("zlib-%d   " % lv + " ".join("%6.3f" % x for x in o) + "  mix %.3f (%.1fx)" % (t, 32 / t))
    for nm in names:
        t, o = run(VARIANTS[nm])
        print("%-9s" % nm + " ".join("%6.3f" % x for x in o) + "  mix %.3f (%.1fx)" % (t, 32 / t))
