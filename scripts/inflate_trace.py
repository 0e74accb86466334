"""Ad-hoc: token-level trace of one raw deflate stream (dynamic Huffman block)."""
import sys

class Bits:
    def __init__(self, data): self.d = data; self.pos = 0
    def get(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.pos >> 3] >> (self.pos & 7)) & 1) << i
            self.pos += 1
        return v

def build(lengths):
    codes = {}
    code = 0
    maxl = max(lengths) if lengths else 0
    bl_count = [0] * (maxl + 2)
    for l in lengths:
        if l: bl_count[l] += 1
    nxt = [0] * (maxl + 2)
    for b in range(1, maxl + 1):
        code = (code + bl_count[b - 1]) << 1
        nxt[b] = code
    for s, l in enumerate(lengths):
        if l:
            codes[(l, nxt[l])] = s
            nxt[l] += 1
    return codes

def decode(bits, codes):
    code = 0; l = 0
    while True:
        code = (code << 1) | bits.get(1); l += 1
        if (l, code) in codes: return codes[(l, code)], l
        if l > 15: raise ValueError("bad code at bit %d" % bits.pos)

LEN_BASE = [3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LEN_EXTRA = [0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]
DIST_BASE = [1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577]
DIST_EXTRA = [0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13]

def trace(data, max_tokens=60, skip=0):
    b = Bits(data)
    final, btype = b.get(1), b.get(2)
    hlit, hdist, hclen = b.get(5) + 257, b.get(5) + 1, b.get(4) + 4
    order = [16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15]
    cl = [0] * 19
    for i in range(hclen): cl[order[i]] = b.get(3)
    clc = build(cl)
    lens = []
    while len(lens) < hlit + hdist:
        s, _ = decode(b, clc)
        if s < 16: lens.append(s)
        elif s == 16: lens += [lens[-1]] * (3 + b.get(2))
        elif s == 17: lens += [0] * (3 + b.get(3))
        else: lens += [0] * (11 + b.get(7))
    ll, dl = lens[:hlit], lens[hlit:]
    print("header ends at bit", b.pos, "hlit", hlit, "hdist", hdist, "dist lens", dl)
    llc, dc = build(ll), build(dl)
    out = bytearray(); n = 0
    while True:
        p0 = b.pos
        s, l = decode(b, llc)
        if s < 256:
            out.append(s); desc = "lit %r (%d bits)" % (chr(s), l)
        elif s == 256:
            print("bit %d EOB" % p0); break
        else:
            i = s - 257; ln = LEN_BASE[i] + b.get(LEN_EXTRA[i])
            ds, dlb = decode(b, dc) if len(dc) > 1 or True else (0, 0)
            dist = DIST_BASE[ds] + b.get(DIST_EXTRA[ds])
            for _ in range(ln): out.append(out[-dist])
            desc = "match len %d dist %d (%d bits)" % (ln, dist, b.pos - p0)
        if n >= skip and n < skip + max_tokens: print("bit %d: %s" % (p0, desc))
        n += 1
    return bytes(out)
