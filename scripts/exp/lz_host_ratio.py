"""Bits per call of the LZ tiers' host twin (dnaf_debug_lz_block) on the reference MAF mix, N = 20000 (two segments)."""
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from dna_factory_b200 import _native
from scripts.exp.lzproto import REP, wrep
sys.path.insert(0, "/root/repo/tests")
from tests.test_lz_host import _segment
levels = [int(x) for x in sys.argv[1:]] or [4, 5, 6, 7, 8, 9]
for lv in levels:
    t0 = time.time(); tot = 0; row = []
    for i, (p, w) in enumerate(zip(REP, wrep)):
        b = 0
        for seg, nc in enumerate((10048, 9952)):
            bits, body = _segment(nc, p, 1000 + 2 * i + seg, True, seg == 1)
            b += 8 * len(_native.debug_lz_block(p, lv, bits, nc, b"1\t12345678\trs1234567\tA\tC\t40\tPASS\t.\tGT\t" if seg == 0 else b"", seg == 1))
        row.append(b / 20000); tot += w * row[-1]
    print("z%d " % lv + " ".join("%5.3f" % x for x in row) + "  mix %.3f (%.2fx incl. header)  %.1fs" % (tot, 4.002 * 8 / tot, time.time() - t0))
