// k_lz: the autosome-row kernel of the higher compression tiers (-z 4..9).  Same job, block geometry, draws, CRC
// and output slots as k_auto (k_auto.cuh); what changes is the deflate parse.  k_auto predicts every byte by the byte
// 4 back ("P4"); here a block is parsed with real LZ77 matches over its whole 32 KiB window -- in the ALLELE-BIT
// domain: the text of an autosome row is  a0 '/' a1 '\t' ...  with a_j in {'0','1'}, so the block is a string of
// allele bits with fixed separators, and a deflate match (length 2k [+1], distance 4D bytes) is a run of k equal bits
// 2D positions back.  Comparing 64 alleles (128 text bytes) costs one 64-bit XOR.
//
// Reference behaviour restated: Bio.bgzf.BgzfWriter(compresslevel=z) -- pop_factory.py:403 passes the -z value
// (default 6, pop_factory.py:656-658) to zlib.compressobj(level, DEFLATED, -15) for every 64 KiB block.  zlib's
// levels are hash chains of growing depth (4, 8, 16, 32, 128, 4096 candidates at levels 4, 5, 6, 7, 8, 9; lazy
// evaluation from level 4 on); the tiers here follow that ladder (LzCfg).  Compressed bytes are outside the parity
// contract; the decompressed text is identical at every level.
//
// Match finder (deterministic, built per block in shared memory):
//   * key of allele position a = its next L alleles (L <= 10, per MAF bucket) and a's parity (distances are even)
//   * a block is cut into REGIONS of 4096 alleles = the 32 spans of one warp.  Every warp links the positions of its
//     region into per-key chains in position order, 32 consecutive positions per step: every lane reads
//     head[region][key] (the last occurrence before this step) into prev[a], then the step's positions go into the
//     heads with atomicMax.  Positions of one step that share a key all point at the same predecessor and only the
//     last of them is reachable from later steps -- which thins exactly the over-populated keys (the inside of a long
//     run of reference alleles) and loses ~1 % of the entries elsewhere.  max() does not depend on the order the
//     lanes arrive in and reads and updates are separated by a warp barrier: the same block always gives the same
//     chains.  (A first version linked equal keys inside a step exactly with __match_any_sync; MATCH.ANY takes time
//     proportional to the number of distinct keys in the warp and made this phase 4x the cost of everything else.)
//   * a lookup at position s walks prev[] in its own region, then enters the regions before it through their heads
//     (all of a previous region precedes s): candidates come nearest first, like zlib's hash chains, at most
//     `chain` of them; plus the two near distances 4 and 8 bytes that need no table.
//   * one thread parses one span (64 cells) sequentially, greedy longest match (ties: nearest), optional one-step
//     lazy evaluation; tokens never cross spans, sources may lie anywhere earlier in the block.
#pragma once
#include "k_auto.cuh"

namespace dnaf {

constexpr int kLStage = 14;            // staged words per span before the block falls back to direct emission
constexpr uint32_t kLzNone = 0xFFFFu;      // prev[] entry / head value of "no earlier position" (heads hold 0 = none, position + 1 otherwise)
constexpr uint32_t kLzRegion = 4096;   // alleles per region (one warp's 32 spans)
constexpr uint32_t kLzMaxKey = 9;      // alleles in a key (head tables have 2^(key+1) 32-bit entries per region)
constexpr uint32_t kLzMaxDist = 16384; // alleles = 32768 bytes

struct LzCfg {
    uint32_t chain;   // far candidates examined per lookup
    uint32_t lazy;    // 0 / 1: one-step lazy evaluation
    uint32_t key;     // alleles per key (<= kLzMaxKey)
    uint32_t nice;    // alleles: a match this long ends the search
};

// -z level -> parse parameters (zlib's own ladder: max_chain 16 / 32 / 128 / 256 / 1024 / 4096 for 4 .. 9, quartered
// here because every chain entry is already a >= 2*key byte match)
__host__ __device__ inline LzCfg lz_cfg(int level, uint32_t key) {
    LzCfg c;
    c.key = key;
    c.chain = level <= 4 ? 1u : level == 5 ? 4u : level == 6 ? 8u : level == 7 ? 16u : level == 8 ? 32u : 128u;
    c.lazy = level >= 6 ? 1u : 0u;
    c.nice = level <= 4 ? 16u : level == 5 ? 32u : level <= 7 ? 64u : 128u;   // zlib: 16, 32, 128, 128, 258, 258 bytes
    return c;
}

// Static code tables of one (MAF bucket, with/without prefix, tier).
struct LzTable {
    // ---- the first kLTabWords words are copied to shared memory by every block (one bulk copy)
    uint32_t len_tok[260];   // [len 3..258]: length code | extra << code bits | total bits << 24 (no distance part)
    uint32_t lit[8];         // cell literals by id: code | bits << 24
    uint32_t eob;
    uint32_t hdr_bits;
    uint32_t dist_tok[32];   // [distance symbol]: code | bits << 24
    uint32_t key_alleles;
    uint32_t pad;
    uint32_t hdr[96];        // serialized dynamic-block header
    // ----
    uint32_t pre_lit[256];   // literal codes of prefix bytes
};
constexpr uint32_t kLTabWords = 260 + 8 + 2 + 32 + 2 + 96;   // 400 words = 100 x 16 bytes
static_assert(kLTabWords % 4 == 0 && sizeof(LzTable) % 16 == 0, "LzTable must copy in 16-byte units");

__host__ __device__ __forceinline__ uint32_t lz_fsr(uint32_t lo, uint32_t hi, uint32_t sh) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
__host__ __device__ __forceinline__ uint32_t lz_ctz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffsll((long long)x) - 1u;
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

// distance (bytes, 1..32768) -> deflate distance symbol, number of extra bits, extra value
__host__ __device__ __forceinline__ void lz_dist_sym(uint32_t d, uint32_t& sym, uint32_t& eb, uint32_t& ev) {
    if (d <= 4u) { sym = d - 1u; eb = 0; ev = 0; return; }
    const uint32_t x = d - 1u;
#ifdef __CUDA_ARCH__
    const uint32_t hb = 31u - (uint32_t)__clz((int)x);
#else
    const uint32_t hb = 31u - (uint32_t)__builtin_clz(x);
#endif
    sym = 2u * hb + ((x >> (hb - 1u)) & 1u);
    eb = hb - 1u;
    ev = x & ((1u << eb) - 1u);
}

// ---- the span grammar of the LZ tiers (host table builder, device fast and slow paths all run this function) ----
// Block coordinates: allele a of the block is text byte 2a, its separator byte 2a+1 ('/' after even a, '\t' after
// odd a); byte -1 is the separator before the block's first allele (the prefix's last tab, or the tab that ended
// the previous segment -- the block owns it).  A span of nc cells starting at allele a0 owns bytes 2*a0 - 1 ..
// 2*(a0 + 2nc) - 2 (its last allele): the separator BEFORE its first cell, not the one after its last.
//   first span of a block: ['\t' unless the block starts a row] a0 '/' a1 as literals (no history yet)
//   then greedy: at byte p, s = first allele at or after p; candidates = distances 4, 8 and the key chain of s;
//   k = most alleles equal to those 2D back (up to the span's end); the match covers bytes p .. 2(s+k)-1, without
//   that last separator when it is not the span's; >= 3 bytes -> match, else one literal byte.
//   ['\n' if the span ends the row] [EOB if it ends the block]
// Mem: word(i) = allele bits 32i .. 32i+31 of the block (zero past the end), prev(a), head(region, key).
// Sink: lit(id), match(len bytes, dist bytes), eob().
template <class Mem>
__host__ __device__ __forceinline__ uint32_t lz_win32(const Mem& mem, uint32_t a) {
    const uint32_t w = a >> 5;
    return lz_fsr(mem.word(w), mem.word(w + 1u), a & 31u);
}

// The parse is ONE flat loop: every trip compares one 32-allele chunk of one candidate.  All lanes of a warp run the
// same instructions whatever token or candidate each of them is at -- nested per-token / per-candidate loops made
// the warp pay for every lane's trip counts in turn (18 of 32 threads active in the first version of this kernel).
//   stage 0: distance 4 bytes (j = s-2)   stage 1: distance 8 (j = s-4)   stage 3: the key chain of s
// nice: a match of that many alleles ends the search (zlib's nice_length).
template <class Mem, class Sink>
__host__ __device__ __forceinline__ void lz_span_tokens(const Mem& mem, uint32_t a0, int nc, bool first_in_block, bool starts_row,
                                                        bool ends_row, bool ends_block, uint32_t nall, const LzCfg cfg, Sink& sink) {
    const uint32_t aend = a0 + 2u * (uint32_t)nc;   // one past the span's last allele
    auto bit = [&](uint32_t a) { return (int)((mem.word(a >> 5) >> (a & 31u)) & 1u); };
    int p = 2 * (int)a0 - 1;                        // next byte to emit (block coordinates)
    if (first_in_block) {
        if (!starts_row) sink.lit(kLitTab);
        sink.lit(bit(a0));
        sink.lit(kLitSlash);
        sink.lit(bit(a0 + 1u));
        p = 2 * (int)a0 + 3;
    }
    const int pend = 2 * (int)aend - 1;             // one past the span's last byte
    const uint32_t kmask = (1u << cfg.key) - 1u;
    // lookup state
    uint32_t s = 0, limit = 0, ws0 = 0, stage = 0, j = 0, off = 0, left = 0, reg = 0, key = 0;
    uint32_t best_k = 0, best_da = 0;
    // lazy evaluation: the finished lookup at the byte before (held while the next allele is examined)
    uint32_t held_len = 0, held_da = 0;
    bool held = false, fresh = true;
    while (p < pend) {
        if (fresh) {   // set up the lookup at byte p
            s = (uint32_t)(p + 1) >> 1;
            limit = aend - s;                        // s < aend here: the span's last byte is an allele
            ws0 = lz_win32(mem, s);
            best_k = 0;
            best_da = 0;
            off = 0;
            stage = s >= 2u ? 0u : 3u;
            j = s - 2u;
            left = (s + cfg.key <= nall) ? cfg.chain : 0u;
            reg = s / kLzRegion;
            key = (ws0 & kmask) | ((s & 1u) << cfg.key);
            fresh = false;
            if (stage == 3u) j = left ? mem.prev(s) : kLzNone;
        }
        bool done = false;
        if (stage == 3u) {
            // resolve the next chain entry: own region first, then the regions before it through their heads
            while (left && j == kLzNone && reg) {
                --reg;
                j = mem.head(reg, key);
            }
            if (!left || j == kLzNone || s - j > kLzMaxDist) done = true;
        }
        if (!done) {
            // one 32-allele chunk of candidate j
            const uint32_t a = off ? lz_win32(mem, s + off) : ws0;
            const uint32_t x = a ^ lz_win32(mem, j + off);
            const uint32_t room = limit - off;
#ifdef __CUDA_ARCH__
            uint32_t c = x ? (uint32_t)__ffs((int)x) - 1u : 32u;
#else
            uint32_t c = x ? (uint32_t)__builtin_ctz(x) : 32u;
#endif
            c = c < room ? c : room;
            if (c == 32u && room > 32u) {
                off += 32u;                          // same candidate, next chunk
            } else {
                const uint32_t k = off + c;
                if (k > best_k) {
                    best_k = k;
                    best_da = s - j;
                }
                off = 0;
                // next candidate
                if (best_k >= limit || best_k >= cfg.nice) done = true;
                else if (stage == 0u) {
                    if (s >= 4u) { stage = 1u; j = s - 4u; }
                    else { stage = 3u; j = left ? mem.prev(s) : kLzNone; }
                } else if (stage == 1u) {
                    stage = 3u;
                    j = left ? mem.prev(s) : kLzNone;
                } else {
                    --left;
                    j = mem.prev(j);
                }
            }
        }
        if (done) {
            const uint32_t odd = (uint32_t)p & 1u;
            // bytes p .. 2(s+k)-1, minus the separator after the span's last allele
            const int len = 2 * (int)(s + best_k) - p - (s + best_k == aend ? 1 : 0);
            if (held) {
                // p is the byte after the held lookup's: take the literal + this match if that is longer
                held = false;
                if (len > (int)held_len + 1) {
                    sink.lit(bit(s - 1u));
                    sink.match(len, 2 * (int)best_da);
                    p += len;
                } else {
                    sink.match((int)held_len, 2 * (int)held_da);
                    p += (int)held_len - 1;
                }
            } else if (len >= 3) {
                if (cfg.lazy && !odd && s + 1u < aend && s + best_k < aend && best_k < cfg.nice) {
                    held = true;                     // would a literal now buy a longer match from the next byte on?
                    held_len = (uint32_t)len;
                    held_da = best_da;
                    p += 1;
                } else {
                    sink.match(len, 2 * (int)best_da);
                    p += len;
                }
            } else {
                if (odd) sink.lit((s & 1u) ? kLitSlash : kLitTab);   // separator before allele s
                else sink.lit(bit(s));
                p += 1;
            }
            fresh = true;
        }
    }
    if (ends_row) sink.lit(kLitNl);
    if (ends_block) sink.eob();
}

#ifdef __CUDACC__
struct LzArgs {
    AutoArgs a;
    const LzTable* tables;   // [2 * bucket + (segment 0 ? 0 : 1)]
    uint32_t chain, lazy, nice;
};

struct LzMemDev {
    const uint32_t* bits;
    const uint16_t* prv;
    const uint32_t* hd;      // [region][2^hbits]: last position of the key in the region + 1, 0 = none
    uint32_t hbits;
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return bits[i]; }
    __device__ __forceinline__ uint32_t prev(uint32_t a) const { return prv[a]; }
    __device__ __forceinline__ uint32_t head(uint32_t reg, uint32_t key) const { return (hd[(reg << hbits) + key] - 1u) & 0xFFFFu; }
};

// token-level adapters of the bit sinks (AStageT / AEmit, k_auto.cuh)
template <class Bits>
struct LzTokSink {
    Bits& b;
    const uint32_t* len_tok;
    const uint32_t* lits;
    const uint32_t* dist_tok;
    uint32_t eob_tok;
    __device__ __forceinline__ void lit(int id) { const uint32_t t = lits[id]; b.put64(t & 0xFFFFFFu, 0u, t >> 24); }
    __device__ __forceinline__ void match(int len, int dist) {
        uint32_t sym, eb, ev;
        lz_dist_sym((uint32_t)dist, sym, eb, ev);
        const uint32_t dt = dist_tok[sym];
        const uint32_t dn = dt >> 24;
        b.put_tok_code(len_tok[len], (dt & 0xFFFFFFu) | (ev << dn), 0u, dn + eb);   // <= 20 + 28 bits
    }
    __device__ __forceinline__ void eob() { b.put64(eob_tok & 0xFFFFFFu, 0u, eob_tok >> 24); }
};

// dynamic shared memory carve-up
__host__ __device__ inline uint32_t lz_smem_bytes(uint32_t nthr, uint32_t hbits) {
    return kLTabWords * 4u + (4u * nthr + 8u) * 4u + (nthr / 32u) * (4u << hbits) + nthr * 256u +
           ((uint32_t)(kLStage + 2) * nthr + 24u) * 4u + 16u;
}

__global__ void __launch_bounds__(256) k_lz(const LzArgs la) {
    const AutoArgs& a = la.a;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem_raw);     // LzTable's first kLTabWords words
    uint32_t* s_bits = s_tab + kLTabWords;                       // allele bits of the block + 8 zero guard words
    uint32_t* s_stage = s_bits + 4u * nthr + 8u;
    uint32_t* s_misc = s_stage + (kLStage + 2) * nthr;           // as in k_auto: [0..7] warp bits, [8..9] prefix bits, [10] const crc, [16] crc, [17] overflow, [22..23] mbarrier
    uint16_t* s_prev = reinterpret_cast<uint16_t*>(s_misc + 24u + 4u);
    const uint32_t* s_len = s_tab;
    const uint32_t* s_lits = s_tab + 260;
    const uint32_t* s_dist = s_tab + 270;
    const uint32_t* s_hdr = s_tab + 304;

    FusedDesc d;
    if (a.desc) {
        d = a.desc[blockIdx.x];
    } else {
        const uint32_t rl = a.nseg_magic ? __umulhi(blockIdx.x, a.nseg_magic) : blockIdx.x, sg = blockIdx.x - rl * a.nseg;
        d.row = a.row0 + rl;
        d.cell0 = __ldg(&a.seginfo[3u * sg]);
        d.ncells = __ldg(&a.seginfo[3u * sg + 1u]);
        d.body_crc = __ldg(&a.seginfo[3u * sg + 2u]);
        d.slot = blockIdx.x;
        d.flags = (sg == 0u ? 1u : 0u) | (sg + 1u == a.nseg ? 2u : 0u);
        d.ovr_first = __ldg(&a.ovr_first[d.row]);
        d.ovr_count = __ldg(&a.ovr_first[d.row + 1]) - d.ovr_first;
        d.table = 2u * __ldg(&a.bucket[d.row]) + (sg == 0u ? 0u : 1u);
    }
    const LzTable* __restrict__ tb = la.tables + d.table;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_misc + 22);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(kLTabWords * 4u) : "memory");
        bulk_g2s(s_tab, tb->len_tok, kLTabWords * 4u, mbar);
    }
    if (tid == 6) { s_misc[16] = 0; s_misc[17] = 0; }
    const bool starts_row = d.flags & 1u, ends_row = (d.flags >> 1) & 1u;
    const uint64_t pb = a.nv.pre_off[d.row];
    const uint32_t plen = starts_row ? (uint32_t)(a.nv.pre_off[d.row + 1] - pb) : 0u;
    const uint32_t lead = starts_row ? plen : 1u;
    const uint32_t n = lead + 4u * d.ncells - (ends_row ? 0u : 1u);
    const uint32_t nspans = (d.ncells + 63u) / 64u;
    const uint32_t nall = 2u * d.ncells;

    // ---- draws, overrides, CRC share: exactly k_auto's (shared functions)
    const uint32_t cs = d.cell0 + 64u * tid;
    int nc = 0;
    if (tid < nspans) nc = (int)min(64u, d.ncells - 64u * tid);
    uint32_t m[4] = {0, 0, 0, 0};
    auto_draw_span(a, d, cs, nc, m);
    uint32_t crc = auto_span_crc(a, d, m, nc, tid, nspans);
    uint32_t pre_tok = 0;
    if (tid < plen) pre_tok = __ldg(&tb->pre_lit[a.nv.prefix[pb + tid]]);
    if (tid == 0) {
        uint32_t c0 = d.body_crc ^ __ldg(&a.xinit[n]);
        if (starts_row) c0 ^= mul_tab(ends_row ? a.mpre + 1024 : a.mpre, __ldg(&a.pre_crc[d.row]));
        s_misc[10] = c0;
    }
    crc = warp_xor(crc);

    // ---- allele bits of the block -> shared memory (bits past the block's last cell are zero)
    {
        uint32_t mm[4] = {m[0], m[1], m[2], m[3]};
        if (nc < 64) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int lo = 2 * nc - 32 * w;
                mm[w] &= lo >= 32 ? 0xFFFFFFFFu : (lo <= 0 ? 0u : ((1u << lo) - 1u));
            }
        }
        *reinterpret_cast<uint4*>(s_bits + 4u * tid) = make_uint4(mm[0], mm[1], mm[2], mm[3]);
        if (tid < 8) s_bits[4u * nthr + tid] = 0u;
    }
    const uint32_t key_alleles = __ldg(&tb->key_alleles);
    const uint32_t hbits = key_alleles + 1u;
    uint32_t* s_head = reinterpret_cast<uint32_t*>(s_prev + 128u * nthr);
    {   // heads of this warp's region start empty
        uint32_t* h32 = s_head + ((size_t)wid << hbits);
        for (uint32_t i = lane; i < (1u << hbits); i += 32u) h32[i] = 0u;
    }
    __syncthreads();        // bits complete; also orders the mbarrier's initialisation before the waits
    if (lane == 0 && crc) atomicXor(&s_misc[16], crc);

    // ---- chains of this warp's region, 32 consecutive positions per step (see the header comment).  Positions whose
    // key runs past the block's end are linked too (their keys hold guard zeros): no lookup can reach them, every
    // candidate lies before a position whose own key fits.
    {
        const uint32_t rbase = kLzRegion * wid;                  // first allele of the region
        const uint32_t rend = min(nall, rbase + kLzRegion);
        uint32_t* hd = s_head + ((size_t)wid << hbits);
        const uint32_t kmask = (1u << key_alleles) - 1u, par = (lane & 1u) << key_alleles;
        uint32_t w0 = s_bits[rbase >> 5];
        for (uint32_t pos = rbase; pos < rend; pos += 32u) {
            const uint32_t w1 = s_bits[(pos >> 5) + 1u];
            const uint32_t key = (__funnelshift_r(w0, w1, lane) & kmask) | par;
            w0 = w1;
            const uint32_t before = hd[key];                     // last occurrence in the steps before this one (+ 1)
            s_prev[pos + lane] = (uint16_t)(before - 1u);        // 0 -> kLzNone
            __syncwarp();
            atomicMax(&hd[key], pos + lane + 1u);
            __syncwarp();
        }
    }
    mbar_wait(mbar, 0u);    // the code tables have landed
    __syncthreads();        // every region's chains are complete

    // ---- pass 1: this span's tokens, staged privately
    const uint32_t eob = s_tab[268];
    const bool worker = nc > 0;
    const bool p_last = worker && 64u * tid + (uint32_t)nc == d.ncells;
    const bool p_end = ends_row && p_last;
    const LzCfg cfg{la.chain, la.lazy, key_alleles, la.nice};
    const LzMemDev mem{s_bits, s_prev, s_head, hbits};
    AStageT<kLStage> st{s_stage + tid, nthr, 0u, 0u, 0u};
    if (worker) {
        LzTokSink<AStageT<kLStage>> ts{st, s_len, s_lits, s_dist, eob};
        lz_span_tokens(mem, 128u * tid, nc, tid == 0, starts_row, p_end, p_last, nall, cfg, ts);
        if (st.bits() > 32u * kLStage) s_misc[17] = 1;
    }
    // ---- exclusive scans over the CTA: prefix literal bits (warps 0-1) and span bits
    const uint32_t pre_bits = pre_tok >> 24;
    const uint32_t my_bits = worker ? st.bits() : 0u;
    uint32_t v0 = pre_bits, v1 = my_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u1 = __shfl_up_sync(0xFFFFFFFFu, v1, o);
        if (lane >= (uint32_t)o) v1 += u1;
    }
    if (wid < 2) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u0 = __shfl_up_sync(0xFFFFFFFFu, v0, o);
            if (lane >= (uint32_t)o) v0 += u0;
        }
    }
    if (lane == 31u) {
        s_misc[wid] = v1;
        if (wid < 2) s_misc[8 + wid] = v0;
    }
    __syncthreads();
    uint32_t span_off = v1 - my_bits, total_span = 0;
    const uint32_t nw = nthr >> 5;
    for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t t = s_misc[w];
        if (w < wid) span_off += t;
        total_span += t;
    }
    const uint32_t total_pre = plen ? s_misc[8] + s_misc[9] : 0u;
    const uint32_t pre_off = v0 - pre_bits + (wid == 1 ? s_misc[8] : 0u);
    const uint32_t hdr_bits = s_tab[269];
    const uint32_t data_bits = hdr_bits + total_pre + total_span;
    const uint32_t payload = (data_bits + 7u) / 8u;
    const uint32_t out_words = (data_bits + 31u) / 32u;
    const bool stored = payload > n + 5u;
    uint8_t* blk = a.slots + (uint64_t)d.slot * a.slot_stride + kSlotLead;
    uint32_t* words = reinterpret_cast<uint32_t*>(blk + 18);

    uint32_t out_payload;
    if (!stored) {
        const uint32_t hdr_words = (hdr_bits + 31u) / 32u;
        const uint32_t hw4 = (hdr_words + 3u) & ~3u;   // <= 96 <= nthr is NOT guaranteed: loop
        for (uint32_t i = tid; i < hw4; i += nthr) words[i] = i < hdr_words ? s_hdr[i] : 0u;
        const bool overflow = s_misc[17] != 0;
        const uint32_t ppos = hdr_bits + pre_off;
        const uint32_t dst = hdr_bits + total_pre + span_off;
        if (overflow) {
            uint4* w4 = reinterpret_cast<uint4*>(words + hw4);
            const uint32_t n4 = out_words + 2u > hw4 ? (out_words + 2u - hw4 + 3u) / 4u : 0u;
            for (uint32_t i = tid; i < n4; i += nthr) w4[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            if (pre_bits) {
                const uint32_t wi = ppos >> 5;
                if (wi >= hw4) words[wi] = 0u;
                if (wi + 1u >= hw4) words[wi + 1u] = 0u;
            }
            if (worker) {
                const uint32_t first = dst >> 5, last = (dst + my_bits - 1u) >> 5;
                if (first >= hw4) words[first] = 0u;
                if (last >= hw4) words[last] = 0u;
            }
        }
        __syncthreads();
        if (pre_bits) {
            const uint32_t wi = ppos >> 5, sh = ppos & 31u, v = pre_tok & 0xFFFFFFu;
            atomicOr(&words[wi], v << sh);
            if (sh + pre_bits > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
        }
        if (worker) {
            if (!overflow) {
                const uint32_t sh = dst & 31u;
                const uint32_t nsrc = (my_bits + 31u) / 32u;
                const uint32_t ndst = (sh + my_bits + 31u) / 32u;
                uint32_t* o = words + (dst >> 5);
                const uint32_t* sp = s_stage + tid;
                uint32_t prev = my_bits ? sp[0] : 0u;
                atomicOr(&o[0], prev << sh);
                uint32_t k = 1;
                for (; k + 1u < ndst; ++k) {
                    const uint32_t cur = sp[k * nthr];
                    o[k] = __funnelshift_l(prev, cur, sh);
                    prev = cur;
                }
                if (k < ndst) {
                    const uint32_t cur = k < nsrc ? sp[k * nthr] : 0u;
                    atomicOr(&o[k], __funnelshift_l(prev, cur, sh));
                }
            } else {
                AEmit em{words, dst};
                LzTokSink<AEmit> ts{em, s_len, s_lits, s_dist, eob};
                lz_span_tokens(mem, 128u * tid, nc, tid == 0, starts_row, p_end, p_last, nall, cfg, ts);
            }
        }
        out_payload = payload;
    } else {
        // stored deflate block (rare safety net), as in k_auto
        if (tid == 0) {
            blk[18] = 1;
            blk[19] = (uint8_t)n; blk[20] = (uint8_t)(n >> 8);
            blk[21] = (uint8_t)~n; blk[22] = (uint8_t)((~n) >> 8);
            if (!starts_row) blk[23] = '\t';
        }
        if (tid < plen) blk[23 + tid] = a.nv.prefix[pb + tid];
        if (nc > 0) {
            uint8_t* p = blk + 23 + lead + 256u * tid;
            for (int c = 0; c < nc; ++c) {
                const uint32_t bits = (pick4(m, c >> 4) >> (2 * (c & 15))) & 3u;
                p[4 * c] = '0' + (bits & 1u);
                p[4 * c + 1] = '/';
                p[4 * c + 2] = '0' + (bits >> 1);
                if (!(p_last && c == nc - 1)) p[4 * c + 3] = '\t';
                else if (ends_row) p[4 * c + 3] = '\n';
            }
        }
        out_payload = n + 5u;
    }
    __syncthreads();
    if (tid < 26) {
        uint32_t delta = s_misc[16];
        if (ends_row) delta = __ldg(&a.crctab[delta & 0xFFu]) ^ (delta >> 8);
        const uint32_t crc32 = ~(delta ^ s_misc[10]);
        const uint32_t bsize = out_payload + 25u;
        if (tid < 16) {
            const uint8_t head[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00};
            blk[tid] = head[tid];
        } else if (tid < 18) {
            blk[tid] = (uint8_t)(bsize >> (8 * (tid - 16)));
        } else if (tid < 22) {
            blk[18 + out_payload + (tid - 18)] = (uint8_t)(crc32 >> (8 * (tid - 18)));
        } else {
            blk[18 + out_payload + (tid - 18)] = (uint8_t)(n >> (8 * (tid - 22)));
        }
        if (tid == 0) {
            a.sizes[d.slot] = out_payload + 26u;
            a.crcs[d.slot] = crc32;
        }
    }
}
#endif  // __CUDACC__

}  // namespace dnaf
