// Fused kernel for X-chromosome rows (K <= 2), in the mask domain like k_fused_auto.
//
// On X a male prints "a\t" (2 bytes, slot 2i only) and a female "a/b\t" (4 bytes) -- pop_factory.py:488-494,
// common/snp.py:109 -- so cells have two widths and the byte-4-back predictor no longer lines up with
// allele slots.  But every even byte of the row body is an allele and every odd byte a separator, so in
// the COMPACTED allele sequence c[k] (slots of the row with the unused second slot of males removed) the
// text is  c[0] S[0] c[1] S[1] ...  with S[k] in {'/', '\t'} fixed by the sex vector alone, and
//     byte 2k   is predicted  <=>  c[k] == c[k-2]          (data:   x = c ^ (c << 2), as on autosomes)
//     byte 2k+1 is predicted  <=>  S[k] == S[k-2]          (static: one mask per population)
// Per span of 64 samples the host precomputes the compaction (a 5-step software PEXT with constant move
// masks), the separator-kind and separator-mismatch masks; the kernel compacts the drawn masks and runs
// the same token / staging / CRC machinery as k_fused_auto (template CRC + 16 table lookups on c).
#pragma once
#include "k_fused.cuh"

namespace dnaf {

struct XSpan {          // static per population, one per 64 samples
    uint32_t used[4];   // slots that print: bit 2i always, bit 2i+1 when sample i is female (existing samples only)
    uint32_t mv[4][5];  // move masks of the 5 compress steps, per 32-slot word
    uint32_t sk[4];     // separator kind over compacted positions: 1 = '/', 0 = '\t'
    uint32_t sm[4];     // separator mismatch: S[k] != S[k-2] (bits 0,1 of a row's first span are set)
    uint32_t len[4];    // compacted bits per word (16..32)
    uint32_t byte_off;  // body byte offset of the span's first sample
    uint32_t pad[3];
};

// compacted allele bits of one span (4 words -> 128-bit little-endian sequence of total length L)
__host__ __device__ __forceinline__ uint32_t compact_span(const uint32_t m[4], const XSpan& xs, uint32_t c[4]) {
    uint32_t cw[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t x = m[w] & xs.used[w];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t t = x & xs.mv[w][i];
            x = (x ^ t) | (t >> (1 << i));
        }
        cw[w] = x;
    }
    const uint32_t l0 = xs.len[0], l1 = xs.len[1], l2 = xs.len[2], l3 = xs.len[3];
    const uint64_t lo = (uint64_t)cw[0] | ((uint64_t)cw[1] << l0);      // l0 + l1 <= 64
    const uint64_t hi = (uint64_t)cw[2] | ((uint64_t)cw[3] << l2);
    const uint32_t sl = l0 + l1;                                         // 0..64
    uint64_t c01 = lo, c23 = 0;
    if (sl < 64) {
        c01 |= hi << sl;
        c23 = sl ? (hi >> (64 - sl)) : 0;
    } else {
        c23 = hi;
    }
    c[0] = (uint32_t)c01; c[1] = (uint32_t)(c01 >> 32); c[2] = (uint32_t)c23; c[3] = (uint32_t)(c23 >> 32);
    return l0 + l1 + l2 + l3;
}

__host__ __device__ __forceinline__ void mask_to_len(uint32_t x[4], int L) {
    if (L < 128) x[3] = L > 96 ? (x[3] & (0xFFFFFFFFu >> (128 - L))) : 0u;
    if (L < 96) x[2] = L > 64 ? (x[2] & (0xFFFFFFFFu >> (96 - L))) : 0u;
    if (L < 64) x[1] = L > 32 ? (x[1] & (0xFFFFFFFFu >> (64 - L))) : 0u;
    if (L < 32) x[0] = L > 0 ? (x[0] & (0xFFFFFFFFu >> (32 - L))) : 0u;
}

struct PairMasks {
    uint32_t xc[4];  // allele mismatches over compacted positions
    uint32_t sm[4];  // separator mismatches
};

__host__ __device__ __forceinline__ PairMasks pair_mismatches(const uint32_t c[4], const uint32_t sm_static[4], uint32_t carry,
                                                              bool has_prev, int L, bool ends_row) {
    PairMasks p;
    p.xc[0] = c[0] ^ ((c[0] << 2) | (carry & 3u));
    p.xc[1] = c[1] ^ ((c[1] << 2) | (c[0] >> 30));
    p.xc[2] = c[2] ^ ((c[2] << 2) | (c[1] >> 30));
    p.xc[3] = c[3] ^ ((c[3] << 2) | (c[2] >> 30));
    for (int w = 0; w < 4; ++w) p.sm[w] = sm_static[w];
    if (!has_prev) { p.xc[0] |= 3u; p.sm[0] |= 3u; }   // block start: nothing 4 bytes back
    mask_to_len(p.xc, L);
    // the row's final separator is '\n': always emitted as a literal by the tail code, not as a pair event
    mask_to_len(p.sm, ends_row ? L - 1 : L);
    return p;
}

// Tokens of a span given as (compacted alleles, separator kinds, mismatch masks).  Sink: gap_tok(gap, lit id)
// fused3(gap, id_a, id_b, id): [match gap>=3 | literals id_a (gap>=1), id_b (gap==2)][literal id]; lit(id); match(len).
template <class Sink>
__host__ __device__ __forceinline__ void tokenize_pairs(const uint32_t c[4], const uint32_t sk[4], const PairMasks& pm, int L,
                                                        bool ends_row, Sink& sink) {
    const int nbytes = 2 * L;
    const int end = ends_row ? nbytes - 1 : nbytes;
    int prev_end = 0;
    auto lit_id_at = [&](int q) {
        const uint32_t bit = (pick4((q & 1) ? sk : c, q >> 6) >> ((q >> 1) & 31)) & 1u;
        return (q & 1) ? (bit ? kLitSlash : kLitTab) : (int)bit;
    };
    // One mismatching byte per iteration (an allele byte before the separator of the same pair), every case
    // through the same straight-line code: [match | 1-2 spelled-out predicted bytes | nothing][literal].
    auto word_of = [&](const uint32_t* a, const uint32_t* b2, int w) { return pick4(a, w) | pick4(b2, w); };
    auto next_word = [&](int from) {  // first word with an event, index >= from; 4 when none
        return (from <= 0 && word_of(pm.xc, pm.sm, 0)) ? 0 : ((from <= 1 && word_of(pm.xc, pm.sm, 1)) ? 1
             : ((from <= 2 && word_of(pm.xc, pm.sm, 2)) ? 2 : ((from <= 3 && word_of(pm.xc, pm.sm, 3)) ? 3 : 4)));
    };
    int cw = next_word(0);
    uint32_t ca = pick4(pm.xc, cw), cs = pick4(pm.sm, cw), cc = pick4(c, cw), ck = pick4(sk, cw);
    while (cw < 4) {
        const uint32_t ce = ca | cs;
#ifdef __CUDA_ARCH__
        const int b = __ffs((int)ce) - 1;
#else
        const int b = __builtin_ctz(ce);
#endif
        const uint32_t is_a = (ca >> b) & 1u;
        ca &= ~(is_a << b);
        cs &= ~((is_a ^ 1u) << b);
        const int p = 2 * (32 * cw + b) + (int)(is_a ^ 1u);
        const int id = is_a ? (int)((cc >> b) & 1u) : (((ck >> b) & 1u) ? kLitSlash : kLitTab);
        const int gap = p - prev_end;
        sink.fused3(gap, lit_id_at(prev_end), lit_id_at(prev_end + 1), id);
        prev_end = p + 1;
        if (!(ca | cs)) {
            cw = next_word(cw + 1);
            ca = pick4(pm.xc, cw);
            cs = pick4(pm.sm, cw);
            cc = pick4(c, cw);
            ck = pick4(sk, cw);
        }
    }
    {
        const int gap = end - prev_end;
        if (gap >= 3) sink.match(gap);
        else
            for (int q = prev_end; q < end; ++q) sink.lit(lit_id_at(q));
    }
    if (ends_row) sink.lit(kLitNl);
}

struct XArgs {
    FusedArgs f;
    const XSpan* xspans;   // [ceil(N/64)]
};

__global__ void __launch_bounds__(kFusedMaxThreads, 3) k_fused_x(const XArgs xa) {
    const FusedArgs& a = xa.f;
    __shared__ FusedSmem s;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const FusedDesc d = a.desc[blockIdx.x];
    const FusedTable* __restrict__ tb = a.tables + d.table;
    const bool has_prefix = d.flags & 1u, ends_row = (d.flags >> 1) & 1u;
    const uint64_t pb = a.nv.pre_off[d.row];
    const uint32_t plen = has_prefix ? (uint32_t)(a.nv.pre_off[d.row + 1] - pb) : 0u;
    const uint32_t body0 = a.sv.xoff[d.cell0];
    const uint32_t n = plen + (a.sv.xoff[d.cell0 + d.ncells] - body0);  // text bytes of this block
    const uint32_t nspans = (d.ncells + 63u) / 64u;

    for (uint32_t i = tid; i < 260; i += nthr) s.len_tok[i] = tb->len_tok[i];
    if (tid < 8) s.lit_tok[tid] = tb->lit[tid];
    for (uint32_t i = tid; i < 132; i += nthr) s.cnt[i] = 0;
    if (tid == 0) { s.crc_acc = 0; s.overflow = 0; }

    // ---- draw this span's 128 slots, force the overrides, compact to the printed alleles
    const uint32_t cs = d.cell0 + 64u * tid;
    int nc = 0;
    if (tid < nspans) nc = (int)min(64u, d.ncells - 64u * tid);
    uint32_t c[4] = {0, 0, 0, 0};
    uint32_t L = 0;
    if (nc > 0) {
        uint32_t m[4] = {0, 0, 0, 0};
        if (a.nv.k[d.row] == 2) {
            const uint32_t thr = a.nv.thr[d.row * 4];
            const uint64_t prow = a.row_base + d.row;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t g = (cs >> 4) + w;
                if (32u * g < 2u * a.sv.n) {
                    const uint32_t slots = 2u * a.sv.n - 32u * g;
                    const uint32_t valid = slots >= 32u ? 0xFFFFFFFFu : ((1u << slots) - 1u);
                    uint32_t p1;
                    draw_group<2>(g, prow, a.k0, a.k1, &thr, valid, m[w], p1);
                }
            }
        }
        for (uint32_t o = 0; o < d.ovr_count; ++o) {
            const uint32_t i = a.osamp[d.ovr_first + o];
            if (i >= cs && i < cs + (uint32_t)nc) {
                const uint32_t j = 2u * (i - cs);
                const uint32_t bit = 3u << (j & 31u);
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if ((j >> 5) == (uint32_t)w) m[w] |= bit;
            }
        }
        L = compact_span(m, xa.xspans[cs >> 6], c);
    }
    // the two alleles before this span (for x = c ^ (c << 2)): top two valid bits of the previous span
    {
        uint32_t top2 = 0;
        if (L >= 2) {
            const uint32_t hiw = pick4(c, (int)((L - 1) >> 5)), b = (L - 1) & 31u;
            const uint32_t prevw = b == 0 ? pick4(c, (int)((L - 2) >> 5)) : hiw;
            top2 = (((hiw >> b) & 1u) << 1) | ((prevw >> ((L - 2) & 31u)) & 1u);
        }
        s.last_bits[tid] = top2;
    }
    __syncthreads();
    const uint32_t carry = tid ? s.last_bits[tid - 1] : 0u;

    // ---- counting sort of the spans by event count (descending)
    uint32_t key = 0, rank_in_bin = 0;
    const XSpan* xs = nc > 0 ? &xa.xspans[cs >> 6] : nullptr;
    const bool my_end_row = ends_row && nc > 0 && 64u * tid + (uint32_t)nc == d.ncells;
    if (nc > 0) {
        const PairMasks pmk = pair_mismatches(c, xs->sm, carry, tid > 0, (int)L, my_end_row);
        key = (__popc(pmk.xc[0]) + __popc(pmk.sm[0]) + __popc(pmk.xc[1]) + __popc(pmk.sm[1]) + __popc(pmk.xc[2]) +
               __popc(pmk.sm[2]) + __popc(pmk.xc[3]) + __popc(pmk.sm[3]) + 1) >> 1;   // mismatching bytes / 2: 0..128
        rank_in_bin = atomicAdd(&s.cnt[key], 1u);
    }
    // ---- CRC32 share of this span: template ^ delta, shifted to the block end
    uint32_t crc = 0;
    if (nc > 0 && (c[0] | c[1] | c[2] | c[3])) {
        uint32_t mm[4] = {c[0], c[1], c[2], c[3]};
        if (L < 128) {  // align the span's end with the table's span end (128-bit left shift by 128 - L)
            const uint32_t sh = 128u - L;
            const uint32_t ws = sh >> 5, bs = sh & 31u;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (ws > (uint32_t)k) { mm[3] = mm[2]; mm[2] = mm[1]; mm[1] = mm[0]; mm[0] = 0; }
            mm[3] = __funnelshift_l(mm[2], mm[3], bs);
            mm[2] = __funnelshift_l(mm[1], mm[2], bs);
            mm[1] = __funnelshift_l(mm[0], mm[1], bs);
            mm[0] = mm[0] << bs;
        }
        uint32_t sp = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int k = 0; k < 4; ++k) sp ^= __ldg(&a.etab[(4 * w + k) * 256 + ((mm[w] >> (8 * k)) & 0xFFu)]);
        const uint32_t span_end = plen + (xs->byte_off - body0) + 2u * L;
        crc = gf2_mulmod(a.xpow8[n - span_end], sp);
    }
    uint32_t pre_tok = 0;
    if (tid < plen) {
        const uint8_t ch = a.nv.prefix[pb + tid];
        pre_tok = tb->pre_lit[ch];
        crc ^= gf2_mulmod(a.xpow8[n - 1u - tid], __ldg(&a.crctab[ch]));
    }
    if (tid == 0) crc ^= d.body_crc ^ gf2_mulmod(a.xpow8[n], 0xFFFFFFFFu);
    crc = warp_xor(crc);
    if ((tid & 31u) == 0 && crc) atomicXor(&s.crc_acc, crc);
    __syncthreads();
    if (tid < 32) {  // bin starts, heaviest spans first (keys 0..128)
        uint32_t c4[5], tot = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t k = 128u - (5u * tid + i);
            c4[i] = (5u * tid + i <= 128u) ? s.cnt[k] : 0u;
            tot += c4[i];
        }
        uint32_t v = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (tid >= (uint32_t)o) v += u;
        }
        uint32_t run = v - tot;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (5u * tid + i <= 128u) s.cnt[128u - (5u * tid + i)] = run;
            run += c4[i];
        }
    }
    __syncthreads();
    if (nc > 0) {
        const uint32_t pos = s.cnt[key] + rank_in_bin;
        s.sm[0][pos] = c[0];
        s.sm[1][pos] = c[1];
        s.sm[2][pos] = c[2];
        s.sm[3][pos] = c[3];
        s.smeta[pos] = tid | (carry << 16) | (L << 20);
    }
    s.span_bits[tid] = 0;
    __syncthreads();

    // ---- pass 1: tokens of the span dealt to this thread, staged privately
    const uint32_t eob = tb->eob;
    FusedStage st{s.len_tok, s.stage + tid, nthr, s.lit_tok[kLit0], s.lit_tok[kLit1], s.lit_tok[kLitSlash],
                  s.lit_tok[kLitTab], s.lit_tok[kLitNl], 0, 0, 0};
    uint32_t pc[4] = {0, 0, 0, 0};
    uint32_t psp = 0;
    int pL = 0;
    bool p_end = false, p_last = false;
    PairMasks ppm;
    const XSpan* pxs = nullptr;
    const bool worker = tid < nspans;
    if (worker) {
        pc[0] = s.sm[0][tid]; pc[1] = s.sm[1][tid]; pc[2] = s.sm[2][tid]; pc[3] = s.sm[3][tid];
        const uint32_t meta = s.smeta[tid];
        psp = meta & 0xFFFFu;
        const uint32_t pcarry = (meta >> 16) & 3u;
        pL = (int)(meta >> 20);
        const uint32_t pnc = min(64u, d.ncells - 64u * psp);
        p_last = 64u * psp + pnc == d.ncells;
        p_end = ends_row && p_last;
        pxs = &xa.xspans[(d.cell0 >> 6) + psp];
        ppm = pair_mismatches(pc, pxs->sm, pcarry, psp > 0, pL, p_end);
        tokenize_pairs(pc, pxs->sk, ppm, pL, p_end, st);
        if (p_last) st.put(eob & 0xFFFFFFu, eob >> 24);
        st.finish();
        if (st.bits() > 32u * kStageWords) s.overflow = 1;
        s.span_bits[psp] = st.bits();
    }
    __syncthreads();
    const uint32_t pre_bits = pre_tok >> 24;
    const uint32_t my_bits = s.span_bits[tid];
    uint32_t pre_off, total_pre, total_span;
    {
        uint32_t v0 = pre_bits, v1 = my_bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u0 = __shfl_up_sync(0xFFFFFFFFu, v0, o);
            const uint32_t u1 = __shfl_up_sync(0xFFFFFFFFu, v1, o);
            if ((tid & 31u) >= (uint32_t)o) { v0 += u0; v1 += u1; }
        }
        if ((tid & 31u) == 31u) { s.warp_pre[tid >> 5] = v0; s.warp_span[tid >> 5] = v1; }
        __syncthreads();
        uint32_t b0 = 0, b1 = 0, t0 = 0, t1 = 0;
        const uint32_t nw = nthr >> 5;
        for (uint32_t wv = 0; wv < nw; ++wv) {
            if (wv < (tid >> 5)) { b0 += s.warp_pre[wv]; b1 += s.warp_span[wv]; }
            t0 += s.warp_pre[wv];
            t1 += s.warp_span[wv];
        }
        pre_off = b0 + v0 - pre_bits;
        total_pre = t0;
        total_span = t1;
        s.span_bits[tid] = b1 + v1 - my_bits;
    }
    const uint32_t hdr_bits = tb->hdr_bits;
    const uint32_t data_bits = hdr_bits + total_pre + total_span;
    const uint32_t payload = (data_bits + 7u) / 8u;
    const uint32_t out_words = (data_bits + 31u) / 32u;
    const bool stored = payload > n + 5u;
    uint8_t* blk = a.slots + (uint64_t)d.slot * kSlot + kSlotLead;
    uint32_t* words = reinterpret_cast<uint32_t*>(blk + 18);

    uint32_t out_payload;
    if (!stored) {
        const uint32_t hdr_words = (hdr_bits + 31u) / 32u;
        for (uint32_t i = tid; i < out_words + 1u; i += nthr) words[i] = i < hdr_words ? tb->hdr[i] : 0u;
        __syncthreads();
        const bool overflow = s.overflow != 0;
        if (pre_bits) {
            const uint32_t pos = hdr_bits + pre_off, wi = pos >> 5, sh = pos & 31u, v = pre_tok & 0xFFFFFFu;
            atomicOr(&words[wi], v << sh);
            if (sh + pre_bits > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
        }
        if (worker) {
            const uint32_t dst = hdr_bits + total_pre + s.span_bits[psp];
            if (!overflow) {
                const uint32_t nb = st.bits();
                const uint32_t sh = dst & 31u;
                const uint32_t nsrc = (nb + 31u) / 32u;
                const uint32_t ndst = (sh + nb + 31u) / 32u;
                uint32_t* o = words + (dst >> 5);
                uint32_t prev = 0;
                for (uint32_t k = 0; k < ndst; ++k) {
                    const uint32_t cur = k < nsrc ? s.stage[k * nthr + tid] : 0u;
                    const uint32_t v = __funnelshift_l(prev, cur, sh);
                    if (k == 0 || k == ndst - 1) atomicOr(&o[k], v);
                    else o[k] = v;
                    prev = cur;
                }
            } else {
                FusedEmit em{s.len_tok, st.lit0, st.lit1, st.lit_slash, st.lit_tab, st.lit_nl, words, dst};
                tokenize_pairs(pc, pxs->sk, ppm, pL, p_end, em);
                if (p_last) em.put(eob & 0xFFFFFFu, eob >> 24);
            }
        }
        out_payload = payload;
    } else {
        // stored deflate block: format the text itself (rare safety net)
        if (tid == 0) {
            blk[18] = 1;
            blk[19] = (uint8_t)n; blk[20] = (uint8_t)(n >> 8);
            blk[21] = (uint8_t)~n; blk[22] = (uint8_t)((~n) >> 8);
        }
        if (tid < plen) blk[23 + tid] = a.nv.prefix[pb + tid];
        if (nc > 0) {
            uint8_t* p = blk + 23 + plen + (xs->byte_off - body0);
            for (uint32_t k = 0; k < L; ++k) {
                p[2 * k] = '0' + ((pick4(c, (int)(k >> 5)) >> (k & 31u)) & 1u);
                const bool slash = (pick4(xs->sk, (int)(k >> 5)) >> (k & 31u)) & 1u;
                p[2 * k + 1] = (my_end_row && k + 1 == L) ? '\n' : (slash ? '/' : '\t');
            }
        }
        out_payload = n + 5u;
    }
    __syncthreads();
    if (tid < 26) {
        const uint32_t crc32 = ~s.crc_acc;
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

}  // namespace dnaf
