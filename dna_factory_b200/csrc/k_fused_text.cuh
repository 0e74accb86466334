// The general fused kernel: any chromosome class (X with its 2/4-byte cells, Y with '.', MT, and autosome
// rows with more than two alleles), one CTA per BGZF block, no HBM round trip for genotypes or text.
// Unlike k_fused_auto it does materialise the text -- but only in shared memory: thread t draws the
// alleles of the samples that overlap its 256-byte span of the row body, writes their bytes into its
// private shared-memory window, then tokenises that window word-wise with the same P4 rule, static
// Huffman tables per (class, MAF bucket), and a slicing-by-4 CRC32 merged with the GF(2) shift operator.
//
// Reference behaviour restated: pop_factory.py:471-508 (row loop incl. the Y/female '.' at :481-484, the
// haploid branch :488-490, forced minors :495-499) and BgzfWriter framing (call site pop_factory.py:449).
#pragma once
#include "k_fused.cuh"

namespace dnaf {

constexpr int kTextStageWords = 40;   // per-thread staging (1280 bits) before the slow path is taken
constexpr int kTextStride = 65;       // words per thread window: 64 + 1 pad (bank-conflict free)

// ---- word-wise P4 tokeniser over a span of bytes; shared with the host's table builder ----
// wat(k) returns the k-th little-endian 32-bit word of the span; prev is the word just before the span.
template <class WordAt, class Sink>
__host__ __device__ __forceinline__ void tokenize_words(WordAt wat, int nbytes, bool has_prev, uint32_t prev, Sink& sink) {
    int run = 0;
    auto byte_at = [&](int p) { return (uint8_t)(wat(p >> 2) >> (8 * (p & 3))); };
    auto flush = [&](int pos) {  // `run` predicted bytes end at pos
        if (run >= 3) {
            sink.match(run);
        } else {
            for (int q = pos - run; q < pos; ++q) sink.lit_byte(byte_at(q));
        }
        run = 0;
    };
    const int nw = (nbytes + 3) >> 2;
    for (int k = 0; k < nw; ++k) {
        const uint32_t w = wat(k);
        const int nb = nbytes - 4 * k < 4 ? nbytes - 4 * k : 4;
        const uint32_t x = has_prev ? (w ^ prev) : 0xFFFFFFFFu;
        if (x == 0 && nb == 4) {
            run += 4;
        } else {
            for (int j = 0; j < nb; ++j) {
                if (((x >> (8 * j)) & 0xFFu) == 0) {
                    ++run;
                } else {
                    flush(4 * k + j);
                    sink.lit_byte((uint8_t)(w >> (8 * j)));
                }
            }
        }
        prev = w;
        has_prev = true;
    }
    flush(nbytes);
}

struct TextDesc {
    uint64_t row;
    uint32_t byte0;      // first body byte of the segment (multiple of 256)
    uint32_t nbytes;     // body bytes in the segment
    uint32_t slot;
    uint32_t flags;      // bit0: has prefix, bit1: ends the row
    uint32_t ovr_first, ovr_count;
    uint32_t table;
    uint32_t pad;
};

struct TextStage {
    const uint32_t* len_tok;
    const uint32_t* lit_all;
    uint32_t* stage;
    uint32_t stride;
    uint32_t wi, nacc, bits;
    uint64_t acc;
    __device__ void put(uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << nacc;
        nacc += n;
        bits += n;
        if (nacc >= 32) {
            if (wi < (uint32_t)kTextStageWords) stage[wi * stride] = (uint32_t)acc;
            ++wi;
            acc >>= 32;
            nacc -= 32;
        }
    }
    __device__ void finish() {
        if (nacc && wi < (uint32_t)kTextStageWords) stage[wi * stride] = (uint32_t)acc;
    }
    __device__ void lit_byte(uint8_t b) { put(lit_all[b] & 0xFFFFFFu, lit_all[b] >> 24); }
    __device__ void match(int len) { put(len_tok[len] & 0xFFFFFFu, len_tok[len] >> 24); }
};

struct TextEmit {
    const uint32_t* len_tok;
    const uint32_t* lit_all;
    uint32_t* words;
    uint32_t pos;
    __device__ void put(uint32_t v, uint32_t n) {
        const uint32_t wi = pos >> 5, sh = pos & 31u;
        atomicOr(&words[wi], v << sh);
        if (sh + n > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
        pos += n;
    }
    __device__ void lit_byte(uint8_t b) { put(lit_all[b] & 0xFFFFFFu, lit_all[b] >> 24); }
    __device__ void match(int len) { put(len_tok[len] & 0xFFFFFFu, len_tok[len] >> 24); }
};

struct TextSmem {
    uint32_t text[kTextStride * kFusedMaxThreads];
    uint32_t stage[kTextStageWords * kFusedMaxThreads];
    uint32_t len_tok[260];
    uint32_t lit_all[256];
    uint32_t warp_pre[8], warp_span[8];
    uint32_t crc_acc;
    uint32_t overflow;
};

struct TextArgs {
    SampleView sv;
    SnpView nv;
    const TextDesc* desc;
    const FusedTable* tables;
    const uint32_t* crc4;       // [4][256] slicing-by-4 tables
    const uint32_t* xpow8;
    const uint32_t* osamp;
    const uint32_t* xspan;      // X rows: sample that holds body byte 256*k
    uint64_t row_base;
    uint32_t k0, k1;
    uint8_t* slots;
    uint32_t slot_stride;      // bytes per output slot of this pass (>= the longest block's text + 96)
    uint32_t* sizes;
    uint32_t* crcs;
};

__global__ void __launch_bounds__(kFusedMaxThreads, 2) k_fused_text(const TextArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TextSmem& s = *reinterpret_cast<TextSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const TextDesc d = a.desc[blockIdx.x];
    const FusedTable* __restrict__ tb = a.tables + d.table;
    const bool has_prefix = d.flags & 1u, ends_row = (d.flags >> 1) & 1u;
    const uint64_t pb = a.nv.pre_off[d.row];
    const uint32_t plen = has_prefix ? (uint32_t)(a.nv.pre_off[d.row + 1] - pb) : 0u;
    const uint32_t n = plen + d.nbytes;
    const uint8_t cls = a.nv.cls[d.row];
    const uint32_t N = a.sv.n;

    for (uint32_t i = tid; i < 260; i += nthr) s.len_tok[i] = tb->len_tok[i];
    for (uint32_t i = tid; i < 256; i += nthr) s.lit_all[i] = tb->pre_lit[i];
    if (tid == 0) { s.crc_acc = 0; s.overflow = 0; }

    // ---- this thread's window of the row body
    const uint32_t w0 = 256u * tid;                          // segment-relative
    int nbytes = 0;
    if (w0 < d.nbytes) nbytes = (int)min(256u, d.nbytes - w0);
    const uint32_t span_start = d.byte0 + w0, span_end = span_start + (uint32_t)nbytes;  // body-relative
    uint32_t* win = s.text + kTextStride * tid;
    uint8_t* wb = reinterpret_cast<uint8_t*>(win);
    if (nbytes > 0) {
        uint32_t i = cls == kAuto ? span_start >> 2 : (cls == kX ? a.xspan[span_start >> 8] : span_start >> 1);
        const uint32_t first = i;
        // forced-minor samples inside this window (at most 128 samples)
        uint32_t ovm[4] = {0, 0, 0, 0};
        for (uint32_t o = 0; o < d.ovr_count; ++o) {
            const uint32_t smp = a.osamp[d.ovr_first + o];
            if (smp >= first && smp < first + 128u) {
                const uint32_t j = smp - first;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if ((j >> 5) == (uint32_t)w) ovm[w] |= 1u << (j & 31u);
            }
        }
        const uint4 t4 = reinterpret_cast<const uint4*>(a.nv.thr)[d.row];
        const uint32_t thr[4] = {t4.x, t4.y, t4.z, t4.w};
        const int K = a.nv.k[d.row];
        const uint64_t prow = a.row_base + d.row;
        uint32_t gcur = 0xFFFFFFFFu, p0 = 0, p1 = 0;
        for (; i < N; ++i) {
            const bool male = a.sv.sex[i] == 1;
            const uint32_t o = cls == kAuto ? 4u * i : (cls == kX ? a.sv.xoff[i] : 2u * i);
            if (o >= span_end) break;
            const uint32_t g = (2u * i) >> 5;
            if (g != gcur) {
                const uint32_t slots = 2u * N - 32u * g;
                const uint32_t valid = slots >= 32u ? 0xFFFFFFFFu : ((1u << slots) - 1u);
                draw_group_k(K, g, prow, a.k0, a.k1, thr, valid, p0, p1);
                gcur = g;
            }
            const uint32_t sh = (2u * i) & 31u;
            uint32_t al0 = ((p0 >> sh) & 1u) | (((p1 >> sh) & 1u) << 1);
            uint32_t al1 = ((p0 >> (sh + 1)) & 1u) | (((p1 >> (sh + 1)) & 1u) << 1);
            const uint32_t j = i - first;
            if ((pick4(ovm, (int)(j >> 5)) >> (j & 31u)) & 1u) { al0 = 1; al1 = 1; }
            const bool wide = cls == kAuto || (cls == kX && !male);
            const uint8_t term = (i + 1 == N) ? '\n' : '\t';
            uint8_t cell[4];
            cell[0] = (cls == kY && !male) ? '.' : (uint8_t)('0' + al0);
            cell[1] = wide ? '/' : term;
            cell[2] = (uint8_t)('0' + al1);
            cell[3] = term;
            const uint32_t width = wide ? 4u : 2u;
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                const uint32_t p = o + q;
                if (q < width && p >= span_start && p < span_end) wb[p - span_start] = cell[q];
            }
        }
    }
    __syncthreads();
    const bool has_prev = tid > 0;
    const uint32_t prev = has_prev ? s.text[kTextStride * (tid - 1) + 63] : 0u;
    const bool last_span = nbytes > 0 && w0 + (uint32_t)nbytes == d.nbytes;
    const uint32_t eob = tb->eob;
    auto wat = [&](int k) { return win[k]; };

    // ---- pass 1: tokens of this window, staged privately
    TextStage st{s.len_tok, s.lit_all, s.stage + tid, nthr, 0, 0, 0, 0};
    uint32_t crc = 0;
    if (nbytes > 0) {
        tokenize_words(wat, nbytes, has_prev, prev, st);
        if (last_span) st.put(eob & 0xFFFFFFu, eob >> 24);
        st.finish();
        if (st.bits > 32u * kTextStageWords) s.overflow = 1;
        // CRC of the window (slicing-by-4), shifted to the block end
        uint32_t c = 0;
        const int full = nbytes >> 2;
        for (int k = 0; k < full; ++k) {
            const uint32_t v = c ^ win[k];
            c = __ldg(&a.crc4[768 + (v & 0xFFu)]) ^ __ldg(&a.crc4[512 + ((v >> 8) & 0xFFu)]) ^
                __ldg(&a.crc4[256 + ((v >> 16) & 0xFFu)]) ^ __ldg(&a.crc4[v >> 24]);
        }
        for (int q = 4 * full; q < nbytes; ++q) c = __ldg(&a.crc4[(c ^ wb[q]) & 0xFFu]) ^ (c >> 8);
        if (c) crc = gf2_mulmod(a.xpow8[n - (plen + w0 + (uint32_t)nbytes)], c);
    }
    uint32_t pre_tok = 0;
    if (tid < plen) {
        const uint8_t c = a.nv.prefix[pb + tid];
        pre_tok = tb->pre_lit[c];
        crc ^= gf2_mulmod(a.xpow8[n - 1u - tid], __ldg(&a.crc4[c]));
    }
    const uint32_t pre_bits = pre_tok >> 24;
    uint32_t pre_off, span_off, total_pre, total_span;
    {
        uint32_t v0 = pre_bits, v1 = st.bits;
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
        span_off = b1 + v1 - st.bits;
        total_pre = t0;
        total_span = t1;
    }
    if (tid == 0) crc ^= gf2_mulmod(a.xpow8[n], 0xFFFFFFFFu);
    crc = warp_xor(crc);
    if ((tid & 31u) == 0 && crc) atomicXor(&s.crc_acc, crc);

    const uint32_t hdr_bits = tb->hdr_bits;
    const uint32_t data_bits = hdr_bits + total_pre + total_span;
    const uint32_t payload = (data_bits + 7u) / 8u;
    const uint32_t out_words = (data_bits + 31u) / 32u;
    const bool stored = payload > n + 5u;
    uint8_t* blk = a.slots + (uint64_t)d.slot * a.slot_stride + kSlotLead;
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
        const uint32_t dst = hdr_bits + total_pre + span_off;
        if (!overflow) {
            const uint32_t nb = st.bits;
            if (nb) {
                const uint32_t sh = dst & 31u;
                const uint32_t nsrc = (nb + 31u) / 32u;
                const uint32_t ndst = (sh + nb + 31u) / 32u;
                uint32_t* o = words + (dst >> 5);
                uint32_t pv = 0;
                for (uint32_t k = 0; k < ndst; ++k) {
                    const uint32_t cur = k < nsrc ? s.stage[k * nthr + tid] : 0u;
                    const uint32_t v = __funnelshift_l(pv, cur, sh);
                    if (k == 0 || k == ndst - 1) atomicOr(&o[k], v);
                    else o[k] = v;
                    pv = cur;
                }
            }
        } else if (nbytes > 0) {
            TextEmit em{s.len_tok, s.lit_all, words, dst};
            tokenize_words(wat, nbytes, has_prev, prev, em);
            if (last_span) em.put(eob & 0xFFFFFFu, eob >> 24);
        }
        out_payload = payload;
    } else {
        if (tid == 0) {
            blk[18] = 1;
            blk[19] = (uint8_t)n; blk[20] = (uint8_t)(n >> 8);
            blk[21] = (uint8_t)~n; blk[22] = (uint8_t)((~n) >> 8);
        }
        if (tid < plen) blk[23 + tid] = a.nv.prefix[pb + tid];
        for (int q = 0; q < nbytes; ++q) blk[23 + plen + w0 + q] = wb[q];
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
