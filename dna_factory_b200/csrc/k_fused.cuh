// The fused hot-path kernel: allele draws -> (virtual) VCF text -> deflate tokens -> BGZF block, one CTA
// per BGZF block, for rows whose cells are all 4 bytes wide (autosomes, K <= 2 alleles).  The VCF text is
// never materialised: every thread owns one 256-byte span of the row (64 genotype cells = 128 allele
// bits = four 32-bit masks straight out of the bit-sliced Philox draw) and derives from those masks
//   * the deflate tokens of the P4 parse (see k_deflate.cuh): a byte is predicted by the byte 4 back, i.e.
//     by the same allele slot of the previous sample, so the mismatching bytes are exactly the set bits
//     of  m ^ (m << 2)  -- literals there, one distance-4 match per gap in between
//   * the CRC32 of its span: CRC is affine, so CRC(text) = CRC(all-reference template) ^ L(delta) where
//     delta has 0x01 at every minor-allele byte; L(delta) is 16 table lookups on the mask bytes, and the
//     template's CRC depends only on (row prefix, segment) and is precomputed
// Huffman codes are static per block: the host precomputes one code (and its serialized dynamic-block
// header) per minor-allele-frequency bucket from the token statistics the Bernoulli(maf) model implies,
// so no histogram pass, no on-device tree construction.
//
// Reference behaviour restated: pop_factory.py:471-508 (row loop + row text) and the BgzfWriter framing
// (call site pop_factory.py:449).  HBM traffic per block: the compressed bytes plus ~100 bytes of metadata.
#pragma once
#include "dnaf_device.cuh"
#include "k_sample_format.cuh"

namespace dnaf {

constexpr int kFusedThreads = 256;
constexpr uint32_t kFusedOutWords = 5120;  // 20 KiB of payload staged in shared memory; larger -> global path

// literal symbol ids of the five bytes a genotype cell can hold
enum : int { kLit0 = 0, kLit1 = 1, kLitSlash = 2, kLitTab = 3, kLitNl = 4 };

struct FusedDesc {
    uint64_t row;        // global row
    uint32_t cell0;      // first sample of the segment (multiple of 64)
    uint32_t ncells;     // samples in the segment (<= 255 * 64)
    uint32_t slot;       // output slot / block index inside the pass
    uint32_t flags;      // bit0: segment starts the row (has the prefix), bit1: segment ends the row
    uint32_t ovr_first;  // overrides of this row: [ovr_first, ovr_first + ovr_count)
    uint32_t ovr_count;
    uint32_t table;      // index into the FusedTable array
    uint32_t body_crc;   // L(template body of this segment)
};

struct FusedTable {
    uint32_t len_tok[260];   // match length -> (code | extra | distance bit) | total bits << 24
    uint32_t lit[8];         // cell literals by id: code | bits << 24
    uint32_t eob;            // end-of-block code | bits << 24
    uint32_t hdr_bits;       // serialized dynamic-block header
    uint32_t hdr[62];
    uint32_t pre_lit[256];   // literal codes for prefix bytes (segments that start a row)
};

__host__ __device__ __forceinline__ uint32_t pick4(const uint32_t m[4], int w) {
    return w == 0 ? m[0] : (w == 1 ? m[1] : (w == 2 ? m[2] : m[3]));
}

// ---- mask-domain P4 tokeniser, shared by the kernel and by the host's table builder ----
// m[0..3]: minor-allele bits of the span's 128 allele slots; carry: the two bits of the previous cell.
template <class Sink>
__host__ __device__ __forceinline__ void tokenize_cells(const uint32_t m[4], uint32_t carry, bool has_prev, int ncells,
                                                        bool ends_row, Sink& sink) {
    const int span_len = 4 * ncells;
    const int end = ends_row ? span_len - 1 : span_len;  // the final '\n' is always a literal
    int prev_end = 0;
    uint32_t x[4];
    x[0] = m[0] ^ ((m[0] << 2) | (carry & 3u));
    x[1] = m[1] ^ ((m[1] << 2) | (m[0] >> 30));
    x[2] = m[2] ^ ((m[2] << 2) | (m[1] >> 30));
    x[3] = m[3] ^ ((m[3] << 2) | (m[2] >> 30));
    if (!has_prev) {  // nothing 4 bytes back that is a cell: the first cell goes out as literals
        sink.lit((int)(m[0] & 1u));
        sink.lit(kLitSlash);
        sink.lit((int)((m[0] >> 1) & 1u));
        if (ncells == 1 && ends_row) {
            sink.lit(kLitNl);
            return;
        }
        sink.lit(kLitTab);
        prev_end = 4;
        x[0] &= ~3u;
    }
    auto gap_to = [&](int p) {  // bytes [prev_end, p) are predicted
        const int gap = p - prev_end;
        if (gap >= 3) {
            sink.match(gap);
        } else {
            for (int q = prev_end; q < p; ++q) {
                if (q & 1) sink.lit((q & 3) == 1 ? kLitSlash : kLitTab);
                else sink.lit((int)((pick4(m, q >> 6) >> ((q >> 1) & 31)) & 1u));
            }
        }
    };
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t xw = x[w];
        if (32 * w + 32 > 2 * ncells) xw &= (2 * ncells > 32 * w) ? (0xFFFFFFFFu >> (32 * w + 32 - 2 * ncells)) : 0u;
        while (xw) {
#ifdef __CUDA_ARCH__
            const int b = __ffs((int)xw) - 1;
#else
            const int b = __builtin_ctz(xw);
#endif
            xw &= xw - 1;
            const int p = 2 * (32 * w + b);
            gap_to(p);
            sink.lit((int)((m[w] >> b) & 1u));
            prev_end = p + 1;
        }
    }
    gap_to(end);
    if (ends_row) sink.lit(kLitNl);
}

struct FusedCount {
    const uint32_t* len_tok;
    const uint32_t* lit_tok;
    uint32_t bits;
    __host__ __device__ void lit(int id) { bits += lit_tok[id] >> 24; }
    __host__ __device__ void match(int len) { bits += len_tok[len] >> 24; }
};

// Appends bits at an arbitrary bit offset; the first and the last word it touches are shared with the
// neighbouring threads (atomic OR into zeroed memory), the words in between are owned.
template <bool kGlobal>
struct FusedEmit {
    const uint32_t* len_tok;
    const uint32_t* lit_tok;
    uint32_t* words;
    uint32_t wi;
    uint32_t nacc;
    uint64_t acc;
    bool first;
    __device__ void init(uint32_t* w, uint32_t bitpos) {
        words = w;
        wi = bitpos >> 5;
        nacc = bitpos & 31u;
        acc = 0;
        first = true;
    }
    __device__ void put(uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            if (first || kGlobal) atomicOr(&words[wi], (uint32_t)acc);
            else words[wi] = (uint32_t)acc;
            first = false;
            ++wi;
            acc >>= 32;
            nacc -= 32;
        }
    }
    __device__ void finish() {
        if (nacc) atomicOr(&words[wi], (uint32_t)acc);
    }
    __device__ void lit(int id) { put(lit_tok[id] & 0xFFFFFFu, lit_tok[id] >> 24); }
    __device__ void match(int len) { put(len_tok[len] & 0xFFFFFFu, len_tok[len] >> 24); }
};

struct FusedSmem {
    uint32_t out[kFusedOutWords];
    uint32_t len_tok[260];
    uint32_t lit_tok[8];
    uint32_t last_bits[kFusedThreads];
    uint32_t warp_tmp[8];
    uint32_t crc_acc;
    uint32_t total_bits;
};

struct FusedArgs {
    SampleView sv;
    SnpView nv;
    const FusedDesc* desc;
    const FusedTable* tables;
    const uint32_t* etab;     // [16][256] span-local CRC contributions of mask bytes
    const uint32_t* crctab;   // [256]
    const uint32_t* xpow8;    // [kBlk+1]
    const uint64_t* orow;
    const uint32_t* osamp;
    uint64_t row_base;
    uint32_t k0, k1;
    uint8_t* slots;
    uint32_t* sizes;
    uint32_t* crcs;
};

__global__ void __launch_bounds__(kFusedThreads, 4) k_fused_auto(const FusedArgs a) {
    __shared__ FusedSmem s;
    const uint32_t tid = threadIdx.x;
    const FusedDesc d = a.desc[blockIdx.x];
    const FusedTable* __restrict__ tb = a.tables + d.table;
    const bool has_prefix = d.flags & 1u, ends_row = (d.flags >> 1) & 1u;
    const uint64_t pb = a.nv.pre_off[d.row];
    const uint32_t plen = has_prefix ? (uint32_t)(a.nv.pre_off[d.row + 1] - pb) : 0u;
    const uint32_t n = plen + 4u * d.ncells;  // text bytes of this block

    for (uint32_t i = tid; i < 260; i += kFusedThreads) s.len_tok[i] = tb->len_tok[i];
    if (tid < 8) s.lit_tok[tid] = tb->lit[tid];
    if (tid == 0) s.crc_acc = 0;

    // ---- draw this span's 128 allele bits
    const int span = (int)tid - 1;  // thread 0 owns the prefix
    const uint32_t cs = d.cell0 + 64u * (uint32_t)(span < 0 ? 0 : span);
    int nc = 0;
    if (span >= 0 && 64u * (uint32_t)span < d.ncells) nc = (int)min(64u, d.ncells - 64u * (uint32_t)span);
    uint32_t m[4] = {0, 0, 0, 0};
    if (nc > 0 && a.nv.k[d.row] == 2) {
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
    // forced-minor cells (pop_factory.py:495-499)
    if (nc > 0) {
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
    }
    s.last_bits[tid] = m[3] >> 30;
    __syncthreads();
    const uint32_t carry = tid ? s.last_bits[tid - 1] : 0u;
    const bool has_prev = span > 0;
    const bool my_end = ends_row && nc > 0 && 64u * (uint32_t)span + (uint32_t)nc == d.ncells;

    // ---- pass 1: bits this thread will emit
    FusedCount cnt{s.len_tok, s.lit_tok, 0};
    uint32_t pre_crc = 0;
    if (tid == 0) {
        for (uint32_t i = 0; i < plen; ++i) {
            const uint8_t c = a.nv.prefix[pb + i];
            cnt.bits += tb->pre_lit[c] >> 24;
            pre_crc = a.crctab[(pre_crc ^ c) & 0xFFu] ^ (pre_crc >> 8);
        }
    } else if (nc > 0) {
        tokenize_cells(m, carry, has_prev, nc, my_end, cnt);
    }
    {   // exclusive scan over the CTA
        uint32_t v = cnt.bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if ((tid & 31u) >= (uint32_t)o) v += u;
        }
        if ((tid & 31u) == 31u) s.warp_tmp[tid >> 5] = v;
        __syncthreads();
        uint32_t base = tb->hdr_bits;
        for (uint32_t wv = 0; wv < (tid >> 5); ++wv) base += s.warp_tmp[wv];
        cnt.bits = base + v - cnt.bits;  // now: this thread's first bit
        if (tid == kFusedThreads - 1) s.total_bits = base + v;
    }
    __syncthreads();
    const uint32_t eob = tb->eob;
    const uint32_t data_bits = s.total_bits + (eob >> 24);
    const uint32_t payload = (data_bits + 7u) / 8u;
    const uint32_t out_words = (data_bits + 31u) / 32u;
    const bool stored = payload > n + 5u;  // cannot happen with sane tables; keeps BSIZE <= 64 KiB regardless
    const bool in_smem = out_words <= kFusedOutWords;
    uint8_t* slot = a.slots + (uint64_t)d.slot * kSlot;
    uint32_t* gwords = reinterpret_cast<uint32_t*>(slot + 20);  // 4-byte aligned staging inside the slot

    // ---- CRC32: template ^ delta (affine), every thread shifts its span's share to the block end
    uint32_t crc = 0;
    if (tid == 0) {
        crc = d.body_crc ^ gf2_mulmod(a.xpow8[n], 0xFFFFFFFFu);
        if (pre_crc) crc ^= gf2_mulmod(a.xpow8[n - plen], pre_crc);
    } else if (nc > 0 && (m[0] | m[1] | m[2] | m[3])) {
        uint32_t mm[4] = {m[0], m[1], m[2], m[3]};
        if (nc < 64) {  // partial last span: align its end with the table's span end (128-bit left shift)
            const uint32_t sh = 2u * (64u - (uint32_t)nc);
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
        const uint32_t span_end = plen + 4u * (64u * (uint32_t)span + (uint32_t)nc);
        crc = gf2_mulmod(a.xpow8[n - span_end], sp);
    }
    crc = warp_xor(crc);
    if ((tid & 31u) == 0 && crc) atomicXor(&s.crc_acc, crc);

    uint32_t out_payload;
    if (!stored) {
        uint32_t* words = in_smem ? s.out : gwords;
        for (uint32_t i = tid; i < out_words + 1; i += kFusedThreads) words[i] = i < (tb->hdr_bits + 31u) / 32u ? tb->hdr[i] : 0u;
        if (!in_smem) __threadfence();
        __syncthreads();
        // ---- pass 2: emit
        if (in_smem) {
            FusedEmit<false> em{s.len_tok, s.lit_tok};
            em.init(words, cnt.bits);
            if (tid == 0) {
                for (uint32_t i = 0; i < plen; ++i) {
                    const uint32_t c = tb->pre_lit[a.nv.prefix[pb + i]];
                    em.put(c & 0xFFFFFFu, c >> 24);
                }
            } else if (nc > 0) {
                tokenize_cells(m, carry, has_prev, nc, my_end, em);
            }
            if (tid == kFusedThreads - 1) em.put(eob & 0xFFFFFFu, eob >> 24);
            em.finish();
        } else {
            FusedEmit<true> em{s.len_tok, s.lit_tok};
            em.init(words, cnt.bits);
            if (tid == 0) {
                for (uint32_t i = 0; i < plen; ++i) {
                    const uint32_t c = tb->pre_lit[a.nv.prefix[pb + i]];
                    em.put(c & 0xFFFFFFu, c >> 24);
                }
            } else if (nc > 0) {
                tokenize_cells(m, carry, has_prev, nc, my_end, em);
            }
            if (tid == kFusedThreads - 1) em.put(eob & 0xFFFFFFu, eob >> 24);
            em.finish();
            __threadfence();
        }
        __syncthreads();
        // payload goes to slot + 18; staged words sit at slot + 20 (global path) or in shared memory
        if (in_smem) {
            const uint8_t* ob = reinterpret_cast<const uint8_t*>(s.out);
            // slot + 18 is 2-byte aligned: move 16-bit units
            const uint16_t* o16 = reinterpret_cast<const uint16_t*>(ob);
            uint16_t* d16 = reinterpret_cast<uint16_t*>(slot + 18);
            for (uint32_t i = tid; i < (payload + 1u) / 2u; i += kFusedThreads) d16[i] = o16[i];
        } else {
            // shift down by two bytes in place, front to back, one CTA-wide stripe at a time
            uint16_t* p16 = reinterpret_cast<uint16_t*>(slot + 18);
            for (uint32_t base = 0; base < (payload + 1u) / 2u; base += kFusedThreads) {
                const uint32_t i = base + tid;
                uint16_t v = 0;
                if (i < (payload + 1u) / 2u) v = p16[i + 1];
                __syncthreads();
                if (i < (payload + 1u) / 2u) p16[i] = v;
                __syncthreads();
            }
        }
        out_payload = payload;
    } else {
        // stored deflate block: format the text itself (rare safety net)
        if (tid == 0) {
            slot[18] = 1;
            slot[19] = (uint8_t)n; slot[20] = (uint8_t)(n >> 8);
            slot[21] = (uint8_t)~n; slot[22] = (uint8_t)((~n) >> 8);
            for (uint32_t i = 0; i < plen; ++i) slot[23 + i] = a.nv.prefix[pb + i];
        } else if (nc > 0) {
            uint8_t* p = slot + 23 + plen + 256u * (uint32_t)span;
            for (int c = 0; c < nc; ++c) {
                const uint32_t bits = (pick4(m, c >> 4) >> (2 * (c & 15))) & 3u;
                p[4 * c] = '0' + (bits & 1u);
                p[4 * c + 1] = '/';
                p[4 * c + 2] = '0' + (bits >> 1);
                p[4 * c + 3] = (my_end && c == nc - 1) ? '\n' : '\t';
            }
        }
        out_payload = n + 5u;
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t crc32 = ~s.crc_acc;
        const uint32_t bsize = out_payload + 25u;
        const uint8_t head[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00,
                                  (uint8_t)bsize, (uint8_t)(bsize >> 8)};
        for (int i = 0; i < 18; ++i) slot[i] = head[i];
        uint8_t* tail = slot + 18 + out_payload;
        for (int i = 0; i < 4; ++i) tail[i] = (uint8_t)(crc32 >> (8 * i));
        for (int i = 0; i < 4; ++i) tail[4 + i] = (uint8_t)(n >> (8 * i));
        a.sizes[d.slot] = out_payload + 26u;
        a.crcs[d.slot] = crc32;
    }
}

}  // namespace dnaf
