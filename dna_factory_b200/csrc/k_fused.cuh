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
// Sink interface: gap_lit(gap, odd, bit) = [gap predicted bytes][literal '0'+bit] with gap in {0,1,3,5,..}
// (gap 1 is the separator before the allele: '/' when the allele slot is odd, '\t' when even),
// lit(id) and match(len) for the rare rest.  One loop over all 128 mismatch bits keeps a warp's trip count
// at the maximum of the per-thread totals instead of the sum of per-word maxima.
struct CellMasks {
    uint32_t x[4];   // mismatch bits (allele differs from the same slot of the previous cell), validity-masked
    int first_lits;  // 1 when the first cell has no predecessor and goes out as literals
};

__host__ __device__ __forceinline__ CellMasks cell_mismatches(const uint32_t m[4], uint32_t carry, bool has_prev, int ncells) {
    CellMasks c;
    c.x[0] = m[0] ^ ((m[0] << 2) | (carry & 3u));
    c.x[1] = m[1] ^ ((m[1] << 2) | (m[0] >> 30));
    c.x[2] = m[2] ^ ((m[2] << 2) | (m[1] >> 30));
    c.x[3] = m[3] ^ ((m[3] << 2) | (m[2] >> 30));
    const int na = 2 * ncells;  // only the first 2*ncells allele slots exist
    if (na < 128) c.x[3] = na > 96 ? (c.x[3] & (0xFFFFFFFFu >> (128 - na))) : 0u;
    if (na < 96) c.x[2] = na > 64 ? (c.x[2] & (0xFFFFFFFFu >> (96 - na))) : 0u;
    if (na < 64) c.x[1] = na > 32 ? (c.x[1] & (0xFFFFFFFFu >> (64 - na))) : 0u;
    if (na < 32) c.x[0] &= 0xFFFFFFFFu >> (32 - na);
    c.first_lits = has_prev ? 0 : 1;
    if (!has_prev) c.x[0] &= ~3u;
    return c;
}

template <class Sink>
__host__ __device__ __forceinline__ void tokenize_cells(const uint32_t m[4], uint32_t carry, bool has_prev, int ncells,
                                                        bool ends_row, Sink& sink) {
    const int span_len = 4 * ncells;
    const int end = ends_row ? span_len - 1 : span_len;  // the final '\n' is always a literal
    int prev_end = 0;
    const CellMasks c = cell_mismatches(m, carry, has_prev, ncells);
    if (c.first_lits) {  // nothing 4 bytes back that is a cell: the first cell goes out as literals
        sink.lit((int)(m[0] & 1u));
        sink.lit(kLitSlash);
        sink.lit((int)((m[0] >> 1) & 1u));
        if (ncells == 1 && ends_row) {
            sink.lit(kLitNl);
            return;
        }
        sink.lit(kLitTab);
        prev_end = 4;
    }
    auto lit_at = [&](int q) {  // literal for text byte q of the span
        if (q & 1) sink.lit((q & 3) == 1 ? kLitSlash : kLitTab);
        else sink.lit((int)((pick4(m, q >> 6) >> ((q >> 1) & 31)) & 1u));
    };
    // walk the set bits of the 128-bit mismatch mask; moving on to the next non-empty word is a few predicated
    // selects at the end of an iteration, not an iteration of its own, so sparse spans do not pay for it
    auto next_word = [&](int from) {  // first non-empty word with index >= from, 4 when none
        return (from <= 0 && c.x[0]) ? 0 : ((from <= 1 && c.x[1]) ? 1 : ((from <= 2 && c.x[2]) ? 2 : ((from <= 3 && c.x[3]) ? 3 : 4)));
    };
    int cw = next_word(0);
    uint32_t cx = pick4(c.x, cw), cm = pick4(m, cw);
    while (cw < 4) {
#ifdef __CUDA_ARCH__
        const int b = __ffs((int)cx) - 1;
#else
        const int b = __builtin_ctz(cx);
#endif
        cx &= cx - 1;
        const int p = 64 * cw + 2 * b;
        int gap = p - prev_end;
        if (gap == 2) {  // only right after a span start: predicted allele + separator, too short for a match
            lit_at(prev_end);
            lit_at(prev_end + 1);
            gap = 0;
        }
        sink.gap_lit(gap, b & 1, (int)((cm >> b) & 1u));
        prev_end = p + 1;
        if (!cx) {
            cw = next_word(cw + 1);
            cx = pick4(c.x, cw);
            cm = pick4(m, cw);
        }
    }
    {   // tail: bytes [prev_end, end) are predicted
        const int gap = end - prev_end;
        if (gap >= 3) sink.match(gap);
        else
            for (int q = prev_end; q < end; ++q) lit_at(q);
    }
    if (ends_row) sink.lit(kLitNl);
}

constexpr int kStageWords = 16;  // per-thread staging capacity (512 bits); beyond it the block re-emits

// Pass-1 sink: appends this thread's bits to its private staging words stage[wi * stride] (word-interleaved
// across threads, so a warp writing word wi hits 32 different banks).  Branch-free: the current word is
// stored after every token, and the word index advances by (bit count >> 5).
struct FusedStage {
    const uint32_t* len_tok;
    uint32_t* stage;   // already offset by the thread id
    uint32_t stride;
    uint32_t lit0, lit1, lit_slash, lit_tab, lit_nl;
    uint32_t wi, nacc;
    uint64_t acc;
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // n <= 32
        acc |= (uint64_t)v << nacc;
        nacc += n;
        if (wi < (uint32_t)kStageWords) stage[wi * stride] = (uint32_t)acc;
        const uint32_t adv = nacc >> 5;
        wi += adv;
        acc >>= (adv << 5);
        nacc &= 31u;
    }
    __device__ __forceinline__ void finish() {  // the word begun by the last advance has not been stored yet
        if (nacc && wi < (uint32_t)kStageWords) stage[wi * stride] = (uint32_t)acc;
    }
    __device__ __forceinline__ uint32_t bits() const { return 32u * wi + nacc; }
    __device__ __forceinline__ uint32_t lit_tok(int id) const {
        return id == kLit0 ? lit0 : (id == kLit1 ? lit1 : (id == kLitSlash ? lit_slash : (id == kLitTab ? lit_tab : lit_nl)));
    }
    __device__ __forceinline__ void lit(int id) {
        const uint32_t t = lit_tok(id);
        put(t & 0xFFFFFFu, t >> 24);
    }
    __device__ __forceinline__ void match(int len) {
        const uint32_t t = len_tok[len];
        put(t & 0xFFFFFFu, t >> 24);
    }
    __device__ __forceinline__ void gap_lit(int gap, int odd, int bit) {
        // len_tok[0] is the empty token, len_tok[1] / len_tok[2] hold the '\t' / '/' literal (gap 2 never gets here)
        const uint32_t t1 = len_tok[gap + (gap == 1 ? odd : 0)];
        const uint32_t t2 = bit ? lit1 : lit0;
        const uint32_t n1 = t1 >> 24;
        put((t1 & 0xFFFFFFu) | ((t2 & 0xFFFFFFu) << n1), n1 + (t2 >> 24));   // tables keep n1 + n2 <= 32
    }
    // [gap predicted bytes as one match, gap == 0 or >= 3][literal id]  (k_fused_x.cuh)
    __device__ __forceinline__ void gap_tok(int gap, int id) {
        const uint32_t t1 = gap ? len_tok[gap] : 0u;
        const uint32_t t2 = lit_tok(id);
        const uint32_t n1 = t1 >> 24;
        put((t1 & 0xFFFFFFu) | ((t2 & 0xFFFFFFu) << n1), n1 + (t2 >> 24));
    }
    // [match of gap >= 3 bytes | literal id_a if gap >= 1, literal id_b if gap == 2 | nothing][literal id] as ONE
    // token, branch-free (tables keep cell literals <= 10 bits and matches <= 21, so the sum fits 32 bits)
    __device__ __forceinline__ void fused3(int gap, int id_a, int id_b, int id) {
        const uint32_t t1 = gap >= 3 ? len_tok[gap] : (gap >= 1 ? lit_tok(id_a) : 0u);
        const uint32_t t2 = gap == 2 ? lit_tok(id_b) : 0u;
        const uint32_t t3 = lit_tok(id);
        const uint32_t n1 = t1 >> 24, n2 = n1 + (t2 >> 24);
        put((t1 & 0xFFFFFFu) | ((t2 & 0xFFFFFFu) << n1) | ((t3 & 0xFFFFFFu) << n2), n2 + (t3 >> 24));
    }
};

// Slow-path sink (a thread overflowed its staging): ORs every token straight into the zeroed output words.
struct FusedEmit {
    const uint32_t* len_tok;
    uint32_t lit0, lit1, lit_slash, lit_tab, lit_nl;
    uint32_t* words;
    uint32_t pos;
    __device__ void put(uint32_t v, uint32_t n) {
        if (!n) return;
        const uint32_t wi = pos >> 5, sh = pos & 31u;
        atomicOr(&words[wi], v << sh);
        if (sh + n > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
        pos += n;
    }
    __device__ uint32_t lit_tok(int id) const {
        return id == kLit0 ? lit0 : (id == kLit1 ? lit1 : (id == kLitSlash ? lit_slash : (id == kLitTab ? lit_tab : lit_nl)));
    }
    __device__ void lit(int id) { const uint32_t t = lit_tok(id); put(t & 0xFFFFFFu, t >> 24); }
    __device__ void match(int len) { const uint32_t t = len_tok[len]; put(t & 0xFFFFFFu, t >> 24); }
    __device__ void gap_lit(int gap, int odd, int bit) {
        if (gap == 1) lit(odd ? kLitSlash : kLitTab);
        else if (gap) match(gap);
        lit(bit);
    }
    __device__ void gap_tok(int gap, int id) {
        if (gap) match(gap);
        lit(id);
    }
    __device__ void fused3(int gap, int id_a, int id_b, int id) {
        if (gap >= 3) match(gap);
        else {
            if (gap >= 1) lit(id_a);
            if (gap == 2) lit(id_b);
        }
        lit(id);
    }
};

constexpr int kFusedMaxThreads = 256;

struct FusedSmem {
    uint32_t stage[kStageWords * kFusedMaxThreads];
    uint32_t len_tok[260];
    uint32_t lit_tok[8];
    uint32_t sm[4][kFusedMaxThreads];      // masks in processing (sorted) order
    uint32_t smeta[kFusedMaxThreads];      // span index | carry << 16
    uint32_t span_bits[kFusedMaxThreads];  // by span: bits, then destination bit offset
    uint32_t last_bits[kFusedMaxThreads];
    uint32_t cnt[132];                     // counting sort by mismatch count
    uint32_t warp_pre[8], warp_span[8];
    uint32_t crc_acc;
    uint32_t overflow;
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


// One CTA per BGZF block.  Thread t draws span t (64 samples) and owns prefix byte t; the tokenisation of
// the spans is then dealt out by mismatch count (a counting sort inside the CTA), so that the 32 lanes of
// a warp run token loops of nearly equal length.
__global__ void __launch_bounds__(kFusedMaxThreads, 4) k_fused_auto(const FusedArgs a) {
    __shared__ FusedSmem s;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const FusedDesc d = a.desc[blockIdx.x];
    const FusedTable* __restrict__ tb = a.tables + d.table;
    const bool has_prefix = d.flags & 1u, ends_row = (d.flags >> 1) & 1u;
    const uint64_t pb = a.nv.pre_off[d.row];
    const uint32_t plen = has_prefix ? (uint32_t)(a.nv.pre_off[d.row + 1] - pb) : 0u;
    const uint32_t n = plen + 4u * d.ncells;  // text bytes of this block
    const uint32_t nspans = (d.ncells + 63u) / 64u;

    for (uint32_t i = tid; i < 260; i += nthr) s.len_tok[i] = tb->len_tok[i];
    if (tid < 8) s.lit_tok[tid] = tb->lit[tid];
    for (uint32_t i = tid; i < 132; i += nthr) s.cnt[i] = 0;
    if (tid == 0) { s.crc_acc = 0; s.overflow = 0; }

    // ---- draw this span's 128 allele bits
    const uint32_t cs = d.cell0 + 64u * tid;
    int nc = 0;
    if (tid < nspans) nc = (int)min(64u, d.ncells - 64u * tid);
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

    // ---- counting sort of the spans by mismatch count (descending): who tokenises what
    uint32_t key = 0, rank_in_bin = 0;
    const uint32_t carry = tid ? s.last_bits[tid - 1] : 0u;
    if (nc > 0) {
        const CellMasks cmk = cell_mismatches(m, carry, tid > 0, nc);
        key = __popc(cmk.x[0]) + __popc(cmk.x[1]) + __popc(cmk.x[2]) + __popc(cmk.x[3]);
        rank_in_bin = atomicAdd(&s.cnt[key], 1u);
    }
    // ---- CRC32 share of this span: template ^ delta (affine), shifted to the block end
    uint32_t crc = 0;
    if (nc > 0 && (m[0] | m[1] | m[2] | m[3])) {
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
        const uint32_t span_end = plen + 4u * (64u * tid + (uint32_t)nc);
        crc = gf2_mulmod(a.xpow8[n - span_end], sp);
    }
    // prefix byte of this thread: literal code, and its share of the CRC
    uint32_t pre_tok = 0;
    if (tid < plen) {
        const uint8_t c = a.nv.prefix[pb + tid];
        pre_tok = tb->pre_lit[c];
        crc ^= gf2_mulmod(a.xpow8[n - 1u - tid], __ldg(&a.crctab[c]));
    }
    if (tid == 0) crc ^= d.body_crc ^ gf2_mulmod(a.xpow8[n], 0xFFFFFFFFu);
    crc = warp_xor(crc);
    if ((tid & 31u) == 0 && crc) atomicXor(&s.crc_acc, crc);
    __syncthreads();
    if (tid < 32) {  // bin starts, heaviest spans first (keys 0..128)
        uint32_t c4[5], tot = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t k = 128u - (5u * tid + i);            // lane 0 holds keys 128..124, ...
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
        s.sm[0][pos] = m[0];
        s.sm[1][pos] = m[1];
        s.sm[2][pos] = m[2];
        s.sm[3][pos] = m[3];
        s.smeta[pos] = tid | (carry << 16);
    }
    s.span_bits[tid] = 0;
    __syncthreads();

    // ---- pass 1: tokens of the span dealt to this thread, staged privately
    const uint32_t eob = tb->eob;
    FusedStage st{s.len_tok, s.stage + tid, nthr, s.lit_tok[kLit0], s.lit_tok[kLit1], s.lit_tok[kLitSlash],
                  s.lit_tok[kLitTab], s.lit_tok[kLitNl], 0, 0, 0};
    uint32_t pm[4] = {0, 0, 0, 0};
    uint32_t psp = 0, pcarry = 0;
    int pnc = 0;
    bool p_end = false, p_last = false;
    const bool worker = tid < nspans;
    if (worker) {
        pm[0] = s.sm[0][tid]; pm[1] = s.sm[1][tid]; pm[2] = s.sm[2][tid]; pm[3] = s.sm[3][tid];
        const uint32_t meta = s.smeta[tid];
        psp = meta & 0xFFFFu;
        pcarry = meta >> 16;
        pnc = (int)min(64u, d.ncells - 64u * psp);
        p_last = 64u * psp + (uint32_t)pnc == d.ncells;
        p_end = ends_row && p_last;
        tokenize_cells(pm, pcarry, psp > 0, pnc, p_end, st);
        if (p_last) st.put(eob & 0xFFFFFFu, eob >> 24);
        st.finish();
        if (st.bits() > 32u * kStageWords) s.overflow = 1;
        s.span_bits[psp] = st.bits();
    }
    __syncthreads();
    // ---- exclusive scans over the CTA, in span order: prefix literal bits and span bits
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
        s.span_bits[tid] = b1 + v1 - my_bits;  // now: offset of span `tid` among the span bits
    }
    const uint32_t hdr_bits = tb->hdr_bits;
    const uint32_t data_bits = hdr_bits + total_pre + total_span;
    const uint32_t payload = (data_bits + 7u) / 8u;
    const uint32_t out_words = (data_bits + 31u) / 32u;
    const bool stored = payload > n + 5u;  // cannot happen with sane tables; keeps BSIZE <= 64 KiB regardless
    uint8_t* blk = a.slots + (uint64_t)d.slot * kSlot + kSlotLead;
    uint32_t* words = reinterpret_cast<uint32_t*>(blk + 18);  // 4-byte aligned

    uint32_t out_payload;
    if (!stored) {
        // zero the words that will be OR-ed into (header words are written, not OR-ed)
        const uint32_t hdr_words = (hdr_bits + 31u) / 32u;
        for (uint32_t i = tid; i < out_words + 1u; i += nthr) words[i] = i < hdr_words ? tb->hdr[i] : 0u;
        __syncthreads();
        const bool overflow = s.overflow != 0;
        if (pre_bits) {  // prefix literal: at most 15 bits
            const uint32_t pos = hdr_bits + pre_off, wi = pos >> 5, sh = pos & 31u, v = pre_tok & 0xFFFFFFu;
            atomicOr(&words[wi], v << sh);
            if (sh + pre_bits > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
        }
        if (worker) {
            const uint32_t dst = hdr_bits + total_pre + s.span_bits[psp];  // first bit of that span in the block
            if (!overflow) {
                // ---- pass 2 (fast): move the staged bits to their final position
                const uint32_t nb = st.bits();
                const uint32_t sh = dst & 31u;
                const uint32_t nsrc = (nb + 31u) / 32u;
                const uint32_t ndst = (sh + nb + 31u) / 32u;
                uint32_t* o = words + (dst >> 5);
                uint32_t prev = 0;
                for (uint32_t k = 0; k < ndst; ++k) {
                    const uint32_t cur = k < nsrc ? s.stage[k * nthr + tid] : 0u;
                    const uint32_t v = __funnelshift_l(prev, cur, sh);  // (cur:prev) << sh, upper word
                    if (k == 0 || k == ndst - 1) atomicOr(&o[k], v);
                    else o[k] = v;
                    prev = cur;
                }
            } else {
                // ---- pass 2 (slow): re-tokenise straight into the output words
                FusedEmit em{s.len_tok, st.lit0, st.lit1, st.lit_slash, st.lit_tab, st.lit_nl, words, dst};
                tokenize_cells(pm, pcarry, psp > 0, pnc, p_end, em);
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
            uint8_t* p = blk + 23 + plen + 256u * tid;
            const bool my_end = ends_row && 64u * tid + (uint32_t)nc == d.ncells;
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
    if (tid < 26) {  // 18-byte BGZF header, CRC32, ISIZE
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
