// Device-side building blocks shared by the kernels: Philox4x32-10, the lazily evaluated
// bit-sliced allele draw, CRC32 arithmetic in GF(2)[x]/P and deflate constant tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dnaf {

constexpr int kKmax = 4;
constexpr uint32_t kBlk = 65280;      // uncompressed bytes per BGZF block (htslib's choice; stored fallback fits)
constexpr uint32_t kSpan = 256;       // bytes of text one thread tokenises; a whole span is one match at most
constexpr uint32_t kSlot = 65536;     // a BGZF block never exceeds 64 KiB
constexpr uint32_t kCrcPoly = 0xEDB88320u;
// Slot layout: the BGZF block starts at slot + 14, so that its deflate payload (slot + 32) is 16-byte aligned.
constexpr uint32_t kSlotLead = 14;

enum : uint8_t { kAuto = 0, kX = 1, kY = 2, kMT = 3 };

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  One call = 4 output words.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// ------------------------------------------------------------------------------------------------
// Allele draw for one group of 32 consecutive allele slots j = 32g .. 32g+31 of one SNP row.
//
// Replay RNG spec: bit (31-b) of the 32-bit uniform U_j is bit (j&31) of
//   W(row,g,b) = philox4x32_10(ctr=(g, b>>2, row_lo, row_hi), key=seed)[b&3].
// The reference's pick (pop_factory.py:92-95) is "first k with cum_k >= u"  <=>  first k with
// U_j <= T_k, T_k = floor(cum_k*2^32).  Because lane j's uniform is spread over bit j of 32 words, the
// compare runs MSB-first on all 32 lanes at once and stops as soon as every lane has met a bit in which
// U and T differ (2 bits per lane on average): only the Philox words that are needed get computed.
// Returns the allele index as two bit planes (idx = p0 | p1<<1 per lane).
template <int K>
__device__ __forceinline__ void draw_group(uint32_t g, uint64_t row, uint32_t k0, uint32_t k1, const uint32_t* thr,
                                           uint32_t valid, uint32_t& p0, uint32_t& p1) {
    constexpr int C = K - 1;  // thresholds that need a compare; the last one is saturated (always true)
    uint32_t eq[C > 0 ? C : 1], lt[C > 0 ? C : 1];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        eq[k] = valid;
        lt[k] = 0;
    }
    if (C > 0) {
        for (uint32_t q = 0; q < 8; ++q) {
            uint32_t live = 0;
#pragma unroll
            for (int k = 0; k < C; ++k) live |= eq[k];
            if (!live) break;
            const uint4 w = philox4x32_10(g, q, (uint32_t)row, (uint32_t)(row >> 32), k0, k1);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t sh = 31u - (4u * q + i);
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const uint32_t m = 0u - ((thr[k] >> sh) & 1u);  // T's bit broadcast to all lanes
                    const uint32_t d = eq[k] & (ww[i] ^ m);         // lanes whose U differs from T here
                    lt[k] |= d & m;                                 // T has 1, U has 0  ->  U < T
                    eq[k] &= ~d;
                }
            }
        }
    }
    uint32_t found = 0;
    p0 = 0;
    p1 = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t le = (k < C) ? (lt[k < C ? k : 0] | eq[k < C ? k : 0]) : 0xFFFFFFFFu;
        const uint32_t sel = le & ~found;
        found |= le;
        if (k & 1) p0 |= sel;
        if (k & 2) p1 |= sel;
    }
    p0 &= valid;
    p1 &= valid;
}

__device__ __forceinline__ void draw_group_k(int K, uint32_t g, uint64_t row, uint32_t k0, uint32_t k1,
                                             const uint32_t* thr, uint32_t valid, uint32_t& p0, uint32_t& p1) {
    switch (K) {
        case 1: p0 = 0; p1 = 0; break;
        case 2: draw_group<2>(g, row, k0, k1, thr, valid, p0, p1); break;
        case 3: draw_group<3>(g, row, k0, k1, thr, valid, p0, p1); break;
        default: draw_group<4>(g, row, k0, k1, thr, valid, p0, p1); break;
    }
}

// ------------------------------------------------------------------------------------------------
// CRC32 (zlib polynomial, reflected).  crc32(D) = ~( shift_n(0xFFFFFFFF) ^ L(D) ) with L the pure linear
// part (register starts at 0, no final xor) and shift_k(c) = c * x^(8k) mod P.  L(A||B) = shift_|B|(L(A)) ^ L(B),
// which is what lets every thread checksum its own slice and the slices be merged afterwards.
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        p ^= b & (0u - (a >> 31));
        a <<= 1;
        b = (b >> 1) ^ (kCrcPoly & (0u - (b & 1u)));
    }
    return p;
}

__device__ __forceinline__ uint32_t warp_xor(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// deflate length codes: symbol 257+i covers lengths kLenBase[i] .. with kLenExtra[i] extra bits
__constant__ uint16_t c_len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                        31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// index 0..28 of the length code for a match length 3..258
__device__ __forceinline__ int len_code_index(int len) {
    if (len == 258) return 28;
    const int l = len - 3;  // 0..254
    if (l < 8) return l;
    const int hb = 31 - __clz(l);  // 3..7
    return 4 * (hb - 1) + ((l >> (hb - 2)) & 3);
}

__device__ __forceinline__ uint32_t bit_reverse(uint32_t code, int len) { return __brev(code) >> (32 - len); }

}  // namespace dnaf
