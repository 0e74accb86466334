// Kernel 3: the generic BGZF block encoder (any bytes in, one BGZF block out per CTA), plus the
// scan / compaction kernels that turn per-block slots into one contiguous stream.
//
// Replaces Bio.bgzf.BgzfWriter._write_block (third party; call sites pop_factory.py:403,405,449,458):
// raw deflate + 18-byte BGZF header + CRC32 + ISIZE.  Compressed bytes are not part of the parity
// contract (only the decompressed text is); the deflate parse used here is "P4":
//   * the block is cut into spans of 256 bytes (first span = [0,phase), so that spans start on genotype
//     cell boundaries); one thread tokenises one span, tokens never cross spans
//   * byte p is "predicted" when it equals byte p-4 (one genotype cell back: "0/0\t" is 4 bytes, "0\t0\t"
//     has period 2); maximal predicted runs of >= 3 bytes become one match (distance 4, length <= 258),
//     everything else is a literal
//   * literals / lengths get a per-block dynamic Huffman code (true Huffman, depth-limited to 15 by
//     frequency scaling); only distance code 3 is ever used
//   * CRC32 is computed per span and merged with the GF(2) shift operator (crc32_combine arithmetic)
#pragma once
#include "dnaf_device.cuh"

namespace dnaf {

struct BlockDesc {
    uint64_t off;    // byte offset of the block's text inside the text buffer
    uint32_t len;    // 1..kBlk
    uint32_t phase;  // length of span 0 (0 = spans start at the block start)
};

constexpr int kNumLit = 286;
constexpr int kHufNodes = 2 * kNumLit;

struct DeflateSmem {
    uint32_t out[kSlot / 4];           // payload bits (zero-initialised, OR-ed into)
    uint32_t hist[kNumLit + 2];        // literal/length frequencies
    uint32_t code[kNumLit + 2];        // reversed code | length << 24
    uint32_t span_bits[257];           // bits per span, then exclusive prefix
    uint32_t crctab[256];
    uint32_t w[kHufNodes];             // Huffman scratch: node weights
    uint16_t parent[kHufNodes];
    uint16_t order[kNumLit + 2];
    uint8_t depth[kHufNodes];
    uint8_t lens[kNumLit + 32];        // code lengths lit/len then dist
    uint8_t cl_sym[kNumLit + 32];      // run-length coded code-length sequence
    uint8_t cl_ext[kNumLit + 32];
    uint32_t cl_hist[19];
    uint32_t cl_code[19];
    uint8_t cl_len[19];
    uint32_t n_match;
    uint32_t header_bits;
    uint32_t total_bits;
    uint32_t crc_acc;
    uint32_t warp_tmp[8];
    alignas(16) uint8_t text[kBlk + 48];
};

// ---- single-thread helpers (thread 0 builds the codes; alphabets here have a few dozen used symbols) ----
// Huffman code lengths for freq[0..nsym), depth <= maxbits.  lens[] gets 0 for unused symbols.
__device__ inline void huff_lengths(const uint32_t* freq, int nsym, int maxbits, uint8_t* lens, DeflateSmem& s) {
    int n = 0;
    for (int i = 0; i < nsym; ++i) {
        lens[i] = 0;
        if (freq[i]) s.order[n++] = (uint16_t)i;
    }
    if (n == 0) return;
    if (n == 1) {
        lens[s.order[0]] = 1;
        return;
    }
    for (int i = 1; i < n; ++i) {  // insertion sort by (freq, symbol)
        const uint16_t v = s.order[i];
        const uint32_t fv = freq[v];
        int j = i - 1;
        while (j >= 0 && freq[s.order[j]] > fv) {
            s.order[j + 1] = s.order[j];
            --j;
        }
        s.order[j + 1] = v;
    }
    for (int shift = 0;; ++shift) {
        for (int i = 0; i < n; ++i) {
            const uint32_t f = freq[s.order[i]] >> shift;
            s.w[i] = f ? f : 1u;
        }
        int li = 0, ni = n, next = n;  // leaf queue head, internal queue head, next internal node
        while (next < 2 * n - 1) {
            int pick[2];
            for (int k = 0; k < 2; ++k) {
                if (li < n && (ni >= next || s.w[li] <= s.w[ni])) pick[k] = li++;
                else pick[k] = ni++;
            }
            s.w[next] = s.w[pick[0]] + s.w[pick[1]];
            s.parent[pick[0]] = (uint16_t)next;
            s.parent[pick[1]] = (uint16_t)next;
            ++next;
        }
        s.depth[2 * n - 2] = 0;
        int maxd = 0;
        for (int k = 2 * n - 3; k >= 0; --k) {
            const int d = s.depth[s.parent[k]] + 1;
            s.depth[k] = (uint8_t)d;
            if (k < n && d > maxd) maxd = d;
        }
        if (maxd <= maxbits) break;
    }
    for (int i = 0; i < n; ++i) lens[s.order[i]] = s.depth[i];
}

// Canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first stream; out[i] = code | len << 24.
__device__ inline void huff_codes(const uint8_t* lens, int nsym, uint32_t* out) {
    uint32_t count[16], next[16];
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int i = 0; i < nsym; ++i) count[lens[i]]++;
    count[0] = 0;
    uint32_t code = 0;
    next[0] = 0;
    for (int b = 1; b < 16; ++b) {
        code = (code + count[b - 1]) << 1;
        next[b] = code;
    }
    for (int i = 0; i < nsym; ++i) {
        const int l = lens[i];
        out[i] = l ? (bit_reverse(next[l]++, l) | ((uint32_t)l << 24)) : 0u;
    }
}

struct BitWriter {  // thread-private sequential writer (header); ORs into zeroed words
    uint32_t* words;
    uint32_t pos;
    __device__ void put(uint32_t v, int n) {
        if (n == 0) return;
        const uint32_t wi = pos >> 5, sh = pos & 31u;
        words[wi] |= v << sh;
        if (sh + n > 32) words[wi + 1] |= v >> (32 - sh);
        pos += n;
    }
};

// Builds both Huffman codes and writes the dynamic block header; returns the header length in bits.
__device__ inline uint32_t build_codes_and_header(DeflateSmem& s) {
    s.hist[256] = 1;  // end of block
    huff_lengths(s.hist, kNumLit, 15, s.lens, s);
    huff_codes(s.lens, kNumLit, s.code);
    int nlit = kNumLit;
    while (nlit > 257 && s.lens[nlit - 1] == 0) --nlit;
    const int ndist = s.n_match ? 4 : 1;
    for (int i = 0; i < ndist; ++i) s.lens[nlit + i] = 0;
    if (s.n_match) s.lens[nlit + 3] = 1;  // only distance 4 (code 3); a lone 1-bit code is legal
    const int total = nlit + ndist;
    // run-length code the length sequence (RFC 1951 3.2.7)
    int m = 0;
    for (int i = 0; i < 19; ++i) s.cl_hist[i] = 0;
    for (int i = 0; i < total;) {
        const int v = s.lens[i];
        int run = 1;
        while (i + run < total && s.lens[i + run] == v) ++run;
        i += run;
        if (v == 0) {
            while (run >= 11) {
                const int c = run > 138 ? 138 : run;
                s.cl_sym[m] = 18; s.cl_ext[m++] = (uint8_t)(c - 11); s.cl_hist[18]++;
                run -= c;
            }
            if (run >= 3) {
                s.cl_sym[m] = 17; s.cl_ext[m++] = (uint8_t)(run - 3); s.cl_hist[17]++;
                run = 0;
            }
            while (run-- > 0) { s.cl_sym[m] = 0; s.cl_ext[m++] = 0; s.cl_hist[0]++; }
        } else {
            s.cl_sym[m] = (uint8_t)v; s.cl_ext[m++] = 0; s.cl_hist[v]++;
            --run;
            while (run >= 3) {
                const int c = run > 6 ? 6 : run;
                s.cl_sym[m] = 16; s.cl_ext[m++] = (uint8_t)(c - 3); s.cl_hist[16]++;
                run -= c;
            }
            while (run-- > 0) { s.cl_sym[m] = (uint8_t)v; s.cl_ext[m++] = 0; s.cl_hist[v]++; }
        }
    }
    huff_lengths(s.cl_hist, 19, 7, s.cl_len, s);
    {   // the code-length code must be complete: a single used symbol needs a partner
        int used = 0, only = 0;
        for (int i = 0; i < 19; ++i)
            if (s.cl_len[i]) { ++used; only = i; }
        if (used == 1) s.cl_len[only == 0 ? 1 : 0] = 1;
    }
    huff_codes(s.cl_len, 19, s.cl_code);
    int ncl = 19;
    while (ncl > 4 && s.cl_len[c_cl_order[ncl - 1]] == 0) --ncl;
    BitWriter bw{s.out, 0};
    bw.put(1, 1);  // BFINAL
    bw.put(2, 2);  // BTYPE = dynamic
    bw.put(nlit - 257, 5);
    bw.put(ndist - 1, 5);
    bw.put(ncl - 4, 4);
    for (int i = 0; i < ncl; ++i) bw.put(s.cl_len[c_cl_order[i]], 3);
    for (int i = 0; i < m; ++i) {
        const int sym = s.cl_sym[i];
        bw.put(s.cl_code[sym] & 0xFFFFFFu, (int)(s.cl_code[sym] >> 24));
        if (sym == 16) bw.put(s.cl_ext[i], 2);
        else if (sym == 17) bw.put(s.cl_ext[i], 3);
        else if (sym == 18) bw.put(s.cl_ext[i], 7);
    }
    return bw.pos;
}

// ---- the P4 tokeniser: calls sink.lit(byte) / sink.match(len) for bytes [b,e) of t ----
template <class Sink>
__device__ __forceinline__ void tokenize_span(const uint8_t* t, int b, int e, Sink& sink) {
    int p = b;
    while (p < e) {
        if (p >= 4 && t[p] == t[p - 4]) {
            int q = p + 1;
            while (q < e && t[q] == t[q - 4]) ++q;
            int run = q - p;
            while (run >= 3) {
                const int c = run > 258 ? 258 : run;
                sink.match(c);
                p += c;
                run -= c;
            }
            while (run-- > 0) sink.lit(t[p++]);
        } else {
            sink.lit(t[p++]);
        }
    }
}

struct HistSink {
    uint32_t* hist;
    uint32_t matches;
    __device__ void lit(uint8_t b) { atomicAdd(&hist[b], 1u); }
    __device__ void match(int len) {
        atomicAdd(&hist[257 + len_code_index(len)], 1u);
        ++matches;
    }
};
struct CountSink {
    const uint32_t* code;
    uint32_t bits;
    __device__ void lit(uint8_t b) { bits += code[b] >> 24; }
    __device__ void match(int len) {
        const int ci = len_code_index(len);
        bits += (code[257 + ci] >> 24) + c_len_extra[ci] + 1;
    }
};
struct EmitSink {
    const uint32_t* code;
    uint32_t* out;
    uint32_t pos;
    __device__ void put(uint32_t v, int n) {
        const uint32_t wi = pos >> 5, sh = pos & 31u;
        atomicOr(&out[wi], v << sh);
        if (sh + n > 32) atomicOr(&out[wi + 1], v >> (32 - sh));
        pos += n;
    }
    __device__ void lit(uint8_t b) { put(code[b] & 0xFFFFFFu, (int)(code[b] >> 24)); }
    __device__ void match(int len) {
        const int ci = len_code_index(len);
        const uint32_t c = code[257 + ci];
        const int cl = (int)(c >> 24), ex = c_len_extra[ci];
        // length code, extra bits, then the 1-bit distance code (value 0)
        put((c & 0xFFFFFFu) | ((uint32_t)(len - c_len_base[ci]) << cl), cl + ex + 1);
    }
};

__device__ __forceinline__ void span_bounds(uint32_t t, uint32_t len, uint32_t phase, int& b, int& e) {
    // span 0 = [0,phase) (empty when phase == 0), span k>=1 = 256-byte windows after it
    if (phase == 0) {
        b = (int)min(len, t * kSpan);
        e = (int)min(len, (t + 1) * kSpan);
    } else if (t == 0) {
        b = 0;
        e = (int)min(len, phase);
    } else {
        b = (int)min(len, phase + (t - 1) * kSpan);
        e = (int)min(len, phase + t * kSpan);
    }
}

// One CTA (256 threads) per BGZF block.
__global__ void __launch_bounds__(256, 1) k_bgzf_generic(const uint8_t* __restrict__ text, const BlockDesc* __restrict__ blocks,
                                                        const uint32_t* __restrict__ slot_idx,
                                                        const uint32_t* __restrict__ g_crctab,
                                                        const uint32_t* __restrict__ g_xpow8, uint8_t* __restrict__ slots,
                                                        uint32_t slot_stride, uint32_t* __restrict__ sizes,
                                                        uint32_t* __restrict__ crcs) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    DeflateSmem& s = *reinterpret_cast<DeflateSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const BlockDesc bd = blocks[blockIdx.x];
    const uint32_t n = bd.len;

    // -- stage the block's text in shared memory (16-byte loads from the aligned-down base)
    const uint8_t* src = text + bd.off;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4* src16 = reinterpret_cast<const uint4*>(src - mis);
    uint4* dst16 = reinterpret_cast<uint4*>(s.text);
    const uint32_t n16 = (mis + n + 15u) / 16u;
    for (uint32_t i = tid; i < n16; i += 256) dst16[i] = src16[i];
    const uint8_t* t = s.text + mis;
    for (uint32_t i = tid; i < kSlot / 4; i += 256) s.out[i] = 0;
    for (uint32_t i = tid; i < kNumLit + 2; i += 256) s.hist[i] = 0;
    s.crctab[tid] = g_crctab[tid];
    if (tid == 0) { s.n_match = 0; s.crc_acc = 0; }
    __syncthreads();

    int b, e;
    span_bounds(tid, n, bd.phase, b, e);

    // -- pass 1: symbol statistics
    HistSink hs{s.hist, 0};
    tokenize_span(t, b, e, hs);
    if (hs.matches) atomicAdd(&s.n_match, hs.matches);
    // -- CRC of this span, shifted to the end of the block
    uint32_t crc = 0;
    for (int p = b; p < e; ++p) crc = s.crctab[(crc ^ t[p]) & 0xFFu] ^ (crc >> 8);
    if (crc) crc = gf2_mulmod(g_xpow8[n - (uint32_t)e], crc);
    crc = warp_xor(crc);
    if ((tid & 31u) == 0 && crc) atomicXor(&s.crc_acc, crc);
    __syncthreads();

    if (tid == 0) s.header_bits = build_codes_and_header(s);
    __syncthreads();

    // -- pass 2: bits per span, exclusive scan
    CountSink cs{s.code, 0};
    tokenize_span(t, b, e, cs);
    {
        uint32_t v = cs.bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if ((tid & 31u) >= (uint32_t)o) v += u;
        }
        if ((tid & 31u) == 31u) s.warp_tmp[tid >> 5] = v;
        __syncthreads();
        uint32_t base = s.header_bits;
        for (uint32_t wv = 0; wv < (tid >> 5); ++wv) base += s.warp_tmp[wv];
        s.span_bits[tid] = base + v - cs.bits;
        if (tid == 255) s.total_bits = base + v;
    }
    __syncthreads();
    const uint32_t eob = s.code[256];
    const uint32_t data_bits = s.total_bits + (eob >> 24);
    const uint32_t payload = (data_bits + 7u) / 8u;
    const bool stored = payload > n + 5u || payload > kSlot - 26u;

    const uint32_t slot_no = slot_idx ? slot_idx[blockIdx.x] : blockIdx.x;
    uint8_t* slot = slots + (uint64_t)slot_no * slot_stride + kSlotLead;  // see dnaf_device.cuh
    uint32_t out_payload;
    if (!stored) {
        // -- pass 3: emit
        EmitSink es{s.code, s.out, s.span_bits[tid]};
        tokenize_span(t, b, e, es);
        if (tid == 255) es.put(eob & 0xFFFFFFu, (int)(eob >> 24));
        __syncthreads();
        const uint8_t* ob = reinterpret_cast<const uint8_t*>(s.out);
        for (uint32_t i = tid; i < payload; i += 256) slot[18 + i] = ob[i];
        out_payload = payload;
    } else {
        // incompressible input: one stored block (n <= 65280 < 65535)
        if (tid == 0) {
            slot[18] = 1;
            slot[19] = (uint8_t)n; slot[20] = (uint8_t)(n >> 8);
            slot[21] = (uint8_t)~n; slot[22] = (uint8_t)((~n) >> 8);
        }
        for (uint32_t i = tid; i < n; i += 256) slot[23 + i] = t[i];
        out_payload = n + 5u;
    }
    if (tid == 0) {
        const uint32_t crc32 = ~(gf2_mulmod(g_xpow8[n], 0xFFFFFFFFu) ^ s.crc_acc);
        const uint32_t bsize = out_payload + 25u;
        const uint8_t head[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00,
                                  (uint8_t)bsize, (uint8_t)(bsize >> 8)};
        for (int i = 0; i < 18; ++i) slot[i] = head[i];
        uint8_t* tail = slot + 18 + out_payload;
        for (int i = 0; i < 4; ++i) tail[i] = (uint8_t)(crc32 >> (8 * i));
        for (int i = 0; i < 4; ++i) tail[4 + i] = (uint8_t)(n >> (8 * i));
        sizes[slot_no] = out_payload + 26u;
        crcs[slot_no] = crc32;
    }
}

// Copies one block's bytes from its slot to its place in the contiguous stream.
__device__ __forceinline__ void copy_block(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t n,
                                           uint32_t tid, uint32_t nthr) {
    // 16-byte stores to the destination; the source is read as ALIGNED 16-byte words too and shifted into place
    // (reads may run up to 31 bytes past the block's last byte: still inside its 64 KiB slot)
    const uint32_t head = min(n, (uint32_t)((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    for (uint32_t i = tid; i < head; i += nthr) dst[i] = src[i];
    const uint32_t body = (n - head) / 16u;
    uint4* d16 = reinterpret_cast<uint4*>(dst + head);
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src + head);
    const uint4* s16 = reinterpret_cast<const uint4*>(sa & ~uintptr_t(15));
    const uint32_t mis = (uint32_t)(sa & 15u), wsh = mis >> 2, bsh = (mis & 3u) * 8u;   // uniform over the block
#pragma unroll 4
    for (uint32_t i = tid; i < body; i += nthr) {
        const uint4 A = s16[i];
        uint4 B = A;
        if (mis) B = s16[i + 1];
        const uint32_t w[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, 0u};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t lo = wsh == 0 ? w[k] : (wsh == 1 ? w[k + 1] : (wsh == 2 ? w[k + 2] : w[k + 3]));
            const uint32_t hi = wsh == 0 ? w[k + 1] : (wsh == 1 ? w[k + 2] : (wsh == 2 ? w[k + 3] : w[k + 4]));
            o[k] = __funnelshift_r(lo, hi, bsh);
        }
        d16[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    for (uint32_t i = head + body * 16u + tid; i < n; i += nthr) dst[i] = src[i];
}

// Scan + gather without any inter-CTA dependency:
//   k_size_partials: one CTA per 256 slots sums their byte counts and xors their CRC32s
//   k_gather:        a CTA takes kTile slots; the bytes in front of them are the partial sums in front of their
//                    256-slot group plus the sizes inside the group in front of the tile -- at most 105 + 255
//                    values, read in parallel -- then one warp per slot copies.  The CTA of the last tile stores
//                    the two totals to `host_totals` (mapped page-locked memory): a DMA read-back would queue
//                    behind the previous pass's data copy on the device-to-host copy engine.
// state: [g] bytes of slot group g, [ngroups + g] xor of its CRC32s (every word is written before it is read).
constexpr uint32_t kTile = 8;
constexpr uint32_t kGroup = 256;
__global__ void __launch_bounds__(kGroup) k_size_partials(const uint32_t* __restrict__ sizes, const uint32_t* __restrict__ crcs,
                                                         uint32_t nb, unsigned long long* __restrict__ state) {
    __shared__ uint32_t s_sum[8], s_xor[8];
    const uint32_t b = blockIdx.x * kGroup + threadIdx.x;
    uint32_t v = b < nb ? sizes[b] : 0u, x = b < nb ? crcs[b] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    x = warp_xor(x);
    if ((threadIdx.x & 31u) == 0) { s_sum[threadIdx.x >> 5] = v; s_xor[threadIdx.x >> 5] = x; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        uint32_t xx = 0;
        for (int w = 0; w < 8; ++w) { t += s_sum[w]; xx ^= s_xor[w]; }
        state[blockIdx.x] = t;
        state[gridDim.x + blockIdx.x] = xx;
    }
}

__global__ void __launch_bounds__(256) k_gather(const uint8_t* __restrict__ slots, uint32_t slot_stride,
                                               const uint32_t* __restrict__ sizes, uint32_t nb,
                                               const unsigned long long* __restrict__ state,
                                               volatile unsigned long long* __restrict__ host_totals,
                                               uint8_t* __restrict__ out) {
    __shared__ uint64_t s_part[8];
    __shared__ uint32_t s_xor[8];
    __shared__ uint64_t s_off[kTile];
    __shared__ uint32_t s_size[kTile];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t tile = blockIdx.x, first = tile * kTile, group = first / kGroup, g0 = group * kGroup;
    // bytes of the slot groups in front (strided over the CTA), plus the slots of this group in front of the tile
    uint64_t acc = 0;
    for (uint32_t g = tid; g < group; g += 256u) acc += state[g];
    const uint32_t b = g0 + tid;
    if (b < first) acc += sizes[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) s_part[wid] = acc;
    const uint32_t sz = (tid < kTile && first + tid < nb) ? sizes[first + tid] : 0u;
    const bool last_tile = first + kTile >= nb;
    const uint32_t ngroups = (nb + kGroup - 1u) / kGroup;
    uint32_t xx = 0;
    if (last_tile) {   // checksum of checksums
        for (uint32_t g = tid; g < ngroups; g += 256u) xx ^= (uint32_t)state[ngroups + g];
        xx = warp_xor(xx);
        if (lane == 0) s_xor[wid] = xx;
    }
    __syncthreads();
    if (tid < 32) {
        uint64_t before = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) before += s_part[w];
        uint32_t v = sz;
#pragma unroll
        for (int o = 1; o < (int)kTile; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (tid >= (uint32_t)o) v += u;
        }
        if (tid < kTile) {
            s_off[tid] = before + v - sz;
            s_size[tid] = sz;
        }
        if (tid == kTile - 1u && last_tile) {   // totals
            uint32_t x = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) x ^= s_xor[w];
            host_totals[0] = before + v;
            host_totals[1] = x;
            __threadfence_system();
        }
    }
    __syncthreads();
    {   // one warp per slot of the tile: eight independent copies in flight per CTA
        const uint32_t k = wid, bb = first + k;
        if (bb < nb) copy_block(slots + (uint64_t)bb * slot_stride + kSlotLead, out + s_off[k], s_size[k], lane, 32u);
    }
}

}  // namespace dnaf
