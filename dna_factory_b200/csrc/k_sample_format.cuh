// Kernel 1 (allele draws -> packed genotype planes), the override pass, kernel 2 (planes -> VCF text)
// and the genotype export used by the bit-exact parity gate.
//
// Reference behaviour restated (ochrzan/dna-factory):
//   pop_factory.py:477      one uniform per allele slot j = 2*i + s of the SNP row
//   pop_factory.py:92-95    allele = first k with cum_k >= u
//   pop_factory.py:481-499  '.' for females on Y; haploid cells use slot 0 only; forced "1" / "1/1" for
//                           cases whose deleterious set holds the SNP
//   pop_factory.py:503-508  row = 9-column prefix + "\t".join(cells) + "\n"
#pragma once
#include "dnaf_device.cuh"

namespace dnaf {

struct SampleView {
    uint32_t n;            // samples
    uint32_t groups;       // ceil(2n/32) words per row and plane
    const uint8_t* sex;    // 1 = male
    const uint32_t* xoff;  // [n+1] byte offset of sample i inside an X-row body
    uint32_t body[4];      // body bytes per chromosome class (incl. the trailing '\n')
};

struct SnpView {
    const uint8_t* cls;      // chromosome class per global row
    const uint8_t* k;        // alleles per global row
    const uint32_t* thr;     // [row][4]
    const uint8_t* prefix;   // concatenated row prefixes
    const uint64_t* pre_off; // [S+1]
};

// One thread per (row, group of 32 allele slots).  Local row r is global row row0 + (row_idx ? row_idx[r] : r).
__global__ void __launch_bounds__(256) k_sample(SampleView sv, SnpView nv, uint64_t row0, const uint32_t* __restrict__ row_idx,
                                               uint64_t row_base, uint32_t n_rows, uint32_t k0, uint32_t k1,
                                               uint32_t* __restrict__ plane0, uint32_t* __restrict__ plane1) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * sv.groups) return;
    const uint32_t r = (uint32_t)(idx / sv.groups);
    const uint32_t g = (uint32_t)(idx % sv.groups);
    const uint64_t row = row0 + (row_idx ? row_idx[r] : r);
    const uint4 t4 = reinterpret_cast<const uint4*>(nv.thr)[row];
    const uint32_t thr[4] = {t4.x, t4.y, t4.z, t4.w};
    const uint32_t slots = 2u * sv.n - 32u * g;
    const uint32_t valid = slots >= 32u ? 0xFFFFFFFFu : ((1u << slots) - 1u);
    uint32_t p0, p1;
    draw_group_k(nv.k[row], g, row_base + row, k0, k1, thr, valid, p0, p1);
    plane0[idx] = p0;
    if (plane1) plane1[idx] = p1;
}

// Forced-minor cells: allele index 1 in both slots (the formatter ignores slot 1 of haploid cells and
// everything on '.' cells, which is the order of the tests at pop_factory.py:481-499).
// olocal[t] = local row (index into the plane buffers), osamp[t] = sample.
__global__ void k_overrides(const uint32_t* __restrict__ olocal, const uint32_t* __restrict__ osamp, uint64_t count,
                            uint32_t groups, uint32_t n_samples, uint32_t* __restrict__ plane0,
                            uint32_t* __restrict__ plane1) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t r = olocal[t];
    const uint32_t i = osamp[t];
    if (i >= n_samples) return;
    const uint64_t w = r * groups + (i >> 4);
    const uint32_t m = 3u << ((2u * i) & 31u);
    atomicOr(&plane0[w], m);
    if (plane1) atomicAnd(&plane1[w], ~m);
}

__device__ __forceinline__ uint32_t cell_width(uint8_t cls, bool male) {
    return (cls == kAuto || (cls == kX && !male)) ? 4u : 2u;
}

// Kernel 2: one CTA per row; threads stride over samples.  Row r goes to text + sub_off[r] when a row
// subset is formatted (row_idx != null), else to text + row_off[row] - text0.
__global__ void __launch_bounds__(256) k_format(SampleView sv, SnpView nv, uint64_t row0, const uint32_t* __restrict__ row_idx,
                                               const uint64_t* __restrict__ sub_off, const uint64_t* __restrict__ row_off,
                                               uint64_t text0, const uint32_t* __restrict__ plane0,
                                               const uint32_t* __restrict__ plane1, uint8_t* __restrict__ text) {
    const uint32_t r = blockIdx.x;
    const uint64_t row = row0 + (row_idx ? row_idx[r] : r);
    const uint8_t cls = nv.cls[row];
    const uint64_t pb = nv.pre_off[row];
    const uint32_t plen = (uint32_t)(nv.pre_off[row + 1] - pb);
    uint8_t* out = text + (row_idx ? sub_off[r] : row_off[row] - text0);
    for (uint32_t i = threadIdx.x; i < plen; i += blockDim.x) out[i] = nv.prefix[pb + i];
    uint8_t* body = out + plen;
    if (sv.n == 0) {
        if (threadIdx.x == 0) body[0] = '\n';
        return;
    }
    const uint32_t* w0 = plane0 + (uint64_t)r * sv.groups;
    const uint32_t* w1 = plane1 ? plane1 + (uint64_t)r * sv.groups : nullptr;
    const bool aligned4 = (reinterpret_cast<uintptr_t>(body) & 3u) == 0;
    for (uint32_t i = threadIdx.x; i < sv.n; i += blockDim.x) {
        const bool male = sv.sex[i] == 1;
        const uint32_t sh = (2u * i) & 31u;
        uint32_t bits = (w0[i >> 4] >> sh) & 3u;
        uint32_t a = bits & 1u, b = bits >> 1;
        if (w1) {
            const uint32_t hb = (w1[i >> 4] >> sh) & 3u;
            a |= (hb & 1u) << 1;
            b |= (hb >> 1) << 1;
        }
        const uint8_t term = (i + 1 == sv.n) ? '\n' : '\t';
        if (cls == kAuto) {
            const uint32_t word = ('0' + a) | ('/' << 8) | (('0' + b) << 16) | ((uint32_t)term << 24);
            uint8_t* p = body + 4ull * i;
            if (aligned4) {
                *reinterpret_cast<uint32_t*>(p) = word;
            } else {
                p[0] = (uint8_t)word; p[1] = '/'; p[2] = (uint8_t)(word >> 16); p[3] = term;
            }
        } else if (cls == kX) {
            uint8_t* p = body + sv.xoff[i];
            if (male) {
                p[0] = '0' + a; p[1] = term;
            } else {
                p[0] = '0' + a; p[1] = '/'; p[2] = '0' + b; p[3] = term;
            }
        } else {
            uint8_t* p = body + 2ull * i;
            p[0] = (cls == kY && !male) ? '.' : (uint8_t)('0' + a);
            p[1] = term;
        }
    }
}

// Parity gate export: out[(r*n+i)*2+s] = allele index or 0xFF.
__global__ void __launch_bounds__(256) k_export_genotypes(SampleView sv, SnpView nv, uint64_t row0, uint32_t n_rows,
                                                         const uint32_t* __restrict__ plane0,
                                                         const uint32_t* __restrict__ plane1, uint8_t* __restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * sv.n) return;
    const uint32_t r = (uint32_t)(idx / sv.n);
    const uint32_t i = (uint32_t)(idx % sv.n);
    const uint8_t cls = nv.cls[row0 + r];
    const bool male = sv.sex[i] == 1;
    const uint32_t sh = (2u * i) & 31u;
    const uint64_t w = (uint64_t)r * sv.groups + (i >> 4);
    const uint32_t bits = (plane0[w] >> sh) & 3u;
    uint32_t a = bits & 1u, b = bits >> 1;
    if (plane1) {
        const uint32_t hb = (plane1[w] >> sh) & 3u;
        a |= (hb & 1u) << 1;
        b |= (hb >> 1) << 1;
    }
    uint8_t o0 = (uint8_t)a, o1 = (uint8_t)b;
    if (cls == kY && !male) { o0 = 0xFF; o1 = 0xFF; }
    else if (cell_width(cls, male) == 2) o1 = 0xFF;
    out[idx * 2] = o0;
    out[idx * 2 + 1] = o1;
}

}  // namespace dnaf
