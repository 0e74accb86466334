// GPU SNP selection: SnpFactory.random_snp_tuples (pop_factory.py:160-193) as an inverse-CDF sampler, and the
// (chromosome string, position) sort of pop_factory.py:245 as a stable radix sort.
//
// Reference behaviour restated, per SNP n (0-based draw index, id = first_id + n), in the reference's draw order:
//   chromosome  numpy.random.choice(CHROMOSOME_LIST, p=CHROMOSOME_PROB)      common/snp.py:8-34, pop_factory.py:169-170
//   maf         numpy.random.choice(sorted_maf[start:], p=pdf[start:]/sum)   pop_factory.py:160-167
//   position    int(numpy.random.random() * CHROMOSOME_MAX_POSITION[c])      pop_factory.py:181,186
//   ref         numpy.random.choice(["A","T","C","G"])                       pop_factory.py:182
//   alt         random.choice(the three others, A,T,C,G order)               pop_factory.py:188-190
// numpy's choice(p=...) is  cdf = p.cumsum(); cdf /= cdf[-1]; cdf.searchsorted(u, side="right")  and its
// choice without p is randint(len): the host passes the two cdf arrays computed with numpy itself, the kernel
// counts the entries <= u in float64, exactly as searchsorted does.
//
// Replay RNG spec for this stream: for draw n
//   A = philox4x32_10(ctr = (n_lo, n_hi, 0x534E5000, 0xFFFFFFFF), key = seed)   u_chrom, u_maf, u_pos, u_ref = A[0..3] * 2^-32
//   B = philox4x32_10(ctr = (n_lo, n_hi, 0x534E5001, 0xFFFFFFFF), key = seed)   u_alt = B[0] * 2^-32
// (counter word 3 = 0xFFFFFFFF keeps the stream disjoint from the genotype stream, whose word 3 is row >> 32).
#pragma once
#include <cub/cub.cuh>

#include "dnaf_device.cuh"

namespace dnaf {

constexpr uint32_t kSelTagA = 0x534E5000u, kSelTagB = 0x534E5001u;
constexpr int kSelMaxChrom = 64, kSelMaxMaf = 256;

struct SelectArgs {
    uint64_t n;
    uint32_t k0, k1;
    uint32_t n_chrom, n_maf;
    const double* chrom_cdf;      // [n_chrom]
    const double* chrom_max_pos;  // [n_chrom]
    const uint8_t* chrom_rank;    // [n_chrom] rank of the label in string order
    const double* maf_cdf;        // [n_maf]
    uint64_t* key;                // [n] rank << 32 | position
    uint32_t* idx;                // [n] draw index
    uint8_t* chrom;               // [n] unsorted columns
    uint8_t* maf;
    uint32_t* pos;
    uint8_t* ref;
    uint8_t* alt;
};

__global__ void __launch_bounds__(256) k_select_snps(const SelectArgs a) {
    __shared__ double s_ccdf[kSelMaxChrom], s_cmax[kSelMaxChrom], s_mcdf[kSelMaxMaf];
    __shared__ uint8_t s_rank[kSelMaxChrom];
    for (uint32_t i = threadIdx.x; i < a.n_chrom; i += blockDim.x) {
        s_ccdf[i] = a.chrom_cdf[i];
        s_cmax[i] = a.chrom_max_pos[i];
        s_rank[i] = a.chrom_rank[i];
    }
    for (uint32_t i = threadIdx.x; i < a.n_maf; i += blockDim.x) s_mcdf[i] = a.maf_cdf[i];
    __syncthreads();
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n) return;
    const uint4 A = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), kSelTagA, 0xFFFFFFFFu, a.k0, a.k1);
    const uint4 B = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), kSelTagB, 0xFFFFFFFFu, a.k0, a.k1);
    const double s = 1.0 / 4294967296.0;
    const double u_chrom = A.x * s, u_maf = A.y * s, u_pos = A.z * s, u_ref = A.w * s, u_alt = B.x * s;
    uint32_t c = 0, m = 0;
    for (uint32_t i = 0; i < a.n_chrom; ++i) c += s_ccdf[i] <= u_chrom;   // searchsorted(side="right")
    for (uint32_t i = 0; i < a.n_maf; ++i) m += s_mcdf[i] <= u_maf;
    c = min(c, a.n_chrom - 1u);
    m = min(m, a.n_maf - 1u);
    const uint32_t pos = (uint32_t)(long long)(u_pos * s_cmax[c]);        // int(u * max_position): truncation
    const uint32_t ref_idx = min((uint32_t)(u_ref * 4.0), 3u);            // A,T,C,G
    const uint32_t pick = min((uint32_t)(u_alt * 3.0), 2u);
    const uint32_t alt_idx = pick + (pick >= ref_idx ? 1u : 0u);          // list.remove(ref) keeps A,T,C,G order
    const char nt[4] = {'A', 'T', 'C', 'G'};
    a.key[n] = ((uint64_t)s_rank[c] << 32) | pos;
    a.idx[n] = (uint32_t)n;
    a.chrom[n] = (uint8_t)c;
    a.maf[n] = (uint8_t)m;
    a.pos[n] = pos;
    a.ref[n] = (uint8_t)nt[ref_idx];
    a.alt[n] = (uint8_t)nt[alt_idx];
}

// columns in sorted order: out[r] = in[order[r]]
__global__ void __launch_bounds__(256) k_select_gather(uint64_t n, const uint32_t* __restrict__ order,
                                                      const uint8_t* __restrict__ chrom, const uint8_t* __restrict__ maf,
                                                      const uint32_t* __restrict__ pos, const uint8_t* __restrict__ ref,
                                                      const uint8_t* __restrict__ alt, uint8_t* __restrict__ o_chrom,
                                                      uint8_t* __restrict__ o_maf, uint32_t* __restrict__ o_pos,
                                                      uint8_t* __restrict__ o_ref, uint8_t* __restrict__ o_alt) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t j = order[r];
    o_chrom[r] = chrom[j];
    o_maf[r] = maf[j];
    o_pos[r] = pos[j];
    o_ref[r] = ref[j];
    o_alt[r] = alt[j];
}

}  // namespace dnaf
