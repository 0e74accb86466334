// Declarations shared by the fused kernels (k_auto.cuh, k_x.cuh, k_fused_text.cuh) and the host's table builder
// (fused_host.h): the block descriptor the planner fills, the static code table of k_fused_text, literal ids.
#pragma once
#include "dnaf_device.cuh"
#include "k_sample_format.cuh"

namespace dnaf {

// literal symbol ids of the five bytes a genotype cell can hold
enum : int { kLit0 = 0, kLit1 = 1, kLitSlash = 2, kLitTab = 3, kLitNl = 4 };

constexpr int kFusedMaxThreads = 256;

// One BGZF block of a row whose spans are whole groups of 64 samples (k_auto, k_x).
struct FusedDesc {
    uint64_t row;        // global row
    uint32_t cell0;      // first sample of the segment (multiple of 64)
    uint32_t ncells;     // samples in the segment (<= 254 * 64)
    uint32_t slot;       // output slot / block index inside the pass
    uint32_t flags;      // bit0: segment starts the row (has the prefix), bit1: segment ends the row
    uint32_t ovr_first;  // overrides of this row: [ovr_first, ovr_first + ovr_count)
    uint32_t ovr_count;
    uint32_t table;      // index into the kernel's table array
    uint32_t body_crc;   // L(all-reference template body of this segment)
};

// Static codes of one (MAF bucket, chromosome class, with/without prefix) for k_fused_text; also the
// intermediate form the host derives AutoTable / XTable from.
struct FusedTable {
    uint32_t len_tok[260];   // match length -> (code | extra | distance bit) | total bits << 24
    uint32_t lit[8];         // cell literals by id: code | bits << 24
    uint32_t eob;            // end-of-block code | bits << 24
    uint32_t hdr_bits;       // serialized dynamic-block header
    uint32_t hdr[62];
    uint32_t pre_lit[256];   // literal codes for every byte value (prefix bytes; all text bytes for k_fused_text)
};

__host__ __device__ __forceinline__ uint32_t pick4(const uint32_t m[4], int w) {
    return w == 0 ? m[0] : (w == 1 ? m[1] : (w == 2 ? m[2] : m[3]));
}

}  // namespace dnaf
