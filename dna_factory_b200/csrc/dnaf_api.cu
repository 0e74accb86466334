// C ABI (include/dnaf_b200.h) and host orchestration of the B200 hot path.
//
// Replaces, for the population-generation path of ochrzan/dna-factory:
//   PopulationFactory.write_vcf_snps   pop_factory.py:417-469  (worker pool + ordered writer)
//   PopulationFactory.queue_vcf_snps   pop_factory.py:471-513  (row loop)
//   Bio.bgzf.BgzfWriter.write          call site pop_factory.py:449
// Work is cut into passes of at most `chunk_bytes` of uncompressed text; each pass is
//   sample -> (overrides) -> format -> BGZF encode -> scan -> compact -> D2H -> sink      (generic path)
// or the fused kernel (k_fused.cuh) followed by scan -> compact -> D2H -> sink.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dnaf_b200.h"
#include "dnaf_device.cuh"
#include "k_deflate.cuh"
#include "k_sample_format.cuh"

using namespace dnaf;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { release(); }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

uint32_t host_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : (b >> 1);
    }
    return p;
}

}  // namespace

struct dnaf_ctx {
    int dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t chunk_bytes = 256ull << 20;
    int fused = 1;
    uint64_t row_base = 0;

    // samples
    bool have_samples = false;
    uint32_t n = 0, males = 0;
    DevBuf d_sex, d_xoff;
    uint32_t body[4] = {1, 1, 1, 1};

    // snps
    bool have_snps = false;
    uint64_t S = 0;
    bool any_multi = false;
    DevBuf d_cls, d_k, d_thr, d_prefix, d_pre_off, d_row_off;
    std::vector<uint8_t> h_cls;
    std::vector<uint32_t> h_plen;
    std::vector<uint64_t> h_row_off;  // valid when layout_ok
    bool layout_ok = false;

    // overrides
    uint64_t P = 0;
    DevBuf d_orow, d_osamp;
    std::vector<uint64_t> h_orow;

    // constant tables
    DevBuf d_crctab, d_xpow8;

    // scratch
    DevBuf d_plane0, d_plane1, d_text, d_slots, d_sizes, d_crcs, d_offsets, d_totals, d_blocks, d_out, d_geno;
    PinnedBuf h_out, h_totals, h_blocks;
    std::vector<BlockDesc> plan;

    cudaEvent_t ev[8] = {};
    bool attr_done = false;
};

namespace {

int fail(dnaf_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail((c), e__ == cudaErrorMemoryAllocation ? DNAF_E_NOMEM : DNAF_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                     \
    } while (0)

template <class T>
int upload(dnaf_ctx* c, DevBuf& b, const T* src, size_t count) {
    CU(c, b.reserve(std::max<size_t>(count, 1) * sizeof(T) + 64));
    if (count) CU(c, cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return DNAF_OK;
}

SampleView sample_view(const dnaf_ctx* c) {
    SampleView v;
    v.n = c->n;
    v.groups = (2u * c->n + 31u) / 32u;
    v.sex = c->d_sex.as<uint8_t>();
    v.xoff = c->d_xoff.as<uint32_t>();
    for (int i = 0; i < 4; ++i) v.body[i] = c->body[i];
    return v;
}

SnpView snp_view(const dnaf_ctx* c) {
    SnpView v;
    v.cls = c->d_cls.as<uint8_t>();
    v.k = c->d_k.as<uint8_t>();
    v.thr = c->d_thr.as<uint32_t>();
    v.prefix = c->d_prefix.as<uint8_t>();
    v.pre_off = c->d_pre_off.as<uint64_t>();
    return v;
}

// Text offset of every row (prefix + class body), host and device copies.
int ensure_layout(dnaf_ctx* c) {
    if (!c->have_samples || !c->have_snps) return fail(c, DNAF_E_ARG, "set_samples and set_snps must be called first");
    if (c->layout_ok) return DNAF_OK;
    c->h_row_off.resize(c->S + 1);
    uint64_t acc = 0;
    for (uint64_t r = 0; r < c->S; ++r) {
        c->h_row_off[r] = acc;
        acc += (uint64_t)c->h_plen[r] + c->body[c->h_cls[r]];
    }
    c->h_row_off[c->S] = acc;
    int rc = upload(c, c->d_row_off, c->h_row_off.data(), c->S + 1);
    if (rc) return rc;
    c->layout_ok = true;
    return DNAF_OK;
}

inline uint32_t row_len(const dnaf_ctx* c, uint64_t r) { return (uint32_t)(c->h_row_off[r + 1] - c->h_row_off[r]); }

// BGZF block plan for rows [r0,r1); offsets relative to the text of row r0.  See k_deflate.cuh.
void plan_blocks(dnaf_ctx* c, uint64_t r0, uint64_t r1) {
    c->plan.clear();
    const uint64_t text0 = c->h_row_off[r0];
    uint64_t r = r0;
    while (r < r1) {
        const uint64_t off = c->h_row_off[r] - text0;
        const uint64_t len = c->h_row_off[r + 1] - c->h_row_off[r];
        const uint32_t plen = c->h_plen[r];
        if (len > kBlk) {
            uint64_t done = 0;
            if (plen + kSpan <= kBlk) {  // first segment: prefix + whole 256-byte spans of the body
                const uint64_t first = plen + (uint64_t)((kBlk - plen) / kSpan) * kSpan;
                c->plan.push_back({off, (uint32_t)std::min<uint64_t>(first, len), plen});
                done = std::min<uint64_t>(first, len);
            }
            while (done < len) {
                const uint32_t piece = (uint32_t)std::min<uint64_t>(kBlk, len - done);
                c->plan.push_back({off + done, piece, 0});
                done += piece;
            }
            ++r;
        } else {
            uint64_t acc = 0;
            while (r < r1 && acc + (c->h_row_off[r + 1] - c->h_row_off[r]) <= kBlk) {
                acc += c->h_row_off[r + 1] - c->h_row_off[r];
                ++r;
            }
            c->plan.push_back({off, (uint32_t)acc, std::min<uint32_t>(plen, (uint32_t)acc)});
        }
    }
}

struct Sink {
    dnaf_sink_fn fn = nullptr;
    void* user = nullptr;
    uint8_t* buf = nullptr;  // host buffer mode
    uint64_t cap = 0, used = 0;
    bool device_only = false;
};

int deliver(dnaf_ctx* c, Sink& s, const uint8_t* data, uint64_t n) {
    if (s.fn) {
        if (s.fn(s.user, data, n) != 0) return fail(c, DNAF_E_SINK, "sink callback failed");
    } else if (s.buf) {
        if (s.used + n > s.cap) return fail(c, DNAF_E_SPACE, "output buffer too small: need more than %llu bytes",
                                            (unsigned long long)s.cap);
        memcpy(s.buf + s.used, data, n);
    }
    s.used += n;
    return DNAF_OK;
}

// Encode the blocks in c->plan from d_text, compact, and hand the bytes to the sink.
int encode_plan(dnaf_ctx* c, Sink& sink, dnaf_stats* st, cudaEvent_t ev_begin, cudaEvent_t ev_end) {
    const uint32_t nb = (uint32_t)c->plan.size();
    if (nb == 0) return DNAF_OK;
    CU(c, c->d_blocks.reserve(nb * sizeof(BlockDesc)));
    CU(c, c->h_blocks.reserve(nb * sizeof(BlockDesc)));
    memcpy(c->h_blocks.p, c->plan.data(), nb * sizeof(BlockDesc));
    CU(c, cudaMemcpyAsync(c->d_blocks.p, c->h_blocks.p, nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, c->stream));
    CU(c, c->d_slots.reserve((size_t)nb * kSlot));
    CU(c, c->d_sizes.reserve(nb * sizeof(uint32_t)));
    CU(c, c->d_crcs.reserve(nb * sizeof(uint32_t)));
    CU(c, c->d_offsets.reserve(nb * sizeof(uint64_t)));
    CU(c, c->d_totals.reserve(2 * sizeof(uint64_t)));
    CU(c, c->d_out.reserve((size_t)nb * kSlot));
    CU(c, c->h_totals.reserve(2 * sizeof(uint64_t)));
    if (!c->attr_done) {
        CU(c, cudaFuncSetAttribute(k_bgzf_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateSmem)));
        c->attr_done = true;
    }
    CU(c, cudaEventRecord(ev_begin, c->stream));
    k_bgzf_generic<<<nb, 256, sizeof(DeflateSmem), c->stream>>>(c->d_text.as<uint8_t>(), c->d_blocks.as<BlockDesc>(),
                                                               c->d_crctab.as<uint32_t>(), c->d_xpow8.as<uint32_t>(),
                                                               c->d_slots.as<uint8_t>(), c->d_sizes.as<uint32_t>(),
                                                               c->d_crcs.as<uint32_t>());
    k_scan_sizes<<<1, 1024, 0, c->stream>>>(c->d_sizes.as<uint32_t>(), nb, c->d_offsets.as<uint64_t>(),
                                            c->d_crcs.as<uint32_t>(), c->d_totals.as<uint64_t>());
    k_compact<<<nb, 256, 0, c->stream>>>(c->d_slots.as<uint8_t>(), kSlot, c->d_sizes.as<uint32_t>(),
                                         c->d_offsets.as<uint64_t>(), c->d_out.as<uint8_t>());
    CU(c, cudaEventRecord(ev_end, c->stream));
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->h_totals.p, c->d_totals.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const uint64_t bytes = c->h_totals.as<uint64_t>()[0];
    const uint32_t cx = (uint32_t)c->h_totals.as<uint64_t>()[1];
    if (st) {
        st->bgzf_bytes += bytes;
        st->bgzf_blocks += nb;
        st->crc_xor ^= cx;
        st->kernel_launches += 3;
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        st->ms_deflate += ms;
    }
    if (!sink.device_only) {
        CU(c, c->h_out.reserve(bytes));
        CU(c, cudaMemcpyAsync(c->h_out.p, c->d_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return deliver(c, sink, c->h_out.as<uint8_t>(), bytes);
    }
    sink.used += bytes;
    return DNAF_OK;
}

// sample (+ overrides) for rows [r0,r1) into the plane buffers
int run_sample(dnaf_ctx* c, uint64_t r0, uint64_t r1, uint64_t seed, dnaf_stats* st) {
    const SampleView sv = sample_view(c);
    const uint32_t rows = (uint32_t)(r1 - r0);
    const uint64_t words = (uint64_t)rows * sv.groups;
    CU(c, c->d_plane0.reserve(std::max<uint64_t>(words, 1) * 4));
    if (c->any_multi) CU(c, c->d_plane1.reserve(std::max<uint64_t>(words, 1) * 4));
    uint32_t* p1 = c->any_multi ? c->d_plane1.as<uint32_t>() : nullptr;
    if (words) {
        const uint32_t grid = (uint32_t)((words + 255) / 256);
        k_sample<<<grid, 256, 0, c->stream>>>(sv, snp_view(c), r0, c->row_base, rows, (uint32_t)seed, (uint32_t)(seed >> 32),
                                              c->d_plane0.as<uint32_t>(), p1);
        if (st) st->kernel_launches += 1;
        const uint64_t o0 = std::lower_bound(c->h_orow.begin(), c->h_orow.end(), r0) - c->h_orow.begin();
        const uint64_t o1 = std::lower_bound(c->h_orow.begin(), c->h_orow.end(), r1) - c->h_orow.begin();
        if (o1 > o0) {
            k_overrides<<<(uint32_t)((o1 - o0 + 255) / 256), 256, 0, c->stream>>>(
                c->d_orow.as<uint64_t>(), c->d_osamp.as<uint32_t>(), o0, o1 - o0, r0, sv.groups, c->n,
                c->d_plane0.as<uint32_t>(), p1);
            if (st) st->kernel_launches += 1;
        }
    }
    CU(c, cudaGetLastError());
    return DNAF_OK;
}

int run_format(dnaf_ctx* c, uint64_t r0, uint64_t r1, dnaf_stats* st) {
    const uint64_t text0 = c->h_row_off[r0];
    const uint64_t bytes = c->h_row_off[r1] - text0;
    CU(c, c->d_text.reserve(bytes + 64));
    const uint32_t rows = (uint32_t)(r1 - r0);
    k_format<<<rows, 256, 0, c->stream>>>(sample_view(c), snp_view(c), r0, c->d_row_off.as<uint64_t>(), text0,
                                          c->d_plane0.as<uint32_t>(),
                                          c->any_multi ? c->d_plane1.as<uint32_t>() : nullptr, c->d_text.as<uint8_t>());
    if (st) st->kernel_launches += 1;
    CU(c, cudaGetLastError());
    return DNAF_OK;
}

uint64_t next_chunk_end(const dnaf_ctx* c, uint64_t r0, uint64_t row_end, uint64_t budget) {
    const uint64_t lim = c->h_row_off[r0] + budget;
    uint64_t r1 = std::upper_bound(c->h_row_off.begin() + r0, c->h_row_off.begin() + row_end + 1, lim) -
                  c->h_row_off.begin() - 1;
    if (r1 <= r0) r1 = r0 + 1;
    return std::min(r1, row_end);
}

int generate_impl(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int rng_mode, int level,
                  Sink& sink, dnaf_stats* st) {
    if (!c) return DNAF_E_ARG;
    if (rng_mode != 0 && rng_mode != 1) return fail(c, DNAF_E_ARG, "rng_mode must be 0 (replay) or 1 (native)");
    if (level < 1 || level > 9) return fail(c, DNAF_E_ARG, "level must be 1..9");
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    CU(c, cudaSetDevice(c->dev));
    dnaf_stats local;
    memset(&local, 0, sizeof local);
    uint64_t r0 = row_begin;
    while (r0 < row_end) {
        const uint64_t r1 = next_chunk_end(c, r0, row_end, c->chunk_bytes);
        CU(c, cudaEventRecord(c->ev[0], c->stream));
        rc = run_sample(c, r0, r1, seed, &local);
        if (rc) return rc;
        CU(c, cudaEventRecord(c->ev[1], c->stream));
        rc = run_format(c, r0, r1, &local);
        if (rc) return rc;
        CU(c, cudaEventRecord(c->ev[2], c->stream));
        plan_blocks(c, r0, r1);
        rc = encode_plan(c, sink, &local, c->ev[3], c->ev[4]);
        if (rc) return rc;
        float a = 0, b = 0, t = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&t, c->ev[0], c->ev[4]);
        local.ms_sample += a;
        local.ms_format += b;
        local.ms_total += t;
        local.rows += r1 - r0;
        local.text_bytes += c->h_row_off[r1] - c->h_row_off[r0];
        r0 = r1;
    }
    local.calls = local.rows * c->n;
    if (st) *st = local;
    return DNAF_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int dnaf_abi_version(void) { return DNAF_ABI_VERSION; }

const char* dnaf_last_error(const dnaf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dnaf_create(int device_ordinal, dnaf_ctx** out) {
    if (!out) return fail(nullptr, DNAF_E_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DNAF_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= count) return fail(nullptr, DNAF_E_ARG, "device ordinal out of range");
    dnaf_ctx* c = new dnaf_ctx();
    c->dev = device_ordinal;
    auto bail = [&](const char* what, cudaError_t err) {
        fail(nullptr, DNAF_E_CUDA, "%s: %s", what, cudaGetErrorString(err));
        delete c;
        return DNAF_E_CUDA;
    };
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    c->own_stream = true;
    for (auto& ev : c->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    // CRC tables: byte table and x^(8k) mod P for k = 0..kBlk
    std::vector<uint32_t> tab(256), xp(kBlk + 1);
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t v = i;
        for (int k = 0; k < 8; ++k) v = (v & 1u) ? (v >> 1) ^ kCrcPoly : (v >> 1);
        tab[i] = v;
    }
    xp[0] = 0x80000000u;  // x^0
    for (uint32_t k = 1; k <= kBlk; ++k) xp[k] = host_mulmod(xp[k - 1], 0x00800000u /* x^8 */);
    int rc = upload(c, c->d_crctab, tab.data(), tab.size());
    if (!rc) rc = upload(c, c->d_xpow8, xp.data(), xp.size());
    if (rc) {
        g_create_error = c->err;
        delete c;
        return rc;
    }
    *out = c;
    return DNAF_OK;
}

void dnaf_destroy(dnaf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->dev);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int dnaf_set_stream(dnaf_ctx* c, void* cuda_stream) {
    if (!c) return DNAF_E_ARG;
    if (c->own_stream && c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    c->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return DNAF_OK;
}

int dnaf_set_chunk_bytes(dnaf_ctx* c, uint64_t text_bytes) {
    if (!c) return DNAF_E_ARG;
    if (text_bytes < 4096) return fail(c, DNAF_E_ARG, "chunk must be at least 4096 bytes");
    c->chunk_bytes = std::min<uint64_t>(text_bytes, 8ull << 30);
    return DNAF_OK;
}

int dnaf_set_row_base(dnaf_ctx* c, uint64_t row_base) {
    if (!c) return DNAF_E_ARG;
    c->row_base = row_base;
    return DNAF_OK;
}

int dnaf_set_fused(dnaf_ctx* c, int enable) {
    if (!c) return DNAF_E_ARG;
    c->fused = enable ? 1 : 0;
    return DNAF_OK;
}

int dnaf_set_samples(dnaf_ctx* c, uint32_t n, const uint8_t* sex, const uint8_t* is_control) {
    if (!c) return DNAF_E_ARG;
    if (n && (!sex || !is_control)) return fail(c, DNAF_E_ARG, "sex / is_control are NULL");
    if (n > (1u << 30)) return fail(c, DNAF_E_ARG, "too many samples");
    CU(c, cudaSetDevice(c->dev));
    std::vector<uint32_t> xoff(n + 1);
    uint32_t males = 0, acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        xoff[i] = acc;
        const bool male = sex[i] == 1;
        males += male;
        acc += male ? 2u : 4u;
    }
    xoff[n] = acc;
    c->n = n;
    c->males = males;
    // body bytes incl. the '\n' that replaces the last '\t' ("\t".join(...) + "\n", pop_factory.py:508)
    c->body[kAuto] = n ? 4u * n : 1u;
    c->body[kX] = n ? acc : 1u;
    c->body[kY] = n ? 2u * n : 1u;
    c->body[kMT] = n ? 2u * n : 1u;
    int rc = upload(c, c->d_sex, sex, n);
    if (!rc) rc = upload(c, c->d_xoff, xoff.data(), xoff.size());
    if (rc) return rc;
    c->have_samples = true;
    c->layout_ok = false;
    return DNAF_OK;
}

int dnaf_set_snps(dnaf_ctx* c, uint64_t S, const uint8_t* cls, const uint8_t* k, const uint32_t* thr,
                  const uint8_t* prefix, const uint64_t* pre_off) {
    if (!c) return DNAF_E_ARG;
    if (S && (!cls || !k || !thr || !prefix || !pre_off)) return fail(c, DNAF_E_ARG, "NULL SNP array");
    CU(c, cudaSetDevice(c->dev));
    bool multi = false;
    for (uint64_t r = 0; r < S; ++r) {
        if (cls[r] > kMT) return fail(c, DNAF_E_ARG, "row %llu: bad chromosome class %u", (unsigned long long)r, cls[r]);
        if (k[r] < 1 || k[r] > kKmax)
            return fail(c, DNAF_E_ARG, "row %llu: %u alleles (supported: 1..%d)", (unsigned long long)r, k[r], kKmax);
        if (thr[r * 4 + k[r] - 1] != 0xFFFFFFFFu)
            return fail(c, DNAF_E_INPUT,
                        "row %llu: cumulative allele probabilities do not reach 1.0 "
                        "(the reference's pick_allele_index would return None)", (unsigned long long)r);
        if (pre_off[r + 1] < pre_off[r] || pre_off[r + 1] - pre_off[r] > (1u << 20))
            return fail(c, DNAF_E_ARG, "row %llu: bad prefix offsets", (unsigned long long)r);
        multi |= k[r] > 2;
    }
    c->S = S;
    c->any_multi = multi;
    c->h_cls.assign(cls, cls + S);
    c->h_plen.resize(S);
    for (uint64_t r = 0; r < S; ++r) c->h_plen[r] = (uint32_t)(pre_off[r + 1] - pre_off[r]);
    int rc = upload(c, c->d_cls, cls, S);
    if (!rc) rc = upload(c, c->d_k, k, S);
    if (!rc) rc = upload(c, c->d_thr, thr, S * 4);
    if (!rc) rc = upload(c, c->d_prefix, prefix, S ? pre_off[S] : 0);
    if (!rc) rc = upload(c, c->d_pre_off, pre_off, S + 1);
    if (rc) return rc;
    if (S == 0) {
        const uint64_t zero = 0;
        rc = upload(c, c->d_pre_off, &zero, 1);
        if (rc) return rc;
    }
    c->have_snps = true;
    c->layout_ok = false;
    return DNAF_OK;
}

int dnaf_set_overrides(dnaf_ctx* c, uint64_t P, const uint64_t* rows, const uint32_t* samples) {
    if (!c) return DNAF_E_ARG;
    if (P && (!rows || !samples)) return fail(c, DNAF_E_ARG, "NULL override array");
    for (uint64_t i = 1; i < P; ++i)
        if (rows[i] < rows[i - 1]) return fail(c, DNAF_E_ARG, "override pairs must be sorted by row");
    CU(c, cudaSetDevice(c->dev));
    c->h_orow.assign(rows, rows + P);
    int rc = upload(c, c->d_orow, rows, P);
    if (!rc) rc = upload(c, c->d_osamp, samples, P);
    if (rc) return rc;
    c->P = P;
    return DNAF_OK;
}

uint64_t dnaf_bgzf_bound(uint64_t text_bytes) {
    // every block carries <= kBlk bytes of text and at most 26 + 5 bytes of framing beyond them;
    // row-aligned cutting can leave blocks partly filled, so count blocks generously
    const uint64_t blocks = text_bytes / (kBlk / 2) + 2;
    return text_bytes + blocks * 64 + 1024;
}

int dnaf_bgzf_eof(uint8_t* out28) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43,
                                    0x02, 0x00, 0x1b, 0x00, 0x03, 0,    0, 0,    0,    0,    0,    0,    0, 0};
    if (!out28) return DNAF_E_ARG;
    memcpy(out28, eof, 28);
    return DNAF_OK;
}

int dnaf_plan(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t* text_bytes, uint64_t* bgzf_bound) {
    if (!c) return DNAF_E_ARG;
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    const uint64_t t = c->h_row_off[row_end] - c->h_row_off[row_begin];
    if (text_bytes) *text_bytes = t;
    if (bgzf_bound) *bgzf_bound = dnaf_bgzf_bound(t) + (row_end - row_begin) * 64;
    return DNAF_OK;
}

int dnaf_generate(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int rng_mode, int level,
                  uint8_t* out, uint64_t out_cap, dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (!out && out_cap) return fail(c, DNAF_E_ARG, "out is NULL");
    Sink s;
    s.buf = out;
    s.cap = out_cap;
    return generate_impl(c, row_begin, row_end, seed, rng_mode, level, s, stats);
}

int dnaf_generate_stream(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int rng_mode, int level,
                         dnaf_sink_fn sink, void* user, dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (!sink) return fail(c, DNAF_E_ARG, "sink is NULL");
    Sink s;
    s.fn = sink;
    s.user = user;
    return generate_impl(c, row_begin, row_end, seed, rng_mode, level, s, stats);
}

int dnaf_generate_device(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int rng_mode, int level,
                         dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    Sink s;
    s.device_only = true;
    return generate_impl(c, row_begin, row_end, seed, rng_mode, level, s, stats);
}

int dnaf_genotypes(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, uint8_t* out, uint64_t cap) {
    if (!c) return DNAF_E_ARG;
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    const uint64_t need = (row_end - row_begin) * c->n * 2ull;
    if (need > cap) return fail(c, DNAF_E_SPACE, "genotype buffer too small: need %llu bytes", (unsigned long long)need);
    if (need == 0) return DNAF_OK;
    if (!out) return fail(c, DNAF_E_ARG, "out is NULL");
    CU(c, cudaSetDevice(c->dev));
    const uint64_t rows_per = std::max<uint64_t>(1, (64ull << 20) / (2ull * c->n));
    for (uint64_t r0 = row_begin; r0 < row_end; r0 += rows_per) {
        const uint64_t r1 = std::min(row_end, r0 + rows_per);
        rc = run_sample(c, r0, r1, seed, nullptr);
        if (rc) return rc;
        const uint64_t cells = (r1 - r0) * c->n;
        CU(c, c->d_geno.reserve(cells * 2));
        k_export_genotypes<<<(uint32_t)((cells + 255) / 256), 256, 0, c->stream>>>(
            sample_view(c), snp_view(c), r0, (uint32_t)(r1 - r0), c->d_plane0.as<uint32_t>(),
            c->any_multi ? c->d_plane1.as<uint32_t>() : nullptr, c->d_geno.as<uint8_t>());
        CU(c, cudaGetLastError());
        CU(c, cudaMemcpyAsync(out + (r0 - row_begin) * c->n * 2ull, c->d_geno.p, cells * 2, cudaMemcpyDeviceToHost,
                              c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return DNAF_OK;
}

int dnaf_text(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, uint8_t* out, uint64_t cap,
              uint64_t* n_bytes) {
    if (!c) return DNAF_E_ARG;
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    const uint64_t need = c->h_row_off[row_end] - c->h_row_off[row_begin];
    if (n_bytes) *n_bytes = need;
    if (need > cap) return fail(c, DNAF_E_SPACE, "text buffer too small: need %llu bytes", (unsigned long long)need);
    if (need == 0) return DNAF_OK;
    if (!out) return fail(c, DNAF_E_ARG, "out is NULL");
    CU(c, cudaSetDevice(c->dev));
    uint64_t r0 = row_begin, done = 0;
    while (r0 < row_end) {
        const uint64_t r1 = next_chunk_end(c, r0, row_end, c->chunk_bytes);
        rc = run_sample(c, r0, r1, seed, nullptr);
        if (!rc) rc = run_format(c, r0, r1, nullptr);
        if (rc) return rc;
        const uint64_t bytes = c->h_row_off[r1] - c->h_row_off[r0];
        CU(c, cudaMemcpyAsync(out + done, c->d_text.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        done += bytes;
        r0 = r1;
    }
    return DNAF_OK;
}

int dnaf_bgzf_compress(dnaf_ctx* c, const uint8_t* text, uint64_t n, int level, uint8_t* out, uint64_t cap,
                       dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (level < 1 || level > 9) return fail(c, DNAF_E_ARG, "level must be 1..9");
    if (n && !text) return fail(c, DNAF_E_ARG, "text is NULL");
    CU(c, cudaSetDevice(c->dev));
    dnaf_stats local;
    memset(&local, 0, sizeof local);
    Sink s;
    s.buf = out;
    s.cap = cap;
    uint64_t done = 0;
    while (done < n) {
        const uint64_t piece = std::min<uint64_t>(n - done, (c->chunk_bytes / kBlk + 1) * (uint64_t)kBlk);
        CU(c, c->d_text.reserve(piece + 64));
        CU(c, cudaMemcpyAsync(c->d_text.p, text + done, piece, cudaMemcpyHostToDevice, c->stream));
        c->plan.clear();
        for (uint64_t o = 0; o < piece; o += kBlk)
            c->plan.push_back({o, (uint32_t)std::min<uint64_t>(kBlk, piece - o), 0});
        int rc = encode_plan(c, s, &local, c->ev[3], c->ev[4]);
        if (rc) return rc;
        local.text_bytes += piece;
        done += piece;
    }
    local.ms_total = local.ms_deflate;
    if (stats) *stats = local;
    return DNAF_OK;
}

}  // extern "C"
