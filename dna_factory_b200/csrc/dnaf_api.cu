// C ABI (include/dnaf_b200.h) and host orchestration of the B200 hot path.
//
// Replaces, for the population-generation path of ochrzan/dna-factory:
//   PopulationFactory.write_vcf_snps   pop_factory.py:417-469  (worker pool + ordered writer)
//   PopulationFactory.queue_vcf_snps   pop_factory.py:471-513  (row loop)
//   Bio.bgzf.BgzfWriter.write          call site pop_factory.py:449
// Work is cut into passes of at most `chunk_bytes` of uncompressed text; each pass is
//   sample -> (overrides) -> format -> BGZF encode -> scan -> compact -> D2H -> sink      (generic path)
// or the fused kernel (k_fused.cuh) followed by scan -> compact -> D2H -> sink.
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dnaf_b200.h"
#include "dnaf_device.cuh"
#include "fused_host.h"
#include "k_deflate.cuh"
#include "k_fused.cuh"
#include "k_fused_text.cuh"
#include "k_sample_format.cuh"
#include "k_select.cuh"
#include "snps_json.h"
#include <map>
#include <unordered_map>

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
        const size_t want = std::max<size_t>(bytes + bytes / 2, 64 << 10);   // geometric: sizes settle after a few passes
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
        const size_t want = std::max<size_t>(bytes + bytes / 2, 64 << 10);
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
    uint64_t chunk_bytes = 1024ull << 20;
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
    std::vector<uint8_t> h_cls, h_k;
    std::vector<uint32_t> h_thr0;
    std::vector<uint32_t> h_plen;
    std::vector<uint64_t> h_row_off;  // valid when layout_ok
    bool layout_ok = false;

    // overrides
    uint64_t P = 0;
    DevBuf d_orow, d_osamp;
    std::vector<uint64_t> h_orow;
    std::vector<uint32_t> h_osamp;

    // constant tables
    DevBuf d_crctab, d_xpow8;

    // scratch
    DevBuf d_plane0, d_plane1, d_text, d_blocks, d_geno;
    struct SlotBuf {                       // block slots of a pass; two sets, so that the compaction of pass i (own
        DevBuf slots, sizes, crcs;         // stream) overlaps the kernels of pass i+1
        cudaEvent_t ev_free = nullptr;     // compaction that read this set has finished
    } sbuf[2];
    int sb = 0;
    cudaStream_t comp = nullptr;           // k_size_partials + k_gather run here
    PinnedBuf h_blocks;
    std::vector<BlockDesc> plan;

    cudaStream_t side = nullptr;           // k_fused_text runs here, concurrently with k_auto
    cudaStream_t side2 = nullptr;          // k_x runs here
    cudaEvent_t ev_join2 = nullptr;
    cudaStream_t copy = nullptr;           // D2H of pass i overlaps the kernels of pass i+1
    struct OutBuf {                        // what must outlive a pass while the next one runs
        DevBuf d_out, d_totals;
        PinnedBuf h_totals, h_out, h_stage;  // h_stage: descriptor uploads of the pass (truly asynchronous H2D)
        size_t stage_used = 0;
        cudaEvent_t ev[6] = {};
        cudaEvent_t ev_auto[2] = {};         // around the k_auto launch
        uint64_t auto_text = 0;              // text bytes of the pass's k_auto blocks (0: no k_auto launch)
        cudaEvent_t ev_copied = nullptr;
        uint32_t nb = 0;
        uint64_t rows = 0, text = 0;
        bool gen = false, fused = false, generic_blocks = false;
        int copy_mode = 0;                   // 0 nothing in flight, 1 DMA into the caller's pinned buffer, 2 via h_out
        uint64_t copy_bytes = 0;
        const uint8_t* copy_dst = nullptr;   // mode 1: where in the caller's buffer the pass lands
    } ob[3];
    // optional record of every BGZF block handed to a host sink by dnaf_generate* (dnaf_block_log)
    bool log_blocks = false;
    std::vector<uint32_t> log_csize, log_usize;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool attr_done = false;

    // fused path (k_fused.cuh): per-bucket static codes, CRC helper tables, per-segment template CRCs
    bool fused_ok = false;
    DevBuf d_ftables, d_etab, d_fdesc, d_gslot, d_grow, d_goff, d_olocal, d_osub;
    std::vector<uint16_t> h_bucket;        // per row
    std::vector<uint32_t> h_seg_crc;       // L(template body) per autosome segment index
    std::vector<uint32_t> h_seg_cell0;     // first cell of every segment (+ end sentinel)
    std::vector<FusedDesc> fplan;
    std::vector<uint32_t> gslot, grow, olocal, osub;
    std::vector<uint64_t> goff;
    uint64_t gen_text_bytes = 0;
    uint32_t pass_blocks = 0;
    uint32_t slot_stride = kSlot;          // per pass: the longest block's text + room for framing, rounded to 256
    uint64_t pass_text = 0;
    uint32_t fused_threads = 256;
    int cur_ob = 0;
    std::map<std::pair<uint64_t, uint64_t>, FusedTable> table_cache;
    bool etab_ok = false;
    std::vector<double> bucket_p;          // minor-allele probability per bucket
    std::unordered_map<uint32_t, int> bucket_of;
    int bucket_shift = 0;
    std::vector<uint64_t> ph;              // prefix byte model
    uint64_t ph_hash = 0;
    uint64_t samples_epoch = 0, seg_epoch = ~0ull;
    std::vector<uint64_t> tables_sig;      // what d_ftables currently holds
    std::vector<uint8_t> h_sex;
    DevBuf d_crc4, d_xspan, d_tdesc, d_xspans, d_xdesc;
    // k_auto (k_auto.cuh): code tables + byte LUTs per (bucket, starts-row), CRC move tables, per-row prefix CRCs
    DevBuf d_atables, d_etab2, d_mtab, d_mtail, d_mpre, d_xinit, d_pre_crc;
    DevBuf d_xtables, d_mspan, d_mpre_x;   // k_x (k_x.cuh)
    DevBuf d_bucket, d_ovr_first, d_seginfo;   // implicit block descriptors of all-autosome passes (k_auto.cuh)
    std::vector<uint32_t> h_other;          // [S+1]: rows before r that do NOT take k_auto
    bool implicit_pass = false;
    std::map<std::pair<uint64_t, uint64_t>, XTable> xtable_cache;
    std::vector<uint32_t> h_mspan, h_mpre_x;
    std::map<std::pair<uint64_t, uint64_t>, AutoTable> atable_cache;
    // k_lz (k_lz.cuh): code tables per (bucket, starts-row) of the LZ tier in use (-z 4..9)
    DevBuf d_ltables;
    std::map<std::pair<std::pair<uint64_t, uint64_t>, int>, LzTable> ltable_cache;
    std::vector<uint64_t> ltables_sig;     // what d_ltables currently holds
    bool lz_ok = false;
    bool lz_attr_done = false;
    std::vector<uint32_t> h_mtail, h_mpre;
    DevBuf d_pfx_state;
    PinnedBuf h_present;                   // byte values seen in the row prefixes (written by k_prefix_crc)
    uint32_t present_sticky[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // prefix byte values seen so far (a superset keeps table keys stable)
    std::vector<uint8_t> need_sticky;      // (bucket, variant) tables ever needed: the uploaded set only grows
    std::vector<uint32_t> bk_key; std::vector<int> bk_val;   // direct-mapped cache in front of bucket_of
    bool seg_tabs_dirty = false;
    std::vector<uint8_t> h_pfx_tab;        // per row: prefix ends with '\t' (k_auto's first match may reach into it)
    std::vector<XSpan> h_xspans;
    std::vector<uint32_t> h_seg_crc_x;     // L(template body) per X segment
    std::vector<uint32_t> h_xoff;
    std::vector<FusedDesc> xplan;
    std::vector<TextDesc> tplan;
    std::vector<uint32_t> seg_byte0[4];    // k_fused_text segments per chromosome class (+ end sentinel)
    uint32_t text_threads = 64;
    bool text_attr_done = false;
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

static bool g_trace = getenv("DNAF_TRACE") != nullptr;
static std::chrono::steady_clock::time_point g_t0;
static void trace(const char* what, int pass) {
    if (!g_trace) return;
    fprintf(stderr, "[dnaf] %8.3f ms  %s %d\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count(), what, pass);
}

template <class T>
int upload(dnaf_ctx* c, DevBuf& b, const T* src, size_t count, bool sync = true) {
    CU(c, b.reserve(std::max<size_t>(count, 1) * sizeof(T) + 64));
    if (count) CU(c, cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    if (sync) CU(c, cudaStreamSynchronize(c->stream));
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

void build_segments(dnaf_ctx* c);
int ensure_tables(dnaf_ctx* c);
int ensure_implicit(dnaf_ctx* c);

// Text offset of every row (prefix + class body), host and device copies.
int ensure_layout(dnaf_ctx* c) {
    if (!c->have_samples || !c->have_snps) return fail(c, DNAF_E_ARG, "set_samples and set_snps must be called first");
    if (c->layout_ok) return DNAF_OK;
    if (g_trace) { g_t0 = std::chrono::steady_clock::now(); trace("ensure_layout begins", 0); }
    c->h_row_off.resize(c->S + 1);
    uint64_t acc = 0;
    for (uint64_t r = 0; r < c->S; ++r) {
        c->h_row_off[r] = acc;
        acc += (uint64_t)c->h_plen[r] + c->body[c->h_cls[r]];
    }
    c->h_row_off[c->S] = acc;
    trace("row offsets summed", 0);
    int rc = upload(c, c->d_row_off, c->h_row_off.data(), c->S + 1);
    if (rc) return rc;
    trace("row offsets uploaded", 0);
    build_segments(c);
    if (c->seg_tabs_dirty) {
        rc = upload(c, c->d_mtail, c->h_mtail.data(), c->h_mtail.size());
        if (!rc) rc = upload(c, c->d_mpre, c->h_mpre.data(), c->h_mpre.size());
        if (!rc) rc = upload(c, c->d_mspan, c->h_mspan.data(), c->h_mspan.size());
        if (!rc) rc = upload(c, c->d_mpre_x, c->h_mpre_x.data(), c->h_mpre_x.size());
        if (rc) return rc;
        c->seg_tabs_dirty = false;
    }
    rc = ensure_tables(c);
    if (rc) return rc;
    trace("tables ensured", 0);
    rc = ensure_implicit(c);
    if (rc) return rc;
    trace("implicit descriptors ready", 0);
    c->layout_ok = true;
    return DNAF_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused-path set-up: MAF buckets -> static Huffman tables; CRC helper tables; template CRCs per segment.
#ifndef DNAF_MIN_FUSED
#define DNAF_MIN_FUSED 4096
#endif
constexpr uint32_t kFusedMinRowBytes = DNAF_MIN_FUSED;  // rows of at least 1024 diploid samples get a block of their own (fused kernels);
                                                         // shorter ones are packed several to a block by the generic path

uint32_t raw_crc(const uint8_t* p, size_t n, const uint32_t* tab) {
    uint32_t c = 0;
    for (size_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c;
}

constexpr int kVariants = 12;  // tables per MAF bucket: [0,1] k_auto with / without prefix, [2..9] k_fused_text (class x {prefix, no prefix}), [10,11] k_x

// Called from set_snps: bucket every row by its first threshold, remember which prefix bytes occur.
int prepare_buckets(dnaf_ctx* c, const uint8_t* kk, const uint32_t* thr) {
    c->h_bucket.assign(c->S, 0);
    if (c->S == 0) return DNAF_OK;
    // bucket key: the first threshold (minor-allele probability = 1 - (T+1)/2^32), coarsened (shift) if a
    // population ever shows more than 512 distinct values.  The key -> bucket map persists across set_snps
    // calls, so bucket ids -- and with them the uploaded tables -- stay put when successive SNP batches arrive.
    for (;;) {
        bool ok = true;
        uint32_t last_key = 0;
        int last_bucket = -1;
        c->bk_key.assign(1024, 0);
        c->bk_val.assign(1024, -1);
        for (uint64_t r = 0; r < c->S; ++r) {
            const uint32_t key = (kk[r] >= 2 ? thr[r * 4] : 0xFFFFFFFFu) >> c->bucket_shift;
            if (last_bucket < 0 || key != last_key) {
                const uint32_t h = (key * 2654435761u) >> 22;
                if (c->bk_val[h] >= 0 && c->bk_key[h] == key) {
                    last_key = key;
                    last_bucket = c->bk_val[h];
                    c->h_bucket[r] = (uint16_t)last_bucket;
                    continue;
                }
                auto it = c->bucket_of.find(key);
                if (it == c->bucket_of.end()) {
                    if (c->bucket_of.size() >= 512) { ok = false; break; }
                    const uint64_t lo = (uint64_t)key << c->bucket_shift;
                    const uint64_t hi = std::min<uint64_t>(0xFFFFFFFFull, lo + ((1ull << c->bucket_shift) - 1));
                    const double t_mid = 0.5 * ((double)lo + (double)hi);
                    c->bucket_p.push_back(std::min(1.0, std::max(0.0, 1.0 - (t_mid + 1.0) / 4294967296.0)));
                    it = c->bucket_of.emplace(key, (int)c->bucket_of.size()).first;
                }
                last_key = key;
                last_bucket = it->second;
                c->bk_key[h] = key;
                c->bk_val[h] = last_bucket;
            }
            c->h_bucket[r] = (uint16_t)last_bucket;
        }
        if (ok) break;
        c->bucket_shift += 2;  // too many distinct thresholds: merge neighbours and start over
        c->bucket_of.clear();
        c->bucket_p.clear();
        c->tables_sig.clear();
        c->need_sticky.clear();
    }
    return DNAF_OK;
}

// prefix byte model (x16 fixed point per row): which bytes occur (from k_prefix_crc), weighted by kind --
// deliberately not the exact counts, so that tables can be cached across set_snps calls with similar prefixes
void prefix_model(dnaf_ctx* c, const uint32_t* present) {
    c->ph.assign(256, 0);
    c->ph_hash = 1469598103934665603ull;
    for (int b = 0; b < 256; ++b) {
        if ((present[b >> 5] >> (b & 31)) & 1u) c->ph[b] = b == '\t' ? 144 : ((b >= '0' && b <= '9') ? 24 : 16);
        c->ph_hash = (c->ph_hash ^ c->ph[b]) * 1099511628211ull;
    }
}

// Called from ensure_layout (samples and SNPs known): static Huffman tables for every (bucket, variant) in use.
int ensure_tables(dnaf_ctx* c) {
    c->fused_ok = false;
    if (c->S == 0 || c->n == 0) return DNAF_OK;
    const int nb = (int)c->bucket_p.size();
    std::vector<uint8_t> need((size_t)nb * kVariants, 0);
    for (uint64_t r = 0; r < c->S; ++r) {
        const int b = c->h_bucket[r];
        if (c->h_cls[r] == kAuto && c->h_k[r] <= 2) {
            need[b * kVariants + 0] = need[b * kVariants + 1] = 1;
        } else if (c->h_cls[r] == kX && c->h_k[r] <= 2) {
            need[b * kVariants + 10] = need[b * kVariants + 11] = 1;
        } else {
            need[b * kVariants + 2 + 2 * c->h_cls[r]] = need[b * kVariants + 3 + 2 * c->h_cls[r]] = 1;
        }
    }
    // Tables stay once they have been needed, and a bucket gets every variant any bucket has needed: successive SNP
    // batches differ in which buckets their (few) X / Y rows hit, and the uploaded set must settle quickly.
    if (c->need_sticky.size() < need.size()) c->need_sticky.resize(need.size(), 0);
    {
        uint8_t var_seen[kVariants] = {0};
        std::vector<uint8_t> bucket_seen(nb, 0);
        for (int b = 0; b < nb; ++b)
            for (int v = 0; v < kVariants; ++v)
                if (need[b * kVariants + v] | c->need_sticky[b * kVariants + v]) var_seen[v] = bucket_seen[b] = 1;
        for (int b = 0; b < nb; ++b)
            for (int v = 0; v < kVariants; ++v)
                need[b * kVariants + v] = c->need_sticky[b * kVariants + v] = bucket_seen[b] && var_seen[v];
    }
    std::vector<uint64_t> sig;
    sig.reserve(need.size() + 2);
    sig.push_back(c->ph_hash);
    sig.push_back(c->samples_epoch);
    for (int b = 0; b < nb; ++b) {
        uint64_t pbits;
        memcpy(&pbits, &c->bucket_p[b], 8);
        for (int v = 0; v < kVariants; ++v) sig.push_back(need[b * kVariants + v] ? pbits : 0);
    }
    if (sig != c->tables_sig) {
        std::vector<FusedTable> tabs((size_t)nb * kVariants);
        memset(tabs.data(), 0, tabs.size() * sizeof(FusedTable));
        std::vector<AutoTable> atabs((size_t)nb * 2);
        memset(atabs.data(), 0, atabs.size() * sizeof(AutoTable));
        std::vector<XTable> xtabs;
        bool any_x = false;
        for (int b = 0; b < nb; ++b) any_x |= need[b * kVariants + 10] || need[b * kVariants + 11];
        if (any_x) {
            xtabs.resize((size_t)nb * 2);
            memset(xtabs.data(), 0, xtabs.size() * sizeof(XTable));
        }
        const int per_block = (int)std::max<size_t>(1, c->h_seg_cell0.size() > 1
                                                           ? (c->h_seg_cell0[1] - c->h_seg_cell0[0] + 63) / 64 : 1);
        {   // build the tables the caches do not hold yet, on all host threads (about 5 ms each, several hundred
            // on a first call); the loop below then finds every table cached
            struct Job { int v; double p; std::pair<uint64_t, uint64_t> key; };
            std::vector<Job> jobs;
            std::map<std::pair<int, std::pair<uint64_t, uint64_t>>, int> seen;
            for (int b = 0; b < nb; ++b) {
                const double p = c->bucket_p[b];
                uint64_t pbits;
                memcpy(&pbits, &p, 8);
                for (int v = 0; v < kVariants; ++v) {
                    if (!need[b * kVariants + v]) continue;
                    const bool with_prefix = (v & 1) == 0;
                    const uint64_t base = (with_prefix ? c->ph_hash : 0x5bd1e995ull) * 31;
                    std::pair<uint64_t, uint64_t> key;
                    bool cached;
                    int family;
                    if (v < 2) {
                        key = {pbits, base + (uint64_t)per_block};
                        cached = c->atable_cache.count(key) != 0;
                        family = 0;
                    } else if (v >= 10) {
                        key = {pbits, base + (uint64_t)per_block + c->samples_epoch * 0x9E3779B97F4A7C15ull};
                        cached = c->xtable_cache.count(key) != 0;
                        family = 1;
                    } else {
                        const int cls = (v - 2) / 2;
                        key = {pbits, base + (uint64_t)(cls + 1) * 1000003ull + c->samples_epoch * 0x9E3779B97F4A7C15ull};
                        cached = c->table_cache.count(key) != 0;
                        family = 2 + cls;
                    }
                    if (!cached && seen.emplace(std::make_pair(family, key), 1).second) jobs.push_back({v, p, key});
                }
            }
            if (!jobs.empty()) {
                std::vector<AutoTable> ra(jobs.size());
                std::vector<XTable> rx;
                std::vector<FusedTable> rf(jobs.size());
                bool need_x = false;
                for (const Job& j : jobs) need_x |= j.v >= 10;
                if (need_x) rx.resize(jobs.size());
                std::atomic<size_t> next{0};
                auto work = [&]() {
                    for (size_t i = next++; i < jobs.size(); i = next++) {
                        const Job& j = jobs[i];
                        const bool with_prefix = (j.v & 1) == 0;
                        const uint64_t* hist = with_prefix ? c->ph.data() : nullptr;
                        if (j.v < 2) ra[i] = hosttab::make_auto_table(j.p, hist, per_block, with_prefix);
                        else if (j.v >= 10) rx[i] = hosttab::make_x_table(j.p, c->h_xspans, per_block, hist);
                        else rf[i] = hosttab::make_text_table((j.v - 2) / 2, j.p, c->h_sex.data(), c->n, hist);
                    }
                };
                const unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 16u, (unsigned)jobs.size()}));
                std::vector<std::thread> pool;
                for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work);
                work();
                for (auto& t : pool) t.join();
                for (size_t i = 0; i < jobs.size(); ++i) {
                    const Job& j = jobs[i];
                    const uint32_t hb = j.v < 2 ? ra[i].hdr_bits : (j.v >= 10 ? rx[i].hdr_bits : rf[i].hdr_bits);
                    if (hb == 0xFFFFFFFFu) return DNAF_OK;  // header too long: stay on the generic path
                    if (j.v < 2) c->atable_cache.emplace(j.key, ra[i]);
                    else if (j.v >= 10) c->xtable_cache.emplace(j.key, rx[i]);
                    else c->table_cache.emplace(j.key, rf[i]);
                }
            }
        }
        for (int b = 0; b < nb; ++b) {
            const double p = c->bucket_p[b];
            uint64_t pbits;
            memcpy(&pbits, &p, 8);
            for (int v = 0; v < kVariants; ++v) {
                if (!need[b * kVariants + v]) continue;
                const bool with_prefix = (v & 1) == 0;
                if (v < 2) {   // autosome rows: k_auto's tables
                    auto key = std::make_pair(pbits, (with_prefix ? c->ph_hash : 0x5bd1e995ull) * 31 + (uint64_t)per_block);
                    auto it = c->atable_cache.find(key);
                    if (it == c->atable_cache.end()) {
                        AutoTable t = hosttab::make_auto_table(p, with_prefix ? c->ph.data() : nullptr, per_block, with_prefix);
                        if (t.hdr_bits == 0xFFFFFFFFu) return DNAF_OK;  // header too long: stay on the generic path
                        it = c->atable_cache.emplace(key, t).first;
                    }
                    atabs[(size_t)b * 2 + v] = it->second;
                    continue;
                }
                if (v >= 10) {   // X rows: k_x's tables
                    auto key = std::make_pair(pbits, (with_prefix ? c->ph_hash : 0x5bd1e995ull) * 31 + (uint64_t)per_block +
                                                         c->samples_epoch * 0x9E3779B97F4A7C15ull);
                    auto it = c->xtable_cache.find(key);
                    if (it == c->xtable_cache.end()) {
                        XTable t = hosttab::make_x_table(p, c->h_xspans, per_block, with_prefix ? c->ph.data() : nullptr);
                        if (t.hdr_bits == 0xFFFFFFFFu) return DNAF_OK;
                        it = c->xtable_cache.emplace(key, t).first;
                    }
                    xtabs[(size_t)b * 2 + (v - 10)] = it->second;
                    continue;
                }
                const int cls = v < 2 ? -1 : (v >= 10 ? 100 : (v - 2) / 2);   // -1: autosome cells, 100: X cells
                const uint64_t vkey = (with_prefix ? c->ph_hash : 0x5bd1e995ull) * 31 + (uint64_t)(cls + 1) * 1000003ull +
                                      (cls >= 0 ? c->samples_epoch * 0x9E3779B97F4A7C15ull : 0);
                auto key = std::make_pair(pbits, vkey);
                auto it = c->table_cache.find(key);
                if (it == c->table_cache.end()) {
                    const uint64_t* hist = with_prefix ? c->ph.data() : nullptr;
                        FusedTable t = hosttab::make_text_table(cls, p, c->h_sex.data(), c->n, hist);
                    if (t.hdr_bits == 0xFFFFFFFFu) return DNAF_OK;  // header too long: stay on the generic path
                    it = c->table_cache.emplace(key, t).first;
                }
                tabs[(size_t)b * kVariants + v] = it->second;
            }
        }
        int rc = upload(c, c->d_ftables, tabs.data(), tabs.size());
        if (!rc) rc = upload(c, c->d_atables, atabs.data(), atabs.size());
        if (!rc && any_x) rc = upload(c, c->d_xtables, xtabs.data(), xtabs.size());
        if (rc) return rc;
        c->tables_sig = sig;
    }
    if (!c->etab_ok) {
        // E tables: contribution of mask byte b at byte k of word w to the span's linear CRC (span end aligned)
        std::vector<uint32_t> tab(256), xp(257);
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t v = i;
            for (int k = 0; k < 8; ++k) v = (v & 1u) ? (v >> 1) ^ kCrcPoly : (v >> 1);
            tab[i] = v;
        }
        xp[0] = 0x80000000u;
        for (int k = 1; k <= 256; ++k) xp[k] = hosttab::mulmod(xp[k - 1], 0x00800000u);
        std::vector<uint32_t> etab(16 * 256, 0);
        for (int w = 0; w < 4; ++w)
            for (int k = 0; k < 4; ++k)
                for (int b = 0; b < 256; ++b) {
                    uint32_t v = 0;
                    for (int i = 0; i < 8; ++i)
                        if ((b >> i) & 1) {
                            const int j = 32 * w + 8 * k + i;             // allele slot, byte 2j of the span
                            v ^= hosttab::mulmod(xp[255 - 2 * j], tab[1]);
                        }
                    etab[(4 * w + k) * 256 + b] = v;
                }
        int rc = upload(c, c->d_etab, etab.data(), etab.size());
        if (rc) return rc;
        {   // k_auto: E table measured to one byte before the span's cell end; moves by whole spans; init terms
            std::vector<uint32_t> etab2(16 * 256, 0);
            for (int w = 0; w < 4; ++w)
                for (int k = 0; k < 4; ++k)
                    for (int b = 0; b < 256; ++b) {
                        uint32_t v = 0;
                        for (int i = 0; i < 8; ++i)
                            if ((b >> i) & 1) {
                                const int j = 32 * w + 8 * k + i;
                                v ^= hosttab::mulmod(xp[254 - 2 * j], tab[1]);
                            }
                        etab2[(4 * w + k) * 256 + b] = v;
                    }
            rc = upload(c, c->d_etab2, etab2.data(), etab2.size());
            if (rc) return rc;
            std::vector<uint32_t> mtab((size_t)254 * 1024);
            uint32_t xj = 0x80000000u;   // x^(8*256*j)
            for (int j = 0; j < 254; ++j) {
                hosttab::fill_mul_table(xj, &mtab[(size_t)j * 1024]);
                xj = hosttab::mulmod(xj, xp[256]);
            }
            rc = upload(c, c->d_mtab, mtab.data(), mtab.size());
            if (rc) return rc;
            std::vector<uint32_t> xinit(kBlk + 1);
            xinit[0] = 0xFFFFFFFFu;
            for (uint32_t i = 1; i <= kBlk; ++i) xinit[i] = tab[xinit[i - 1] & 0xFFu] ^ (xinit[i - 1] >> 8);
            rc = upload(c, c->d_xinit, xinit.data(), xinit.size());
            if (rc) return rc;
        }
        // slicing-by-4 tables for k_fused_text
        std::vector<uint32_t> c4(1024);
        for (int i = 0; i < 256; ++i) c4[i] = tab[i];
        for (int t = 1; t < 4; ++t)
            for (int i = 0; i < 256; ++i) c4[256 * t + i] = (c4[256 * (t - 1) + i] >> 8) ^ tab[c4[256 * (t - 1) + i] & 0xFFu];
        rc = upload(c, c->d_crc4, c4.data(), c4.size());
        if (rc) return rc;
        c->etab_ok = true;
    }
    c->fused_ok = true;
    return DNAF_OK;
}

// The LZ tiers' code tables (k_lz.cuh) for the level of this call: one per (bucket with autosome rows, starts-row).
// Built lazily at the first dnaf_generate* call that asks for -z >= 4, cached per (bucket, prefix model, level).
int ensure_lz_tables(dnaf_ctx* c, int level) {
    c->lz_ok = false;
    if (!c->fused_ok || level < 4 || c->h_seg_crc.empty()) return DNAF_OK;
    const int nb = (int)c->bucket_p.size();
    const int per_block = (int)std::max<size_t>(1, c->h_seg_cell0.size() > 1
                                                       ? (c->h_seg_cell0[1] - c->h_seg_cell0[0] + 63) / 64 : 1);
    std::vector<uint64_t> sig;
    sig.push_back(c->ph_hash);
    sig.push_back((uint64_t)level);
    sig.push_back((uint64_t)per_block);
    for (int b = 0; b < nb; ++b) {
        uint64_t pbits;
        memcpy(&pbits, &c->bucket_p[b], 8);
        sig.push_back(c->need_sticky[(size_t)b * kVariants] ? pbits : 0);
    }
    if (sig == c->ltables_sig) {
        c->lz_ok = true;
        return DNAF_OK;
    }
    struct Job { int b, v; std::pair<std::pair<uint64_t, uint64_t>, int> key; };
    std::vector<Job> jobs;
    std::map<std::pair<std::pair<uint64_t, uint64_t>, int>, int> seen;
    auto key_of = [&](int b, int v) {
        uint64_t pbits;
        memcpy(&pbits, &c->bucket_p[b], 8);
        return std::make_pair(std::make_pair(pbits, (v == 0 ? c->ph_hash : 0x5bd1e995ull) * 31 + (uint64_t)per_block), level);
    };
    for (int b = 0; b < nb; ++b) {
        if (!c->need_sticky[(size_t)b * kVariants]) continue;
        for (int v = 0; v < 2; ++v) {
            auto key = key_of(b, v);
            if (!c->ltable_cache.count(key) && seen.emplace(key, 1).second) jobs.push_back({b, v, key});
        }
    }
    if (!jobs.empty()) {
        std::vector<LzTable> res(jobs.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t i = next++; i < jobs.size(); i = next++)
                res[i] = hosttab::make_lz_table(c->bucket_p[jobs[i].b], jobs[i].v == 0 ? c->ph.data() : nullptr, per_block,
                                                jobs[i].v == 0, level);
        };
        const unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 16u, (unsigned)jobs.size()}));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        for (size_t i = 0; i < jobs.size(); ++i) {
            if (res[i].hdr_bits == 0xFFFFFFFFu) return DNAF_OK;   // header too long: the call stays on k_auto
            c->ltable_cache.emplace(jobs[i].key, res[i]);
        }
    }
    std::vector<LzTable> tabs((size_t)nb * 2);
    memset(tabs.data(), 0, tabs.size() * sizeof(LzTable));
    for (int b = 0; b < nb; ++b) {
        if (!c->need_sticky[(size_t)b * kVariants]) continue;
        for (int v = 0; v < 2; ++v) tabs[(size_t)b * 2 + v] = c->ltable_cache.at(key_of(b, v));
    }
    const int rc = upload(c, c->d_ltables, tabs.data(), tabs.size());
    if (rc) return rc;
    c->ltables_sig = sig;
    c->lz_ok = true;
    return DNAF_OK;
}

// Segments of an autosome row (balanced, at most 254 spans of 64 samples each) and the linear CRC of their
// all-reference template bodies.
void build_segments(dnaf_ctx* c) {
    if (c->seg_epoch == c->samples_epoch) return;  // depends on the sample set only
    c->seg_epoch = c->samples_epoch;
    c->h_seg_cell0.clear();
    c->h_seg_crc.clear();
    c->fused_threads = 64;
    for (auto& v : c->seg_byte0) v.clear();
    if (c->n == 0) return;
    std::vector<uint32_t> tab(256);
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t v = i;
        for (int k = 0; k < 8; ++k) v = (v & 1u) ? (v >> 1) ^ kCrcPoly : (v >> 1);
        tab[i] = v;
    }
    std::vector<uint8_t> body((size_t)4 * c->n);
    for (uint32_t i = 0; i < c->n; ++i) memcpy(&body[4ull * i], "0/0\t", 4);
    body.back() = '\n';
    const uint32_t spans = (c->n + 63u) / 64u;
    const uint32_t nseg = (spans + 253u) / 254u;
    const uint32_t per = (spans + nseg - 1u) / nseg;  // spans per segment: <= 254, so prefix + body <= kBlk
    for (uint32_t sg = 0; sg < nseg; ++sg) {
        const uint32_t cell = std::min(c->n, sg * per * 64u);
        const uint32_t cnt = std::min(c->n, (sg + 1) * per * 64u) - cell;
        if (!cnt) break;
        c->h_seg_cell0.push_back(cell);
        // k_auto's blocks: a segment starts with the separator that ended the previous one and stops before its own
        // last separator, unless it ends the row
        const uint64_t b0 = sg ? 4ull * cell - 1 : 0, b1 = (cell + cnt == c->n) ? 4ull * c->n : 4ull * (cell + cnt) - 1;
        c->h_seg_crc.push_back(raw_crc(&body[b0], b1 - b0, tab.data()));
    }
    c->h_seg_cell0.push_back(c->n);
    c->fused_threads = std::max(64u, (per + 31u) / 32u * 32u);
    {   // CRC move tables that depend on the sample count: the short last span, and prefix -> end of segment 0
        std::vector<uint32_t> xp(4ull * c->n + 2);
        xp[0] = 0x80000000u;
        for (size_t k = 1; k < xp.size(); ++k) xp[k] = (xp[k - 1] >> 8) ^ tab[xp[k - 1] & 0xFFu];   // times x^8
        c->h_mtail.assign(1024, 0);
        c->h_mpre.assign(2048, 0);
        hosttab::fill_mul_table(xp[4u * (c->n & 63u)], c->h_mtail.data());
        const uint32_t cells0 = c->h_seg_cell0[1] - c->h_seg_cell0[0];
        hosttab::fill_mul_table(xp[4ull * cells0 - 1], c->h_mpre.data());
        hosttab::fill_mul_table(xp[4ull * cells0], c->h_mpre.data() + 1024);
        // k_x: per span, the distance from the end of its text to the end of its segment's text; prefix -> end of segment 0
        const size_t nsp = c->h_xspans.size();
        c->h_mspan.assign(nsp * 1024, 0);
        for (size_t sg = 0; sg + 1 < c->h_seg_cell0.size(); ++sg) {
            const uint32_t seg_end = c->h_xoff[c->h_seg_cell0[sg + 1]];
            for (size_t sp = c->h_seg_cell0[sg] / 64; sp < (c->h_seg_cell0[sg + 1] + 63u) / 64u && sp < nsp; ++sp) {
                const uint32_t span_end = c->h_xspans[sp].byte_off + 2u * c->h_xspans[sp].L;
                hosttab::fill_mul_table(xp[seg_end - span_end], &c->h_mspan[sp * 1024]);
            }
        }
        c->h_mpre_x.assign(1024, 0);
        hosttab::fill_mul_table(xp[c->h_xoff[c->h_seg_cell0[1]]], c->h_mpre_x.data());
        c->seg_tabs_dirty = true;
    }
    {   // X rows use the same sample segments; their template is the all-reference X body
        std::vector<uint8_t> xbody;
        xbody.reserve(c->body[kX]);
        for (uint32_t i = 0; i < c->n; ++i) {
            xbody.push_back('0');
            if (c->h_sex[i] != 1) { xbody.push_back('/'); xbody.push_back('0'); }
            xbody.push_back(i + 1 == c->n ? '\n' : '\t');
        }
        c->h_seg_crc_x.clear();
        for (size_t sg = 0; sg + 1 < c->h_seg_cell0.size(); ++sg) {
            const uint32_t b0 = c->h_xoff[c->h_seg_cell0[sg]], b1 = c->h_xoff[c->h_seg_cell0[sg + 1]];
            c->h_seg_crc_x.push_back(raw_crc(xbody.data() + b0, b1 - b0, tab.data()));
        }
    }
    // k_fused_text: balanced byte segments (multiples of 256 bytes) of every class body
    c->text_threads = 64;
    for (int cls = 0; cls < 4; ++cls) {
        c->seg_byte0[cls].clear();
        const uint32_t body = c->body[cls];
        const uint32_t sp = (body + 255u) / 256u;
        const uint32_t ns = (sp + 253u) / 254u;
        const uint32_t pr = (sp + ns - 1u) / ns;
        for (uint32_t sg = 0; sg < ns; ++sg)
            if (sg * pr * 256u < body) c->seg_byte0[cls].push_back(sg * pr * 256u);
        c->seg_byte0[cls].push_back(body);
        c->text_threads = std::max(c->text_threads, (pr + 31u) / 32u * 32u);
    }
}

// 0 = generic three-kernel path, 1 = k_auto, 2 = k_fused_text, 3 = k_x
inline int row_kind(const dnaf_ctx* c, uint64_t r, const uint8_t* hk) {
    if (!c->fused || !c->fused_ok || c->h_plen[r] < 1 || c->h_plen[r] > 64 || c->n == 0) return 0;
    if (c->body[c->h_cls[r]] < kFusedMinRowBytes) return 0;
    if (hk[r] <= 2 && c->h_cls[r] == kAuto) return c->h_pfx_tab[r] ? 1 : 2;
    if (hk[r] <= 2 && c->h_cls[r] == kX) return 3;
    return 2;
}
inline bool row_is_fused(const dnaf_ctx* c, uint64_t r, const uint8_t* hk) { return row_kind(c, r, hk) != 0; }

// What k_auto needs to derive its block descriptors itself when a pass holds autosome rows only: which rows those
// are (host: prefix count of the others), the bucket and first override of every row, the segment table.
int ensure_implicit(dnaf_ctx* c) {
    c->h_other.assign(c->S + 1, 0);
    uint32_t others = 0;
    for (uint64_t r = 0; r < c->S; ++r) {
        c->h_other[r] = others;
        others += row_kind(c, r, c->h_k.data()) != 1;
    }
    c->h_other[c->S] = others;
    if (!c->fused_ok || c->S == 0 || others == c->S) return DNAF_OK;
    std::vector<uint32_t> first(c->S + 1);
    size_t o = 0;
    for (uint64_t r = 0; r <= c->S; ++r) {
        while (o < c->h_orow.size() && c->h_orow[o] < r) ++o;
        first[r] = (uint32_t)o;
    }
    std::vector<uint32_t> seg;
    for (size_t sg = 0; sg + 1 < c->h_seg_cell0.size(); ++sg) {
        seg.push_back(c->h_seg_cell0[sg]);
        seg.push_back(c->h_seg_cell0[sg + 1] - c->h_seg_cell0[sg]);
        seg.push_back(c->h_seg_crc[sg]);
    }
    int rc = upload(c, c->d_bucket, c->h_bucket.data(), c->h_bucket.size(), false);
    if (!rc) rc = upload(c, c->d_ovr_first, first.data(), first.size(), false);
    if (!rc) rc = upload(c, c->d_seginfo, seg.data(), seg.size());
    return rc;
}

// BGZF block plan of one pass (rows [r0,r1)): fused segments and generic blocks, slots in row order.
void plan_pass(dnaf_ctx* c, uint64_t r0, uint64_t r1, const uint8_t* hk) {
    c->fplan.clear();
    c->xplan.clear();
    c->tplan.clear();
    c->plan.clear();
    c->gslot.clear();
    c->grow.clear();
    c->goff.clear();
    c->olocal.clear();
    c->osub.clear();
    uint64_t gtext = 0;
    uint32_t slot = 0;
    size_t o = std::lower_bound(c->h_orow.begin(), c->h_orow.end(), r0) - c->h_orow.begin();
    uint64_t r = r0;
    while (r < r1) {
        while (o < c->h_orow.size() && c->h_orow[o] < r) ++o;
        const int kind = row_kind(c, r, hk);
        if (kind == 2) {
            size_t oe = o;
            while (oe < c->h_orow.size() && c->h_orow[oe] == r) ++oe;
            const std::vector<uint32_t>& sb = c->seg_byte0[c->h_cls[r]];
            const size_t nseg = sb.size() - 1;
            for (size_t sgi = 0; sgi < nseg; ++sgi) {
                TextDesc d;
                d.row = r;
                d.byte0 = sb[sgi];
                d.nbytes = sb[sgi + 1] - sb[sgi];
                d.slot = slot++;
                d.flags = (sgi == 0 ? 1u : 0u) | (sgi + 1 == nseg ? 2u : 0u);
                d.ovr_first = (uint32_t)o;
                d.ovr_count = (uint32_t)(oe - o);
                d.table = (uint32_t)c->h_bucket[r] * kVariants + 2u + 2u * c->h_cls[r] + (sgi == 0 ? 0u : 1u);
                d.pad = 0;
                c->tplan.push_back(d);
            }
            o = oe;
            ++r;
            continue;
        }
        if (kind == 1 || kind == 3) {
            size_t oe = o;
            while (oe < c->h_orow.size() && c->h_orow[oe] == r) ++oe;
            const size_t nseg = c->h_seg_crc.size();
            for (size_t sgi = 0; sgi < nseg; ++sgi) {
                FusedDesc d;
                d.row = r;
                d.cell0 = c->h_seg_cell0[sgi];
                d.ncells = c->h_seg_cell0[sgi + 1] - d.cell0;
                d.slot = slot++;
                d.flags = (sgi == 0 ? 1u : 0u) | (sgi + 1 == nseg ? 2u : 0u);
                d.ovr_first = (uint32_t)o;
                d.ovr_count = (uint32_t)(oe - o);
                if (kind == 1) {
                    d.table = (uint32_t)c->h_bucket[r] * 2u + (sgi == 0 ? 0u : 1u);
                    d.body_crc = c->h_seg_crc[sgi];
                    c->fplan.push_back(d);
                } else {
                    d.table = (uint32_t)c->h_bucket[r] * 2u + (sgi == 0 ? 0u : 1u);
                    d.body_crc = c->h_seg_crc_x[sgi];
                    c->xplan.push_back(d);
                }
            }
            o = oe;
            ++r;
            continue;
        }
        // a run of consecutive generic rows: text laid out back to back in the generic text buffer
        const uint64_t run_begin = r;
        while (r < r1 && !row_is_fused(c, r, hk)) {
            c->grow.push_back((uint32_t)(r - r0));
            c->goff.push_back(gtext);
            while (o < c->h_orow.size() && c->h_orow[o] == r) {
                c->olocal.push_back((uint32_t)(c->grow.size() - 1));
                c->osub.push_back(c->h_osamp[o]);
                ++o;
            }
            gtext += c->h_row_off[r + 1] - c->h_row_off[r];
            ++r;
        }
        uint64_t q = run_begin;
        size_t gi = c->grow.size() - (size_t)(r - run_begin);
        while (q < r) {
            const uint64_t off = c->goff[gi];
            const uint64_t len = c->h_row_off[q + 1] - c->h_row_off[q];
            const uint32_t plen = c->h_plen[q];
            if (len > kBlk) {
                uint64_t done = 0;
                if (plen + kSpan <= kBlk) {
                    const uint64_t first = plen + (uint64_t)((kBlk - plen) / kSpan) * kSpan;
                    c->plan.push_back({off, (uint32_t)std::min<uint64_t>(first, len), plen});
                    c->gslot.push_back(slot++);
                    done = std::min<uint64_t>(first, len);
                }
                while (done < len) {
                    const uint32_t piece = (uint32_t)std::min<uint64_t>(kBlk, len - done);
                    c->plan.push_back({off + done, piece, 0});
                    c->gslot.push_back(slot++);
                    done += piece;
                }
                ++q;
                ++gi;
            } else {
                uint64_t acc = 0;
                while (q < r && acc + (c->h_row_off[q + 1] - c->h_row_off[q]) <= kBlk) {
                    acc += c->h_row_off[q + 1] - c->h_row_off[q];
                    ++q;
                    ++gi;
                }
                c->plan.push_back({off, (uint32_t)acc, std::min<uint32_t>(plen, (uint32_t)acc)});
                c->gslot.push_back(slot++);
            }
        }
    }
    c->gen_text_bytes = gtext;
    c->pass_blocks = slot;
    // Slot stride of the pass: the longest block's text (a stored block is the worst case: text + 5) plus the slot
    // lead, BGZF framing, the zero-fill / copy overrun of the kernels (< 64 bytes), rounded up to 256.
    uint32_t longest = 0;
    for (const FusedDesc& d : c->fplan) longest = std::max(longest, 4u * d.ncells + 66u);           // prefix <= 64, +1 lead, +1
    for (const FusedDesc& d : c->xplan) longest = std::max(longest, 4u * d.ncells + 66u);
    for (const TextDesc& d : c->tplan) longest = std::max(longest, d.nbytes + 66u);
    for (const BlockDesc& b : c->plan) longest = std::max(longest, b.len);
    c->slot_stride = std::min<uint32_t>(kSlot, (longest + 128u + 255u) & ~255u);
    c->pass_text = c->h_row_off[r1] - c->h_row_off[r0];
}

struct Sink {
    dnaf_sink_fn fn = nullptr;
    void* user = nullptr;
    uint8_t* buf = nullptr;  // host buffer mode
    uint64_t cap = 0, used = 0;
    bool device_only = false;
    bool pinned = false;     // buf is page-locked host memory
    int fd = -1;             // file descriptor mode: write() straight from the page-locked staging buffer
    int64_t fd_off = -1;     // >= 0: pwrite() at this file offset instead (advanced as pieces land)
    bool log = false;        // append the blocks to the context's block log as they reach the host
};

// Walks whole BGZF blocks in [data, data+n): compressed size from BSIZE (the BC subfield), text size from ISIZE.
// Returns the number of bytes covered by well-formed blocks (== n for a clean stream).
template <class F>
uint64_t walk_bgzf(const uint8_t* data, uint64_t n, F&& on_block) {
    uint64_t o = 0;
    while (o + 28 <= n) {
        const uint8_t* h = data + o;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C') break;
        const uint32_t csize = (uint32_t)(h[16] | (h[17] << 8)) + 1u;
        if (csize < 26 || o + csize > n) break;
        const uint8_t* t = h + csize - 4;
        const uint32_t usize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        on_block(csize, usize);
        o += csize;
    }
    return o;
}

int append_block_log(dnaf_ctx* c, const uint8_t* data, uint64_t n) {
    const uint64_t covered = walk_bgzf(data, n, [&](uint32_t cs, uint32_t us) {
        c->log_csize.push_back(cs);
        c->log_usize.push_back(us);
    });
    if (covered != n) return fail(c, DNAF_E_CUDA, "block log: pass output is not a whole number of BGZF blocks");
    return DNAF_OK;
}

int deliver(dnaf_ctx* c, Sink& s, const uint8_t* data, uint64_t n) {
    if (s.fd >= 0) {
        uint64_t done = 0;
        while (done < n) {
            const size_t piece = (size_t)std::min<uint64_t>(n - done, 1u << 30);
            const ssize_t w = s.fd_off >= 0 ? ::pwrite(s.fd, data + done, piece, (off_t)(s.fd_off + (int64_t)done)) : ::write(s.fd, data + done, piece);
            if (w < 0) {
                if (errno == EINTR) continue;
                return fail(c, DNAF_E_SINK, "write to file descriptor %d failed: %s", s.fd, strerror(errno));
            }
            done += (uint64_t)w;
        }
        if (s.fd_off >= 0) s.fd_off += (int64_t)n;
    } else if (s.fn) {
        if (s.fn(s.user, data, n) != 0) return fail(c, DNAF_E_SINK, "sink callback failed");
    } else if (s.buf) {
        if (s.used + n > s.cap) return fail(c, DNAF_E_SPACE, "output buffer too small: need more than %llu bytes",
                                            (unsigned long long)s.cap);
        memcpy(s.buf + s.used, data, n);
    }
    s.used += n;
    return DNAF_OK;
}

// Uploads a host vector through the pass's page-locked staging arena, so the copy is asynchronous and the
// host can go on planning while the previous pass still runs.  reserve_stage() sizes the arena up front.
template <class T>
int upload_async(dnaf_ctx* c, DevBuf& b, const std::vector<T>& v) {
    CU(c, b.reserve(std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (v.empty()) return DNAF_OK;
    dnaf_ctx::OutBuf& B = c->ob[c->cur_ob];
    const size_t bytes = v.size() * sizeof(T);
    const size_t at = (B.stage_used + 63) & ~size_t(63);
    if (at + bytes > B.h_stage.cap) {  // not planned for: fall back to a pageable (synchronising) copy
        CU(c, cudaMemcpyAsync(b.p, v.data(), bytes, cudaMemcpyHostToDevice, c->stream));
        return DNAF_OK;
    }
    memcpy(B.h_stage.as<uint8_t>() + at, v.data(), bytes);
    B.stage_used = at + bytes;
    CU(c, cudaMemcpyAsync(b.p, B.h_stage.as<uint8_t>() + at, bytes, cudaMemcpyHostToDevice, c->stream));
    return DNAF_OK;
}

int reserve_stage(dnaf_ctx* c, dnaf_ctx::OutBuf& B) {
    const size_t need = (c->fplan.size() + c->xplan.size()) * sizeof(FusedDesc) + c->tplan.size() * sizeof(TextDesc) +
                        c->plan.size() * sizeof(BlockDesc) + (c->gslot.size() + c->grow.size() + c->olocal.size() +
                        c->osub.size()) * 4 + c->goff.size() * 8 + 1024;
    if (need > B.h_stage.cap) {
        CU(c, cudaEventSynchronize(B.ev[5]));  // the arena may still feed the previous use of this buffer
        CU(c, B.h_stage.reserve(need * 2));
    }
    B.stage_used = 0;
    return DNAF_OK;
}

int reserve_outputs(dnaf_ctx* c, dnaf_ctx::OutBuf& B, uint32_t nb_exact, uint64_t text_bytes) {
    // whole multiples of 2048 blocks: passes of a job differ a little in block count, buffers must not be
    // re-allocated (cudaMalloc synchronises the device) every time one is a few blocks larger than the last
    const uint32_t nb = nb_exact > 256u ? (nb_exact + 2047u) / 2048u * 2048u : nb_exact;
    CU(c, c->sbuf[c->sb].slots.reserve((size_t)nb * c->slot_stride + 256));
    CU(c, c->sbuf[c->sb].sizes.reserve(nb * sizeof(uint32_t)));
    CU(c, c->sbuf[c->sb].crcs.reserve(nb * sizeof(uint32_t)));
    CU(c, B.d_totals.reserve((2 + 2 * (size_t)((nb + kGroup - 1u) / kGroup)) * sizeof(uint64_t)));   // state of k_size_partials / k_gather
    CU(c, B.d_out.reserve(text_bytes + (size_t)nb * 64 + 256));   // worst case: every block stored
    CU(c, B.h_totals.reserve(2 * sizeof(uint64_t)));
    if (!c->attr_done) {
        CU(c, cudaFuncSetAttribute(k_bgzf_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateSmem)));
        c->attr_done = true;
    }
    return DNAF_OK;
}

// generic encoder over c->plan (text in d_text), slots from c->gslot (or 0..n-1 when empty)
int launch_generic(dnaf_ctx* c, dnaf_stats* st) {
    const uint32_t nb = (uint32_t)c->plan.size();
    if (!nb) return DNAF_OK;
    int rc = upload_async(c, c->d_blocks, c->plan);
    if (!rc) rc = upload_async(c, c->d_gslot, c->gslot);
    if (rc) return rc;
    k_bgzf_generic<<<nb, 256, sizeof(DeflateSmem), c->stream>>>(
        c->d_text.as<uint8_t>(), c->d_blocks.as<BlockDesc>(), c->gslot.empty() ? nullptr : c->d_gslot.as<uint32_t>(),
        c->d_crctab.as<uint32_t>(), c->d_xpow8.as<uint32_t>(), c->sbuf[c->sb].slots.as<uint8_t>(), c->slot_stride, c->sbuf[c->sb].sizes.as<uint32_t>(),
        c->sbuf[c->sb].crcs.as<uint32_t>());
    if (st) st->kernel_launches += 1;
    CU(c, cudaGetLastError());
    return DNAF_OK;
}

// Compaction of the pass whose kernels were just queued on the main stream (ev[4] marks their end): sizes -> offsets ->
// gather into B.d_out, on the compaction stream, so that it overlaps the next pass's kernels.  ev[5] = pass done.
int close_pass(dnaf_ctx* c, dnaf_ctx::OutBuf& B, uint32_t nb, dnaf_stats* st) {
    B.nb = nb;
    dnaf_ctx::SlotBuf& S = c->sbuf[c->sb];
    CU(c, cudaStreamWaitEvent(c->comp, B.ev[4], 0));
    if (nb) {
        const uint32_t ntiles = (nb + kTile - 1u) / kTile, ngroups = (nb + kGroup - 1u) / kGroup;
        k_size_partials<<<ngroups, kGroup, 0, c->comp>>>(S.sizes.as<uint32_t>(), S.crcs.as<uint32_t>(), nb,
                                                         reinterpret_cast<unsigned long long*>(B.d_totals.p));
        k_gather<<<ntiles, 256, 0, c->comp>>>(S.slots.as<uint8_t>(), c->slot_stride, S.sizes.as<uint32_t>(), nb,
                                              reinterpret_cast<unsigned long long*>(B.d_totals.p),
                                              reinterpret_cast<unsigned long long*>(B.h_totals.p), B.d_out.as<uint8_t>());
        if (st) st->kernel_launches += 2;
    }
    CU(c, cudaEventRecord(B.ev[5], c->comp));
    CU(c, cudaEventRecord(S.ev_free, c->comp));
    CU(c, cudaGetLastError());
    c->sb ^= 1;   // the next pass writes the other slot set
    return DNAF_OK;
}

// A closed pass: wait for its kernels and totals, account it, and START moving its bytes to the host (straight
// into a page-locked caller buffer when there is one).  finish_copy() completes the move.
int start_copy(dnaf_ctx* c, dnaf_ctx::OutBuf& B, Sink& sink, dnaf_stats* st) {
    trace("start_copy: wait kernels", (int)B.nb);
    CU(c, cudaEventSynchronize(B.ev[5]));
    trace("start_copy: kernels done", (int)B.nb);
    const uint64_t bytes = B.nb ? *reinterpret_cast<volatile uint64_t*>(B.h_totals.p) : 0;
    if (st) {
        st->bgzf_bytes += bytes;
        st->bgzf_blocks += B.nb;
        if (B.nb) st->crc_xor ^= (uint32_t)reinterpret_cast<volatile uint64_t*>(B.h_totals.p)[1];
        float t01 = 0, t12 = 0, t23 = 0, t34 = 0, t45 = 0, t05 = 0;
        cudaEventElapsedTime(&t01, B.ev[0], B.ev[1]);
        cudaEventElapsedTime(&t12, B.ev[1], B.ev[2]);
        cudaEventElapsedTime(&t23, B.ev[2], B.ev[3]);
        cudaEventElapsedTime(&t34, B.ev[3], B.ev[4]);
        cudaEventElapsedTime(&t45, B.ev[4], B.ev[5]);
        cudaEventElapsedTime(&t05, B.ev[0], B.ev[5]);
        if (B.gen) {
            st->ms_sample += t01;
            st->ms_format += t12;
        }
        st->ms_deflate += (B.generic_blocks ? t23 : 0.f) + t45;
        if (B.fused) st->ms_fused += t34;
        st->ms_total += t05;
        if (B.auto_text) {
            float ta = 0;
            cudaEventElapsedTime(&ta, B.ev_auto[0], B.ev_auto[1]);
            st->ms_auto += ta;
            st->auto_launches += 1;
            st->auto_text_bytes += B.auto_text;
        }
        st->rows += B.rows;
        st->text_bytes += B.text;
    }
    B.copy_mode = 0;
    B.copy_bytes = bytes;
    if (sink.device_only || !bytes) {
        sink.used += bytes;
        return DNAF_OK;
    }
    if (sink.buf && sink.pinned) {  // no staging copy: DMA straight into the caller's page-locked buffer
        if (sink.used + bytes > sink.cap)
            return fail(c, DNAF_E_SPACE, "output buffer too small: need more than %llu bytes", (unsigned long long)sink.cap);
        CU(c, cudaMemcpyAsync(sink.buf + sink.used, B.d_out.p, bytes, cudaMemcpyDeviceToHost, c->copy));
        CU(c, cudaEventRecord(B.ev_copied, c->copy));
        B.copy_dst = sink.buf + sink.used;
        sink.used += bytes;
        B.copy_mode = 1;
        return DNAF_OK;
    }
    CU(c, B.h_out.reserve(bytes));
    CU(c, cudaMemcpyAsync(B.h_out.p, B.d_out.p, bytes, cudaMemcpyDeviceToHost, c->copy));
    CU(c, cudaEventRecord(B.ev_copied, c->copy));
    B.copy_mode = 2;
    return DNAF_OK;
}

int finish_copy(dnaf_ctx* c, dnaf_ctx::OutBuf& B, Sink& sink) {
    const int mode = B.copy_mode;
    B.copy_mode = 0;
    if (!mode) return DNAF_OK;
    trace("finish_copy: wait", (int)B.nb);
    CU(c, cudaEventSynchronize(B.ev_copied));
    trace("finish_copy: done", (int)B.nb);
    if (sink.log) {
        const int rc = append_block_log(c, mode == 2 ? B.h_out.as<uint8_t>() : B.copy_dst, B.copy_bytes);
        if (rc) return rc;
    }
    if (mode == 2) return deliver(c, sink, B.h_out.as<uint8_t>(), B.copy_bytes);
    return DNAF_OK;
}

// sample (+ overrides) `rows` rows into the plane buffers; row list optional (d_grow), overrides as local pairs
int run_sample(dnaf_ctx* c, uint64_t r0, uint32_t rows, const uint32_t* d_row_idx, uint64_t seed, uint64_t n_over,
               const uint32_t* d_olocal, const uint32_t* d_osamp, dnaf_stats* st) {
    const SampleView sv = sample_view(c);
    const uint64_t words = (uint64_t)rows * sv.groups;
    CU(c, c->d_plane0.reserve(std::max<uint64_t>(words, 1) * 4));
    if (c->any_multi) CU(c, c->d_plane1.reserve(std::max<uint64_t>(words, 1) * 4));
    uint32_t* p1 = c->any_multi ? c->d_plane1.as<uint32_t>() : nullptr;
    if (words) {
        const uint32_t grid = (uint32_t)((words + 255) / 256);
        k_sample<<<grid, 256, 0, c->stream>>>(sv, snp_view(c), r0, d_row_idx, c->row_base, rows, (uint32_t)seed,
                                              (uint32_t)(seed >> 32), c->d_plane0.as<uint32_t>(), p1);
        if (st) st->kernel_launches += 1;
        if (n_over) {
            k_overrides<<<(uint32_t)((n_over + 255) / 256), 256, 0, c->stream>>>(d_olocal, d_osamp, n_over, sv.groups, c->n,
                                                                              c->d_plane0.as<uint32_t>(), p1);
            if (st) st->kernel_launches += 1;
        }
    }
    CU(c, cudaGetLastError());
    return DNAF_OK;
}

// overrides of rows [r0,r1) as (local row, sample) device arrays (all rows, no subset)
int stage_overrides_all(dnaf_ctx* c, uint64_t r0, uint64_t r1, uint64_t* n_over) {
    const size_t o0 = std::lower_bound(c->h_orow.begin(), c->h_orow.end(), r0) - c->h_orow.begin();
    const size_t o1 = std::lower_bound(c->h_orow.begin(), c->h_orow.end(), r1) - c->h_orow.begin();
    c->olocal.clear();
    c->osub.clear();
    for (size_t o = o0; o < o1; ++o) {
        c->olocal.push_back((uint32_t)(c->h_orow[o] - r0));
        c->osub.push_back(c->h_osamp[o]);
    }
    *n_over = o1 - o0;
    return DNAF_OK;
}

// uploads the (local row, sample) override pairs staged in c->olocal / c->osub
int upload_overrides(dnaf_ctx* c) {
    if (c->olocal.empty()) return DNAF_OK;
    int rc = upload_async(c, c->d_olocal, c->olocal);
    if (!rc) rc = upload_async(c, c->d_osub, c->osub);
    return rc;
}

int run_format(dnaf_ctx* c, uint64_t r0, uint32_t rows, const uint32_t* d_row_idx, const uint64_t* d_sub_off,
               uint64_t text_bytes, dnaf_stats* st) {
    CU(c, c->d_text.reserve(text_bytes + 64));
    if (!rows) return DNAF_OK;
    k_format<<<rows, 256, 0, c->stream>>>(sample_view(c), snp_view(c), r0, d_row_idx, d_sub_off,
                                          c->d_row_off.as<uint64_t>(), c->h_row_off[r0], c->d_plane0.as<uint32_t>(),
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

int generate_passes(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                    Sink& sink, dnaf_stats* st);

// The pass pipeline keeps up to three passes in flight.  When a call fails half way (sink error, caller's buffer
// too small, CUDA error) nothing of it may still be running when the error is returned: a copy could be landing in a
// caller buffer that is about to be freed, and the next call must find an idle pipeline.
int generate_impl(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                  Sink& sink, dnaf_stats* st) {
    const int rc = generate_passes(c, row_begin, row_end, seed, level, sink, st);
    if (rc && c) {
        cudaSetDevice(c->dev);
        cudaStreamSynchronize(c->stream);   // may be the caller's stream (dnaf_set_stream), NULL = the default stream
        for (cudaStream_t s : {c->side, c->side2, c->comp, c->copy})
            if (s) cudaStreamSynchronize(s);
        cudaGetLastError();
        for (auto& b : c->ob) b.copy_mode = 0;
    }
    return rc;
}

int generate_passes(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                    Sink& sink, dnaf_stats* st) {
    if (!c) return DNAF_E_ARG;
    if (level < 1 || level > 9) return fail(c, DNAF_E_ARG, "level must be 1..9");
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    CU(c, cudaSetDevice(c->dev));
    // -z 1..3: the byte-4-back parse (k_auto); -z 4..9: LZ77 tiers of growing search depth (k_lz) on autosome rows
    static const int lz_off = getenv("DNAF_NO_LZ") ? 1 : 0;
    rc = ensure_lz_tables(c, lz_off ? 1 : level);
    if (rc) return rc;
    const bool use_lz = c->lz_ok;
    dnaf_stats local;
    memset(&local, 0, sizeof local);
    sink.log = c->log_blocks && !sink.device_only;
    if (sink.buf) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, sink.buf) == cudaSuccess) sink.pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    uint64_t r0 = row_begin;
    // Three output buffers in rotation.  Pass i is launched as soon as the copy of pass i-3 (same buffer) has landed,
    // i.e. without waiting for anything recent, so the GPU runs ahead; then the copy of pass i-1 is queued behind
    // the copy of pass i-2 that is still in flight, so the copy engine never waits for the host either.
    for (auto& b : c->ob) b.copy_mode = 0;
    int npass = 0;
    if (g_trace) { g_t0 = std::chrono::steady_clock::now(); trace("generate begins", 0); }
    while (r0 < row_end) {
        const int cur = npass % 3;
        rc = finish_copy(c, c->ob[cur], sink);
        if (rc) return rc;
        // the first passes of a call are short (1/4, 1/2 of a chunk; 1/2 when nothing leaves the device): the GPU starts
        // while the host still plans and the first D2H copy starts early.  Measured (bench.py, 32768-row calls): ramp
        // 3 / 2 / 1 / 0 -> 4.90 / 5.00 / 5.09 / 5.09e11 calls/s on the device, end to end unchanged.
        static const int ramp_env = getenv("DNAF_RAMP") ? atoi(getenv("DNAF_RAMP")) : -1;
        const int ramp = ramp_env >= 0 ? ramp_env : (sink.device_only ? 1 : 2);
        const uint64_t r1 = next_chunk_end(c, r0, row_end, std::max<uint64_t>(c->chunk_bytes >> std::max(0, ramp - npass), 4096));
        dnaf_ctx::OutBuf& B = c->ob[cur];
        const auto t_plan0 = std::chrono::steady_clock::now();
        c->implicit_pass = c->fused_ok && c->h_other[r1] == c->h_other[r0] && !c->h_seg_crc.empty();
        if (c->implicit_pass) {   // autosome rows only: k_auto derives its descriptors, the host plans nothing
            c->fplan.clear(); c->xplan.clear(); c->tplan.clear(); c->plan.clear(); c->gslot.clear(); c->grow.clear();
            c->goff.clear(); c->olocal.clear(); c->osub.clear();
            c->gen_text_bytes = 0;
            const uint32_t nseg = (uint32_t)c->h_seg_crc.size();
            c->pass_blocks = (uint32_t)(r1 - r0) * nseg;
            uint32_t longest = 0;
            for (uint32_t sg = 0; sg < nseg; ++sg) longest = std::max(longest, 4u * (c->h_seg_cell0[sg + 1] - c->h_seg_cell0[sg]) + 66u);
            c->slot_stride = std::min<uint32_t>(kSlot, (longest + 128u + 255u) & ~255u);
            c->pass_text = c->h_row_off[r1] - c->h_row_off[r0];
        } else {
            plan_pass(c, r0, r1, c->h_k.data());
        }
        const auto t_plan1 = std::chrono::steady_clock::now();
        c->cur_ob = cur;
        rc = reserve_outputs(c, B, c->pass_blocks, c->pass_text);
        if (!rc) rc = reserve_stage(c, B);
        if (rc) return rc;
        CU(c, cudaStreamWaitEvent(c->stream, c->sbuf[c->sb].ev_free, 0));   // compaction two passes ago read this slot set
        CU(c, cudaEventRecord(B.ev[0], c->stream));
        const uint32_t grows = (uint32_t)c->grow.size();
        if (grows) {
            rc = upload_async(c, c->d_grow, c->grow);
            if (!rc) rc = upload_async(c, c->d_goff, c->goff);
            if (!rc) rc = upload_overrides(c);
            if (!rc) rc = run_sample(c, r0, grows, c->d_grow.as<uint32_t>(), seed, c->olocal.size(),
                                     c->d_olocal.as<uint32_t>(), c->d_osub.as<uint32_t>(), &local);
            if (rc) return rc;
        }
        CU(c, cudaEventRecord(B.ev[1], c->stream));
        if (grows) {
            rc = run_format(c, r0, grows, c->d_grow.as<uint32_t>(), c->d_goff.as<uint64_t>(), c->gen_text_bytes, &local);
            if (rc) return rc;
        }
        CU(c, cudaEventRecord(B.ev[2], c->stream));
        rc = launch_generic(c, &local);
        if (rc) return rc;
        CU(c, cudaEventRecord(B.ev[3], c->stream));
        // descriptors of the three fused kernels go up first; the few, long blocks of k_fused_text / k_fused_x then start
        // on the high-priority side stream and k_auto fills the rest of the chip from the main stream
        if (!c->tplan.empty()) rc = upload_async(c, c->d_tdesc, c->tplan);
        if (!rc && !c->xplan.empty()) rc = upload_async(c, c->d_xdesc, c->xplan);
        if (!rc && !c->fplan.empty()) rc = upload_async(c, c->d_fdesc, c->fplan);
        if (rc) return rc;
        const bool side_work = !c->tplan.empty() || !c->xplan.empty();
        if (side_work) {
            CU(c, cudaEventRecord(c->ev_fork, c->stream));
            if (!c->tplan.empty()) CU(c, cudaStreamWaitEvent(c->side, c->ev_fork, 0));
            if (!c->xplan.empty()) CU(c, cudaStreamWaitEvent(c->side2, c->ev_fork, 0));
        }
        if (!c->tplan.empty()) {
            if (!c->text_attr_done) {
                CU(c, cudaFuncSetAttribute(k_fused_text, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TextSmem)));
                c->text_attr_done = true;
            }
            TextArgs ta;
            ta.sv = sample_view(c);
            ta.nv = snp_view(c);
            ta.desc = c->d_tdesc.as<TextDesc>();
            ta.tables = c->d_ftables.as<FusedTable>();
            ta.crc4 = c->d_crc4.as<uint32_t>();
            ta.xpow8 = c->d_xpow8.as<uint32_t>();
            ta.osamp = c->d_osamp.as<uint32_t>();
            ta.xspan = c->d_xspan.as<uint32_t>();
            ta.row_base = c->row_base;
            ta.k0 = (uint32_t)seed;
            ta.k1 = (uint32_t)(seed >> 32);
            ta.slots = c->sbuf[c->sb].slots.as<uint8_t>();
            ta.slot_stride = c->slot_stride;
            ta.sizes = c->sbuf[c->sb].sizes.as<uint32_t>();
            ta.crcs = c->sbuf[c->sb].crcs.as<uint32_t>();
            k_fused_text<<<(uint32_t)c->tplan.size(), c->text_threads, sizeof(TextSmem), c->side>>>(ta);
            local.kernel_launches += 1;
            CU(c, cudaGetLastError());
        }
        if (!c->xplan.empty()) {
            XArgs xa;
            xa.sv = sample_view(c);
            xa.nv = snp_view(c);
            xa.desc = c->d_xdesc.as<FusedDesc>();
            xa.tables = c->d_xtables.as<XTable>();
            xa.xspans = c->d_xspans.as<XSpan>();
            xa.etab = c->d_etab.as<uint32_t>();
            xa.mspan = c->d_mspan.as<uint32_t>();
            xa.mpre = c->d_mpre_x.as<uint32_t>();
            xa.xinit = c->d_xinit.as<uint32_t>();
            xa.pre_crc = c->d_pre_crc.as<uint32_t>();
            xa.orow = c->d_orow.as<uint64_t>();
            xa.osamp = c->d_osamp.as<uint32_t>();
            xa.row_base = c->row_base;
            xa.k0 = (uint32_t)seed;
            xa.k1 = (uint32_t)(seed >> 32);
            xa.slots = c->sbuf[c->sb].slots.as<uint8_t>();
            xa.slot_stride = c->slot_stride;
            xa.sizes = c->sbuf[c->sb].sizes.as<uint32_t>();
            xa.crcs = c->sbuf[c->sb].crcs.as<uint32_t>();
            k_x<<<(uint32_t)c->xplan.size(), c->fused_threads, x_smem_bytes(c->fused_threads), c->side2>>>(xa);
            local.kernel_launches += 1;
            CU(c, cudaGetLastError());
        }
        if (!c->tplan.empty()) CU(c, cudaEventRecord(c->ev_join, c->side));
        if (!c->xplan.empty()) CU(c, cudaEventRecord(c->ev_join2, c->side2));
        if (!c->fplan.empty() || c->implicit_pass) {
            AutoArgs fa;
            fa.sv = sample_view(c);
            fa.nv = snp_view(c);
            fa.desc = c->implicit_pass ? nullptr : c->d_fdesc.as<FusedDesc>();
            fa.row0 = r0;
            fa.nseg = (uint32_t)c->h_seg_crc.size();
            fa.nseg_magic = fa.nseg > 1 ? (uint32_t)(((1ull << 32) + fa.nseg - 1) / fa.nseg) : 0u;
            fa.seginfo = c->d_seginfo.as<uint32_t>();
            fa.bucket = c->d_bucket.as<uint16_t>();
            fa.ovr_first = c->d_ovr_first.as<uint32_t>();
            fa.tables = c->d_atables.as<AutoTable>();
            fa.etab = c->d_etab2.as<uint32_t>();
            fa.mtab = c->d_mtab.as<uint32_t>();
            fa.mtail = c->d_mtail.as<uint32_t>();
            fa.mpre = c->d_mpre.as<uint32_t>();
            fa.crctab = c->d_crctab.as<uint32_t>();
            fa.xinit = c->d_xinit.as<uint32_t>();
            fa.pre_crc = c->d_pre_crc.as<uint32_t>();
            fa.orow = c->d_orow.as<uint64_t>();
            fa.osamp = c->d_osamp.as<uint32_t>();
            fa.row_base = c->row_base;
            fa.k0 = (uint32_t)seed;
            fa.k1 = (uint32_t)(seed >> 32);
            fa.slots = c->sbuf[c->sb].slots.as<uint8_t>();
            fa.slot_stride = c->slot_stride;
            fa.sizes = c->sbuf[c->sb].sizes.as<uint32_t>();
            fa.crcs = c->sbuf[c->sb].crcs.as<uint32_t>();
            CU(c, cudaEventRecord(B.ev_auto[0], c->stream));
            const uint32_t ablocks = c->implicit_pass ? c->pass_blocks : (uint32_t)c->fplan.size();
            if (use_lz) {
                const uint32_t smem = lz_smem_bytes(c->fused_threads, kLzMaxKey + 1u);
                if (!c->lz_attr_done) {
                    CU(c, cudaFuncSetAttribute(k_lz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz_smem_bytes(256, kLzMaxKey + 1u)));
                    c->lz_attr_done = true;
                }
                LzArgs la;
                la.a = fa;
                la.tables = c->d_ltables.as<LzTable>();
                const LzCfg cfg = lz_cfg(level, kLzMaxKey);
                la.chain = cfg.chain;
                la.lazy = cfg.lazy;
                la.nice = cfg.nice;
                k_lz<<<ablocks, c->fused_threads, smem, c->stream>>>(la);
            } else {
                k_auto<<<ablocks, c->fused_threads, auto_smem_bytes(c->fused_threads), c->stream>>>(fa);
            }
            CU(c, cudaEventRecord(B.ev_auto[1], c->stream));
            if (c->implicit_pass) {
                B.auto_text = c->pass_text;
            } else {   // text of the planned k_auto blocks: prefix on a row's first block, cells, '\n' for '\t' at the row's end
                uint64_t t = 0;
                for (const FusedDesc& d : c->fplan) t += 4ull * d.ncells + ((d.flags & 1u) ? c->h_plen[d.row] : 0u);
                B.auto_text = t;
            }
            local.kernel_launches += 1;
            CU(c, cudaGetLastError());
        }
        if (!c->tplan.empty()) CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        if (!c->xplan.empty()) CU(c, cudaStreamWaitEvent(c->stream, c->ev_join2, 0));
        CU(c, cudaEventRecord(B.ev[4], c->stream));
        if (c->fplan.empty() && !c->implicit_pass) B.auto_text = 0;
        B.rows = r1 - r0;
        B.text = c->h_row_off[r1] - c->h_row_off[r0];
        B.gen = grows != 0;
        B.generic_blocks = !c->plan.empty();
        B.fused = c->implicit_pass || !c->fplan.empty() || !c->tplan.empty() || !c->xplan.empty();
        rc = close_pass(c, B, c->pass_blocks, &local);
        if (rc) return rc;
        if (g_trace) {
            const auto t_l = std::chrono::steady_clock::now();
            fprintf(stderr, "[dnaf] pass rows %llu: plan %.0f us, launch %.0f us\n", (unsigned long long)(r1 - r0),
                    std::chrono::duration<double, std::micro>(t_plan1 - t_plan0).count(),
                    std::chrono::duration<double, std::micro>(t_l - t_plan1).count());
            trace("launched pass", npass);
        }
        if (npass >= 1) {
            rc = start_copy(c, c->ob[(npass - 1) % 3], sink, &local);
            if (rc) return rc;
        }
        ++npass;
        r0 = r1;
    }
    if (npass >= 1) {
        rc = start_copy(c, c->ob[(npass - 1) % 3], sink, &local);
        if (rc) return rc;
    }
    for (int k = std::max(0, npass - 3); k < npass; ++k) {   // in pass order: staged sinks are delivered here
        rc = finish_copy(c, c->ob[k % 3], sink);
        if (rc) return rc;
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
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&c->side2, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    if ((e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    {   // high priority: the compaction of pass i must not queue behind the blocks of pass i+1's kernels
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((e = cudaStreamCreateWithPriority(&c->comp, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    for (auto& sbf : c->sbuf)
        if ((e = cudaEventCreateWithFlags(&sbf.ev_free, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (auto& b : c->ob) {
        for (auto& ev : b.ev)
            if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
        for (auto& ev : b.ev_auto)
            if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&b.ev_copied, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    if ((e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
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
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->side2) { cudaStreamSynchronize(c->side2); cudaStreamDestroy(c->side2); }
    if (c->ev_join2) cudaEventDestroy(c->ev_join2);
    if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
    if (c->comp) { cudaStreamSynchronize(c->comp); cudaStreamDestroy(c->comp); }
    for (auto& sbf : c->sbuf)
        if (sbf.ev_free) cudaEventDestroy(sbf.ev_free);
    for (auto& b : c->ob) {
        for (auto& ev : b.ev)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : b.ev_auto)
            if (ev) cudaEventDestroy(ev);
        if (b.ev_copied) cudaEventDestroy(b.ev_copied);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
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
    c->layout_ok = false;   // which rows take which kernel changes
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
    c->h_sex.assign(sex, sex + n);
    c->h_xoff = xoff;
    c->samples_epoch++;
    c->h_xspans = hosttab::build_xspans(sex, n, xoff.data());
    rc = upload(c, c->d_xspans, c->h_xspans.data(), c->h_xspans.size());
    if (rc) return rc;
    {   // X rows: the sample that holds body byte 256*k (for k_fused_text)
        std::vector<uint32_t> xspan((size_t)acc / 256 + 2, 0);
        uint32_t i = 0;
        for (size_t k = 0; k < xspan.size(); ++k) {
            const uint64_t byte = 256ull * k;
            while (i + 1 < n && xoff[i + 1] <= byte) ++i;
            xspan[k] = i;
        }
        rc = upload(c, c->d_xspan, xspan.data(), xspan.size());
        if (rc) return rc;
    }
    c->have_samples = true;
    c->layout_ok = false;
    return DNAF_OK;
}

int dnaf_set_snps(dnaf_ctx* c, uint64_t S, const uint8_t* cls, const uint8_t* k, const uint32_t* thr,
                  const uint8_t* prefix, const uint64_t* pre_off) {
    if (!c) return DNAF_E_ARG;
    if (S && (!cls || !k || !thr || !prefix || !pre_off)) return fail(c, DNAF_E_ARG, "NULL SNP array");
    CU(c, cudaSetDevice(c->dev));
    if (g_trace) { g_t0 = std::chrono::steady_clock::now(); trace("set_snps begins", 0); }
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
    c->h_k.assign(k, k + S);
    c->h_plen.resize(S);
    c->h_pfx_tab.resize(S);
    for (uint64_t r = 0; r < S; ++r) {
        c->h_plen[r] = (uint32_t)(pre_off[r + 1] - pre_off[r]);
        c->h_pfx_tab[r] = c->h_plen[r] && prefix[pre_off[r + 1] - 1] == '\t';
    }
    trace("validated, host copies made", 0);
    // pageable sources: each copy returns once the data is staged, one synchronise covers them all
    int rc = upload(c, c->d_cls, cls, S, false);
    if (!rc) rc = upload(c, c->d_k, k, S, false);
    if (!rc) rc = upload(c, c->d_thr, thr, S * 4, false);
    if (!rc) rc = upload(c, c->d_prefix, prefix, S ? pre_off[S] : 0, false);
    if (!rc) rc = upload(c, c->d_pre_off, pre_off, S + 1, S == 0);
    if (rc) return rc;
    if (S == 0) {
        const uint64_t zero = 0;
        rc = upload(c, c->d_pre_off, &zero, 1);
        if (rc) return rc;
    } else {   // per-row linear CRC of the prefix (k_auto adds it to its blocks' checksums with four lookups)
        CU(c, c->d_pre_crc.reserve(S * sizeof(uint32_t)));
        CU(c, c->d_pfx_state.reserve(16 * sizeof(uint32_t)));
        CU(c, c->h_present.reserve(8 * sizeof(uint32_t)));
        CU(c, cudaMemsetAsync(c->d_pfx_state.p, 0, 16 * sizeof(uint32_t), c->stream));
        k_prefix_crc<<<(uint32_t)((S + 255) / 256), 256, 0, c->stream>>>(c->d_prefix.as<uint8_t>(), c->d_pre_off.as<uint64_t>(), S,
                                                                       c->d_crctab.as<uint32_t>(), c->d_pre_crc.as<uint32_t>(),
                                                                       c->d_pfx_state.as<uint32_t>(), c->h_present.as<uint32_t>());
        CU(c, cudaGetLastError());
        CU(c, cudaStreamSynchronize(c->stream));   // the uploads above and the kernel
        uint32_t present[8];
        for (int w = 0; w < 8; ++w) present[w] = c->present_sticky[w] |= reinterpret_cast<volatile uint32_t*>(c->h_present.p)[w];
        prefix_model(c, present);
    }
    trace("uploaded", 0);
    c->have_snps = true;
    c->layout_ok = false;
    rc = prepare_buckets(c, k, thr);
    if (rc) return rc;
    trace("buckets prepared", 0);
    return DNAF_OK;
}

int dnaf_set_overrides(dnaf_ctx* c, uint64_t P, const uint64_t* rows, const uint32_t* samples) {
    if (!c) return DNAF_E_ARG;
    if (P && (!rows || !samples)) return fail(c, DNAF_E_ARG, "NULL override array");
    for (uint64_t i = 1; i < P; ++i)
        if (rows[i] < rows[i - 1]) return fail(c, DNAF_E_ARG, "override pairs must be sorted by row");
    CU(c, cudaSetDevice(c->dev));
    c->h_orow.assign(rows, rows + P);
    c->h_osamp.assign(samples, samples + P);
    int rc = upload(c, c->d_orow, rows, P);
    if (!rc) rc = upload(c, c->d_osamp, samples, P);
    if (rc) return rc;
    c->P = P;
    c->layout_ok = false;   // the per-row override index is part of the layout
    return DNAF_OK;
}

int dnaf_select_snps(dnaf_ctx* c, uint64_t n, uint64_t seed, uint32_t n_chrom, const double* chrom_cdf,
                     const double* chrom_max_pos, const uint8_t* chrom_rank, uint32_t n_maf, const double* maf_cdf,
                     int sorted, uint32_t* order, uint8_t* chrom_idx, uint8_t* maf_bin, uint32_t* position, uint8_t* ref,
                     uint8_t* alt) {
    if (!c) return DNAF_E_ARG;
    if (n == 0) return DNAF_OK;
    if (n > 0xFFFFFFFFull) return fail(c, DNAF_E_ARG, "at most 2^32-1 SNPs per call");
    if (!chrom_cdf || !chrom_max_pos || !chrom_rank || !maf_cdf || !order || !chrom_idx || !maf_bin || !position || !ref || !alt)
        return fail(c, DNAF_E_ARG, "NULL array");
    if (n_chrom < 1 || n_chrom > (uint32_t)kSelMaxChrom || n_maf < 1 || n_maf > (uint32_t)kSelMaxMaf)
        return fail(c, DNAF_E_ARG, "1..%d chromosomes and 1..%d MAF bins", kSelMaxChrom, kSelMaxMaf);
    for (uint32_t i = 0; i < n_chrom; ++i)
        if (!(chrom_max_pos[i] >= 0.0) || chrom_max_pos[i] >= 4294967296.0) return fail(c, DNAF_E_ARG, "chromosome length out of range");
    CU(c, cudaSetDevice(c->dev));
    DevBuf d_par, d_key, d_key2, d_idx, d_idx2, d_col, d_col2, d_tmp;
    const size_t par_bytes = (2 * (size_t)n_chrom + n_maf) * sizeof(double) + n_chrom;
    std::vector<uint8_t> par(par_bytes);
    memcpy(par.data(), chrom_cdf, n_chrom * sizeof(double));
    memcpy(par.data() + n_chrom * sizeof(double), chrom_max_pos, n_chrom * sizeof(double));
    memcpy(par.data() + 2 * n_chrom * sizeof(double), maf_cdf, n_maf * sizeof(double));
    memcpy(par.data() + (2 * (size_t)n_chrom + n_maf) * sizeof(double), chrom_rank, n_chrom);
    int rc = upload(c, d_par, par.data(), par.size(), false);
    if (rc) return rc;
    const size_t col_bytes = n * 8 + 64;   // chrom, maf, ref, alt (1 byte each) + pos (4 bytes), 16-byte aligned pieces
    auto col_at = [&](DevBuf& b, int which) {   // 0 pos, 1 chrom, 2 maf, 3 ref, 4 alt
        uint8_t* p = b.as<uint8_t>();
        const size_t n4 = (n * 4 + 15) & ~size_t(15), n1 = (n + 15) & ~size_t(15);
        return which == 0 ? p : p + n4 + (size_t)(which - 1) * n1;
    };
    CU(c, d_key.reserve(n * 8));
    CU(c, d_key2.reserve(n * 8));
    CU(c, d_idx.reserve(n * 4));
    CU(c, d_idx2.reserve(n * 4));
    CU(c, d_col.reserve(col_bytes + 64));
    CU(c, d_col2.reserve(col_bytes + 64));
    SelectArgs a;
    a.n = n;
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.n_chrom = n_chrom;
    a.n_maf = n_maf;
    a.chrom_cdf = d_par.as<double>();
    a.chrom_max_pos = d_par.as<double>() + n_chrom;
    a.maf_cdf = d_par.as<double>() + 2 * n_chrom;
    a.chrom_rank = d_par.as<uint8_t>() + (2 * (size_t)n_chrom + n_maf) * sizeof(double);
    a.key = d_key.as<uint64_t>();
    a.idx = d_idx.as<uint32_t>();
    a.pos = reinterpret_cast<uint32_t*>(col_at(d_col, 0));
    a.chrom = col_at(d_col, 1);
    a.maf = col_at(d_col, 2);
    a.ref = col_at(d_col, 3);
    a.alt = col_at(d_col, 4);
    k_select_snps<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(a);
    CU(c, cudaGetLastError());
    const DevBuf* res = &d_col;
    const uint32_t* d_order = d_idx.as<uint32_t>();
    if (sorted) {
        // stable LSD radix sort on (string rank of the chromosome, position): ties keep draw order, like list.sort
        size_t tmp = 0;
        CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_key.as<uint64_t>(), d_key2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                              d_idx2.as<uint32_t>(), (int)n, 0, 40, c->stream));
        CU(c, d_tmp.reserve(tmp + 16));
        CU(c, cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp, d_key.as<uint64_t>(), d_key2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                              d_idx2.as<uint32_t>(), (int)n, 0, 40, c->stream));
        k_select_gather<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(
            n, d_idx2.as<uint32_t>(), a.chrom, a.maf, a.pos, a.ref, a.alt, col_at(d_col2, 1), col_at(d_col2, 2),
            reinterpret_cast<uint32_t*>(col_at(d_col2, 0)), col_at(d_col2, 3), col_at(d_col2, 4));
        CU(c, cudaGetLastError());
        res = &d_col2;
        d_order = d_idx2.as<uint32_t>();
    }
    DevBuf& R = const_cast<DevBuf&>(*res);
    CU(c, cudaMemcpyAsync(order, d_order, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(position, col_at(R, 0), n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(chrom_idx, col_at(R, 1), n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(maf_bin, col_at(R, 2), n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(ref, col_at(R, 3), n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(alt, col_at(R, 4), n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return DNAF_OK;
}

int64_t dnaf_parse_snps_jsonl(const char* text, uint64_t n_bytes, uint64_t cap, int64_t* ids, int32_t* chrom_idx,
                              int64_t* position, uint8_t* n_alleles, uint8_t* nts, double* cum, char* chrom_labels,
                              uint32_t max_labels, uint32_t* n_labels) {
    if (!text || !ids || !chrom_idx || !position || !n_alleles || !nts || !cum || !chrom_labels || !n_labels) return 0;
    return snpsjson::parse(text, n_bytes, cap, ids, chrom_idx, position, n_alleles, nts, cum, chrom_labels, max_labels, n_labels);
}

uint64_t dnaf_format_prefixes(uint64_t n, const int32_t* chrom_idx, const char* labels, const int64_t* position,
                              const int64_t* ids, const uint8_t* n_alleles, const uint8_t* nts, char* out, uint64_t* off) {
    return snpsfmt::prefixes(n, chrom_idx, labels, position, ids, n_alleles, nts, out, off);
}

uint64_t dnaf_format_snps_jsonl(uint64_t n, const int32_t* chrom_idx, const char* labels, const int64_t* position,
                                const int64_t* ids, const uint8_t* n_alleles, const uint8_t* nts, const uint32_t* repr_idx,
                                const char* reprs, const uint32_t* repr_off, char* out) {
    return snpsfmt::jsonl(n, chrom_idx, labels, position, ids, n_alleles, nts, repr_idx, reprs, repr_off, out);
}

uint64_t dnaf_bgzf_bound(uint64_t text_bytes) {
    // every block carries <= kBlk bytes of text and at most 26 + 5 bytes of framing beyond them;
    // row-aligned cutting can leave blocks partly filled, so count blocks generously
    const uint64_t blocks = text_bytes / (kBlk / 2) + 2;
    return text_bytes + blocks * 64 + 1024;
}

int dnaf_block_log(dnaf_ctx* c, int enable) {
    if (!c) return DNAF_E_ARG;
    c->log_blocks = enable != 0;
    c->log_csize.clear();
    c->log_usize.clear();
    return DNAF_OK;
}

int dnaf_block_log_get(dnaf_ctx* c, const uint32_t** csize, const uint32_t** usize, uint64_t* n_blocks) {
    if (!c) return DNAF_E_ARG;
    if (!csize || !usize || !n_blocks) return fail(c, DNAF_E_ARG, "NULL output pointer");
    *csize = c->log_csize.data();
    *usize = c->log_usize.data();
    *n_blocks = c->log_csize.size();
    return DNAF_OK;
}

int dnaf_bgzf_scan(const uint8_t* data, uint64_t n_bytes, uint32_t* csize, uint32_t* usize, uint64_t cap,
                   uint64_t* n_blocks) {
    if (!n_blocks || (n_bytes && !data)) return DNAF_E_ARG;
    uint64_t k = 0;
    const uint64_t covered = walk_bgzf(data, n_bytes, [&](uint32_t cs, uint32_t us) {
        if (k < cap) {
            if (csize) csize[k] = cs;
            if (usize) usize[k] = us;
        }
        ++k;
    });
    *n_blocks = k;
    if (covered != n_bytes) return DNAF_E_INPUT;
    return k > cap && (csize || usize) ? DNAF_E_SPACE : DNAF_OK;
}

int64_t dnaf_debug_lz_block(double p_minor, int level, const uint32_t* allele_bits, uint32_t n_cells, const uint8_t* prefix,
                            uint32_t prefix_len, int ends_row, uint8_t* out, uint64_t out_cap) {
    if (level < 4 || level > 9 || !allele_bits || !out || n_cells == 0 || n_cells > 254u * 64u || prefix_len > 64 ||
        (prefix_len && !prefix))
        return DNAF_E_ARG;
    uint64_t hist[256] = {0};
    for (uint32_t i = 0; i < prefix_len; ++i) hist[prefix[i]] += 16;
    const int per_block = (int)((n_cells + 63u) / 64u);
    const LzTable t = hosttab::make_lz_table(p_minor, prefix_len ? hist : nullptr, per_block, prefix_len != 0, level);
    if (t.hdr_bits == 0xFFFFFFFFu) return DNAF_E_INPUT;
    const std::vector<uint8_t> enc = hosttab::lz_encode_block_host(t, allele_bits, n_cells, prefix, prefix_len, ends_row != 0, level);
    if (enc.size() > out_cap) return DNAF_E_SPACE;
    memcpy(out, enc.data(), enc.size());
    return (int64_t)enc.size();
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

int dnaf_row_offsets(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t* out) {
    if (!c) return DNAF_E_ARG;
    int rc = ensure_layout(c);
    if (rc) return rc;
    if (row_begin > row_end || row_end > c->S) return fail(c, DNAF_E_ARG, "row range out of bounds");
    if (!out) return fail(c, DNAF_E_ARG, "out is NULL");
    const uint64_t base = c->h_row_off[row_begin];
    for (uint64_t r = row_begin; r <= row_end; ++r) out[r - row_begin] = c->h_row_off[r] - base;
    return DNAF_OK;
}

int dnaf_generate(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                  uint8_t* out, uint64_t out_cap, dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (!out && out_cap) return fail(c, DNAF_E_ARG, "out is NULL");
    Sink s;
    s.buf = out;
    s.cap = out_cap;
    return generate_impl(c, row_begin, row_end, seed, level, s, stats);
}

int dnaf_generate_stream(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                         dnaf_sink_fn sink, void* user, dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (!sink) return fail(c, DNAF_E_ARG, "sink is NULL");
    Sink s;
    s.fn = sink;
    s.user = user;
    return generate_impl(c, row_begin, row_end, seed, level, s, stats);
}

int dnaf_generate_fd(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level, int fd,
                     dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (fd < 0) return fail(c, DNAF_E_ARG, "bad file descriptor");
    Sink s;
    s.fd = fd;
    return generate_impl(c, row_begin, row_end, seed, level, s, stats);
}

int dnaf_generate_fd_at(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level, int fd, uint64_t file_offset,
                        dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    if (fd < 0) return fail(c, DNAF_E_ARG, "bad file descriptor");
    if (file_offset > (uint64_t)INT64_MAX) return fail(c, DNAF_E_ARG, "file offset out of range");
    Sink s;
    s.fd = fd;
    s.fd_off = (int64_t)file_offset;
    return generate_impl(c, row_begin, row_end, seed, level, s, stats);
}

int dnaf_generate_device(dnaf_ctx* c, uint64_t row_begin, uint64_t row_end, uint64_t seed, int level,
                         dnaf_stats* stats) {
    if (!c) return DNAF_E_ARG;
    Sink s;
    s.device_only = true;
    return generate_impl(c, row_begin, row_end, seed, level, s, stats);
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
        uint64_t n_over = 0;
        stage_overrides_all(c, r0, r1, &n_over);
        rc = upload_overrides(c);
        if (!rc) rc = run_sample(c, r0, (uint32_t)(r1 - r0), nullptr, seed, n_over, c->d_olocal.as<uint32_t>(),
                                 c->d_osub.as<uint32_t>(), nullptr);
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
        const uint64_t bytes = c->h_row_off[r1] - c->h_row_off[r0];
        uint64_t n_over = 0;
        stage_overrides_all(c, r0, r1, &n_over);
        rc = upload_overrides(c);
        if (!rc) rc = run_sample(c, r0, (uint32_t)(r1 - r0), nullptr, seed, n_over, c->d_olocal.as<uint32_t>(),
                                 c->d_osub.as<uint32_t>(), nullptr);
        if (!rc) rc = run_format(c, r0, (uint32_t)(r1 - r0), nullptr, nullptr, bytes, nullptr);
        if (rc) return rc;
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
        c->gslot.clear();
        for (uint64_t o = 0; o < piece; o += kBlk)
            c->plan.push_back({o, (uint32_t)std::min<uint64_t>(kBlk, piece - o), 0});
        dnaf_ctx::OutBuf& B = c->ob[0];
        c->cur_ob = 0;
        c->slot_stride = kSlot;
        int rc = reserve_outputs(c, B, (uint32_t)c->plan.size(), piece);
        if (!rc) rc = reserve_stage(c, B);
        if (rc) return rc;
        CU(c, cudaStreamWaitEvent(c->stream, c->sbuf[c->sb].ev_free, 0));
        for (int e = 0; e < 3; ++e) CU(c, cudaEventRecord(B.ev[e], c->stream));
        B.rows = 0;
        B.text = 0;
        B.gen = false;
        B.fused = false;
        B.generic_blocks = true;
        B.auto_text = 0;
        rc = launch_generic(c, &local);
        if (rc) return rc;
        for (int e = 3; e < 5; ++e) CU(c, cudaEventRecord(B.ev[e], c->stream));   // ev[4]: the compaction stream waits for it
        rc = close_pass(c, B, (uint32_t)c->plan.size(), &local);
        if (!rc) rc = start_copy(c, B, s, &local);
        if (!rc) rc = finish_copy(c, B, s);
        if (rc) return rc;
        local.text_bytes += piece;
        done += piece;
    }
    local.ms_total = local.ms_deflate;
    if (stats) *stats = local;
    return DNAF_OK;
}

}  // extern "C"
