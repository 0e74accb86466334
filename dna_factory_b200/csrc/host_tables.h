// Host-side table builders: MAF buckets -> static Huffman tables (k_auto / k_x / k_fused_text / k_lz), CRC helper
// tables, autosome / X / text segments of a row and their template CRCs.  Included by dnaf_api.cu after host_ctx.h.
#pragma once

namespace {

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
        std::lock_guard<std::mutex> tables_lock(g_tables_mu);
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
                        key = {pbits, base + (uint64_t)per_block + c->samples_hash * 0x9E3779B97F4A7C15ull};
                        cached = c->xtable_cache.count(key) != 0;
                        family = 1;
                    } else {
                        const int cls = (v - 2) / 2;
                        key = {pbits, base + (uint64_t)(cls + 1) * 1000003ull + c->samples_hash * 0x9E3779B97F4A7C15ull};
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
                                                         c->samples_hash * 0x9E3779B97F4A7C15ull);
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
                                      (cls >= 0 ? c->samples_hash * 0x9E3779B97F4A7C15ull : 0);
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
    if (!c->fused_ok || level < 3 || c->h_seg_crc.empty()) return DNAF_OK;
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
    std::lock_guard<std::mutex> tables_lock(g_tables_mu);
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
            if (res[i].hdr_bits == 0xFFFFFFFFu || res[i].key_alleles > kLzMaxKey) return DNAF_OK;   // header too long: the call stays on k_auto
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

}  // namespace
