// Fast ingest of snps.json(.gz, already inflated) -- the file PopulationFactory.output_snps writes and
// load_snps_file reads back (pop_factory.py:118-133,258-272): one JSON object per line,
//   {"id": 329, "chromosome": "1", "position": 1798996, "tuples": {"T": 0.98, "A": 1.0}}
// straight into the column form the device path consumes, without building a Python object per SNP.
// Host code only.  Records the column form cannot hold (string ids, multi-character alleles, more than DNAF_KMAX
// alleles, escapes in strings) make the parser stop with the line number; the caller then falls back to json.loads.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace dnaf {
namespace snpsjson {

struct Cursor {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    // "key" without escapes -> [s, s+n)
    bool str(const char*& s, size_t& n) {
        ws();
        if (p >= end || *p != '"') return false;
        s = ++p;
        while (p < end && *p != '"' && *p != '\\' && *p != '\n') ++p;
        if (p >= end || *p != '"') return false;
        n = (size_t)(p - s);
        ++p;
        return true;
    }
    bool integer(int64_t& v) {
        ws();
        const char* s = p;
        bool neg = false;
        if (p < end && *p == '-') { neg = true; ++p; }
        if (p >= end || *p < '0' || *p > '9') { p = s; return false; }
        uint64_t a = 0;
        int digits = 0;
        while (p < end && *p >= '0' && *p <= '9') { a = a * 10 + (uint64_t)(*p - '0'); ++p; if (++digits > 18) return false; }
        if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) { p = s; return false; }
        v = neg ? -(int64_t)a : (int64_t)a;
        return true;
    }
    bool number(double& v) {   // strtod parses exactly like Python's float() for JSON numbers
        ws();
        char buf[64];
        size_t n = 0;
        while (p + n < end && n < 63 && (strchr("+-.eE", p[n]) || (p[n] >= '0' && p[n] <= '9'))) { buf[n] = p[n]; ++n; }
        if (!n) return false;
        buf[n] = 0;
        char* e = nullptr;
        v = strtod(buf, &e);
        if (e != buf + n) return false;
        p += n;
        return true;
    }
};

// Returns the number of records parsed, or -(line number) of the first line it cannot hold in column form.
// chrom_labels: up to max_labels distinct labels of at most 7 characters, 8 bytes each (NUL padded), in order of
// first appearance; chrom_idx[r] indexes them.
inline int64_t parse(const char* text, size_t n_bytes, uint64_t cap, int64_t* ids, int32_t* chrom_idx, int64_t* position,
                     uint8_t* n_alleles, uint8_t* nts /*[cap][4]*/, double* cum /*[cap][4]*/, char* chrom_labels,
                     uint32_t max_labels, uint32_t* n_labels) {
    Cursor c{text, text + n_bytes};
    uint64_t r = 0;
    int64_t line = 0;
    uint32_t nl = 0;
    while (c.p < c.end) {
        ++line;
        c.ws();
        if (c.p < c.end && *c.p == '\n') { ++c.p; continue; }
        if (c.p >= c.end) break;
        if (r >= cap) return -line;
        if (!c.eat('{')) return -line;
        bool have_id = false, have_chr = false, have_pos = false;
        uint8_t k = 0;
        for (int q = 0; q < 4; ++q) { nts[4 * r + q] = 0; cum[4 * r + q] = 2.0; }
        for (;;) {
            const char* key; size_t kn;
            if (!c.str(key, kn) || !c.eat(':')) return -line;
            if (kn == 2 && !memcmp(key, "id", 2)) {
                if (!c.integer(ids[r])) return -line;
                have_id = true;
            } else if (kn == 10 && !memcmp(key, "chromosome", 10)) {
                const char* s; size_t sn;
                if (!c.str(s, sn) || sn == 0 || sn > 7) return -line;
                uint32_t i = 0;
                for (; i < nl; ++i)
                    if (!strncmp(chrom_labels + 8 * i, s, sn) && chrom_labels[8 * i + sn] == 0) break;
                if (i == nl) {
                    if (nl >= max_labels) return -line;
                    memset(chrom_labels + 8 * nl, 0, 8);
                    memcpy(chrom_labels + 8 * nl, s, sn);
                    ++nl;
                }
                chrom_idx[r] = (int32_t)i;
                have_chr = true;
            } else if (kn == 8 && !memcmp(key, "position", 8)) {
                if (!c.integer(position[r])) return -line;
                have_pos = true;
            } else if (kn == 6 && !memcmp(key, "tuples", 6)) {
                if (!c.eat('{')) return -line;
                if (!c.eat('}')) {
                    for (;;) {
                        const char* s; size_t sn;
                        double v;
                        if (!c.str(s, sn) || sn != 1 || !c.eat(':') || !c.number(v) || k >= 4) return -line;
                        nts[4 * r + k] = (uint8_t)s[0];
                        cum[4 * r + k] = v;
                        ++k;
                        if (c.eat(',')) continue;
                        if (c.eat('}')) break;
                        return -line;
                    }
                }
            } else {
                return -line;   // unknown key: let json.loads decide
            }
            if (c.eat(',')) continue;
            if (c.eat('}')) break;
            return -line;
        }
        if (!have_id || !have_chr || !have_pos || k < 1) return -line;
        n_alleles[r] = k;
        c.ws();
        if (c.p < c.end && *c.p != '\n') return -line;
        if (c.p < c.end) ++c.p;
        ++r;
    }
    *n_labels = nl;
    return (int64_t)r;
}

}  // namespace snpsjson
}  // namespace dnaf

// ---- emitters (host code): the row prefixes of pop_factory.py:503-507 and the lines of SNPTuples.__str__ ----
namespace dnaf {
namespace snpsfmt {

inline char* put_uint(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char* put_str(char* p, const char* s) {
    while (*s) *p++ = *s++;
    return p;
}

// "%s\t%i\trs%s\t%s\t%s\t40\tPASS\t.\tGT\t" % (chromosome, position, id, ref, alt_alleles())  -- pop_factory.py:503-507
// with alt_alleles() of pop_factory.py:111-116: tuples[1] for K == 2, the REF itself for K == 1, comma-joined tuples[1:] else.
// labels: 8 bytes per chromosome label, NUL padded.  Returns bytes written; off[r] .. off[r+1] delimit row r.
inline uint64_t prefixes(uint64_t n, const int32_t* chrom_idx, const char* labels, const int64_t* position, const int64_t* ids,
                         const uint8_t* k, const uint8_t* nts, char* out, uint64_t* off) {
    char* p = out;
    for (uint64_t r = 0; r < n; ++r) {
        off[r] = (uint64_t)(p - out);
        p = put_str(p, labels + 8 * chrom_idx[r]);
        *p++ = '\t';
        p = put_uint(p, (uint64_t)position[r]);
        p = put_str(p, "\trs");
        p = put_uint(p, (uint64_t)ids[r]);
        *p++ = '\t';
        *p++ = (char)nts[4 * r];
        *p++ = '\t';
        if (k[r] == 1) *p++ = (char)nts[4 * r];
        else
            for (int j = 1; j < k[r]; ++j) {
                if (j > 1) *p++ = ',';
                *p++ = (char)nts[4 * r + j];
            }
        p = put_str(p, "\t40\tPASS\t.\tGT\t");
    }
    off[n] = (uint64_t)(p - out);
    return (uint64_t)(p - out);
}

// {"id": 329, "chromosome": "1", "position": 1798996, "tuples": {"T": 0.98, "A": 1.0}}\n   -- json.dumps of
// pop_factory.py:118-124.  Floats are printed by the caller (Python's repr): repr_idx[r*4+j] indexes the table of
// NUL-terminated strings `reprs` (repr_off[i] = start of string i).
inline uint64_t jsonl(uint64_t n, const int32_t* chrom_idx, const char* labels, const int64_t* position, const int64_t* ids,
                      const uint8_t* k, const uint8_t* nts, const uint32_t* repr_idx, const char* reprs, const uint32_t* repr_off,
                      char* out) {
    char* p = out;
    for (uint64_t r = 0; r < n; ++r) {
        p = put_str(p, "{\"id\": ");
        p = put_uint(p, (uint64_t)ids[r]);
        p = put_str(p, ", \"chromosome\": \"");
        p = put_str(p, labels + 8 * chrom_idx[r]);
        p = put_str(p, "\", \"position\": ");
        p = put_uint(p, (uint64_t)position[r]);
        if (k[r]) {
            p = put_str(p, ", \"tuples\": {");
            for (int j = 0; j < k[r]; ++j) {
                if (j) p = put_str(p, ", ");
                *p++ = '"';
                *p++ = (char)nts[4 * r + j];
                p = put_str(p, "\": ");
                p = put_str(p, reprs + repr_off[repr_idx[4 * r + j]]);
            }
            *p++ = '}';
        }
        *p++ = '}';
        *p++ = '\n';
    }
    return (uint64_t)(p - out);
}

}  // namespace snpsfmt
}  // namespace dnaf
