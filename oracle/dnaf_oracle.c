/*
 * dnaf_oracle.c -- CPU restatement of the dna-factory population-generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the timed CPU baseline.  The product path (dna_factory_b200)
 * never imports, links or executes this file.
 *
 * What it restates (file:line in the upstream ochrzan/dna-factory tree):
 *   - SNPTuples.pick_allele_index            pop_factory.py:92-95   (first i with cum[i] >= u, float64)
 *   - is_haploid                             common/snp.py:102-109  ((X and male) or MT or Y)
 *   - PopulationFactory.queue_vcf_snps       pop_factory.py:471-513 (per-SNP / per-sample row loop + row text)
 *   - SNPTuples.alt_alleles / ref_allele     pop_factory.py:104-116 (ALT column) -- done by the Python wrapper
 *   - Bio.bgzf.BgzfWriter (third party, NOT in the reference tree; biopython, unpinned in
 *     requirements.txt:4): restated from its published behaviour -- cut the text every 65536
 *     bytes, raw deflate (zlib level z, wbits -15, memLevel 8), 18-byte BGZF header, CRC32,
 *     ISIZE, 28-byte EOF block.  Call sites pop_factory.py:13,403,405,449,458.
 *
 * Parity pinning: this file is checked (tests/test_oracle_golden.py) against rows produced by
 * the UNMODIFIED reference run in the build container (tests/golden/make_golden.py imports
 * /root/reference/pop_factory.py and calls PopulationFactory.queue_vcf_snps with
 * numpy.random.rand patched to the Philox stream below), and against the reference's own
 * known-answer test for pick_allele_index (test/unit/pop_factory_test.py:24-28).
 *
 * The uniform stream ("replay RNG spec"):
 *   W(row, g, b) = philox4x32_10(ctr = (g, b>>2, row_lo, row_hi), key = (seed_lo, seed_hi))[b & 3]
 *   U_j (32-bit) : bit (31-b) of U_j = bit (j & 31) of W(row, j>>5, b),  b = 0..31
 *   u_j = U_j * 2^-32   -- what numpy.random.rand(2N)[j] returns for that SNP row
 * The oracle always assembles all 32 bits (no lazy evaluation) and compares in float64 like
 * the reference does; the CUDA path uses lazily evaluated bit-sliced integer compares.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DNAF_OR_KMAX 4

/* ---------------------------------------------------------------- Philox4x32-10 */
void dnaf_or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 32-bit integer uniforms U_j for one SNP row, j = 0 .. n_alleles-1 (n_alleles = 2N). */
void dnaf_or_uniform_bits(uint64_t seed, uint64_t row, uint32_t n_alleles, uint32_t* U) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t groups = (n_alleles + 31u) / 32u;
    for (uint32_t g = 0; g < groups; ++g) {
        uint32_t W[32];
        for (uint32_t q = 0; q < 8; ++q) {
            uint32_t ctr[4] = {g, q, (uint32_t)row, (uint32_t)(row >> 32)};
            dnaf_or_philox4x32_10(ctr, key, &W[4 * q]);
        }
        for (uint32_t l = 0; l < 32; ++l) {
            uint32_t j = g * 32u + l;
            if (j >= n_alleles) break;
            uint32_t u = 0;
            for (uint32_t b = 0; b < 32; ++b) u |= ((W[b] >> l) & 1u) << (31u - b);
            U[j] = u;
        }
    }
}

/* float64 uniforms, exactly what the patched numpy.random.rand(2N) hands the reference. */
void dnaf_or_uniforms(uint64_t seed, uint64_t row, uint32_t n_alleles, double* u) {
    uint32_t* U = (uint32_t*)malloc(sizeof(uint32_t) * (n_alleles ? n_alleles : 1));
    dnaf_or_uniform_bits(seed, row, n_alleles, U);
    for (uint32_t j = 0; j < n_alleles; ++j) u[j] = (double)U[j] * (1.0 / 4294967296.0);
    free(U);
}

/* pop_factory.py:92-95 -- first index whose cumulative probability is >= the roll; -1 = None. */
int dnaf_or_pick_allele_index(const double* cum, int k, double roll) {
    for (int i = 0; i < k; ++i)
        if (cum[i] >= roll) return i;
    return -1;
}

/* common/snp.py:102-109 */
static int is_haploid(const char* chromo, int is_male) {
    return (strcmp(chromo, "X") == 0 && is_male) || strcmp(chromo, "MT") == 0 || strcmp(chromo, "Y") == 0;
}

/*
 * One VCF data row, pop_factory.py:474-508.  `prefix` is the already formatted 9-column lead
 * ("%s\t%i\trs%s\t%s\t%s\t40\tPASS\t.\tGT\t", built by the wrapper exactly as :503-507).
 * `is_del[i]` != 0 marks the samples for which `snp.id in sample.deleterious_snps` holds
 * (computed by the caller with Python dict semantics so that quirk R8 is preserved).
 * Returns bytes written, or -1 when pick_allele_index would return None (reference raises).
 */
int64_t dnaf_or_row(const char* chromo, const uint8_t* prefix, uint32_t prefix_len,
                    const double* cum, int k, uint32_t n_samples, const uint8_t* sex,
                    const uint8_t* is_control, const uint8_t* is_del, const double* randoms,
                    uint8_t* out) {
    uint8_t* p = out;
    memcpy(p, prefix, prefix_len);
    p += prefix_len;
    int chrom_is_y = strcmp(chromo, "Y") == 0;
    for (uint32_t i = 0; i < n_samples; ++i) {
        int is_male = sex[i] == 1;                       /* pop_factory.py:70-71 */
        if (i) *p++ = '\t';                              /* "\t".join(sample_values) */
        if (!is_male && chrom_is_y) {                    /* :481-484 */
            *p++ = '.';
            continue;
        }
        if (is_control[i] || !(is_del && is_del[i])) {   /* :485 */
            int a = dnaf_or_pick_allele_index(cum, k, randoms[2 * i]);
            if (a < 0) return -1;
            if (is_haploid(chromo, is_male)) {           /* :488-490 */
                p += sprintf((char*)p, "%i", a);
                continue;
            }
            int b = dnaf_or_pick_allele_index(cum, k, randoms[2 * i + 1]);
            if (b < 0) return -1;
            p += sprintf((char*)p, "%i/%i", a, b);       /* :494 */
        } else {
            if (is_haploid(chromo, is_male)) *p++ = '1'; /* :496-499 */
            else { *p++ = '1'; *p++ = '/'; *p++ = '1'; }
        }
    }
    *p++ = '\n';
    return (int64_t)(p - out);
}

/*
 * Rows [row_begin, row_begin + n_rows) of a population (global sorted row index = Philox row
 * counter).  Arrays are indexed by local row 0..n_rows-1.  Override pairs (local row, sample)
 * must be sorted by row.  `row_off` (n_rows+1) receives the byte offset of every row.
 * Returns total bytes, -1 on a None pick, -2 if `cap` is too small.
 */
int64_t dnaf_or_rows(uint64_t seed, uint64_t row_begin, uint64_t n_rows,
                     const uint8_t* chrom_bytes, const uint64_t* chrom_off,
                     const uint8_t* prefix_bytes, const uint64_t* prefix_off,
                     const uint8_t* kk, const double* cum /* [n_rows][KMAX] */,
                     uint32_t n_samples, const uint8_t* sex, const uint8_t* is_control,
                     uint64_t n_over, const uint64_t* over_row, const uint32_t* over_sample,
                     uint8_t* out, uint64_t cap, uint64_t* row_off, int n_threads) {
    /* worst-case row length: prefix + 4 bytes per sample; exact lengths need the sex vector */
    uint64_t* lens = (uint64_t*)calloc(n_rows + 1, sizeof(uint64_t));
    uint32_t males = 0;
    for (uint32_t i = 0; i < n_samples; ++i) males += sex[i] == 1;
    for (uint64_t r = 0; r < n_rows; ++r) {
        char chromo[32];
        uint64_t cl = chrom_off[r + 1] - chrom_off[r];
        if (cl > 31) cl = 31;
        memcpy(chromo, chrom_bytes + chrom_off[r], cl);
        chromo[cl] = 0;
        uint64_t body;
        if (!strcmp(chromo, "Y") || !strcmp(chromo, "MT")) body = 2ull * n_samples;
        else if (!strcmp(chromo, "X")) body = 2ull * males + 4ull * (n_samples - males);
        else body = 4ull * n_samples;
        if (n_samples == 0) body = 1; /* just the newline */
        lens[r] = (prefix_off[r + 1] - prefix_off[r]) + body;
    }
    uint64_t tot = 0;
    for (uint64_t r = 0; r < n_rows; ++r) { row_off[r] = tot; tot += lens[r]; }
    row_off[n_rows] = tot;
    free(lens);
    if (tot > cap) return -2;
    /* first override index per row */
    uint64_t* ofirst = (uint64_t*)malloc(sizeof(uint64_t) * (n_rows + 1));
    {
        uint64_t o = 0;
        for (uint64_t r = 0; r <= n_rows; ++r) {
            while (o < n_over && over_row[o] < r) ++o;
            ofirst[r] = o;
        }
    }
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
    {
        double* u = (double*)malloc(sizeof(double) * (2ull * n_samples + 1));
        uint8_t* del = (uint8_t*)malloc(n_samples + 1);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t r = 0; r < (int64_t)n_rows; ++r) {
            char chromo[32];
            uint64_t cl = chrom_off[r + 1] - chrom_off[r];
            if (cl > 31) cl = 31;
            memcpy(chromo, chrom_bytes + chrom_off[r], cl);
            chromo[cl] = 0;
            dnaf_or_uniforms(seed, row_begin + (uint64_t)r, 2u * n_samples, u); /* :477 */
            const uint8_t* delp = NULL;
            if (ofirst[r + 1] > ofirst[r]) {
                memset(del, 0, n_samples);
                for (uint64_t o = ofirst[r]; o < ofirst[r + 1]; ++o)
                    if (over_sample[o] < n_samples) del[over_sample[o]] = 1;
                delp = del;
            }
            int64_t n = dnaf_or_row(chromo, prefix_bytes + prefix_off[r],
                                    (uint32_t)(prefix_off[r + 1] - prefix_off[r]),
                                    cum + (uint64_t)r * DNAF_OR_KMAX, kk[r], n_samples, sex,
                                    is_control, delp, u, out + row_off[r]);
            if (n < 0 || (uint64_t)n != row_off[r + 1] - row_off[r]) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                bad = 1;
            }
        }
        free(u);
        free(del);
    }
    free(ofirst);
    return bad ? -1 : (int64_t)tot;
}

/* ---------------------------------------------------------------- BGZF (Biopython BgzfWriter) */
static const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43,
                                     0x02, 0x00, 0x1b, 0x00, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* One BGZF block from `n` (<= 65536) bytes; returns block size or -1 (did not fit in 64 KiB). */
static int64_t bgzf_block(const uint8_t* in, uint32_t n, int level, uint8_t* out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8 /* DEF_MEM_LEVEL */, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = (Bytef*)in;
    zs.avail_in = n;
    zs.next_out = out + 18;
    zs.avail_out = 65536 - 26;
    int rc = deflate(&zs, Z_FINISH);
    uint32_t clen = (uint32_t)zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    static const uint8_t head[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00};
    memcpy(out, head, 16);
    uint32_t bsize = clen + 25;
    out[16] = (uint8_t)bsize;
    out[17] = (uint8_t)(bsize >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, n);
    uint8_t* t = out + 18 + clen;
    for (int i = 0; i < 4; ++i) t[i] = (uint8_t)(crc >> (8 * i));
    for (int i = 0; i < 4; ++i) t[4 + i] = (uint8_t)(n >> (8 * i));
    return (int64_t)clen + 26;
}

/* Worst-case output size for dnaf_or_bgzf. */
uint64_t dnaf_or_bgzf_bound(uint64_t n) { return (n / 65536 + 2) * 65536ull + 28; }

/*
 * BgzfWriter.write()/close(): cut every 65536 input bytes, compress each piece independently,
 * append the EOF marker when `with_eof`.  Blocks are compressed in parallel (the reference's
 * writer is a single thread; this port is deliberately generous to the CPU side).
 */
int64_t dnaf_or_bgzf(const uint8_t* text, uint64_t n, int level, int with_eof, uint8_t* out,
                     uint64_t cap, int n_threads) {
    uint64_t nblk = (n + 65535) / 65536;
    int64_t* sizes = (int64_t*)malloc(sizeof(int64_t) * (nblk + 1));
    uint8_t* scratch = (uint8_t*)malloc((nblk ? nblk : 1) * 65536ull);
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < (int64_t)nblk; ++b) {
        uint64_t o = (uint64_t)b * 65536;
        uint32_t len = (uint32_t)((n - o) < 65536 ? (n - o) : 65536);
        sizes[b] = bgzf_block(text + o, len, level, scratch + (uint64_t)b * 65536);
        if (sizes[b] < 0) bad = 1;
    }
    int64_t tot = 0;
    if (!bad) {
        for (uint64_t b = 0; b < nblk; ++b) {
            if ((uint64_t)tot + (uint64_t)sizes[b] > cap) { bad = 1; break; }
            memcpy(out + tot, scratch + b * 65536, (size_t)sizes[b]);
            tot += sizes[b];
        }
        if (!bad && with_eof) {
            if ((uint64_t)tot + 28 > cap) bad = 1;
            else { memcpy(out + tot, BGZF_EOF, 28); tot += 28; }
        }
    }
    free(sizes);
    free(scratch);
    return bad ? -1 : tot;
}

uint32_t dnaf_or_crc32(const uint8_t* p, uint64_t n) { return (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n); }

int dnaf_or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
