/* Deterministic, multi-threaded synthetic-data generators for the bam2db / freq hot paths.
 *
 *   - 10x-v3-style BAM (BGZF, zlib level-6 raw deflate, records never split across BGZF blocks,
 *     header in its own block, 28-byte EOF block) + barcodes.tsv + features.tsv
 *   - R1 FASTQ (BGZF-compressed so blocks inflate independently; lines DO straddle blocks)
 *
 * The record shape follows SURVEY.md section 8(d) ("Synthetic BAM" / "Synthetic FASTQ").  Every
 * read is a pure function of (seed, read index) through a counter-based hash, and the BGZF block
 * boundaries are a pure function of the fixed super-chunk size, so the bytes do not depend on the
 * number of worker threads.
 *
 * This is data tooling (bench + tests); it is not part of the reference's hot path.
 */
#define _GNU_SOURCE
#include "synth.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define BGZF_MAX_PAYLOAD 0xff00
#define SUPER_CHUNK_READS 32768

/* ---------- counter-based randomness ---------- */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { r->s += 0x9e3779b97f4a7c15ULL; uint64_t z = r->s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static const char BASES[4] = {'A', 'C', 'G', 'T'};

/* barcode k of the list: distinct random 16-mers, sorted ascending (so the file is sorted). We draw
 * n distinct 32-bit codes by stratifying the 2^32 space into n equal strata (sorted by construction). */
static void barcode_string(const fastf_synth_params *p, uint32_t k, char out[17])
{
    uint64_t stratum = ((uint64_t)1 << 32) / p->n_cells;
    uint64_t h = mix64(p->seed * 0x51ed270b7a3fULL + 0xbc0deULL + k);
    uint64_t code = (uint64_t)k * stratum + (h % stratum);
    for (int i = 0; i < 16; i++) out[i] = BASES[(code >> (2 * (15 - i))) & 3];
    out[16] = 0;
}

/* ---------- BGZF block writer ---------- */
typedef struct { uint8_t *p; size_t n, cap; } buf_t;
static void buf_reserve(buf_t *b, size_t extra)
{
    if (b->n + extra > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : (1 << 20);
        while (nc < b->n + extra) nc *= 2;
        b->p = (uint8_t *)realloc(b->p, nc);
        if (!b->p) { fprintf(stderr, "synth: out of memory\n"); abort(); }
        b->cap = nc;
    }
}
static void bgzf_emit_block(buf_t *out, const uint8_t *payload, unsigned len, int level, uint32_t *nblocks)
{
    buf_reserve(out, (size_t)len + 1024 + len / 1000 + 64);
    uint8_t *dst = out->p + out->n;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(dst, hdr, 16);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) abort();
    zs.next_in = (Bytef *)payload;
    zs.avail_in = len;
    zs.next_out = dst + 18;
    zs.avail_out = (uInt)(out->cap - out->n - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "synth: deflate failed\n"); abort(); }
    unsigned clen = (unsigned)zs.total_out;
    deflateEnd(&zs);
    unsigned bsize = 18 + clen + 8;
    if (bsize > 65536) { fprintf(stderr, "synth: BGZF block overflow\n"); abort(); }
    dst[16] = (uint8_t)((bsize - 1) & 0xff);
    dst[17] = (uint8_t)((bsize - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), payload, len);
    uint8_t *t = dst + 18 + clen;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = (crc >> 24) & 0xff;
    t[4] = len & 0xff; t[5] = (len >> 8) & 0xff; t[6] = (len >> 16) & 0xff; t[7] = (len >> 24) & 0xff;
    out->n += bsize;
    if (nblocks) (*nblocks)++;
}
static const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* ---------- BAM record synthesis ---------- */
static inline void put32(uint8_t *p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff; }
static inline void put16(uint8_t *p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; }
static int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
static inline uint8_t *aux_C(uint8_t *q, const char *tag, unsigned v) { q[0] = tag[0]; q[1] = tag[1]; q[2] = 'C'; q[3] = (uint8_t)v; return q + 4; }
static inline uint8_t *aux_A(uint8_t *q, const char *tag, char v) { q[0] = tag[0]; q[1] = tag[1]; q[2] = 'A'; q[3] = (uint8_t)v; return q + 4; }
static inline uint8_t *aux_Z(uint8_t *q, const char *tag, const char *s, size_t n) { q[0] = tag[0]; q[1] = tag[1]; q[2] = 'Z'; memcpy(q + 3, s, n); q[3 + n] = 0; return q + 4 + n; }

#define L_SEQ 91
#define REF_LEN 248956422

/* Writes one BAM record (including its 4-byte block_size) for read index i; returns its length. */
static unsigned synth_record(const fastf_synth_params *p, uint64_t i, uint8_t *dst)
{
    rng_t r = {mix64(p->seed ^ (i * 0x2545f4914f6cdd1dULL))};
    /* molecule -> (cell, gene, umi); reads pick a molecule uniformly, so duplicates are Poisson */
    uint64_t n_mol = p->n_molecules ? p->n_molecules : p->n_reads;
    uint64_t mol = rng_next(&r) % n_mol;
    rng_t m = {mix64(p->seed * 0x100000001b3ULL + 0x6d6f6cULL + mol)};
    uint32_t cell = rng_below(&m, p->n_cells);
    uint32_t gene = rng_below(&m, p->n_genes);
    char umi[FASTF_SYNTH_MAX_UMI + 1];
    for (unsigned k = 0; k < p->umi_len; k++) umi[k] = BASES[rng_below(&m, 4)];
    umi[p->umi_len] = 0;

    uint8_t *c = dst + 4;
    int32_t pos = (int32_t)(i % (REF_LEN - 200));
    char qname[64];
    int l_qname = snprintf(qname, sizeof qname, "A00519:%03u:HW%05uDSXX:%u:%04u:%05u:%05u", (unsigned)(p->seed % 1000), (unsigned)((i >> 24) & 0xffff),
                           1 + (unsigned)((i >> 22) & 3), 1101 + (unsigned)((i >> 12) & 1023) % 600, (unsigned)(rng_below(&r, 32000)), (unsigned)(rng_below(&r, 36000))) + 1;
    put32(c + 0, 0);                 /* refID */
    put32(c + 4, (uint32_t)pos);     /* pos */
    c[8] = (uint8_t)l_qname;
    c[9] = 255;                      /* mapq */
    put16(c + 10, (uint32_t)reg2bin(pos, pos + L_SEQ));
    put16(c + 12, 1);                /* n_cigar_op */
    put16(c + 14, (rng_next(&r) & 1) ? 16 : 0);
    put32(c + 16, L_SEQ);
    put32(c + 20, 0xffffffffu);      /* next_refID */
    put32(c + 24, 0xffffffffu);      /* next_pos */
    put32(c + 28, 0);                /* tlen */
    uint8_t *q = c + 32;
    memcpy(q, qname, (size_t)l_qname); q += l_qname;
    put32(q, (L_SEQ << 4) | 0); q += 4;   /* 91M */
    for (int k = 0; k < (L_SEQ + 1) / 2; k++) {
        uint64_t v = rng_next(&r);
        uint8_t hi = (uint8_t)(1u << (v & 3)), lo = (uint8_t)(1u << ((v >> 2) & 3));
        if (2 * k + 1 >= L_SEQ) lo = 0;
        *q++ = (uint8_t)((hi << 4) | lo);
    }
    for (int k = 0; k < L_SEQ; k++) *q++ = (uint8_t)(2 + rng_below(&r, 39));   /* phred 2..40 uniform */

    /* aux */
    q = aux_C(q, "NH", 1); q = aux_C(q, "HI", 1); q = aux_C(q, "AS", 80 + rng_below(&r, 10)); q = aux_C(q, "nM", rng_below(&r, 3));
    q = aux_Z(q, "RG", "synth:0:1:HW00000DSXX:1", 23);
    double ug = rng_unit(&r);
    int has_gx = ug < p->p_gx25 + p->p_gx17;
    int xf = ug < p->p_gx25 ? 25 : (has_gx ? 17 : 0);
    q = aux_A(q, "RE", has_gx ? 'E' : 'I');
    if (has_gx) {
        char gx[32], gn[32];
        int lg = snprintf(gx, sizeof gx, "ENSG%011u", gene + 1);
        int ln = snprintf(gn, sizeof gn, "GENE%u", gene + 1);
        q = aux_Z(q, "GX", gx, (size_t)lg);
        q = aux_Z(q, "GN", gn, (size_t)ln);
    }
    q = aux_C(q, "xf", (unsigned)xf);
    double uc = rng_unit(&r);
    char cb[20], cr[17];
    if (uc < p->p_cb_in_list) barcode_string(p, cell, cr);
    else for (int k = 0; k < 16; k++) cr[k] = BASES[rng_below(&r, 4)];
    cr[16] = 0;
    memcpy(cb, cr, 16); cb[16] = '-'; cb[17] = '1'; cb[18] = 0;
    int has_cb = uc < p->p_cb_in_list + p->p_cb_not_in_list;
    static const char QF[] = "FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF";
    q = aux_Z(q, "CR", cr, 16);
    q = aux_Z(q, "CY", QF, 16);
    if (has_cb) q = aux_Z(q, "CB", cb, 18);
    q = aux_Z(q, "UR", umi, p->umi_len);
    q = aux_Z(q, "UY", QF, p->umi_len);
    char ub[FASTF_SYNTH_MAX_UMI + 1];
    memcpy(ub, umi, (size_t)p->umi_len + 1);
    if (p->p_umi_n > 0 && rng_unit(&r) < p->p_umi_n) ub[rng_below(&r, p->umi_len)] = 'N';
    if (has_cb) q = aux_Z(q, "UB", ub, p->umi_len);
    unsigned block_size = (unsigned)(q - c);
    put32(dst, block_size);
    return block_size + 4;
}

/* ---------- parallel driver: one job per super-chunk ---------- */
typedef struct {
    const fastf_synth_params *p;
    int kind;                 /* 0 = BAM, 1 = FASTQ */
    uint64_t n_chunks;
    uint64_t next;            /* atomic job cursor */
    buf_t *chunks;            /* compressed bytes per chunk */
    uint64_t *chunk_inflated; /* inflated bytes per chunk */
    uint32_t *chunk_blocks;
    /* FASTQ only */
    const char *true_barcodes; /* n_true x 16 */
} job_t;

static void bam_chunk(job_t *J, uint64_t ci)
{
    const fastf_synth_params *p = J->p;
    uint64_t r0 = ci * SUPER_CHUNK_READS, r1 = r0 + SUPER_CHUNK_READS;
    if (r1 > p->n_reads) r1 = p->n_reads;
    uint8_t payload[BGZF_MAX_PAYLOAD + 1024];
    unsigned fill = 0;
    buf_t out = {0, 0, 0};
    uint64_t infl = 0;
    uint32_t nb = 0;
    for (uint64_t i = r0; i < r1; i++) {
        uint8_t rec[1024];
        unsigned n = synth_record(p, i, rec);
        if (fill + n > BGZF_MAX_PAYLOAD) { bgzf_emit_block(&out, payload, fill, p->zlevel, &nb); infl += fill; fill = 0; }
        memcpy(payload + fill, rec, n);
        fill += n;
    }
    if (fill) { bgzf_emit_block(&out, payload, fill, p->zlevel, &nb); infl += fill; }
    J->chunks[ci] = out;
    J->chunk_inflated[ci] = infl;
    J->chunk_blocks[ci] = nb;
}

/* FASTQ record i: "@r%09llu\n<16bp CB><umi>\n+\n<F x (16+umi)>\n" */
static unsigned synth_fastq_record(const fastf_synth_params *p, const char *true_bc, uint64_t i, char *dst)
{
    rng_t r = {mix64(p->seed * 0x9e3779b1ULL + 0xfa57ULL + i * 0x2545f4914f6cdd1dULL)};
    int n = sprintf(dst, "@r%09llu\n", (unsigned long long)i);
    char *q = dst + n;
    uint32_t k = rng_below(&r, p->n_cells);
    memcpy(q, true_bc + (size_t)k * 16, 16);
    if (rng_unit(&r) < p->p_bc_error) {
        unsigned pos = rng_below(&r, 16);
        char c;
        do c = BASES[rng_below(&r, 4)]; while (c == q[pos]);
        q[pos] = c;
    }
    if (p->p_umi_n > 0 && rng_unit(&r) < p->p_umi_n) q[rng_below(&r, 16)] = 'N';
    q += 16;
    for (unsigned u = 0; u < p->umi_len; u++) *q++ = BASES[rng_below(&r, 4)];
    *q++ = '\n'; *q++ = '+'; *q++ = '\n';
    for (unsigned u = 0; u < 16 + p->umi_len; u++) *q++ = 'F';
    *q++ = '\n';
    return (unsigned)(q - dst);
}

static void fastq_chunk(job_t *J, uint64_t ci)
{
    const fastf_synth_params *p = J->p;
    uint64_t r0 = ci * SUPER_CHUNK_READS, r1 = r0 + SUPER_CHUNK_READS;
    if (r1 > p->n_reads) r1 = p->n_reads;
    /* text of the whole chunk, then cut at fixed payload size: lines straddle BGZF blocks */
    size_t cap = (size_t)(r1 - r0) * (64 + 2 * (size_t)p->umi_len) + 64;
    char *text = (char *)malloc(cap);
    size_t n = 0;
    for (uint64_t i = r0; i < r1; i++) n += synth_fastq_record(p, J->true_barcodes, i, text + n);
    buf_t out = {0, 0, 0};
    uint32_t nb = 0;
    for (size_t o = 0; o < n; o += BGZF_MAX_PAYLOAD) {
        unsigned len = (unsigned)((n - o) < BGZF_MAX_PAYLOAD ? (n - o) : BGZF_MAX_PAYLOAD);
        bgzf_emit_block(&out, (const uint8_t *)text + o, len, p->zlevel, &nb);
    }
    free(text);
    J->chunks[ci] = out;
    J->chunk_inflated[ci] = n;
    J->chunk_blocks[ci] = nb;
}

static void *worker(void *arg)
{
    job_t *J = (job_t *)arg;
    for (;;) {
        uint64_t ci = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (ci >= J->n_chunks) break;
        if (J->kind == 0) bam_chunk(J, ci); else fastq_chunk(J, ci);
    }
    return NULL;
}

static int run_jobs(job_t *J, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > J->n_chunks) nthreads = (int)(J->n_chunks ? J->n_chunks : 1);
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; i++) pthread_create(&t[i], NULL, worker, J);
    for (int i = 0; i < nthreads; i++) pthread_join(t[i], NULL);
    free(t);
    return 0;
}

static void default_fill(fastf_synth_params *p)
{
    if (p->zlevel == 0) p->zlevel = 6;
    if (p->umi_len == 0) p->umi_len = 12;
    if (p->umi_len > FASTF_SYNTH_MAX_UMI) p->umi_len = FASTF_SYNTH_MAX_UMI;
    if (p->n_cells == 0) p->n_cells = 1000;
    if (p->n_genes == 0) p->n_genes = 2000;
}

void fastf_synth_defaults(fastf_synth_params *p)
{
    memset(p, 0, sizeof *p);
    p->n_reads = 100000; p->n_cells = 1000; p->n_genes = 2000; p->seed = 11; p->umi_len = 12; p->zlevel = 6;
    p->p_cb_in_list = 0.96; p->p_cb_not_in_list = 0.02; p->p_gx25 = 0.85; p->p_gx17 = 0.05; p->p_umi_n = 0.0; p->p_bc_error = 0.05;
}

static size_t bam_header_bytes(uint8_t *dst)
{
    static const char text[] = "@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:248956422\n@RG\tID:synth:0:1:HW00000DSXX:1\tSM:synth\n@PG\tID:fastf_synth\tPN:fastf_synth\n";
    size_t lt = sizeof(text) - 1, n = 0;
    memcpy(dst, "BAM\1", 4); n = 4;
    put32(dst + n, (uint32_t)lt); n += 4;
    memcpy(dst + n, text, lt); n += lt;
    put32(dst + n, 1); n += 4;
    put32(dst + n, 5); n += 4;
    memcpy(dst + n, "chr1\0", 5); n += 5;
    put32(dst + n, REF_LEN); n += 4;
    return n;
}

static int assemble(job_t *J, const uint8_t *head, size_t head_n, uint8_t **out, size_t *out_n, fastf_synth_stats *st)
{
    size_t total = head_n + sizeof BGZF_EOF;
    for (uint64_t i = 0; i < J->n_chunks; i++) total += J->chunks[i].n;
    uint8_t *o = (uint8_t *)malloc(total + 64);
    if (!o) return 1;
    size_t n = 0;
    if (head_n) { memcpy(o, head, head_n); n = head_n; }
    for (uint64_t i = 0; i < J->n_chunks; i++) {
        memcpy(o + n, J->chunks[i].p, J->chunks[i].n);
        n += J->chunks[i].n;
        free(J->chunks[i].p);
        if (st) { st->inflated_bytes += J->chunk_inflated[i]; st->n_blocks += J->chunk_blocks[i]; }
    }
    memcpy(o + n, BGZF_EOF, sizeof BGZF_EOF);
    n += sizeof BGZF_EOF;
    memset(o + n, 0, 64);
    if (st) { st->n_blocks += 1; st->compressed_bytes = n; }
    *out = o; *out_n = n;
    return 0;
}

int fastf_synth_bam(const fastf_synth_params *pin, int nthreads, uint8_t **out, size_t *out_n, fastf_synth_stats *st)
{
    fastf_synth_params P = *pin;
    default_fill(&P);
    if (st) memset(st, 0, sizeof *st);
    job_t J;
    memset(&J, 0, sizeof J);
    J.p = &P; J.kind = 0;
    J.n_chunks = (P.n_reads + SUPER_CHUNK_READS - 1) / SUPER_CHUNK_READS;
    J.chunks = (buf_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(buf_t));
    J.chunk_inflated = (uint64_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(uint64_t));
    J.chunk_blocks = (uint32_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(uint32_t));
    run_jobs(&J, nthreads);
    uint8_t hdr[1024];
    size_t hn = bam_header_bytes(hdr);
    buf_t hb = {0, 0, 0};
    uint32_t nb = 0;
    bgzf_emit_block(&hb, hdr, (unsigned)hn, P.zlevel, &nb);
    int rc = assemble(&J, hb.p, hb.n, out, out_n, st);
    if (st) { st->n_blocks += 1; st->inflated_bytes += hn; st->n_reads = P.n_reads; }
    free(hb.p); free(J.chunks); free(J.chunk_inflated); free(J.chunk_blocks);
    return rc;
}

int fastf_synth_fastq(const fastf_synth_params *pin, int nthreads, uint8_t **out, size_t *out_n, fastf_synth_stats *st)
{
    fastf_synth_params P = *pin;
    default_fill(&P);
    if (st) memset(st, 0, sizeof *st);
    char *bc = (char *)malloc((size_t)P.n_cells * 16 + 17);
    for (uint32_t k = 0; k < P.n_cells; k++) { char s[17]; barcode_string(&P, k, s); memcpy(bc + (size_t)k * 16, s, 16); }
    job_t J;
    memset(&J, 0, sizeof J);
    J.p = &P; J.kind = 1; J.true_barcodes = bc;
    J.n_chunks = (P.n_reads + SUPER_CHUNK_READS - 1) / SUPER_CHUNK_READS;
    J.chunks = (buf_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(buf_t));
    J.chunk_inflated = (uint64_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(uint64_t));
    J.chunk_blocks = (uint32_t *)calloc(J.n_chunks ? J.n_chunks : 1, sizeof(uint32_t));
    run_jobs(&J, nthreads);
    int rc = assemble(&J, NULL, 0, out, out_n, st);
    if (st) st->n_reads = P.n_reads;
    free(bc); free(J.chunks); free(J.chunk_inflated); free(J.chunk_blocks);
    return rc;
}

/* barcodes.tsv ("<16bp>-1\n" per cell, sorted) and features.tsv ("ENSG%011u\tGENE%u\tGene Expression\n") as plain text */
int fastf_synth_barcodes(const fastf_synth_params *pin, char **out, size_t *out_n)
{
    fastf_synth_params P = *pin;
    default_fill(&P);
    char *o = (char *)malloc((size_t)P.n_cells * 19 + 1);
    size_t n = 0;
    for (uint32_t k = 0; k < P.n_cells; k++) { barcode_string(&P, k, o + n); n += 16; o[n++] = '-'; o[n++] = '1'; o[n++] = '\n'; }
    o[n] = 0;
    *out = o; *out_n = n;
    return 0;
}
int fastf_synth_features(const fastf_synth_params *pin, char **out, size_t *out_n)
{
    fastf_synth_params P = *pin;
    default_fill(&P);
    char *o = (char *)malloc((size_t)P.n_genes * 64 + 1);
    size_t n = 0;
    for (uint32_t g = 0; g < P.n_genes; g++) n += (size_t)sprintf(o + n, "ENSG%011u\tGENE%u\tGene Expression\n", g + 1, g + 1);
    *out = o; *out_n = n;
    return 0;
}
void fastf_synth_free(void *p) { free(p); }

#ifdef FASTF_SYNTH_MAIN
#include <unistd.h>
static int write_file(const char *path, const void *p, size_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); return 1; }
    size_t w = fwrite(p, 1, n, f);
    fclose(f);
    return w != n;
}
static int write_gz(const char *path, const void *p, size_t n)
{
    gzFile g = gzopen(path, "wb");
    if (!g) { perror(path); return 1; }
    size_t o = 0;
    while (o < n) { unsigned c = (unsigned)((n - o) > (1u << 30) ? (1u << 30) : (n - o)); if (gzwrite(g, (const char *)p + o, c) <= 0) { gzclose(g); return 1; } o += c; }
    gzclose(g);
    return 0;
}
/* fastf_synth bam|fastq --out DIR [--reads N --cells C --genes G --seed S --umi L --threads T --molecules M --umi-n P --level Z] */
int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: fastf_synth bam|fastq --out DIR [--reads N --cells C --genes G --seed S --umi L --threads T --molecules M --umi-n P --bc-error P --level Z]\n"); return 2; }
    fastf_synth_params p;
    fastf_synth_defaults(&p);
    const char *outdir = ".";
    int threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    for (int i = 2; i + 1 < argc; i += 2) {
        const char *k = argv[i], *v = argv[i + 1];
        if (!strcmp(k, "--out")) outdir = v;
        else if (!strcmp(k, "--reads")) p.n_reads = strtoull(v, 0, 10);
        else if (!strcmp(k, "--cells")) p.n_cells = (uint32_t)strtoul(v, 0, 10);
        else if (!strcmp(k, "--genes")) p.n_genes = (uint32_t)strtoul(v, 0, 10);
        else if (!strcmp(k, "--seed")) p.seed = strtoull(v, 0, 10);
        else if (!strcmp(k, "--umi")) p.umi_len = (uint32_t)strtoul(v, 0, 10);
        else if (!strcmp(k, "--threads")) threads = atoi(v);
        else if (!strcmp(k, "--molecules")) p.n_molecules = strtoull(v, 0, 10);
        else if (!strcmp(k, "--umi-n")) p.p_umi_n = atof(v);
        else if (!strcmp(k, "--bc-error")) p.p_bc_error = atof(v);
        else if (!strcmp(k, "--level")) p.zlevel = atoi(v);
        else { fprintf(stderr, "unknown option %s\n", k); return 2; }
    }
    char path[2048];
    uint8_t *buf; size_t n; fastf_synth_stats st;
    if (!strcmp(argv[1], "bam")) {
        if (fastf_synth_bam(&p, threads, &buf, &n, &st)) return 1;
        snprintf(path, sizeof path, "%s/synth.bam", outdir);
        if (write_file(path, buf, n)) return 1;
        free(buf);
        char *t; size_t tn;
        fastf_synth_barcodes(&p, &t, &tn);
        snprintf(path, sizeof path, "%s/barcodes.tsv.gz", outdir);
        if (write_gz(path, t, tn)) return 1;
        free(t);
        fastf_synth_features(&p, &t, &tn);
        snprintf(path, sizeof path, "%s/features.tsv.gz", outdir);
        if (write_gz(path, t, tn)) return 1;
        free(t);
    } else if (!strcmp(argv[1], "fastq")) {
        if (fastf_synth_fastq(&p, threads, &buf, &n, &st)) return 1;
        snprintf(path, sizeof path, "%s/R1.fastq.gz", outdir);
        if (write_file(path, buf, n)) return 1;
        free(buf);
    } else return 2;
    fprintf(stderr, "synth %s: reads=%llu blocks=%llu inflated=%llu compressed=%llu (%.1f B/read inflated, %.1f compressed)\n", argv[1],
            (unsigned long long)st.n_reads, (unsigned long long)st.n_blocks, (unsigned long long)st.inflated_bytes, (unsigned long long)st.compressed_bytes,
            (double)st.inflated_bytes / (double)(st.n_reads ? st.n_reads : 1), (double)st.compressed_bytes / (double)(st.n_reads ? st.n_reads : 1));
    return 0;
}
#endif
