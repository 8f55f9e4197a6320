/* TEST INFRASTRUCTURE ONLY -- see fastf_oracle.h.  Plain-C CPU restatement of the reference's
 * bam2db and freq hot paths.  Every function cites the reference file:line it restates.
 * htslib (absent from /root/reference and this image) is restated from the SAM/BAM spec.
 */
#define _GNU_SOURCE
#include "fastf_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ====================================================================================== */
/* MT19937 -- Matsumoto & Nishimura mt19937ar; reference src/mt19937ar.c:60-73 (seeding),   */
/* :111-129 (twist), :133-137 (temper).                                                   */
/* ====================================================================================== */
void oracle_mt_init(oracle_mt *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) {
        uint32_t prev = s->mt[i - 1];
        s->mt[i] = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
    }
    s->mti = 624;
}
static void mt_twist(oracle_mt *s)
{
    uint32_t *x = s->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % 624] & 0x7fffffffu);
        uint32_t v = x[(k + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        x[k] = v;
    }
    s->mti = 0;
}
uint32_t oracle_mt_next(oracle_mt *s)
{
    if (s->mti >= 624) mt_twist(s);
    uint32_t y = s->mt[s->mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
void oracle_mt_stream(uint32_t seed, uint64_t n, uint32_t *out)
{
    oracle_mt s;
    oracle_mt_init(&s, seed);
    for (uint64_t i = 0; i < n; i++) out[i] = oracle_mt_next(&s);
}
/* genrand_real1 (src/mt19937ar.c:149-153) compared as at src/bam2db_ds.c:385-390 (rate is a float
 * promoted to double). */
int oracle_depth_keep(uint32_t u, float rate_depth)
{
    double r = (double)u * (1.0 / 4294967295.0);
    return !(r >= (double)rate_depth);
}

/* ====================================================================================== */
/* Cell sampling -- src/bam2db_ds.c:240-244 + src/utils.c:29-75.                           */
/* ====================================================================================== */
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
uint64_t oracle_sample_cells(uint64_t n_cells, float rate_cell, uint32_t seed, uint64_t *out, uint64_t *d0)
{
    /* size_t * float is evaluated in float32 (src/bam2db_ds.c:241) */
    float prod = (float)n_cells * rate_cell;
    uint64_t ns = (uint64_t)prod;
    *d0 = 0;
    if (ns > n_cells) return (uint64_t)-1;   /* reference prints an error and exit(1) (src/utils.c:40-45) */
    if (ns == n_cells) {                     /* early return, no draw consumed (src/utils.c:48-51) */
        for (uint64_t i = 0; i < n_cells; i++) out[i] = i;
        return ns;
    }
    oracle_mt s;
    oracle_mt_init(&s, seed);                /* SampleInt re-seeds (src/utils.c:32) */
    uint64_t *pool = (uint64_t *)malloc(sizeof(uint64_t) * (n_cells ? n_cells : 1));
    for (uint64_t i = 0; i < n_cells; i++) pool[i] = i;
    uint64_t remaining = n_cells;
    for (uint64_t i = 0; i < ns; i++) {      /* partial Fisher-Yates, swap-with-last (src/utils.c:53-62) */
        uint64_t j = (uint64_t)oracle_mt_next(&s) % remaining;
        out[i] = pool[j];
        pool[j] = pool[remaining - 1];
        remaining--;
    }
    free(pool);
    *d0 = ns;
    qsort(out, ns, sizeof(uint64_t), cmp_u64);   /* src/bam2db_ds.c:244 */
    return ns;
}

/* ====================================================================================== */
/* 2-bit codec -- src/bam2db_ds.c:5-51: A=0 C=1 G=2 T=3, MSB first, 4 bases per byte.       */
/* ====================================================================================== */
int oracle_encode_dna(const char *s, uint8_t *out, size_t cap)
{
    size_t len = strlen(s);
    size_t nb = (len + 3) / 4;
    if (nb > cap) return -2;
    memset(out, 0, nb);
    for (size_t i = 0; i < len; i++) {
        unsigned code;
        switch (s[i]) {
        case 'A': code = 0; break;
        case 'C': code = 1; break;
        case 'G': code = 2; break;
        case 'T': code = 3; break;
        default: return -1;
        }
        out[i / 4] |= (uint8_t)(code << (6 - 2 * (i % 4)));
    }
    return (int)nb;
}

/* ====================================================================================== */
/* BGZF (SAMv1 section 4.1): gzip members with a 'BC' extra subfield holding BSIZE-1.       */
/* Restates what htslib's sam_open/sam_read1 do underneath (call sites src/bam2db_ds.c:141,360). */
/* ====================================================================================== */
static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t rd16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

int oracle_bgzf_inflate(const uint8_t *file, size_t n, uint8_t **out, size_t *out_n,
                        uint64_t **blk_in_off, uint32_t **blk_in_len, uint32_t **blk_isize, uint64_t *n_blocks)
{
    size_t cap = n * 4 + 65536, used = 0;
    uint8_t *o = (uint8_t *)malloc(cap);
    uint64_t nb = 0, bcap = 1024;
    uint64_t *bo = (uint64_t *)malloc(sizeof(uint64_t) * bcap);
    uint32_t *bl = (uint32_t *)malloc(sizeof(uint32_t) * bcap);
    uint32_t *bi = (uint32_t *)malloc(sizeof(uint32_t) * bcap);
    size_t pos = 0;
    int rc = 0;
    while (pos < n) {
        if (n - pos < 18) { rc = 1; break; }
        const uint8_t *h = file + pos;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { rc = 2; break; }
        uint32_t xlen = rd16(h + 10);
        if (n - pos < 12 + (size_t)xlen) { rc = 1; break; }
        uint32_t bsize = 0;
        int found = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *sf = h + 12 + x;
            uint32_t slen = rd16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) { bsize = rd16(sf + 4) + 1; found = 1; }
            x += 4 + slen;
        }
        if (!found || n - pos < bsize || bsize < 12 + xlen + 8) { rc = 3; break; }
        const uint8_t *cdata = h + 12 + xlen;
        uint32_t clen = bsize - 12 - xlen - 8;
        uint32_t crc = rd32(h + bsize - 8), isize = rd32(h + bsize - 4);
        if (isize > 65536) { rc = 4; break; }
        if (used + isize > cap) { cap = cap * 2 + isize; o = (uint8_t *)realloc(o, cap); }
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { rc = 5; break; }
        zs.next_in = (Bytef *)cdata; zs.avail_in = clen;
        zs.next_out = o + used; zs.avail_out = isize;
        int zr = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (zr != Z_STREAM_END || zs.total_out != isize) { rc = 6; break; }
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), o + used, isize) != crc) { rc = 7; break; }
        if (nb == bcap) {
            bcap *= 2;
            bo = (uint64_t *)realloc(bo, sizeof(uint64_t) * bcap);
            bl = (uint32_t *)realloc(bl, sizeof(uint32_t) * bcap);
            bi = (uint32_t *)realloc(bi, sizeof(uint32_t) * bcap);
        }
        bo[nb] = pos; bl[nb] = bsize; bi[nb] = isize; nb++;
        used += isize;
        pos += bsize;
    }
    if (rc) { free(o); free(bo); free(bl); free(bi); return rc; }
    *out = o; *out_n = used;
    if (blk_in_off) *blk_in_off = bo; else free(bo);
    if (blk_in_len) *blk_in_len = bl; else free(bl);
    if (blk_isize) *blk_isize = bi; else free(bi);
    if (n_blocks) *n_blocks = nb;
    return 0;
}

static int read_whole_file(const char *path, uint8_t **buf, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = (uint8_t *)malloc((size_t)sz + 64);
    size_t got = fread(b, 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz) { free(b); return 1; }
    memset(b + sz, 0, 64);
    *buf = b; *n = (size_t)sz;
    return 0;
}

/* ====================================================================================== */
/* Exact-string membership tables (semantics of src/hashtable.c:70-115: insert refuses a    */
/* duplicate key, lookup is strcmp-exact).  Open addressing here; only the semantics matter. */
/* ====================================================================================== */
typedef struct { char **keys; uint64_t *vals; size_t cap; } strmap;
static uint64_t fnv1a(const char *s) { uint64_t h = 1469598103934665603ULL; for (; *s; s++) { h ^= (uint8_t)*s; h *= 1099511628211ULL; } return h; }
static void strmap_init(strmap *m, size_t n) { size_t c = 64; while (c < n * 2 + 8) c <<= 1; m->cap = c; m->keys = (char **)calloc(c, sizeof(char *)); m->vals = (uint64_t *)calloc(c, sizeof(uint64_t)); }
static int strmap_get(const strmap *m, const char *k, uint64_t *v)
{
    if (!k) return 0;
    size_t i = (size_t)fnv1a(k) & (m->cap - 1);
    while (m->keys[i]) { if (!strcmp(m->keys[i], k)) { *v = m->vals[i]; return 1; } i = (i + 1) & (m->cap - 1); }
    return 0;
}
static int strmap_put(strmap *m, const char *k, uint64_t v)   /* returns 0 if the key exists */
{
    size_t i = (size_t)fnv1a(k) & (m->cap - 1);
    while (m->keys[i]) { if (!strcmp(m->keys[i], k)) return 0; i = (i + 1) & (m->cap - 1); }
    m->keys[i] = strdup(k); m->vals[i] = v;
    return 1;
}
static void strmap_free(strmap *m) { for (size_t i = 0; i < m->cap; i++) free(m->keys[i]); free(m->keys); free(m->vals); }

/* ====================================================================================== */
/* BAM aux access restated from the SAM/BAM spec (htslib semantics: first match; corrupt -> absent) */
/* ====================================================================================== */
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end)   /* s points at the type byte */
{
    if (s >= end) return NULL;
    int t = *s++;
    size_t sz = 0;
    switch (t) {
    case 'A': case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    case 'd': sz = 8; break;
    case 'Z': case 'H':
        while (s < end && *s) s++;
        return s < end ? s + 1 : NULL;
    case 'B': {
        if (end - s < 5) return NULL;
        int sub = *s;
        uint32_t cnt = rd32(s + 1);
        s += 5;
        size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
        if (!es || (uint64_t)(end - s) < (uint64_t)es * cnt) return NULL;
        return s + (size_t)es * cnt;
    }
    default: return NULL;
    }
    if ((size_t)(end - s) < sz) return NULL;
    return s + sz;
}
static const uint8_t *aux_find(const uint8_t *aux, const uint8_t *end, char a, char b)   /* -> type byte or NULL */
{
    const uint8_t *s = aux;
    while (s && end - s >= 3) {
        const uint8_t *next = aux_skip(s + 2, end);
        if (s[0] == (uint8_t)a && s[1] == (uint8_t)b) return next ? s + 2 : NULL;
        s = next;
    }
    return NULL;
}
static const char *aux_as_Z(const uint8_t *t) { return (t && (*t == 'Z' || *t == 'H')) ? (const char *)(t + 1) : NULL; }
static int64_t aux_as_int(const uint8_t *t)
{
    if (!t) return 0;
    const uint8_t *v = t + 1;
    switch (*t) {
    case 'c': return (int8_t)v[0];
    case 'C': return v[0];
    case 's': return (int16_t)rd16(v);
    case 'S': return (uint16_t)rd16(v);
    case 'i': return (int32_t)rd32(v);
    case 'I': return rd32(v);
    default: return 0;
    }
}

/* ====================================================================================== */
/* bam2db -- src/bam2db_ds.c:106-573                                                       */
/* ====================================================================================== */
typedef struct { uint32_t cell, gene; int32_t nb; uint64_t umi; } row_t;
static int row_cmp(const void *a, const void *b)
{
    const row_t *x = (const row_t *)a, *y = (const row_t *)b;
    if (x->cell != y->cell) return x->cell < y->cell ? -1 : 1;
    if (x->gene != y->gene) return x->gene < y->gene ? -1 : 1;
    /* NULL first, then memcmp order (left-aligned big-endian u64 compares like memcmp), then length */
    if ((x->nb < 0) != (y->nb < 0)) return x->nb < 0 ? -1 : 1;
    if (x->umi != y->umi) return x->umi < y->umi ? -1 : 1;
    if (x->nb != y->nb) return x->nb < y->nb ? -1 : 1;
    return 0;
}

void oracle_bam2db_free(oracle_bam2db_result *r)
{
    free(r->row_cell); free(r->row_gene); free(r->row_umi_nbytes); free(r->row_umi);
    free(r->m_gene); free(r->m_cell); free(r->m_count);
    memset(r, 0, sizeof *r);
}

int oracle_bam2db(const char *bam_file, const char *barcodes_file, const char *features_file,
                  float rate_cell, float rate_depth, unsigned seed, const char *out_dir,
                  oracle_bam2db_result *res)
{
    memset(res, 0, sizeof *res);
    char line[1024];

    /* --- barcodes: count lines (src/bam2db_ds.c:229-237) --- */
    gzFile g = gzopen(barcodes_file, "r");
    if (!g) return 1;
    uint64_t n_cells = 0;
    while (gzgets(g, line, 1024) != NULL) n_cells++;
    res->n_cells = n_cells;
    if (n_cells == 0) { gzclose(g); return 2; }   /* reference walks off the end here */

    /* --- sample (src/bam2db_ds.c:240-246) --- */
    uint64_t *samp = (uint64_t *)malloc(sizeof(uint64_t) * n_cells);
    uint64_t d0 = 0;
    uint64_t ns = oracle_sample_cells(n_cells, rate_cell, seed, samp, &d0);
    if (ns == (uint64_t)-1) { gzclose(g); free(samp); return 3; }
    res->n_cells_sampled = ns;
    res->d0 = d0;

    /* --- second pass over barcodes (src/bam2db_ds.c:248-289) --- */
    gzrewind(g);
    strmap cells;
    strmap_init(&cells, ns);
    char **cell_rows = (char **)malloc(sizeof(char *) * (ns ? ns : 1));
    uint64_t cell_index = 1, nth = 0;
    while (gzgets(g, line, 1024) != NULL && cell_index <= ns) {
        nth++;
        if (nth - 1 != samp[cell_index - 1]) continue;
        line[strcspn(line, "\n\r\t")] = 0;
        if (strmap_put(&cells, line, cell_index)) { cell_rows[cell_index - 1] = strdup(line); cell_index++; }
        /* duplicate: index not advanced, so no later line can match (src/bam2db_ds.c:260,281-285) */
    }
    gzclose(g);
    free(samp);
    uint64_t n_cell_rows = cell_index - 1;
    res->n_cell_rows = n_cell_rows;

    /* --- features (src/bam2db_ds.c:296-337) --- */
    g = gzopen(features_file, "r");
    if (!g) return 1;
    size_t fcap = 1024, nf = 0;
    char **f_id = (char **)malloc(sizeof(char *) * fcap), **f_name = (char **)malloc(sizeof(char *) * fcap), **f_type = (char **)malloc(sizeof(char *) * fcap);
    /* two passes so the map can be sized */
    uint64_t nlines = 0;
    while (gzgets(g, line, 1024) != NULL) nlines++;
    gzrewind(g);
    strmap feats;
    strmap_init(&feats, nlines);
    memset(line, 0, sizeof line);
    while (gzgets(g, line, 1024) != NULL) {
        char *id = strtok(line, "\t");
        char *name = strtok(NULL, "\t");
        char *type = strtok(NULL, "\t");
        if (!id || !name || !type) { gzclose(g); return 4; }   /* reference dereferences NULL here */
        type[strcspn(type, "\n\r\t")] = 0;
        if (strmap_put(&feats, line, nf + 1)) {      /* key is the buffer up to its first NUL (src/bam2db_ds.c:313) */
            if (nf == fcap) { fcap *= 2; f_id = (char **)realloc(f_id, sizeof(char *) * fcap); f_name = (char **)realloc(f_name, sizeof(char *) * fcap); f_type = (char **)realloc(f_type, sizeof(char *) * fcap); }
            f_id[nf] = strdup(id); f_name[nf] = strdup(name); f_type[nf] = strdup(type);
            nf++;
        }
    }
    gzclose(g);
    res->n_features = nf;

    /* --- BAM loop (src/bam2db_ds.c:340-438) --- */
    uint8_t *file, *bam;
    size_t fn, bn;
    if (read_whole_file(bam_file, &file, &fn)) return 1;
    int irc = oracle_bgzf_inflate(file, fn, &bam, &bn, NULL, NULL, NULL, NULL);
    free(file);
    if (irc) return 10 + irc;
    if (bn < 12 || memcmp(bam, "BAM\1", 4)) { free(bam); return 20; }
    size_t p = 8 + (size_t)rd32(bam + 4);
    uint32_t n_ref = rd32(bam + p);
    p += 4;
    for (uint32_t i = 0; i < n_ref; i++) p += 8 + (size_t)rd32(bam + p);

    oracle_mt mt;
    oracle_mt_init(&mt, seed);                       /* src/bam2db_ds.c:122, then SampleInt re-seeds with the same seed */
    for (uint64_t i = 0; i < d0; i++) oracle_mt_next(&mt);   /* draws consumed by SampleInt */

    size_t rcap = 1 << 16, nrows = 0;
    row_t *rows = (row_t *)malloc(sizeof(row_t) * rcap);
    uint64_t total = 0, cbv = 0, sampled = 0, valid = 0;
    while (p + 4 <= bn) {                            /* sam_read1 >= 0 */
        int32_t bs = (int32_t)rd32(bam + p);
        if (bs < 32 || p + 4 + (size_t)bs > bn) break;
        const uint8_t *c = bam + p + 4, *end = c + bs;
        p += 4 + (size_t)bs;
        int64_t aoff = 32 + (int64_t)c[8] + 4 * (int64_t)rd16(c + 12) + ((int64_t)(int32_t)rd32(c + 16) + 1) / 2 + (int32_t)rd32(c + 16);
        if ((int32_t)rd32(c + 16) < 0 || aoff > bs) break;
        const uint8_t *aux = c + aoff;
        total++;
        const char *cb = aux_as_Z(aux_find(aux, end, 'C', 'B'));      /* :366-380 */
        uint64_t cidx;
        if (!strmap_get(&cells, cb, &cidx)) continue;
        cbv++;
        uint32_t u = oracle_mt_next(&mt);                              /* :385 */
        if (!oracle_depth_keep(u, rate_depth)) continue;               /* :387-390 */
        sampled++;
        const uint8_t *xfp = aux_find(aux, end, 'x', 'f');             /* :394-400 (absent xf crashes the reference; treated as 0 here) */
        int xf = (int)aux_as_int(xfp);
        if (!(xf == 25 || xf == 17)) continue;
        const char *gx = aux_as_Z(aux_find(aux, end, 'G', 'X'));      /* :403-410 */
        uint64_t gidx;
        if (!strmap_get(&feats, gx, &gidx)) continue;
        const uint8_t *ubp = aux_find(aux, end, 'U', 'B');             /* :412-416 */
        if (!ubp) continue;
        const char *ub = aux_as_Z(ubp);
        if (!ub) continue;                                             /* non-Z UB crashes the reference; skipped here */
        uint8_t enc[64];
        int nb = oracle_encode_dna(ub, enc, 8);                        /* :418-419 */
        if (nb == -2) { free(bam); free(rows); return 30; }            /* UMI longer than 32 bases: outside the oracle's range */
        if (nrows == rcap) { rcap *= 2; rows = (row_t *)realloc(rows, sizeof(row_t) * rcap); }
        row_t *r = &rows[nrows++];
        r->cell = (uint32_t)cidx; r->gene = (uint32_t)gidx; r->nb = nb < 0 ? -1 : nb; r->umi = 0;
        for (int k = 0; k < nb; k++) r->umi |= (uint64_t)enc[k] << (56 - 8 * k);
        valid++;                                                       /* :435 */
    }
    free(bam);
    res->total = total; res->cb_valid = cbv; res->sampled = sampled; res->valid = valid;

    res->n_rows = nrows;
    res->row_cell = (uint32_t *)malloc(sizeof(uint32_t) * (nrows ? nrows : 1));
    res->row_gene = (uint32_t *)malloc(sizeof(uint32_t) * (nrows ? nrows : 1));
    res->row_umi_nbytes = (int32_t *)malloc(sizeof(int32_t) * (nrows ? nrows : 1));
    res->row_umi = (uint64_t *)malloc(sizeof(uint64_t) * (nrows ? nrows : 1));
    for (size_t i = 0; i < nrows; i++) { res->row_cell[i] = rows[i].cell; res->row_gene[i] = rows[i].gene; res->row_umi_nbytes[i] = rows[i].nb; res->row_umi[i] = rows[i].umi; }

    /* --- COUNT(DISTINCT encoded_umi) GROUP BY cell_index, feature_index (src/bam2db_ds.c:480-483) --- */
    qsort(rows, nrows, sizeof(row_t), row_cmp);
    res->m_gene = (uint32_t *)malloc(sizeof(uint32_t) * (nrows ? nrows : 1));
    res->m_cell = (uint32_t *)malloc(sizeof(uint32_t) * (nrows ? nrows : 1));
    res->m_count = (uint32_t *)malloc(sizeof(uint32_t) * (nrows ? nrows : 1));
    uint64_t nnz = 0;
    for (size_t i = 0; i < nrows;) {
        size_t j = i;
        uint32_t cnt = 0;
        while (j < nrows && rows[j].cell == rows[i].cell && rows[j].gene == rows[i].gene) {
            if (rows[j].nb >= 0 && (j == i || rows[j - 1].nb < 0 || rows[j].umi != rows[j - 1].umi || rows[j].nb != rows[j - 1].nb)) cnt++;
            j++;
        }
        res->m_gene[nnz] = rows[i].gene; res->m_cell[nnz] = rows[i].cell; res->m_count[nnz] = cnt;
        nnz++;
        i = j;
    }
    res->nnz = nnz;
    free(rows);

    /* --- text outputs (src/bam2db_ds.c:498-525; decompressed bytes of the three .gz files) --- */
    if (out_dir) {
        char path[2048];
        snprintf(path, sizeof path, "%s/matrix.mtx", out_dir);
        FILE *f = fopen(path, "w");
        if (!f) return 40;
        fprintf(f, "%%%%MatrixMarket matrix coordinate integer general\n%%metadata_json: \n%%{\n"
                   "%%\t\"software_version\": \"fastF-1.0.0\",\n%%\t\"format_version\": 1,\n%%\t\"parent_bam\": \"%s\",\n"
                   "%%\t\"rate_cell\": %.3f,\n%%\t\"rate_depth\": %.3f,\n%%\t\"total_n_FastQ\": %zu,\n%%\t\"sampled_n_FastQ\": %zu,\n"
                   "%%\t\"sampled_valid_n_FastQ\": %zu\n%%}\n",
                bam_file, rate_cell, rate_depth, (size_t)total, (size_t)sampled, (size_t)valid);
        fprintf(f, "%zu %zu %zu\n", nf, (size_t)n_cell_rows, (size_t)nnz);
        for (uint64_t i = 0; i < nnz; i++) fprintf(f, "%d %d %d\n", (int)res->m_gene[i], (int)res->m_cell[i], (int)res->m_count[i]);
        fclose(f);
        snprintf(path, sizeof path, "%s/barcodes.tsv", out_dir);
        f = fopen(path, "w");
        if (!f) return 40;
        for (uint64_t i = 0; i < n_cell_rows; i++) fprintf(f, "%s\n", cell_rows[i]);
        fclose(f);
        snprintf(path, sizeof path, "%s/features.tsv", out_dir);
        f = fopen(path, "w");
        if (!f) return 40;
        for (size_t i = 0; i < nf; i++) fprintf(f, "%s\t%s\t%s\n", f_id[i], f_name[i], f_type[i]);
        fclose(f);
    }
    for (uint64_t i = 0; i < n_cell_rows; i++) free(cell_rows[i]);
    free(cell_rows);
    for (size_t i = 0; i < nf; i++) { free(f_id[i]); free(f_name[i]); free(f_type[i]); }
    free(f_id); free(f_name); free(f_type);
    strmap_free(&cells); strmap_free(&feats);
    return 0;
}

/* ====================================================================================== */
/* freq -- src/count.c:3-21; reader src/filter.c:15-37; key src/filter.c:260-275;           */
/* histogram = unbalanced BST in read order (src/filter.c:105-124), printed in PRE-order    */
/* (src/filter.c:139-148).                                                                 */
/* ====================================================================================== */
typedef struct fnode { char *key; long count; struct fnode *lo, *hi; } fnode;

int oracle_freq(const char *r1_file, size_t len_cellbarcode, size_t len_umi, const char *out_path,
                uint64_t *n_reads, uint64_t *n_keys)
{
    gzFile g = gzopen(r1_file, "r");
    if (!g) return 1;
    gzbuffer(g, 1 << 20);
    size_t klen = len_cellbarcode + len_umi;
    char id[1024], seq[1024], plus[1024], qual[1024];
    char *key = (char *)malloc(klen + 1);
    fnode *root = NULL;
    uint64_t reads = 0, keys = 0;
    for (;;) {
        if (gzgets(g, id, 1024) == NULL) break;       /* EOF when the id line cannot be read (src/filter.c:23,29-34) */
        seq[0] = 0;
        gzgets(g, seq, 1024);
        gzgets(g, plus, 1024);
        gzgets(g, qual, 1024);
        strncpy(key, seq, klen);                      /* src/filter.c:270: short lines keep their '\n', then NUL padding */
        key[klen] = 0;
        fnode **slot = &root;                         /* iterative form of insert_tree */
        while (*slot) {
            int c = strcmp(key, (*slot)->key);
            if (c == 0) { (*slot)->count++; break; }
            slot = c < 0 ? &(*slot)->lo : &(*slot)->hi;
        }
        if (!*slot) { fnode *nn = (fnode *)calloc(1, sizeof(fnode)); nn->key = strdup(key); nn->count = 1; *slot = nn; keys++; }
        reads++;
    }
    gzclose(g);
    free(key);
    FILE *f = fopen(out_path, "w");
    if (!f) return 2;
    /* pre-order with an explicit stack: node, left subtree, right subtree */
    size_t scap = 1024, sp = 0;
    fnode **stack = (fnode **)malloc(sizeof(fnode *) * scap);
    if (root) stack[sp++] = root;
    while (sp) {
        fnode *nd = stack[--sp];
        fprintf(f, "%s,%ld\n", nd->key, nd->count);
        if (sp + 2 > scap) { scap *= 2; stack = (fnode **)realloc(stack, sizeof(fnode *) * scap); }
        if (nd->hi) stack[sp++] = nd->hi;
        if (nd->lo) stack[sp++] = nd->lo;
        free(nd->key); free(nd);
    }
    free(stack);
    fclose(f);
    if (n_reads) *n_reads = reads;
    if (n_keys) *n_keys = keys;
    return 0;
}

/* ====================================================================================== */
/* crb / extract -- src/extract.c:4-216 (two more per-record histograms over the same BAM reader; */
/* output order is again the pre-order of an unbalanced BST built in read order)                  */
/* ====================================================================================== */
typedef struct cbnode { char *cb; fnode *cr; struct cbnode *lo, *hi; } cbnode;
static void fnode_insert(fnode **root, const char *key)             /* insert_tree, src/filter.c:105-124 */
{
    fnode **slot = root;
    while (*slot) {
        int c = strcmp(key, (*slot)->key);
        if (c == 0) { (*slot)->count++; return; }
        slot = c < 0 ? &(*slot)->lo : &(*slot)->hi;
    }
    fnode *nn = (fnode *)calloc(1, sizeof(fnode));
    nn->key = strdup(key);
    nn->count = 1;
    *slot = nn;
}
static void fnode_print(fnode *nd, FILE *f, const char *fmt)        /* print_tree / print_tree_same_row: node, left, right */
{
    if (!nd) return;
    fprintf(f, fmt, nd->key, nd->count);
    fnode_print(nd->lo, f, fmt);
    fnode_print(nd->hi, f, fmt);
    free(nd->key);
    free(nd);
}
static void cbnode_print(cbnode *nd, FILE *f)                        /* print_CB_node, src/extract.c:47-62 */
{
    if (!nd) return;
    fprintf(f, "%s;", nd->cb);
    fnode_print(nd->cr, f, "%s,%ld;");
    fprintf(f, "\n");
    cbnode_print(nd->lo, f);
    cbnode_print(nd->hi, f);
    free(nd->cb);
    free(nd);
}
/* walks the records of an inflated BAM; returns the offset of the first record or 0 on a bad header */
static size_t bam_first_record(const uint8_t *bam, size_t bn)
{
    if (bn < 12 || memcmp(bam, "BAM\1", 4)) return 0;
    size_t p = 8 + (size_t)rd32(bam + 4);
    uint32_t n_ref = rd32(bam + p);
    p += 4;
    for (uint32_t i = 0; i < n_ref; i++) p += 8 + (size_t)rd32(bam + p);
    return p;
}
static const uint8_t *bam_next_aux(const uint8_t *bam, size_t bn, size_t *p, const uint8_t **end)   /* sam_read1 >= 0 */
{
    if (*p + 4 > bn) return NULL;
    int32_t bs = (int32_t)rd32(bam + *p);
    if (bs < 32 || *p + 4 + (size_t)bs > bn) return NULL;
    const uint8_t *c = bam + *p + 4;
    *end = c + bs;
    *p += 4 + (size_t)bs;
    int64_t aoff = 32 + (int64_t)c[8] + 4 * (int64_t)rd16(c + 12) + ((int64_t)(int32_t)rd32(c + 16) + 1) / 2 + (int32_t)rd32(c + 16);
    if ((int32_t)rd32(c + 16) < 0 || aoff > bs) return NULL;
    return c + aoff;
}
static int load_bam(const char *bam_file, uint8_t **bam, size_t *bn, size_t *first)
{
    uint8_t *file;
    size_t fn;
    if (read_whole_file(bam_file, &file, &fn)) return 1;
    int irc = oracle_bgzf_inflate(file, fn, bam, bn, NULL, NULL, NULL, NULL);
    free(file);
    if (irc) return 10 + irc;
    *first = bam_first_record(*bam, *bn);
    if (!*first) { free(*bam); return 20; }
    return 0;
}

/* read_bam + print_CB_node (src/extract.c:64-133, src/main.c:231-286): out_path receives the DECOMPRESSED bytes of the
 * reference's gz output.  Returns 3 where the reference dereferences NULL (CB present but CR absent, or either not a string). */
int oracle_crb(const char *bam_file, const char *out_path, uint64_t *n_reads)
{
    uint8_t *bam;
    size_t bn, p;
    int rc = load_bam(bam_file, &bam, &bn, &p);
    if (rc) return rc;
    cbnode *root = NULL;
    uint64_t reads = 0;
    const uint8_t *aux, *end;
    while ((aux = bam_next_aux(bam, bn, &p, &end)) != NULL) {
        const uint8_t *cbp = aux_find(aux, end, 'C', 'B'), *crp = aux_find(aux, end, 'C', 'R');
        if (cbp) {                                                   /* src/extract.c:92-103 */
            const char *cb = aux_as_Z(cbp), *cr = aux_as_Z(crp);
            if (!cb || !cr) { free(bam); return 3; }
            cbnode **slot = &root;                                   /* insert_CB_node, src/extract.c:4-31 */
            while (*slot) {
                int c = strcmp(cb, (*slot)->cb);
                if (c == 0) break;
                slot = c < 0 ? &(*slot)->lo : &(*slot)->hi;
            }
            if (!*slot) { cbnode *nn = (cbnode *)calloc(1, sizeof(cbnode)); nn->cb = strdup(cb); *slot = nn; }
            fnode_insert(&(*slot)->cr, cr);
        }
        reads++;
    }
    free(bam);
    FILE *f = fopen(out_path, "w");
    if (!f) return 2;
    cbnode_print(root, f);
    fclose(f);
    if (n_reads) *n_reads = reads;
    return 0;
}

/* extract_bam (src/extract.c:135-216): histogram of one aux tag; type 0 = string (bam_aux2Z), 1 = integer printed with "%d".
 * *total is the reference's total_count, which it increments TWICE per record (src/extract.c:162,164).  Returns 3 where the
 * reference dereferences NULL (type 0 on a tag that is not Z/H). */
int oracle_extract(const char *bam_file, const char *tag, int type, const char *out_path, uint64_t *total, uint64_t *valid)
{
    uint8_t *bam;
    size_t bn, p;
    int rc = load_bam(bam_file, &bam, &bn, &p);
    if (rc) return rc;
    fnode *root = NULL;
    uint64_t tot = 0, val = 0;
    const uint8_t *aux, *end;
    while ((aux = bam_next_aux(bam, bn, &p, &end)) != NULL) {
        tot += 2;
        const uint8_t *tp = aux_find(aux, end, tag[0], tag[1]);
        if (!tp) continue;
        val++;
        if (type == 0) {
            const char *z = aux_as_Z(tp);
            if (!z) { free(bam); return 3; }
            fnode_insert(&root, z);
        } else {
            char buf[32];
            snprintf(buf, sizeof buf, "%d", (int)aux_as_int(tp));     /* "%d" of the int64 bam_aux2i(): its low 32 bits */
            fnode_insert(&root, buf);
        }
    }
    free(bam);
    FILE *f = fopen(out_path, "w");
    if (!f) return 2;
    fnode_print(root, f, "%s,%ld\n");
    fclose(f);
    if (total) *total = tot;
    if (valid) *valid = val;
    return 0;
}

#ifdef FASTF_ORACLE_MAIN
int main(int argc, char **argv)
{
    if (argc >= 9 && !strcmp(argv[1], "bam2db")) {
        oracle_bam2db_result r;
        int rc = oracle_bam2db(argv[2], argv[3], argv[4], strtof(argv[5], 0), strtof(argv[6], 0), (unsigned)strtoul(argv[7], 0, 10), argv[8], &r);
        if (rc) { fprintf(stderr, "oracle bam2db failed: %d\n", rc); return 1; }
        printf("total=%llu cb_valid=%llu sampled=%llu valid=%llu nnz=%llu cells=%llu/%llu d0=%llu\n", (unsigned long long)r.total, (unsigned long long)r.cb_valid,
               (unsigned long long)r.sampled, (unsigned long long)r.valid, (unsigned long long)r.nnz, (unsigned long long)r.n_cell_rows, (unsigned long long)r.n_cells, (unsigned long long)r.d0);
        oracle_bam2db_free(&r);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "freq")) {
        uint64_t nr, nk;
        int rc = oracle_freq(argv[2], strtoul(argv[3], 0, 10), strtoul(argv[4], 0, 10), argv[5], &nr, &nk);
        if (rc) { fprintf(stderr, "oracle freq failed: %d\n", rc); return 1; }
        printf("reads=%llu keys=%llu\n", (unsigned long long)nr, (unsigned long long)nk);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "crb")) {
        uint64_t nr;
        int rc = oracle_crb(argv[2], argv[3], &nr);
        if (rc) { fprintf(stderr, "oracle crb failed: %d\n", rc); return 1; }
        printf("reads=%llu\n", (unsigned long long)nr);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "extract")) {
        uint64_t t, v;
        int rc = oracle_extract(argv[2], argv[3], atoi(argv[4]), argv[5], &t, &v);
        if (rc) { fprintf(stderr, "oracle extract failed: %d\n", rc); return 1; }
        printf("total=%llu valid=%llu\n", (unsigned long long)t, (unsigned long long)v);
        return 0;
    }
    fprintf(stderr, "usage: oracle_cli bam2db BAM BARCODES FEATURES RATE_CELL RATE_DEPTH SEED OUTDIR | freq R1 L U OUT\n");
    return 2;
}
#endif
