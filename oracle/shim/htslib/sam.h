/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <htslib/sam.h>.
 *
 * htslib is a third-party dependency of the reference that is NOT vendored under
 * /root/reference (CMakeLists.txt:13-14 points at an absent external/htslib; .gitignore:4
 * hints at htslib-1.17) and is not installed in this image.  This header implements, from the
 * SAM/BAM specification (SAMv1 section 4: BGZF + BAM record layout + aux encoding), the entry
 * points the reference calls (src/bam2db_ds.c:141,340-341,360,366,374,394-395,403-404,412,417,
 * 447-449; src/extract.c:67-130,138-215) so that the UNMODIFIED reference sources compile where
 * they lie.  BGZF is a series of gzip members, which zlib's gzread() inflates transparently.
 * Semantics kept from htslib: bam_aux_get returns the FIRST matching tag (pointer to its type
 * byte); bam_aux2Z returns NULL unless the type is Z or H; bam_aux2i accepts c C s S i I and
 * returns 0 otherwise; sam_read1 returns -1 at clean EOF and < -1 on a truncated record. */
#ifndef FASTF_ORACLE_HTSLIB_SAM_SHIM_H
#define FASTF_ORACLE_HTSLIB_SAM_SHIM_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { gzFile fp; } htsFile;
typedef htsFile samFile;
typedef struct { int32_t n_targets; char *text; uint32_t l_text; } bam_hdr_t;
typedef bam_hdr_t sam_hdr_t;
typedef struct {
    uint8_t *data;   /* the record bytes after block_size (32-byte core first) */
    int l_data;      /* = block_size */
    int m_data;
    int aux_off;     /* offset of the first aux field inside data */
} bam1_t;

static inline htsFile *hts_open(const char *fn, const char *mode)
{
    (void)mode;
    gzFile g = gzopen(fn, "rb");
    if (!g) return NULL;
    gzbuffer(g, 1 << 20);
    htsFile *f = (htsFile *)calloc(1, sizeof(htsFile));
    f->fp = g;
    return f;
}
#define sam_open(fn, mode) hts_open((fn), (mode))
static inline int hts_close(htsFile *f) { if (!f) return -1; gzclose(f->fp); free(f); return 0; }
#define sam_close(f) hts_close(f)

static inline int shim_read_exact(gzFile g, void *buf, unsigned n)
{
    unsigned got = 0;
    while (got < n) {
        int r = gzread(g, (char *)buf + got, n - got);
        if (r <= 0) break;
        got += (unsigned)r;
    }
    return (int)got;
}
static inline int32_t shim_le32(const uint8_t *p) { return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)); }

static inline bam_hdr_t *sam_hdr_read(samFile *f)
{
    uint8_t b[8];
    if (shim_read_exact(f->fp, b, 8) != 8 || memcmp(b, "BAM\1", 4) != 0) return NULL;
    bam_hdr_t *h = (bam_hdr_t *)calloc(1, sizeof(bam_hdr_t));
    h->l_text = (uint32_t)shim_le32(b + 4);
    h->text = (char *)malloc((size_t)h->l_text + 1);
    if (shim_read_exact(f->fp, h->text, h->l_text) != (int)h->l_text) { free(h->text); free(h); return NULL; }
    h->text[h->l_text] = 0;
    if (shim_read_exact(f->fp, b, 4) != 4) { free(h->text); free(h); return NULL; }
    h->n_targets = shim_le32(b);
    for (int32_t i = 0; i < h->n_targets; i++) {
        if (shim_read_exact(f->fp, b, 4) != 4) break;
        int32_t l_name = shim_le32(b);
        char *tmp = (char *)malloc((size_t)l_name + 4);
        shim_read_exact(f->fp, tmp, (unsigned)l_name + 4);
        free(tmp);
    }
    return h;
}
static inline void bam_hdr_destroy(bam_hdr_t *h) { if (h) { free(h->text); free(h); } }
static inline bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
static inline void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

static inline int sam_read1(samFile *f, bam_hdr_t *h, bam1_t *b)
{
    (void)h;
    uint8_t l[4];
    int r = shim_read_exact(f->fp, l, 4);
    if (r == 0) return -1;          /* clean EOF */
    if (r != 4) return -2;          /* truncated */
    int32_t bs = shim_le32(l);
    if (bs < 32) return -4;
    if (bs > b->m_data) { b->m_data = bs + 64; b->data = (uint8_t *)realloc(b->data, (size_t)b->m_data); }
    if (shim_read_exact(f->fp, b->data, (unsigned)bs) != bs) return -4;
    b->l_data = bs;
    const uint8_t *c = b->data;
    int l_read_name = c[8];
    int n_cigar = c[12] | (c[13] << 8);
    int32_t l_seq = shim_le32(c + 16);
    int64_t off = 32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar + ((int64_t)l_seq + 1) / 2 + l_seq;
    if (l_seq < 0 || off > bs) return -4;
    b->aux_off = (int)off;
    return bs + 4;
}

/* size of the value of an aux field whose type byte is at s[0]; returns pointer past the value or NULL */
static inline uint8_t *shim_skip_aux(uint8_t *s, uint8_t *end)
{
    if (s >= end) return NULL;
    int type = *s++;
    size_t sz;
    switch (type) {
    case 'A': case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    case 'd': sz = 8; break;
    case 'Z': case 'H':
        while (s < end && *s) s++;
        return s < end ? s + 1 : NULL;
    case 'B': {
        if (end - s < 5) return NULL;
        int sub = *s++;
        uint32_t n = (uint32_t)shim_le32(s);
        s += 4;
        size_t es;
        switch (sub) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break;
                       case 'i': case 'I': case 'f': es = 4; break; default: return NULL; }
        if ((size_t)(end - s) < es * (size_t)n) return NULL;
        return s + es * (size_t)n;
    }
    default: return NULL;
    }
    if ((size_t)(end - s) < sz) return NULL;
    return s + sz;
}
static inline uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
    uint8_t *s = b->data + b->aux_off, *end = b->data + b->l_data;
    while (s != NULL && end - s >= 3) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) {
            uint8_t *e = shim_skip_aux(s + 2, end);
            if (e == NULL) return NULL;   /* htslib: corrupt aux -> NULL */
            return s + 2;
        }
        s = shim_skip_aux(s + 2, end);
    }
    return NULL;
}
static inline char *bam_aux2Z(const uint8_t *s)
{
    int type = *s++;
    if (type == 'Z' || type == 'H') return (char *)s;
    return NULL;
}
static inline int64_t bam_aux2i(const uint8_t *s)
{
    int type = *s++;
    switch (type) {
    case 'c': return (int8_t)s[0];
    case 'C': return s[0];
    case 's': return (int16_t)(s[0] | (s[1] << 8));
    case 'S': return (uint16_t)(s[0] | (s[1] << 8));
    case 'i': return shim_le32(s);
    case 'I': return (uint32_t)shim_le32(s);
    default: return 0;
    }
}
#endif
