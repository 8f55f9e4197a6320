/* Host writers that keep up with the device: text tables formatted in bulk and gzip-compressed by all host cores.
 * The reference writes matrix.mtx.gz / umi.tsv.gz one gzprintf() per line (src/bam2db_ds.c:516, :556); the files it produces are
 * defined by their DECOMPRESSED bytes, so here the text is cut into pieces, every piece becomes its own gzip member (RFC 1952
 * allows concatenated members; gzread / gunzip / R / scanpy read them as one stream) and the members are written in order. */
#include "fastf_host.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int fastf_host_threads(void)
{
    const char *e = getenv("FASTF_HOST_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

/* decimal text of v at p (no terminator); returns the number of characters */
unsigned fastf_fmt_u64(char *p, uint64_t v)
{
    char t[24];
    unsigned n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (unsigned i = 0; i < n; i++) p[i] = t[n - 1 - i];
    return n;
}
unsigned fastf_fmt_i64(char *p, int64_t v)
{
    if (v < 0) { *p = '-'; return 1 + fastf_fmt_u64(p + 1, (uint64_t)(-(v + 1)) + 1u); }
    return fastf_fmt_u64(p, (uint64_t)v);
}

void fastf_textbuf_reserve(fastf_textbuf *b, size_t extra)
{
    if (b->n + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : (size_t)1 << 20;
    while (nc < b->n + extra) nc *= 2;
    b->p = (char *)realloc(b->p, nc);
    if (!b->p) { fprintf(stderr, "out of host memory\n"); exit(1); }
    b->cap = nc;
}
void fastf_textbuf_free(fastf_textbuf *b) { free(b->p); memset(b, 0, sizeof *b); }

typedef struct { const char *src; size_t n; uint8_t *dst; size_t dst_n, dst_cap; int rc; } gz_piece;
typedef struct { gz_piece *pieces; size_t n_pieces; size_t next; pthread_mutex_t mu; } gz_job;

static void gz_compress_piece(gz_piece *pc)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    pc->rc = 1;
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return;
    pc->dst_cap = deflateBound(&zs, (uLong)pc->n) + 64;
    pc->dst = (uint8_t *)malloc(pc->dst_cap);
    if (pc->dst) {
        zs.next_in = (Bytef *)pc->src; zs.avail_in = (uInt)pc->n;
        zs.next_out = pc->dst; zs.avail_out = (uInt)pc->dst_cap;
        if (deflate(&zs, Z_FINISH) == Z_STREAM_END) { pc->dst_n = zs.total_out; pc->rc = 0; }
    }
    deflateEnd(&zs);
}
static void *gz_worker(void *arg)
{
    gz_job *J = (gz_job *)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        size_t i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->n_pieces) return NULL;
        gz_compress_piece(&J->pieces[i]);
    }
}

/* writes text[0, n) to path as a gzip file (one member per 4 MiB piece, compressed on all host threads).  Returns 0 / 1. */
int fastf_gz_write_parallel(const char *path, const char *text, size_t n)
{
    const size_t PIECE = (size_t)4 << 20;
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    int rc = 0;
    size_t n_pieces = n ? (n + PIECE - 1) / PIECE : 1;   /* an empty file is one empty member, like gzclose() on nothing written */
    /* bounded memory: a window of pieces at a time */
    const size_t WINDOW = 64;
    gz_piece *pieces = (gz_piece *)calloc(WINDOW, sizeof *pieces);
    int nt = fastf_host_threads();
    pthread_t th[64];
    for (size_t base = 0; base < n_pieces && !rc; base += WINDOW) {
        size_t cnt = n_pieces - base < WINDOW ? n_pieces - base : WINDOW;
        for (size_t i = 0; i < cnt; i++) {
            size_t o = (base + i) * PIECE;
            memset(&pieces[i], 0, sizeof pieces[i]);
            pieces[i].src = text + o;
            pieces[i].n = n - o < PIECE ? n - o : PIECE;
        }
        gz_job J = {pieces, cnt, 0, PTHREAD_MUTEX_INITIALIZER};
        int started = 0;
        int want = nt < (int)cnt ? nt : (int)cnt;
        for (int t = 1; t < want; t++) if (pthread_create(&th[started], NULL, gz_worker, &J) == 0) started++;
        gz_worker(&J);
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        for (size_t i = 0; i < cnt; i++) {
            if (pieces[i].rc || fwrite(pieces[i].dst, 1, pieces[i].dst_n, f) != pieces[i].dst_n) rc = 1;
            free(pieces[i].dst);
        }
    }
    free(pieces);
    if (fclose(f)) rc = 1;
    return rc;
}

/* ---- text tables: rows formatted AND compressed on all host threads ----
 * fmt(ctx, i, p) writes the line of row i at p (at most FASTF_LINE_MAX bytes) and returns its length. */
typedef struct { fastf_line_fmt fmt; void *ctx; uint64_t lo, hi; uint8_t *dst; size_t dst_n; int rc; } line_piece;
typedef struct { line_piece *pieces; size_t n_pieces; size_t next; pthread_mutex_t mu; } line_job;
static void *line_worker(void *arg)
{
    line_job *J = (line_job *)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        size_t k = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (k >= J->n_pieces) return NULL;
        line_piece *pc = &J->pieces[k];
        pc->rc = 1;
        char *text = (char *)malloc((size_t)(pc->hi - pc->lo) * FASTF_LINE_MAX + 1);
        if (!text) continue;
        char *q = text;
        for (uint64_t i = pc->lo; i < pc->hi; i++) q += pc->fmt(pc->ctx, i, q);
        gz_piece g;
        memset(&g, 0, sizeof g);
        g.src = text; g.n = (size_t)(q - text);
        gz_compress_piece(&g);
        free(text);
        pc->dst = g.dst; pc->dst_n = g.dst_n; pc->rc = g.rc;
    }
}
int fastf_gz_write_lines_parallel(const char *path, const char *head, size_t head_n, uint64_t n_rows, fastf_line_fmt fmt, void *ctx)
{
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    int rc = 0;
    {
        gz_piece g;
        memset(&g, 0, sizeof g);
        g.src = head; g.n = head_n;
        gz_compress_piece(&g);
        if (g.rc || fwrite(g.dst, 1, g.dst_n, f) != g.dst_n) rc = 1;
        free(g.dst);
    }
    const uint64_t ROWS = 1u << 18;   /* rows per gzip member: ~4 MB of text */
    const size_t WINDOW = 64;
    line_piece *pieces = (line_piece *)calloc(WINDOW, sizeof *pieces);
    const int nt = fastf_host_threads();
    pthread_t th[64];
    for (uint64_t base = 0; base < n_rows && !rc; base += ROWS * WINDOW) {
        size_t cnt = 0;
        for (uint64_t lo = base; lo < n_rows && cnt < WINDOW; lo += ROWS, cnt++) {
            memset(&pieces[cnt], 0, sizeof pieces[cnt]);
            pieces[cnt].fmt = fmt; pieces[cnt].ctx = ctx; pieces[cnt].lo = lo; pieces[cnt].hi = lo + ROWS < n_rows ? lo + ROWS : n_rows;
        }
        line_job J = {pieces, cnt, 0, PTHREAD_MUTEX_INITIALIZER};
        int started = 0;
        const int want = nt < (int)cnt ? nt : (int)cnt;
        for (int t = 1; t < want; t++) if (pthread_create(&th[started], NULL, line_worker, &J) == 0) started++;
        line_worker(&J);
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        for (size_t k = 0; k < cnt; k++) {
            if (pieces[k].rc || fwrite(pieces[k].dst, 1, pieces[k].dst_n, f) != pieces[k].dst_n) rc = 1;
            free(pieces[k].dst);
        }
    }
    free(pieces);
    if (fclose(f)) rc = 1;
    return rc;
}

/* ---- file -> pinned memory with several readers (a single fread() tops out far below what PCIe takes) ---- */
typedef struct { int fd; char *dst; off_t off; size_t n; ssize_t got; } rd_part;
static void *rd_worker(void *arg)
{
    rd_part *p = (rd_part *)arg;
    size_t done = 0;
    while (done < p->n) {
        ssize_t r = pread(p->fd, p->dst + done, p->n - done, p->off + (off_t)done);
        if (r < 0) { p->got = -1; return NULL; }
        if (r == 0) break;
        done += (size_t)r;
    }
    p->got = (ssize_t)done;
    return NULL;
}
/* reads up to n bytes at file offset off into dst; returns the bytes read (short only at end of file) or -1 */
ssize_t fastf_pread_parallel(int fd, void *dst, size_t n, off_t off)
{
    int nt = fastf_host_threads();
    if (nt > 8) nt = 8;
    const size_t MIN_PART = (size_t)8 << 20;
    if ((size_t)nt > n / MIN_PART) nt = (int)(n / MIN_PART);
    if (nt < 1) nt = 1;
    rd_part parts[8];
    pthread_t th[8];
    size_t per = (n / (size_t)nt + 4095) & ~(size_t)4095;
    int used = 0;
    for (size_t o = 0; o < n; o += per, used++) {
        parts[used].fd = fd; parts[used].dst = (char *)dst + o; parts[used].off = off + (off_t)o; parts[used].n = n - o < per ? n - o : per; parts[used].got = 0;
    }
    int started = 0;
    for (int t = 1; t < used; t++) if (pthread_create(&th[started], NULL, rd_worker, &parts[t]) == 0) started++; else rd_worker(&parts[t]);
    rd_worker(&parts[0]);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    ssize_t total = 0;
    for (int t = 0; t < used; t++) {
        if (parts[t].got < 0) return -1;
        total += parts[t].got;
        if ((size_t)parts[t].got < parts[t].n) break;   /* end of file inside this part */
    }
    return total;
}
