/* Bulk loader for the big tables of the bam2db database (`umi`: one row per kept read, reference src/bam2db_ds.c:351-352,421-435;
 * `mtx`: src/bam2db_ds.c:480-483; `numi`: :527-530).  The reference pushes every row through sqlite3_step(); at 10^8 rows that is
 * minutes of host time behind a device job of seconds.  Here the rows are laid out directly as a table b-tree in the SQLite file
 * format (https://sqlite.org/fileformat2.html: leaf pages 0x0d, interior pages 0x05, records with minimal-width integer serial
 * types, rowids 1..n in row order) and appended to the database file that sqlite itself created (schema, small tables).  The
 * result is an ordinary database: `PRAGMA integrity_check` is "ok" and every SELECT gives what the row-by-row inserts would.
 *
 * Usage: create the (empty) table through the sqlite API, look up its root page, CLOSE the connection, then
 * fastf_sqlite_bulk_begin -> _row ... -> _end.  Several tables are filled one after the other. */
#include "fastf_host.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct fastf_sqlite_bulk {
    FILE *f;
    uint32_t page_size, usable, root, n_pages;   /* n_pages: pages of the file so far (next new page = n_pages + 1) */
    uint32_t change_counter;
    uint64_t rowid;                              /* last rowid written */
    /* leaf under construction */
    uint8_t *leaf;
    uint32_t leaf_cells, leaf_top;               /* cells so far; cell content grows down from leaf_top */
    /* finished leaves wait in a write buffer; their (page number, largest rowid) feed the interior levels */
    uint8_t *wbuf;
    size_t wbuf_pages, wbuf_cap_pages;
    uint32_t wbuf_first_page;
    uint32_t *kid_page; uint64_t *kid_key; size_t n_kids, cap_kids;
    int failed;
};

static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static void put_be16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static uint32_t get_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static uint32_t get_be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }

/* SQLite varint: big-endian base 128, at most 9 bytes (the 9th carries 8 bits) */
static unsigned put_varint(uint8_t *p, uint64_t v)
{
    if (v <= 0x7f) { p[0] = (uint8_t)v; return 1; }
    if (v <= 0x3fff) { p[0] = (uint8_t)((v >> 7) | 0x80); p[1] = (uint8_t)(v & 0x7f); return 2; }
    if (v >> 56) {
        p[8] = (uint8_t)v;
        v >>= 8;
        for (int i = 7; i >= 0; i--) { p[i] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; }
        return 9;
    }
    uint8_t tmp[10];
    unsigned n = 0;
    do { tmp[n++] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; } while (v);
    tmp[0] &= 0x7f;
    for (unsigned i = 0; i < n; i++) p[i] = tmp[n - 1 - i];
    return n;
}

/* integer column: serial type and big-endian body of the smallest width that holds v (schema format 4: 0 and 1 have no body) */
static unsigned put_int_col(int64_t v, uint8_t *type, uint8_t *body)
{
    if (v == 0) { *type = 8; return 0; }
    if (v == 1) { *type = 9; return 0; }
    unsigned n;
    if (v >= -128 && v <= 127) { *type = 1; n = 1; }
    else if (v >= -32768 && v <= 32767) { *type = 2; n = 2; }
    else if (v >= -8388608 && v <= 8388607) { *type = 3; n = 3; }
    else if (v >= -2147483648LL && v <= 2147483647LL) { *type = 4; n = 4; }
    else if (v >= -140737488355328LL && v <= 140737488355327LL) { *type = 5; n = 6; }
    else { *type = 6; n = 8; }
    for (unsigned i = 0; i < n; i++) body[i] = (uint8_t)((uint64_t)v >> (8 * (n - 1 - i)));
    return n;
}

unsigned fastf_sqlite_int_col(int64_t v, uint8_t *type, uint8_t *body) { return put_int_col(v, type, body); }

static void leaf_reset(fastf_sqlite_bulk *b)
{
    memset(b->leaf, 0, b->page_size);
    b->leaf_cells = 0;
    b->leaf_top = b->usable;
}

static int flush_wbuf(fastf_sqlite_bulk *b)
{
    if (!b->wbuf_pages) return 0;
    if (fseeko(b->f, (off_t)(b->wbuf_first_page - 1) * b->page_size, SEEK_SET) || fwrite(b->wbuf, b->page_size, b->wbuf_pages, b->f) != b->wbuf_pages) return b->failed = 1;
    b->wbuf_first_page += (uint32_t)b->wbuf_pages;
    b->wbuf_pages = 0;
    return 0;
}

static int push_kid(fastf_sqlite_bulk *b, uint32_t page, uint64_t key)
{
    if (b->n_kids == b->cap_kids) {
        b->cap_kids = b->cap_kids ? b->cap_kids * 2 : 4096;
        b->kid_page = (uint32_t *)realloc(b->kid_page, b->cap_kids * sizeof(uint32_t));
        b->kid_key = (uint64_t *)realloc(b->kid_key, b->cap_kids * sizeof(uint64_t));
        if (!b->kid_page || !b->kid_key) return b->failed = 1;
    }
    b->kid_page[b->n_kids] = page;
    b->kid_key[b->n_kids++] = key;
    return 0;
}

/* the leaf under construction is complete: header, then into the write buffer as the next new page of the file */
static int leaf_finish(fastf_sqlite_bulk *b)
{
    uint8_t *p = b->leaf;
    p[0] = 0x0d;
    put_be16(p + 3, b->leaf_cells);
    put_be16(p + 5, b->leaf_top == 65536 ? 0 : b->leaf_top);
    if (b->wbuf_pages == b->wbuf_cap_pages && flush_wbuf(b)) return 1;
    memcpy(b->wbuf + b->wbuf_pages * b->page_size, p, b->page_size);
    b->wbuf_pages++;
    if (push_kid(b, ++b->n_pages, b->rowid)) return 1;
    leaf_reset(b);
    return 0;
}

fastf_sqlite_bulk *fastf_sqlite_bulk_begin(const char *db_file, unsigned root_page)
{
    fastf_sqlite_bulk *b = (fastf_sqlite_bulk *)calloc(1, sizeof *b);
    uint8_t h[100];
    if (!b) return NULL;
    b->f = fopen(db_file, "r+b");
    if (!b->f || fread(h, 1, 100, b->f) != 100 || memcmp(h, "SQLite format 3", 16)) goto fail;
    b->page_size = get_be16(h + 16);
    if (b->page_size == 1) b->page_size = 65536;
    b->usable = b->page_size - h[20];
    b->change_counter = get_be32(h + 24);
    b->n_pages = get_be32(h + 28);
    b->root = root_page;
    if (b->page_size < 512 || root_page < 2 || root_page > b->n_pages || get_be32(h + 92) != b->change_counter) goto fail;   /* in-header size must be valid */
    {
        /* the table must be empty: its root is a leaf without cells */
        uint8_t rh[8];
        if (fseeko(b->f, (off_t)(root_page - 1) * b->page_size, SEEK_SET) || fread(rh, 1, 8, b->f) != 8 || rh[0] != 0x0d || get_be16(rh + 3) != 0) goto fail;
    }
    b->leaf = (uint8_t *)malloc(b->page_size);
    b->wbuf_cap_pages = ((size_t)16 << 20) / b->page_size;
    b->wbuf = (uint8_t *)malloc(b->wbuf_cap_pages * b->page_size);
    if (!b->leaf || !b->wbuf) goto fail;
    b->wbuf_first_page = b->n_pages + 1;
    leaf_reset(b);
    return b;
fail:
    if (b->f) fclose(b->f);
    free(b->leaf); free(b->wbuf); free(b);
    return NULL;
}

/* the record (ncol serial types < 128, nb body bytes) becomes the next cell, rowid = previous + 1 */
static int append_record(fastf_sqlite_bulk *b, const uint8_t *types, unsigned ncol, const uint8_t *body, unsigned nb);

/* one row: n_int integer columns, then (has_tail) one more column that is a blob of tail_len bytes, or NULL when tail == NULL */
int fastf_sqlite_bulk_row(fastf_sqlite_bulk *b, const int64_t *ints, unsigned n_int, int has_tail, const void *tail, unsigned tail_len)
{
    uint8_t types[16], body[96];
    unsigned nb = 0;
    if (b->failed || n_int > 8 || tail_len > 16) return b->failed = 1;
    for (unsigned i = 0; i < n_int; i++) nb += put_int_col(ints[i], &types[i], body + nb);
    unsigned ncol = n_int;
    if (has_tail) {
        types[ncol++] = tail ? (uint8_t)(12 + 2 * tail_len) : 0;   /* all serial types here are < 128: one varint byte each */
        if (tail) { memcpy(body + nb, tail, tail_len); nb += tail_len; }
    }
    return append_record(b, types, ncol, body, nb);
}

/* one row of table numi: two integers, a blob or NULL, one integer (reference src/bam2db_ds.c:527-530) */
int fastf_sqlite_bulk_row4(fastf_sqlite_bulk *b, const int64_t head[2], const void *blob, unsigned blob_len, int64_t last)
{
    uint8_t types[4], body[64];
    unsigned nb = 0;
    if (b->failed || blob_len > 16) return b->failed = 1;
    nb += put_int_col(head[0], &types[0], body + nb);
    nb += put_int_col(head[1], &types[1], body + nb);
    types[2] = blob ? (uint8_t)(12 + 2 * blob_len) : 0;
    if (blob) { memcpy(body + nb, blob, blob_len); nb += blob_len; }
    nb += put_int_col(last, &types[3], body + nb);
    return append_record(b, types, 4, body, nb);
}

static int append_record(fastf_sqlite_bulk *b, const uint8_t *types, unsigned ncol, const uint8_t *body, unsigned nb)
{
    uint8_t cell[160];
    const unsigned hdr = 1 + ncol, payload = hdr + nb;
    unsigned n = put_varint(cell, payload);
    n += put_varint(cell + n, b->rowid + 1);
    cell[n++] = (uint8_t)hdr;
    memcpy(cell + n, types, ncol); n += ncol;
    memcpy(cell + n, body, nb); n += nb;
    /* room: 8-byte page header + 2-byte pointer per cell + the cell bodies */
    if (8 + 2 * (b->leaf_cells + 1) + n > b->leaf_top && leaf_finish(b)) return 1;
    b->leaf_top -= n;
    memcpy(b->leaf + b->leaf_top, cell, n);
    put_be16(b->leaf + 8 + 2 * b->leaf_cells, b->leaf_top);
    b->leaf_cells++;
    b->rowid++;
    return 0;
}

/* ---- many rows at once, encoded on all host threads ----
 * Table leaf pages hold no page numbers, so threads lay out the leaves of disjoint row ranges independently (a range starts a
 * fresh leaf); the pages are then written in row order and get their page numbers as they go. */
typedef struct {
    const fastf_sqlite_bulk *b;
    fastf_row_encoder enc; void *ctx;
    uint64_t lo, hi, rowid0;        /* rows [lo, hi) of this call; row lo gets rowid rowid0 + 1 */
    uint8_t *pages; size_t n_pages, cap_pages;
    uint64_t *last_rowid;           /* per page */
    int failed;
} leaf_range;

static void *leaf_range_worker(void *arg)
{
    leaf_range *R = (leaf_range *)arg;
    const uint32_t ps = R->b->page_size, usable = R->b->usable;
    uint8_t *pg = NULL;
    uint32_t cells = 0, top = usable;
    uint64_t rowid = R->rowid0;
    for (uint64_t i = R->lo; i < R->hi; i++) {
        uint8_t cell[192], types[16], body[128];
        unsigned ncol = 0, nb = 0;
        R->enc(R->ctx, i, types, &ncol, body, &nb);
        const unsigned hdr = 1 + ncol, payload = hdr + nb;
        unsigned n = put_varint(cell, payload);
        n += put_varint(cell + n, rowid + 1);
        cell[n++] = (uint8_t)hdr;
        memcpy(cell + n, types, ncol); n += ncol;
        memcpy(cell + n, body, nb); n += nb;
        if (!pg || 8 + 2 * (cells + 1) + n > top) {
            if (pg) { pg[0] = 0x0d; put_be16(pg + 3, cells); put_be16(pg + 5, top == 65536 ? 0 : top); R->last_rowid[R->n_pages - 1] = rowid; }
            if (R->n_pages == R->cap_pages) {
                R->cap_pages = R->cap_pages ? R->cap_pages * 2 : 256;
                R->pages = (uint8_t *)realloc(R->pages, R->cap_pages * ps);
                R->last_rowid = (uint64_t *)realloc(R->last_rowid, R->cap_pages * sizeof(uint64_t));
                if (!R->pages || !R->last_rowid) { R->failed = 1; return NULL; }
            }
            pg = R->pages + R->n_pages++ * ps;
            memset(pg, 0, ps);
            cells = 0; top = usable;
        }
        top -= n;
        memcpy(pg + top, cell, n);
        put_be16(pg + 8 + 2 * cells, top);
        cells++;
        rowid++;
    }
    if (pg) { pg[0] = 0x0d; put_be16(pg + 3, cells); put_be16(pg + 5, top == 65536 ? 0 : top); R->last_rowid[R->n_pages - 1] = rowid; }
    return NULL;
}

int fastf_sqlite_bulk_rows_parallel(fastf_sqlite_bulk *b, uint64_t n, fastf_row_encoder enc, void *ctx)
{
    if (b->failed) return 1;
    if (n == 0) return 0;
    int nt = fastf_host_threads();
    const uint64_t SMALL = 20000;
    if (n < SMALL * 2 || nt == 1) {
        /* not worth threads (and small tables keep one tight leaf chain): the sequential path */
        for (uint64_t i = 0; i < n; i++) {
            uint8_t types[16], body[128];
            unsigned ncol = 0, nb = 0;
            enc(ctx, i, types, &ncol, body, &nb);
            if (append_record(b, types, ncol, body, nb)) return 1;
        }
        return 0;
    }
    if (b->leaf_cells && leaf_finish(b)) return 1;
    if (flush_wbuf(b)) return 1;
    const uint64_t BATCH = (uint64_t)nt * (1u << 20);   /* rows per round: bounds the pages held in memory */
    leaf_range R[64];
    pthread_t th[64];
    for (uint64_t base = 0; base < n && !b->failed; base += BATCH) {
        const uint64_t cnt = n - base < BATCH ? n - base : BATCH;
        int used = (int)((cnt + SMALL - 1) / SMALL) < nt ? (int)((cnt + SMALL - 1) / SMALL) : nt;
        const uint64_t per = (cnt + (uint64_t)used - 1) / (uint64_t)used;
        for (int t = 0; t < used; t++) {
            memset(&R[t], 0, sizeof R[t]);
            R[t].b = b; R[t].enc = enc; R[t].ctx = ctx;
            R[t].lo = base + per * (uint64_t)t;
            R[t].hi = R[t].lo + per < base + cnt ? R[t].lo + per : base + cnt;
            if (R[t].lo > R[t].hi) R[t].lo = R[t].hi;
            R[t].rowid0 = b->rowid + (R[t].lo - base);
        }
        int started = 0;
        for (int t = 1; t < used; t++) { if (pthread_create(&th[started], NULL, leaf_range_worker, &R[t]) == 0) started++; else leaf_range_worker(&R[t]); }
        leaf_range_worker(&R[0]);
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        for (int t = 0; t < used; t++) {
            if (R[t].failed) b->failed = 1;
            if (!b->failed && R[t].n_pages) {
                if (fseeko(b->f, (off_t)b->n_pages * b->page_size, SEEK_SET) || fwrite(R[t].pages, b->page_size, R[t].n_pages, b->f) != R[t].n_pages) b->failed = 1;
                for (size_t k = 0; k < R[t].n_pages && !b->failed; k++) push_kid(b, ++b->n_pages, R[t].last_rowid[k]);
            }
            free(R[t].pages); free(R[t].last_rowid);
        }
        b->rowid += cnt;
        b->wbuf_first_page = b->n_pages + 1;
    }
    return b->failed;
}

int fastf_sqlite_bulk_end(fastf_sqlite_bulk *b)
{
    int rc = 1;
    uint8_t *page = NULL;
    if (b->failed) goto done;
    page = (uint8_t *)malloc(b->page_size);
    if (!page) goto done;
    if (b->n_kids == 0) {
        /* everything fits the root leaf (also the empty table) */
        b->leaf[0] = 0x0d;
        put_be16(b->leaf + 3, b->leaf_cells);
        put_be16(b->leaf + 5, b->leaf_top == 65536 ? 0 : b->leaf_top);
        if (fseeko(b->f, (off_t)(b->root - 1) * b->page_size, SEEK_SET) || fwrite(b->leaf, b->page_size, 1, b->f) != 1) goto done;
    } else {
        if (b->leaf_cells && leaf_finish(b)) goto done;
        if (flush_wbuf(b)) goto done;
        /* interior levels bottom-up until one page is left: that one is written over the table's root page */
        const unsigned per_page = (b->usable - 12) / (2 + 4 + 9);   /* cell = left child + rowid varint (<= 9 bytes) */
        while (1) {
            const size_t n = b->n_kids;
            const size_t n_parents = (n + per_page) / (per_page + 1);   /* a page with k cells has k + 1 children */
            const int top = n_parents == 1;
            size_t out = 0;
            for (size_t i = 0; i < n;) {
                size_t take = n - i < (size_t)per_page + 1 ? n - i : (size_t)per_page + 1;
                if (n - i - take == 1) take--;   /* never leave a parent with a single child and no cell */
                memset(page, 0, b->page_size);
                page[0] = 0x05;
                unsigned topo = b->usable;
                for (size_t c = 0; c + 1 < take; c++) {
                    uint8_t cell[16];
                    put_be32(cell, b->kid_page[i + c]);
                    unsigned cn = 4 + put_varint(cell + 4, b->kid_key[i + c]);
                    topo -= cn;
                    memcpy(page + topo, cell, cn);
                    put_be16(page + 12 + 2 * c, topo);
                }
                put_be16(page + 3, (uint32_t)(take - 1));
                put_be16(page + 5, topo == 65536 ? 0 : topo);
                put_be32(page + 8, b->kid_page[i + take - 1]);
                const uint32_t pno = top ? b->root : ++b->n_pages;
                if (fseeko(b->f, (off_t)(pno - 1) * b->page_size, SEEK_SET) || fwrite(page, b->page_size, 1, b->f) != 1) goto done;
                b->kid_page[out] = pno;
                b->kid_key[out++] = b->kid_key[i + take - 1];
                i += take;
            }
            b->n_kids = out;
            if (top) break;
        }
    }
    {
        /* header: new size, bumped change counter, "version valid for" = change counter so that the in-header size is trusted */
        uint8_t h[100];
        if (fseeko(b->f, 0, SEEK_SET) || fread(h, 1, 100, b->f) != 100) goto done;
        put_be32(h + 24, b->change_counter + 1);
        put_be32(h + 28, b->n_pages);
        put_be32(h + 92, b->change_counter + 1);
        if (fseeko(b->f, 0, SEEK_SET) || fwrite(h, 1, 100, b->f) != 100) goto done;
    }
    rc = 0;
done:
    if (fclose(b->f)) rc = 1;
    free(page); free(b->leaf); free(b->wbuf); free(b->kid_page); free(b->kid_key); free(b);
    return rc;
}
