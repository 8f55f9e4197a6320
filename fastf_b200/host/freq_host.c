/* freq: cell_counts() + print_tree() of the reference (src/count.c:3-21, src/filter.c:105-148) on top of fastf_freq_gpu.
 * The device returns the pure-ACGT keys (2-bit packed, ascending = strcmp order) with count and first-occurrence ordinal, and the
 * raw bytes of the few reads whose key holds another byte; those are merged here by byte order, then everything is printed in
 * the pre-order of the reference's BST (fastf_cartesian_preorder). */
#include "fastf_host.h"
#include "../../include/fastf_gpu.h"
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

typedef struct { char key[36]; uint32_t len; uint32_t first; uint64_t count; } exc_t;
static int exc_cmp(const void *a, const void *b)
{
    const exc_t *x = (const exc_t *)a, *y = (const exc_t *)b;
    int c = strcmp(x->key, y->key);
    if (c) return c;
    return x->first < y->first ? -1 : x->first > y->first;
}
static void decode_key(uint64_t k, uint32_t klen, char *out)
{
    for (uint32_t i = 0; i < klen; i++) out[i] = "ACGT"[(k >> (2 * (klen - 1 - i))) & 3];
    out[klen] = 0;
}

int freq_whitelist(const char *r1_path, size_t len_cellbarcode, size_t len_umi, FILE *fp)
{
    int rc = 1;
    fastf_ctx *ctx = NULL;
    fastf_freq_result res;
    memset(&res, 0, sizeof res);
    void *buf = NULL;
    exc_t *exc = NULL;
    uint32_t *first_all = NULL, *src = NULL;
    uint64_t *order = NULL;
    const uint32_t klen = (uint32_t)(len_cellbarcode + len_umi);
    const int f = open(r1_path, O_RDONLY);
    struct stat sb;
    if (f < 0 || fstat(f, &sb)) { fprintf(stderr, "Cannot open file %s \n", r1_path); if (f >= 0) close(f); return 1; }
    long long n = (long long)sb.st_size;
    if (fastf_ctx_create(fastf_device, &ctx)) { fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(NULL)); goto done; }
    if (fastf_host_alloc(ctx, (size_t)(n > 0 ? n : 1), &buf)) { fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(ctx)); goto done; }
    if (n > 0 && fastf_pread_parallel(f, buf, (size_t)n, 0) != (ssize_t)n) { fprintf(stderr, "Cannot read file %s \n", r1_path); goto done; }   /* several readers into pinned memory */
    if (fastf_freq_gpu(ctx, buf, (size_t)n, klen, 0, &res)) { fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(ctx)); goto done; }

    /* exceptional reads -> distinct keys sorted by bytes */
    uint64_t ne = res.n_exceptions, counted = ne;
    for (uint64_t i = 0; i < res.n_keys; i++) counted += res.count[i];
    uint64_t extra_empty = res.n_reads > counted ? res.n_reads - counted : 0;   /* trailing record without a sequence line: empty key */
    exc = (exc_t *)calloc(ne + 1, sizeof(exc_t));
    for (uint64_t i = 0; i < ne; i++) {
        const uint8_t *raw = res.exc_bytes + i * res.exc_stride;
        uint32_t l = 0;
        /* key = bytes of the sequence line up to klen, stopping after the first '\n' (the NUL of gzgets' buffer follows it) or at NUL */
        while (l < klen && raw[l] != 0) { exc[i].key[l] = (char)raw[l]; l++; if (raw[l - 1] == '\n') break; }
        exc[i].key[l] = 0; exc[i].len = l; exc[i].first = res.exc_ordinal[i]; exc[i].count = 1;
    }
    uint64_t nex = ne;
    if (extra_empty) { exc[nex].key[0] = 0; exc[nex].len = 0; exc[nex].first = (uint32_t)counted; exc[nex].count = extra_empty; nex++; }
    qsort(exc, nex, sizeof(exc_t), exc_cmp);
    uint64_t nd = 0;
    for (uint64_t i = 0; i < nex; i++) {
        if (nd && !strcmp(exc[nd - 1].key, exc[i].key)) exc[nd - 1].count += exc[i].count;
        else exc[nd++] = exc[i];
    }
    /* merge with the device keys */
    uint64_t total = res.n_keys + nd;
    first_all = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
    src = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));      /* bit 31 set: index into exc[] */
    order = (uint64_t *)malloc(sizeof(uint64_t) * (total ? total : 1));
    if (total >= 0x7fffffffull) { fprintf(stderr, "\x1b[31mError:\x1b[0m too many distinct keys\n"); goto done; }
    {
        uint64_t a = 0, b = 0, o = 0;
        char dk[36];
        while (a < res.n_keys || b < nd) {
            int take_dev;
            if (a >= res.n_keys) take_dev = 0;
            else if (b >= nd) take_dev = 1;
            else { decode_key(res.key[a], klen, dk); take_dev = strcmp(dk, exc[b].key) < 0; }
            if (take_dev) { first_all[o] = res.first[a]; src[o] = (uint32_t)a; a++; }
            else { first_all[o] = exc[b].first; src[o] = 0x80000000u | (uint32_t)b; b++; }
            o++;
        }
    }
    if (fastf_cartesian_preorder(first_all, total, order)) { fprintf(stderr, "\x1b[31mError:\x1b[0m pre-order failed\n"); goto done; }
    {
        /* "key,count\n" lines formatted into a block buffer: 10^8 fprintf calls would cost more than the device job */
        const size_t BLK = (size_t)4 << 20;
        char *out = (char *)malloc(BLK + 128);
        size_t o = 0;
        if (!out) { fprintf(stderr, "\x1b[31mError:\x1b[0m out of host memory\n"); goto done; }
        for (uint64_t i = 0; i < total; i++) {
            uint32_t s = src[order[i]];
            if (s & 0x80000000u) {
                const exc_t *e = &exc[s & 0x7fffffffu];
                memcpy(out + o, e->key, e->len); o += e->len;
                out[o++] = ',';
                o += fastf_fmt_u64(out + o, e->count);
            } else {
                decode_key(res.key[s], klen, out + o); o += klen;
                out[o++] = ',';
                o += fastf_fmt_u64(out + o, res.count[s]);
            }
            out[o++] = '\n';
            if (o >= BLK) { fwrite(out, 1, o, fp); o = 0; }
        }
        if (o) fwrite(out, 1, o, fp);
        free(out);
    }
    rc = 0;
done:
    close(f);
    free(exc); free(first_all); free(src); free(order);
    fastf_freq_result_free(&res);
    if (ctx) { fastf_host_free(ctx, buf); fastf_ctx_destroy(ctx); }
    return rc;
}
