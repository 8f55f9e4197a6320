/* Host side of `fastF crb` and `fastF extract` (reference src/extract.c, src/main.c:231-286,364-402): the record loop runs on
 * the GPU (fastf_taghist_gpu: inflate, record framing, aux walk, grouping, first occurrences); what is left here is the order
 * in which the reference's unbalanced BSTs print their nodes -- pre-order = Cartesian tree of the first occurrences over the
 * bytewise-sorted values (fastf_cartesian_preorder) -- and the output files.  No CPU path: without a CUDA device both fail. */
#include "fastf_host.h"
#include "../../include/fastf_gpu.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { const char *a, *b; uint32_t a_len, b_len, first, count; } grp_t;

static int cmp_bytes(const char *x, uint32_t nx, const char *y, uint32_t ny)   /* strcmp order of NUL-free strings */
{
    uint32_t n = nx < ny ? nx : ny;
    int c = n ? memcmp(x, y, n) : 0;
    if (c) return c;
    return nx < ny ? -1 : nx > ny;
}
static int grp_cmp(const void *p, const void *q)
{
    const grp_t *x = (const grp_t *)p, *y = (const grp_t *)q;
    int c = cmp_bytes(x->a, x->a_len, y->a, y->a_len);
    if (c) return c;
    return cmp_bytes(x->b, x->b_len, y->b, y->b_len);
}

typedef struct {
    fastf_ctx *ctx;
    void *buf;
    fastf_taghist_result res;
    grp_t *g;         /* groups sorted by (A, B) bytes */
    char *itext;      /* integer mode: the printed values */
} taghist_t;

static void taghist_close(taghist_t *t)
{
    free(t->g);
    free(t->itext);
    fastf_taghist_result_free(&t->res);
    if (t->ctx) { fastf_host_free(t->ctx, t->buf); fastf_ctx_destroy(t->ctx); }
    memset(t, 0, sizeof *t);
}

/* reads the file into pinned memory and runs the device histogram */
static int taghist_open(taghist_t *t, const char *bam_file, const char *tag_a, uint32_t mode, const char *tag_b)
{
    memset(t, 0, sizeof *t);
    FILE *f = fopen(bam_file, "rb");
    if (!f) { fprintf(stderr, "ERROR: Cannot open bam file %s\n", bam_file); return 1; }
    fseek(f, 0, SEEK_END);
    long long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    int rc = 1;
    if (fastf_ctx_create(fastf_device, &t->ctx)) { fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(NULL)); goto done; }
    if (fastf_host_alloc(t->ctx, (size_t)(n > 0 ? n : 1), &t->buf)) { fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(t->ctx)); goto done; }
    if (n > 0 && fread(t->buf, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "ERROR: Cannot read bam file %s\n", bam_file); goto done; }
    if (fastf_taghist_gpu(t->ctx, t->buf, (size_t)n, tag_a, mode, tag_b, 0, &t->res)) {
        /* records that cross BGZF blocks (htsjdk / STAR writers): once more with record starts guessed and verified per block */
        if (!strstr(fastf_last_error(t->ctx), "record-straddles-bgzf-block") || fastf_taghist_gpu(t->ctx, t->buf, (size_t)n, tag_a, mode, tag_b, FASTF_BAM_STRADDLE, &t->res)) {
            fprintf(stderr, "\x1b[31mError:\x1b[0m %s\n", fastf_last_error(t->ctx));
            goto done;
        }
    }
    const uint64_t ng = t->res.n_groups;
    t->g = (grp_t *)calloc(ng + 1, sizeof(grp_t));
    if (mode == FASTF_TAG_INT) t->itext = (char *)calloc(ng + 1, 16);
    if (!t->g || (mode == FASTF_TAG_INT && !t->itext)) { fprintf(stderr, "\x1b[31mError:\x1b[0m out of memory\n"); goto done; }
    for (uint64_t i = 0; i < ng; i++) {
        grp_t *g = &t->g[i];
        g->first = t->res.first[i];
        g->count = t->res.count[i];
        if (mode == FASTF_TAG_INT) {
            g->a = t->itext + 16 * i;
            g->a_len = (uint32_t)snprintf(t->itext + 16 * i, 16, "%d", (int)t->res.ivalue[i]);   /* sprintf(tag_str, "%d", bam_aux2i(..)), src/extract.c:190 */
            g->b = g->a + g->a_len;
        } else {
            g->a = t->res.strings + t->res.a_off[i];
            g->a_len = t->res.a_len[i];
            g->b = g->a + g->a_len;
            g->b_len = t->res.b_len[i];
        }
    }
    qsort(t->g, ng, sizeof(grp_t), grp_cmp);
    rc = 0;
done:
    fclose(f);
    if (rc) taghist_close(t);
    return rc;
}

/* order[k] = index (into a run of n sorted values with first occurrences t[]) of the k-th node print_tree visits */
static int preorder(const uint32_t *t, uint64_t n, uint64_t *order)
{
    if (n == 0) return 0;
    return fastf_cartesian_preorder(t, n, order);
}

/* reference: void extract_bam(char *bam_file, const char *tag, int type)  (src/extract.c:135-216); writes ./tag_summary.csv */
int extract_bam(char *bam_file, const char *tag, int type)
{
    if (!tag || strlen(tag) < 2) { fprintf(stderr, "\x1b[31mError:\x1b[0m --tag takes a two-character tag.\n"); return 1; }
    taghist_t t;
    if (taghist_open(&t, bam_file, tag, type ? FASTF_TAG_INT : FASTF_TAG_STRING, NULL)) return 1;
    int rc = 1;
    const uint64_t ng = t.res.n_groups;
    uint32_t *first = (uint32_t *)malloc(sizeof(uint32_t) * (ng + 1));
    uint64_t *order = (uint64_t *)malloc(sizeof(uint64_t) * (ng + 1));
    FILE *fp = fopen("tag_summary.csv", "w");
    if (!first || !order || !fp) { fprintf(stderr, "\x1b[31mError:\x1b[0m cannot write tag_summary.csv\n"); goto done; }
    for (uint64_t i = 0; i < ng; i++) first[i] = t.g[i].first;
    if (preorder(first, ng, order)) { fprintf(stderr, "\x1b[31mError:\x1b[0m pre-order failed\n"); goto done; }
    for (uint64_t k = 0; k < ng; k++) {
        const grp_t *g = &t.g[order[k]];
        fwrite(g->a, 1, g->a_len, fp);
        fprintf(fp, ",%ld\n", (long)g->count);                                 /* print_tree, src/filter.c:139-148 */
    }
    printf("Processed all %lu reads\n", (unsigned long)(2 * t.res.n_records));   /* total_count++ twice per record, src/extract.c:162,164 */
    printf("Valid reads: %lu\n", (unsigned long)t.res.n_hits);
    rc = 0;
done:
    if (fp) fclose(fp);
    free(first); free(order);
    taghist_close(&t);
    return rc;
}

/* reference: read_bam(bam) + print_CB_node(tree, gz) (src/extract.c:47-133), called from cmd_crb (src/main.c:272-278) */
int crb_write(char *bam_file, gzFile out)
{
    taghist_t t;
    if (taghist_open(&t, bam_file, "CB", FASTF_TAG_STRING, "CR")) return 1;
    int rc = 1;
    const uint64_t ng = t.res.n_groups;
    /* runs of equal CB inside the (CB, CR)-sorted groups */
    uint64_t ncb = 0;
    uint64_t *cb_start = (uint64_t *)malloc(sizeof(uint64_t) * (ng + 2));
    uint32_t *first = (uint32_t *)malloc(sizeof(uint32_t) * (ng + 1));
    uint64_t *order = (uint64_t *)malloc(sizeof(uint64_t) * (ng + 1)), *order2 = (uint64_t *)malloc(sizeof(uint64_t) * (ng + 1));
    if (!cb_start || !first || !order || !order2) { fprintf(stderr, "\x1b[31mError:\x1b[0m out of memory\n"); goto done; }
    for (uint64_t i = 0; i < ng; i++)
        if (i == 0 || cmp_bytes(t.g[i].a, t.g[i].a_len, t.g[i - 1].a, t.g[i - 1].a_len)) cb_start[ncb++] = i;
    cb_start[ncb] = ng;
    for (uint64_t c = 0; c < ncb; c++) {                       /* a CB node is created by its first read */
        uint32_t m = 0xffffffffu;
        for (uint64_t i = cb_start[c]; i < cb_start[c + 1]; i++) if (t.g[i].first < m) m = t.g[i].first;
        first[c] = m;
    }
    if (preorder(first, ncb, order)) { fprintf(stderr, "\x1b[31mError:\x1b[0m pre-order failed\n"); goto done; }
    printf("Processed all %lu reads\n", (unsigned long)t.res.n_records);   /* src/extract.c:123 */
    printf("Writing to file...\n");                                          /* src/main.c:274 */
    for (uint64_t k = 0; k < ncb; k++) {
        const uint64_t lo = cb_start[order[k]], n = cb_start[order[k] + 1] - lo;
        gzwrite(out, t.g[lo].a, t.g[lo].a_len);
        gzputc(out, ';');                                      /* gzprintf(fp, "%s;", root->CB) */
        for (uint64_t i = 0; i < n; i++) first[i] = t.g[lo + i].first;   /* (first[] of the CB level is no longer needed) */
        if (preorder(first, n, order2)) { fprintf(stderr, "\x1b[31mError:\x1b[0m pre-order failed\n"); goto done; }
        for (uint64_t i = 0; i < n; i++) {
            const grp_t *g = &t.g[lo + order2[i]];
            gzwrite(out, g->b, g->b_len);
            gzprintf(out, ",%ld;", (long)g->count);            /* print_tree_same_row, src/filter.c:161-170 */
        }
        gzputc(out, '\n');
    }
    rc = 0;
done:
    free(cb_start); free(first); free(order); free(order2);
    taghist_close(&t);
    return rc;
}
