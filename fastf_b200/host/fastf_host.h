/* Host side of the operators, with the reference's own signatures (yuw444/fastF):
 *   int bam2db(char *bam_file, char *db_file, char *path_out, char *barcodes_file, char *features_file,
 *              float rate_cell, float rate_depth, unsigned int seed);            reference src/bam2db_ds.h:62-70
 *   freq: cell_counts(gzFile, l, u) + print_tree(node*, FILE*)                   reference src/count.h:6, src/filter.h:77
 * The per-read loop and the aggregation run on the GPU through include/fastf_gpu.h; nothing here has a CPU fallback. */
#ifndef FASTF_HOST_H
#define FASTF_HOST_H
#include <stddef.h>
#include <stdio.h>
#include <zlib.h>
extern int _umi_copies_flag;   /* reference src/bam2db_ds.h:23 */
extern int fastf_device;       /* CUDA device ordinal (env FASTF_DEVICE, default 0) */
int bam2db(char *bam_file, char *db_file, char *path_out, char *barcodes_file, char *features_file, float rate_cell, float rate_depth, unsigned int seed);
/* cell_counts + print_tree in one call: histogram of the first l+u bases of every read of R1 (BGZF or plain text), written to fp
 * in the reference's BST pre-order.  Returns 0 / 1. */
int freq_whitelist(const char *r1_path, size_t len_cellbarcode, size_t len_umi, FILE *fp);
/* reference extract_bam(bam_file, tag, type) (src/extract.h, src/extract.c:135-216): histogram of one aux tag -> ./tag_summary.csv; type 0 = string, 1 = integer.
 * Returns 0 / 1 (the reference returns void and exits on errors). */
int extract_bam(char *bam_file, const char *tag, int type);
/* reference read_bam() + print_CB_node() (src/extract.c:47-133): per cell barcode, its raw barcodes (CR) with counts, one gz line each */
int crb_write(char *bam_file, gzFile out);
#endif
