/* Synthetic 10x-v3-style BAM / R1 FASTQ generators (data tooling for bench + tests). */
#ifndef FASTF_SYNTH_H
#define FASTF_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define FASTF_SYNTH_MAX_UMI 16
typedef struct {
    uint64_t n_reads;
    uint64_t n_molecules;      /* 0 = n_reads; each read picks a molecule uniformly -> (cell, gene, UMI) */
    uint64_t seed;
    uint32_t n_cells;          /* size of the barcode list (BAM) / number of true barcodes (FASTQ) */
    uint32_t n_genes;
    uint32_t umi_len;          /* 12 (10x v3) */
    int32_t zlevel;            /* zlib level, 6 */
    double p_cb_in_list;       /* 0.96 */
    double p_cb_not_in_list;   /* 0.02 ; remainder: CB (and UB) tag absent */
    double p_gx25;             /* 0.85 GX + xf:25 */
    double p_gx17;             /* 0.05 GX + xf:17 ; remainder: no GX, xf:0 */
    double p_umi_n;            /* fraction of reads whose UB (BAM) / barcode (FASTQ) gets an 'N' */
    double p_bc_error;         /* FASTQ: fraction of reads with one substitution in the barcode (0.05) */
} fastf_synth_params;
typedef struct {
    uint64_t n_reads, n_blocks, inflated_bytes, compressed_bytes;
} fastf_synth_stats;
void fastf_synth_defaults(fastf_synth_params *p);
int fastf_synth_bam(const fastf_synth_params *p, int nthreads, uint8_t **out, size_t *out_n, fastf_synth_stats *st);
int fastf_synth_fastq(const fastf_synth_params *p, int nthreads, uint8_t **out, size_t *out_n, fastf_synth_stats *st);
int fastf_synth_barcodes(const fastf_synth_params *p, char **out, size_t *out_n);
int fastf_synth_features(const fastf_synth_params *p, char **out, size_t *out_n);
void fastf_synth_free(void *p);
#ifdef __cplusplus
}
#endif
#endif
