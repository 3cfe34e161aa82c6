/*
 * libdemux_io -- C ABI of the native input stage (host code, no GPU): a streaming BGZF/BAM reader and the
 * per-region counting loop of `count_snps`.
 *
 * Replaces, for the built-in read filters, the Python loop of the reference's
 * `count_call_variants_for_chromosome` (demuxalot/snp_counter.py:234-276) including
 * `compress_groups_of_molecule_reads` (:195-226), `compress_molecule_reads_to_snips` (:142-192),
 * `ChromosomeSNPLookup.get_snps` (:38-69) and `parse_read` (cellranger_specific.py:13-36,
 * BDRhapsody_specific.py:13-36).  Output records have the packed layouts of `CompressedSNPCalls`
 * (snp_counter.py:88-98): molecules 12 B (compressed_cb i4, compressed_ub i4, p_group_misaligned f4) and
 * snp_calls 13 B (molecule_index i4, snp_position i4, base_index u1, p_base_wrong f4).
 * Bound with ctypes by demuxalot_b200/counting.py.  All pointers are host pointers.
 */
#ifndef DEMUX_IO_H
#define DEMUX_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmxio_result dmxio_result;

const char* dmxio_last_error(void);

/* Counts one region task.  Reads are taken from `start_voffset` (a BGZF virtual offset at or before the first read
 * of interest, e.g. from the .bai linear index) until the reference changes or a read starts at/after `stop`;
 * reads ending at/before `start` are skipped (start/stop < 0: unbounded).  `positions`: sorted 0-based SNP positions.
 * Whitelist: n_barcodes keys concatenated in `barcode_keys` (key k = bytes [offsets[k], offsets[k+1]); with use_rg the
 * key is CB + 0x1f + RG) with their compressed_cb in `barcode_indices`.  An empty `nhits_tag` disables the
 * multi-mapping check (BD Rhapsody flavour).  Returns NULL on error (dmxio_last_error()). */
dmxio_result* dmxio_count_region(const char* bam_path, int32_t ref_id, uint64_t start_voffset, int64_t start,
                                 int64_t stop, const int64_t* positions, int64_t n_positions,
                                 const char* barcode_keys, const int64_t* barcode_key_offsets,
                                 const int32_t* barcode_indices, int64_t n_barcodes, const char* cb_tag, int32_t use_rg,
                                 const char* umi_tag, const char* nhits_tag, const char* score_tag,
                                 int32_t score_diff_max, int32_t mapq_threshold, double p_misaligned_default);

/* Per-base coverage of [start, stop) of one reference: counts[b * (stop - start) + (pos - start)] += 1 for every
 * aligned (M / = / X) base b in A, C, G, T with base quality >= quality_threshold of every mapped read that passes
 * the built-in read filter -- what `np.asarray(pysam.AlignmentFile.count_coverage(..., read_callback=lambda r:
 * parse_read(r) is not None))` returns for `detect_snps_for_chromosome` (snp_detection.py:33-42).  `counts` is
 * caller-owned int32 [4, stop - start], zero-initialised.  Returns 0, or -1 on error (dmxio_last_error()). */
int dmxio_count_coverage(const char* bam_path, int32_t ref_id, uint64_t start_voffset, int64_t start, int64_t stop,
                         const char* umi_tag, const char* nhits_tag, const char* score_tag, int32_t score_diff_max,
                         int32_t mapq_threshold, int32_t quality_threshold, int32_t* counts);

int64_t dmxio_n_molecules(const dmxio_result* r);
int64_t dmxio_n_calls(const dmxio_result* r);
int64_t dmxio_n_reads_seen(const dmxio_result* r);
/* copies 12 * n_molecules and 13 * n_calls bytes */
void dmxio_copy(const dmxio_result* r, void* molecules_out, void* calls_out);
void dmxio_free(dmxio_result* r);

#ifdef __cplusplus
}
#endif
#endif /* DEMUX_IO_H */
