/*
 * libdemux_b200 -- C ABI of the B200-native likelihood / EM core for demuxalot.
 *
 * The reference (arogozhnikov/demuxalot v0.4.3) is pure Python and has no FFI seam; its boundary for
 * this path is `demuxalot.Demultiplexer` (demuxalot/demux.py:24-392).  Each entry point below replaces
 * one numpy stage of that class and cites it.  The Python host (`demuxalot_b200/demultiplexer.py`) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is a DEVICE pointer into caller-owned memory
 *    unless the parameter name starts with `h_` (host);
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream unless stated;
 *  - return value: 0 = OK, negative = error; text via dmx_last_error() (thread-local);
 *  - nothing throws across the ABI; the library keeps no global state, scratch memory is passed in
 *    by the caller (`*_workspace_bytes` queries), so calls on distinct streams/devices are independent;
 *  - table / posterior matrices are row-major float32 with an explicit leading dimension (`ld*`, in
 *    elements).
 */
#ifndef DEMUX_B200_H
#define DEMUX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMX_ABI_VERSION 4

/* E-step arithmetic flavours (see DESIGN.md "E-step") */
#define DMX_ESTEP_EXACT 0 /* per-term float32 argument roundings + logf of demux.py:261, float64 accumulation */
#define DMX_ESTEP_FAST 1  /* a = fma(P, 1-e, e'), float32 products of 8, 16 or 32 row factors with exact exponent
                             bookkeeping, one lg2 per pair and work item, float64 combination */
#define DMX_ESTEP_AUTO 2  /* EXACT for the singlet-only E-step (doublet_prior == 0: the cases where FAST missed the 1e-6
                             posterior bar; 1.3-2.4 x slower there), FAST whenever there are doublet columns */

/* ---- boundary smoke ------------------------------------------------------------------------------- */
int dmx_abi_version(void);
const char* dmx_last_error(void);
/* fills sm_count, cc_major, cc_minor, l2_bytes, total_mem_bytes for `device` */
int dmx_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes,
                    int64_t* total_mem_bytes);

/* ---- (a2) variant matching: demux.py:334-358 --------------------------------------------------------
 * Unpacks one chromosome's packed records (snp_counter.py:88-98: snp_calls 13 B, molecules 12 B),
 * builds key = chrom_id << 40 | pos << 8 | base and binary-searches it in the sorted genotype keys.
 * Writes, for call k, out_variant[k] (variant id or -1), out_cb[k] (molecules[mol].compressed_cb) and
 * out_e[k] (p_base_wrong, bit pattern preserved).
 * `molecule_stride` = 12: molecules_packed holds the 12-byte records; 4: it holds only their compressed_cb column
 * as a plain int32 array (what dmx_host_gather_cb produces: a third of the bytes on the wire).
 */
int dmx_unpack_match_calls(const uint8_t* snp_calls_packed, int64_t n_calls,
                           const uint8_t* molecules_packed, int64_t n_molecules, int32_t molecule_stride,
                           int64_t chrom_id,
                           const int64_t* geno_keys_sorted, const int32_t* geno_vids_sorted, int64_t n_variants,
                           int32_t* out_variant, int32_t* out_cb, float* out_e, void* stream);

/* HOST function (both pointers are host memory): copies the compressed_cb column of n_molecules packed 12-byte
 * molecule records (snp_counter.py:77-86) into a contiguous int32 array with n_threads threads, so that only the
 * field the hot path reads (demux.py:352) is uploaded.  Pure data movement; no reference counterpart. */
int dmx_host_gather_cb(const uint8_t* h_molecules_packed, int64_t n_molecules, int32_t* h_out_cb, int32_t n_threads);

/* ---- (a3) group + UMI-combine: demux.py:276-300, 362-363, 381 ---------------------------------------
 * Input: molecule-level calls (variant or -1, barcode, p_base_wrong) in original call order.
 * Output (capacity n_calls each, the first *h_n_rows entries are valid):
 *   rows in the reference's order, ascending (variant_id, compressed_cb)  ["CSC", used by the M-step]:
 *     csc_variant, csc_cb, csc_e (= ordered float32 product, no flush-to-zero), csc_count (group size),
 *     variant_offsets int64 [n_variants + 1]
 *   the same rows stably re-sorted by barcode                              ["CSR", used by the E-step]:
 *     csr_variant, csr_e, csr_row (index of the row in CSC order), barcode_offsets int64 [n_barcodes + 1]
 *   n_mol_per_variant int64 [n_variants] = matched molecule-level calls per variant (demux.py:381); may be
 *     NULL when the data prior is not needed (predict_posteriors)
 * Only calls whose barcode lies in [barcode_lo, barcode_hi) become rows (barcode sharding across GPUs; pass
 * 0, n_barcodes for everything); n_mol_per_variant always counts every matched call.
 * Synchronises `stream` once to return *h_n_rows and *h_n_matched on the host.
 */
int64_t dmx_build_rows_workspace_bytes(int64_t n_calls, int64_t n_variants, int64_t n_barcodes);
int dmx_build_rows(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                   int64_t n_variants, int64_t n_barcodes, int64_t barcode_lo, int64_t barcode_hi,
                   void* workspace, int64_t workspace_bytes,
                   int32_t* csc_variant, int32_t* csc_cb, float* csc_e, int32_t* csc_count,
                   int64_t* variant_offsets,
                   int32_t* csr_variant, float* csr_e, int32_t* csr_row, int64_t* barcode_offsets,
                   int64_t* n_mol_per_variant,
                   int64_t* h_n_rows, int64_t* h_n_matched, void* stream);

/* Launch schedule for the barcode-parallel E-step: order[k] = k-th barcode by descending row count, so the
 * deepest barcodes start first and the tail of the grid is made of short ones (no reference counterpart; the
 * results do not depend on it). */
int64_t dmx_barcode_schedule_workspace_bytes(int64_t n_barcodes);
int dmx_barcode_schedule(const int64_t* barcode_offsets, int64_t n_barcodes, int32_t* order, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* ---- (a4) regularised betas: demux.py:367-390 --------------------------------------------------------
 * out[v, g] = raw[v, g] + float32((1 + [n_mol] n_mol[v] / (sum_snp n_mol + 100)
 *                                   + rowsum(raw)[v] / (sum_snp rowsum + 100)) * default_prior)
 * rowsum is numpy's float32 pairwise sum; the rest is float64.  n_mol_per_variant may be NULL
 * (predict_posteriors).  snp_offsets/snp_variants: CSR SNP -> variants, ascending variant id.
 * scratch: float32 [n_variants].
 */
int dmx_prior_betas(const float* raw_betas, int64_t ld_raw, int64_t n_variants, int32_t n_genotypes,
                    const int32_t* snp_offsets, const int32_t* snp_variants, int64_t n_snps,
                    const int64_t* n_mol_per_variant, double default_prior,
                    float* scratch_rowsum, float* out_betas, int64_t ld_out, void* stream);

/* ---- (a6) probability table: demux.py:267-274 --------------------------------------------------------
 * b = betas (+ addition, float32 add, demux.py:90);  den[s, g] = float64 sum over the SNP's variants;
 * P[v, g] = clip(float32(double(b) / max(den, 1e-7)), clip_lo, clip_hi); columns g in [G, ld_table) = 1.
 */
int dmx_probs_from_betas(const float* betas, int64_t ld_betas, const float* addition, int64_t ld_addition,
                         int64_t n_variants, int32_t n_genotypes,
                         const int32_t* snp_offsets, const int32_t* snp_variants, int64_t n_snps,
                         float clip_lo, float clip_hi, float* table, int64_t ld_table, void* stream);

/* ---- (a7-a10) E-step: demux.py:246-265, 158-191, 97-101 ----------------------------------------------
 * logits[b, c] = float32(pen[c] + sum_{rows r of b} log(p_c[v_r] (1 - e_r) + max(e_r, 1e-4))), columns:
 * G singlets then (doublet_prior != 0) the G(G-1)/2 pairs i < j, i-major.  Optional prior_logits [B, C]
 * (float64, combined as numpy does: float32(float64(logit) + prior)) is added after the sum (demux.py:97-99).
 * Then a row softmax (float32).
 * Outputs (each may be NULL): logits [B, ld_logits], posteriors [B, ld_post], singlet posteriors
 * [B, ld_singlet] (first G columns of the softmax; the only part the M-step reads, demux.py:115).
 * `table` is the output of dmx_probs_from_betas with ld_table a multiple of 4.
 * workspace (query below): float32 [n_barcodes * C] when `logits` is NULL, plus float64 [n_items * C] partial sums
 * when a plan with multi-segment barcodes (n_items > n_barcodes) is given.
 * `table_floor`: a lower bound of the table entries (the clip_lo given to dmx_probs_from_betas), or 0 if
 * unknown; the FAST flavour uses it to decide how many row factors it may multiply between exponent flushes.
 * With a plan (dmx_estep_plan) and 25..32 or 57..64 genotypes the row softmax of every barcode that is a single work
 * item is computed by the E-step kernel itself (no logits round trip through memory); `logits` is then only written
 * when it is not NULL.  The outputs are the same either way.
 */
int64_t dmx_estep_workspace_bytes(int64_t n_barcodes, int32_t n_genotypes, double doublet_prior, int64_t n_items,
                                  int32_t need_logits_scratch);
int dmx_estep(const int64_t* barcode_offsets, const int32_t* barcode_order /* dmx_barcode_schedule, or NULL */,
              const int32_t* csr_variant, const float* csr_e, int64_t n_barcodes,
              const float* table, int64_t ld_table, int32_t n_genotypes,
              double doublet_prior, const double* prior_logits, int64_t ld_prior,
              float* logits, int64_t ld_logits, float* posteriors, int64_t ld_post,
              float* singlet_posteriors, int64_t ld_singlet,
              void* workspace, int64_t workspace_bytes, int32_t flavour, float table_floor,
              const int32_t* seg_prefix, const int32_t* item_slot, int64_t n_items, int32_t seg_rows /* dmx_estep_plan,
              or NULL, NULL, 0, 0 */, void* stream);

/* Work items of the warp-per-item E-step kernels (FAST flavour; dmx_estep_plan_supported tells whether a
 * configuration is served by them: doublet_prior != 0 with 1..40, 49..56 or 65..256 genotypes, and
 * doublet_prior == 0 with up to 8 genotypes): a
 * barcode with more than seg_rows rows is cut into ceil(rows / seg_rows) segments so that no single warp carries a
 * deep barcode alone; the segments' float64 partial sums are added in segment order (deterministic).  Outputs, by
 * schedule slot s (barcode = barcode_order[s], or s when barcode_order is NULL): seg_prefix int32 [n_barcodes + 1]
 * (first item of slot s; seg_prefix[n_barcodes] = number of items) and item_slot int32 [item_capacity]
 * (item -> slot); item_capacity >= n_barcodes + n_rows / seg_rows always suffices.  No reference counterpart; the
 * results do not depend on the plan beyond float64 regrouping.  Synchronises `stream` once to return *h_n_items.
 * workspace: dmx_estep_plan_workspace_bytes(n_barcodes). */
int dmx_estep_plan_supported(int32_t n_genotypes, double doublet_prior, int32_t flavour);
int64_t dmx_estep_plan_workspace_bytes(int64_t n_barcodes);
int dmx_estep_plan(const int64_t* barcode_offsets, const int32_t* barcode_order, int64_t n_barcodes, int32_t seg_rows,
                   int32_t* seg_prefix, int32_t* item_slot, int64_t item_capacity, void* workspace,
                   int64_t workspace_bytes, int64_t* h_n_items, void* stream);

/* Introspection (HOST only, no device work): the assignment of the pair triangle to the lanes of the strip kernel
 * (csrc/estep_pairs_strip.cu; 25..32 genotypes: n_blocks = 4, 8 slots; 57..64: n_blocks = 8, 32 slots).  h_slots receives 8
 * int16 per slot: float offsets p0, p1, p2 (pairs of strips 0..2), q1 (block of strips 0, 1), d (block of strip 2 and
 * unit 3), P1, P2 (pairs of unit 3; P2 < 0: unit 3 is a plain strip), 0.  Lets a CPU test check that every genotype pair is
 * produced exactly once. */
int dmx_estep_strip_layout(int32_t n_blocks, int16_t* h_slots, int32_t* h_n_slots);

/* row softmax only (scipy.special.softmax(x, axis=-1), demux.py:101,152); outputs as in dmx_estep */
int dmx_softmax_rows(const float* logits, int64_t ld_logits, int64_t n_rows, int32_t n_cols,
                     float* posteriors, int64_t ld_post, float* singlet_posteriors, int64_t ld_singlet,
                     int32_t n_singlets, void* stream);

/* ---- (a11) M-step: demux.py:113-118 -------------------------------------------------------------------
 * addition[v, g] = float32(sum_{rows r of v, ascending} (post[cb_r, g] (1 - e_r)) ^ power), g < G,
 * float32 terms accumulated in float64.  Variants in [variant_lo, variant_hi) are processed so the caller
 * can tile the variant range and overlap the all-reduce of finished tiles.  When `addition64` is not
 * NULL the unrounded float64 sums are stored there as well (multi-GPU partials).
 */
int dmx_mstep(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
              const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
              float* addition, int64_t ld_addition, double* addition64, int64_t ld_addition64,
              int64_t variant_lo, int64_t variant_hi, void* stream);

/* Planned M-step: same result as dmx_mstep, scheduled by rows per variant.  dmx_mstep_plan classifies the variants
 * once per packed data set (the row counts do not change between EM iterations): variants with at most 128 rows are
 * walked by groups of lanes in strict row order (np.bincount's order), longer ones by a warp each, and variants with
 * more than 4096 rows are cut into chunks whose float64 partial sums are added in chunk order, so no work item is
 * longer than 4096 rows.  `plan`: device buffer of dmx_mstep_plan_bytes(n_rows) bytes, opaque to the caller;
 * h_counts[3] (host) receives n_medium, n_heavy_variants, n_heavy_items, to be passed back to dmx_mstep_planned
 * together with a float64 scratch of n_heavy_items * n_genotypes elements.  dmx_mstep_plan synchronises `stream`. */
int64_t dmx_mstep_plan_bytes(int64_t n_rows);
int dmx_mstep_plan(const int64_t* variant_offsets, int64_t n_variants, int64_t n_rows, void* plan, int64_t plan_bytes,
                   int64_t* h_counts, void* stream);
int dmx_mstep_planned(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
                      const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
                      float* addition, int64_t ld_addition, double* addition64, int64_t ld_addition64,
                      int64_t variant_lo, int64_t variant_hi, const void* plan, int64_t n_rows, int64_t n_medium,
                      int64_t n_heavy_variants, int64_t n_heavy_items, double* heavy_scratch, void* stream);

/* ---- per-(barcode, SNP) regularised likelihood: demux.py:204-244 (`Demultiplexer.aggregate_on_snps = True`) ----
 * The reference's experimental alternative to (a9): matched molecule-level calls are grouped by
 * (compressed_cb, snp_id) (FeatureLookup, utils.py:207-265); per group and column the float32 logs
 * log(p_c[variant] + p_base_wrong) are summed in float64, divided by count ** compensation, log-softmaxed in
 * float32, mixed with a uniform 0.01 / C "bad SNP" mass, log-softmaxed in float64 and summed per barcode.
 * Logits and posteriors of this branch are float64, as in the reference (np.logaddexp against a float64 scalar).
 *
 * dmx_build_snp_groups: inputs are the outputs of dmx_unpack_match_calls over all chromosomes (variant -1 =
 * unmatched) and variant2snp (genotypes.py:56-66) on the device; calls of barcodes outside [cb_lo, cb_hi) are
 * skipped.  Outputs: the kept calls in group order -- ascending (barcode, SNP), original call order inside a
 * group (stable sort) -- as grouped_variant / grouped_e [n_calls], group_offsets [n_calls + 1] (first n_groups + 1
 * entries valid) and barcode_group_offsets [n_barcodes + 1] (groups of barcodes < b).  Synchronises `stream`;
 * h_n_matched / h_n_groups (host) receive the counts.  n_calls < 2^31. */
int64_t dmx_snp_groups_workspace_bytes(int64_t n_calls);
int dmx_build_snp_groups(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                         const int32_t* variant2snp, int64_t n_snps, int64_t n_barcodes, int64_t cb_lo, int64_t cb_hi,
                         void* workspace, int64_t workspace_bytes, int32_t* grouped_variant, float* grouped_e,
                         int64_t* group_offsets, int64_t* barcode_group_offsets, int64_t* h_n_matched,
                         int64_t* h_n_groups, void* stream);
/* logits[b, c] float64 [n_barcodes, C], C as in dmx_estep; `table` is the output of dmx_probs_from_betas.
 * One warp per barcode, groups in order, no atomics: deterministic. */
int64_t dmx_snp_logits_workspace_bytes(int64_t n_barcodes, int32_t n_cols);
int dmx_snp_logits(const int64_t* barcode_group_offsets, const int64_t* group_offsets, const int32_t* grouped_variant,
                   const float* grouped_e, int64_t n_barcodes, const float* table, int64_t ld_table,
                   int32_t n_genotypes, double doublet_prior, double compensation, double* logits, int64_t ld_logits,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* float64 row softmax (demux.py:101,152 on this branch's logits).  prior_logits (may be NULL) is first added to
 * `logits` in place (demux.py:97-99); posteriors (float64, may be NULL) and singlet_posteriors (float32 [n_rows,
 * ld_singlet], the M-step's input, may be NULL) as in dmx_softmax_rows. */
int dmx_softmax_rows_f64(double* logits, int64_t ld_logits, const double* prior_logits, int64_t ld_prior, int64_t n_rows,
                         int32_t n_cols, double* posteriors, int64_t ld_post, float* singlet_posteriors,
                         int64_t ld_singlet, int32_t n_singlets, void* stream);

/* ---- sharded pack (multi-GPU) --------------------------------------------------------------
 * The reference packs on one host (demux.py:302-392; its only split is the per-chromosome loop at :334-358).  Across
 * GPUs every rank uploads and unpacks a contiguous SLICE of each chromosome's calls; the matched calls then move to
 * the rank that owns their barcode (contiguous barcode ranges balanced by matched calls), so that a rank sorts and
 * keeps only its shard.  The exchange itself is an all-to-all the host issues (torch.distributed / NCCL).
 *
 * dmx_barcode_histogram: histogram[b] += matched calls of barcode b (int64 [n_barcodes], accumulated: zero it first).
 * dmx_route_calls: stable partition of the matched calls by owner rank; cuts (device, int64 [world + 1]) are the
 * barcode range boundaries, cuts[0] = 0, cuts[world] = n_barcodes.  Outputs (capacity n_calls): the kept calls grouped
 * by destination rank, original order inside a group, with out_cb LOCAL to the owner's range (cb - cuts[dest]).
 * h_counts (host, int64 [world + 1]): calls per destination; the last entry counts matched calls whose compressed_cb
 * lies outside [0, n_barcodes) (an input error the caller must raise on every rank).  Synchronises `stream`.
 */
int dmx_barcode_histogram(const int32_t* call_variant, const int32_t* call_cb, int64_t n_calls, int64_t n_variants,
                          int64_t n_barcodes, int64_t* histogram, void* stream);
int64_t dmx_route_calls_workspace_bytes(int64_t n_calls);
int dmx_route_calls(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                    int64_t n_variants, int64_t n_barcodes, const int64_t* cuts, int32_t world, void* workspace,
                    int64_t workspace_bytes, int32_t* out_variant, int32_t* out_cb, float* out_e, int64_t* h_counts,
                    void* stream);

/* ---- (f) M-step + cross-GPU sum: demux.py:113-118 on barcode shards ---------------------------------------------------
 * One communicator per process (one process per GPU).  dmx_comm_unique_id (rank 0; h_id128 = 128 host bytes, to be
 * broadcast by the host's own means) and dmx_comm_init wrap ncclGetUniqueId / ncclCommInitRank; NCCL is bound with
 * dlopen at the first call, there is no link-time dependency.
 *
 * dmx_mstep_allreduce = dmx_mstep[_planned] over n_tiles tiles of the variant range, each finished tile handed to NCCL
 * on the communicator's own stream while the kernel of the next tile runs; `stream` waits for the last collective
 * before the call's successors run.  On return (stream order) `addition` holds the float32 GLOBAL sums on every rank.
 *   wire_float64 != 0: reduce-scatter of float64 partials (partial64 [v_pad, G]), this rank's slice rounded to
 *     float32 once (slice64: float64 scratch of ceil(v_pad / n_tiles + world) * G / world elements), all-gather of
 *     the float32 slices;  wire_float64 == 0: in-place float32 all-reduce (partial64 / slice64 may be NULL).
 * `addition` (and partial64) must hold v_pad = dmx_mstep_allreduce_padded_variants(n_variants, world) rows of G
 * contiguous floats, rows [n_variants, v_pad) zeroed by the caller once.  plan == NULL: unplanned M-step.
 */
/* Cross-GPU sum of the M-step partials over peer memory (NVLink / NVSwitch), one kernel, no library collective.
 * h_partials[r] / h_outputs[r] (HOST arrays of `world` DEVICE pointers): rank r's float32 partial table and output table
 * as mapped into THIS process (peer mappings: CUDA IPC / fabric handles, e.g. torch symmetric memory).  This rank sums
 * slice `rank` of the n_elements floats over all partials (rank order, float64, one rounding) and stores the result
 * into slice `rank` of every output: reduce-scatter + all-gather in one pass, the same bits on every rank.
 * The caller brackets the call with two cross-rank barriers on `stream`: all partials complete before it, all
 * outputs complete after it.  n_elements must be a multiple of 4 * world; world <= 16. */
int dmx_peer_sum_f32(const void* const* h_partials, void* const* h_outputs, int32_t rank, int32_t world,
                     int64_t n_elements, void* stream);
int dmx_comm_unique_id(uint8_t* h_id128);
int dmx_comm_init(const uint8_t* h_id128, int32_t rank, int32_t world, void** h_comm);
int dmx_comm_destroy(void* comm);
int64_t dmx_mstep_allreduce_padded_variants(int64_t n_variants, int32_t world);
int dmx_mstep_allreduce(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
                        const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
                        float* addition, double* partial64, double* slice64, int64_t n_variants, const void* plan,
                        int64_t n_rows, int64_t n_medium, int64_t n_heavy_variants, int64_t n_heavy_items,
                        double* heavy_scratch, void* comm, int32_t n_tiles, int32_t wire_float64, void* stream);

/* float32(out) = float32(in64) elementwise over a [rows, cols] matrix (after an all-reduce of float64 partials) */
int dmx_round_f64_to_f32(const double* in64, int64_t ld_in, float* out, int64_t ld_out, int64_t n_rows,
                         int32_t n_cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMUX_B200_H */
