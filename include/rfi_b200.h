/*
 * rfi_b200.h -- C ABI of librfi_b200.so, the B200 (sm_100a) implementation of the
 * rfi_toolbox preprocessing / evaluation hot path.
 *
 * The reference (preshanth/rfi_toolbox 0.2.0) is pure Python and has no FFI layer of
 * its own; each entry point below names the reference function
 * (file:line under rfi_toolbox/) whose arithmetic it replaces.  The Python host layer in
 * rfi_toolbox_b200/ binds these with ctypes and mirrors the reference's public API.
 *
 * Conventions
 *   - plain C: device pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - no allocation inside, re-entrant per stream; no host synchronisation, with one exception:
 *     rfi_tile_stats for P = 256 / 512 / 1024 reads one int back (how many groups need the
 *     generic select) before it returns;
 *   - every call returns 0 on success or a negative RFI_E_* code; rfi_last_error_string()
 *     gives the text for the calling thread's last failure;
 *   - all device buffers are owned by the caller.
 */
#ifndef RFI_B200_H
#define RFI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFI_B200_ABI_VERSION 6

/* status codes */
#define RFI_OK 0
#define RFI_E_INVALID (-1)     /* bad argument / unsupported combination */
#define RFI_E_UNSUPPORTED (-2) /* legal in the reference, not built on this path yet */
#define RFI_E_CUDA (-3)        /* CUDA runtime error (text in rfi_last_error_string) */

/* element types of the visibility cube (preprocessor.py:175-196 accepts any of them) */
#define RFI_F32 0
#define RFI_F64 1
#define RFI_C64 2
#define RFI_C128 3

/* stretch (preprocessor.py:672-706) */
#define RFI_STRETCH_NONE 0
#define RFI_STRETCH_SQRT 1
#define RFI_STRETCH_LOG10 2

/* where the labels come from (preprocessor.py:315-334) */
#define RFI_FLAGS_CUSTOM 0    /* caller's flags, rotated + tiled like the data */
#define RFI_FLAGS_MAD 1       /* median +- sigma * MAD of the processed patch */
#define RFI_FLAGS_INFERENCE 2 /* all-zero labels */

/* Plan of one create_dataset call over a cube (B, Npol, C, T), C-order.  Three code paths
 * (rfi_plan_path), same results:
 *   fast     P = 128, C and T multiples of P: one CTA per tile, tile resident on chip;
 *   big      P = 256 / 512 / 1024, C and T multiples of P, float32 / complex64 through the real
 *            branch: sub-tile + group launches, 4-CTA cluster writer;
 *   generic  any other size, C or T not a multiple of P (the reference zero-pads bottom/right after
 *            the rotation, preprocessor.py:527-550), waterfalls no larger than the patch (patchify
 *            skipped, preprocessor.py:261-269), float64 at P >= 256: segmented multi-pass radix
 *            select over global memory.  (A caller can keep a padded geometry on the fast / big
 *            kernels by running one rotations = 1 plan per view over rfi_rotate_pad copies -- the
 *            Python layer does.)
 * All of them take the workspace rfi_plan_workspace_bytes() reports. */
typedef struct rfi_plan {
    int32_t dtype;       /* RFI_F32 .. RFI_C128 */
    int32_t magnitude;   /* complex input only: 1 = take |z| on load and run the real branch
                            (reference fed np.abs(data), ms_loader.py:556-561);
                            0 = complex branch (preprocessor.py:285-292, 562-606) */
    int64_t n_waterfalls; /* B * Npol */
    int64_t channels;    /* C (rows of a waterfall)  */
    int64_t times;       /* T (columns of a waterfall) */
    int32_t patch;       /* P */
    int32_t rotations;   /* effective views per waterfall: 1, 2 or 4 (preprocessor.py:413-446) */
    int32_t stretch;     /* RFI_STRETCH_* */
    int32_t norm_before; /* preprocessor.py:295-297 */
    int32_t norm_after;  /* preprocessor.py:309-311 */
    int32_t flag_mode;   /* RFI_FLAGS_* */
    double sigma;        /* flag_sigma, applied in the data's precision */
} rfi_plan_t;

/* Per ORIGINAL tile statistics produced by rfi_tile_stats and consumed by
 * rfi_write_patches (they are rotation invariant, SURVEY.md section 8 identity (i)).
 * Stored as doubles; for float32/complex64 input each holds an exactly representable
 * float32 value.  88 bytes. */
typedef struct rfi_tile_stat {
    double median_before; /* nanmedian of the raw tile           (preprocessor.py:663) */
    double inf_fill;      /* MAD of the finite stretched values  (preprocessor.py:697-702) */
    double median_after;  /* nanmedian after the stretch         (preprocessor.py:311) */
    double centre;        /* nanmedian of the processed tile     (preprocessor.py:737) */
    double mad;           /* MAD of the processed tile           (preprocessor.py:736) */
    double thr_lo;        /* centre - mad * sigma                (preprocessor.py:740) */
    double thr_hi;        /* centre + mad * sigma                (preprocessor.py:739) */
    int32_t n_valid;      /* non-NaN samples of the raw tile */
    int32_t n_inf;        /* +-inf samples after the stretch */
    int32_t n_flagged;    /* samples flagged (MAD mode) or non-zero custom flags */
    int32_t route;        /* RFI_TILE_* bits: how the tile was measured, what phase 2 may use */
    double raw_lo;        /* RFI_TILE_RAW_THRESHOLDS: a sample is flagged iff raw < raw_lo ... */
    double raw_hi;        /* ... or raw > raw_hi, raw = the loaded value (|z| for magnitude) */
} rfi_tile_stat_t;

/* rfi_tile_stat_t.route */
#define RFI_TILE_RAW_THRESHOLDS 1 /* all samples >= +0 and finite after the stretch: every stage is
                                     monotone, so thr_lo / thr_hi were mapped back EXACTLY to the raw
                                     domain (raw_lo / raw_hi) and phase 2 labels without normalising */
#define RFI_TILE_RAW_FILL 8       /* with RFI_TILE_RAW_THRESHOLDS, LOG10 tiles whose smallest samples stretch to -inf
                                   * (exact-zero bandpass rows): a sample <= raw_zero IS the fill value `inf_fill`
                                   * (flagged iff inf_fill lies outside [thr_lo, thr_hi]); raw_zero travels in
                                   * `median_after`, which such a tile does not use (no second normalisation) */
#define RFI_TILE_GENERAL 2        /* measured by the general kernel (negative, infinite or inf-filled
                                     samples, or a sampled bracket that missed) */

/* Host arithmetic only.
 * rfi_plan_num_tiles    entries of the statistics array = statistic groups: one per original
 *                       tile when C and T are multiples of P (the R rotated patches share
 *                       their statistics), one per output patch (canonical order) when the
 *                       reference pads, because the pad follows the flip;
 * rfi_plan_num_patches  patches before blank removal = R * B * Npol * ceil(C/P) * ceil(T/P)
 *                       (R * B * Npol when patchify is skipped);
 * rfi_plan_workspace_bytes  device scratch the two phases share.  Fast path, float32 / complex64:
 *                       64 KB per tile (the tile's keys, thread-private in phase 1), laid out tile
 *                       by tile, so a call over a sub-range of the waterfalls may be given the
 *                       matching sub-range of the buffer; for complex64 with magnitude = 1 phase 2
 *                       reads the exact magnitudes back from it (4 B / px instead of 8 B / px and
 *                       |z| again; NULL there = read the cube); 0 for float64 / complex128. */
#define RFI_PATH_FAST 0     /* P = 128, dims multiples of 128: one CTA per tile, tile on chip */
#define RFI_PATH_BIG 1      /* P = 256 / 512 / 1024, dims multiples of P, float32 arithmetic, real branch */
#define RFI_PATH_GENERIC 2  /* everything else */
int rfi_plan_path(const rfi_plan_t* plan); /* which of the three the plan takes (-1: bad plan) */
int64_t rfi_plan_num_tiles(const rfi_plan_t* plan);
int64_t rfi_plan_num_patches(const rfi_plan_t* plan);
size_t rfi_plan_workspace_bytes(const rfi_plan_t* plan);

/* Phase 1 -- replaces _normalize / _apply_stretch statistics / _generate_mad_flags
 * (preprocessor.py:646-745) and the `any()` of _remove_blank_patches (:749).
 *   data   device, cube in the plan's dtype
 *   flags  device, uint8/bool cube of the same shape (RFI_FLAGS_CUSTOM) or NULL
 *   stats  device, rfi_plan_num_tiles() entries, written
 *   workspace device, rfi_plan_workspace_bytes() bytes (NULL when that is 0); the same buffer,
 *          untouched in between, must be passed to rfi_write_patches
 * P = 128: one CTA per original tile, exact order statistics by sampled brackets on the raw keys
 * (register-resident radix select as the in-CTA fallback).  P = 256 / 512 / 1024: the same
 * selection split over sub-tile and group launches.  Everything else: segmented radix select. */
int rfi_tile_stats(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                   rfi_tile_stat_t* stats, void* workspace, void* stream);

/* Phase 2 -- replaces _apply_rotations, patchify, _normalize, _apply_stretch, flag
 * application, _extract_channels_from_{real,complex}, ImageNet normalisation and the
 * compaction + shuffle gathers (preprocessor.py:22-42, 413-446, 562-783).
 *   dest_slot device int64[rfi_plan_num_patches()], canonical patch index -> output slot,
 *             -1 = dropped (blank / beyond num_patches)
 *   images    device float32 (N, P, P, 3), labels device uint8 (N, P, P)
 *             ((N, C, T, 3) / (N, C, T) when patchify is skipped) */
int rfi_write_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                      const rfi_tile_stat_t* stats, const int64_t* dest_slot,
                      float* images, uint8_t* labels, void* workspace, void* stream);

/* Phase 1 + phase 2 in ONE launch (P = 128 on-chip path, float32 arithmetic, real branch) -- replaces
 * the same reference lines as rfi_tile_stats + rfi_write_patches (preprocessor.py:22-42, 413-446,
 * 562-783) for a caller that knows `dest_slot` BEFORE the statistics exist: inference_mode
 * (:345-353, no compaction, no shuffle), or MAD flags with the slots of the all-kept case drawn ahead
 * and the flag counts in `stats` checked afterwards (if a tile came out blank the caller completes
 * through rfi_write_patches with these `stats`, workspace NULL).  One CTA per original tile keeps the
 * tile in shared memory from the 128-bit loads to the bulk (TMA) stores of its R patches: the cube
 * is read once, nothing but the output is written (60 B / px for complex64, R = 4).
 *   rfi_plan_fusable  1 if the plan can take this path (RFI_PATH_FAST, RFI_F32 or RFI_C64 with
 *                     magnitude, RFI_FLAGS_MAD or RFI_FLAGS_INFERENCE), else 0
 *   stats     device, rfi_plan_num_tiles() entries, written (as by rfi_tile_stats)
 *   other arguments as for rfi_write_patches; no workspace */
int rfi_plan_fusable(const rfi_plan_t* plan);
int rfi_fused_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                      rfi_tile_stat_t* stats, const int64_t* dest_slot, float* images,
                      uint8_t* labels, void* stream);

/* Confusion counts -- replaces the boolean reductions of evaluation/metrics.py:36-40,
 * 63-68, 95-99, 142-147.  `elem_*` is the element size in bytes (1, 2, 4 or 8) and
 * `is_float_*` selects float semantics (x != 0, NaN counts as True) over integer ones.
 *   counts device uint64[3] = {TP, FP, FN}, ACCUMULATED into (caller zeroes). */
int rfi_confusion_counts(const void* pred, int elem_pred, int is_float_pred,
                         const void* truth, int elem_true, int is_float_true,
                         int64_t n, unsigned long long* counts, void* stream);

/* Confusion counts SUMMED OVER THE RANKS of one box in the same kernel (masks sharded by baseline,
 * SURVEY.md section 8e) -- replaces rfi_confusion_counts + an NCCL all-reduce.  Every rank calls it
 * with the same `epoch` (1, 2, 3 ... per call); the last CTA of the reduction stores the rank's
 * totals into every peer's exchange buffer over NVLink, publishes the epoch behind a system fence,
 * waits for every peer's epoch in its own buffer and writes the sums.
 *   peers   host array of `world` (<= 16) device pointers: the exchange buffers of all ranks
 *           (own: rfi_peer_alloc; others: rfi_peer_open of the handle that rank exported)
 *   counts  device uint64[4], WRITTEN: {TP, FP, FN} over all ranks, [3] != 0 if a peer never
 *           arrived (the wait gives up after ~30 s instead of hanging the GPU) */
int rfi_confusion_counts_allreduce(const void* pred, int elem_pred, int is_float_pred,
                                   const void* truth, int elem_true, int is_float_true, int64_t n,
                                   void* const* peers, int world, int rank, uint64_t epoch,
                                   unsigned long long* counts, void* stream);

/* Exchange buffers (RFI_PEER_BYTES used; allocated by the library so that the 64-byte CUDA IPC
 * handle maps exactly this buffer).  Handles travel between the ranks by any host channel
 * (torch.distributed.all_gather_object in the Python layer). */
#define RFI_PEER_BYTES 4096
int rfi_peer_alloc(void** buf, unsigned char* handle64);
int rfi_peer_open(const unsigned char* handle64, void** peer);
int rfi_peer_close(void* peer);
int rfi_peer_free(void* buf);

/* Same, one triple per consecutive segment of `seg` elements (per-pair sweep,
 * BASELINE config 4).  counts device uint64[n_seg][3], written. */
int rfi_confusion_counts_segmented(const void* pred, int elem_pred, int is_float_pred,
                                   const void* truth, int elem_true, int is_float_true,
                                   int64_t n_seg, int64_t seg, unsigned long long* counts,
                                   void* stream);

/* Robust statistics -- replaces compute_statistics (evaluation/statistics.py:16-56): stats
 * of |data| over the samples whose flag byte is zero (all samples when flags is NULL),
 * computed in the data's own precision T and widened to double on output. */
typedef struct rfi_stats {
    double mean;       /* float64-accumulated sum / count, rounded to T  (statistics.py:50) */
    double median;     /* mean of the two middle order statistics in T    (statistics.py:51) */
    double std;        /* population std (ddof 0), two-pass               (statistics.py:52) */
    double mad;        /* median(|x - median|)                            (statistics.py:10-13) */
    int64_t count;     /* unflagged samples                               (statistics.py:54) */
    int64_t n_flagged; /* non-zero flag bytes                             (statistics.py:34) */
    int64_t n_nan;     /* NaNs among the unflagged samples; if > 0 median and mad are NaN */
    double max;        /* np.max of the unflagged magnitudes (NaN if any is NaN)   (statistics.py:161) */
} rfi_stats_t;

size_t rfi_statistics_workspace_bytes(void);

/*   data      device, n samples of `dtype` (RFI_F32 .. RFI_C128)
 *   flags     device uint8[n] or NULL
 *   out       device rfi_stats_t, written (stream-ordered)
 *   workspace device, rfi_statistics_workspace_bytes() bytes */
int rfi_statistics(const void* data, int dtype, const uint8_t* flags, int64_t n,
                   rfi_stats_t* out, void* workspace, void* stream);

/* Both sets of compute_ffi in one call -- replaces compute_statistics(data) AND compute_statistics(data,
 * flags) (statistics.py:16-56, called back to back by compute_ffi, :73-74): out[0] = all samples, out[1] =
 * the samples whose flag byte is zero (== out[0] when flags is NULL).  RFI_F32 / RFI_C64: three passes over
 * the data (cube -> 4 B / px key scratch; median brackets; MAD brackets) instead of forty, exact order
 * statistics from sampled brackets; a missed bracket (probability ~1e-5) is reported as out[0].count = -1
 * and the caller repeats the call through rfi_statistics.  RFI_F64 / RFI_C128: rfi_statistics twice.
 *   data       device, 16-byte aligned;  flags  device uint8[n] (4-byte aligned) or NULL
 *   workspace  device, rfi_statistics2_workspace_bytes(dtype, n) bytes (~4.6 B per sample) */
size_t rfi_statistics2_workspace_bytes(int dtype, int64_t n);
int rfi_statistics2(const void* data, int dtype, const uint8_t* flags, int64_t n, rfi_stats_t* out,
                    void* workspace, void* stream);

/* The same two sets of statistics over a cube that is SHARDED over ranks by baseline ("f64 {n,
 * sum x, sum x^2} plus radix histograms for a global FFI"): the caller (torch.distributed in the Python layer)
 * sums what these two entry points leave between the calls.  RFI_F32 / RFI_C64 only.
 *   rfi_statistics_shard_begin   pass A over the rank's shard (keys -> workspace scratch); state = device
 *       double[8] {n, n_flagged, n_nan[0], n_nan[1], sum[0], sum[1], maxkey[0], maxkey[1]} (set 0 = all samples,
 *       set 1 = unflagged; [0..5] are summed over the ranks, [6..7] maximised)
 *   rfi_statistics_shard_count   one pass over the rank's key scratch.  mode 0: counts[s * 16 + t] = number of
 *       keys of set s below prefix[s] | (t + 1) << shift, t = 0 .. 14 (MSB-first radix select, 4 bits per pass;
 *       all-reduce SUM); mode 1: counts[s * 16] = keys <= prefix[s] (SUM), counts[s * 16 + 1] = smallest key
 *       above it (MIN).  dev_mode 1: the key of a sample is |x - centre[s]| (the MAD).  sumsq (or NULL): device
 *       double[2], sum of (x - mean[s])^2 over the rank's samples of each set.
 *   centre / mean / prefix are HOST arrays of 2; counts device uint64[32]; workspace as for rfi_statistics2. */
int rfi_statistics_shard_begin(const void* data, int dtype, const uint8_t* flags, int64_t n,
                               void* workspace, double* state, void* stream);
int rfi_statistics_shard_count(const uint8_t* flags, int64_t n, void* workspace, int mode, int dev_mode,
                               const float* centre, const float* mean, const uint32_t* prefix, int shift,
                               unsigned long long* counts, double* sumsq, void* stream);

/* Per-pair sweep (BASELINE config 4): the same statistics for every consecutive segment of `seg`
 * samples (seg <= 16384, e.g. one 128 x 128 patch) in one launch, one CTA per segment.
 *   out  device rfi_stats_t[n_seg][2]: [i][0] over all samples of segment i (statistics.py:73),
 *        [i][1] over its unflagged samples (:74), n_flagged filled in [i][1] */
int rfi_statistics_segmented(const void* data, int dtype, const uint8_t* flags, int64_t n_seg,
                             int64_t seg, rfi_stats_t* out, void* stream);

/* Fused per-pair sweep (BASELINE config 4) -- replaces, for every pair i, compute_ffi(data[i],
 * flags[i]) (statistics.py:59-97, i.e. compute_statistics before and after flagging, :16-56) AND the
 * boolean reductions behind evaluate_segmentation(flags[i], truth[i]) (metrics.py:36-40, 63-68, 95-99,
 * 142-147, 155-172) in ONE launch that reads data, flags and truth once (10 B / px for complex64).
 *   data     device, n_pairs x seg samples, RFI_F32 or RFI_C64 (|z| fused into the load); seg <= 16384
 *   flags    device uint8 / bool [n_pairs x seg], the predicted mask (non-zero = flagged), or NULL
 *   truth    device uint8 / bool [n_pairs x seg], the ground-truth mask, or NULL (counts stay 0)
 *   stats    device rfi_stats_t[n_pairs][2] as rfi_statistics_segmented writes them, or NULL
 *   results  device rfi_pair_result_t[n_pairs], or NULL
 * status: 0 = ffi valid; 1 = the reference's guard (all flagged, or NaN among the unflagged samples,
 * statistics.py:77-78: ffi = reductions = 0, flagged_fraction = 1); 2 = constant data (MAD or std of 0
 * before flagging: the reference raises ZeroDivisionError; the fields hold NaN). */
typedef struct rfi_pair_result {
    double ffi, mad_reduction, std_reduction, flagged_fraction;   /* statistics.py:80-97, float64 like Python floats */
    double iou, precision, recall, f1, dice;                      /* metrics.py:25-152 from the counts, float64 */
    uint32_t tp, fp, fn;                                          /* pred & true, pred & ~true, ~pred & true */
    int32_t status;
} rfi_pair_result_t;

int rfi_pair_sweep(const void* data, int dtype, const uint8_t* flags, const uint8_t* truth,
                   int64_t n_pairs, int64_t seg, rfi_stats_t* stats, rfi_pair_result_t* results,
                   void* stream);

/* Host helper (no CUDA): np.random.permutation(n) of NumPy's legacy MT19937 generator --
 * replaces the shuffle of preprocessor.py:758-763 at ~3 ns per element instead of ~30.
 *   mt_key  host uint32[624], the generator key  (np.random.get_state()[1]), advanced in place
 *   mt_pos  host, position in the key              (np.random.get_state()[2]), advanced in place
 *   out     host int64[n], the permutation
 * The caller writes (mt_key, mt_pos) back with np.random.set_state(), so the stream position
 * after the call is what the reference leaves behind. */
int rfi_legacy_permutation(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int64_t* out);

/* Host helper (no CUDA): blank-patch removal, shuffle and truncation (preprocessor.py:746-763,
 * :356-359) between the two phases.
 *   n_flagged    host, per statistic group (rfi_plan_num_tiles() entries), `stride_bytes` apart
 *                (pass the n_flagged field of the copied-back rfi_tile_stat_t array, stride 88)
 *   shuffle      0 = inference mode: canonical order, nothing dropped, RNG untouched
 *   mt_key/pos   as for rfi_legacy_permutation (advanced in place)
 *   num_patches  <= 0: keep all
 *   order        host int64[rfi_plan_num_patches()]: canonical index of output patch k, k < *n_out
 *   dest         host int64[rfi_plan_num_patches()]: output slot of canonical patch q, -1 = dropped */
int rfi_plan_slots(const rfi_plan_t* plan, const int32_t* n_flagged, int64_t stride_bytes, int shuffle,
                   uint32_t* mt_key, int32_t* mt_pos, int64_t num_patches, int64_t* order,
                   int64_t* dest, int64_t* n_out);

/* Device-side synthetic visibilities -- the step BEFORE the hot path: replaces the per-pixel work of
 * SyntheticDataGenerator._generate_single_sample / _generate_bandpass
 * (data_generation/synthetic_generator.py:520-656, 657-673).  Distribution parity only (the
 * reference's host MT19937 stream cannot be reproduced on a device): every pixel owns a
 * Philox4x32-10 counter (key = seed; counter = pixel, first_baseline + b, stream), so the cube is a
 * pure function of (seed, baseline index, shape) however the baselines are sharded over GPUs.
 * The RFI events of a baseline (drawn on the host as the reference's _add_* do, :675-815) arrive in
 * separable form -- all of them are rows x times rectangles or sweeps. */
typedef struct rfi_synth {
    int64_t channels, times;
    int32_t n_pol;           /* pol 0 full RFI, pol 1 pol_corr x RFI, pols >= 2 noise only (:616-636) */
    int32_t enable_bandpass; /* polynomial roll-off over 10 % of the channels at each edge (:657-673) */
    int32_t bandpass_order;
    int32_t n_bands;         /* narrow bands with a time profile per baseline (<= 64) */
    int32_t n_sweeps;        /* frequency sweeps per baseline */
    float noise_level;       /* clean amplitude ~ N(noise_level, 0.1 noise_level) (:549) */
    float pol_corr;
    uint64_t seed;
} rfi_synth_t;

/*   row_amp   device float32 [n_baselines][C]          summed amplitude of the full-row events
 *   col_amp   device float32 [n_baselines][T]          summed amplitude of the full-column events
 *   band_rows device int32   [n_baselines][n_bands][2] row range [r0, r1) of band k
 *   band_amp  device float32 [n_baselines][n_bands][T] amplitude of band k at time t (0 = off)
 *   sweep     device float32 [n_baselines][n_sweeps][6] start row, end row, width, order (1|2), amplitude, 0
 *   cube      device complex64 [n_baselines][n_pol][C][T], written
 *   mask      device uint8     same shape, written (exact RFI locations, pols 0 and 1) */
int rfi_synth_waterfalls(const rfi_synth_t* sp, int64_t n_baselines, int64_t first_baseline,
                         const float* row_amp, const float* col_amp, const int32_t* band_rows,
                         const float* band_amp, const float* sweep, void* cube, uint8_t* mask,
                         void* stream);

/* Raw complex patches -- replaces the tiling / blank removal / shuffle gathers of
 * GPUPreprocessor.create_raw_patches (preprocessor.py:846-940, :942-972): non-overlapping P x P
 * tiles of every waterfall (patchify with step P: remainders dropped, no padding), or the whole
 * waterfall when it is no larger than the patch (:885-890); masks are the caller's flags or
 * |z| > 0 (:881-883).  dtype RFI_C64 / RFI_C128 only (:832-836).
 *   rfi_raw_num_tiles    tiles before blank removal (-1: bad arguments) and the tile shape
 *   rfi_raw_tile_counts  counts device int32[tiles]: flagged samples per tile (mask.any(), :906)
 *   rfi_raw_gather       dest_slot device int64[tiles] (-1 = dropped) -> patches device
 *                        (N, rows, cols) in the input dtype, masks device uint8 (N, rows, cols) */
int64_t rfi_raw_num_tiles(int dtype, int64_t n_waterfalls, int64_t channels, int64_t times, int32_t patch,
                          int32_t* tile_rows, int32_t* tile_cols);
int rfi_raw_tile_counts(const void* data, int dtype, const uint8_t* flags, int64_t n_waterfalls,
                        int64_t channels, int64_t times, int32_t patch, int32_t* counts, void* stream);
int rfi_raw_gather(const void* data, int dtype, const uint8_t* flags, int64_t n_waterfalls,
                   int64_t channels, int64_t times, int32_t patch, const int64_t* dest_slot,
                   void* patches, uint8_t* masks, void* stream);

/* Rotated, zero-padded copies of the waterfalls -- view `rotation` of _apply_rotations
 * (preprocessor.py:413-446: 0 = X, 1 = X[::-1, :], 2 = X.T, 3 = X.T[::-1, :]) followed by the bottom /
 * right zero pad of _create_patches (:527-550).  A geometry whose dims are not multiples of P then runs
 * through the on-chip kernels as single-view plans over these copies.
 *   in   device (n_waterfalls, channels, times), elements of elem_bytes (1, 4, 8 or 16)
 *   out  device (n_waterfalls, out_rows, out_cols) >= the rotated view, written */
int rfi_rotate_pad(const void* in, void* out, int elem_bytes, int64_t n_waterfalls, int64_t channels,
                   int64_t times, int64_t out_rows, int64_t out_cols, int rotation, void* stream);

/* Self test (used by tests/): counts the inputs t in [1, 2] (all 2^23 + 1 float32 values) for
 * which the range-restricted square root of the magnitude kernel differs from sqrt.rn.f32.
 *   mismatches_dev  device uint64, ACCUMULATED into (caller zeroes); must end up 0 */
/* Preprocessor.patches -- the side-effect attribute of preprocessor.py:194, 272-311, 345-359: the PROCESSED
 * patches (normalised / stretched / inf-filled samples in the data's precision on the real branch, the raw
 * complex samples on the complex branch) in the dataset's final order.  The hot path never materialises
 * them; this rebuilds them on demand from the cube and the statistics rfi_tile_stats left.
 *   stats  device, as written by rfi_tile_stats / rfi_fused_patches for this plan
 *   order  device int64[n_out], canonical patch index of every output patch (rfi_plan_slots' `order`)
 *   out    device, n_out x P x P samples of T (real branch) or of the input's complex type
 * Only for plans on RFI_PATH_FAST / RFI_PATH_BIG (dims multiples of the patch size). */
int rfi_processed_patches(const rfi_plan_t* plan, const void* data, const rfi_tile_stat_t* stats,
                          const int64_t* order, int64_t n_out, void* out, void* stream);

/* Ingestion of the reference loaders' complex128 / float64 cubes (io/ms_loader.py:202-238,
 * data_generation/synthetic_generator.py:648) as complex64 / float32: every component rounded once to
 * nearest-even (= ndarray.astype(np.complex64)).  Opt-in (Preprocessor(..., compute_dtype="float32")): the
 * reference itself computes such input in float64 up to preprocessor.py:376.
 *   in   device, n samples of RFI_F64 / RFI_C128 (16-byte aligned);  out  device, n samples of RFI_F32 / RFI_C64 */
int rfi_downcast(const void* in, void* out, int dtype_in, int64_t n, void* stream);

int rfi_selftest_sqrt_unit(unsigned long long* mismatches_dev, void* stream);
int rfi_selftest_cabs_fast(unsigned long long* mismatches_dev, void* stream);

const char* rfi_last_error_string(void);
int rfi_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RFI_B200_H */
