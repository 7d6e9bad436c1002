/*
 * mdb200.h -- C ABI of the B200-native deconvolution hot path (libmdb200.so).
 *
 * This is the drop-in boundary: exactly the entry points a `metabodecon-sys` FFI crate would
 * bind so that metabodecon's Rust API (and through it the PyO3 bindings) routes the
 * deconvolution path to the GPU.  Every declaration cites the reference interface it replaces;
 * citations are relative to the reference repo root (metabodecon/src/... unless noted).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.
 *  - inputs are borrowed for the duration of the call and never retained.
 *  - every function returns an mdb_status; MDB_OK == 0.
 *  - thread-safe and re-entrant: a Deconvoluter is `Send + Sync + Clone` in the reference
 *    (deconvolution/deconvoluter.rs:913-917) and so is an mdb_deconvoluter for const calls.
 *  - device selection follows the calling thread's current CUDA device (cudaSetDevice /
 *    torch.cuda.set_device); one process per GPU is the intended multi-GPU model.
 *  - there is NO CPU fallback: without a usable CUDA device every compute call returns
 *    MDB_ERR_CUDA and mdb_last_error_message() says why.
 */
#ifndef MDB200_H
#define MDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB200_ABI_VERSION 2

/* ---------------------------------------------------------------------------------------------
 * Status codes.  1..7 mirror deconvolution::error::Kind (deconvolution/error.rs:39-95),
 * 10..14 mirror the spectrum::error::Kind variants that Spectrum::new can raise
 * (spectrum/spectrum.rs:179-200, 756-905).
 * ------------------------------------------------------------------------------------------- */
typedef enum mdb_status {
    MDB_OK = 0,
    MDB_ERR_NO_PEAKS_DETECTED = 1,          /* Kind::NoPeaksDetected        detector.rs:107-109 */
    MDB_ERR_EMPTY_SIGNAL_REGION = 2,        /* Kind::EmptySignalRegion      noise_score_filter.rs:105-107,121-123 */
    MDB_ERR_EMPTY_SIGNAL_FREE_REGION = 3,   /* Kind::EmptySignalFreeRegion  noise_score_filter.rs:102-104 */
    MDB_ERR_INVALID_SMOOTHING_SETTINGS = 4, /* smoothing/smoother.rs:84-100 */
    MDB_ERR_INVALID_SELECTION_SETTINGS = 5, /* peak_selection/selector.rs:82-98 */
    MDB_ERR_INVALID_FITTING_SETTINGS = 6,   /* fitting/fitter.rs:76-90 */
    MDB_ERR_INVALID_IGNORE_REGION = 7,      /* deconvoluter.rs:439-444 */
    MDB_ERR_EMPTY_DATA = 10,                /* spectrum Kind::EmptyData */
    MDB_ERR_DATA_LENGTH_MISMATCH = 11,      /* spectrum Kind::DataLengthMismatch */
    MDB_ERR_NON_UNIFORM_SPACING = 12,       /* spectrum Kind::NonUniformSpacing */
    MDB_ERR_INVALID_INTENSITIES = 13,       /* spectrum Kind::InvalidIntensities */
    MDB_ERR_INVALID_SIGNAL_BOUNDARIES = 14, /* spectrum Kind::InvalidSignalBoundaries */
    MDB_ERR_REFERENCE_PANIC = 100,          /* input on which the reference panics (slice out of range / usize underflow) */
    MDB_ERR_CUDA = 200,                     /* driver/runtime failure or no device; see mdb_last_error_message */
    MDB_ERR_INVALID_ARGUMENT = 201,         /* NULL pointer, unsupported size (n < 5 or n >= 2^31) */
    MDB_ERR_UNSUPPORTED = 202               /* valid in the reference but outside this build's limits */
} mdb_status;

/* Lorentzian {sfhw, hw2, maxp}: field order and meaning of deconvolution/lorentzian.rs:138-145. */
typedef struct mdb_lorentzian {
    double sfhw; /* scale factor * half width */
    double hw2;  /* half width squared */
    double maxp; /* maximum position (ppm) */
} mdb_lorentzian;

/* SmoothingSettings (smoothing/smoother.rs:27-56), SelectionSettings (peak_selection/
 * selector.rs:32-57), ScoringMethod (peak_selection/scorer.rs:21-29), FittingSettings
 * (fitting/fitter.rs:34-57) as plain tagged PODs. */
enum { MDB_SMOOTHING_IDENTITY = 0, MDB_SMOOTHING_MOVING_AVERAGE = 1 };
enum { MDB_SELECTION_DETECTOR_ONLY = 0, MDB_SELECTION_NOISE_SCORE_FILTER = 1 };
enum { MDB_SCORING_MINIMUM_SUM = 0 };
enum { MDB_FITTING_ANALYTICAL = 0 };

typedef struct mdb_smoothing_settings {
    int32_t kind;
    uint64_t iterations;  /* MovingAverage only; default 3 */
    uint64_t window_size; /* MovingAverage only; default 3 */
} mdb_smoothing_settings;

typedef struct mdb_selection_settings {
    int32_t kind;
    int32_t scoring_method; /* NoiseScoreFilter only */
    double threshold;       /* NoiseScoreFilter only; default 5.0 */
} mdb_selection_settings;

typedef struct mdb_fitting_settings {
    int32_t kind;
    uint64_t iterations; /* default 10 */
} mdb_fitting_settings;

/* One spectrum as the path consumes it: the four accessors of spectrum/spectrum.rs the hot path
 * reads (chemical_shifts, intensities, signal_boundaries -- already ordered to match the axis'
 * monotonicity as Spectrum::new stores them, spectrum.rs:854-863 -- and the length). */
typedef struct mdb_spectrum_view {
    const double *chemical_shifts;
    const double *intensities;
    size_t len;
    double signal_boundaries[2];
} mdb_spectrum_view;

/* Where the arrays of an mdb_spectrum_view / superposition call live. */
enum { MDB_MEM_HOST = 0, MDB_MEM_DEVICE = 1 };

/* ---------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------- */
uint32_t mdb_abi_version(void);
/* Thread-local, NUL-terminated description of the last non-OK status on this thread. */
const char *mdb_last_error_message(void);
/* Number of CUDA devices visible; 0 when there is none (compute calls then fail loudly). */
int mdb_device_count(void);
/* GPUs used by ONE mdb_deconvolute_spectra call on host-memory batches: 1 (default) = the calling
 * thread's current device; 0 = every visible device; n = CUDA devices 0..n-1.  With more than one,
 * the batch is cut into contiguous shards, one host thread and one pipeline per GPU, no exchange
 * between them (the in-process form of deconvoluter.rs:699-710's rayon-over-spectra); host-memory
 * mdb_superposition_vec calls cut their grid the same way (lorentzian.rs:656-663).  The
 * environment variable MDB_DEVICES ("all" or a number) overrides it. */
mdb_status mdb_set_device_count(int n);
/* Arithmetic of the two superposition kernels that do NOT feed back into the fit: the full-grid
 * superposition behind Deconvolution::mse (deconvoluter.rs:540-543, 828-862) and
 * Lorentzian::superposition_vec / par_superposition_vec (lorentzian.rs:631-663).
 *   MDB_SUPERPOSITION_EXACT  every operator of the reference replayed with one IEEE rounding each:
 *                            values and MSE carry the reference's bit patterns;
 *   MDB_SUPERPOSITION_FAST   6 instead of 12 FP64 instructions per evaluation (fused denominator,
 *                            reciprocal to 2^-53 + 2^-60, fused accumulate): each term within about
 *                            2 ulp, same summation order over the Lorentzians; the squared residuals
 *                            of the MSE are summed by a fixed tree instead of one left fold.
 *                            Superposition values agree with the exact mode to about 1e-15
 *                            relative, the MSE to about 1e-13 (contract: 1e-9).
 * Peak sets and Lorentzian parameters are bit-identical in both modes (the refinement always uses
 * exact arithmetic), and so is the choice made by mdb_deconvoluter_optimize_settings, which
 * compares MSEs and therefore always computes them exactly.
 *
 * The mode is PER DECONVOLUTER (mdb_deconvoluter_set_superposition_mode pins it) and an explicit
 * argument of mdb_superposition_vec_mode: two deconvoluters of one process may differ and run
 * concurrently.  mdb_set_superposition_mode only sets the process DEFAULT -- what a deconvoluter that
 * was never pinned uses (read once at the start of each call) and what the argument-less
 * mdb_superposition_vec uses.  Its initial value is MDB_SUPERPOSITION_FAST unless the environment
 * variable MDB_SUPERPOSITION says "exact" or "fast"; any other value makes the calls that need the
 * default fail with MDB_ERR_INVALID_ARGUMENT (never a silent fallback).  mdb_superposition_mode
 * returns the default (negative when the environment variable is malformed). */
enum { MDB_SUPERPOSITION_EXACT = 0, MDB_SUPERPOSITION_FAST = 1 };
mdb_status mdb_set_superposition_mode(int mode);
int mdb_superposition_mode(void);
/* FP64 instruction rate of the current device, measured: independent DFMA / DADD streams on every
 * SM for a few milliseconds, CUDA events on the launching stream.  Instructions (not flops) per
 * second; the denominator bench.py's fp64_pipe_util is quoted against. */
mdb_status mdb_measure_fp64_rate(double *dfma_per_s, double *dadd_per_s);
/* Page-locked host memory for callers that want full-rate H2D (optional helper). */
mdb_status mdb_host_alloc(void **ptr, size_t bytes);
mdb_status mdb_host_free(void *ptr);
/* Release cached device workspaces of the current device. */
mdb_status mdb_release_workspaces(void);
/* Counters: kernels launched by this library in this process since the last reset. */
uint64_t mdb_kernel_launch_count(void);
/* Host->device and device->host bytes copied by this library since the last reset (e2e accounting). */
void mdb_transfer_bytes(uint64_t *h2d, uint64_t *d2h);
/* Resets the launch counter and the transfer byte counters. */
void mdb_reset_kernel_launch_count(void);

/* Optional per-kernel timing for roofline reports: CUDA events on the launching stream around
 * every kernel, accumulated per kernel family.  work = algorithmic units of that family:
 * bytes for SMOOTH / DETECT / MSE_REDUCE, Lorentzian evaluations for FIT_ITER /
 * MSE_SUPERPOSITION / SUPERPOSITION_VEC, items otherwise. */
enum {
    MDB_KERNEL_SMOOTH = 0,
    MDB_KERNEL_DETECT = 1,
    MDB_KERNEL_SELECT = 2,
    MDB_KERNEL_FIT_INIT = 3,
    MDB_KERNEL_FIT_ITER = 4,
    MDB_KERNEL_RETAIN = 5,
    MDB_KERNEL_MSE_SUPERPOSITION = 6,
    MDB_KERNEL_MSE_REDUCE = 7,
    MDB_KERNEL_SUPERPOSITION_VEC = 8,
    MDB_KERNEL_SMALL_FUSED = 9,
    MDB_KERNEL_COUNT = 10
};
void mdb_profile_enable(int on);
void mdb_profile_reset(void);
mdb_status mdb_profile_read(int kernel, double *ms, uint64_t *launches, double *work);

/* ---------------------------------------------------------------------------------------------
 * Spectrum validation: Spectrum::new (spectrum/spectrum.rs:179-200).  On success writes the
 * boundaries re-ordered to the axis monotonicity into ordered_boundaries[2].
 * ------------------------------------------------------------------------------------------- */
mdb_status mdb_spectrum_validate(const double *chemical_shifts, size_t n_shifts,
                                 const double *intensities, size_t n_intensities,
                                 const double signal_boundaries[2], double ordered_boundaries[2]);

/* ---------------------------------------------------------------------------------------------
 * Deconvoluter (deconvolution/deconvoluter.rs:117-127): settings + ignore regions.
 * ------------------------------------------------------------------------------------------- */
typedef struct mdb_deconvoluter mdb_deconvoluter;

/* Deconvoluter::default()  deconvoluter.rs:129-138 */
mdb_status mdb_deconvoluter_default(mdb_deconvoluter **out);
/* Deconvoluter::new(smoothing, selection, fitting)  deconvoluter.rs:172-207 */
mdb_status mdb_deconvoluter_new(const mdb_smoothing_settings *smoothing,
                                const mdb_selection_settings *selection,
                                const mdb_fitting_settings *fitting, mdb_deconvoluter **out);
/* Clone  (derive(Clone), deconvoluter.rs:117) */
mdb_status mdb_deconvoluter_clone(const mdb_deconvoluter *src, mdb_deconvoluter **out);
void mdb_deconvoluter_free(mdb_deconvoluter *d);

/* getters  deconvoluter.rs:229-298 */
mdb_status mdb_deconvoluter_smoothing_settings(const mdb_deconvoluter *d, mdb_smoothing_settings *out);
mdb_status mdb_deconvoluter_selection_settings(const mdb_deconvoluter *d, mdb_selection_settings *out);
mdb_status mdb_deconvoluter_fitting_settings(const mdb_deconvoluter *d, mdb_fitting_settings *out);
/* Number of ignore regions, or -1 for Option::None.  Copies up to cap (lo, hi) pairs. */
int64_t mdb_deconvoluter_ignore_regions(const mdb_deconvoluter *d, double *pairs, size_t cap);

/* setters  deconvoluter.rs:323-401; validation errors leave the deconvoluter unchanged */
mdb_status mdb_deconvoluter_set_smoothing_settings(mdb_deconvoluter *d, const mdb_smoothing_settings *s);
mdb_status mdb_deconvoluter_set_selection_settings(mdb_deconvoluter *d, const mdb_selection_settings *s);
mdb_status mdb_deconvoluter_set_fitting_settings(mdb_deconvoluter *d, const mdb_fitting_settings *s);
/* add_ignore_region / clear_ignore_regions  deconvoluter.rs:438-492 (sort + merge included) */
mdb_status mdb_deconvoluter_add_ignore_region(mdb_deconvoluter *d, double lo, double hi);
void mdb_deconvoluter_clear_ignore_regions(mdb_deconvoluter *d);
/* Pins the arithmetic of the MSE superposition of THIS deconvoluter (MDB_SUPERPOSITION_*); clones
 * inherit it.  The getter returns the pinned mode, or the process default for a deconvoluter that was
 * never pinned.  No counterpart in the reference (which has one arithmetic); see the enum above. */
mdb_status mdb_deconvoluter_set_superposition_mode(mdb_deconvoluter *d, int mode);
int mdb_deconvoluter_superposition_mode(const mdb_deconvoluter *d);
/* Arithmetic of the refinement passes (fitter_analytical.rs:39-66).  MDB_FIT_EXACT (default, the
 * product): the reference's operators, one IEEE rounding each -- Lorentzian parameters carry the
 * reference's bit patterns.  The other two are OPT-IN experiments whose measured deviation is
 * recorded in DESIGN.md section 2:
 *   MDB_FIT_CORRECTED  the division keeps its Markstein correction but drops the second Newton step
 *                      of the reciprocal (10 instead of 12 FP64 instructions per evaluation); the
 *                      quotient is the correctly rounded one except when a/b lies within ~2^-103
 *                      of a rounding boundary (about one quotient in 2^49);
 *   MDB_FIT_ULP        the 6-instruction few-ulp evaluation of MDB_SUPERPOSITION_FAST inside the
 *                      refinement: parameters drift by far more than 1e-9 relative on real spectra
 *                      (the analytic solve amplifies ulp noise).
 * The environment variable MDB_FIT_ARITHMETIC ("exact" | "corrected" | "ulp") sets the initial value
 * of new deconvoluters. */
enum { MDB_FIT_EXACT = 0, MDB_FIT_CORRECTED = 1, MDB_FIT_ULP = 2 };
mdb_status mdb_deconvoluter_set_fit_arithmetic(mdb_deconvoluter *d, int kind);
int mdb_deconvoluter_fit_arithmetic(const mdb_deconvoluter *d);

/* ---------------------------------------------------------------------------------------------
 * Deconvolution results (deconvolution/deconvolution.rs:45-56).  A batch result owns S
 * deconvolutions; element accessors return borrowed pointers valid until mdb_batch_free.
 * ------------------------------------------------------------------------------------------- */
typedef struct mdb_batch mdb_batch;

size_t mdb_batch_len(const mdb_batch *b);
/* Per-spectrum status (MDB_OK or the Kind the reference would return for that spectrum). */
mdb_status mdb_batch_status(const mdb_batch *b, size_t i);
size_t mdb_batch_n_lorentzians(const mdb_batch *b, size_t i);
const mdb_lorentzian *mdb_batch_lorentzians(const mdb_batch *b, size_t i);
double mdb_batch_mse(const mdb_batch *b, size_t i);
/* Diagnostics the reference does not surface but parity tests need (bit-exact peak sets). */
size_t mdb_batch_n_peaks(const mdb_batch *b, size_t i);
/* Selected peaks as (left, center, right) int32 triples, ascending by centre. */
const int32_t *mdb_batch_peaks(const mdb_batch *b, size_t i);
void mdb_batch_free(mdb_batch *b);
/* Bulk accessors for bindings with per-call overhead: totals over the batch, then one export of
 * every per-spectrum field (any output pointer may be NULL).  status/n_lorentzians/n_peaks/mse hold
 * mdb_batch_len entries; lorentzians / peaks are the per-spectrum arrays concatenated in order. */
void mdb_batch_totals(const mdb_batch *b, size_t *n_lorentzians, size_t *n_peaks);
mdb_status mdb_batch_export(const mdb_batch *b, int32_t *status, uint64_t *n_lorentzians, uint64_t *n_peaks,
                            double *mse, mdb_lorentzian *lorentzians, int32_t *peaks);

/*
 * Deconvoluter::deconvolute_spectra / par_deconvolute_spectra  (deconvoluter.rs:651-661,
 * 699-710) and, with n_spectra == 1, deconvolute_spectrum / par_deconvolute_spectrum
 * (:530-552, :590-613).  The serial and rayon variants of the reference give identical results
 * (same per-point summation order), so one GPU entry point serves all four.
 *
 * Return value: MDB_OK, or the status of the FIRST failing spectrum in index order -- the
 * `collect::<Result<Vec<_>>>` semantics of deconvoluter.rs:655-658.  *out is still produced when
 * a per-spectrum Kind error occurs so callers can inspect every spectrum's status; it is NULL
 * only for MDB_ERR_CUDA / MDB_ERR_INVALID_ARGUMENT.
 *
 * memory: MDB_MEM_HOST (pageable or pinned host arrays; the call performs H2D, compute, D2H) or
 * MDB_MEM_DEVICE (arrays already resident in this device's HBM; no input copies).
 * Spectra sharing the same chemical_shifts pointer share one device copy of the axis.
 */
mdb_status mdb_deconvolute_spectra(const mdb_deconvoluter *d, const mdb_spectrum_view *spectra,
                                   size_t n_spectra, int memory, mdb_batch **out);

/*
 * Deconvoluter::optimize_settings  (deconvoluter.rs:761-825): deconvolutes `reference` under 27
 * smoothing x 10 selection x 3 fitting settings (810 combinations, same grids and iteration order
 * as the reference), stores the combination with the lowest MSE in the deconvoluter (first minimum
 * wins) and writes that MSE to *mse.  Any failing combination aborts with its status and leaves
 * the deconvoluter unchanged.  Ignore regions of the deconvoluter apply.
 */
mdb_status mdb_deconvoluter_optimize_settings(mdb_deconvoluter *d, const mdb_spectrum_view *reference,
                                              int memory, double *mse);

/*
 * Lorentzian::superposition_vec / par_superposition_vec  (lorentzian.rs:631-635, 656-663):
 * out[i] = sum_j sfhw_j / (hw2_j + (x_i - maxp_j)^2), j ascending, one rounding per operation.
 * memory applies to x, lorentzians and out alike.
 */
mdb_status mdb_superposition_vec(const double *x, size_t n, const mdb_lorentzian *lorentzians,
                                 size_t n_lorentzians, double *out, int memory);
/* The same with the arithmetic stated by the caller (MDB_SUPERPOSITION_*) instead of the process
 * default.  Host-memory calls honour mdb_set_device_count: the grid is cut into contiguous slices,
 * one per GPU, the parameter table is replicated, and every slice streams through its GPU in
 * chunks (H2D, kernel and D2H of neighbouring chunks overlap) -- the form of par_superposition_vec's
 * rayon-over-points (lorentzian.rs:656-663).  Slices are bit-identical to a one-shot evaluation. */
mdb_status mdb_superposition_vec_mode(const double *x, size_t n, const mdb_lorentzian *lorentzians,
                                      size_t n_lorentzians, double *out, int memory, int mode);

/* ---------------------------------------------------------------------------------------------
 * Stage entry points.  The reference keeps these behind crate-private traits (Smoother,
 * Selector, Fitter); they are exported so the parity suite can compare every intermediate with
 * the oracle bit for bit.  Host memory only.
 * ------------------------------------------------------------------------------------------- */
/* MovingAverage::smooth_values  smoothing/moving_average.rs:53-83 (out may alias values). */
mdb_status mdb_stage_smooth(const double *values, size_t n, uint64_t iterations,
                            uint64_t window_size, double *out);
/* K1 on `count` equally long spectra already in DEVICE memory (rows of `stride` doubles), one
 * launch; *ms (optional) = kernel time from CUDA events.  Measurement aid: shows the launch size at
 * which the exact-recurrence smoothing turns from latency bound to bandwidth bound. */
mdb_status mdb_stage_smooth_batch(const double *values_dev, size_t n, size_t count, size_t stride,
                                  uint64_t iterations, uint64_t window_size, double *out_dev, double *ms);
/* second_derivative + Detector::detect_peaks  peak_selection/common.rs:5-10, detector.rs:99-113,
 * plus ScorerMinimumSum::score_peak (scorer.rs:65-74) for every detected triplet.
 * peaks: 3*cap int32 (left, center, right); scores: cap doubles; *n_found may exceed cap. */
mdb_status mdb_stage_detect(const double *smoothed, size_t n, int32_t *peaks, double *scores,
                            size_t cap, size_t *n_found);
/* Selector::select_peaks for the deconvoluter's selection settings
 * (noise_score_filter.rs:32-54, detector_only.rs:16-39).  ignore_idx: n_ignore (start,end) index
 * pairs, has_ignore = Option::is_some.  peaks: 3*cap int32.  mean_sd (optional): 2 doubles. */
mdb_status mdb_stage_select(const mdb_deconvoluter *d, const double *smoothed, size_t n,
                            size_t sb0, size_t sb1, int has_ignore, const size_t *ignore_idx,
                            size_t n_ignore, int32_t *peaks, size_t cap, size_t *n_selected,
                            double *mean_sd);
/* FitterAnalytical::fit_lorentzian  fitting/fitter_analytical.rs:19-72.  peaks: 3*n_peaks int32.
 * out: n_peaks lorentzians (first *n_retained valid).  trace (optional): (iterations+1)*n_peaks
 * lorentzians -- the initial solve and the state after each refinement pass, before retain. */
mdb_status mdb_stage_fit(const double *x, const double *y, size_t n, const int32_t *peaks,
                         size_t n_peaks, uint64_t iterations, mdb_lorentzian *out,
                         size_t *n_retained, mdb_lorentzian *trace);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H */
