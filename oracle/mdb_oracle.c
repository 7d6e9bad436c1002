/*
 * mdb_oracle.c -- CPU restatement of metabodecon's deconvolution hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (metabodecon_rust_b200/csrc, libmdb200.so) never links, imports or calls anything here.
 *
 * Parity status: the reference (Rust) cannot be compiled in this image (no cargo/rustc),
 * so this file is a scalar-f64 restatement of the reference's arithmetic, in the
 * reference's operation order, built with -ffp-contract=off (rustc never contracts to FMA).
 * It is PINNED against every known-answer unit test the reference holds for this path
 * (tests/test_oracle_kats.py lists them with file:line), against the survey's independent
 * NumPy checkpoints on blood_01, and against a second restatement written separately in NumPy
 * (oracle/numpy_restatement.py; bit-for-bit agreement on sim_01, blood_01 and synthetic spectra).  END-TO-END PARITY IS UNPINNED: the
 * reference commits no golden deconvolution output (metabodecon/tests/deconvoluter.rs
 * only writes JSON to a temp dir), so there is nothing end-to-end to pin against.
 *
 * All file:line citations are relative to /root/reference/metabodecon/src/.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_NO_PEAKS_DETECTED 1        /* deconvolution/error.rs Kind::NoPeaksDetected */
#define ORC_EMPTY_SIGNAL_REGION 2      /* Kind::EmptySignalRegion */
#define ORC_EMPTY_SIGNAL_FREE_REGION 3 /* Kind::EmptySignalFreeRegion */
#define ORC_PANIC 100                  /* the reference would panic (index out of range / underflow) */

#define ORC_SMOOTH_IDENTITY 0
#define ORC_SMOOTH_MOVING_AVERAGE 1
#define ORC_SELECT_DETECTOR_ONLY 0
#define ORC_SELECT_NOISE_SCORE_FILTER 1

/* lib.rs:277  CHECK_PRECISION = 1.0e+3 * f64::EPSILON */
static const double ORC_EPSILON = 2.220446049250313e-16;
#define ORC_CHECK_PRECISION (1.0e+3 * ORC_EPSILON)

/* Rust `f as usize`: saturating, NaN -> 0. */
static size_t f64_as_usize(double f)
{
    if (!(f == f)) return 0;
    if (f <= 0.0) return 0;
    if (f >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)f;
}

/* ------------------------------------------------------------------------------------------
 * Smoothing: smoothing/moving_average.rs:53-83, FIFO semantics smoothing/circular_buffer.rs:34-59.
 * One running sum per pass; add happens before subtract; `div` is a rounded reciprocal that is
 * multiplied, and it persists between steps.
 * ---------------------------------------------------------------------------------------- */
void orc_smooth_values(double *values, size_t n, size_t iterations, size_t window_size)
{
    size_t right = window_size / 2; /* moving_average.rs:114 */
    double *fifo = (double *)malloc(window_size * sizeof(double));
    for (size_t it = 0; it < iterations; ++it) {
        size_t head = 0, len = 0;
        double div = 1.0;
        double sum = 0.0;
        for (size_t k = 0; k < right && k < n; ++k) { /* :58-61 */
            fifo[(head + len) % window_size] = values[k];
            ++len;
            sum += values[k];
        }
        for (size_t i = 0; i + right < n; ++i) { /* :62-70 */
            double incoming = values[i + right];
            sum += incoming;
            if (len == window_size) { /* circular_buffer.rs:35-40: pop front, then push back */
                double popped = fifo[head];
                head = (head + 1) % window_size;
                fifo[(head + len - 1) % window_size] = incoming;
                sum -= popped;
            } else {
                fifo[(head + len) % window_size] = incoming;
                ++len;
                div = 1.0 / (double)len;
            }
            values[i] = sum * div;
        }
        for (size_t i = (n >= right ? n - right : 0); i < n; ++i) { /* :71-79 */
            if (len > 0) {
                double popped = fifo[head];
                head = (head + 1) % window_size;
                --len;
                sum -= popped;
                div = 1.0 / (double)len;
                values[i] = sum * div;
            }
        }
    }
    free(fifo);
}

/* peak_selection/common.rs:5-10   d2[j] = (y[j] - 2*y[j+1]) + y[j+2] */
void orc_second_derivative(const double *y, size_t n, double *d2)
{
    for (size_t j = 0; j + 2 < n; ++j) d2[j] = y[j] - 2.0 * y[j + 1] + y[j + 2];
}

/* peak_selection/detector.rs:150-154 (find_right_border) on the slice t = d2[c-1..] */
static size_t find_right_border(const double *t, size_t len)
{
    for (size_t p = 0; p + 2 < len; ++p) {
        const double *w = t + p;
        if (w[1] > w[0] && (w[1] >= w[2] || (w[1] < 0. && w[2] >= 0.))) return p + 1;
    }
    return len;
}

/* peak_selection/detector.rs:158-164 (find_left_border) on the slice u = d2[..c], windows reversed */
static size_t find_left_border(const double *u, size_t len)
{
    if (len >= 3) {
        size_t p = 0;
        for (size_t q = len - 3;; --q, ++p) {
            const double *w = u + q;
            if (w[1] > w[2] && (w[1] >= w[0] || (w[1] < 0. && w[0] >= 0.))) return p + 1;
            if (q == 0) break;
        }
    }
    return len;
}

/* Exported forms of the detector's private helpers so that the reference's own unit tests
 * (detector.rs:170-228) can be replayed against this file verbatim. */
size_t orc_find_right_border(const double *t, size_t len) { return find_right_border(t, len); }
size_t orc_find_left_border(const double *u, size_t len) { return find_left_border(u, len); }

/* detector.rs:120-127 */
size_t orc_find_peak_centers(const double *d2, size_t m, size_t *centers, size_t cap)
{
    size_t count = 0;
    for (size_t i = 0; i + 2 < m; ++i) {
        const double *w = d2 + i;
        if (w[1] < 0. && w[1] < w[0] && w[1] < w[2]) {
            if (count < cap) centers[count] = i + 2;
            ++count;
        }
    }
    return count;
}

/* detector.rs:133-146: (left, right) pairs for the given centres, sentinels included */
void orc_find_peak_borders(const double *d2, size_t m, const size_t *centers, size_t nc, size_t *borders)
{
    for (size_t k = 0; k < nc; ++k) {
        size_t c = centers[k];
        borders[2 * k] = c - find_left_border(d2, c);
        borders[2 * k + 1] = c + find_right_border(d2 + (c - 1), m - (c - 1));
    }
}

/* peak_selection/common.rs:26-40, exported for common.rs:61-68 */
static void peak_region_boundaries(const size_t *center, size_t np, size_t sb0, size_t sb1,
                                   size_t *out_left, size_t *out_right);
void orc_peak_region_boundaries(const size_t *center, size_t np, size_t sb0, size_t sb1, size_t *lr)
{
    peak_region_boundaries(center, np, sb0, sb1, &lr[0], &lr[1]);
}

/* peak_selection/detector.rs:99-127  -> triplets ascending by centre.  Returns count; if it
 * exceeds cap only the first cap are stored.  m = len(d2). */
size_t orc_detect_peaks(const double *d2, size_t m, size_t *left, size_t *center, size_t *right,
                        size_t cap)
{
    size_t count = 0;
    for (size_t i = 0; i + 2 < m; ++i) {
        const double *w = d2 + i;
        if (w[1] < 0. && w[1] < w[0] && w[1] < w[2]) { /* :124 */
            size_t c = i + 2;
            size_t l = c - find_left_border(d2, c);                   /* :138 */
            size_t r = c + find_right_border(d2 + (c - 1), m - (c - 1)); /* :139 */
            if (l != 0 && r != m + 1) {                               /* :105 */
                if (count < cap) {
                    left[count] = l;
                    center[count] = c;
                    right[count] = r;
                }
                ++count;
            }
        }
    }
    return count;
}

/* peak_selection/scorer.rs:65-74 on a = |d2|; ascending ordered sums starting from zero. */
double orc_score_peak(const double *a, size_t left, size_t center, size_t right)
{
    double ls = 0.0, rs = 0.0;
    for (size_t j = left - 1; j < center; ++j) ls += a[j];
    for (size_t j = center - 1; j < right; ++j) rs += a[j];
    return fmin(ls, rs); /* f64::min ignores NaN like fmin */
}

/* peak_selection/common.rs:26-40 */
static void peak_region_boundaries(const size_t *center, size_t np, size_t sb0, size_t sb1,
                                   size_t *out_left, size_t *out_right)
{
    size_t l = 0;
    for (size_t i = 0; i < np; ++i)
        if (center[i] > sb0) { l = i; break; }
    size_t r = np - 1; /* wraps when np == 0; the reference would then panic on slicing */
    for (size_t i = l; i < np; ++i)
        if (center[i] > sb1) { r = i; break; }
    *out_left = l;
    *out_right = r;
}

/* peak_selection/noise_score_filter.rs:129-138 */
void orc_mean_sd_scores(const double *scores, size_t n, double *mean, double *sd)
{
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += scores[i];
    double mu = s / (double)n;
    double v = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double d = scores[i] - mu;
        v += d * d;
    }
    v = v / (double)n;
    *mean = mu;
    *sd = sqrt(v);
}

static int peak_ignored(size_t l, size_t r, const size_t *ig, size_t n_ig)
{
    /* noise_score_filter.rs:41-48 / detector_only.rs:31-38 */
    for (size_t k = 0; k < n_ig; ++k) {
        size_t s = ig[2 * k], e = ig[2 * k + 1];
        if ((l >= s && l < e) || (r >= s && r < e)) return 1;
    }
    return 0;
}

/*
 * Selector::select_peaks for both selectors.
 *   smoothed[n]      intensities after smoothing
 *   sb0, sb1         Spectrum::signal_boundaries_indices (spectrum/spectrum.rs:741-746)
 *   ig[2*n_ig]       ignore regions as index pairs, has_ig = Option::is_some
 * Outputs (all caller-allocated with capacity cap >= n/2):
 *   sel_{left,center,right}, *n_sel
 *   optional diagnostics: n_detected, n_after_ignore, region split, mean, sd, scores of the
 *   detected-after-ignore peaks (cand_score) together with the candidates themselves.
 */
typedef struct {
    size_t n_detected;
    size_t n_after_ignore;
    size_t region_left, region_right;
    size_t n_sfr;
    double mean, sd;
} orc_select_info;

int orc_select_peaks(const double *smoothed, size_t n, int selector, double threshold, size_t sb0,
                     size_t sb1, int has_ig, const size_t *ig, size_t n_ig, size_t cap,
                     size_t *sel_left, size_t *sel_center, size_t *sel_right, size_t *n_sel,
                     size_t *cand_left, size_t *cand_center, size_t *cand_right,
                     double *cand_score, orc_select_info *info)
{
    if (n < 3) return ORC_PANIC;
    size_t m = n - 2;
    double *d2 = (double *)malloc(m * sizeof(double));
    size_t *pl = (size_t *)malloc(3 * cap * sizeof(size_t));
    size_t *pc = pl + cap, *pr = pc + cap;
    int status = ORC_OK;
    *n_sel = 0;
    orc_second_derivative(smoothed, n, d2);
    size_t np = orc_detect_peaks(d2, m, pl, pc, pr, cap);
    if (info) memset(info, 0, sizeof(*info));
    if (info) info->n_detected = np;
    if (np > cap) { status = ORC_PANIC; goto done; }
    if (np == 0) { status = ORC_NO_PEAKS_DETECTED; goto done; } /* detector.rs:107-109 */

    if (selector == ORC_SELECT_DETECTOR_ONLY) {
        /* detector_only.rs:16-39 */
        size_t k = 0;
        for (size_t i = 0; i < np; ++i) {
            if (!(pl[i] >= sb0 && pr[i] <= sb1)) continue;
            if (has_ig && peak_ignored(pl[i], pr[i], ig, n_ig)) continue;
            sel_left[k] = pl[i]; sel_center[k] = pc[i]; sel_right[k] = pr[i];
            ++k;
        }
        *n_sel = k;
        goto done;
    }

    /* noise_score_filter.rs:32-54 */
    if (has_ig) {
        size_t k = 0;
        for (size_t i = 0; i < np; ++i) {
            if (peak_ignored(pl[i], pr[i], ig, n_ig)) continue;
            pl[k] = pl[i]; pc[k] = pc[i]; pr[k] = pr[i];
            ++k;
        }
        np = k;
    }
    if (info) info->n_after_ignore = np;
    for (size_t j = 0; j < m; ++j) d2[j] = fabs(d2[j]);
    if (np == 0) { status = ORC_PANIC; goto done; } /* peaks.len() - 1 underflows, slicing panics */

    /* filter_peaks  noise_score_filter.rs:91-126 */
    size_t bl, br;
    peak_region_boundaries(pc, np, sb0, sb1, &bl, &br);
    if (info) { info->region_left = bl; info->region_right = br; }
    if (bl == 0 && br >= np) { status = ORC_EMPTY_SIGNAL_FREE_REGION; goto done; } /* :102-104 */
    if (br < bl) { status = ORC_PANIC; goto done; }                                /* cannot happen */
    if (bl == br) { status = ORC_EMPTY_SIGNAL_REGION; goto done; }                 /* :105-107 */
    {
        size_t n_sfr = bl + (np - br);
        double *sfr = (double *)malloc(n_sfr * sizeof(double));
        size_t k = 0;
        for (size_t i = 0; i < bl; ++i) sfr[k++] = orc_score_peak(d2, pl[i], pc[i], pr[i]);
        for (size_t i = br; i < np; ++i) sfr[k++] = orc_score_peak(d2, pl[i], pc[i], pr[i]);
        double mean, sd;
        orc_mean_sd_scores(sfr, n_sfr, &mean, &sd);
        free(sfr);
        if (info) { info->n_sfr = n_sfr; info->mean = mean; info->sd = sd; }
        if (cand_score) {
            for (size_t i = 0; i < np; ++i) {
                cand_left[i] = pl[i]; cand_center[i] = pc[i]; cand_right[i] = pr[i];
                cand_score[i] = orc_score_peak(d2, pl[i], pc[i], pr[i]);
            }
        }
        k = 0;
        for (size_t i = bl; i < br; ++i) {
            double sc = orc_score_peak(d2, pl[i], pc[i], pr[i]);
            if (sc >= mean + threshold * sd) { /* :118 (>=, no FMA) */
                sel_left[k] = pl[i]; sel_center[k] = pc[i]; sel_right[k] = pr[i];
                ++k;
            }
        }
        *n_sel = k;
        if (k == 0) status = ORC_EMPTY_SIGNAL_REGION; /* :121-123 */
    }
done:
    free(pl);
    free(d2);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Lorentzian: deconvolution/lorentzian.rs:138-145 {sfhw, hw2, maxp};
 * evaluate :546-548, superposition :606-611 (ordered sum from zero), superposition_vec :631-635.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double sfhw, hw2, maxp; } orc_lorentzian;

static inline double lor_eval(const orc_lorentzian *l, double x)
{
    double d = x - l->maxp;
    return l->sfhw / (l->hw2 + d * d);
}

double orc_superposition(double x, const orc_lorentzian *l, size_t p)
{
    double s = 0.0;
    for (size_t j = 0; j < p; ++j) s += lor_eval(l + j, x);
    return s;
}

void orc_superposition_vec(const double *x, size_t n, const orc_lorentzian *l, size_t p, double *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = orc_superposition(x[i], l, p);
}

/* par_superposition_vec (lorentzian.rs:656-663): rayon over x; per-point order unchanged. */
void orc_par_superposition_vec(const double *x, size_t n, const orc_lorentzian *l, size_t p,
                               double *out)
{
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) out[i] = orc_superposition(x[i], l, p);
}

/* ------------------------------------------------------------------------------------------
 * Analytical fitter: fitting/fitter_analytical.rs:19-72 (+ par variant :77-130),
 * fitting/peak_stencil.rs:27-36,113-131, fitting/reduced_spectrum.rs:16-42.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double x1, x2, x3, y1, y2, y3; } orc_stencil;

void orc_mirror_shoulder(orc_stencil *s) /* peak_stencil.rs:113-131 */
{
    int increasing = s->y1 <= s->y2 && s->y2 <= s->y3;
    int decreasing = s->y1 >= s->y2 && s->y2 >= s->y3;
    if (increasing) {
        s->y3 = s->y1;
        s->x3 = 2.0 * s->x2 - s->x1;
    } else if (decreasing) {
        s->y1 = s->y3;
        s->x1 = 2.0 * s->x2 - s->x3;
    }
}

static double maximum_position(const orc_stencil *p) /* fitter_analytical.rs:147-155 */
{
    double numerator = p->x1 * p->x1 * p->y1 * (p->y2 - p->y3)
                     + p->x2 * p->x2 * p->y2 * (p->y3 - p->y1)
                     + p->x3 * p->x3 * p->y3 * (p->y1 - p->y2);
    double divisor = 2.0 * (p->x1 - p->x2) * p->y1 * p->y2
                   + 2.0 * (p->x2 - p->x3) * p->y2 * p->y3
                   + 2.0 * (p->x3 - p->x1) * p->y3 * p->y1;
    return numerator / divisor;
}

static double half_width2(const orc_stencil *p, double maxp) /* fitter_analytical.rs:159-165 */
{
    double d1 = (p->x1 - maxp) * (p->x1 - maxp);
    double d2 = (p->x2 - maxp) * (p->x2 - maxp);
    double d3 = (p->x3 - maxp) * (p->x3 - maxp);
    double left = (p->y1 * d1 - p->y2 * d2) / (p->y2 - p->y1);
    double right = (p->y2 * d2 - p->y3 * d3) / (p->y3 - p->y2);
    return fmax((left + right) / 2.0, ORC_EPSILON); /* f64::max: NaN -> EPSILON, as fmax */
}

static double scale_factor_half_width(const orc_stencil *p, double maxp, double hw2)
{
    return p->y2 * (hw2 + (p->x2 - maxp) * (p->x2 - maxp)); /* fitter_analytical.rs:170-172 */
}

void orc_solve_stencil(const orc_stencil *s, orc_lorentzian *out)
{
    double maxp = maximum_position(s);
    double hw2 = half_width2(s, maxp);
    double sfhw = scale_factor_half_width(s, maxp, hw2);
    out->sfhw = sfhw;
    out->hw2 = hw2;
    out->maxp = maxp;
}

/*
 * fit_lorentzian.  x, y are the ORIGINAL (unsmoothed) spectrum arrays.  `lor` has capacity np.
 * Returns the number retained (sfhw > CHECK_PRECISION && hw2 > CHECK_PRECISION).
 * If trace != NULL it receives (iterations+1)*np parameter triples: the initial solve followed
 * by the state after every refinement pass, before the final retain.  `parallel` selects the
 * OpenMP-over-points form (par_fit_lorentzian); results are identical by construction.
 */
size_t orc_fit_lorentzian(const double *x, const double *y, const size_t *left,
                          const size_t *center, const size_t *right, size_t np, size_t iterations,
                          orc_lorentzian *lor, orc_lorentzian *trace, int parallel)
{
    if (np == 0) return 0;
    double *rx = (double *)malloc(6 * np * sizeof(double));
    double *ry = rx + 3 * np;
    orc_stencil *st = (orc_stencil *)malloc(np * sizeof(orc_stencil));
    double *sup = (double *)malloc(3 * np * sizeof(double));
    for (size_t k = 0; k < np; ++k) { /* reduced_spectrum.rs:16-42, peak_stencil.rs:27-36 */
        rx[3 * k] = x[left[k]]; rx[3 * k + 1] = x[center[k]]; rx[3 * k + 2] = x[right[k]];
        ry[3 * k] = y[left[k]]; ry[3 * k + 1] = y[center[k]]; ry[3 * k + 2] = y[right[k]];
        st[k].x1 = rx[3 * k]; st[k].x2 = rx[3 * k + 1]; st[k].x3 = rx[3 * k + 2];
        st[k].y1 = ry[3 * k]; st[k].y2 = ry[3 * k + 1]; st[k].y3 = ry[3 * k + 2];
        orc_mirror_shoulder(&st[k]);
    }
    for (size_t k = 0; k < np; ++k) orc_solve_stencil(&st[k], &lor[k]);
    if (trace) memcpy(trace, lor, np * sizeof(orc_lorentzian));
    for (size_t it = 0; it < iterations; ++it) { /* :39-66 */
        if (parallel) orc_par_superposition_vec(rx, 3 * np, lor, np, sup);
        else orc_superposition_vec(rx, 3 * np, lor, np, sup);
        for (size_t k = 0; k < np; ++k) {
            double r0 = ry[3 * k] / sup[3 * k];
            double r1 = ry[3 * k + 1] / sup[3 * k + 1];
            double r2 = ry[3 * k + 2] / sup[3 * k + 2];
            st[k].y1 = st[k].y1 * r0;
            st[k].y2 = st[k].y2 * r1;
            st[k].y3 = st[k].y3 * r2;
            orc_mirror_shoulder(&st[k]);
        }
        for (size_t k = 0; k < np; ++k) orc_solve_stencil(&st[k], &lor[k]);
        if (trace) memcpy(trace + (it + 1) * np, lor, np * sizeof(orc_lorentzian));
    }
    size_t kept = 0; /* :67-69 */
    for (size_t k = 0; k < np; ++k)
        if (lor[k].sfhw > ORC_CHECK_PRECISION && lor[k].hw2 > ORC_CHECK_PRECISION) lor[kept++] = lor[k];
    free(sup);
    free(st);
    free(rx);
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * Index helpers and MSE: spectrum/spectrum.rs:633-635,741-746; deconvoluter.rs:828-904.
 * ---------------------------------------------------------------------------------------- */
void orc_signal_boundaries_indices(const double *x, double sb0, double sb1, size_t *i0, size_t *i1)
{
    double step = x[1] - x[0];
    *i0 = f64_as_usize(floor((sb0 - x[0]) / step));
    *i1 = f64_as_usize(ceil((sb1 - x[0]) / step));
}

/* deconvoluter.rs:865-904.  regions[2*n_regions] in ppm as stored by add_ignore_region (each pair
 * (min,max), list sorted by start, merged).  Output pairs in the same list order. */
size_t orc_ignore_region_indices(const double *x, double sb0, double sb1, const double *regions,
                                 size_t n_regions, size_t *out)
{
    double step = x[1] - x[0];
    double first = x[0];
    double lower_boundary = fmin(sb0, sb1), upper_boundary = fmax(sb0, sb1);
    size_t b0, b1;
    orc_signal_boundaries_indices(x, sb0, sb1, &b0, &b1);
    size_t lower = b0 < b1 ? b0 : b1, upper = b0 < b1 ? b1 : b0;
    size_t k = 0;
    for (size_t r = 0; r < n_regions; ++r) {
        double start = regions[2 * r], end = regions[2 * r + 1];
        if ((start < lower_boundary && end < lower_boundary)
            || (start > upper_boundary && end > upper_boundary))
            continue;
        size_t fi = f64_as_usize(floor((start - first) / step));
        size_t si = f64_as_usize(ceil((end - first) / step));
        if (fi < lower) fi = lower;
        if (si > upper) si = upper;
        size_t lo = fi < si ? fi : si, hi = fi < si ? si : fi;
        if (lo < hi - 1) { /* usize arithmetic: hi == 0 wraps in release builds */
            out[2 * k] = lo;
            out[2 * k + 1] = hi;
            ++k;
        }
    }
    return k;
}

/* deconvoluter.rs:828-862.  Returns ORC_PANIC if a range is reversed or out of bounds. */
int orc_compute_mse(const double *sup, const double *y, size_t n, size_t sb0, size_t sb1,
                    int has_ig, const size_t *ig, size_t n_ig, double *mse)
{
    size_t n_pts = 2 + (has_ig ? 2 * n_ig : 0);
    size_t *pts = (size_t *)malloc(n_pts * sizeof(size_t));
    size_t k = 0;
    pts[k++] = sb0;
    if (has_ig)
        for (size_t r = 0; r < n_ig; ++r) { pts[k++] = ig[2 * r]; pts[k++] = ig[2 * r + 1]; }
    pts[k++] = sb1;
    double residuals = 0.0;
    size_t length = 0;
    int status = ORC_OK;
    for (size_t r = 0; r + 1 < n_pts; r += 2) {
        size_t s = pts[r], e = pts[r + 1];
        if (s > e || e > n) { status = ORC_PANIC; break; }
        double part = 0.0;
        for (size_t i = s; i < e; ++i) {
            double d = sup[i] - y[i];
            part += d * d;
        }
        residuals += part;
        length += e - s; /* usize; cannot wrap once s <= e */
    }
    free(pts);
    *mse = residuals / (double)length;
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Deconvoluter::deconvolute_spectrum / par_deconvolute_spectrum  deconvoluter.rs:530-552,590-613
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int smoothing_kind;       /* ORC_SMOOTH_* */
    size_t smoothing_iterations;
    size_t smoothing_window;
    int selection_kind;       /* ORC_SELECT_* */
    double threshold;
    size_t fitting_iterations;
    int has_ignore_regions;   /* Option::is_some */
    size_t n_ignore_regions;
    const double *ignore_regions; /* 2*n, ppm, merged + sorted */
} orc_settings;

typedef struct {
    int status;
    size_t n_selected;
    size_t n_lorentzians;
    double mse;
    orc_select_info info;
} orc_result;

/* lor_out capacity n/2; sel_* (optional) capacity n/2; smoothed_out (optional) n. */
int orc_deconvolute_spectrum(const orc_settings *st, const double *x, const double *y, size_t n,
                             double sb0, double sb1, int parallel, orc_lorentzian *lor_out,
                             size_t *sel_left, size_t *sel_center, size_t *sel_right,
                             double *smoothed_out, orc_result *res)
{
    memset(res, 0, sizeof(*res));
    size_t cap = n / 2 + 1;
    double *sm = (double *)malloc(n * sizeof(double));
    memcpy(sm, y, n * sizeof(double));
    if (st->smoothing_kind == ORC_SMOOTH_MOVING_AVERAGE)
        orc_smooth_values(sm, n, st->smoothing_iterations, st->smoothing_window);
    if (smoothed_out) memcpy(smoothed_out, sm, n * sizeof(double));

    size_t i0, i1;
    orc_signal_boundaries_indices(x, sb0, sb1, &i0, &i1);
    size_t *ig = NULL, n_ig = 0;
    if (st->has_ignore_regions) {
        ig = (size_t *)malloc((2 * st->n_ignore_regions + 2) * sizeof(size_t));
        n_ig = orc_ignore_region_indices(x, sb0, sb1, st->ignore_regions, st->n_ignore_regions, ig);
    }
    size_t *own = NULL;
    if (!sel_left) {
        own = (size_t *)malloc(3 * cap * sizeof(size_t));
        sel_left = own; sel_center = own + cap; sel_right = own + 2 * cap;
    }
    size_t n_sel = 0;
    int status = orc_select_peaks(sm, n, st->selection_kind, st->threshold, i0, i1,
                                  st->has_ignore_regions, ig, n_ig, cap, sel_left, sel_center,
                                  sel_right, &n_sel, NULL, NULL, NULL, NULL, &res->info);
    res->status = status;
    res->n_selected = n_sel;
    if (status == ORC_OK) {
        size_t kept = orc_fit_lorentzian(x, y, sel_left, sel_center, sel_right, n_sel,
                                         st->fitting_iterations, lor_out, NULL, parallel);
        res->n_lorentzians = kept;
        double *sup = (double *)malloc(n * sizeof(double));
        if (parallel) orc_par_superposition_vec(x, n, lor_out, kept, sup);
        else orc_superposition_vec(x, n, lor_out, kept, sup);
        int ms = orc_compute_mse(sup, y, n, i0, i1, st->has_ignore_regions, ig, n_ig, &res->mse);
        if (ms != ORC_OK) res->status = ms;
        free(sup);
    }
    free(own);
    free(ig);
    free(sm);
    return res->status;
}

/*
 * par_deconvolute_spectra (deconvoluter.rs:699-710): rayon over spectra, nested rayon inside.
 * Restated with OpenMP over spectra (the inner loops run serially inside each worker; per-point
 * summation order is unchanged, so results are identical).  All spectra share n here (bench use).
 * lor_out: [S][cap] with cap = n/2+1; results[S].  Returns the first non-OK status in index order
 * (collect::<Result<Vec<_>>> semantics), ORC_OK otherwise.
 */
int orc_par_deconvolute_spectra(const orc_settings *st, size_t n_spectra, const double *const *xs,
                                const double *const *ys, size_t n, const double *sb,
                                orc_lorentzian *lor_out, orc_result *results)
{
    size_t cap = n / 2 + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (ptrdiff_t s = 0; s < (ptrdiff_t)n_spectra; ++s)
        orc_deconvolute_spectrum(st, xs[s], ys[s], n, sb[2 * s], sb[2 * s + 1], 0,
                                 lor_out + (size_t)s * cap, NULL, NULL, NULL, NULL, &results[s]);
    for (size_t s = 0; s < n_spectra; ++s)
        if (results[s].status != ORC_OK) return results[s].status;
    return ORC_OK;
}

/*
 * Deconvoluter::optimize_settings  deconvoluter.rs:761-825.  27 smoothing settings (iterations
 * 2..=10 outer, window 3,5,7 inner) x 10 thresholds (5.0 + c*(8.0-5.0)/9.0) x 3 fit iteration
 * counts (5, 10, 15); rayon over the smoothing settings -> OpenMP here.  The minimum is the FIRST
 * smallest MSE in iteration order (Iterator::min_by).  best[4] = {iterations, window, threshold,
 * fit iterations}.  Returns the first non-OK status in iteration order, ORC_PANIC for a NaN MSE.
 */
int orc_optimize_settings(const orc_settings *base, const double *x, const double *y, size_t n, double sb0,
                          double sb1, double *best, double *best_mse, double *all_mse)
{
    enum { NS = 27, NT = 10, NF = 3 };
    int status[NS * NT * NF];
    double mse[NS * NT * NF];
    size_t cap = n / 2 + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < NS; ++s) {
        orc_lorentzian *lor = (orc_lorentzian *)malloc(cap * sizeof(orc_lorentzian));
        for (int c = 0; c < NT; ++c)
            for (int f = 0; f < NF; ++f) {
                orc_settings st = *base;
                st.smoothing_kind = ORC_SMOOTH_MOVING_AVERAGE;
                st.smoothing_iterations = (size_t)(2 + s / 3);
                st.smoothing_window = (size_t)(3 + 2 * (s % 3));
                st.selection_kind = ORC_SELECT_NOISE_SCORE_FILTER;
                st.threshold = 5.0 + (double)c * (8.0 - 5.0) / 9.0;
                st.fitting_iterations = (size_t)(5 + 5 * f);
                orc_result res;
                orc_deconvolute_spectrum(&st, x, y, n, sb0, sb1, 0, lor, NULL, NULL, NULL, NULL, &res);
                status[(s * NT + c) * NF + f] = res.status;
                mse[(s * NT + c) * NF + f] = res.mse;
            }
        free(lor);
    }
    int best_i = 0;
    for (int i = 0; i < NS * NT * NF; ++i) {
        if (all_mse) all_mse[i] = mse[i];
        if (status[i] != ORC_OK) return status[i];
        if (!(mse[i] == mse[i])) return ORC_PANIC;
        if (mse[i] < mse[best_i]) best_i = i;
    }
    int s = best_i / (NT * NF), c = (best_i / NF) % NT, f = best_i % NF;
    best[0] = (double)(2 + s / 3);
    best[1] = (double)(3 + 2 * (s % 3));
    best[2] = 5.0 + (double)c * (8.0 - 5.0) / 9.0;
    best[3] = (double)(5 + 5 * f);
    *best_mse = mse[best_i];
    return ORC_OK;
}

/* bench.py's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: let the caller
 * ask for all host cores explicitly. */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
