// api.cu -- host side of libmdb200.so: the C ABI of include/mdb200.h, the Deconvoluter settings
// logic (validation, ignore-region merge, ppm -> index helpers) and the chunked GPU pipeline.
//
// There is no CPU compute path in this file: every stage of the deconvolution runs in the
// kernels of kernels.cuh, and every entry point fails with MDB_ERR_CUDA when no device is usable.
//
// Reference citations are relative to /root/reference/metabodecon/src/.
#include "../../include/mdb200.h"
#include "kernels.cuh"
#include "smooth_lanes.cuh"
#include "small_fused.cuh"
#include "smooth_stream.cuh"
#include "smooth_split.cuh"
#include "fit_wide.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

using namespace mdb;

// ---------------------------------------------------------------------------------------------
// errors, counters
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};  // host<->device bytes moved by this library

static void count_transfer(size_t bytes, cudaMemcpyKind kind)
{
    if (kind == cudaMemcpyHostToDevice) g_h2d_bytes += bytes;
    else if (kind == cudaMemcpyDeviceToHost) g_d2h_bytes += bytes;
}
static cudaError_t counted_memcpy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
    count_transfer(bytes, kind);
    return cudaMemcpyAsync(dst, src, bytes, kind, stream);
}
static cudaError_t counted_memcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    count_transfer(bytes, kind);
    return cudaMemcpy(dst, src, bytes, kind);
}

static mdb_status fail(mdb_status st, const std::string &msg)
{
    g_last_error = msg;
    return st;
}

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            cudaGetLastError(); /* reported here: do not leave it for the next launch check */    \
            return fail(MDB_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(e__)  \
                                          + " (" __FILE__ ":" + std::to_string(__LINE__) + ")");  \
        }                                                                                         \
    } while (0)

#define LAUNCH_CHECK()                                                                            \
    do {                                                                                          \
        ++g_launches;                                                                             \
        CUDA_TRY(cudaGetLastError());                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Optional per-kernel timing (bench.py's roofline numbers): CUDA events recorded on the stream the
// kernel is launched on, resolved when the chunk is finished.  Off by default.
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_profile{0};
struct ProfAcc { double ms = 0.0; uint64_t launches = 0; double work = 0.0; };
static std::mutex g_prof_mutex;
static ProfAcc g_prof[MDB_KERNEL_COUNT];
struct ProfSpan { int kernel; cudaEvent_t e0, e1; double work; };

// Timeline mode (MDB_TIMELINE=<file>, tools/timeline.py): the spans are recorded even when per-kernel
// profiling is off, and every span is also written out as (kernel, start, end) on the GPU's clock.
static thread_local cudaEvent_t t_timeline_base = nullptr;  // non-null: timeline mode
struct TraceRow { int kernel; float t0, t1; };
static thread_local std::vector<TraceRow> t_trace;

static void prof_begin(std::vector<ProfSpan> *spans, int kernel, cudaStream_t stream)
{
    if (!spans || !(g_profile.load() || t_timeline_base)) return;
    ProfSpan sp{kernel, nullptr, nullptr, 0.0};
    if (cudaEventCreate(&sp.e0) != cudaSuccess || cudaEventCreate(&sp.e1) != cudaSuccess) return;
    cudaEventRecord(sp.e0, stream);
    spans->push_back(sp);
}
static void prof_end(std::vector<ProfSpan> *spans, cudaStream_t stream, double work)
{
    if (!spans || !(g_profile.load() || t_timeline_base) || spans->empty()) return;
    ProfSpan &sp = spans->back();
    sp.work = work;
    cudaEventRecord(sp.e1, stream);
}
// call after the stream has been synchronised
static void prof_resolve(std::vector<ProfSpan> *spans)
{
    if (!spans) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (ProfSpan &sp : *spans) {
        float ms = 0.f;
        if (t_timeline_base) {
            TraceRow row{sp.kernel, -1.f, -1.f};
            if (cudaEventElapsedTime(&row.t0, t_timeline_base, sp.e0) == cudaSuccess
                && cudaEventElapsedTime(&row.t1, t_timeline_base, sp.e1) == cudaSuccess)
                t_trace.push_back(row);
            else
                cudaGetLastError();
        }
        if (g_profile.load() && cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess && sp.kernel >= 0 && sp.kernel < MDB_KERNEL_COUNT) {
            g_prof[sp.kernel].ms += ms;
            g_prof[sp.kernel].launches += 1;
            g_prof[sp.kernel].work += sp.work;
        }
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    spans->clear();
}
extern "C" void mdb_profile_enable(int on) { g_profile.store(on ? 1 : 0); }
extern "C" void mdb_profile_reset(void)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto &a : g_prof) a = ProfAcc();
}
extern "C" mdb_status mdb_profile_read(int kernel, double *ms, uint64_t *launches, double *work)
{
    if (kernel < 0 || kernel >= MDB_KERNEL_COUNT) return fail(MDB_ERR_INVALID_ARGUMENT, "unknown kernel id");
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (ms) *ms = g_prof[kernel].ms;
    if (launches) *launches = g_prof[kernel].launches;
    if (work) *work = g_prof[kernel].work;
    return MDB_OK;
}

// CUDA multiplexes streams over CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8) and streams
// that share a queue serialise behind each other.  The variable is read when the process creates
// its CUDA context, so ask for the maximum as soon as the library is loaded -- unless the user chose
// a value.  (No effect when the context already exists; the pipeline then still works, with fewer
// of its streams truly independent.)
__attribute__((constructor)) static void mdb_library_loaded() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" uint32_t mdb_abi_version(void) { return MDB200_ABI_VERSION; }
extern "C" void mdb_transfer_bytes(uint64_t *h2d, uint64_t *d2h)
{
    if (h2d) *h2d = g_h2d_bytes.load();
    if (d2h) *d2h = g_d2h_bytes.load();
}
extern "C" const char *mdb_last_error_message(void) { return g_last_error.c_str(); }
extern "C" uint64_t mdb_kernel_launch_count(void) { return g_launches.load(); }
extern "C" void mdb_reset_kernel_launch_count(void)
{
    g_launches.store(0);
    g_h2d_bytes.store(0);
    g_d2h_bytes.store(0);
}

extern "C" int mdb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" mdb_status mdb_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_host_alloc: null out pointer");
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
    return MDB_OK;
}

extern "C" mdb_status mdb_host_free(void *ptr)
{
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// Rust numeric semantics used by the index helpers
// ---------------------------------------------------------------------------------------------
static const double F64_EPSILON = 2.220446049250313e-16;
static const double CHECK_PRECISION = 1.0e+3 * F64_EPSILON;  // lib.rs:277

static size_t f64_as_usize(double f)  // `as usize`: saturating, NaN -> 0
{
    if (!(f == f) || f <= 0.0) return 0;
    if (f >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)f;
}

static int clamp_idx(size_t v) { return v > (size_t)INT_MAX ? INT_MAX : (int)v; }

// Spectrum::signal_boundaries_indices  spectrum/spectrum.rs:741-746 (step: :633-635)
static void signal_boundaries_indices(double x0, double x1, double sb0, double sb1, size_t *i0, size_t *i1)
{
    const double step = x1 - x0;
    *i0 = f64_as_usize(std::floor((sb0 - x0) / step));
    *i1 = f64_as_usize(std::ceil((sb1 - x0) / step));
}

// Deconvoluter::ignore_region_indices  deconvoluter.rs:865-904
static std::vector<std::pair<size_t, size_t>>
ignore_region_indices(double x0, double x1, double sb0, double sb1,
                      const std::vector<std::pair<double, double>> &regions)
{
    std::vector<std::pair<size_t, size_t>> out;
    const double step = x1 - x0, first = x0;
    const double lower_boundary = std::fmin(sb0, sb1), upper_boundary = std::fmax(sb0, sb1);
    size_t b0, b1;
    signal_boundaries_indices(x0, x1, sb0, sb1, &b0, &b1);
    const size_t lower = std::min(b0, b1), upper = std::max(b0, b1);
    for (const auto &reg : regions) {
        const double start = reg.first, end = reg.second;
        if ((start < lower_boundary && end < lower_boundary) || (start > upper_boundary && end > upper_boundary))
            continue;
        const size_t fi = std::max(f64_as_usize(std::floor((start - first) / step)), lower);
        const size_t si = std::min(f64_as_usize(std::ceil((end - first) / step)), upper);
        const size_t lo = std::min(fi, si), hi = std::max(fi, si);
        if (lo < hi - 1) out.emplace_back(lo, hi);  // usize arithmetic: wraps for hi == 0 (release build)
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Settings validation + Deconvoluter object
// ---------------------------------------------------------------------------------------------
static mdb_status validate_smoothing(const mdb_smoothing_settings &s)  // smoother.rs:84-100
{
    if (s.kind == MDB_SMOOTHING_IDENTITY) return MDB_OK;
    if (s.kind != MDB_SMOOTHING_MOVING_AVERAGE)
        return fail(MDB_ERR_INVALID_SMOOTHING_SETTINGS, "unknown smoothing kind");
    if (s.iterations == 0 || s.window_size <= 1)
        return fail(MDB_ERR_INVALID_SMOOTHING_SETTINGS,
                    "invalid smoothing settings: iterations must be > 0 and window size > 1");
    return MDB_OK;
}

static mdb_status validate_selection(const mdb_selection_settings &s)  // selector.rs:82-98
{
    if (s.kind == MDB_SELECTION_DETECTOR_ONLY) return MDB_OK;
    if (s.kind != MDB_SELECTION_NOISE_SCORE_FILTER || s.scoring_method != MDB_SCORING_MINIMUM_SUM)
        return fail(MDB_ERR_INVALID_SELECTION_SETTINGS, "unknown selection kind or scoring method");
    if (s.threshold <= 0.0 || !std::isfinite(s.threshold))
        return fail(MDB_ERR_INVALID_SELECTION_SETTINGS,
                    "invalid selection settings: threshold must be finite and > 0");
    return MDB_OK;
}

static mdb_status validate_fitting(const mdb_fitting_settings &s)  // fitter.rs:76-90
{
    if (s.kind != MDB_FITTING_ANALYTICAL) return fail(MDB_ERR_INVALID_FITTING_SETTINGS, "unknown fitting kind");
    if (s.iterations == 0)
        return fail(MDB_ERR_INVALID_FITTING_SETTINGS, "invalid fitting settings: iterations must be > 0");
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// Arithmetic of K7 (the MSE superposition) and K8 (superposition_vec).  The mode is a property of
// each mdb_deconvoluter (and an explicit argument of mdb_superposition_vec_mode); the process-wide
// value below is only the DEFAULT a new deconvoluter starts with and what the argument-less
// mdb_superposition_vec uses.  MDB_SUPERPOSITION ("exact" | "fast") sets its initial value; any
// other value is an error reported by the first call that needs it -- never a silent fallback.
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_superposition_default{-1};

static int initial_superposition_mode()  // -2: MDB_SUPERPOSITION holds an unknown value
{
    const char *env = std::getenv("MDB_SUPERPOSITION");
    if (!env || !env[0]) return MDB_SUPERPOSITION_FAST;
    if (std::strcmp(env, "exact") == 0) return MDB_SUPERPOSITION_EXACT;
    if (std::strcmp(env, "fast") == 0) return MDB_SUPERPOSITION_FAST;
    return -2;
}

extern "C" int mdb_superposition_mode(void)
{
    int m = g_superposition_default.load();
    if (m == -1) {
        m = initial_superposition_mode();
        int expected = -1;
        if (!g_superposition_default.compare_exchange_strong(expected, m)) m = expected;
    }
    return m;  // -2 when the environment variable is malformed
}

static mdb_status default_superposition_mode(int *mode)
{
    const int m = mdb_superposition_mode();
    if (m < 0)
        return fail(MDB_ERR_INVALID_ARGUMENT, std::string("MDB_SUPERPOSITION must be \"exact\" or \"fast\", not \"")
                                                  + (std::getenv("MDB_SUPERPOSITION") ? std::getenv("MDB_SUPERPOSITION") : "") + "\"");
    *mode = m;
    return MDB_OK;
}

extern "C" mdb_status mdb_set_superposition_mode(int mode)
{
    if (mode != MDB_SUPERPOSITION_EXACT && mode != MDB_SUPERPOSITION_FAST)
        return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_set_superposition_mode: unknown mode");
    g_superposition_default.store(mode);
    return MDB_OK;
}

struct mdb_deconvoluter {
    mdb_smoothing_settings smoothing;
    mdb_selection_settings selection;
    mdb_fitting_settings fitting;
    bool has_ignore;
    std::vector<std::pair<double, double>> ignore;
    int sup_mode = -1;  // arithmetic of K7: -1 = the process default at call time, else pinned (mdb_deconvoluter_set_superposition_mode)
    int fit_arith = MDB_FIT_EXACT;          // arithmetic of K6; anything but EXACT is an opt-in experiment
};

extern "C" mdb_status mdb_deconvoluter_new(const mdb_smoothing_settings *sm, const mdb_selection_settings *se,
                                           const mdb_fitting_settings *fi, mdb_deconvoluter **out)
{
    if (!sm || !se || !fi || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvoluter_new: null argument");
    mdb_status st;
    if ((st = validate_smoothing(*sm)) != MDB_OK) return st;
    if ((st = validate_selection(*se)) != MDB_OK) return st;
    if ((st = validate_fitting(*fi)) != MDB_OK) return st;
    auto *d = new mdb_deconvoluter();
    if (const char *fa = std::getenv("MDB_FIT_ARITHMETIC")) {  // experiments only; the product default is EXACT
        if (std::strcmp(fa, "corrected") == 0) d->fit_arith = MDB_FIT_CORRECTED;
        else if (std::strcmp(fa, "ulp") == 0) d->fit_arith = MDB_FIT_ULP;
        else if (std::strcmp(fa, "exact") != 0) { delete d; return fail(MDB_ERR_INVALID_ARGUMENT, "MDB_FIT_ARITHMETIC must be exact, corrected or ulp"); }
    }
    d->smoothing = *sm;
    d->selection = *se;
    d->fitting = *fi;
    d->has_ignore = false;
    *out = d;
    return MDB_OK;
}

extern "C" mdb_status mdb_deconvoluter_default(mdb_deconvoluter **out)
{
    // SmoothingSettings::default (smoother.rs:58-65), SelectionSettings::default (selector.rs:59-66),
    // FittingSettings::default (fitter.rs:59-63)
    mdb_smoothing_settings sm = {MDB_SMOOTHING_MOVING_AVERAGE, 3, 3};
    mdb_selection_settings se = {MDB_SELECTION_NOISE_SCORE_FILTER, MDB_SCORING_MINIMUM_SUM, 5.0};
    mdb_fitting_settings fi = {MDB_FITTING_ANALYTICAL, 10};
    return mdb_deconvoluter_new(&sm, &se, &fi, out);
}

extern "C" mdb_status mdb_deconvoluter_clone(const mdb_deconvoluter *src, mdb_deconvoluter **out)
{
    if (!src || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvoluter_clone: null argument");
    *out = new mdb_deconvoluter(*src);
    return MDB_OK;
}

extern "C" void mdb_deconvoluter_free(mdb_deconvoluter *d) { delete d; }

extern "C" mdb_status mdb_deconvoluter_set_superposition_mode(mdb_deconvoluter *d, int mode)
{
    if (!d) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (mode != MDB_SUPERPOSITION_EXACT && mode != MDB_SUPERPOSITION_FAST)
        return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvoluter_set_superposition_mode: unknown mode");
    d->sup_mode = mode;
    return MDB_OK;
}
extern "C" int mdb_deconvoluter_superposition_mode(const mdb_deconvoluter *d)
{
    if (!d) return -1;
    return d->sup_mode >= 0 ? d->sup_mode : mdb_superposition_mode();
}

extern "C" mdb_status mdb_deconvoluter_set_fit_arithmetic(mdb_deconvoluter *d, int kind)
{
    if (!d) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (kind != MDB_FIT_EXACT && kind != MDB_FIT_CORRECTED && kind != MDB_FIT_ULP)
        return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvoluter_set_fit_arithmetic: unknown kind");
    d->fit_arith = kind;
    return MDB_OK;
}
extern "C" int mdb_deconvoluter_fit_arithmetic(const mdb_deconvoluter *d) { return d ? d->fit_arith : -1; }

extern "C" mdb_status mdb_deconvoluter_smoothing_settings(const mdb_deconvoluter *d, mdb_smoothing_settings *out)
{
    if (!d || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = d->smoothing;
    return MDB_OK;
}
extern "C" mdb_status mdb_deconvoluter_selection_settings(const mdb_deconvoluter *d, mdb_selection_settings *out)
{
    if (!d || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = d->selection;
    return MDB_OK;
}
extern "C" mdb_status mdb_deconvoluter_fitting_settings(const mdb_deconvoluter *d, mdb_fitting_settings *out)
{
    if (!d || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = d->fitting;
    return MDB_OK;
}
extern "C" int64_t mdb_deconvoluter_ignore_regions(const mdb_deconvoluter *d, double *pairs, size_t cap)
{
    if (!d || !d->has_ignore) return -1;
    for (size_t i = 0; i < d->ignore.size() && i < cap && pairs; ++i) {
        pairs[2 * i] = d->ignore[i].first;
        pairs[2 * i + 1] = d->ignore[i].second;
    }
    return (int64_t)d->ignore.size();
}

extern "C" mdb_status mdb_deconvoluter_set_smoothing_settings(mdb_deconvoluter *d, const mdb_smoothing_settings *s)
{
    if (!d || !s) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    mdb_status st = validate_smoothing(*s);
    if (st == MDB_OK) d->smoothing = *s;
    return st;
}
extern "C" mdb_status mdb_deconvoluter_set_selection_settings(mdb_deconvoluter *d, const mdb_selection_settings *s)
{
    if (!d || !s) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    mdb_status st = validate_selection(*s);
    if (st == MDB_OK) d->selection = *s;
    return st;
}
extern "C" mdb_status mdb_deconvoluter_set_fitting_settings(mdb_deconvoluter *d, const mdb_fitting_settings *s)
{
    if (!d || !s) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    mdb_status st = validate_fitting(*s);
    if (st == MDB_OK) d->fitting = *s;
    return st;
}

// add_ignore_region  deconvoluter.rs:438-472: order the pair, insert, sort by start, merge
// neighbours that overlap or touch within CHECK_PRECISION until none is left.
extern "C" mdb_status mdb_deconvoluter_add_ignore_region(mdb_deconvoluter *d, double a, double b)
{
    if (!d) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (!std::isfinite(a) || !std::isfinite(b) || std::fabs(a - b) < CHECK_PRECISION)
        return fail(MDB_ERR_INVALID_IGNORE_REGION, "invalid ignore region: bounds must be finite and distinct");
    const std::pair<double, double> reg(std::fmin(a, b), std::fmax(a, b));
    if (!d->has_ignore) {
        d->has_ignore = true;
        d->ignore.assign(1, reg);
        return MDB_OK;
    }
    auto &v = d->ignore;
    v.push_back(reg);
    std::sort(v.begin(), v.end(), [](const std::pair<double, double> &p, const std::pair<double, double> &q) {
        return p.first < q.first;
    });
    for (;;) {
        size_t pos = v.size();
        for (size_t i = 0; i + 1 < v.size(); ++i) {
            if (v[i + 1].first < v[i].second || std::fabs(v[i].second - v[i + 1].first) < CHECK_PRECISION) {
                pos = i;
                break;
            }
        }
        if (pos == v.size()) break;
        const std::pair<double, double> combined(std::fmin(v[pos].first, v[pos + 1].first),
                                                 std::fmax(v[pos].second, v[pos + 1].second));
        v.erase(v.begin() + pos, v.begin() + pos + 2);
        v.insert(v.begin() + pos, combined);
    }
    return MDB_OK;
}

extern "C" void mdb_deconvoluter_clear_ignore_regions(mdb_deconvoluter *d)
{
    if (!d) return;
    d->has_ignore = false;
    d->ignore.clear();
}

// ---------------------------------------------------------------------------------------------
// Spectrum::new validation  spectrum/spectrum.rs:179-200, 756-905; meta/monotonicity.rs:29-38
// ---------------------------------------------------------------------------------------------
extern "C" mdb_status mdb_spectrum_validate(const double *x, size_t nx, const double *y, size_t ny,
                                            const double sb[2], double ordered[2])
{
    if (!sb || !ordered || (nx && !x) || (ny && !y)) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (nx == 0 || ny == 0) return fail(MDB_ERR_EMPTY_DATA, "empty data");
    if (nx != ny) return fail(MDB_ERR_DATA_LENGTH_MISMATCH, "chemical shifts and intensities differ in length");
    if (nx < 2) return fail(MDB_ERR_REFERENCE_PANIC, "a single-point spectrum makes the reference index out of range");
    const double step = x[1] - x[0];
    if (std::fabs(step) < CHECK_PRECISION) return fail(MDB_ERR_NON_UNIFORM_SPACING, "step size is zero");
    for (size_t i = 0; i + 1 < nx; ++i) {
        const double dd = x[i + 1] - x[i];
        if (std::fabs(dd - step) > CHECK_PRECISION || !std::isfinite(dd))
            return fail(MDB_ERR_NON_UNIFORM_SPACING, "chemical shifts are not uniformly spaced at index " + std::to_string(i));
    }
    for (size_t i = 0; i < ny; ++i)
        if (!std::isfinite(y[i])) return fail(MDB_ERR_INVALID_INTENSITIES, "non-finite intensity at index " + std::to_string(i));
    const bool increasing = x[0] < x[1];
    const double width = sb[0] - sb[1];
    if (std::fabs(width) < CHECK_PRECISION || !std::isfinite(width))
        return fail(MDB_ERR_INVALID_SIGNAL_BOUNDARIES, "signal boundaries are not distinct finite numbers");
    const double lo = std::fmin(sb[0], sb[1]), hi = std::fmax(sb[0], sb[1]);
    const double first = x[0], last = x[nx - 1];
    if (increasing) {
        ordered[0] = lo; ordered[1] = hi;
        if (ordered[0] < first || ordered[1] > last)
            return fail(MDB_ERR_INVALID_SIGNAL_BOUNDARIES, "signal boundaries outside the chemical shift range");
    } else {
        ordered[0] = hi; ordered[1] = lo;
        if (ordered[0] > first || ordered[1] < last)
            return fail(MDB_ERR_INVALID_SIGNAL_BOUNDARIES, "signal boundaries outside the chemical shift range");
    }
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// Device / pinned buffers and workspaces
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct Workspace {
    int device = -1;
    // Two streams per workspace.  `stream_a` (highest priority) carries stage A -- input copies, K1
    // smoothing, K2/K3 detection, K4 selection: short, latency- or bandwidth-bound kernels; `stream`
    // (lowest priority) carries stage B -- the FP64-bound fit and MSE kernels.  Grids of different
    // chunks are dispatched in priority order as SM slots free up, so the stage-A kernels of the next
    // chunks slip in beside the long FP64 grids of the current ones instead of queueing behind
    // them (their CTAs need few registers and little FP64 issue, so they overlap for free).
    cudaStream_t stream = nullptr, stream_a = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // stage A
    DevBuf x, y, ys, tmp, tile_cnt, pk, sc, sfr, sel, ig, desc, sel_out;
    PinBuf h_desc, h_ig, h_sel_out, h_stage;
    // stage B
    DevBuf fdesc, segs, fit_state, par_a, par_b, lor, n_kept, resid, mse, peaks_dense, fit_queue, blk_off;
    PinBuf h_fdesc, h_segs, h_lor, h_n_kept, h_mse, h_peaks, h_blk_off;
    // small-spectrum path (run_small)
    DevBuf small_in;
    PinBuf h_small_in, h_small_out;
    void release()
    {
        for (DevBuf *b : {&x, &y, &ys, &tmp, &tile_cnt, &pk, &sc, &sfr, &sel, &ig, &desc, &sel_out, &fdesc,
                          &segs, &fit_state, &par_a, &par_b, &lor, &n_kept, &resid, &mse, &peaks_dense, &fit_queue, &blk_off,
                          &small_in})
            b->release();
        for (PinBuf *b : {&h_desc, &h_ig, &h_sel_out, &h_stage, &h_fdesc, &h_segs, &h_lor, &h_n_kept, &h_mse, &h_peaks, &h_blk_off,
                          &h_small_in, &h_small_out})
            b->release();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (stream) cudaStreamDestroy(stream);
        if (stream_a) cudaStreamDestroy(stream_a);
        ev_a = ev_b = nullptr;
        stream = stream_a = nullptr;
    }
};

// Per-process pool of idle workspaces, keyed by device.  Calls check workspaces out and back in,
// so concurrent callers (the reference's Deconvoluter is Sync) never share one.
static std::mutex g_pool_mutex;
static std::multimap<int, Workspace *> g_pool;

static mdb_status acquire_workspace(Workspace **out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    {
        // most recently released first: it is the one whose buffers match the current chunk size
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto range = g_pool.equal_range(dev);
        if (range.first != range.second) {
            auto it = std::prev(range.second);
            *out = it->second;
            g_pool.erase(it);
            return MDB_OK;
        }
    }
    auto *ws = new Workspace();
    ws->device = dev;
    int prio_least = 0, prio_greatest = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (const char *env = std::getenv("MDB_STREAM_PRIORITIES")) if (env[0] == '0') prio_greatest = prio_least;  // measurement aid
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ws->stream, cudaStreamNonBlocking, prio_least);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ws->stream_a, cudaStreamNonBlocking, prio_greatest);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ws->ev_a, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ws->ev_b, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        ws->release();
        delete ws;
        return fail(MDB_ERR_CUDA, std::string("cannot create CUDA stream/events: ") + cudaGetErrorString(e));
    }
    *out = ws;
    return MDB_OK;
}

static void release_workspace(Workspace *ws)
{
    if (!ws) return;
    cudaStreamSynchronize(ws->stream_a);  // callers synchronise `stream`; nothing may be pending on either
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool.emplace(ws->device, ws);
}

static void release_pipe_stream_pool(int dev);  // defined with PipeStreams below

extern "C" mdb_status mdb_release_workspaces(void)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::vector<Workspace *> mine;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto range = g_pool.equal_range(dev);
        for (auto it = range.first; it != range.second; ++it) mine.push_back(it->second);
        g_pool.erase(range.first, range.second);
    }
    for (Workspace *ws : mine) {
        ws->release();
        delete ws;
    }
    release_pipe_stream_pool(dev);
    return MDB_OK;
}

static mdb_status require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MDB_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU fallback): ")
                                      + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// Per-spectrum host bookkeeping
// ---------------------------------------------------------------------------------------------
struct HostSpec {
    const double *x = nullptr, *y = nullptr;  // caller pointers (host or device)
    size_t n = 0;
    double sb[2] = {0, 0};
    double x0 = 0, x1 = 0;
    int sb_i0 = 0, sb_i1 = 0;
    std::vector<int> ig;                       // flattened (start, end) index pairs
    std::vector<std::pair<int, int>> ranges;   // MSE ranges (deconvoluter.rs:828-845)
    int pre_status = MDB_OK;                   // status known before any launch
    // per-spectrum overrides (optimize_settings runs one spectrum under many settings)
    const double *ys_dev = nullptr;            // already-smoothed intensities in device memory
    double threshold = 0.0;                    // selection threshold
    int fit_iters = 0;                         // refinement passes
};

struct SpecResult {
    int status = MDB_OK;
    std::vector<mdb_lorentzian> lor;
    std::vector<int32_t> peaks;  // triples
    double mse = 0.0;
    SelectOut info{};
};

struct mdb_batch {
    std::vector<SpecResult> r;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host part of a spectrum that does not depend on the GPU: index helpers, MSE ranges.
static void precompute_spec(HostSpec &h, const mdb_deconvoluter &dc)
{
    size_t i0, i1;
    signal_boundaries_indices(h.x0, h.x1, h.sb[0], h.sb[1], &i0, &i1);
    h.sb_i0 = clamp_idx(i0);
    h.sb_i1 = clamp_idx(i1);
    std::vector<std::pair<size_t, size_t>> ig;
    if (dc.has_ignore) ig = ignore_region_indices(h.x0, h.x1, h.sb[0], h.sb[1], dc.ignore);
    h.ig.clear();
    for (auto &p : ig) {
        h.ig.push_back(clamp_idx(p.first));
        h.ig.push_back(clamp_idx(p.second));
    }
    // compute_mse ranges: sb0, [ig start, ig end]*, sb1 paired up (deconvoluter.rs:829-845)
    std::vector<size_t> pts;
    pts.push_back(i0);
    if (dc.has_ignore)
        for (auto &p : ig) { pts.push_back(p.first); pts.push_back(p.second); }
    pts.push_back(i1);
    h.ranges.clear();
    for (size_t k = 0; k + 1 < pts.size(); k += 2) {
        const size_t s = pts[k], e = pts[k + 1];
        if (s > e || e > h.n) { h.pre_status = MDB_ERR_REFERENCE_PANIC; h.ranges.clear(); return; }  // slice panics
        h.ranges.emplace_back((int)s, (int)e);
    }
}

// ---------------------------------------------------------------------------------------------
// Pageable host rows (what a drop-in caller holds: Spectrum owns Arc<[f64]>, spectrum/spectrum.rs:
// 101-116; NumPy arrays from Python): cudaMemcpyAsync would bounce them through the driver's own
// staging buffer on the calling thread at ~10 GB/s and block it.  Instead a few host threads gather
// the rows of a chunk into the chunk workspace's page-locked staging area, one chunk AHEAD of the
// pipeline (the job is submitted when the chunk is created, one iteration before its stage A), in
// parts of ~32 MB so that the DMA of part p runs while part p+1 is still being gathered.
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("hidden"))) void mdb_host_row_copy(double *dst, const double *src, size_t n);  // hostcopy.cpp
struct HostSpec;
struct StageJob {
    std::vector<const double *> src;     // row pointers
    std::vector<size_t> len, off;        // elements per row, element offset in the staging area
    double *dst = nullptr;
    std::vector<size_t> part_end;        // row index (exclusive) closing each part
    std::vector<int> part_of_row;
    std::unique_ptr<std::atomic<int>[]> part_left;  // rows of the part not yet copied
    std::atomic<size_t> next{0};
    std::mutex m;
    std::condition_variable cv;
    void wait_part(size_t p)
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return part_left[p].load() == 0; });
    }
    bool all_done() const
    {
        for (size_t p = 0; p < part_end.size(); ++p)
            if (part_left[p].load() != 0) return false;
        return true;
    }
};

class Stager {
  public:
    explicit Stager(size_t n_threads)
    {
        for (size_t t = 0; t < n_threads; ++t) threads_.emplace_back([this] { worker(); });
    }
    ~Stager()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void submit(const std::shared_ptr<StageJob> &j)
    {
        { std::lock_guard<std::mutex> lk(m_); q_.push_back(j); }
        cv_.notify_all();
    }
  private:
    void worker()
    {
        for (;;) {
            std::shared_ptr<StageJob> j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;  // stop requested and nothing left
                j = q_.front();
            }
            const size_t rows = j->src.size();
            for (;;) {
                const size_t r = j->next.fetch_add(1);
                if (r >= rows) break;
                mdb_host_row_copy(j->dst + j->off[r], j->src[r], j->len[r]);  // non-temporal stores (hostcopy.cpp)
                if (j->part_left[j->part_of_row[r]].fetch_sub(1) == 1) {
                    std::lock_guard<std::mutex> lk(j->m);
                    j->cv.notify_all();
                }
            }
            std::lock_guard<std::mutex> lk(m_);
            if (!q_.empty() && q_.front() == j) q_.pop_front();  // every row has been claimed
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::shared_ptr<StageJob>> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// Number of pipelines of this process that run side by side (the in-process multi-GPU sharder sets it
// in its worker threads): the staging threads of one pipeline are sized so that all of them together
// stay within the host's cores.
static thread_local int t_pipeline_peers = 1;

// ---------------------------------------------------------------------------------------------
// The pipeline over one chunk of spectra
// ---------------------------------------------------------------------------------------------
struct Chunk {
    Workspace *ws = nullptr;
    size_t first = 0, count = 0;       // range in the batch
    std::vector<SpecDesc> desc;        // host mirror
    std::vector<FitDesc> fdesc;
    std::vector<Segment> segs;
    long long p_total = 0;             // selected peaks in the chunk
    long long res_total = 0;
    int max_tiles = 0, max_peaks = 0, max_seg_len = 0;
    bool stage_b_launched = false, finished = false;
    double est_evals = 0.0;            // Lorentzian evaluations of this chunk's fit + MSE kernels (from the counts)
    std::shared_ptr<StageJob> stage_job;  // pageable host rows being gathered into ws->h_stage (null: none)
    bool startup = false;                 // one of the first chunks of a call: its stage A runs on an otherwise idle GPU
    // Streams of this chunk (null: the workspace's own).  The pipeline hands out streams from a small
    // per-call set (PipeStreams) instead of one pair per workspace, see run_pipeline.
    cudaStream_t s_copy = nullptr, s_a = nullptr, s_b = nullptr, s_m = nullptr;
    cudaEvent_t ev_in = nullptr;   // inputs landed (recorded on s_copy, awaited by s_a)
    cudaEvent_t ev_fit = nullptr;  // refinement done (recorded on s_b, awaited by s_m)
    // host clock (ms since the pipeline started) at which stage A was queued, the counts had arrived,
    // stage B was queued and the results had arrived; written out by MDB_TIMELINE=<file> (tools/timeline.py)
    double t_a = 0.0, t_counts = 0.0, t_b = 0.0, t_done = 0.0;
    // the same on the GPU's clock (timeline mode only): inputs landed, smoothing done, stage A done, first
    // stage-B kernel started, stage B done -- CUDA events, ms after the pipeline's base event
    cudaEvent_t g_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<ProfSpan> spans;       // per-kernel timing, resolved in finish_chunk
};

// Device attributes used by the launch heuristics, cached PER DEVICE (the in-process sharder runs one
// host thread per GPU, and callers may sit on any device) behind a once-flag each.
struct DeviceInfo { std::once_flag once; int smem_optin = 0; int sms = 148; };
static DeviceInfo g_device_info[64];

static const DeviceInfo &device_info()
{
    int dev = 0;
    cudaGetDevice(&dev);
    DeviceInfo &di = g_device_info[(dev >= 0 && dev < 64) ? dev : 0];
    std::call_once(di.once, [&]() {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) di.smem_optin = v;
        v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) di.sms = v;
    });
    return di;
}
static int smem_optin_limit() { return device_info().smem_optin; }
// MDB_SUPERPOSITION_FAST: four Lorentzians per reciprocal (lorentz_quad_ulp) unless MDB_SUP_GROUP=1 asks for
// the one-at-a-time few-ulp form (measurement aid; read at every launch so that a sweep can flip it)
// MDB_STREAM_GENERIC=1 (tests): the latency forms of K1 (smooth_stream.cuh, small_fused.cuh) run their any-window
// interior loop even for the windows they have a specialised loop for
static bool stream_generic_env()
{
    const char *env = std::getenv("MDB_STREAM_GENERIC");
    return env && env[0] == '1';
}
// MDB_FIT_BLOCK=1 (experiment, tested): chunks of spectra with up to 1 024 peaks run every refinement pass in ONE
// launch, one CTA per spectrum (fit_block_kernel).  Bit-identical, but measured 5 % slower on config 3 (4 000
// spectra: 38.0 k against 40.0 k spectra/s, profiles/sweep_r2.txt): CTAs that hold their SM for a millisecond
// interleave worse with the MSE kernels of the other chunks than many short one-warp CTAs do.
static bool fit_block_enabled()
{
    const char *env = std::getenv("MDB_FIT_BLOCK");
    return env && env[0] == '1';
}
static bool sup_quad()
{
    const char *env = std::getenv("MDB_SUP_GROUP");
    return !(env && env[0] == '1');
}
constexpr int FIT_WARP_CTA_MAX = 1024;  // K6: spectra with at most this many selected peaks run one-warp CTAs
static int sm_count() { return device_info().sms; }

// K1 dispatch: the lane-per-pass TMA-staged kernel when the settings are covered (window 2..9,
// up to 32 iterations) and every input row is 16-byte aligned, else one generic pass per launch.
static mdb_status launch_smooth(cudaStream_t stream, const SpecDesc *d_desc, const std::vector<SpecDesc> &descs,
                                int iters, int window, std::vector<ProfSpan> *spans, size_t latency_form_up_to = 32)
{
    const size_t S = descs.size();
    double pts = 0.0;
    bool aligned = true;
    for (const SpecDesc &d : descs) {
        pts += (double)d.n;
        aligned = aligned && (((uintptr_t)d.y & 15) == 0) && (((uintptr_t)d.ys & 15) == 0);
    }
    int stride = 0, in_stages = 0;
    SmoothLanesFn fn = smooth_lanes_lookup(window, iters, S, sm_count(), &stride, &in_stages);
    const char *force = std::getenv("MDB_SMOOTH_GENERIC");
    // A few long spectra: the latency form (smooth_stream.cuh), one CTA per spectrum.  Only for launches
    // of up to 32 spectra: its CTAs hold 66 KB of shared memory each for the whole smoothing, which in
    // the chunk pipeline (40+ spectra per chunk, eight chunks in flight) would crowd the FP64 kernels
    // of the other chunks off the SMs; there the lane-packed kernel (10 chains per warp) stays.
    const char *stream_env = std::getenv("MDB_SMOOTH_STREAM");
    size_t stream_max = latency_form_up_to;
    if (const char *env = std::getenv("MDB_STREAM_MAX_SPECTRA")) if (std::atoi(env) >= 1) stream_max = (size_t)std::atoi(env);  // sweeps
    bool streamable = iters >= 1 && iters <= STREAM_MAX_ITERS && window >= 1 && window <= 64 && S <= stream_max
                      && !(stream_env && stream_env[0] == '0') && !(force && force[0] == '1');
    for (const SpecDesc &d : descs) streamable = streamable && d.n >= 4096;
    prof_begin(spans, MDB_KERNEL_SMOOTH, stream);
    if (streamable) {
        // windows 3, 5, 7 and up to six passes: one chain warp per pass, the multiply done by the mover warps (smooth_split.cuh)
        const char *split_env = std::getenv("MDB_SMOOTH_SPLIT");
        if ((window == 3 || window == 5 || window == 7) && iters <= SPLIT_MAX_ITERS && !stream_generic_env()
            && !(split_env && split_env[0] == '0')) {
            auto kern = window == 5 ? smooth_split_kernel<5> : window == 3 ? smooth_split_kernel<3> : smooth_split_kernel<7>;
            const size_t smem = smooth_split_smem_bytes(iters);
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smooth_split_smem_bytes(SPLIT_MAX_ITERS)));
            kern<<<(unsigned)S, smooth_split_threads(iters), smem, stream>>>(d_desc, iters);
            LAUNCH_CHECK();
            prof_end(spans, stream, 16.0 * pts);  // algorithmic bytes: read 8N + write 8N
            return MDB_OK;
        }
        const size_t smem = smooth_stream_smem_bytes(iters);
        // interior loop specialised for the default window and the ones optimize_settings tries (smooth_stream.cuh)
        auto kern = window == 5 ? smooth_stream_kernel<5> : window == 3 ? smooth_stream_kernel<3> : window == 7 ? smooth_stream_kernel<7> : smooth_stream_kernel<0>;
        if (stream_generic_env()) kern = smooth_stream_kernel<0>;  // tests: the any-window form
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)S, STREAM_THREADS, smem, stream>>>(d_desc, iters, window);
        LAUNCH_CHECK();
    } else if (fn && aligned && !(force && force[0] == '1')) {
        const int groups = 32 / iters;  // spectra per warp: one lane per (spectrum, pass)
        const size_t smem = smooth_lanes_smem_bytes(stride, in_stages, groups);
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fn<<<(unsigned)((S + groups - 1) / groups), SL_THREADS, smem, stream>>>(d_desc, (int)S, iters);
        LAUNCH_CHECK();
    } else {
        for (int p = 0; p < iters; ++p) {
            smooth_pass_generic_kernel<<<(unsigned)((S + 31) / 32), 32, 0, stream>>>(d_desc, (int)S, p, iters, window);
            LAUNCH_CHECK();
        }
    }
    prof_end(spans, stream, 16.0 * pts);  // algorithmic bytes: read 8N + write 8N
    return MDB_OK;
}

// True when the intensity rows of the chunk live in page-locked (or managed) host memory.  Rows
// are probed individually (a batch may mix allocations); the probe costs about a microsecond.
static bool host_rows_pinned(const std::vector<HostSpec> &hs, size_t first, size_t count)
{
    for (size_t s = 0; s < count; ++s) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, hs[first + s].y) != cudaSuccess) { cudaGetLastError(); return false; }
        if (attr.type == cudaMemoryTypeUnregistered) return false;
    }
    return true;
}

static thread_local std::chrono::steady_clock::time_point t_pipeline_origin;
static void timeline_stamp(Chunk &ck, int which, cudaStream_t stream)
{
    if (!t_timeline_base) return;
    if (cudaEventCreate(&ck.g_ev[which]) == cudaSuccess) cudaEventRecord(ck.g_ev[which], stream);
}
static double pipeline_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pipeline_origin).count();
}

// Stage A: inputs -> device, smoothing, detection, selection, counts back to the host.
static mdb_status stage_a(Chunk &ck, const std::vector<HostSpec> &hs, const mdb_deconvoluter &dc, int memory,
                          bool skip_smoothing_input_is_smoothed)
{
    Workspace &ws = *ck.ws;
    const size_t S = ck.count;
    // stage A's kernels run on `sa`; the input copies go through `sc` -- ONE copy stream per pipeline, so
    // that the chunks' H2D copies happen in chunk order (copies of different streams are interleaved by
    // the hardware in no particular order: chunk 1's rows were seen landing after chunk 3's), and `sa`
    // waits for them through an event
    const cudaStream_t sa = ck.s_a ? ck.s_a : ws.stream_a;
    const cudaStream_t sc = ck.s_copy ? ck.s_copy : sa;
    // ---- layout
    size_t y_elems = 0, cand_elems = 0, tile_elems = 0, ig_elems = 0;
    std::vector<size_t> y_off(S), cand_off(S), tile_off(S), ig_off(S);
    // (caller x pointer, length) -> element offset in ws.x.  The length is part of the key: two views of
    // one axis array (big[:1000], big[:2000]) share the pointer but not the row.
    std::map<std::pair<const double *, size_t>, size_t> x_map;
    size_t x_elems = 0;
    ck.max_tiles = 0;
    for (size_t s = 0; s < S; ++s) {
        const HostSpec &h = hs[ck.first + s];
        y_off[s] = y_elems;
        y_elems += align_up(h.n, 16);
        const size_t tiles = (h.n + DETECT_TILE - 1) / DETECT_TILE;
        cand_off[s] = cand_elems;
        cand_elems += tiles * DETECT_CAP;  // candidate capacity: DETECT_CAP records per detection tile
        tile_off[s] = tile_elems;
        tile_elems += align_up(tiles, 4);
        ig_off[s] = ig_elems;
        ig_elems += h.ig.size();
        ck.max_tiles = std::max(ck.max_tiles, (int)tiles);
        if (memory == MDB_MEM_HOST && !x_map.count({h.x, h.n})) {
            x_map[{h.x, h.n}] = x_elems;
            x_elems += align_up(h.n, 16);
        }
    }
    const bool preset = hs[ck.first].ys_dev != nullptr;  // smoothed rows supplied by the caller (all or none)
    const bool ma = dc.smoothing.kind == MDB_SMOOTHING_MOVING_AVERAGE && !skip_smoothing_input_is_smoothed && !preset;
    const bool need_tmp = ma && dc.smoothing.iterations >= 2;
    if (memory == MDB_MEM_HOST) {
        CUDA_TRY(ws.x.ensure(x_elems * 8));
        CUDA_TRY(ws.y.ensure(y_elems * 8));
    }
    if (!preset) CUDA_TRY(ws.ys.ensure(y_elems * 8));
    if (need_tmp) CUDA_TRY(ws.tmp.ensure(y_elems * 8));
    CUDA_TRY(ws.tile_cnt.ensure(tile_elems * 4));
    CUDA_TRY(ws.pk.ensure(cand_elems * 12));
    CUDA_TRY(ws.sc.ensure(cand_elems * 8));
    CUDA_TRY(ws.sfr.ensure(cand_elems * 8));
    CUDA_TRY(ws.sel.ensure(cand_elems * 12));
    CUDA_TRY(ws.ig.ensure(std::max<size_t>(ig_elems, 1) * 4));
    CUDA_TRY(ws.desc.ensure(S * sizeof(SpecDesc)));
    CUDA_TRY(ws.sel_out.ensure(S * sizeof(SelectOut)));
    CUDA_TRY(ws.h_desc.ensure(S * sizeof(SpecDesc)));
    CUDA_TRY(ws.h_ig.ensure(std::max<size_t>(ig_elems, 1) * 4));
    CUDA_TRY(ws.h_sel_out.ensure(S * sizeof(SelectOut)));

    // ---- descriptors
    ck.desc.resize(S);
    int *h_ig = ws.h_ig.as<int>();
    for (size_t s = 0; s < S; ++s) {
        const HostSpec &h = hs[ck.first + s];
        SpecDesc &d = ck.desc[s];
        if (memory == MDB_MEM_HOST) {
            d.x = ws.x.as<double>() + x_map[{h.x, h.n}];
            d.y = ws.y.as<double>() + y_off[s];
        } else {
            d.x = h.x;
            d.y = h.y;
        }
        d.ys = h.ys_dev ? const_cast<double *>(h.ys_dev) : ws.ys.as<double>() + y_off[s];
        d.threshold = h.threshold;
        d.tmp = need_tmp ? ws.tmp.as<double>() + y_off[s] : nullptr;
        d.tile_cnt = ws.tile_cnt.as<int>() + tile_off[s];
        d.pk = ws.pk.as<int>() + 3 * cand_off[s];
        d.sc = ws.sc.as<double>() + cand_off[s];
        d.sfr = ws.sfr.as<double>() + cand_off[s];
        d.sel = ws.sel.as<int>() + 3 * cand_off[s];
        d.ig = ws.ig.as<int>() + ig_off[s];
        d.n = (int)h.n;
        d.n_tiles = (int)((h.n + DETECT_TILE - 1) / DETECT_TILE);
        d.sb0 = h.sb_i0;
        d.sb1 = h.sb_i1;
        d.n_ig = (int)(h.ig.size() / 2);
        d.has_ig = dc.has_ignore ? 1 : 0;
        for (size_t q = 0; q < h.ig.size(); ++q) h_ig[ig_off[s] + q] = h.ig[q];
    }
    std::memcpy(ws.h_desc.p, ck.desc.data(), S * sizeof(SpecDesc));
    CUDA_TRY(counted_memcpy_async(ws.desc.p, ws.h_desc.p, S * sizeof(SpecDesc), cudaMemcpyHostToDevice, sc));
    if (ig_elems)
        CUDA_TRY(counted_memcpy_async(ws.ig.p, ws.h_ig.p, ig_elems * 4, cudaMemcpyHostToDevice, sc));

    // ---- inputs to the device (rows that are adjacent on the host and on the device go as one copy)
    double *y_dst = skip_smoothing_input_is_smoothed ? ws.ys.as<double>() : ws.y.as<double>();
    if (memory == MDB_MEM_HOST) {
        for (auto &kv : x_map)
            CUDA_TRY(counted_memcpy_async(ws.x.as<double>() + kv.second, kv.first.first, kv.first.second * 8,
                                          cudaMemcpyHostToDevice, sc));
        if (ck.stage_job) {
            // pageable rows, gathered ahead of time by the staging threads (prepare_staging): one DMA per part
            StageJob &job = *ck.stage_job;
            size_t row0 = 0;
            for (size_t pt = 0; pt < job.part_end.size(); ++pt) {
                job.wait_part(pt);
                const size_t row1 = job.part_end[pt];
                const size_t e0 = y_off[row0], e1 = row1 < S ? y_off[row1] : y_elems;
                CUDA_TRY(counted_memcpy_async(y_dst + e0, job.dst + e0, (e1 - e0) * 8, cudaMemcpyHostToDevice, sc));
                row0 = row1;
            }
            ck.stage_job.reset();
        } else if (!host_rows_pinned(hs, ck.first, S)) {
            // pageable rows of a single-chunk entry point (the stage_* functions): gather on this thread
            CUDA_TRY(ws.h_stage.ensure(y_elems * 8));
            double *stage = ws.h_stage.as<double>();
            CUDA_TRY(cudaStreamSynchronize(sc));  // the previous DMA out of this staging area has finished
            for (size_t s = 0; s < S; ++s)
                mdb_host_row_copy(stage + y_off[s], hs[ck.first + s].y, hs[ck.first + s].n);
            CUDA_TRY(counted_memcpy_async(y_dst, stage, y_elems * 8, cudaMemcpyHostToDevice, sc));
        } else {
            size_t s = 0;
            while (s < S) {
                size_t e = s + 1;
                const HostSpec &h0 = hs[ck.first + s];
                size_t bytes = h0.n * 8;
                while (e < S) {
                    const HostSpec &hp = hs[ck.first + e - 1], &hn = hs[ck.first + e];
                    if (hp.n % 16 == 0 && hn.y == hp.y + hp.n) { bytes += hn.n * 8; ++e; }
                    else break;
                }
                cudaError_t ce = counted_memcpy_async(y_dst + y_off[s], h0.y, bytes, cudaMemcpyHostToDevice, sc);
                if (ce == cudaErrorInvalidValue && e > s + 1) {
                    // adjacent in the address space but not one page-locked allocation (separately allocated
                    // rows that happen to touch): a copy may not span two allocations -- row by row instead
                    cudaGetLastError();
                    g_h2d_bytes -= bytes;
                    for (size_t q = s; q < e; ++q)
                        CUDA_TRY(counted_memcpy_async(y_dst + y_off[q], hs[ck.first + q].y, hs[ck.first + q].n * 8,
                                                      cudaMemcpyHostToDevice, sc));
                } else {
                    CUDA_TRY(ce);
                }
                s = e;
            }
        }
    } else if (skip_smoothing_input_is_smoothed) {
        for (size_t s = 0; s < S; ++s)
            CUDA_TRY(counted_memcpy_async(y_dst + y_off[s], hs[ck.first + s].y, hs[ck.first + s].n * 8,
                                     cudaMemcpyDeviceToDevice, sc));
    }

    const SpecDesc *d_desc = ws.desc.as<SpecDesc>();
    if (sc != sa) {  // hand over from the copy stream to this chunk's stage-A stream
        CUDA_TRY(cudaEventRecord(ck.ev_in, sc));
        CUDA_TRY(cudaStreamWaitEvent(sa, ck.ev_in, 0));
    }
    timeline_stamp(ck, 0, sa);
    // ---- K1 smoothing (deconvoluter.rs:531-532)
    if (!skip_smoothing_input_is_smoothed && !preset) {
        if (ma) {
            // The first chunks of a call smooth on an idle GPU: there the latency form (1.4 ms per launch instead of
            // 2.9, one CTA per spectrum) brings the first peak counts -- and with them the first FP64 work -- forward;
            // later chunks keep the lane-packed kernel, whose few warps do not crowd the fit / MSE kernels.
            const char *su = std::getenv("MDB_STARTUP_STREAM");
            const size_t up_to = (ck.startup && !(su && su[0] == '0')) ? 80 : 32;
            mdb_status sst = launch_smooth(sa, d_desc, ck.desc, (int)dc.smoothing.iterations,
                                           (int)dc.smoothing.window_size, &ck.spans, up_to);
            if (sst != MDB_OK) return sst;
        } else {  // Identity (smoothing/identity.rs): the smoothed copy is the input itself
            for (size_t s = 0; s < S; ++s)
                CUDA_TRY(counted_memcpy_async(ck.desc[s].ys, ck.desc[s].y, (size_t)ck.desc[s].n * 8,
                                         cudaMemcpyDeviceToDevice, sa));
        }
    }
    timeline_stamp(ck, 1, sa);
    // ---- K2/K3 detection + scoring
    {
        dim3 grid((unsigned)((ck.max_tiles + DETECT_WARPS - 1) / DETECT_WARPS), (unsigned)S);
        double pts = 0.0;
        for (size_t s = 0; s < S; ++s) pts += (double)ck.desc[s].n;
        prof_begin(&ck.spans, MDB_KERNEL_DETECT, sa);
        detect_kernel<<<grid, DETECT_THREADS, 0, sa>>>(d_desc);
        LAUNCH_CHECK();
        prof_end(&ck.spans, sa, 8.0 * pts);  // algorithmic bytes: read 8N
    }
    // ---- K4 selection
    {
        const size_t smem = (size_t)4 * ck.max_tiles * sizeof(int);
        if ((int)smem + 1024 > smem_optin_limit())
            return fail(MDB_ERR_UNSUPPORTED, "spectrum too long for the selection kernel's shared-memory masks");
        const bool few = S <= 8;  // few spectra: 32 warps per spectrum instead of 8
        auto kern = few ? select_kernel<1024> : select_kernel<SELECT_THREADS>;
        if (smem > 40 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(&ck.spans, MDB_KERNEL_SELECT, sa);
        // MDB_ZEROCOPY_COUNTS=1 (experiment): the kernel stores the counts straight into the mapped pinned
        // buffer instead of a D2H copy behind it
        static const bool zero_copy_counts = std::getenv("MDB_ZEROCOPY_COUNTS") && std::getenv("MDB_ZEROCOPY_COUNTS")[0] == '1';
        SelectOut *so_dst = ws.sel_out.as<SelectOut>();
        if (zero_copy_counts) CUDA_TRY(cudaHostGetDevicePointer((void **)&so_dst, ws.h_sel_out.p, 0));
        kern<<<(unsigned)S, few ? 1024 : SELECT_THREADS, smem, sa>>>(d_desc, so_dst, dc.selection.kind);
        LAUNCH_CHECK();
        prof_end(&ck.spans, sa, (double)S);
        if (zero_copy_counts) count_transfer(S * sizeof(SelectOut), cudaMemcpyDeviceToHost);
        else CUDA_TRY(counted_memcpy_async(ws.h_sel_out.p, ws.sel_out.p, S * sizeof(SelectOut), cudaMemcpyDeviceToHost, sa));
    }
    CUDA_TRY(cudaEventRecord(ws.ev_a, sa));
    timeline_stamp(ck, 2, sa);
    ck.t_a = pipeline_ms();
    return MDB_OK;
}

static void fit_state_pointers(Workspace &ws, long long p_total, FitState &st)
{
    double *base = ws.fit_state.as<double>();
    const size_t stride = align_up((size_t)std::max<long long>(p_total, 1), 16);
    double **fields[] = {&st.ox1, &st.ox2, &st.ox3, &st.oy1, &st.oy2, &st.oy3, &st.sx1, &st.sx3, &st.sy1, &st.sy2, &st.sy3};
    for (size_t i = 0; i < 11; ++i) *fields[i] = base + i * stride;
    st.pa = ws.par_a.as<double>();
    st.pb = ws.par_b.as<double>();
}

// Stage B: fit, retain, MSE, results back to the host.  `trace` (optional, device->host copies
// after every pass) is used by mdb_stage_fit only.  `with_mse` false skips K7.
static mdb_status stage_b(Chunk &ck, const std::vector<HostSpec> &hs, const mdb_deconvoluter &dc,
                          std::vector<SpecResult> &results, bool with_mse, mdb_lorentzian *trace)
{
    Workspace &ws = *ck.ws;
    const size_t S = ck.count;
    const cudaStream_t sb = ck.s_b ? ck.s_b : ws.stream;
    CUDA_TRY(cudaEventSynchronize(ws.ev_a));
    ck.t_counts = pipeline_ms();
    timeline_stamp(ck, 3, sb);
    const SelectOut *so = ws.h_sel_out.as<SelectOut>();
    ck.fdesc.assign(S, FitDesc{});
    ck.segs.clear();
    ck.p_total = 0;
    ck.est_evals = 0.0;
    ck.res_total = 0;
    ck.max_peaks = 0;
    ck.max_seg_len = 0;
    for (size_t s = 0; s < S; ++s) {
        SpecResult &r = results[ck.first + s];
        const HostSpec &h = hs[ck.first + s];
        r.info = so[s];
        r.status = (h.pre_status != MDB_OK) ? h.pre_status : so[s].status;
        FitDesc &f = ck.fdesc[s];
        f.off = ck.p_total;
        f.n_peaks = (r.status == MDB_OK) ? so[s].n_selected : 0;
        f.seg_off = (int)ck.segs.size();
        f.seg_cnt = 0;
        f.n_iters = h.fit_iters;
        if (r.status == MDB_OK && with_mse) {
            for (auto &rg : h.ranges) {
                Segment sg;
                sg.spec = (int)s; sg.start = rg.first; sg.end = rg.second; sg.pad_ = 0;
                sg.res_off = 0;  // assigned below, once the launch shape of K7 is known
                ck.max_seg_len = std::max(ck.max_seg_len, rg.second - rg.first);
                ck.segs.push_back(sg);
                ++f.seg_cnt;
            }
        }
        {
            double seg_pts = 0.0;
            if (r.status == MDB_OK) for (auto &rg : h.ranges) seg_pts += (double)(rg.second - rg.first);
            ck.est_evals += (double)f.n_peaks * (3.0 * (double)f.n_iters * (double)f.n_peaks + seg_pts);
        }
        ck.p_total += (f.n_peaks + 1) & ~1;  // even offsets: every parameter block stays 16-byte aligned (TMA bulk copies)
        ck.max_peaks = std::max(ck.max_peaks, f.n_peaks);
    }
    const size_t P = (size_t)std::max<long long>(ck.p_total, 1);
    const size_t n_seg = ck.segs.size();
    // K7 launch shape.  R = 8 points per thread is the throughput shape; small batches use R = 2 to fill
    // the SMs; the few-ulp form runs best with 16 points per thread (tools/kbench.cu, KBENCH_ULP=1) once
    // that still fills the GPU.  In the few-ulp mode a CTA hands on ONE sum of squared residuals (MODE 2
    // of superposition_kernel), so a range owns ceil(len / per_block) slots of `resid` instead of len.
    const bool ulp = dc.sup_mode == MDB_SUPERPOSITION_FAST;
    long long blocks8 = 0;
    for (const Segment &sg : ck.segs) blocks8 += (sg.end - sg.start + SUP_THREADS * 8 - 1) / (SUP_THREADS * 8);
    const int sup_r = (ulp && blocks8 >= 8 * sm_count() && !std::getenv("MDB_SUP_R8")) ? 16 : blocks8 >= 2 * sm_count() ? 8 : 2;
    const int per_block = SUP_THREADS * sup_r;
    for (Segment &sg : ck.segs) {
        sg.res_off = ck.res_total;
        const size_t len = (size_t)(sg.end - sg.start);
        ck.res_total += (long long)(ulp ? (len + per_block - 1) / per_block : align_up(len, 16));
    }
    CUDA_TRY(ws.fdesc.ensure(S * sizeof(FitDesc)));
    CUDA_TRY(ws.h_fdesc.ensure(S * sizeof(FitDesc)));
    CUDA_TRY(ws.segs.ensure(std::max<size_t>(n_seg, 1) * sizeof(Segment)));
    CUDA_TRY(ws.h_segs.ensure(std::max<size_t>(n_seg, 1) * sizeof(Segment)));
    CUDA_TRY(ws.fit_state.ensure(14 * align_up(P, 16) * 8));  // 11 stencil rows + 3 rows of rescaled y (fit_wide.cuh)
    CUDA_TRY(ws.par_a.ensure(P * 24));
    CUDA_TRY(ws.par_b.ensure(P * 24));
    CUDA_TRY(ws.lor.ensure(P * 24));
    CUDA_TRY(ws.peaks_dense.ensure(P * 12));
    CUDA_TRY(ws.n_kept.ensure(S * 4));
    CUDA_TRY(ws.mse.ensure(S * 8));
    CUDA_TRY(ws.resid.ensure(std::max<size_t>((size_t)ck.res_total, 1) * 8));
    CUDA_TRY(ws.h_lor.ensure(P * 24));
    CUDA_TRY(ws.h_peaks.ensure(P * 12));
    CUDA_TRY(ws.h_n_kept.ensure(S * 4));
    CUDA_TRY(ws.h_mse.ensure(S * 8));
    // The two descriptor tables go up through a tiny KERNEL that reads the page-locked host copies over
    // PCIe (zero-copy), not through cudaMemcpyAsync: a copy command of this low-priority stream would sit
    // in the H2D engine's queue behind hundreds of megabytes of input rows of the chunks ahead (measured:
    // no stage-B kernel ran before the whole batch had been copied in).
    std::memcpy(ws.h_fdesc.p, ck.fdesc.data(), S * sizeof(FitDesc));
    if (n_seg) std::memcpy(ws.h_segs.p, ck.segs.data(), n_seg * sizeof(Segment));
    {
        static_assert(sizeof(FitDesc) % 8 == 0 && sizeof(Segment) % 8 == 0, "descriptor tables are copied in 8-byte words");
        const unsigned long long *src_f = nullptr, *src_s = nullptr;
        CUDA_TRY(cudaHostGetDevicePointer((void **)&src_f, ws.h_fdesc.p, 0));
        CUDA_TRY(cudaHostGetDevicePointer((void **)&src_s, ws.h_segs.p, 0));
        const size_t wf = S * sizeof(FitDesc) / 8, wsg = n_seg * sizeof(Segment) / 8;
        upload_tables_kernel<<<(unsigned)std::min<size_t>((wf + wsg + 255) / 256, 64), 256, 0, sb>>>(
            src_f, ws.fdesc.as<unsigned long long>(), wf, src_s, ws.segs.as<unsigned long long>(), wsg);
        LAUNCH_CHECK();
        count_transfer((wf + wsg) * 8, cudaMemcpyHostToDevice);
    }
    const SpecDesc *d_desc = ws.desc.as<SpecDesc>();
    const FitDesc *d_fd = ws.fdesc.as<FitDesc>();
    FitState st;
    fit_state_pointers(ws, ck.p_total, st);
    int iters = 0;  // refinement launches: the largest per-spectrum count of the chunk
    for (size_t s = 0; s < S; ++s) iters = std::max(iters, ck.fdesc[s].n_peaks > 0 ? ck.fdesc[s].n_iters : 0);
    if (ck.p_total > 0) {
        dim3 grid((unsigned)((ck.max_peaks + FIT_THREADS - 1) / FIT_THREADS), (unsigned)S);
        prof_begin(&ck.spans, MDB_KERNEL_FIT_INIT, sb);
        fit_init_kernel<<<grid, FIT_THREADS, 0, sb>>>(d_desc, d_fd, st, ws.peaks_dense.as<int>());
        LAUNCH_CHECK();
        prof_end(&ck.spans, sb, (double)ck.p_total);
        // trace is a single-spectrum facility (mdb_stage_fit): p_total carries alignment padding, n_peaks does not
        const size_t n_trace = (size_t)ck.fdesc[0].n_peaks;
        if (trace) {
            CUDA_TRY(counted_memcpy_async(trace, st.pa, n_trace * 24, cudaMemcpyDeviceToHost, sb));
        }
        // Default: one launch per refinement pass.  MDB_FIT_PERSISTENT=1 selects the single-launch
        // work-queue form (fit_persistent_kernel); measured 10 % slower per chunk and 6.5 % slower end
        // to end (it holds every SM slot for the whole fit, which starves the other chunk streams),
        // so it stays an experiment.
        const char *persistent = std::getenv("MDB_FIT_PERSISTENT");
        // A few spectra: one pass = superposition per (stencil point, peak) + solve per peak
        // (fit_wide.cuh); three times the CTAs and 8 Lorentzians in flight per thread.
        long long fit_blocks = 0;
        for (size_t s = 0; s < S; ++s) fit_blocks += (ck.fdesc[s].n_peaks + FIT_THREADS - 1) / FIT_THREADS;
        const char *wide_env = std::getenv("MDB_FIT_WIDE");
        // (the opt-in arithmetic experiments exist in the one-thread-per-peak kernel only)
        const bool wide = fit_blocks <= sm_count() && !(wide_env && wide_env[0] == '0') && !(persistent && persistent[0] == '1')
                          && dc.fit_arith == MDB_FIT_EXACT;
        if (wide) {
            const long long yn_stride = (long long)align_up(P, 16);
            double *yn = ws.fit_state.as<double>() + 11 * yn_stride;
            dim3 grid3(grid.x, grid.y, 3);
            // MDB_FIT_WIDE=1 selects the instruction-parallel form (8 Lorentzians per thread), anything else
            // the thread-parallel one (producers / accumulators)
            // (its 256-thread CTAs pay off while they all fit the GPU at once: one blood spectrum is 93 of them;
            // sixteen spectra are ~1 500, and the 128-thread instruction-parallel form is faster again)
            long long wide2_ctas = 0;
            for (size_t s = 0; s < S; ++s) wide2_ctas += 3ll * ((ck.fdesc[s].n_peaks + WIDE2_CHAINS - 1) / WIDE2_CHAINS);
            const bool wide2 = !(wide_env && wide_env[0] == '1') && wide2_ctas <= 4ll * sm_count();
            // the whole parameter set of a spectrum behind the quotient buffers when it fits (up to ~4 000 peaks)
            const size_t wide2_params = align_up((size_t)std::max(ck.max_peaks, 1) * 24, 16);
            const bool wide2_psm = wide2 && WIDE2_SMEM + wide2_params <= (size_t)160 * 1024 && !std::getenv("MDB_WIDE2_GLOBAL_PARAMS");
            const size_t wide2_smem = WIDE2_SMEM + (wide2_psm ? wide2_params : 0);
            if (wide2)
                CUDA_TRY(cudaFuncSetAttribute(wide2_psm ? fit_wide2_superpose_kernel<true> : fit_wide2_superpose_kernel<false>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide2_smem));
            // counters for the solve step fused into the producer / accumulator kernel (MDB_WIDE2_FUSED_SOLVE=0: separate launch)
            int *wide2_done = nullptr;
            if (wide2 && !(std::getenv("MDB_WIDE2_FUSED_SOLVE") && std::getenv("MDB_WIDE2_FUSED_SOLVE")[0] == '0')) {
                const size_t n_counters = S * (size_t)((ck.max_peaks + WIDE2_CHAINS - 1) / WIDE2_CHAINS);
                CUDA_TRY(ws.fit_queue.ensure(n_counters * 4));
                CUDA_TRY(cudaMemsetAsync(ws.fit_queue.p, 0, n_counters * 4, sb));
                wide2_done = ws.fit_queue.as<int>();
            }
            for (int it = 0; it < iters; ++it) {
                double evals = 0.0;
                for (size_t s = 0; s < S; ++s)
                    if (it < ck.fdesc[s].n_iters) evals += 3.0 * (double)ck.fdesc[s].n_peaks * (double)ck.fdesc[s].n_peaks;
                prof_begin(&ck.spans, MDB_KERNEL_FIT_ITER, sb);
                if (wide2) {
                    dim3 g2((unsigned)((ck.max_peaks + WIDE2_CHAINS - 1) / WIDE2_CHAINS), (unsigned)S, 3);
                    if (wide2_psm) fit_wide2_superpose_kernel<true><<<g2, WIDE2_THREADS, wide2_smem, sb>>>(d_fd, st, yn, yn_stride, it, wide2_done);
                    else fit_wide2_superpose_kernel<false><<<g2, WIDE2_THREADS, wide2_smem, sb>>>(d_fd, st, yn, yn_stride, it, wide2_done);
                } else {
                    fit_wide_superpose_kernel<<<grid3, FIT_THREADS, LOR_SMEM_BYTES, sb>>>(d_fd, st, yn, yn_stride, it);
                }
                LAUNCH_CHECK();
                if (!wide2_done) {  // (the producer / accumulator kernel solves in its last CTA per peak block)
                    fit_wide_solve_kernel<<<grid, FIT_THREADS, 0, sb>>>(d_fd, st, yn, yn_stride, it);
                    LAUNCH_CHECK();
                }
                prof_end(&ck.spans, sb, evals);
                if (trace)
                    CUDA_TRY(counted_memcpy_async(trace + (size_t)(it + 1) * n_trace, (it & 1) ? st.pa : st.pb, n_trace * 24,
                                             cudaMemcpyDeviceToHost, sb));
            }
        } else if (!trace && !(persistent && persistent[0] == '1') && iters > 0 && ck.max_peaks <= FIT_BLOCK_MAX && fit_block_enabled()) {
            // spectra with up to 1 024 selected peaks: one CTA per spectrum, every pass in ONE launch (fit_block_kernel)
            double evals = 0.0;
            for (size_t s = 0; s < S; ++s) evals += 3.0 * (double)ck.fdesc[s].n_iters * (double)ck.fdesc[s].n_peaks * (double)ck.fdesc[s].n_peaks;
            const int threads = (int)align_up((size_t)std::max(ck.max_peaks, 1), 32);
            const size_t smem = fit_block_smem_bytes(threads);
            // launch-bound classes: up to 320 threads x 3 CTAs per SM, up to 576 x 2, up to 1 024 x 1
            auto kern = dc.fit_arith == MDB_FIT_ULP ? fit_block_kernel<2, FIT_BLOCK_MAX, 1> : dc.fit_arith == MDB_FIT_CORRECTED ? fit_block_kernel<3, FIT_BLOCK_MAX, 1> : fit_block_kernel<1, FIT_BLOCK_MAX, 1>;
            if (dc.fit_arith == MDB_FIT_EXACT && threads <= 320) kern = fit_block_kernel<1, 320, 3>;
            else if (dc.fit_arith == MDB_FIT_EXACT && threads <= 576) kern = fit_block_kernel<1, 576, 2>;
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fit_block_smem_bytes(FIT_BLOCK_MAX)));
            prof_begin(&ck.spans, MDB_KERNEL_FIT_ITER, sb);
            kern<<<(unsigned)S, threads, smem, sb>>>(d_fd, st);
            LAUNCH_CHECK();
            prof_end(&ck.spans, sb, evals);
        } else if (trace || !(persistent && persistent[0] == '1')) {
            // one launch per refinement pass (needed for the per-pass trace of mdb_stage_fit)
            for (int it = 0; it < iters; ++it) {
                double evals = 0.0;  // E_fit of this pass = sum of 3 * P_s^2 over the spectra still iterating
                for (size_t s = 0; s < S; ++s)
                    if (it < ck.fdesc[s].n_iters) evals += 3.0 * (double)ck.fdesc[s].n_peaks * (double)ck.fdesc[s].n_peaks;
                prof_begin(&ck.spans, MDB_KERNEL_FIT_ITER, sb);
                // CTA shape: one warp per CTA (32 peaks, 128-Lorentzian tiles) for spectra with up to FIT_WARP_CTA_MAX
                // peaks, 128 peaks x 512-Lorentzian tiles above (see fit_iter_kernel); MDB_FIT_CTA=32|128 forces one
                int cta = ck.max_peaks <= FIT_WARP_CTA_MAX ? 32 : FIT_THREADS;
                if (const char *env = std::getenv("MDB_FIT_CTA")) { const int v = std::atoi(env); if (v == 32 || v == FIT_THREADS) cta = v; }
                auto fit_kern = dc.fit_arith == MDB_FIT_ULP ? fit_iter_kernel<2> : dc.fit_arith == MDB_FIT_CORRECTED ? fit_iter_kernel<3> : fit_iter_kernel<1>;
                if (cta == 32)
                    fit_kern = dc.fit_arith == MDB_FIT_ULP ? fit_iter_kernel<2, 32, 128, 32> : dc.fit_arith == MDB_FIT_CORRECTED ? fit_iter_kernel<3, 32, 128, 32> : fit_iter_kernel<1, 32, 128, 32>;
                size_t fit_smem = cta == 32 ? fit_smem_bytes(128) : LOR_SMEM_BYTES;
                if (const char *occ = std::getenv("MDB_FIT_OCC")) {  // experiment: 12 CTAs of 128 threads per SM (<= 40 registers, 256-Lorentzian tiles)
                    if (occ[0] == '1' && cta == FIT_THREADS && dc.fit_arith == MDB_FIT_EXACT) { fit_kern = fit_iter_kernel<1, 128, 256, 12>; fit_smem = fit_smem_bytes(256); }
                }
                const dim3 grid_sf((unsigned)S, (unsigned)((ck.max_peaks + cta - 1) / cta));  // spectrum fastest, peak block slowest
                fit_kern<<<grid_sf, cta, fit_smem, sb>>>(d_fd, st, it);
                LAUNCH_CHECK();
                prof_end(&ck.spans, sb, evals);
                if (trace)
                    CUDA_TRY(counted_memcpy_async(trace + (size_t)(it + 1) * n_trace, (it & 1) ? st.pa : st.pb, n_trace * 24,
                                             cudaMemcpyDeviceToHost, sb));
            }
        } else if (iters > 0) {
            // all passes in ONE persistent launch: work items through an atomic queue, per-(spectrum,
            // pass) completion counters instead of launch boundaries (fit_persistent_kernel)
            CUDA_TRY(ws.blk_off.ensure((S + 1) * 4));
            CUDA_TRY(ws.h_blk_off.ensure((S + 1) * 4));
            int *h_blk = ws.h_blk_off.as<int>();
            double evals = 0.0;
            h_blk[0] = 0;
            for (size_t s = 0; s < S; ++s) {
                h_blk[s + 1] = h_blk[s] + (ck.fdesc[s].n_peaks + FIT_THREADS - 1) / FIT_THREADS;
                evals += 3.0 * (double)ck.fdesc[s].n_iters * (double)ck.fdesc[s].n_peaks * (double)ck.fdesc[s].n_peaks;
            }
            const int blocks_per_pass = h_blk[S];
            const size_t n_counters = 1 + S * (size_t)iters;
            CUDA_TRY(ws.fit_queue.ensure(n_counters * 4));
            CUDA_TRY(counted_memcpy_async(ws.blk_off.p, ws.h_blk_off.p, (S + 1) * 4, cudaMemcpyHostToDevice, sb));
            CUDA_TRY(cudaMemsetAsync(ws.fit_queue.p, 0, n_counters * 4, sb));
            FitQueue q;
            q.next_item = ws.fit_queue.as<int>();
            q.done = ws.fit_queue.as<int>() + 1;
            q.blk_off = ws.blk_off.as<int>();
            q.n_spec = (int)S;
            q.max_iters = iters;
            q.blocks_per_pass = blocks_per_pass;
            const long long items = (long long)blocks_per_pass * iters;
            const unsigned pgrid = (unsigned)std::min<long long>(items, (long long)sm_count() * 8);
            if (pgrid > 0) {
                prof_begin(&ck.spans, MDB_KERNEL_FIT_ITER, sb);
                fit_persistent_kernel<<<pgrid, FIT_THREADS, LOR_SMEM_BYTES, sb>>>(d_fd, st, q);
                LAUNCH_CHECK();
                prof_end(&ck.spans, sb, evals);
            }
        }
    }
    prof_begin(&ck.spans, MDB_KERNEL_RETAIN, sb);
    retain_kernel<<<(unsigned)S, RETAIN_THREADS, 0, sb>>>(d_fd, st.pa, st.pb, ws.lor.as<double>(), ws.n_kept.as<int>());
    LAUNCH_CHECK();
    prof_end(&ck.spans, sb, (double)ck.p_total);
    // The MSE superposition and everything after it run on the chunk's LOWEST-priority stream `sm`, the
    // refinement passes above on a medium-priority one: a refinement pass is a short launch the next pass
    // waits for, K7 is one long launch with thousands of CTAs.  With both at one priority the SMs ran
    // one kind at a time (CTAs are dispatched grid after grid); now the pending pass takes the slots
    // that free up and K7's CTAs -- 16 independent chains per thread -- fill the FP64-pipe bubbles the
    // three-chain refinement warps leave on the same SM.
    const cudaStream_t sm = (ck.s_m && with_mse) ? ck.s_m : sb;
    if (sm != sb) {
        CUDA_TRY(cudaEventRecord(ck.ev_fit, sb));
        CUDA_TRY(cudaStreamWaitEvent(sm, ck.ev_fit, 0));
    }
    if (with_mse && n_seg) {
        const int r = sup_r;
        dim3 grid((unsigned)((ck.max_seg_len + per_block - 1) / per_block), (unsigned)n_seg);
        if (grid.x > 0) {
            prof_begin(&ck.spans, MDB_KERNEL_MSE_SUPERPOSITION, sm);
            auto kern = ulp ? (sup_quad() ? (r == 16 ? superposition_kernel<2, 16, 4> : r == 8 ? superposition_kernel<2, 8, 4> : superposition_kernel<2, 2, 4>)
                                          : (r == 16 ? superposition_kernel<2, 16, 2> : r == 8 ? superposition_kernel<2, 8, 2> : superposition_kernel<2, 2, 2>))
                            : (r == 8 ? superposition_kernel<1, 8, 1> : superposition_kernel<1, 2, 1>);
            kern<<<grid, SUP_THREADS, LOR_SMEM_BYTES, sm>>>(nullptr, 0, nullptr, 0, ws.resid.as<double>(), d_desc, d_fd,
                                                                  ws.segs.as<Segment>(), ws.lor.as<double>(),
                                                                  ws.n_kept.as<int>());
            LAUNCH_CHECK();
            prof_end(&ck.spans, sm, -1.0);  // work = sum(points * kept), known in finish_chunk
        }
        prof_begin(&ck.spans, MDB_KERNEL_MSE_REDUCE, sm);
        if (ulp)  // few-ulp mode: fold the CTA sums K7 left behind (mse_partials_kernel)
            mse_partials_kernel<<<(unsigned)((S + MSE_PART_WARPS - 1) / MSE_PART_WARPS), 32 * MSE_PART_WARPS, 0, sm>>>(
                d_fd, ws.segs.as<Segment>(), ws.resid.as<double>(), ws.mse.as<double>(), (int)S, per_block);
        else
            mse_reduce_kernel<<<(unsigned)((S + MSE_WARPS - 1) / MSE_WARPS), 32 * MSE_WARPS, 0, sm>>>(d_fd, ws.segs.as<Segment>(),
                                                                                  ws.resid.as<double>(), ws.mse.as<double>(), (int)S);
        LAUNCH_CHECK();
        prof_end(&ck.spans, sm, 8.0 * (double)ck.res_total);
        CUDA_TRY(counted_memcpy_async(ws.h_mse.p, ws.mse.p, S * 8, cudaMemcpyDeviceToHost, sm));
    }
    CUDA_TRY(counted_memcpy_async(ws.h_n_kept.p, ws.n_kept.p, S * 4, cudaMemcpyDeviceToHost, sm));
    if (ck.p_total > 0) {
        CUDA_TRY(counted_memcpy_async(ws.h_lor.p, ws.lor.p, (size_t)ck.p_total * 24, cudaMemcpyDeviceToHost, sm));
        CUDA_TRY(counted_memcpy_async(ws.h_peaks.p, ws.peaks_dense.p, (size_t)ck.p_total * 12, cudaMemcpyDeviceToHost, sm));
    }
    CUDA_TRY(cudaEventRecord(ws.ev_b, sm));
    timeline_stamp(ck, 4, sm);
    ck.stage_b_launched = true;
    ck.t_b = pipeline_ms();
    return MDB_OK;
}

// Wait for stage B and unpack into the batch.
static mdb_status finish_chunk(Chunk &ck, std::vector<SpecResult> &results, bool with_mse)
{
    Workspace &ws = *ck.ws;
    CUDA_TRY(cudaEventSynchronize(ws.ev_b));
    ck.t_done = pipeline_ms();
    const int *kept = ws.h_n_kept.as<int>();
    const double *mse = ws.h_mse.as<double>();
    const mdb_lorentzian *lor = ws.h_lor.as<mdb_lorentzian>();
    const int32_t *pk = ws.h_peaks.as<int32_t>();
    if (!ck.spans.empty()) {
        double mse_evals = 0.0;  // E_mse = sum over ranges of points * P_kept (signal-region points only)
        for (const Segment &sg : ck.segs) mse_evals += (double)(sg.end - sg.start) * (double)kept[sg.spec];
        for (ProfSpan &sp : ck.spans)
            if (sp.kernel == MDB_KERNEL_MSE_SUPERPOSITION) sp.work = mse_evals;
        prof_resolve(&ck.spans);
    }
    for (size_t s = 0; s < ck.count; ++s) {
        SpecResult &r = results[ck.first + s];
        if (r.status != MDB_OK) continue;
        const FitDesc &f = ck.fdesc[s];
        r.lor.assign(lor + f.off, lor + f.off + kept[s]);
        r.peaks.assign(pk + 3 * f.off, pk + 3 * (f.off + f.n_peaks));
        r.mse = with_mse ? mse[s] : 0.0;
    }
    ck.finished = true;
    return MDB_OK;
}

static size_t chunk_size_for(const std::vector<HostSpec> &hs)
{
    size_t max_n = 0;
    for (auto &h : hs) max_n = std::max(max_n, h.n);
    const char *env = std::getenv("MDB_CHUNK_SPECTRA");
    if (env && std::atoi(env) > 0) return (size_t)std::atoi(env);
    // About 2^23 points per chunk (64 spectra of 2^17 points): small enough that eight chunks in
    // flight keep every SM busy through each other's kernel tails and host round trips, large
    // enough that a fit-refinement launch still fills the GPU (measured: profiles/depth_sweep_r1_*.txt).
    const size_t per_spec = 48 * max_n + 4096;          // device bytes per spectrum
    const size_t budget = (size_t)1536 << 20;           // per workspace
    size_t c = std::min(budget / per_spec, ((size_t)1 << 23) / std::max<size_t>(max_n, 1));
    c = std::max<size_t>(16, std::min<size_t>(c, 512));
    c = std::max<size_t>(1, std::min(c, budget / per_spec));
    return c;
}

// Validate views, fetch x[0], x[1] and run the host-side precompute.
static mdb_status build_host_specs(const mdb_deconvoluter &dc, const mdb_spectrum_view *sp, size_t n_spec,
                                   int memory, std::vector<HostSpec> &hs)
{
    // the kernels take the smoothing settings as int: larger values are valid in the reference (usize) but
    // would run for years there too -- refuse them instead of letting the casts wrap
    if (dc.smoothing.kind == MDB_SMOOTHING_MOVING_AVERAGE
        && (dc.smoothing.iterations > (uint64_t)INT_MAX || dc.smoothing.window_size > (uint64_t)INT_MAX))
        return fail(MDB_ERR_UNSUPPORTED, "moving average: iterations and window size above 2^31 - 1 are not supported");
    hs.resize(n_spec);
    std::map<const double *, std::pair<double, double>> x01;
    for (size_t s = 0; s < n_spec; ++s) {
        const mdb_spectrum_view &v = sp[s];
        if (!v.chemical_shifts || !v.intensities)
            return fail(MDB_ERR_INVALID_ARGUMENT, "spectrum " + std::to_string(s) + ": null array");
        if (v.len < 5 || v.len >= ((size_t)1 << 31))
            return fail(MDB_ERR_INVALID_ARGUMENT, "spectrum " + std::to_string(s) + ": length must be in [5, 2^31)");
        HostSpec &h = hs[s];
        h.x = v.chemical_shifts;
        h.y = v.intensities;
        h.n = v.len;
        h.sb[0] = v.signal_boundaries[0];
        h.sb[1] = v.signal_boundaries[1];
        h.threshold = dc.selection.threshold;
        h.fit_iters = (int)std::min<uint64_t>(dc.fitting.iterations, (uint64_t)INT_MAX);
        auto it = x01.find(h.x);
        if (it == x01.end()) {
            double two[2];
            if (memory == MDB_MEM_HOST) { two[0] = h.x[0]; two[1] = h.x[1]; }
            else CUDA_TRY(counted_memcpy(two, h.x, 16, cudaMemcpyDeviceToHost));
            it = x01.emplace(h.x, std::make_pair(two[0], two[1])).first;
        }
        h.x0 = it->second.first;
        h.x1 = it->second.second;
        if (dc.smoothing.kind == MDB_SMOOTHING_MOVING_AVERAGE && h.n < dc.smoothing.window_size / 2)
            h.pre_status = MDB_ERR_REFERENCE_PANIC;  // `values_len - self.right` underflows (moving_average.rs:62)
        precompute_spec(h, dc);
    }
    return MDB_OK;
}

// The streams of one pipeline call.  CUDA multiplexes streams over a small number of hardware queues
// (CUDA_DEVICE_MAX_CONNECTIONS, 8 unless the process raised it before creating its context); streams
// that share a queue serialise behind each other, which was measured as stage B of chunk k+2 starting
// only when chunk k had finished with one stream pair per workspace (22 streams).  So a call uses a
// small fixed set, handed to the chunks round robin: ONE copy stream (input copies in chunk order),
// N_A high-priority stage-A streams, and `depth` low-priority stage-B streams.
struct PipeStreams {
    static constexpr int MAX_A = 4, MAX_B = 32;
    int device = -1;
    cudaStream_t copy = nullptr, a[MAX_A] = {}, b[MAX_B] = {}, m[MAX_B] = {};
    cudaEvent_t ev_in[64] = {};   // "inputs landed", one per workspace slot of the ring
    cudaEvent_t ev_fit[64] = {};  // "refinement done", likewise
};
static std::mutex g_pipe_mutex;
static std::multimap<int, PipeStreams *> g_pipe_pool;

static mdb_status acquire_pipe_streams(PipeStreams **out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        auto it = g_pipe_pool.find(dev);
        if (it != g_pipe_pool.end()) {
            *out = it->second;
            g_pipe_pool.erase(it);
            return MDB_OK;
        }
    }
    auto *ps = new PipeStreams();
    ps->device = dev;
    int least = 0, greatest = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    if (const char *env = std::getenv("MDB_STREAM_PRIORITIES")) if (env[0] == '0') greatest = least;  // measurement aid
    CUDA_TRY(cudaStreamCreateWithPriority(&ps->copy, cudaStreamNonBlocking, greatest));
    for (auto &st : ps->a) CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest));
    // three levels: stage A and the copies highest, the refinement passes in the middle, the MSE pass lowest
    // -- only with MDB_FIT_PRIORITY=1; by default refinement and MSE share the lowest level and one stream per chunk
    const char *fp = std::getenv("MDB_FIT_PRIORITY");
    const bool split = fp && fp[0] == '1';  // measured: no gain (profiles/sweep_r2.txt), so off unless asked for
    const int medium = split ? (least + greatest) / 2 : least;
    for (auto &st : ps->b) CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, medium));
    if (split)
        for (auto &st : ps->m) CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, least));
    for (auto &e : ps->ev_in) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ps->ev_fit) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = ps;
    return MDB_OK;
}

static void release_pipe_stream_pool(int dev)  // destroys the idle stream sets of one device
{
    std::vector<PipeStreams *> mine;
    {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        auto range = g_pipe_pool.equal_range(dev);
        for (auto it = range.first; it != range.second; ++it) mine.push_back(it->second);
        g_pipe_pool.erase(range.first, range.second);
    }
    for (PipeStreams *ps : mine) {
        if (ps->copy) cudaStreamDestroy(ps->copy);
        for (auto st : ps->a) if (st) cudaStreamDestroy(st);
        for (auto st : ps->b) if (st) cudaStreamDestroy(st);
        for (auto st : ps->m) if (st) cudaStreamDestroy(st);
        for (auto e : ps->ev_in) if (e) cudaEventDestroy(e);
        for (auto e : ps->ev_fit) if (e) cudaEventDestroy(e);
        delete ps;
    }
}

static void release_pipe_streams(PipeStreams *ps)
{
    if (!ps) return;
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    g_pipe_pool.emplace(ps->device, ps);
}

// Submit the gather of a chunk's pageable rows into its workspace's page-locked staging area.
static mdb_status prepare_staging(Chunk &ck, const std::vector<HostSpec> &hs, Stager &stager)
{
    Workspace &ws = *ck.ws;
    auto job = std::make_shared<StageJob>();
    const size_t S = ck.count;
    job->src.resize(S); job->len.resize(S); job->off.resize(S); job->part_of_row.resize(S);
    size_t elems = 0, part_elems = 0;
    const size_t PART_ELEMS = ((size_t)32 << 20) / 8;
    for (size_t s = 0; s < S; ++s) {
        const HostSpec &h = hs[ck.first + s];
        job->src[s] = h.y; job->len[s] = h.n; job->off[s] = elems;  // the layout of ws.y in stage_a
        elems += align_up(h.n, 16);
        part_elems += align_up(h.n, 16);
        job->part_of_row[s] = (int)job->part_end.size();
        if (part_elems >= PART_ELEMS || s + 1 == S) { job->part_end.push_back(s + 1); part_elems = 0; }
    }
    job->part_left.reset(new std::atomic<int>[job->part_end.size()]);
    for (size_t pt = 0, row0 = 0; pt < job->part_end.size(); ++pt) {
        job->part_left[pt].store((int)(job->part_end[pt] - row0));
        row0 = job->part_end[pt];
    }
    CUDA_TRY(ws.h_stage.ensure(elems * 8));
    job->dst = ws.h_stage.as<double>();
    ck.stage_job = job;
    stager.submit(job);
    return MDB_OK;
}

// The chunked pipeline over a prepared list of spectra.  Up to eight workspaces (two streams each)
// in flight: stage A of chunk k+1 is queued before the host waits for the counts of chunk k, and
// older chunks are unpacked while the younger ones run their fit / MSE kernels.  With pageable
// host rows the chunks are CREATED one further ahead, so that the staging threads gather chunk
// k+2 while chunk k+1 is in stage A.
static mdb_status run_pipeline(const mdb_deconvoluter &dc, const std::vector<HostSpec> &hs, int memory,
                               std::vector<SpecResult> &results)
{
    const size_t n_spectra = hs.size();
    mdb_status st = MDB_OK;
    t_pipeline_origin = std::chrono::steady_clock::now();
    t_timeline_base = nullptr;
    t_trace.clear();
    if (std::getenv("MDB_TIMELINE") && cudaEventCreate(&t_timeline_base) == cudaSuccess) cudaEventRecord(t_timeline_base, nullptr);
    size_t depth = 6;  // chunks in stage B at a time = stage-B streams (sweeps: profiles/sweep_r2.txt)
    if (const char *env = std::getenv("MDB_PIPELINE_DEPTH"))
        if (std::atoi(env) >= 1) depth = (size_t)std::min(std::atoi(env), 32);
    // Pageable rows?  (probed on the first and the last spectrum; a chunk that mixes kinds is still
    // handled, row kinds are re-probed per chunk)
    const bool serial = depth == 1;  // strictly serial (per-kernel profiling): A(k) B(k) finish(k) A(k+1)
    bool any_pageable = false;
    if (memory == MDB_MEM_HOST) any_pageable = !host_rows_pinned(hs, 0, 1) || !host_rows_pinned(hs, n_spectra - 1, 1);
    // Stage A runs `ahead` chunks in front of stage B: the first chunks are small (their peak counts
    // size the later ones), and queueing several of them at once lets their H2D copies and smoothing
    // latencies overlap instead of costing one round trip each while the GPU has nothing else to do.
    const size_t ahead = serial ? 0 : std::min<size_t>(3, depth - 1);
    const size_t look = (any_pageable && !serial) ? 1 : 0;  // chunks created (and gathered) ahead of their stage A
    const size_t ring = depth + ahead + look;               // workspaces: `depth` chunks in stage B, the rest in front of it
    std::unique_ptr<Stager> stager;
    if (any_pageable) {
        size_t bytes = 0;
        for (auto &h : hs) bytes += h.n * 8;
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        // pipelines sharing this host's cores: the in-process sharder's threads, or the other ranks of a torchrun job
        // (LOCAL_WORLD_SIZE) -- eight processes with eight gather threads each oversubscribed a 32-thread host
        int peers = std::max(1, t_pipeline_peers);
        if (const char *lws = std::getenv("LOCAL_WORLD_SIZE")) peers = std::max(peers, std::atoi(lws));
        size_t n_thr = std::max<size_t>(2, std::min<size_t>(8, hw / 2 / (unsigned)peers));
        if (const char *env = std::getenv("MDB_STAGE_THREADS")) if (std::atoi(env) >= 1) n_thr = (size_t)std::min(std::atoi(env), 64);
        // small calls (one blood spectrum is 1 MB): stage A gathers the rows on the calling thread -- starting and
        // joining a helper thread costs more than the copy
        if (bytes >= ((size_t)4 << 20)) stager.reset(new Stager(n_thr));
    }
    // Chunk size: starts at chunk_size_for() and, unless pinned by MDB_CHUNK_SPECTRA, is re-derived
    // from the first chunk's selected-peak counts so that a chunk carries about TARGET Lorentzian
    // evaluations (~6 ms of FP64 work): many-peak spectra get small chunks (fine-grained overlap
    // of kernel tails across streams), few-peak spectra get large ones (launch and smoothing
    // latency amortised).  Measured in profiles/depth_sweep_r1_*.txt and profiles/sweep_r2.txt.
    double TARGET_EVALS = 0.9e10;
    if (const char *env = std::getenv("MDB_TARGET_EVALS")) if (std::atof(env) > 0.0) TARGET_EVALS = std::atof(env);  // sweeps
    const bool pinned_size = std::getenv("MDB_CHUNK_SPECTRA") && std::atoi(std::getenv("MDB_CHUNK_SPECTRA")) > 0;
    size_t csz = chunk_size_for(hs);
    std::vector<Workspace *> wss;
    std::vector<Chunk> chunks;
    chunks.reserve(n_spectra / 16 + 2);
    PipeStreams *pipe = nullptr;
    if ((st = acquire_pipe_streams(&pipe)) != MDB_OK) return st;
    size_t n_a = 2;  // stage-A streams: two chunks' smoothing / detection / selection side by side
    if (const char *env = std::getenv("MDB_STAGE_A_STREAMS")) if (std::atoi(env) >= 1) n_a = (size_t)std::min(std::atoi(env), (int)PipeStreams::MAX_A);
    auto cleanup = [&]() {
        if (pipe) {
            cudaStreamSynchronize(pipe->copy);
            for (size_t i = 0; i < n_a; ++i) cudaStreamSynchronize(pipe->a[i]);
            for (size_t i = 0; i < depth; ++i) cudaStreamSynchronize(pipe->b[i]);
            for (size_t i = 0; i < depth; ++i) if (pipe->m[i]) cudaStreamSynchronize(pipe->m[i]);
            release_pipe_streams(pipe);
            pipe = nullptr;
        }
        stager.reset();  // joins the staging threads before their buffers go back to the pool
        // in reverse: the pool hands out the most recently released workspace first, so the next call of
        // the same shape gives every chunk slot the workspace (and the buffer sizes) it had this time
        for (size_t i = wss.size(); i-- > 0;)
            if (wss[i]) { cudaStreamSynchronize(wss[i]->stream); release_workspace(wss[i]); }
    };
    size_t next_first = 0;
    // Appends the next chunk.  Its workspace is that of chunk k - ring, which is finished first.
    auto make_chunk = [&]() -> mdb_status {
        const size_t k = chunks.size();
        if (k >= ring && !chunks[k - ring].finished) {
            mdb_status s2 = finish_chunk(chunks[k - ring], results, true);
            if (s2 != MDB_OK) return s2;
        }
        if (k < ring) {
            Workspace *w = nullptr;
            mdb_status s2 = acquire_workspace(&w);
            if (s2 != MDB_OK) return s2;
            wss.push_back(w);
        }
        Chunk ck;
        ck.ws = wss[k % ring];
        ck.s_copy = pipe->copy;
        ck.s_a = pipe->a[k % n_a];
        ck.s_b = pipe->b[k % depth];
        ck.s_m = pipe->m[k % depth];  // null when MDB_FIT_PRIORITY=0
        ck.ev_in = pipe->ev_in[k % ring];
        ck.ev_fit = pipe->ev_fit[k % ring];
        ck.first = next_first;
        ck.startup = !serial && k <= ahead;  // the start-up burst (see stage_a: latency form of the smoothing kernel)
        size_t count = std::min(csz, n_spectra - next_first);
        if (n_spectra - next_first - count < csz / 4) count = n_spectra - next_first;  // no tiny straggler chunk
        ck.count = count;
        next_first += count;
        chunks.push_back(std::move(ck));
        if (stager && !host_rows_pinned(hs, chunks.back().first, count)) return prepare_staging(chunks.back(), hs, *stager);
        return MDB_OK;
    };
    auto ensure_created = [&](size_t upto) -> mdb_status {  // chunks 0..upto exist (as far as the batch reaches)
        while (chunks.size() <= upto && next_first < n_spectra) {
            mdb_status s2 = make_chunk();
            if (s2 != MDB_OK) return s2;
        }
        return MDB_OK;
    };
    auto retune = [&](const Chunk &ck) {
        if (pinned_size || ck.count == 0 || ck.est_evals <= 0.0) return;
        const double per_spec = ck.est_evals / (double)ck.count;
        const size_t want = (size_t)std::max(1.0, TARGET_EVALS / per_spec);
        size_t max_n = 0;
        for (auto &h : hs) max_n = std::max(max_n, h.n);
        const size_t mem_cap = std::max<size_t>(1, ((size_t)1536 << 20) / (48 * max_n + 4096));
        csz = std::max<size_t>(std::min<size_t>(48, mem_cap), std::min({want, (size_t)512, mem_cap}));
    };
    size_t a_next = 0;  // the next chunk whose stage A has not been queued
    size_t b_next = 0;  // the next chunk whose stage B has not been launched
    // Stage A queued for chunks 0..upto (as far as the batch reaches).  A chunk whose pageable rows are
    // still being gathered is left for the next round as long as stage B has other chunks to launch:
    // the host thread should be waiting for the GPU's counts, not for a memcpy three chunks ahead.
    auto queue_stage_a = [&](size_t upto, bool may_defer) -> mdb_status {
        mdb_status s2 = ensure_created(upto + look);
        while (s2 == MDB_OK && a_next <= upto && a_next < chunks.size()) {
            Chunk &c = chunks[a_next];
            if (may_defer && c.stage_job && a_next > b_next && !c.stage_job->all_done()) break;
            s2 = stage_a(c, hs, dc, memory, false);
            ++a_next;
        }
        return s2;
    };
    st = queue_stage_a(serial ? 0 : ahead, false);  // the start-up burst: nothing else to do yet
    for (size_t k = 0; st == MDB_OK && k < chunks.size(); ++k) {
        if (serial) {
            st = stage_b(chunks[k], hs, dc, results, true, nullptr);
            if (st == MDB_OK) st = finish_chunk(chunks[k], results, true);
            b_next = k + 1;
            if (st == MDB_OK && k == 0) retune(chunks[0]);
            if (st == MDB_OK) st = queue_stage_a(k + 1, false);
            continue;
        }
        // stage B of chunk k first (its counts are usually on the host already), THEN the stage A of the chunk
        // `ahead` in front: queueing that one may block -- on the staging threads (pageable rows) or on the
        // chunk whose workspace it takes over -- and the FP64 kernels of chunk k should not wait for either
        st = stage_b(chunks[k], hs, dc, results, true, nullptr);
        b_next = k + 1;
        if (st == MDB_OK && k == 0) retune(chunks[0]);
        if (st == MDB_OK) st = queue_stage_a(k + 1 + ahead, true);
    }
    for (size_t k = 0; st == MDB_OK && k < chunks.size(); ++k)
        if (chunks[k].stage_b_launched && !chunks[k].finished) st = finish_chunk(chunks[k], results, true);
    if (const char *path = std::getenv("MDB_TIMELINE")) {  // measurement aid: one line per chunk, appended
        static std::mutex tl_mutex;
        std::lock_guard<std::mutex> lock(tl_mutex);
        if (FILE *f = std::fopen(path, "a")) {
            int dev = 0;
            cudaGetDevice(&dev);
            std::fprintf(f, "{\"device\": %d, \"spectra\": %zu, \"total_ms\": %.3f, \"chunks\": [", dev, n_spectra, pipeline_ms());
            for (size_t k = 0; k < chunks.size(); ++k) {
                float g[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
                for (int q = 0; q < 5; ++q)
                    if (t_timeline_base && chunks[k].g_ev[q] && cudaEventElapsedTime(&g[q], t_timeline_base, chunks[k].g_ev[q]) != cudaSuccess) {
                        cudaGetLastError();
                        g[q] = -1.f;
                    }
                std::fprintf(f, "%s[%zu, %zu, %.3f, %.3f, %.3f, %.3f, %.3f, %.3f, %.3f, %.3f, %.3f]", k ? ", " : "", chunks[k].first,
                             chunks[k].count, chunks[k].t_a, chunks[k].t_counts, chunks[k].t_b, chunks[k].t_done, g[0], g[1], g[2], g[3], g[4]);
            }
            std::fprintf(f, "], \"kernels\": [");
            for (size_t q = 0; q < t_trace.size(); ++q)
                std::fprintf(f, "%s[%d, %.3f, %.3f]", q ? ", " : "", t_trace[q].kernel, t_trace[q].t0, t_trace[q].t1);
            std::fprintf(f, "]}\n");
            std::fclose(f);
        }
    }
    t_trace.clear();
    for (Chunk &ck : chunks)
        for (cudaEvent_t &e : ck.g_ev)
            if (e) { cudaEventDestroy(e); e = nullptr; }
    if (t_timeline_base) { cudaEventDestroy(t_timeline_base); t_timeline_base = nullptr; }
    cleanup();
    return st;
}

// ---------------------------------------------------------------------------------------------
// Small spectra (every spectrum of the call <= SMALL_MAX_N points): small_fused.cuh.
// Per chunk: one packed H2D copy (descriptors, index lists, x rows, y rows), the smoothing launch,
// the fused launch, one stream synchronisation; the kernel stores its results straight into a
// page-locked host slot per spectrum.  MDB_SMALL_PATH=0 sends such calls through the general
// pipeline instead (tests compare the two).
// ---------------------------------------------------------------------------------------------
static bool small_path_applies(const std::vector<HostSpec> &hs)
{
    const char *env = std::getenv("MDB_SMALL_PATH");
    if (env && env[0] == '0') return false;
    if (hs.empty()) return false;
    for (const HostSpec &h : hs) {
        if (h.n > (size_t)SMALL_MAX_N || h.ys_dev) return false;
        long long pts = 0;  // residual slots: the ranges of a valid spectrum are disjoint, but do not rely on it
        for (auto &rg : h.ranges) pts += rg.second - rg.first;
        if (pts > (long long)h.n) return false;
    }
    return true;
}

static mdb_status run_small(const mdb_deconvoluter &dc, const std::vector<HostSpec> &hs, int memory,
                            std::vector<SpecResult> &results)
{
    const size_t n_spectra = hs.size();
    size_t max_n = 0;
    for (auto &h : hs) max_n = std::max(max_n, h.n);
    const int n_al = (int)align_up(max_n, 8);
    const int cap = n_al / 2;
    const size_t slot = small_slot_bytes(cap);
    const size_t smem = small_smem_bytes(n_al);
    if ((int)smem + 1024 > smem_optin_limit()) return fail(MDB_ERR_UNSUPPORTED, "small path: shared memory limit");
    const size_t row = (size_t)n_al;  // elements per x / y / ys row
    size_t chunk_cap = std::min(((size_t)64 << 20) / slot, ((size_t)64 << 20) / (16 * row + 256));
    chunk_cap = std::max<size_t>(1, chunk_cap);
    const bool ma = dc.smoothing.kind == MDB_SMOOTHING_MOVING_AVERAGE;
    // the moving average runs inside the fused kernel (one warp, one lane per pass) unless it has
    // more passes than a warp has lanes; MDB_SMALL_SMOOTH=separate keeps it a launch of its own
    const char *sm_env = std::getenv("MDB_SMALL_SMOOTH");
    const bool smooth_fused = ma && dc.smoothing.iterations <= (uint64_t)SMALL_SMOOTH_MAX_ITERS
                              && dc.smoothing.window_size <= (uint64_t)SMALL_MAX_N && !(sm_env && sm_env[0] == 's');
    const bool need_tmp = ma && !smooth_fused && dc.smoothing.iterations >= 2;

    Workspace *wsp = nullptr;
    mdb_status st = acquire_workspace(&wsp);
    if (st != MDB_OK) return st;
    Workspace &ws = *wsp;
    struct Releaser { Workspace *w; ~Releaser() { cudaStreamSynchronize(w->stream); release_workspace(w); } } releaser{wsp};
    {   // opt in to the largest shared-memory footprint once per device
        static std::mutex attr_mutex;
        static bool attr_done[64] = {};
        std::lock_guard<std::mutex> lock(attr_mutex);
        const int dv = ws.device >= 0 && ws.device < 64 ? ws.device : 0;
        if (!attr_done[dv]) {
            CUDA_TRY(cudaFuncSetAttribute(small_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)small_smem_bytes(SMALL_MAX_N)));
            attr_done[dv] = true;
        }
    }

    std::vector<SpecDesc> descs;
    std::vector<ProfSpan> spans;
    for (size_t first = 0; first < n_spectra; first += chunk_cap) {
        const size_t S = std::min(chunk_cap, n_spectra - first);
        // ---- layout of the packed input blob (same offsets on the host and on the device)
        size_t int_elems = 0;
        std::vector<size_t> ig_off(S), rg_off(S);
        std::map<std::pair<const double *, size_t>, size_t> x_map;  // (caller x pointer, length) -> row index
        for (size_t s = 0; s < S; ++s) {
            const HostSpec &h = hs[first + s];
            ig_off[s] = int_elems; int_elems += h.ig.size();
            rg_off[s] = int_elems; int_elems += 2 * h.ranges.size();
            if (memory == MDB_MEM_HOST && !x_map.count({h.x, h.n})) { const size_t k = x_map.size(); x_map[{h.x, h.n}] = k; }
        }
        const size_t off_small = align_up(S * sizeof(SpecDesc), 16);
        const size_t off_int = off_small + align_up(S * sizeof(SmallDesc), 16);
        const size_t off_x = off_int + align_up(int_elems * 4, 16);
        const size_t off_y = off_x + x_map.size() * row * 8;
        const size_t in_bytes = (memory == MDB_MEM_HOST) ? off_y + S * row * 8 : off_x;
        CUDA_TRY(ws.small_in.ensure(in_bytes));
        CUDA_TRY(ws.h_small_in.ensure(in_bytes));
        CUDA_TRY(ws.h_small_out.ensure(S * slot));
        if (ma && !smooth_fused) CUDA_TRY(ws.ys.ensure(S * row * 8));
        if (need_tmp) CUDA_TRY(ws.tmp.ensure(S * row * 8));
        CUDA_TRY(ws.fit_state.ensure(S * 14 * (size_t)cap * 8));
        unsigned char *out_dev = nullptr;  // device-side address of the page-locked result slots
        CUDA_TRY(cudaHostGetDevicePointer((void **)&out_dev, ws.h_small_out.p, 0));

        unsigned char *hb = ws.h_small_in.as<unsigned char>();
        unsigned char *db = ws.small_in.as<unsigned char>();
        SpecDesc *h_desc = reinterpret_cast<SpecDesc *>(hb);
        SmallDesc *h_small = reinterpret_cast<SmallDesc *>(hb + off_small);
        int *h_int = reinterpret_cast<int *>(hb + off_int);
        const int *d_int = reinterpret_cast<const int *>(db + off_int);
        descs.assign(S, SpecDesc{});
        for (size_t s = 0; s < S; ++s) {
            const HostSpec &h = hs[first + s];
            SpecDesc &d = descs[s];
            if (memory == MDB_MEM_HOST) {
                d.x = reinterpret_cast<const double *>(db + off_x) + x_map[{h.x, h.n}] * row;
                d.y = reinterpret_cast<const double *>(db + off_y) + s * row;
                std::memcpy(hb + off_y + s * row * 8, h.y, h.n * 8);
            } else {
                d.x = h.x;
                d.y = h.y;
            }
            // Identity (smoothing/identity.rs) and the fused moving average read the raw row
            d.ys = (ma && !smooth_fused) ? ws.ys.as<double>() + s * row : const_cast<double *>(d.y);
            d.tmp = need_tmp ? ws.tmp.as<double>() + s * row : nullptr;
            d.pk = nullptr; d.sc = nullptr; d.tile_cnt = nullptr; d.sfr = nullptr; d.sel = nullptr;
            d.ig = d_int + ig_off[s];
            d.n = (int)h.n;
            d.n_tiles = 0;
            d.sb0 = h.sb_i0; d.sb1 = h.sb_i1;
            d.n_ig = (int)(h.ig.size() / 2);
            d.has_ig = dc.has_ignore ? 1 : 0;
            d.threshold = h.threshold;
            for (size_t q = 0; q < h.ig.size(); ++q) h_int[ig_off[s] + q] = h.ig[q];
            for (size_t q = 0; q < h.ranges.size(); ++q) {
                h_int[rg_off[s] + 2 * q] = h.ranges[q].first;
                h_int[rg_off[s] + 2 * q + 1] = h.ranges[q].second;
            }
            SmallDesc &e = h_small[s];
            e.ranges = d_int + rg_off[s];
            e.fit_state = ws.fit_state.as<double>() + s * 14 * (size_t)cap;
            e.out = out_dev + s * slot;
            e.n_ranges = (int)h.ranges.size();
            e.n_iters = h.fit_iters;
            e.cap = cap;
            e.skip = (h.pre_status != MDB_OK) ? 1 : 0;
        }
        std::memcpy(h_desc, descs.data(), S * sizeof(SpecDesc));
        if (memory == MDB_MEM_HOST)
            for (auto &kv : x_map) std::memcpy(hb + off_x + kv.second * row * 8, kv.first.first, kv.first.second * 8);
        // (letting the kernel pull its rows out of the mapped staging blob instead was measured: 32 KB
        // of zero-copy reads cost 13 us inside the kernel, the DMA plus its launch about 6)
        CUDA_TRY(counted_memcpy_async(db, hb, in_bytes, cudaMemcpyHostToDevice, ws.stream));

        const SpecDesc *d_desc = reinterpret_cast<const SpecDesc *>(db);
        if (ma && !smooth_fused) {
            st = launch_smooth(ws.stream, d_desc, descs, (int)dc.smoothing.iterations, (int)dc.smoothing.window_size, &spans);
            if (st != MDB_OK) return st;
        }
        long long *d_stamps = nullptr;  // MDB_SMALL_STAMPS=1: phase clock of CTA 0, printed to stderr (development aid)
        const char *stamps_env = std::getenv("MDB_SMALL_STAMPS");
        if (stamps_env && stamps_env[0] == '1') {
            CUDA_TRY(ws.blk_off.ensure(16 * 8));
            CUDA_TRY(cudaMemsetAsync(ws.blk_off.p, 0, 16 * 8, ws.stream));
            d_stamps = ws.blk_off.as<long long>();
        }
        prof_begin(&spans, MDB_KERNEL_SMALL_FUSED, ws.stream);
        small_fused_kernel<<<(unsigned)S, SMALL_THREADS, smem, ws.stream>>>(
            d_desc, reinterpret_cast<const SmallDesc *>(db + off_small), n_al, dc.selection.kind,
            smooth_fused ? (int)dc.smoothing.iterations : 0, (int)dc.smoothing.window_size,
            (dc.sup_mode == MDB_SUPERPOSITION_FAST ? 1 : 0) | (stream_generic_env() ? 2 : 0), d_stamps);
        LAUNCH_CHECK();
        prof_end(&spans, ws.stream, (double)S);
        CUDA_TRY(cudaStreamSynchronize(ws.stream));
        prof_resolve(&spans);
        if (d_stamps) {
            long long h[16];
            CUDA_TRY(cudaMemcpy(h, d_stamps, sizeof(h), cudaMemcpyDeviceToHost));
            static const char *names[] = {"smooth", "centres", "borders", "select", "fit_init", "fit_iters", "retain+mse", "fold"};
            std::fprintf(stderr, "[mdb small stamps, cycles]");
            for (int k = 1; k <= 8 && h[k] > 0; ++k) std::fprintf(stderr, " %s=%lld", names[k - 1], h[k] - h[k - 1]);
            std::fprintf(stderr, "\n");
        }

        // ---- unpack the result slots
        size_t out_bytes = 0;
        for (size_t s = 0; s < S; ++s) {
            const HostSpec &h = hs[first + s];
            SpecResult &r = results[first + s];
            if (h.pre_status != MDB_OK) { r.status = h.pre_status; continue; }
            const unsigned char *sl = ws.h_small_out.as<unsigned char>() + s * slot;
            const SmallOut *o = reinterpret_cast<const SmallOut *>(sl);
            r.info = o->info;
            r.status = o->info.status;
            out_bytes += sizeof(SmallOut);
            if (r.status != MDB_OK) continue;
            const mdb_lorentzian *lor = reinterpret_cast<const mdb_lorentzian *>(sl + 64);
            const int32_t *pk = reinterpret_cast<const int32_t *>(sl + 64 + (size_t)24 * cap);
            r.lor.assign(lor, lor + o->n_kept);
            r.peaks.assign(pk, pk + 3 * (size_t)o->info.n_selected);
            r.mse = o->mse;
            out_bytes += (size_t)24 * o->n_kept + (size_t)12 * o->info.n_selected;
        }
        count_transfer(out_bytes, cudaMemcpyDeviceToHost);  // written by the kernel over PCIe (zero-copy stores)
    }
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// In-process multi-GPU sharding (host-memory batches only)
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_device_policy{1};  // 1 = the calling thread's current device; 0 = all visible; n = first n

extern "C" mdb_status mdb_set_device_count(int n)
{
    if (n < 0) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_set_device_count: n must be >= 0 (0 = all visible devices)");
    g_device_policy.store(n);
    return MDB_OK;
}

static int devices_to_use(size_t n_spectra)
{
    int want = g_device_policy.load();
    if (const char *env = std::getenv("MDB_DEVICES")) want = std::strcmp(env, "all") == 0 ? 0 : std::atoi(env);
    if (want == 1) return 1;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); return 1; }
    int n = (want <= 0) ? visible : std::min(want, visible);
    while (n > 1 && n_spectra < (size_t)2 * n) --n;  // at least two spectra per device
    return std::max(n, 1);
}

extern "C" size_t mdb_batch_len(const mdb_batch *b) { return b ? b->r.size() : 0; }
extern "C" mdb_status mdb_batch_status(const mdb_batch *b, size_t i)
{
    return (b && i < b->r.size()) ? (mdb_status)b->r[i].status : MDB_ERR_INVALID_ARGUMENT;
}
extern "C" size_t mdb_batch_n_lorentzians(const mdb_batch *b, size_t i) { return (b && i < b->r.size()) ? b->r[i].lor.size() : 0; }
extern "C" const mdb_lorentzian *mdb_batch_lorentzians(const mdb_batch *b, size_t i)
{
    return (b && i < b->r.size()) ? b->r[i].lor.data() : nullptr;
}
extern "C" double mdb_batch_mse(const mdb_batch *b, size_t i) { return (b && i < b->r.size()) ? b->r[i].mse : NAN; }
extern "C" size_t mdb_batch_n_peaks(const mdb_batch *b, size_t i) { return (b && i < b->r.size()) ? b->r[i].peaks.size() / 3 : 0; }
extern "C" const int32_t *mdb_batch_peaks(const mdb_batch *b, size_t i)
{
    return (b && i < b->r.size()) ? b->r[i].peaks.data() : nullptr;
}
extern "C" void mdb_batch_free(mdb_batch *b) { delete b; }

// Bulk accessors: one call instead of five per spectrum (language bindings with per-call overhead).
extern "C" void mdb_batch_totals(const mdb_batch *b, size_t *n_lorentzians, size_t *n_peaks)
{
    size_t nl = 0, np = 0;
    if (b)
        for (const SpecResult &r : b->r) { nl += r.lor.size(); np += r.peaks.size() / 3; }
    if (n_lorentzians) *n_lorentzians = nl;
    if (n_peaks) *n_peaks = np;
}

extern "C" mdb_status mdb_batch_export(const mdb_batch *b, int32_t *status, uint64_t *n_lorentzians, uint64_t *n_peaks,
                                       double *mse, mdb_lorentzian *lorentzians, int32_t *peaks)
{
    if (!b) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_batch_export: null batch");
    size_t ol = 0, op = 0;
    for (size_t i = 0; i < b->r.size(); ++i) {
        const SpecResult &r = b->r[i];
        if (status) status[i] = r.status;
        if (n_lorentzians) n_lorentzians[i] = r.lor.size();
        if (n_peaks) n_peaks[i] = r.peaks.size() / 3;
        if (mse) mse[i] = r.mse;
        if (lorentzians && !r.lor.empty()) std::memcpy(lorentzians + ol, r.lor.data(), r.lor.size() * sizeof(mdb_lorentzian));
        if (peaks && !r.peaks.empty()) std::memcpy(peaks + 3 * op, r.peaks.data(), r.peaks.size() * sizeof(int32_t));
        ol += r.lor.size();
        op += r.peaks.size() / 3;
    }
    return MDB_OK;
}

static const char *status_text(int st)
{
    switch (st) {
    case MDB_ERR_NO_PEAKS_DETECTED: return "no peaks detected in the spectrum";
    case MDB_ERR_EMPTY_SIGNAL_REGION: return "no peaks found in the signal region of the spectrum";
    case MDB_ERR_EMPTY_SIGNAL_FREE_REGION: return "no peaks found in the signal free region of the spectrum";
    case MDB_ERR_REFERENCE_PANIC: return "input on which the reference implementation panics";
    default: return "error";
    }
}

extern "C" mdb_status mdb_deconvolute_spectra(const mdb_deconvoluter *d, const mdb_spectrum_view *spectra,
                                              size_t n_spectra, int memory, mdb_batch **out)
{
    if (out) *out = nullptr;
    if (!d || !out || (n_spectra && !spectra)) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvolute_spectra: null argument");
    if (memory != MDB_MEM_HOST && memory != MDB_MEM_DEVICE) return fail(MDB_ERR_INVALID_ARGUMENT, "unknown memory kind");
    mdb_status st = require_device();
    if (st != MDB_OK) return st;
    std::unique_ptr<mdb_batch> batch(new mdb_batch());
    batch->r.resize(n_spectra);
    if (n_spectra == 0) { *out = batch.release(); return MDB_OK; }
    // the call works on its own copy of the settings: the arithmetic of an unpinned deconvoluter is the
    // process default AS OF NOW, and stays that for the whole call whatever other threads set meanwhile
    mdb_deconvoluter dc_local = *d;
    if (dc_local.sup_mode < 0 && (st = default_superposition_mode(&dc_local.sup_mode)) != MDB_OK) return st;
    d = &dc_local;
    std::vector<HostSpec> hs;
    if ((st = build_host_specs(*d, spectra, n_spectra, memory, hs)) != MDB_OK) return st;

    const int n_dev = (memory == MDB_MEM_HOST) ? devices_to_use(n_spectra) : 1;
    const bool small = small_path_applies(hs);
    auto run = small ? run_small : run_pipeline;
    if (n_dev <= 1) {
        st = run(*d, hs, memory, batch->r);
    } else {
        // One host thread per GPU, contiguous shards, no exchange: the in-process form of the
        // one-process-per-GPU sharding.  Device d of the shard list is CUDA device d.
        std::vector<mdb_status> sts(n_dev, MDB_OK);
        std::vector<std::string> msgs(n_dev);
        std::vector<std::thread> threads;
        for (int dev = 0; dev < n_dev; ++dev) {
            threads.emplace_back([&, dev]() {
                const size_t lo = (size_t)dev * n_spectra / n_dev, hi = (size_t)(dev + 1) * n_spectra / n_dev;
                if (cudaSetDevice(dev) != cudaSuccess) {
                    sts[dev] = MDB_ERR_CUDA;
                    msgs[dev] = "cudaSetDevice(" + std::to_string(dev) + ") failed";
                    return;
                }
                t_pipeline_peers = n_dev;
                std::vector<HostSpec> shard(hs.begin() + lo, hs.begin() + hi);
                std::vector<SpecResult> res(hi - lo);
                sts[dev] = run(*d, shard, memory, res);
                if (sts[dev] != MDB_OK) msgs[dev] = g_last_error;  // thread-local in the worker
                for (size_t i = lo; i < hi; ++i) batch->r[i] = std::move(res[i - lo]);
            });
        }
        for (auto &t : threads) t.join();
        for (int dev = 0; dev < n_dev && st == MDB_OK; ++dev)
            if (sts[dev] != MDB_OK) st = fail(sts[dev], "device " + std::to_string(dev) + ": " + msgs[dev]);
    }
    if (st != MDB_OK) return st;
    mdb_status first = MDB_OK;
    for (size_t s = 0; s < n_spectra; ++s)
        if (batch->r[s].status != MDB_OK) {
            first = (mdb_status)batch->r[s].status;
            fail(first, "spectrum " + std::to_string(s) + ": " + status_text(first));
            break;
        }
    *out = batch.release();
    return first;
}

// ---------------------------------------------------------------------------------------------
// Deconvoluter::optimize_settings  deconvoluter.rs:761-825
// 27 smoothing settings (iterations 2..=10 x window 3,5,7) x 10 thresholds (5 + c*3/9) x 3 fit
// iteration counts (5,10,15) = 810 deconvolutions of one spectrum, argmin MSE (first minimum in
// iteration order, Iterator::min_by).  On the GPU: the output of pass k of a 10-pass smoothing is
// the k-iteration result, so three all-passes runs (one per window) give all 27 smoothed curves;
// the 810 (smoothing, threshold, iterations) variants then go through the ordinary chunk pipeline
// as "spectra" sharing x, y and one of the 27 smoothed rows.
// ---------------------------------------------------------------------------------------------
extern "C" mdb_status mdb_deconvoluter_optimize_settings(mdb_deconvoluter *d, const mdb_spectrum_view *spectrum,
                                                         int memory, double *mse_out)
{
    if (!d || !spectrum || !spectrum->chemical_shifts || !spectrum->intensities)
        return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_deconvoluter_optimize_settings: null argument");
    if (memory != MDB_MEM_HOST && memory != MDB_MEM_DEVICE) return fail(MDB_ERR_INVALID_ARGUMENT, "unknown memory kind");
    const size_t n = spectrum->len;
    if (n < 5 || n >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "length must be in [5, 2^31)");
    mdb_status st = require_device();
    if (st != MDB_OK) return st;

    const int windows[3] = {3, 5, 7};
    const int max_iters = 10;
    const size_t stride = align_up(n, 16);
    DevBuf xb, yb, ysb, jobb;
    struct Free { DevBuf *b[4]; ~Free() { for (DevBuf *p : b) p->release(); } } guard{{&xb, &yb, &ysb, &jobb}};
    CUDA_TRY(ysb.ensure(3 * (size_t)max_iters * stride * 8));
    CUDA_TRY(jobb.ensure(3 * sizeof(SmoothAllJob)));
    const double *dx = spectrum->chemical_shifts, *dy = spectrum->intensities;
    Workspace *prep = nullptr;  // its stream orders the preparatory copies and the all-passes smoothing
    if ((st = acquire_workspace(&prep)) != MDB_OK) return st;
    struct PrepGuard { Workspace *w; ~PrepGuard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } prep_guard{prep};
    cudaStream_t stream = prep->stream;
    if (memory == MDB_MEM_HOST) {
        CUDA_TRY(xb.ensure(stride * 8));
        CUDA_TRY(yb.ensure(stride * 8));
        CUDA_TRY(counted_memcpy_async(xb.p, spectrum->chemical_shifts, n * 8, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(counted_memcpy_async(yb.p, spectrum->intensities, n * 8, cudaMemcpyHostToDevice, stream));
        dx = xb.as<double>();
        dy = yb.as<double>();
    }
    SmoothAllJob jobs[3];
    for (int w = 0; w < 3; ++w) {
        jobs[w].y = dy;
        jobs[w].out = ysb.as<double>() + (size_t)w * max_iters * stride;
        jobs[w].stride = (long long)stride;
        jobs[w].n = (int)n;
        jobs[w].iters = max_iters;
        jobs[w].window = windows[w];
    }
    CUDA_TRY(counted_memcpy_async(jobb.p, jobs, sizeof(jobs), cudaMemcpyHostToDevice, stream));  // `jobs` outlives the sync below
    smooth_all_passes_kernel<<<3, 32, 0, stream>>>(jobb.as<SmoothAllJob>());
    LAUNCH_CHECK();
    CUDA_TRY(cudaStreamSynchronize(stream));

    // ---- the 810 variants in the reference's iteration order: smoothing (iterations outer, window
    // inner), then threshold, then fit iterations
    struct Variant { int iterations, window; double threshold; int fit; };
    std::vector<Variant> variants;
    for (int iterations = 2; iterations <= 10; ++iterations)
        for (int w = 0; w < 3; ++w)
            for (int c = 0; c < 10; ++c)
                for (int fit = 5; fit <= 15; fit += 5)
                    variants.push_back({iterations, w, 5.0 + (double)c * (8.0 - 5.0) / 9.0, fit});
    mdb_deconvoluter work = *d;
    work.sup_mode = MDB_SUPERPOSITION_EXACT;  // the argmin compares MSEs: always the reference's bit patterns
    work.fit_arith = MDB_FIT_EXACT;
    work.selection.kind = MDB_SELECTION_NOISE_SCORE_FILTER;
    work.selection.scoring_method = MDB_SCORING_MINIMUM_SUM;
    // the reference calls set_smoothing_settings(variant) before every run (deconvoluter.rs:787-803): the
    // caller's own window must not decide the `n < window / 2` panic check; the variants' windows are
    // 3, 5 and 7, which no admissible spectrum (n >= 5) trips
    work.smoothing = {MDB_SMOOTHING_MOVING_AVERAGE, 2, 7};
    mdb_spectrum_view view = *spectrum;
    view.chemical_shifts = dx;
    view.intensities = dy;
    std::vector<mdb_spectrum_view> views(variants.size(), view);
    std::vector<HostSpec> hs;
    if ((st = build_host_specs(work, views.data(), views.size(), MDB_MEM_DEVICE, hs)) != MDB_OK) return st;
    for (size_t v = 0; v < variants.size(); ++v) {
        hs[v].ys_dev = ysb.as<double>() + ((size_t)variants[v].window * max_iters + (size_t)(variants[v].iterations - 1)) * stride;
        hs[v].threshold = variants[v].threshold;
        hs[v].fit_iters = variants[v].fit;
    }
    std::vector<SpecResult> results(variants.size());
    if ((st = run_pipeline(work, hs, MDB_MEM_DEVICE, results)) != MDB_OK) return st;
    size_t best = 0;
    for (size_t v = 0; v < variants.size(); ++v) {
        if (results[v].status != MDB_OK)  // any failing combination aborts the optimisation (:805-812)
            return fail((mdb_status)results[v].status, std::string("optimize_settings: ") + status_text(results[v].status));
        if (results[v].mse != results[v].mse)  // partial_cmp(..).unwrap() panics on NaN (:813)
            return fail(MDB_ERR_REFERENCE_PANIC, "optimize_settings: NaN mean squared error");
        if (results[v].mse < results[best].mse) best = v;  // strict: the first minimum wins, as Iterator::min_by
    }
    d->smoothing = {MDB_SMOOTHING_MOVING_AVERAGE, (uint64_t)variants[best].iterations, (uint64_t)windows[variants[best].window]};
    d->selection = {MDB_SELECTION_NOISE_SCORE_FILTER, MDB_SCORING_MINIMUM_SUM, variants[best].threshold};
    d->fitting = {MDB_FITTING_ANALYTICAL, (uint64_t)variants[best].fit};
    if (mse_out) *mse_out = results[best].mse;
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// superposition_vec  (lorentzian.rs:631-635, 656-663)
// ---------------------------------------------------------------------------------------------
using SupKernel = void (*)(const double *, long long, const double *, int, double *, const SpecDesc *, const FitDesc *,
                           const Segment *, const double *, const int *);

// Points per thread from the size of the whole job on this device (every shape gives the same bits).
static int superposition_points_per_thread(size_t n, bool ulp)
{
    const size_t blocks8 = (n + (size_t)SUP_THREADS * 8 - 1) / ((size_t)SUP_THREADS * 8);
    return (ulp && blocks8 >= (size_t)8 * sm_count() && !std::getenv("MDB_SUP_R8")) ? 16 : blocks8 >= (size_t)2 * sm_count() ? 8 : 2;
}
static SupKernel superposition_kernel_for(int r, bool ulp)
{
    return ulp ? (sup_quad() ? (r == 16 ? superposition_kernel<0, 16, 4> : r == 8 ? superposition_kernel<0, 8, 4> : superposition_kernel<0, 2, 4>)
                             : (r == 16 ? superposition_kernel<0, 16, 2> : r == 8 ? superposition_kernel<0, 8, 2> : superposition_kernel<0, 2, 2>))
               : (r == 8 ? superposition_kernel<0, 8, 1> : superposition_kernel<0, 2, 1>);
}

// One slice of a HOST-memory grid on the calling thread's current device: the parameter table goes
// up once, the slice streams through in chunks over three streams (H2D of chunk c+1 and D2H of chunk
// c-1 overlap the kernel of chunk c; the kernels of neighbouring chunks overlap each other's tails).
static mdb_status superposition_host_slice(const double *x, size_t n, const mdb_lorentzian *lor, size_t p, double *out, bool ulp)
{
    constexpr int NS = 3;
    Workspace *ws[NS] = {nullptr, nullptr, nullptr};
    struct Guard {
        Workspace **w;
        ~Guard() { for (int i = 0; i < NS; ++i) if (w[i]) { cudaStreamSynchronize(w[i]->stream); release_workspace(w[i]); } }
    } guard{ws};
    mdb_status st;
    for (int i = 0; i < NS; ++i)
        if ((st = acquire_workspace(&ws[i])) != MDB_OK) return st;
    const int r = superposition_points_per_thread(n, ulp);
    const size_t per_block = (size_t)SUP_THREADS * r;
    // chunk: about an eighth of the slice, between 64 CTAs' worth and 2^21 points, a multiple of a CTA's points
    size_t chunk = std::max<size_t>(64 * per_block, std::min<size_t>((size_t)1 << 21, (n + 7) / 8));
    if (const char *env = std::getenv("MDB_SUP_CHUNK")) if (std::atoll(env) > 0) chunk = (size_t)std::atoll(env);
    chunk = (chunk + per_block - 1) / per_block * per_block;
    if ((chunk + per_block - 1) / per_block > 0x7fffffffull) return fail(MDB_ERR_INVALID_ARGUMENT, "grid too large");
    CUDA_TRY(ws[0]->lor.ensure(std::max<size_t>(p, 1) * 24));
    if (p) CUDA_TRY(counted_memcpy_async(ws[0]->lor.p, lor, p * 24, cudaMemcpyHostToDevice, ws[0]->stream));
    CUDA_TRY(cudaEventRecord(ws[0]->ev_a, ws[0]->stream));
    SupKernel kern = superposition_kernel_for(r, ulp);
    std::vector<ProfSpan> spans;
    size_t c = 0;
    for (size_t lo = 0; lo < n; lo += chunk, ++c) {
        Workspace &w = *ws[c % NS];
        const size_t cnt = std::min(chunk, n - lo);
        if (c < NS) {
            CUDA_TRY(w.x.ensure(std::min(chunk, n) * 8));
            CUDA_TRY(w.ys.ensure(std::min(chunk, n) * 8));
            if (c > 0) CUDA_TRY(cudaStreamWaitEvent(w.stream, ws[0]->ev_a, 0));  // the parameter table has landed
        }
        CUDA_TRY(counted_memcpy_async(w.x.p, x + lo, cnt * 8, cudaMemcpyHostToDevice, w.stream));
        prof_begin(&spans, MDB_KERNEL_SUPERPOSITION_VEC, w.stream);
        kern<<<(unsigned)((cnt + per_block - 1) / per_block), SUP_THREADS, LOR_SMEM_BYTES, w.stream>>>(
            w.x.as<double>(), (long long)cnt, ws[0]->lor.as<double>(), (int)p, w.ys.as<double>(), nullptr, nullptr, nullptr, nullptr, nullptr);
        LAUNCH_CHECK();
        prof_end(&spans, w.stream, (double)cnt * (double)p);
        CUDA_TRY(counted_memcpy_async(out + lo, w.ys.p, cnt * 8, cudaMemcpyDeviceToHost, w.stream));
    }
    for (int i = 0; i < NS; ++i) CUDA_TRY(cudaStreamSynchronize(ws[i]->stream));
    prof_resolve(&spans);
    return MDB_OK;
}

extern "C" mdb_status mdb_superposition_vec_mode(const double *x, size_t n, const mdb_lorentzian *lor, size_t p,
                                                 double *out, int memory, int mode)
{
    if ((n && (!x || !out)) || (p && !lor)) return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_superposition_vec: null argument");
    if (p >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "too many lorentzians");
    if (memory != MDB_MEM_HOST && memory != MDB_MEM_DEVICE) return fail(MDB_ERR_INVALID_ARGUMENT, "unknown memory kind");
    if (mode != MDB_SUPERPOSITION_EXACT && mode != MDB_SUPERPOSITION_FAST)
        return fail(MDB_ERR_INVALID_ARGUMENT, "mdb_superposition_vec_mode: unknown mode");
    mdb_status st = require_device();
    if (st != MDB_OK) return st;
    if (n == 0) return MDB_OK;
    const bool ulp = mode == MDB_SUPERPOSITION_FAST;
    if (memory == MDB_MEM_HOST) {
        // par_superposition_vec's rayon-over-points (lorentzian.rs:660-662) as grid slices over the GPUs
        // mdb_set_device_count allows: at least 2^16 points per device, parameters replicated, no exchange
        const int n_dev = devices_to_use(n >> 15);
        if (n_dev <= 1) return superposition_host_slice(x, n, lor, p, out, ulp);
        std::vector<mdb_status> sts(n_dev, MDB_OK);
        std::vector<std::string> msgs(n_dev);
        std::vector<std::thread> threads;
        for (int dev = 0; dev < n_dev; ++dev)
            threads.emplace_back([&, dev]() {
                const size_t lo = (size_t)dev * n / n_dev, hi = (size_t)(dev + 1) * n / n_dev;
                if (cudaSetDevice(dev) != cudaSuccess) {
                    sts[dev] = MDB_ERR_CUDA;
                    msgs[dev] = "cudaSetDevice(" + std::to_string(dev) + ") failed";
                    return;
                }
                sts[dev] = superposition_host_slice(x + lo, hi - lo, lor, p, out + lo, ulp);
                if (sts[dev] != MDB_OK) msgs[dev] = g_last_error;  // thread-local in the worker
            });
        for (auto &t : threads) t.join();
        for (int dev = 0; dev < n_dev; ++dev)
            if (sts[dev] != MDB_OK) return fail(sts[dev], "device " + std::to_string(dev) + ": " + msgs[dev]);
        return MDB_OK;
    }
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    const double *dl = (const double *)lor;
    if (((uintptr_t)lor & 15) != 0) {  // the kernel's bulk copies need 16-byte alignment
        CUDA_TRY(ws->lor.ensure(std::max<size_t>(p, 1) * 24));
        if (p) CUDA_TRY(counted_memcpy_async(ws->lor.p, lor, p * 24, cudaMemcpyDeviceToDevice, ws->stream));
        dl = ws->lor.as<double>();
    }
    const int r = superposition_points_per_thread(n, ulp);
    const size_t per_block = (size_t)SUP_THREADS * r;
    const size_t blocks = (n + per_block - 1) / per_block;
    if (blocks > 0x7fffffffull) return fail(MDB_ERR_INVALID_ARGUMENT, "grid too large");
    std::vector<ProfSpan> spans;
    prof_begin(&spans, MDB_KERNEL_SUPERPOSITION_VEC, ws->stream);
    superposition_kernel_for(r, ulp)<<<(unsigned)blocks, SUP_THREADS, LOR_SMEM_BYTES, ws->stream>>>(
        x, (long long)n, dl, (int)p, out, nullptr, nullptr, nullptr, nullptr, nullptr);
    LAUNCH_CHECK();
    prof_end(&spans, ws->stream, (double)n * (double)p);
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    prof_resolve(&spans);
    return MDB_OK;
}

extern "C" mdb_status mdb_superposition_vec(const double *x, size_t n, const mdb_lorentzian *lor, size_t p,
                                            double *out, int memory)
{
    int mode = 0;
    mdb_status st = default_superposition_mode(&mode);
    if (st != MDB_OK) return st;
    return mdb_superposition_vec_mode(x, n, lor, p, out, memory, mode);
}

// FP64 instruction rate of the current device (bench.py's measured denominator): every SM fully
// occupied by 16 independent chains per thread of nothing but DFMA, then DADD; CUDA events on the
// launching stream; best of three.
extern "C" mdb_status mdb_measure_fp64_rate(double *dfma_per_s, double *dadd_per_s)
{
    mdb_status st = require_device();
    if (st != MDB_OK) return st;
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    const int blocks = sm_count() * 8, iters = 1 << 14;
    CUDA_TRY(ws->tmp.ensure((size_t)blocks * RATE_THREADS * 8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best[2] = {0.0, 0.0};
    for (int kind = 0; kind < 2; ++kind)
        for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
            CUDA_TRY(cudaEventRecord(e0, ws->stream));
            if (kind == 0) fp64_rate_kernel<0><<<blocks, RATE_THREADS, 0, ws->stream>>>(ws->tmp.as<double>(), iters, 1.0 + rep);
            else fp64_rate_kernel<1><<<blocks, RATE_THREADS, 0, ws->stream>>>(ws->tmp.as<double>(), iters, 1.0 + rep);
            LAUNCH_CHECK();
            CUDA_TRY(cudaEventRecord(e1, ws->stream));
            CUDA_TRY(cudaStreamSynchronize(ws->stream));
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            const double rate = (double)blocks * RATE_THREADS * RATE_CHAINS * (double)iters / ((double)ms * 1e-3);
            if (rep > 0) best[kind] = std::max(best[kind], rate);
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (dfma_per_s) *dfma_per_s = best[0];
    if (dadd_per_s) *dadd_per_s = best[1];
    return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// Stage entry points (parity suite)
// ---------------------------------------------------------------------------------------------
extern "C" mdb_status mdb_stage_smooth(const double *values, size_t n, uint64_t iterations, uint64_t window, double *out)
{
    if (!values || !out) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 5 || n >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "length must be in [5, 2^31)");
    mdb_smoothing_settings sm = {MDB_SMOOTHING_MOVING_AVERAGE, iterations, window};
    mdb_status st = validate_smoothing(sm);
    if (st != MDB_OK) return st;
    if (n < window / 2) return fail(MDB_ERR_REFERENCE_PANIC, "fewer points than half the window");
    if (iterations > (uint64_t)INT_MAX || window > (uint64_t)INT_MAX)
        return fail(MDB_ERR_UNSUPPORTED, "moving average: iterations and window size above 2^31 - 1 are not supported");
    if ((st = require_device()) != MDB_OK) return st;
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    const size_t bytes = align_up(n, 16) * 8;
    CUDA_TRY(ws->y.ensure(bytes));
    CUDA_TRY(ws->ys.ensure(bytes));
    CUDA_TRY(ws->tmp.ensure(bytes));
    CUDA_TRY(ws->desc.ensure(sizeof(SpecDesc)));
    SpecDesc d{};
    d.y = ws->y.as<double>();
    d.ys = ws->ys.as<double>();
    d.tmp = ws->tmp.as<double>();
    d.n = (int)n;
    CUDA_TRY(counted_memcpy_async(ws->y.p, values, n * 8, cudaMemcpyHostToDevice, ws->stream));
    CUDA_TRY(counted_memcpy_async(ws->desc.p, &d, sizeof(d), cudaMemcpyHostToDevice, ws->stream));
    {
        std::vector<SpecDesc> descs(1, d);
        if ((st = launch_smooth(ws->stream, ws->desc.as<SpecDesc>(), descs, (int)iterations, (int)window, nullptr)) != MDB_OK)
            return st;
    }
    CUDA_TRY(counted_memcpy_async(out, ws->ys.p, n * 8, cudaMemcpyDeviceToHost, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    return MDB_OK;
}

// Test / measurement entry: K1 on `count` equally long spectra ALREADY IN DEVICE MEMORY (rows of
// `stride` doubles, 16-byte aligned), one launch.  Lets bench.py show where the exact-recurrence
// smoothing stops being latency bound (thousands of spectra per launch).  *ms (optional) receives
// the kernel time measured with CUDA events on the launching stream.
extern "C" mdb_status mdb_stage_smooth_batch(const double *values_dev, size_t n, size_t count, size_t stride,
                                             uint64_t iterations, uint64_t window, double *out_dev, double *ms)
{
    if (!values_dev || !out_dev || count == 0) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 5 || n >= ((size_t)1 << 31) || stride < n) return fail(MDB_ERR_INVALID_ARGUMENT, "bad length / stride");
    mdb_smoothing_settings sm = {MDB_SMOOTHING_MOVING_AVERAGE, iterations, window};
    mdb_status st = validate_smoothing(sm);
    if (st != MDB_OK) return st;
    if (iterations > (uint64_t)INT_MAX || window > (uint64_t)INT_MAX)
        return fail(MDB_ERR_UNSUPPORTED, "moving average: iterations and window size above 2^31 - 1 are not supported");
    if ((st = require_device()) != MDB_OK) return st;
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    std::vector<SpecDesc> descs(count);
    for (size_t s = 0; s < count; ++s) {
        descs[s] = SpecDesc{};
        descs[s].y = values_dev + s * stride;
        descs[s].ys = out_dev + s * stride;
        descs[s].n = (int)n;
    }
    CUDA_TRY(ws->desc.ensure(count * sizeof(SpecDesc)));
    if (iterations >= 2) {  // the generic fallback ping-pongs through tmp
        CUDA_TRY(ws->tmp.ensure(count * stride * 8));
        for (size_t s = 0; s < count; ++s) descs[s].tmp = ws->tmp.as<double>() + s * stride;
    }
    CUDA_TRY(counted_memcpy_async(ws->desc.p, descs.data(), count * sizeof(SpecDesc), cudaMemcpyHostToDevice, ws->stream));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, ws->stream));
    st = launch_smooth(ws->stream, ws->desc.as<SpecDesc>(), descs, (int)iterations, (int)window, nullptr);
    if (st == MDB_OK) {
        CUDA_TRY(cudaEventRecord(e1, ws->stream));
        CUDA_TRY(cudaStreamSynchronize(ws->stream));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
        if (ms) *ms = (double)t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return st;
}

// Runs stage A on one already-smoothed spectrum held on the host.
static mdb_status stage_a_single(const mdb_deconvoluter &dc, const double *smoothed, size_t n, size_t sb0, size_t sb1,
                                 int has_ig, const size_t *ig, size_t n_ig, Chunk &ck, std::vector<HostSpec> &hs)
{
    hs.assign(1, HostSpec());
    HostSpec &h = hs[0];
    h.x = smoothed;  // placeholder axis: detection and selection never read x
    h.y = smoothed;
    h.n = n;
    h.sb_i0 = clamp_idx(sb0);
    h.sb_i1 = clamp_idx(sb1);
    h.threshold = dc.selection.threshold;
    h.fit_iters = (int)std::min<uint64_t>(dc.fitting.iterations, (uint64_t)INT_MAX);
    if (has_ig)
        for (size_t q = 0; q < 2 * n_ig; ++q) h.ig.push_back(clamp_idx(ig[q]));
    mdb_deconvoluter local = dc;
    local.has_ignore = has_ig != 0;
    ck.first = 0;
    ck.count = 1;
    return stage_a(ck, hs, local, MDB_MEM_HOST, true);
}

extern "C" mdb_status mdb_stage_detect(const double *smoothed, size_t n, int32_t *peaks, double *scores, size_t cap,
                                       size_t *n_found)
{
    if (!smoothed || !n_found || (cap && (!peaks || !scores))) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 5 || n >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "length must be in [5, 2^31)");
    mdb_status st = require_device();
    if (st != MDB_OK) return st;
    mdb_deconvoluter *dc = nullptr;
    if ((st = mdb_deconvoluter_default(&dc)) != MDB_OK) return st;
    std::unique_ptr<mdb_deconvoluter> dc_guard(dc);
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    Chunk ck;
    ck.ws = ws;
    std::vector<HostSpec> hs;
    if ((st = stage_a_single(*dc, smoothed, n, 0, n, 0, nullptr, 0, ck, hs)) != MDB_OK) return st;
    // download the per-tile candidate lists and concatenate them in tile order (test-only path)
    const SpecDesc &d = ck.desc[0];
    std::vector<int> tile_cnt(d.n_tiles);
    CUDA_TRY(cudaStreamSynchronize(ws->stream_a));
    CUDA_TRY(counted_memcpy(tile_cnt.data(), d.tile_cnt, (size_t)d.n_tiles * 4, cudaMemcpyDeviceToHost));
    std::vector<int> pk(3 * (size_t)DETECT_CAP);
    std::vector<double> sc(DETECT_CAP);
    size_t k = 0;
    for (int t = 0; t < d.n_tiles; ++t) {
        const int c = tile_cnt[t];
        if (c == 0) continue;
        CUDA_TRY(counted_memcpy(pk.data(), d.pk + 3 * (size_t)t * DETECT_CAP, (size_t)c * 12, cudaMemcpyDeviceToHost));
        CUDA_TRY(counted_memcpy(sc.data(), d.sc + (size_t)t * DETECT_CAP, (size_t)c * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c; ++i, ++k) {
            if (k < cap) {
                peaks[3 * k] = pk[3 * i]; peaks[3 * k + 1] = pk[3 * i + 1]; peaks[3 * k + 2] = pk[3 * i + 2];
                scores[k] = sc[i];
            }
        }
    }
    *n_found = k;
    return MDB_OK;
}

extern "C" mdb_status mdb_stage_select(const mdb_deconvoluter *dc, const double *smoothed, size_t n, size_t sb0,
                                       size_t sb1, int has_ignore, const size_t *ignore_idx, size_t n_ignore,
                                       int32_t *peaks, size_t cap, size_t *n_selected, double *mean_sd)
{
    if (!dc || !smoothed || !n_selected || (cap && !peaks) || (n_ignore && !ignore_idx))
        return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 5 || n >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "length must be in [5, 2^31)");
    mdb_status st = require_device();
    if (st != MDB_OK) return st;
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    Chunk ck;
    ck.ws = ws;
    std::vector<HostSpec> hs;
    if ((st = stage_a_single(*dc, smoothed, n, sb0, sb1, has_ignore, ignore_idx, n_ignore, ck, hs)) != MDB_OK) return st;
    CUDA_TRY(cudaEventSynchronize(ws->ev_a));
    const SelectOut so = ws->h_sel_out.as<SelectOut>()[0];
    *n_selected = (size_t)so.n_selected;
    if (mean_sd) { mean_sd[0] = so.mean; mean_sd[1] = so.sd; }
    if (so.status != MDB_OK) return fail((mdb_status)so.status, status_text(so.status));
    const size_t ncopy = std::min<size_t>(cap, (size_t)so.n_selected);
    if (ncopy) CUDA_TRY(counted_memcpy(peaks, ck.desc[0].sel, ncopy * 12, cudaMemcpyDeviceToHost));
    return MDB_OK;
}

extern "C" mdb_status mdb_stage_fit(const double *x, const double *y, size_t n, const int32_t *peaks, size_t n_peaks,
                                    uint64_t iterations, mdb_lorentzian *out, size_t *n_retained, mdb_lorentzian *trace)
{
    if (!x || !y || !n_retained || (n_peaks && (!peaks || !out))) return fail(MDB_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 5 || n >= ((size_t)1 << 31)) return fail(MDB_ERR_INVALID_ARGUMENT, "length must be in [5, 2^31)");
    for (size_t q = 0; q < 3 * n_peaks; ++q)
        if (peaks[q] < 0 || (size_t)peaks[q] >= n) return fail(MDB_ERR_REFERENCE_PANIC, "peak index out of range");
    mdb_fitting_settings fs = {MDB_FITTING_ANALYTICAL, iterations};
    mdb_status st = validate_fitting(fs);
    if (st != MDB_OK) return st;
    if ((st = require_device()) != MDB_OK) return st;
    *n_retained = 0;
    if (n_peaks == 0) return MDB_OK;
    mdb_deconvoluter *dc = nullptr;
    if ((st = mdb_deconvoluter_default(&dc)) != MDB_OK) return st;
    std::unique_ptr<mdb_deconvoluter> dc_guard(dc);
    dc->fitting = fs;
    Workspace *ws = nullptr;
    if ((st = acquire_workspace(&ws)) != MDB_OK) return st;
    struct Guard { Workspace *w; ~Guard() { cudaStreamSynchronize(w->stream); release_workspace(w); } } guard{ws};
    // hand-built chunk: x, y and the peak list go straight into the workspace
    Chunk ck;
    ck.ws = ws;
    ck.first = 0;
    ck.count = 1;
    const size_t bytes = align_up(n, 16) * 8;
    CUDA_TRY(ws->x.ensure(bytes));
    CUDA_TRY(ws->y.ensure(bytes));
    CUDA_TRY(ws->sel.ensure(n_peaks * 12));
    CUDA_TRY(ws->desc.ensure(sizeof(SpecDesc)));
    CUDA_TRY(ws->sel_out.ensure(sizeof(SelectOut)));
    CUDA_TRY(ws->h_sel_out.ensure(sizeof(SelectOut)));
    SpecDesc d{};
    d.x = ws->x.as<double>();
    d.y = ws->y.as<double>();
    d.sel = ws->sel.as<int>();
    d.n = (int)n;
    ck.desc.assign(1, d);
    CUDA_TRY(counted_memcpy_async(ws->x.p, x, n * 8, cudaMemcpyHostToDevice, ws->stream));
    CUDA_TRY(counted_memcpy_async(ws->y.p, y, n * 8, cudaMemcpyHostToDevice, ws->stream));
    CUDA_TRY(counted_memcpy_async(ws->sel.p, peaks, n_peaks * 12, cudaMemcpyHostToDevice, ws->stream));
    CUDA_TRY(counted_memcpy_async(ws->desc.p, &d, sizeof(d), cudaMemcpyHostToDevice, ws->stream));
    SelectOut so{};
    so.status = MDB_OK;
    so.n_selected = (int)n_peaks;
    *ws->h_sel_out.as<SelectOut>() = so;
    CUDA_TRY(cudaEventRecord(ws->ev_a, ws->stream));
    std::vector<HostSpec> hs(1);
    hs[0].n = n;
    hs[0].fit_iters = (int)std::min<uint64_t>(iterations, (uint64_t)INT_MAX);
    std::vector<SpecResult> res(1);
    if ((st = stage_b(ck, hs, *dc, res, false, trace)) != MDB_OK) return st;
    if ((st = finish_chunk(ck, res, false)) != MDB_OK) return st;
    *n_retained = res[0].lor.size();
    std::memcpy(out, res[0].lor.data(), res[0].lor.size() * sizeof(mdb_lorentzian));
    return MDB_OK;
}
