"""ctypes wrapper around oracle/libmdb_oracle.so (the CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity status: unit KATs pinned (tests/test_oracle_kats.py); end-to-end parity unpinned because
the reference commits no golden deconvolution output (see oracle/mdb_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmdb_oracle.so")

OK = 0
NO_PEAKS_DETECTED = 1
EMPTY_SIGNAL_REGION = 2
EMPTY_SIGNAL_FREE_REGION = 3
PANIC = 100

SMOOTH_IDENTITY, SMOOTH_MOVING_AVERAGE = 0, 1
SELECT_DETECTOR_ONLY, SELECT_NOISE_SCORE_FILTER = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile if the .so is missing or stale."""
    src = os.path.join(_HERE, "mdb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


class _SelectInfo(C.Structure):
    _fields_ = [("n_detected", C.c_size_t), ("n_after_ignore", C.c_size_t),
                ("region_left", C.c_size_t), ("region_right", C.c_size_t),
                ("n_sfr", C.c_size_t), ("mean", C.c_double), ("sd", C.c_double)]


class _Settings(C.Structure):
    _fields_ = [("smoothing_kind", C.c_int), ("smoothing_iterations", C.c_size_t),
                ("smoothing_window", C.c_size_t), ("selection_kind", C.c_int),
                ("threshold", C.c_double), ("fitting_iterations", C.c_size_t),
                ("has_ignore_regions", C.c_int), ("n_ignore_regions", C.c_size_t),
                ("ignore_regions", C.POINTER(C.c_double))]


class _Result(C.Structure):
    _fields_ = [("status", C.c_int), ("n_selected", C.c_size_t), ("n_lorentzians", C.c_size_t),
                ("mse", C.c_double), ("info", _SelectInfo)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_score_peak.restype = C.c_double
        _lib.orc_superposition.restype = C.c_double
        _lib.orc_detect_peaks.restype = C.c_size_t
        _lib.orc_fit_lorentzian.restype = C.c_size_t
        _lib.orc_ignore_region_indices.restype = C.c_size_t
        _lib.orc_find_right_border.restype = C.c_size_t
        _lib.orc_find_left_border.restype = C.c_size_t
        _lib.orc_find_peak_centers.restype = C.c_size_t
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _sp(a):
    return a.ctypes.data_as(C.POINTER(C.c_size_t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _usz(a):
    return np.ascontiguousarray(a, dtype=np.uintp)


# ----------------------------------------------------------------------------- stage functions
def smooth_values(values, iterations: int, window_size: int) -> np.ndarray:
    v = _f64(values).copy()
    lib().orc_smooth_values(_dp(v), C.c_size_t(v.size), C.c_size_t(iterations), C.c_size_t(window_size))
    return v


def second_derivative(y) -> np.ndarray:
    y = _f64(y)
    d2 = np.empty(max(y.size - 2, 0), dtype=np.float64)
    lib().orc_second_derivative(_dp(y), C.c_size_t(y.size), _dp(d2))
    return d2


def detect_peaks(d2) -> np.ndarray:
    """Returns an (n, 3) uintp array of (left, center, right)."""
    d2 = _f64(d2)
    cap = d2.size // 2 + 2
    l, c, r = (np.zeros(cap, dtype=np.uintp) for _ in range(3))
    n = lib().orc_detect_peaks(_dp(d2), C.c_size_t(d2.size), _sp(l), _sp(c), _sp(r), C.c_size_t(cap))
    return np.stack([l[:n], c[:n], r[:n]], axis=1)


def find_right_border(d2_slice) -> int:
    t = _f64(d2_slice)
    return lib().orc_find_right_border(_dp(t), C.c_size_t(t.size))


def find_left_border(d2_slice) -> int:
    u = _f64(d2_slice)
    return lib().orc_find_left_border(_dp(u), C.c_size_t(u.size))


def find_peak_centers(d2) -> list:
    d2 = _f64(d2)
    out = np.zeros(d2.size + 1, dtype=np.uintp)
    n = lib().orc_find_peak_centers(_dp(d2), C.c_size_t(d2.size), _sp(out), C.c_size_t(out.size))
    return out[:n].tolist()


def find_peak_borders(d2, centers) -> list:
    d2 = _f64(d2)
    c = _usz(centers)
    out = np.zeros(2 * c.size, dtype=np.uintp)
    lib().orc_find_peak_borders(_dp(d2), C.c_size_t(d2.size), _sp(c), C.c_size_t(c.size), _sp(out))
    return [tuple(p) for p in out.reshape(-1, 2).tolist()]


def peak_region_boundaries(centers, sb_idx):
    c = _usz(centers)
    out = np.zeros(2, dtype=np.uintp)
    lib().orc_peak_region_boundaries(_sp(c), C.c_size_t(c.size), C.c_size_t(sb_idx[0]), C.c_size_t(sb_idx[1]), _sp(out))
    return int(out[0]), int(out[1])


def score_peak(abs_d2, left: int, center: int, right: int) -> float:
    a = _f64(abs_d2)
    return lib().orc_score_peak(_dp(a), C.c_size_t(left), C.c_size_t(center), C.c_size_t(right))


def mean_sd_scores(scores):
    s = _f64(scores)
    m, sd = C.c_double(), C.c_double()
    lib().orc_mean_sd_scores(_dp(s), C.c_size_t(s.size), C.byref(m), C.byref(sd))
    return m.value, sd.value


@dataclass
class Selection:
    status: int
    peaks: np.ndarray          # (P, 3) selected
    candidates: np.ndarray     # (C, 3) detected, after ignore filter (NoiseScoreFilter only)
    scores: np.ndarray         # (C,)
    n_detected: int
    n_after_ignore: int
    region: tuple
    n_sfr: int
    mean: float
    sd: float


def select_peaks(smoothed, selector: int, threshold: float, sb_idx, ignore_idx=None) -> Selection:
    sm = _f64(smoothed)
    n = sm.size
    cap = n // 2 + 1
    sel = [np.zeros(cap, dtype=np.uintp) for _ in range(3)]
    cand = [np.zeros(cap, dtype=np.uintp) for _ in range(3)]
    sc = np.zeros(cap, dtype=np.float64)
    nsel = C.c_size_t()
    info = _SelectInfo()
    has_ig = ignore_idx is not None
    ig = _usz(np.asarray(ignore_idx if has_ig else [], dtype=np.uintp).reshape(-1))
    st = lib().orc_select_peaks(_dp(sm), C.c_size_t(n), C.c_int(selector), C.c_double(threshold),
                                C.c_size_t(sb_idx[0]), C.c_size_t(sb_idx[1]), C.c_int(has_ig),
                                _sp(ig), C.c_size_t(ig.size // 2), C.c_size_t(cap),
                                _sp(sel[0]), _sp(sel[1]), _sp(sel[2]), C.byref(nsel),
                                _sp(cand[0]), _sp(cand[1]), _sp(cand[2]), _dp(sc), C.byref(info))
    p = nsel.value
    ncand = info.n_after_ignore if selector == SELECT_NOISE_SCORE_FILTER and info.n_sfr else 0
    return Selection(st, np.stack([a[:p] for a in sel], axis=1),
                     np.stack([a[:ncand] for a in cand], axis=1), sc[:ncand].copy(),
                     info.n_detected, info.n_after_ignore, (info.region_left, info.region_right),
                     info.n_sfr, info.mean, info.sd)


def mirror_shoulder(stencil):
    s = _f64(stencil).copy()
    lib().orc_mirror_shoulder(_dp(s))
    return s


def solve_stencil(stencil):
    """stencil = (x1,x2,x3,y1,y2,y3) -> (sfhw, hw2, maxp)."""
    s = _f64(stencil)
    out = np.zeros(3, dtype=np.float64)
    lib().orc_solve_stencil(_dp(s), _dp(out))
    return out


def superposition(x: float, lorentzians) -> float:
    l = _f64(lorentzians).reshape(-1, 3)
    return lib().orc_superposition(C.c_double(x), _dp(l), C.c_size_t(l.shape[0]))


def superposition_vec(x, lorentzians, parallel: bool = False) -> np.ndarray:
    x = _f64(x)
    l = _f64(lorentzians).reshape(-1, 3)
    out = np.empty_like(x)
    fn = lib().orc_par_superposition_vec if parallel else lib().orc_superposition_vec
    fn(_dp(x), C.c_size_t(x.size), _dp(l), C.c_size_t(l.shape[0]), _dp(out))
    return out


def fit_lorentzian(x, y, peaks, iterations: int, trace: bool = False, parallel: bool = False):
    """Returns (retained (K,3) array, trace or None).  trace: (iterations+1, P, 3)."""
    x, y = _f64(x), _f64(y)
    pk = np.asarray(peaks, dtype=np.uintp).reshape(-1, 3)
    p = pk.shape[0]
    l, c, r = (_usz(pk[:, i]) for i in range(3))
    lor = np.zeros((max(p, 1), 3), dtype=np.float64)
    tr = np.zeros((iterations + 1, max(p, 1), 3), dtype=np.float64) if trace else None
    kept = lib().orc_fit_lorentzian(_dp(x), _dp(y), _sp(l), _sp(c), _sp(r), C.c_size_t(p),
                                    C.c_size_t(iterations), _dp(lor),
                                    _dp(tr) if trace else None, C.c_int(parallel))
    return lor[:kept].copy(), (tr[:, :p] if trace else None)


def signal_boundaries_indices(x, sb):
    x = _f64(x)
    i0, i1 = C.c_size_t(), C.c_size_t()
    lib().orc_signal_boundaries_indices(_dp(x), C.c_double(sb[0]), C.c_double(sb[1]),
                                        C.byref(i0), C.byref(i1))
    return i0.value, i1.value


def ignore_region_indices(x, sb, regions):
    x = _f64(x)
    reg = _f64(regions).reshape(-1)
    out = np.zeros(reg.size + 2, dtype=np.uintp)
    k = lib().orc_ignore_region_indices(_dp(x), C.c_double(sb[0]), C.c_double(sb[1]), _dp(reg),
                                        C.c_size_t(reg.size // 2), _sp(out))
    return out[:2 * k].reshape(-1, 2)


def compute_mse(sup, y, sb_idx, ignore_idx=None):
    sup, y = _f64(sup), _f64(y)
    has_ig = ignore_idx is not None
    ig = _usz(np.asarray(ignore_idx if has_ig else [], dtype=np.uintp).reshape(-1))
    mse = C.c_double()
    st = lib().orc_compute_mse(_dp(sup), _dp(y), C.c_size_t(y.size), C.c_size_t(sb_idx[0]),
                               C.c_size_t(sb_idx[1]), C.c_int(has_ig), _sp(ig),
                               C.c_size_t(ig.size // 2), C.byref(mse))
    return st, mse.value


# ----------------------------------------------------------------------------- whole pipeline
@dataclass
class Settings:
    smoothing_kind: int = SMOOTH_MOVING_AVERAGE
    smoothing_iterations: int = 3
    smoothing_window: int = 3
    selection_kind: int = SELECT_NOISE_SCORE_FILTER
    threshold: float = 5.0
    fitting_iterations: int = 10
    ignore_regions: object = None  # None or list of (lo, hi) ppm pairs, merged + sorted

    def _c(self):
        has = self.ignore_regions is not None
        reg = _f64(np.asarray(self.ignore_regions if has else [], dtype=np.float64).reshape(-1))
        st = _Settings(self.smoothing_kind, self.smoothing_iterations, self.smoothing_window,
                       self.selection_kind, self.threshold, self.fitting_iterations, int(has),
                       reg.size // 2, _dp(reg))
        return st, reg  # keep reg alive


@dataclass
class Deconvolved:
    status: int
    lorentzians: np.ndarray   # (K, 3) sfhw, hw2, maxp
    peaks: np.ndarray         # (P, 3) selected
    mse: float
    smoothed: np.ndarray
    n_detected: int
    n_after_ignore: int
    region: tuple
    n_sfr: int
    mean: float
    sd: float


def deconvolute_spectrum(settings: Settings, x, y, sb, parallel: bool = False) -> Deconvolved:
    x, y = _f64(x), _f64(y)
    n = y.size
    cap = n // 2 + 1
    st, _keep = settings._c()
    lor = np.zeros((cap, 3), dtype=np.float64)
    sel = [np.zeros(cap, dtype=np.uintp) for _ in range(3)]
    sm = np.empty(n, dtype=np.float64)
    res = _Result()
    lib().orc_deconvolute_spectrum(C.byref(st), _dp(x), _dp(y), C.c_size_t(n), C.c_double(sb[0]),
                                   C.c_double(sb[1]), C.c_int(parallel), _dp(lor), _sp(sel[0]),
                                   _sp(sel[1]), _sp(sel[2]), _dp(sm), C.byref(res))
    p = res.n_selected
    i = res.info
    return Deconvolved(res.status, lor[:res.n_lorentzians].copy(),
                       np.stack([a[:p] for a in sel], axis=1), res.mse, sm, i.n_detected,
                       i.n_after_ignore, (i.region_left, i.region_right), i.n_sfr, i.mean, i.sd)


def optimize_settings(settings: Settings, x, y, sb):
    """Deconvoluter::optimize_settings (deconvoluter.rs:761-825).  Returns
    (status, (iterations, window, threshold, fit_iterations), best_mse, all 810 MSEs in iteration order)."""
    x, y = _f64(x), _f64(y)
    st, _keep = settings._c()
    best = np.zeros(4, dtype=np.float64)
    mse = C.c_double()
    all_mse = np.zeros(810, dtype=np.float64)
    lib().orc_optimize_settings.restype = C.c_int
    status = lib().orc_optimize_settings(C.byref(st), _dp(x), _dp(y), C.c_size_t(y.size), C.c_double(sb[0]),
                                         C.c_double(sb[1]), _dp(best), C.byref(mse), _dp(all_mse))
    return status, (int(best[0]), int(best[1]), float(best[2]), int(best[3])), mse.value, all_mse


def par_deconvolute_spectra(settings: Settings, x, ys, sb):
    """Batch form used as the CPU baseline: OpenMP over spectra (shared x, equal n).

    ys: (S, n) array.  Returns (status, list of (K,3) arrays, mse array, n_selected array).
    """
    x = _f64(x)
    ys = _f64(ys)
    s, n = ys.shape
    cap = n // 2 + 1
    st, _keep = settings._c()
    lor = np.zeros((s, cap, 3), dtype=np.float64)
    res = (_Result * s)()
    xp = (C.POINTER(C.c_double) * s)(*[_dp(x)] * s)
    yp = (C.POINTER(C.c_double) * s)(*[ys[i].ctypes.data_as(C.POINTER(C.c_double)) for i in range(s)])
    sbs = _f64(np.tile(np.asarray(sb, dtype=np.float64), (s, 1)))
    status = lib().orc_par_deconvolute_spectra(C.byref(st), C.c_size_t(s), xp, yp, C.c_size_t(n),
                                               _dp(sbs), _dp(lor), res)
    lors = [lor[i, :res[i].n_lorentzians].copy() for i in range(s)]
    mse = np.array([res[i].mse for i in range(s)])
    nsel = np.array([res[i].n_selected for i in range(s)])
    return status, lors, mse, nsel


def max_threads() -> int:
    return lib().orc_max_threads()


def use_cores(n: int) -> int:
    """Size the OpenMP pool to `n` threads (bench.py: one oracle per rank, cores shared out)."""
    lib().orc_set_num_threads(C.c_int(max(1, int(n))))
    return max_threads()


def use_all_cores() -> int:
    """Size the OpenMP pool to every core this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which would otherwise make the CPU baseline single-threaded)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(n))
    return max_threads()
