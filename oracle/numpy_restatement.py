"""Second, independent CPU restatement of the reference's deconvolution path, in NumPy.

TEST INFRASTRUCTURE ONLY (same rules as mdb_oracle.c).  It was written from the reference's Rust
sources and SURVEY.md Appendix A, not from mdb_oracle.c, and exists so that the C oracle is not the
only opinion about what the reference computes: tests/test_oracle_kats.py requires the two to agree
bit for bit, stage by stage, on the bundled fixtures.  (The Rust reference itself cannot be built in
this image -- no cargo/rustc -- so end-to-end parity stays unpinned against a reference binary.)

Conventions that matter for bit-exactness:
  * element-wise NumPy arithmetic on float64 is IEEE round-to-nearest per operation, no FMA;
  * `np.add.accumulate` is a strict left fold (np.sum is pairwise and must not be used);
  * every formula keeps the reference's association order.
Citations are relative to /root/reference/metabodecon/src/.
"""
import math

import numpy as np

EPSILON = 2.220446049250313e-16
CHECK_PRECISION = 1.0e3 * EPSILON  # lib.rs:277


def seq_sum(values) -> float:
    """Iterator::sum over f64: left fold from zero."""
    values = np.asarray(values, dtype=np.float64)
    if values.size == 0:
        return 0.0
    return float(np.add.accumulate(np.concatenate(([0.0], values)))[-1])


def smooth_values(values, iterations: int, window: int) -> np.ndarray:
    """smoothing/moving_average.rs:53-83 with circular_buffer.rs:34-59 (VecDeque of capacity window)."""
    v = np.array(values, dtype=np.float64)
    n, right = v.size, window // 2
    for _ in range(iterations):
        fifo = []
        div, total = 1.0, 0.0
        for k in range(min(right, n)):
            fifo.append(v[k])
            total = total + v[k]
        for i in range(n - right):
            incoming = v[i + right]
            total = total + incoming
            if len(fifo) == window:
                popped = fifo.pop(0)
                fifo.append(incoming)
                total = total - popped
            else:
                fifo.append(incoming)
                div = 1.0 / float(len(fifo))
            v[i] = total * div
        for i in range(max(n - right, 0), n):
            if fifo:
                popped = fifo.pop(0)
                total = total - popped
                div = 1.0 / float(len(fifo)) if fifo else math.inf
                v[i] = total * div
    return v


def second_derivative(y) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    return (y[:-2] - 2.0 * y[1:-1]) + y[2:]  # peak_selection/common.rs:8


def detect_peaks(d2) -> np.ndarray:
    """detector.rs:99-164 -> (left, centre, right) rows, ascending by centre."""
    m = d2.size
    w0, w1, w2 = d2[:-2], d2[1:-1], d2[2:]
    centres = np.flatnonzero((w1 < 0.0) & (w1 < w0) & (w1 < w2)) + 2
    out = []
    for c in centres:
        right_slice = d2[c - 1:]
        rb = right_slice.size
        for p in range(right_slice.size - 2):
            a, b, cc = right_slice[p], right_slice[p + 1], right_slice[p + 2]
            if b > a and (b >= cc or (b < 0.0 and cc >= 0.0)):
                rb = p + 1
                break
        left_slice = d2[:c]
        lb = left_slice.size
        for p, q in enumerate(range(left_slice.size - 3, -1, -1)):
            a, b, cc = left_slice[q], left_slice[q + 1], left_slice[q + 2]
            if b > cc and (b >= a or (b < 0.0 and a >= 0.0)):
                lb = p + 1
                break
        left, right = c - lb, c + rb
        if left != 0 and right != m + 1:
            out.append((left, c, right))
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def score_peak(abs_d2, left, centre, right) -> float:  # scorer.rs:65-74
    return min(seq_sum(abs_d2[left - 1:centre]), seq_sum(abs_d2[centre - 1:right]))


def signal_boundaries_indices(x, sb):  # spectrum/spectrum.rs:741-746
    step = x[1] - x[0]
    return int(max(math.floor((sb[0] - x[0]) / step), 0)), int(max(math.ceil((sb[1] - x[0]) / step), 0))


def ignore_region_indices(x, sb, regions):  # deconvoluter.rs:865-904
    step, first = x[1] - x[0], x[0]
    lo_b, hi_b = min(sb), max(sb)
    i0, i1 = signal_boundaries_indices(x, sb)
    lower, upper = min(i0, i1), max(i0, i1)
    out = []
    for start, end in regions:
        if (start < lo_b and end < lo_b) or (start > hi_b and end > hi_b):
            continue
        fi = max(int(max(math.floor((start - first) / step), 0)), lower)
        si = min(int(max(math.ceil((end - first) / step), 0)), upper)
        a, b = min(fi, si), max(fi, si)
        if a < b - 1:
            out.append((a, b))
    return out


def select_peaks(smoothed, threshold, sb_idx, ignore_idx):
    """noise_score_filter.rs:32-54, 91-138 (NoiseScoreFilter / MinimumSum).  Returns (peaks, mean, sd)."""
    d2 = second_derivative(smoothed)
    peaks = detect_peaks(d2)
    if ignore_idx is not None:
        keep = [not any((s <= l < e) or (s <= r < e) for s, e in ignore_idx) for l, _, r in peaks]
        peaks = peaks[np.array(keep, dtype=bool)] if len(peaks) else peaks
    a = np.abs(d2)
    centres = peaks[:, 1]
    above0 = np.flatnonzero(centres > sb_idx[0])
    left = int(above0[0]) if above0.size else 0
    above1 = np.flatnonzero(centres[left:] > sb_idx[1])
    right = left + int(above1[0]) if above1.size else len(peaks) - 1
    sfr = [score_peak(a, *p) for p in peaks[:left]] + [score_peak(a, *p) for p in peaks[right:]]
    mean = seq_sum(sfr) / float(len(sfr))
    var = seq_sum([(s - mean) * (s - mean) for s in sfr]) / float(len(sfr))
    sd = math.sqrt(var)
    sel = [p for p in peaks[left:right] if score_peak(a, *p) >= mean + threshold * sd]
    return np.array(sel, dtype=np.int64).reshape(-1, 3), mean, sd


def superposition_vec(x, lor) -> np.ndarray:  # lorentzian.rs:546-548, 606-635
    x = np.asarray(x, dtype=np.float64)
    acc = np.zeros_like(x)
    with np.errstate(all="ignore"):
        for sfhw, hw2, maxp in lor:
            d = x - maxp
            acc = acc + sfhw / (hw2 + d * d)
    return acc


def _mirror(st):  # peak_stencil.rs:113-131; st columns x1 x2 x3 y1 y2 y3
    x1, x2, x3, y1, y2, y3 = (st[:, k].copy() for k in range(6))
    inc = (y1 <= y2) & (y2 <= y3)
    dec = ~inc & (y1 >= y2) & (y2 >= y3)
    y3n = np.where(inc, y1, y3)
    x3n = np.where(inc, 2.0 * x2 - x1, x3)
    y1n = np.where(dec, y3, y1)
    x1n = np.where(dec, 2.0 * x2 - x3, x1)
    return np.stack([x1n, x2, x3n, y1n, y2, y3n], axis=1)


def _solve(st):  # fitter_analytical.rs:147-172
    x1, x2, x3, y1, y2, y3 = (st[:, k] for k in range(6))
    with np.errstate(all="ignore"):
        num = ((x1 * x1) * y1) * (y2 - y3) + ((x2 * x2) * y2) * (y3 - y1) + ((x3 * x3) * y3) * (y1 - y2)
        den = ((2.0 * (x1 - x2)) * y1) * y2 + ((2.0 * (x2 - x3)) * y2) * y3 + ((2.0 * (x3 - x1)) * y3) * y1
        maxp = num / den
        d1, d2, d3 = (x1 - maxp) * (x1 - maxp), (x2 - maxp) * (x2 - maxp), (x3 - maxp) * (x3 - maxp)
        left = (y1 * d1 - y2 * d2) / (y2 - y1)
        right = (y2 * d2 - y3 * d3) / (y3 - y2)
        hw2 = (left + right) / 2.0
        hw2 = np.where(np.isnan(hw2) | (hw2 < EPSILON), EPSILON, hw2)  # f64::max(., EPSILON): NaN -> EPSILON
        sfhw = y2 * (hw2 + d2)
    return np.stack([sfhw, hw2, maxp], axis=1)


def fit_lorentzian(x, y, peaks, iterations):
    """fitter_analytical.rs:19-72 -> retained (sfhw, hw2, maxp) rows."""
    idx = np.asarray(peaks, dtype=np.int64).reshape(-1, 3)
    rx, ry = x[idx], y[idx]
    st = _mirror(np.concatenate([rx, ry], axis=1))
    lor = _solve(st)
    for _ in range(iterations):
        sup = superposition_vec(rx.reshape(-1), lor).reshape(-1, 3)
        with np.errstate(all="ignore"):
            st[:, 3:6] = st[:, 3:6] * (ry / sup)
        st = _mirror(st)
        lor = _solve(st)
    keep = (lor[:, 0] > CHECK_PRECISION) & (lor[:, 1] > CHECK_PRECISION)
    return lor[keep]


def compute_mse(sup, y, sb_idx, ignore_idx):  # deconvoluter.rs:828-862
    pts = [sb_idx[0]] + ([v for pair in ignore_idx for v in pair] if ignore_idx is not None else []) + [sb_idx[1]]
    ranges = list(zip(pts[0::2], pts[1::2]))
    residuals = seq_sum([seq_sum((sup[s:e] - y[s:e]) * (sup[s:e] - y[s:e])) for s, e in ranges])
    return residuals / float(sum(e - s for s, e in ranges))


def deconvolute(x, y, sb, smoothing=(3, 3), threshold=5.0, fit_iterations=10, ignore_regions=None):
    """Deconvoluter::deconvolute_spectrum (deconvoluter.rs:530-552), default algorithm choices."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    sm = smooth_values(y, *smoothing)
    sb_idx = signal_boundaries_indices(x, sb)
    ig = ignore_region_indices(x, sb, ignore_regions) if ignore_regions is not None else None
    peaks, mean, sd = select_peaks(sm, threshold, sb_idx, ig)
    lor = fit_lorentzian(x, y, peaks, fit_iterations)
    mse = compute_mse(superposition_vec(x, lor), y, sb_idx, ig)
    return {"smoothed": sm, "peaks": peaks, "mean": mean, "sd": sd, "lorentzians": lor, "mse": mse}
