"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
 and against the survey's independent NumPy checkpoints on blood_01
(SURVEY.md Appendix B).  Citations: /root/reference/metabodecon/src/deconvolution/...

The reference commits no end-to-end golden output, so end-to-end parity is "unpinned"; what the
reference's own tests do pin is all here.
"""
import numpy as np
import pytest

import oracle as O


def approx(a, b, ulps=4):
    return abs(a - b) <= ulps * np.spacing(max(abs(a), abs(b)))


def test_second_derivative_kat():  # peak_selection/common.rs:50-58
    assert O.second_derivative([1.0, 2.0, 3.0, 2.0, 1.0]).tolist() == [0.0, -2.0, 0.0]


def test_peak_region_boundaries_kat():  # common.rs:61-68: centres [2,4,5,8], sb (3,7) -> (1,3)
    assert O.peak_region_boundaries([2, 4, 5, 8], (3, 7)) == (1, 3)
    assert O.peak_region_boundaries([2, 4, 5, 8], (9, 12)) == (0, 3)   # map_or fallbacks: 0 and len-1
    assert O.peak_region_boundaries([2, 4, 5, 8], (0, 100)) == (0, 3)


def test_detector_kats():  # detector.rs:170-228, replayed verbatim
    assert O.find_peak_centers([0.0, -2.0, 0.0]) == [2]
    assert O.find_peak_borders([0.5, -0.5, -1.0, 0.0, 0.5, 0.0], [3]) == [(2, 5)]
    assert O.find_peak_borders([0.0, 0.5, 0.0, -1.0, -0.5, 0.5], [4]) == [(2, 5)]
    assert O.find_peak_borders([1.0, 1.0, 1.0, 1.5, 1.0], [3]) == [(0, 4)]
    assert O.find_peak_borders([1.0, 1.5, 1.0, 1.0, 1.0], [3]) == [(2, 6)]
    assert O.find_peak_borders([1.0, 1.0, 1.0, 1.0, 1.0], [3]) == [(0, 6)]
    assert O.find_right_border([0.0, -2.0, -1.0, -0.5, 0.5][2:]) == 1
    assert O.find_right_border([0.0, -2.0, -1.0, 0.0, 0.5, 0.0][2:]) == 2
    assert O.find_right_border([1.0, 1.0, 1.0, 1.0, 1.0][2:]) == 3
    assert O.find_left_border([0.5, -0.5, -1.0, -2.0, 0.0][0:3]) == 1
    assert O.find_left_border([0.0, 0.5, 0.0, -1.0, -2.0, 0.0][0:4]) == 2
    assert O.find_left_border([1.0, 1.0, 1.0, 1.0, 1.0][0:3]) == 3
    # detect_peaks drops centres whose border search hit a sentinel (detector.rs:105)
    assert O.detect_peaks(np.array([0.0, -2.0, 0.0])).shape[0] == 0
    d2 = np.array([0.5, 1.0, 0.2, -0.4, -2.0, -0.3, 0.8, 0.1, 0.0])
    c = O.find_peak_centers(d2)
    assert c == [5] and O.detect_peaks(d2).tolist() == [[b[0], c[0], b[1]] for b in O.find_peak_borders(d2, c)]


def test_scorer_minimum_sum_kat():  # scorer.rs:94-107
    a = np.array([1.0, 2.0, 4.0, 2.0, 2.0, 5.0, 4.0, 3.0, 2.0])
    assert O.score_peak(a, 1, 3, 4) == 6.0
    assert O.score_peak(a, 5, 6, 9) == 7.0


def test_mean_sd_kat():  # noise_score_filter.rs:153-170 -> (4.0, 1.0)
    a = np.array([1.0, 2.0, 4.0, 2.0, 2.0, 5.0, 4.0, 3.0, 2.0])
    peaks = [(i - 1, i, i + 1) for i in (2, 4, 5, 8)]
    bl, br = 1, 3
    sfr = [O.score_peak(a, *p) for p in peaks[:bl] + peaks[br:]]
    mean, sd = O.mean_sd_scores(sfr)
    assert approx(mean, 4.0) and approx(sd, 1.0)


def test_mirror_shoulder_kats():  # peak_stencil.rs:176-208
    s = O.mirror_shoulder([1.0, 2.0, 3.0, 1.0, 2.0, 3.0])
    assert s.tolist() == [1.0, 2.0, 3.0, 1.0, 2.0, 1.0]
    s = O.mirror_shoulder([1.0, 2.0, 4.0, 3.0, 2.0, 1.0])
    assert s.tolist() == [0.0, 2.0, 4.0, 1.0, 2.0, 1.0]
    s = O.mirror_shoulder([1.0, 2.0, 3.0, 1.0, 5.0, 2.0])  # a true maximum is left alone
    assert s.tolist() == [1.0, 2.0, 3.0, 1.0, 5.0, 2.0]


def test_analytic_solve_kat():  # fitter_analytical.rs:188-196
    sfhw, hw2, maxp = O.solve_stencil([4.0, 8.0, 12.0, 5.0, 10.0, 5.0])
    assert approx(maxp, 8.0) and approx(np.sqrt(hw2), 4.0) and approx(sfhw / np.sqrt(hw2), 40.0)


def test_reduced_spectrum_gather_kat():  # reduced_spectrum.rs:61-86 via the fit's initial state
    x = np.array([1.0 + i for i in range(10)])
    y = np.array([10.0 - i for i in range(10)])
    peaks = [[2, 3, 4], [4, 5, 6], [6, 7, 8]]
    _, tr = O.fit_lorentzian(x, y, peaks, 1, trace=True)
    for k, (l, c, r) in enumerate(peaks):  # decreasing stencils are mirrored: x1 = 2*x2 - x3, y1 = y3
        want = O.solve_stencil([2.0 * x[c] - x[r], x[c], x[r], y[r], y[c], y[r]])
        assert tr[0, k].tolist() == want.tolist()


def test_lorentzian_evaluate_and_superposition_kats():  # lorentzian.rs:707-788, doc-test :593-604
    x = np.array([-5.0 + i for i in range(11)])
    one = [[1.0, 1.0, 0.0]]
    want = [1 / 26, 1 / 17, 1 / 10, 1 / 5, 1 / 2, 1.0, 1 / 2, 1 / 5, 1 / 10, 1 / 17, 1 / 26]
    got = O.superposition_vec(x, one)
    assert all(approx(g, w) for g, w in zip(got, want))
    lor = [[1.0, 0.5, -2.0], [2.0, 0.75, 0.0], [1.0, 0.5, 2.0]]
    want = [1.0 / 9.5 + 2.0 / 25.75 + 1.0 / 49.5, 1.0 / 4.5 + 2.0 / 16.75 + 1.0 / 36.5,
            1.0 / 1.5 + 2.0 / 9.75 + 1.0 / 25.5, 1.0 / 0.5 + 2.0 / 4.75 + 1.0 / 16.5,
            1.0 / 1.5 + 2.0 / 1.75 + 1.0 / 9.5, 1.0 / 4.5 + 2.0 / 0.75 + 1.0 / 4.5,
            1.0 / 9.5 + 2.0 / 1.75 + 1.0 / 1.5, 1.0 / 16.5 + 2.0 / 4.75 + 1.0 / 0.5,
            1.0 / 25.5 + 2.0 / 9.75 + 1.0 / 1.5, 1.0 / 36.5 + 2.0 / 16.75 + 1.0 / 4.5,
            1.0 / 49.5 + 2.0 / 25.75 + 1.0 / 9.5]
    for fn_par in (False, True):
        got = O.superposition_vec(x, lor, parallel=fn_par)
        assert got.tolist() == want  # same operation order -> identical bits
    for xi, wi in zip(x, want):
        assert O.superposition(float(xi), lor) == wi
    trip = [[0.03, 0.0009, 4.8], [0.02, 0.0004, 5.0], [0.03, 0.0009, 5.2]]
    assert abs(O.superposition(5.0, trip) - 51.466992) < 1e-6


def test_circular_buffer_semantics_through_smoothing():  # circular_buffer.rs:66-106 + moving_average.rs:53-83
    # window 3 on integers: exact arithmetic, growing / shrinking edge windows
    v = O.smooth_values([3.0, 6.0, 9.0, 12.0, 15.0], 1, 3)
    assert v.tolist() == [(3 + 6) * (1 / 2), (3 + 6 + 9) * (1 / 3), (6 + 9 + 12) * (1 / 3), (9 + 12 + 15) * (1 / 3), (12 + 15) * (1 / 2)]
    # even window 4 (right = 2): asymmetric window [i-1, i+2]
    v = O.smooth_values([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], 1, 4)
    assert v.tolist() == [6 * (1 / 3), 10 * (1 / 4), 14 * (1 / 4), 18 * (1 / 4), 15 * (1 / 3), 11 * (1 / 2)]
    # the recurrence is a running sum, not a re-summed window: drift is reproduced, not avoided
    y = np.array([1e16, 1.0, -1e16, 1.0, 1.0, 1.0, 1.0])
    v = O.smooth_values(y, 1, 3)
    s = 0.0 + 1e16
    s = s + 1.0
    out0 = s * (1 / 2)
    s = s + -1e16
    out1 = s * (1 / 3)
    s = (s + 1.0) - 1e16
    assert v[0] == out0 and v[1] == out1 and v[2] == s * (1 / 3)


def test_index_helpers(blood_arrays):  # spectrum.rs:741-746, deconvoluter.rs:865-904 (Appendix B)
    x, _ = blood_arrays
    assert float(x[0]).hex() == "0x1.d9f77af64063ap+3" and float(x[1]).hex() == "0x1.d9f63a94e72a4p+3"
    assert O.signal_boundaries_indices(x, (11.8, -2.2)) == (19712, 111354)
    assert O.ignore_region_indices(x, (11.8, -2.2), [(4.7, 4.9)]).tolist() == [[64879, 66187]]
    # doc-test spectrum.rs:725-739: x = 1..5, boundaries (2.25, 3.75) -> (1, 3)
    assert O.signal_boundaries_indices([1.0, 2.0, 3.0, 4.0, 5.0], (2.25, 3.75)) == (1, 3)
    # regions entirely outside the signal region are dropped
    assert O.ignore_region_indices(x, (11.8, -2.2), [(12.5, 13.0)]).shape[0] == 0
    # literal replay of deconvoluter.rs:886-887 on a DECREASING axis: `start` (the smaller ppm) maps
    # to the larger index, so max(.., lower) / min(.., upper) do not clamp a region that sticks out
    # of the signal region; floor((11-x0)/step) = 24949, ceil((13-x0)/step) = 11858
    assert O.ignore_region_indices(x, (11.8, -2.2), [(11.0, 13.0)]).tolist() == [[11858, 24949]]
    # on an increasing axis the same formulas do clamp
    xi = np.arange(0.0, 100.0, 1.0)
    assert O.ignore_region_indices(xi, (10.0, 90.0), [(5.0, 20.0)]).tolist() == [[10, 20]]


def test_blood_checkpoints(blood_arrays):  # SURVEY.md Appendix B, bit for bit
    x, y = blood_arrays
    r = O.deconvolute_spectrum(O.Settings(ignore_regions=[(4.7, 4.9)]), x, y, (11.8, -2.2))
    assert r.status == O.OK
    assert [float(v).hex() for v in r.smoothed[:3]] == ["-0x1.4ff71c71c71c6p+10", "-0x1.5a0aaaaaaaaa9p+10", "-0x1.56ed097b425eap+10"]
    assert float(r.smoothed[65536]).hex() == "0x1.3b35c1097b448p+22" and float(r.smoothed[-1]).hex() == "0x1.26138e3b895b4p+9"
    assert (r.n_detected, r.n_after_ignore, r.region, r.n_sfr) == (16100, 15980, (2527, 13391), 5116)
    assert float(r.mean).hex() == "0x1.1ac34b169a537p+8" and float(r.sd).hex() == "0x1.6a11671179e58p+7"
    assert len(r.peaks) == 981 and r.peaks[0].tolist() == [41583, 41585, 41588] and r.peaks[-1].tolist() == [97895, 97896, 97898]
    assert len(r.lorentzians) == 760
    assert [float(v).hex() for v in r.lorentzians[0]] == ["0x1.084fd4b50b8dfp-4", "0x1.1c9f87d1217eep-22", "0x1.0eac0d0cd9679p+3"]
    assert [float(v).hex() for v in r.lorentzians[2]] == ["0x1.35c5dcf473907p-7", "0x1.40a12ec4d6881p-21", "0x1.06b6158530021p+3"]
    assert float(r.mse).hex() == "0x1.0808a64fe177ep+35"
    r2 = O.deconvolute_spectrum(O.Settings(), x, y, (11.8, -2.2))
    assert (len(r2.peaks), len(r2.lorentzians)) == (992, 766)
    # serial and rayon-shaped variants agree bit for bit
    r3 = O.deconvolute_spectrum(O.Settings(), x, y, (11.8, -2.2), parallel=True)
    assert np.array_equal(r2.lorentzians.view(np.uint64), r3.lorentzians.view(np.uint64)) and r2.mse == r3.mse


def test_sim_spectrum_matches_generating_parameters(sim_arrays, golden_dir):
    import os
    x, y = sim_arrays
    r = O.deconvolute_spectrum(O.Settings(), x, y, (3.55, 3.35))
    assert r.status == O.OK and len(r.lorentzians) > 5
    truth = np.loadtxt(os.path.join(golden_dir, "bruker", "sim_01", "lorentzians.csv"), delimiter=",", skiprows=1)
    step = abs(x[1] - x[0])
    dist = np.min(np.abs(r.lorentzians[:, 2][:, None] - truth[:, 2][None, :]), axis=1)
    assert np.median(dist) < 2 * step


def test_error_paths():
    x = np.linspace(10.0, 0.0, 2000)
    st = O.deconvolute_spectrum(O.Settings(), x, np.zeros(2000), (9.0, 1.0)).status
    assert st == O.NO_PEAKS_DETECTED
    rng = np.random.default_rng(3)
    y = rng.normal(0, 1, 2000)
    r = O.deconvolute_spectrum(O.Settings(), x, y, (5.0, 4.99))
    assert r.status == O.EMPTY_SIGNAL_REGION
    r = O.deconvolute_spectrum(O.Settings(selection_kind=O.SELECT_DETECTOR_ONLY), x, y, (9.0, 1.0))
    assert r.status == O.OK and len(r.peaks) > 0


def test_batch_baseline_matches_single(blood_arrays):
    x, y = blood_arrays
    ys = np.stack([y, y * 0.5 + 3.0])
    st, lors, mse, nsel = O.par_deconvolute_spectra(O.Settings(), x, ys, (11.8, -2.2))
    one = O.deconvolute_spectrum(O.Settings(), x, y, (11.8, -2.2))
    assert st == O.OK and np.array_equal(lors[0].view(np.uint64), one.lorentzians.view(np.uint64)) and mse[0] == one.mse


def test_optimize_settings_grid_and_argmin(sim_arrays):  # deconvoluter.rs:761-825
    """The oracle's optimiser must agree with 810 individual oracle deconvolutions taken in the
    reference's iteration order (smoothing outer, threshold, fit iterations inner; first minimum)."""
    x, y = sim_arrays
    sb = (3.55, 3.35) if x[0] > x[1] else (3.35, 3.55)
    status, best, mse, all_mse = O.optimize_settings(O.Settings(), x, y, sb)
    assert status == O.OK
    want = []
    for iterations in range(2, 11):
        for window in (3, 5, 7):
            for c in range(10):
                for fit in (5, 10, 15):
                    thr = 5.0 + float(c) * (8.0 - 5.0) / 9.0
                    r = O.deconvolute_spectrum(O.Settings(smoothing_iterations=iterations, smoothing_window=window,
                                                          threshold=thr, fitting_iterations=fit), x, y, sb)
                    assert r.status == O.OK
                    want.append(((iterations, window, thr, fit), r.mse))
    assert [m for _, m in want] == all_mse.tolist()
    k = int(np.argmin(all_mse))  # numpy argmin returns the first minimum, like Iterator::min_by
    assert best == want[k][0] and mse == want[k][1]
    assert 5.0 <= best[2] <= 8.0 and best[3] in (5, 10, 15)


# ------------------------------------------------------------------------------------------------
# Two independent restatements (C: oracle/mdb_oracle.c, NumPy: oracle/numpy_restatement.py) must agree
# bit for bit -- the strongest pin available without a Rust toolchain.
# ------------------------------------------------------------------------------------------------
def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _agree(x, y, sb, ignore, what):
    from oracle import numpy_restatement as NR
    sb = tuple(sb)
    got = NR.deconvolute(x, y, sb, ignore_regions=ignore)
    want = O.deconvolute_spectrum(O.Settings(ignore_regions=ignore), x, y, sb)
    assert want.status == O.OK
    assert np.array_equal(_bits(got["smoothed"]), _bits(want.smoothed)), f"{what}: smoothing"
    assert np.array_equal(got["peaks"], want.peaks.astype(np.int64)), f"{what}: selected peaks"
    assert _bits([got["mean"], got["sd"]]).tolist() == _bits([want.mean, want.sd]).tolist(), f"{what}: mean/sd"
    assert got["lorentzians"].shape == want.lorentzians.shape, f"{what}: retained count"
    assert np.array_equal(_bits(got["lorentzians"]), _bits(want.lorentzians)), f"{what}: lorentzians"
    assert _bits([got["mse"]]).tolist() == _bits([want.mse]).tolist(), f"{what}: mse"


def test_numpy_and_c_restatements_agree_sim(sim_arrays):
    x, y = sim_arrays
    _agree(x, y, (3.55, 3.35) if x[0] > x[1] else (3.35, 3.55), None, "sim_01")


def test_numpy_and_c_restatements_agree_blood(blood_arrays):
    x, y = blood_arrays
    _agree(x, y, (11.8, -2.2), [(4.7, 4.9)], "blood_01 with the water region ignored")


def test_numpy_and_c_restatements_agree_synthetic():
    import synth
    n = 16384
    x = synth.axis(n)
    _agree(x, synth.config3(3, n=n, x=x), synth.SIGNAL_BOUNDARIES, None, "synthetic f64")
    _agree(x, synth.config3(4, n=n, x=x, integer=True), synth.SIGNAL_BOUNDARIES, [(1.0, 1.2)], "synthetic integer")


def _random_spectrum(rng, n):
    """Random small spectrum with the features that stress the path: integer values (exact ties in
    d2), flat stretches, narrow and wide Lorentzians, noise."""
    x = 10.0 - np.arange(n) * (12.0 / (n - 1))
    y = rng.normal(0.0, 30.0, n)
    for _ in range(int(rng.integers(3, 25))):
        m, hw, a = rng.uniform(-1.5, 9.5), np.exp(rng.uniform(np.log(3e-3), np.log(8e-2))), np.exp(rng.uniform(np.log(2e3), np.log(1e6)))
        y += a * hw * hw / (hw * hw + (x - m) ** 2)
    if rng.random() < 0.5:
        y = np.rint(y)
    if rng.random() < 0.3:
        lo = int(rng.integers(0, n - 40))
        y[lo:lo + int(rng.integers(5, 40))] = np.rint(y[lo])
    return x, y


def test_randomised_differential_c_vs_numpy():
    """40 random spectra x random settings: every stage of the two restatements must agree in bits
    (or both report the same error condition)."""
    from oracle import numpy_restatement as NR
    rng = np.random.default_rng(20261018)
    checked = 0
    for case in range(40):
        n = int(rng.integers(200, 3000))
        x, y = _random_spectrum(rng, n)
        iters, window = int(rng.integers(1, 6)), int(rng.choice([2, 3, 4, 5, 7, 9]))
        thr, fit = float(rng.uniform(0.5, 8.0)), int(rng.integers(1, 12))
        sb = (float(rng.uniform(7.0, 9.5)), float(rng.uniform(-1.5, 1.0)))
        want = O.deconvolute_spectrum(O.Settings(smoothing_iterations=iters, smoothing_window=window, threshold=thr,
                                                 fitting_iterations=fit), x, y, sb)
        sm = NR.smooth_values(y, iters, window)
        assert np.array_equal(_bits(sm), _bits(want.smoothed)), f"case {case}: smoothing ({iters}, {window})"
        if want.status != O.OK:
            continue
        got = NR.deconvolute(x, y, sb, smoothing=(iters, window), threshold=thr, fit_iterations=fit)
        assert np.array_equal(got["peaks"], want.peaks.astype(np.int64)), f"case {case}: peaks"
        assert got["lorentzians"].shape == want.lorentzians.shape, f"case {case}: retained"
        assert np.array_equal(_bits(got["lorentzians"]), _bits(want.lorentzians)), f"case {case}: lorentzians"
        a, b = _bits([got["mse"]])[0], _bits([want.mse])[0]
        assert a == b or (np.isnan(got["mse"]) and np.isnan(want.mse)), f"case {case}: mse"
        checked += 1
    assert checked >= 25
