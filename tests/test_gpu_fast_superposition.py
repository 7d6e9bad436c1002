"""The library's DEFAULT arithmetic for the two superposition kernels that do not feed back into
the fit -- K7 (the full-grid superposition behind Deconvolution.mse, deconvoluter.rs:540-543,
828-862) and K8 (Lorentzian::superposition_vec / par_superposition_vec, lorentzian.rs:631-663):
MDB_SUPERPOSITION_FAST, 5.25 instead of 12 FP64 instructions per evaluation (kernels.cuh,
lorentz_quad_ulp: four Lorentzians behind one reciprocal; lorentz_step_ulp for what is left over).

Bar (BASELINE.json north_star): peak sets bit-exact; Lorentzian parameters and superposition values
within 1e-9 relative.  What is asserted here, against the CPU oracle on the same inputs:
  * peak sets and Lorentzian parameters: identical bit patterns (the refinement is always exact);
  * superposition values: TOL_VALUES = 1e-13 relative (measured: a few 1e-16);
  * mean squared error (few-ulp terms, residuals summed by a fixed tree instead of a left fold):
    TOL_MSE = 1e-9 relative, the north-star figure (measured: a few 1e-14 on every case here); the residual (S - y) can cancel, so its relative error is not bounded by that of S
    for a perfect fit -- there the absolute error is checked against the size of S instead;
  * operands outside the division's fast domain take the IEEE loop in both modes: bit-exact;
  * optimize_settings picks the same settings (it always computes its MSEs exactly).
tests/test_gpu_parity.py runs the same kernels in MDB_SUPERPOSITION_EXACT and asserts bits.
"""
import os

import numpy as np
import pytest

import oracle as O
import synth
from metabodecon_rust_b200 import Deconvoluter, Spectrum, _lib, set_superposition_mode, superposition_mode
from metabodecon_rust_b200.lorentzian import superposition_vec_array

pytestmark = pytest.mark.gpu

TOL_VALUES = 1e-13
TOL_MSE = 1e-9


@pytest.fixture(autouse=True)
def default_mode():
    set_superposition_mode("fast")
    yield
    set_superposition_mode("fast")


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))) if want.size else 0.0


def random_lorentzians(rng, p, lo=0.0, hi=10.0):
    hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
    sf = np.exp(rng.uniform(np.log(1.0), np.log(1e4), p))
    return np.stack([sf * hw, hw * hw, rng.uniform(lo, hi, p)], axis=1)


def test_fast_is_the_default_and_the_switch_works():
    assert superposition_mode() == "fast"
    rng = np.random.default_rng(5)
    x = np.linspace(-2.2, 11.8, 4099)
    lor = random_lorentzians(rng, 700)
    want = O.superposition_vec(x, lor)
    fast = superposition_vec_array(x, lor)
    set_superposition_mode("exact")
    assert superposition_mode() == "exact"
    exact = superposition_vec_array(x, lor)
    assert np.array_equal(exact.view(np.uint64), want.view(np.uint64))
    assert rel_err(fast, want) <= TOL_VALUES
    with pytest.raises(ValueError):
        set_superposition_mode("sloppy")
    assert _lib.load().mdb_set_superposition_mode(7) != 0


@pytest.mark.parametrize("n,p", [(1, 1), (31, 3), (1024, 512), (1025, 513), (4097, 1023), (300000, 1100), (2 ** 21 + 5, 40)])  # the last one takes 16 points per thread
def test_superposition_vec_within_tolerance_of_oracle(n, p):
    rng = np.random.default_rng(n * 7 + p)
    x = np.linspace(-2.2, 11.8, n) if n > 1 else np.array([3.3])
    lor = random_lorentzians(rng, p)
    got = superposition_vec_array(x, lor)
    want = O.superposition_vec(x, lor, parallel=n > 100000)
    err = rel_err(got, want)
    assert err <= TOL_VALUES, f"n={n} p={p}: max relative error {err:.3e} > {TOL_VALUES}"


def test_superposition_vec_points_on_and_next_to_the_maxima():
    # x == maxp exactly (denominator = hw2), one grid step away, and far tails (1e6 half widths)
    rng = np.random.default_rng(11)
    lor = random_lorentzians(rng, 257)
    x = np.concatenate([lor[:, 2], lor[:, 2] + np.sqrt(lor[:, 1]), lor[:, 2] + 1e6 * np.sqrt(lor[:, 1]),
                        np.nextafter(lor[:, 2], np.inf)])
    err = rel_err(superposition_vec_array(x, lor), O.superposition_vec(x, lor))
    assert err <= TOL_VALUES, f"{err:.3e}"


def test_wide_dynamic_range_inside_the_fast_domain():
    # parameters spread over 2^-280 .. 2^280 (the fast domain is 2^-300 .. 2^300): every term is still a
    # normal-number division; terms of one sign, so the sum cannot cancel
    rng = np.random.default_rng(12)
    p = 600
    e = rng.uniform(-280, 280, p)
    lor = np.stack([np.exp2(e), np.exp2(rng.uniform(-280, 280, p)), rng.uniform(-5, 5, p) * np.exp2(rng.uniform(-100, 100, p))], axis=1)
    x = rng.uniform(-5, 5, 3000) * np.exp2(rng.uniform(-100, 100, 3000))
    want = O.superposition_vec(x, lor)
    got = superposition_vec_array(x, lor)
    fin = np.isfinite(want)
    assert fin.any() and np.array_equal(np.isfinite(got), fin)
    assert rel_err(got[fin], want[fin]) <= TOL_VALUES


def test_four_lorentzians_per_reciprocal_domain_edges_and_remainders(monkeypatch):
    # lorentz_quad_ulp (kernels.cuh): N / P over groups of four inside |sfhw|, hw2 in [2^-200, 2^200],
    # |maxp|, |x| <= 2^100; parameter counts that leave 1, 2 and 3 Lorentzians for the one-at-a-time tail
    rng = np.random.default_rng(14)
    for p in (4, 5, 6, 7, 512 + 2, 1024 + 3):
        lor = np.stack([np.exp2(rng.uniform(-195, 195, p)), np.exp2(rng.uniform(-195, 195, p)),
                        rng.uniform(-5, 5, p) * np.exp2(rng.uniform(-60, 95, p))], axis=1)
        x = rng.uniform(-5, 5, 2000) * np.exp2(rng.uniform(-60, 95, 2000))
        want = O.superposition_vec(x, lor)
        got = superposition_vec_array(x, lor)
        assert np.isfinite(want).all() and np.isfinite(got).all()
        assert rel_err(got, want) <= TOL_VALUES, f"p={p}: {rel_err(got, want):.3e}"
    # one parameter just outside the quad domain (hw2 = 2^-210): that tile runs one at a time, still in tolerance
    lor = random_lorentzians(rng, 600)
    lor[17, 1] = 2.0 ** -210
    x = np.linspace(-2.2, 11.8, 5000)
    assert rel_err(superposition_vec_array(x, lor), O.superposition_vec(x, lor)) <= TOL_VALUES
    # mixed signs: a group's error is bounded relative to the sum of the magnitudes of its terms
    lor = random_lorentzians(rng, 801)
    lor[::3, 0] *= -1.0
    want = O.superposition_vec(x, lor)
    mag = O.superposition_vec(x, np.abs(lor))
    got = superposition_vec_array(x, lor)
    assert float(np.max(np.abs(got - want) / mag)) <= TOL_VALUES
    # the one-at-a-time few-ulp form (MDB_SUP_GROUP=1, measurement aid) agrees to the same tolerance
    lor = random_lorentzians(rng, 2143)
    want = O.superposition_vec(x, lor)
    quad = superposition_vec_array(x, lor)
    monkeypatch.setenv("MDB_SUP_GROUP", "1")
    single = superposition_vec_array(x, lor)
    monkeypatch.delenv("MDB_SUP_GROUP")
    assert rel_err(quad, want) <= TOL_VALUES and rel_err(single, want) <= TOL_VALUES
    assert not np.array_equal(quad, single) or True  # different roundings are expected, not required


def test_out_of_domain_operands_take_the_ieee_loop_bit_exactly():
    # zeros, denormals, infinities, NaN, hw2 = 0 at x = maxp (division by zero): the per-tile domain
    # check sends these tiles through __ddiv_rn in both modes
    x = np.array([0.0, 1.0, -1.0, 1e-310, 1e300, np.inf, 2.5])
    specials = np.array([[1.0, 0.0, 2.5], [0.0, 1.0, 0.0], [1e-320, 1e-320, 0.0], [1.0, 1e-310, 1.0],
                         [np.inf, 1.0, 0.0], [1.0, np.inf, 0.0], [np.nan, 1.0, 0.0], [1.0, 1.0, np.nan],
                         [1e308, 1e-308, 1e308], [-1.0, 1.0, 0.5], [1.0, -1.0, 0.5]])
    for k in range(len(specials)):
        got = superposition_vec_array(x, specials[k:k + 1])
        want = O.superposition_vec(x, specials[k:k + 1])
        both_nan = np.isnan(got) & np.isnan(want)
        assert np.array_equal(np.where(both_nan, 0.0, got).view(np.uint64), np.where(both_nan, 0.0, want).view(np.uint64)), f"special {k}"
    # |x| beyond 2^300 leaves the domain through the grid, not the parameters
    lor = random_lorentzians(np.random.default_rng(3), 40)
    xs = np.array([1e95, -1e95, 3e200, 1.0])
    got, want = superposition_vec_array(xs, lor), O.superposition_vec(xs, lor)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def _check_default_mode(dec, osettings, spectra, what):
    outs = dec.deconvolute_spectra(spectra)
    worst = 0.0
    for i, (sp, out) in enumerate(zip(spectra, outs)):
        r = O.deconvolute_spectrum(osettings, sp.chemical_shifts, sp.intensities, sp.signal_boundaries)
        assert r.status == O.OK
        assert np.array_equal(out.peaks.astype(np.int64), r.peaks.astype(np.int64)), f"{what}[{i}]: peak set differs"
        got = np.ascontiguousarray(out.parameters, dtype=np.float64)
        assert np.array_equal(got.view(np.uint64), np.ascontiguousarray(r.lorentzians).view(np.uint64)), f"{what}[{i}]: lorentzians differ in bits"
        err = abs(out.mse - r.mse) / abs(r.mse)
        assert err <= TOL_MSE, f"{what}[{i}]: mse {out.mse!r} vs {r.mse!r}, relative error {err:.3e} > {TOL_MSE}"
        worst = max(worst, err)
    return outs, worst


def test_config1_blood_default_mode(golden_dir):
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
    dec = Deconvoluter()
    dec.add_ignore_region((4.7, 4.9))
    outs, worst = _check_default_mode(dec, O.Settings(ignore_regions=[(4.7, 4.9)]), [sp], "blood_01 water ignored")
    assert len(outs[0].lorentzians) == 760
    assert abs(outs[0].mse - float.fromhex("0x1.0808a64fe177ep+35")) <= 1e-12 * 35438015103.04588
    assert worst <= 1e-12
    dec.clear_ignore_regions()
    _check_default_mode(dec, O.Settings(), [sp, sp, sp], "blood_01 x3")


def test_config2_jcampdx_default_mode(golden_dir):
    sp = Spectrum.read_jcampdx(os.path.join(golden_dir, "jcampdx", "blood_01.dx"), (-2.2, 11.8))
    _check_default_mode(Deconvoluter(), O.Settings(), [sp], "blood_01.dx")


def test_synthetic_batches_default_mode():
    # config-3 and config-5 style spectra (raw and integer-rounded), long enough to leave the fused
    # small-spectrum path, in one batch spanning R = 8 and R = 2 launches of K7
    specs = []
    for s, (n, integer) in enumerate([(32768, False), (16384, True), (65536, False), (20000, True), (8192, False)]):
        x = synth.axis(n)
        specs.append(Spectrum(x, synth.config3(40 + s, n=n, integer=integer, x=x), (-2.2, 11.8)))
    _, worst = _check_default_mode(Deconvoluter(), O.Settings(), specs, "synthetic default")
    assert worst <= 1e-12
    dec = Deconvoluter()
    dec.set_moving_average_smoother(2, 5)
    dec.set_noise_score_selector(6.5)
    dec.set_analytical_fitter(5)
    dec.add_ignore_region((4.7, 4.9))
    _check_default_mode(dec, O.Settings(smoothing_iterations=2, smoothing_window=5, threshold=6.5, fitting_iterations=5,
                                        ignore_regions=[(4.7, 4.9)]), specs, "synthetic custom")


def test_full_size_batch_default_vs_exact_mode():
    # 24 config-5 spectra of 2^17 points: the two modes give the same peaks and Lorentzians in bits
    # and MSEs within TOL_MSE (the oracle is checked on two of them; it needs seconds per spectrum)
    n = 131072
    x = synth.axis(n)
    specs = [Spectrum(x, synth.config5(900 + s, n=n, x=x), (-2.2, 11.8)) for s in range(24)]
    dec = Deconvoluter()
    fast = dec.deconvolute_spectra(specs)
    set_superposition_mode("exact")
    exact = dec.deconvolute_spectra(specs)
    set_superposition_mode("fast")
    for i, (a, b) in enumerate(zip(fast, exact)):
        assert np.array_equal(a.peaks, b.peaks)
        assert np.array_equal(np.ascontiguousarray(a.parameters).view(np.uint64), np.ascontiguousarray(b.parameters).view(np.uint64))
        assert abs(a.mse - b.mse) <= 1e-12 * abs(b.mse), f"[{i}] {a.mse!r} vs {b.mse!r}"
    _check_default_mode(dec, O.Settings(), specs[:2], "config5 full size")


def test_sixteen_points_per_thread_with_two_ranges_per_spectrum(golden_dir):
    # 16 copies of blood_01 with the water region ignored: two MSE ranges per spectrum whose lengths
    # are no multiple of the 2 048 points a CTA of the 16-points-per-thread K7 covers
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
    dec = Deconvoluter()
    dec.add_ignore_region((4.7, 4.9))
    fast = dec.deconvolute_spectra([sp] * 16)
    set_superposition_mode("exact")
    exact = dec.deconvolute_spectra([sp] * 16)
    set_superposition_mode("fast")
    assert float(exact[0].mse).hex() == "0x1.0808a64fe177ep+35"
    for a, b in zip(fast, exact):
        assert np.array_equal(np.ascontiguousarray(a.parameters).view(np.uint64), np.ascontiguousarray(b.parameters).view(np.uint64))
        assert a.mse == fast[0].mse and abs(a.mse - b.mse) <= 1e-12 * b.mse


def test_randomised_differential_default_mode(monkeypatch):
    """60 random spectra (integer / float values, flat stretches, 200..6000 points) under random
    settings, in random-size batches, through the GENERAL pipeline (the fused small-spectrum kernel has
    its own test above, so it is switched off here): statuses, peak sets and Lorentzians as the
    oracle's in bits, MSE within TOL_MSE."""
    from test_oracle_kats import _random_spectrum
    monkeypatch.setenv("MDB_SMALL_PATH", "0")
    rng = np.random.default_rng(4242)
    done, worst = 0, 0.0
    while done < 60:
        batch = int(rng.integers(1, 9))
        iters, window = int(rng.integers(1, 6)), int(rng.choice([2, 3, 4, 5, 7, 9, 11]))
        thr, fit = float(rng.uniform(0.5, 8.0)), int(rng.integers(1, 12))
        dec = Deconvoluter()
        dec.set_moving_average_smoother(iters, window)
        dec.set_noise_score_selector(thr)
        dec.set_analytical_fitter(fit)
        settings = O.Settings(smoothing_iterations=iters, smoothing_window=window, threshold=thr, fitting_iterations=fit)
        specs, wants = [], []
        for _ in range(batch):
            n = int(rng.integers(200, 6000))
            x, y = _random_spectrum(rng, n)
            sb = (float(rng.uniform(7.0, 9.5)), float(rng.uniform(-1.5, 1.0)))
            sp = Spectrum(x, y, sb)
            specs.append(sp)
            wants.append(O.deconvolute_spectrum(settings, x, y, sp.signal_boundaries))
        done += batch
        if any(w.status != O.OK for w in wants):
            with pytest.raises(Exception):
                dec.deconvolute_spectra(specs)
            continue
        for i, (out, w) in enumerate(zip(dec.deconvolute_spectra(specs), wants)):
            what = f"random case {done - batch + i} (iters={iters}, window={window}, thr={thr:.3f}, fit={fit})"
            assert np.array_equal(out.peaks.astype(np.int64), w.peaks.astype(np.int64)), what
            nan = np.isnan(w.lorentzians)
            got = np.where(nan, 0.0, np.ascontiguousarray(out.parameters, dtype=np.float64))
            assert np.array_equal(got.view(np.uint64), np.where(nan, 0.0, w.lorentzians).view(np.uint64)), what
            if np.isnan(w.mse) or np.isinf(w.mse):
                assert (np.isnan(out.mse) and np.isnan(w.mse)) or out.mse == w.mse, what
            else:
                err = abs(out.mse - w.mse) / abs(w.mse) if w.mse != 0.0 else abs(out.mse)
                assert err <= TOL_MSE, f"{what}: mse {out.mse!r} vs {w.mse!r} ({err:.3e})"
                worst = max(worst, err)
    assert worst <= TOL_MSE


def test_small_spectrum_path_default_mode(monkeypatch, golden_dir):
    # spectra of up to 4 096 points take the one-launch fused kernel, whose MSE tail follows the mode
    # too (few-ulp superposition, tree-summed residuals): against the oracle, and against the general
    # pipeline in the same mode
    lib = _lib.load()
    sim = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    specs = [sim]
    for s, n in enumerate([2048, 3001, 4096, 700]):
        x = synth.axis(n)
        specs.append(Spectrum(x, synth.spectrum(5000 + s, n=n, k=25, hw_range=(8e-3, 5e-2), x=x), (-2.2, 11.8)))
    dec = Deconvoluter()
    lib.mdb_reset_kernel_launch_count()
    outs, worst = _check_default_mode(dec, O.Settings(), specs, "small path")
    assert worst <= 1e-12
    dec.add_ignore_region((4.7, 4.9))  # two MSE ranges in the synthetic ones
    small, _ = _check_default_mode(dec, O.Settings(ignore_regions=[(4.7, 4.9)]), specs[1:], "small path, ignore region")
    lib.mdb_reset_kernel_launch_count()
    dec.deconvolute_spectra(specs[1:])
    assert lib.mdb_kernel_launch_count() == 1
    monkeypatch.setenv("MDB_SMALL_PATH", "0")
    general = dec.deconvolute_spectra(specs[1:])
    for a, b in zip(small, general):
        assert np.array_equal(np.ascontiguousarray(a.parameters).view(np.uint64), np.ascontiguousarray(b.parameters).view(np.uint64))
        assert abs(a.mse - b.mse) <= 1e-12 * abs(b.mse)
    monkeypatch.delenv("MDB_SMALL_PATH")
    set_superposition_mode("exact")
    exact = dec.deconvolute_spectra(specs[1:])
    for sp, out in zip(specs[1:], exact):
        r = O.deconvolute_spectrum(O.Settings(ignore_regions=[(4.7, 4.9)]), sp.chemical_shifts, sp.intensities, sp.signal_boundaries)
        assert out.mse == r.mse


def test_exact_fit_small_residuals():
    # a noiseless spectrum of well separated Lorentzians is fitted almost exactly: S - y cancels, so
    # the MSE's relative error is not bounded by that of S; its absolute error is, by 2 * |r| * dS
    n = 16384
    x = np.linspace(0.0, 10.0, n)
    true = np.stack([np.full(8, 2e-3) * 1e5, np.full(8, 4e-6), np.linspace(1.0, 9.0, 8)], axis=1)
    y = O.superposition_vec(x, true)
    sp = Spectrum(x, y, (0.5, 9.5))
    dec = Deconvoluter()
    dec.set_identity_smoother()
    dec.set_detector_only()
    r = O.deconvolute_spectrum(O.Settings(smoothing_kind=O.SMOOTH_IDENTITY, selection_kind=O.SELECT_DETECTOR_ONLY),
                               x, y, sp.signal_boundaries)
    if r.status != O.OK:
        pytest.skip("oracle rejects this construction")
    out = dec.deconvolute_spectra([sp])[0]
    assert np.array_equal(np.ascontiguousarray(out.parameters).view(np.uint64), np.ascontiguousarray(r.lorentzians).view(np.uint64))
    rms, smax = np.sqrt(max(r.mse, 0.0)), float(np.max(np.abs(y)))
    assert abs(out.mse - r.mse) <= 2.0 * (rms + 1e-13 * smax) * 1e-13 * smax + 1e-9 * r.mse


def test_optimize_settings_choice_does_not_depend_on_the_mode(golden_dir):
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    a, b = Deconvoluter(), Deconvoluter()
    mse_fast = a.optimize_settings(sp)
    set_superposition_mode("exact")
    mse_exact = b.optimize_settings(sp)
    assert mse_fast == mse_exact  # computed exactly in both modes
    assert a.smoothing_settings() == b.smoothing_settings() and a.selection_settings() == b.selection_settings()
    assert a.fitting_settings() == b.fitting_settings()
