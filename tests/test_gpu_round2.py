"""GPU parity tests added in round 2: per-deconvoluter superposition mode, axis views of different
lengths in one batch, pageable / pinned / device inputs, grid sharding and chunked streaming in
superposition_vec, full-size integer-valued spectra (SURVEY.md 8d variant B), and the opt-in
arithmetic experiments of the refinement kernel.  Everything is compared with the oracle (or with
the exact GPU path the oracle has already pinned) through the C ABI."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle as O
import synth
from metabodecon_rust_b200 import Deconvoluter, Spectrum, _lib, exceptions, set_superposition_mode
from metabodecon_rust_b200.lorentzian import Lorentzian, superposition_vec_array

pytestmark = pytest.mark.gpu

SB = (-2.2, 11.8)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same_bits(got, want, what=""):
    got, want = np.ascontiguousarray(got, dtype=np.float64), np.ascontiguousarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    assert np.array_equal(bits(got), bits(want)), f"{what}: bit patterns differ"


@pytest.fixture(autouse=True)
def restore_default_mode():
    set_superposition_mode("fast")
    yield
    set_superposition_mode("fast")


def fast_spectrum(seed, n=131072, k=500, hw_range=(5e-4, 3e-3), integer=False, x=None):
    """A synthetic spectrum of the tests/synth.py family (same parameter and noise streams), the clean
    signal summed by the oracle's OpenMP superposition instead of NumPy broadcasting (seconds -> ms)."""
    rng = np.random.Generator(np.random.PCG64(20260000 + seed))
    x = synth.axis(n) if x is None else x
    maxp = rng.uniform(-2.0, 11.6, k)
    hw = np.exp(rng.uniform(np.log(hw_range[0]), np.log(hw_range[1]), k))
    amp = np.exp(rng.uniform(np.log(1e4), np.log(1e7), k))
    noise = rng.normal(0.0, 300.0, n)
    y = O.superposition_vec(x, np.stack([amp * hw * hw, hw * hw, maxp], axis=1), parallel=True) + noise
    return np.rint(y) if integer else y


def check_vs_oracle(outs, specs, settings=None, exact_mse=True, what=""):
    settings = settings or O.Settings()
    for i, (sp, out) in enumerate(zip(specs, outs)):
        r = O.deconvolute_spectrum(settings, sp.chemical_shifts, sp.intensities, sp.signal_boundaries)
        assert r.status == O.OK
        assert np.array_equal(out.peaks.astype(np.int64), r.peaks.astype(np.int64)), f"{what}[{i}]: peak set differs"
        assert_same_bits(out.parameters, r.lorentzians, f"{what}[{i}]: lorentzians")
        if exact_mse:
            assert out.mse == r.mse, f"{what}[{i}]: mse {out.mse!r} vs {r.mse!r}"
        else:
            assert abs(out.mse - r.mse) <= 1e-9 * abs(r.mse), f"{what}[{i}]: mse {out.mse!r} vs {r.mse!r}"


# ------------------------------------------------------------------------------ advisor: axis views
@pytest.mark.parametrize("n_long", [9000, 4000])  # the general pipeline and the fused small-spectrum path
def test_two_views_of_one_axis_array_with_different_lengths(n_long):
    """big[:n_short] and big[:n_long] share their base pointer: the device copy of the axis must be keyed
    by (pointer, length), or the longer spectrum reads past the shorter one's row."""
    n_short = n_long - 1500
    big = synth.axis(n_long)  # uniform spacing: every prefix is a valid axis
    ys = [synth.spectrum(41 + s, n=n_long, k=60, hw_range=(4e-3, 2e-2), x=big) for s in range(4)]
    lo = float(big[n_short - 1]) + 1e-9
    specs = [Spectrum(big[:n_short], ys[0][:n_short], (lo, 11.8)), Spectrum(big, ys[1], SB),
             Spectrum(big[:n_short], ys[2][:n_short], (lo, 11.8)), Spectrum(big, ys[3], SB)]
    assert specs[0].chemical_shifts.ctypes.data == specs[1].chemical_shifts.ctypes.data
    dec = Deconvoluter()
    dec.set_superposition_mode("exact")
    check_vs_oracle(dec.deconvolute_spectra(specs), specs, what=f"axis views {n_long}")
    # the longer view first, too
    order = [specs[1], specs[0], specs[3], specs[2]]
    check_vs_oracle(dec.deconvolute_spectra(order), order, what=f"axis views {n_long}, long first")


# ------------------------------------------------------------------------------ per-deconvoluter mode
def test_two_deconvoluters_with_different_modes_run_concurrently():
    """The arithmetic of the MSE superposition is a property of the deconvoluter: one pinned to exact and
    one pinned to fast, called concurrently from two threads, each keep their own results -- whatever the
    process default is set to meanwhile."""
    n = 16384
    x = synth.axis(n)
    specs = [Spectrum(x, synth.config3(700 + s, n=n, x=x), SB) for s in range(12)]
    exact, fast = Deconvoluter(), Deconvoluter()
    exact.set_superposition_mode("exact")
    fast.set_superposition_mode("fast")
    assert exact.superposition_mode() == "exact" and fast.superposition_mode() == "fast"
    want_exact = exact.deconvolute_spectra(specs)
    want_fast = fast.deconvolute_spectra(specs)
    check_vs_oracle(want_exact, specs, exact_mse=True, what="pinned exact")
    check_vs_oracle(want_fast, specs, exact_mse=False, what="pinned fast")
    assert any(a.mse != b.mse for a, b in zip(want_exact, want_fast)), "the two modes should differ in the last bits somewhere"
    got, errors = {}, []

    def work(name, dec):
        try:
            for k in range(6):
                set_superposition_mode("exact" if (k & 1) else "fast")  # the process default flips under them
                got[name] = dec.deconvolute_spectra(specs)
        except Exception as err:  # noqa: BLE001
            errors.append(err)

    threads = [threading.Thread(target=work, args=("exact", exact)), threading.Thread(target=work, args=("fast", fast))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for a, b in zip(want_exact, got["exact"]):
        assert a.mse == b.mse and np.array_equal(bits(a.parameters), bits(b.parameters))
    for a, b in zip(want_fast, got["fast"]):
        assert a.mse == b.mse and np.array_equal(bits(a.parameters), bits(b.parameters))
    # an unpinned deconvoluter follows the process default at call time; clones inherit a pin
    free = Deconvoluter()
    set_superposition_mode("exact")
    assert free.superposition_mode() == "exact"
    assert [o.mse for o in free.deconvolute_spectra(specs[:3])] == [o.mse for o in want_exact[:3]]
    set_superposition_mode("fast")
    assert [o.mse for o in free.deconvolute_spectra(specs[:3])] == [o.mse for o in want_fast[:3]]
    lib = _lib.load()
    clone = C.c_void_p()
    assert lib.mdb_deconvoluter_clone(exact._h, C.byref(clone)) == 0
    assert lib.mdb_deconvoluter_superposition_mode(clone) == _lib.MDB_SUPERPOSITION_EXACT
    lib.mdb_deconvoluter_free(clone)
    assert lib.mdb_deconvoluter_set_superposition_mode(exact._h, 5) != 0


def test_single_point_helpers_are_exact_in_every_mode():
    """Lorentzian.evaluate_vec and Lorentzian.superposition always use the exact kernel: they agree bit
    for bit with the scalar evaluate() / the oracle, also while the process default is fast."""
    rng = np.random.default_rng(11)
    lor = [Lorentzian(float(sf), float(hw), float(m)) for sf, hw, m in
           zip(np.exp(rng.uniform(0, 9, 40)), np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), 40)), rng.uniform(0, 10, 40))]
    xs = rng.uniform(-1, 11, 257)
    params = np.array([[l.sfhw, l.hw2, l.maxp] for l in lor])
    for l in lor[:5]:
        assert_same_bits(l.evaluate_vec(xs), [l.evaluate(float(v)) for v in xs], "evaluate_vec vs evaluate")
    for v in xs[:16]:
        assert Lorentzian.superposition(float(v), lor) == O.superposition(float(v), params)
    exact = superposition_vec_array(xs, params, mode="exact")
    fast = superposition_vec_array(xs, params, mode="fast")
    assert_same_bits(exact, O.superposition_vec(xs, params), "explicit exact mode")
    assert np.max(np.abs(fast - exact) / np.abs(exact)) <= 1e-13
    assert _lib.load().mdb_superposition_vec_mode(xs.ctypes.data, xs.size, params.ctypes.data, 40, fast.ctypes.data, 0, 3) != 0


# ------------------------------------------------------------------------------ inputs: pageable, pinned, device
def test_pageable_pinned_and_device_rows_give_identical_results(monkeypatch):
    """The same batch as pageable NumPy rows (gathered by the staging threads one chunk ahead), as
    page-locked rows (direct DMA) and resident in device memory: identical bits, several chunks deep,
    with one and with several staging threads, rows of different lengths included."""
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    n = 32768
    x = synth.axis(n)
    rows = [synth.config3(500 + s, n=n, x=x) for s in range(40)]
    lens = [n if s % 7 else n - 4096 - 8 * s for s in range(40)]  # a few shorter rows (own axis views)
    dec = Deconvoluter()
    dec.set_superposition_mode("exact")
    monkeypatch.setenv("MDB_CHUNK_SPECTRA", "6")  # 7 chunks: the look-ahead ring wraps
    specs = [Spectrum(x[:m], y[:m], (float(x[m - 1]) + 1e-9, 11.8)) for y, m in zip(rows, lens)]
    base = dec.deconvolute_spectra(specs)
    check_vs_oracle(base[:6], specs[:6], what="pageable rows")
    for threads in ("1", "3"):
        monkeypatch.setenv("MDB_STAGE_THREADS", threads)
        again = dec.deconvolute_spectra(specs)
        for a, b in zip(base, again):
            assert np.array_equal(a.peaks, b.peaks) and a.mse == b.mse
            assert_same_bits(a.parameters, b.parameters, f"staging threads = {threads}")
    # pinned host rows and device rows through the raw C ABI
    pinned = [torch.from_numpy(y[:m].copy()).pin_memory() for y, m in zip(rows, lens)]
    xp = torch.from_numpy(x.copy()).pin_memory()
    devs = [p.cuda() for p in pinned]
    xd = xp.cuda()
    for memory, xr, yr in ((_lib.MDB_MEM_HOST, xp, pinned), (_lib.MDB_MEM_DEVICE, xd, devs)):
        views = (_lib.SpectrumView * 40)()
        for s in range(40):
            views[s].chemical_shifts = xr.data_ptr()
            views[s].intensities = yr[s].data_ptr()
            views[s].len = lens[s]
            views[s].signal_boundaries[0], views[s].signal_boundaries[1] = specs[s].signal_boundaries
        batch = C.c_void_p()
        assert lib.mdb_deconvolute_spectra(dec._h, views, 40, memory, C.byref(batch)) == 0, _lib.last_error()
        for s in range(40):
            k = lib.mdb_batch_n_lorentzians(batch, s)
            got = np.ctypeslib.as_array(C.cast(lib.mdb_batch_lorentzians(batch, s), C.POINTER(C.c_double)), (max(k, 1), 3))[:k]
            assert_same_bits(got, base[s].parameters, f"memory kind {memory}, spectrum {s}")
            assert lib.mdb_batch_mse(batch, s) == base[s].mse
        lib.mdb_batch_free(batch)


# ------------------------------------------------------------------------------ superposition_vec: streaming + sharding
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_superposition_host_chunks_and_device_slices_are_bit_identical_to_one_shot(mode, monkeypatch):
    """Host-memory grids stream through the GPU in chunks over three streams and are cut into one slice
    per device under mdb_set_device_count: both must reproduce the one-shot device-memory evaluation bit
    for bit (grid points are independent; lorentzian.rs:656-663)."""
    torch = pytest.importorskip("torch")
    import metabodecon_rust_b200 as M
    lib = _lib.load()
    rng = np.random.default_rng(3)
    p, n = 1500, (1 << 20) + 12345
    hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
    lor = np.ascontiguousarray(np.stack([np.exp(rng.uniform(0, 9, p)) * hw, hw * hw, rng.uniform(0, 10, p)], axis=1))
    x = np.linspace(-2.2, 11.8, n)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lor).cuda()
    od = torch.empty_like(xd)
    code = _lib.MDB_SUPERPOSITION_EXACT if mode == "exact" else _lib.MDB_SUPERPOSITION_FAST
    assert lib.mdb_superposition_vec_mode(xd.data_ptr(), n, ld.data_ptr(), p, od.data_ptr(), _lib.MDB_MEM_DEVICE, code) == 0
    one_shot = od.cpu().numpy()
    assert_same_bits(superposition_vec_array(x, lor, mode=mode), one_shot, "host memory, default chunks")
    for chunk in ("4096", "300000", "99999999"):
        monkeypatch.setenv("MDB_SUP_CHUNK", chunk)
        assert_same_bits(superposition_vec_array(x, lor, mode=mode), one_shot, f"host memory, chunk {chunk}")
    monkeypatch.delenv("MDB_SUP_CHUNK")
    M.set_devices(0)  # every visible device (one slice each; a one-GPU box degrades to one slice)
    try:
        assert_same_bits(superposition_vec_array(x, lor, mode=mode), one_shot, "device slices")
        tiny = superposition_vec_array(x[:100], lor, mode=mode)  # too small to shard: one device
        assert_same_bits(tiny, one_shot[:100], "tiny grid under set_devices(0)")
    finally:
        M.set_devices(1)
    if mode == "exact":
        idx = rng.integers(0, n, 2048)
        assert_same_bits(one_shot[idx], O.superposition_vec(x[idx], lor, parallel=True), "oracle sample")


# ------------------------------------------------------------------------------ SURVEY 8d variant B at full size
@pytest.mark.parametrize("workload", ["config3", "config5"])
def test_full_size_integer_valued_spectra_vs_oracle(workload):
    """2^17-point synthetic spectra ROUNDED TO INTEGERS (Bruker 1r data is int32, bruker.rs:459-475):
    exact ties in the second difference are common there, so peak sets hinge on the smoothing
    recurrence's rounding.  Six spectra per workload against the oracle, bit patterns."""
    n = 131072
    x = synth.axis(n)
    kw = dict(k=500, hw_range=(5e-4, 3e-3)) if workload == "config3" else dict(k=3000, hw_range=(3e-4, 1.5e-3))
    O.use_all_cores()

    def make(seed, n, integer, x):
        return fast_spectrum(seed, n=n, integer=integer, x=x, **kw)

    specs = [Spectrum(x, make(2100 + s, n=n, integer=True, x=x), SB) for s in range(6)]
    assert all(np.array_equal(sp.intensities, np.rint(sp.intensities)) for sp in specs)
    dec = Deconvoluter()
    dec.set_superposition_mode("exact")
    outs = dec.deconvolute_spectra(specs)
    check_vs_oracle(outs, specs, what=f"{workload} integer-valued")
    # the float-valued twins differ from them (the rounding matters) and are exact too
    twins = [Spectrum(x, make(2100 + s, n=n, integer=False, x=x), SB) for s in range(2)]
    check_vs_oracle(dec.deconvolute_spectra(twins), twins, what=f"{workload} float-valued")
    # ties really occur on the integer data: the smoothed curve has exactly equal neighbours in d2
    sm = O.smooth_values(specs[0].intensities, 3, 3)
    d2 = O.second_derivative(sm)
    assert int(np.sum(d2[1:] == d2[:-1])) > 0


# ------------------------------------------------------------------------------ refinement arithmetic experiments
def test_fit_arithmetic_default_is_exact_and_variants_are_opt_in():
    """MDB_FIT_EXACT is the product; MDB_FIT_CORRECTED (one Newton step fewer, Markstein correction kept)
    must reproduce the exact bits on ordinary data (it can differ only when a quotient lies within
    ~2^-103 of a rounding boundary); MDB_FIT_ULP is measurably different (tools/fit_arithmetic.py records
    by how much) -- here only that it stays a valid opt-in with the same peak sets."""
    lib = _lib.load()
    n = 32768
    x = synth.axis(n)
    O.use_all_cores()
    specs = [Spectrum(x, fast_spectrum(800 + s, n=n, k=1200, hw_range=(3e-4, 1.5e-3), x=x), SB) for s in range(64)]  # enough CTAs for the per-peak kernel
    dec = Deconvoluter()
    dec.set_superposition_mode("exact")
    assert lib.mdb_deconvoluter_fit_arithmetic(dec._h) == _lib.MDB_FIT_EXACT
    base = dec.deconvolute_spectra(specs)
    check_vs_oracle(base[:3], specs[:3], what="exact fit")
    assert lib.mdb_deconvoluter_set_fit_arithmetic(dec._h, _lib.MDB_FIT_CORRECTED) == 0
    corrected = dec.deconvolute_spectra(specs)
    for a, b in zip(base, corrected):
        assert_same_bits(b.parameters, a.parameters, "corrected division")
        assert a.mse == b.mse
    assert lib.mdb_deconvoluter_set_fit_arithmetic(dec._h, _lib.MDB_FIT_ULP) == 0
    ulp = dec.deconvolute_spectra(specs)
    for a, b in zip(base, ulp):
        assert np.array_equal(a.peaks, b.peaks)  # selection does not depend on the fit
    assert lib.mdb_deconvoluter_set_fit_arithmetic(dec._h, 7) != 0
    assert lib.mdb_deconvoluter_set_fit_arithmetic(dec._h, _lib.MDB_FIT_EXACT) == 0
    again = dec.deconvolute_spectra(specs[:4])
    for a, b in zip(base, again):
        assert_same_bits(b.parameters, a.parameters, "back to exact")


# ------------------------------------------------------------------------------ limits
def test_unsupported_smoothing_sizes_and_optimize_with_a_large_window(golden_dir):
    import os
    lib = _lib.load()
    n = 8192
    x = synth.axis(n)
    sp = Spectrum(x, synth.config3(3, n=n, x=x), SB)
    dec = Deconvoluter()
    dec.set_moving_average_smoother(2 ** 31, 3)
    with pytest.raises(exceptions.UnexpectedError):
        dec.deconvolute_spectrum(sp)
    assert "2^31" in _lib.last_error()
    dec.set_moving_average_smoother(3, 2 ** 40)
    with pytest.raises(exceptions.UnexpectedError):
        dec.deconvolute_spectrum(sp)
    # optimize_settings replaces the smoothing settings per variant: a deconvoluter that currently holds a
    # window far longer than the spectrum must still optimise (the reference calls set_smoothing_settings
    # before every run, deconvoluter.rs:787-803)
    sim = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    a, b = Deconvoluter(), Deconvoluter()
    b.set_moving_average_smoother(3, 100001)
    assert a.optimize_settings(sim) == b.optimize_settings(sim)
    assert a.smoothing_settings() == b.smoothing_settings()
    assert lib.mdb_kernel_launch_count() > 0


def test_measured_fp64_rate_is_plausible():
    lib = _lib.load()
    dfma, dadd = C.c_double(), C.c_double()
    assert lib.mdb_measure_fp64_rate(C.byref(dfma), C.byref(dadd)) == 0, _lib.last_error()
    # a B200 issues 148 x 64 FP64 instructions per clock at up to 1.965 GHz = 1.86e13/s
    assert 0.5e13 < dfma.value < 2.0e13 and 0.5e13 < dadd.value < 2.0e13


@pytest.mark.parametrize("k_lor,n,count,fit_iters", [(60, 16384, 200, 10), (500, 131072, 150, 10), (900, 131072, 150, 3)])
def test_all_refinement_passes_in_one_launch_match_one_launch_per_pass(k_lor, n, count, fit_iters, monkeypatch):
    """Experiment kept under test (MDB_FIT_BLOCK=1; measured slower in the pipeline): chunks
    whose spectra have at most 1 024 selected peaks run K6 as ONE launch, one CTA per spectrum, the parameter set
    in shared memory (fit_block_kernel; launch-bound classes of 320, 576 and 1 024 threads); the product keeps
    one launch per pass (fit_iter_kernel).  Same bits, and the oracle's
    (fitter_analytical.rs:39-69), including spectra with different peak counts in one chunk and an empty one."""
    x = synth.axis(n)
    base = [fast_spectrum(7000 + 31 * k_lor + s, n=n, k=k_lor + 7 * s, x=x, integer=bool(s & 1)) for s in range(6)]
    flat = np.full(n, 1000.0)  # no peaks at all: its CTA has nothing to do
    ys = [base[s % 6] for s in range(count - 1)] + [flat]
    specs = [Spectrum(x, y, SB) for y in ys]
    dec = Deconvoluter()
    dec.set_analytical_fitter(fit_iters)
    dec.set_superposition_mode("exact")
    lib = _lib.load()
    monkeypatch.setenv("MDB_FIT_BLOCK", "1")
    lib.mdb_reset_kernel_launch_count()
    block = dec.deconvolute_spectra(specs[:-1])
    launches_block = lib.mdb_kernel_launch_count()
    monkeypatch.delenv("MDB_FIT_BLOCK")
    lib.mdb_reset_kernel_launch_count()
    per_pass = dec.deconvolute_spectra(specs[:-1])
    launches_per_pass = lib.mdb_kernel_launch_count()
    assert launches_block < launches_per_pass  # the passes really went into one launch per chunk
    for i, (a, b) in enumerate(zip(block, per_pass)):
        assert np.array_equal(a.peaks, b.peaks), i
        assert_same_bits(a.parameters, b.parameters, f"spectrum {i}")
        assert a.mse == b.mse
    settings = O.Settings(fitting_iterations=fit_iters)
    check_vs_oracle(block[:6], specs[:6], settings, what=f"fit_block k={k_lor}")
    assert max(len(o.peaks) for o in block) <= 1024
    # the spectrum without peaks is an error in the reference (EmptySignalRegion / no peaks); it must not disturb the others
    with pytest.raises(Exception):
        dec.deconvolute_spectra(specs)
