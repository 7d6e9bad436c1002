"""GPU parity: every stage of the CUDA path against the CPU oracle, through the C ABI.

Bar (BASELINE.json north_star): peak indices and selected-peak sets bit-exact; Lorentzian
parameters and superposition values within 1e-9 relative.  The kernels replay the reference's
operation order, so these tests assert the stronger property -- identical bit patterns -- and
state the 1e-9 tolerance only where a weaker check is all the reference itself could promise.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
import synth
from metabodecon_rust_b200 import Deconvoluter, Lorentzian, Spectrum, _lib, exceptions

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9  # north_star tolerance for f64 parameters / superposition values


@pytest.fixture(autouse=True)
def exact_superposition(monkeypatch):
    """This file asserts identical bit patterns everywhere, so it runs the MSE superposition and
    superposition_vec in MDB_SUPERPOSITION_EXACT.  The library's default (the few-ulp form of those
    two kernels; peak sets and Lorentzians are bit-identical in both) is covered against the same
    oracle, with the tolerance written out, in tests/test_gpu_fast_superposition.py."""
    lib = _lib.load()
    before = lib.mdb_superposition_mode()
    monkeypatch.setenv("MDB_SUPERPOSITION", "exact")  # host programs started by a test
    assert lib.mdb_set_superposition_mode(0) == 0
    yield
    assert lib.mdb_set_superposition_mode(before) == 0


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(got, want, what=""):
    got = np.ascontiguousarray(got, dtype=np.float64)
    want = np.ascontiguousarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    neq = bits(got) != bits(want)
    if neq.any():
        i = np.flatnonzero(neq.reshape(-1))[0]
        rel = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
        raise AssertionError(f"{what}: {neq.sum()} of {neq.size} values differ in bits; first at {i}: "
                             f"{got.reshape(-1)[i]!r} vs {want.reshape(-1)[i]!r}; max rel err {rel:.3e} "
                             f"(north-star tolerance {REL_TOL})")


# ------------------------------------------------------------------------------ stage wrappers
def gpu_smooth(y, iterations, window):
    lib = _lib.load()
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(y)
    st = lib.mdb_stage_smooth(y.ctypes.data, y.size, iterations, window, out.ctypes.data)
    assert st == 0, _lib.last_error()
    return out


def gpu_detect(smoothed):
    lib = _lib.load()
    sm = np.ascontiguousarray(smoothed, dtype=np.float64)
    cap = sm.size // 2 + 2
    peaks = np.zeros((cap, 3), dtype=np.int32)
    scores = np.zeros(cap, dtype=np.float64)
    n = C.c_size_t()
    st = lib.mdb_stage_detect(sm.ctypes.data, sm.size, peaks.ctypes.data, scores.ctypes.data, cap, C.byref(n))
    assert st == 0, _lib.last_error()
    return peaks[:n.value], scores[:n.value]


def gpu_select(dec: Deconvoluter, smoothed, sb_idx, ignore_idx=None):
    lib = _lib.load()
    sm = np.ascontiguousarray(smoothed, dtype=np.float64)
    cap = sm.size // 2 + 2
    peaks = np.zeros((cap, 3), dtype=np.int32)
    n = C.c_size_t()
    msd = (C.c_double * 2)()
    has = ignore_idx is not None
    ig = np.ascontiguousarray(np.asarray(ignore_idx if has else [], dtype=np.uintp).reshape(-1))
    st = lib.mdb_stage_select(dec._h, sm.ctypes.data, sm.size, sb_idx[0], sb_idx[1], int(has),
                              ig.ctypes.data if ig.size else None, ig.size // 2, peaks.ctypes.data, cap,
                              C.byref(n), msd)
    return st, peaks[:n.value], (msd[0], msd[1])


def gpu_fit(x, y, peaks, iterations, trace=True):
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    pk = np.ascontiguousarray(peaks, dtype=np.int32).reshape(-1, 3)
    p = pk.shape[0]
    out = np.zeros((max(p, 1), 3), dtype=np.float64)
    tr = np.zeros((iterations + 1, max(p, 1), 3), dtype=np.float64)
    kept = C.c_size_t()
    st = lib.mdb_stage_fit(x.ctypes.data, y.ctypes.data, x.size, pk.ctypes.data, p, iterations,
                           out.ctypes.data, C.byref(kept), tr.ctypes.data if trace else None)
    assert st == 0, _lib.last_error()
    return out[:kept.value], tr[:, :p]


# ------------------------------------------------------------------------------ superposition (K8)
def test_superposition_reference_kats():
    # lorentzian.rs:741-788: 11 points x 3 Lorentzians with closed-form expectations
    lor = [Lorentzian.from_transformed(1.0, 0.5, -2.0), Lorentzian.from_transformed(2.0, 0.75, 0.0),
           Lorentzian.from_transformed(1.0, 0.5, 2.0)]
    x = np.array([-5.0 + i for i in range(11)])
    want = np.array([
        1.0 / 9.5 + 2.0 / 25.75 + 1.0 / 49.5, 1.0 / 4.5 + 2.0 / 16.75 + 1.0 / 36.5,
        1.0 / 1.5 + 2.0 / 9.75 + 1.0 / 25.5, 1.0 / 0.5 + 2.0 / 4.75 + 1.0 / 16.5,
        1.0 / 1.5 + 2.0 / 1.75 + 1.0 / 9.5, 1.0 / 4.5 + 2.0 / 0.75 + 1.0 / 4.5,
        1.0 / 9.5 + 2.0 / 1.75 + 1.0 / 1.5, 1.0 / 16.5 + 2.0 / 4.75 + 1.0 / 0.5,
        1.0 / 25.5 + 2.0 / 9.75 + 1.0 / 1.5, 1.0 / 36.5 + 2.0 / 16.75 + 1.0 / 4.5,
        1.0 / 49.5 + 2.0 / 25.75 + 1.0 / 9.5])
    for fn in (Lorentzian.superposition_vec, Lorentzian.par_superposition_vec):
        got = fn(x, lor)
        np.testing.assert_allclose(got, want, rtol=4e-16)  # float_cmp approx_eq (ULP-level) in the reference
    for xi, wi in zip(x, want):
        assert abs(Lorentzian.superposition(float(xi), lor) - wi) <= 4e-16 * abs(wi)
    # lorentzian.rs:593-604 doc-test
    trip = [Lorentzian.from_transformed(0.03, 0.0009, 4.8), Lorentzian.from_transformed(0.02, 0.0004, 5.0),
            Lorentzian.from_transformed(0.03, 0.0009, 5.2)]
    assert abs(Lorentzian.superposition(5.0, trip) - 51.466992) < 1e-6
    # lorentzian.rs:707-739 evaluate table
    one = Lorentzian.from_transformed(1.0, 1.0, 0.0)
    ev = one.evaluate_vec(x)
    np.testing.assert_allclose(ev, 1.0 / (1.0 + x * x), rtol=4e-16)


@pytest.mark.parametrize("n,p", [(1, 1), (7, 3), (1023, 1), (1025, 1025), (10007, 777), (4096, 2049), (100000, 64)])
def test_superposition_bit_exact_vs_oracle(n, p):
    rng = np.random.default_rng(n * 131 + p)
    x = np.sort(rng.uniform(-2.2, 11.8, n))
    hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
    sf = np.exp(rng.uniform(0, np.log(1e4), p))
    lor = np.stack([sf * hw, hw * hw, rng.uniform(0, 10, p)], axis=1)
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    got = superposition_vec_array(x, lor)
    want = O.superposition_vec(x, lor)
    assert_bit_equal(got, want, f"superposition_vec n={n} p={p}")


def test_superposition_empty_and_special_values():
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    x = np.linspace(0, 1, 17)
    got = superposition_vec_array(x, np.zeros((0, 3)))
    assert np.all(got == 0.0)
    assert superposition_vec_array(np.zeros(0), np.ones((2, 3))).size == 0
    # denominators of 0, inf and NaN parameters take the division slow path; compare to the oracle
    lor = np.array([[1.0, 0.0, 0.5], [np.inf, 1.0, 0.2], [1.0, np.nan, 0.1], [1e-320, 1e300, 0.3], [1e300, 1e-300, 0.7]])
    for k in range(len(lor)):
        assert_bit_equal(superposition_vec_array(x, lor[k:k + 1]), O.superposition_vec(x, lor[k:k + 1]), f"special {k}")


def _random_lorentzians(rng, p):
    hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
    sf = np.exp(rng.uniform(0, np.log(1e4), p))
    return np.stack([sf * hw, hw * hw, rng.uniform(0, 10, p)], axis=1)


@pytest.mark.parametrize("n,p", [(303104, 33), (303105, 513), (400001, 1025), (131072, 1537), (1 << 20, 511)])
def test_superposition_throughput_shape_and_tile_edges(n, p):
    """n >= 2*148*1024 selects the 8-points-per-thread kernel; p around multiples of the 512-entry
    TMA tile (odd p leaves a trailing 8 bytes that is copied by hand)."""
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    rng = np.random.default_rng(n + p)
    x = rng.uniform(-2.2, 11.8, n)
    lor = _random_lorentzians(rng, p)
    assert_bit_equal(superposition_vec_array(x, lor), O.superposition_vec(x, lor, parallel=True), f"n={n} p={p}")


@pytest.mark.parametrize("bad", [
    [0.0, 1e-6, 5.0],            # sfhw == 0: quotient 0, outside the fast domain
    [-3.0, 2e-7, 4.0],           # negative sfhw stays inside it
    [1e-320, 1e-6, 5.0],         # denormal numerator
    [1.0, 0.0, 5.0],             # hw2 == 0 (division by zero at x == maxp)
    [1.0, -1e-6, 5.0],           # negative hw2
    [1.0, 1e-310, 5.0],          # denormal hw2
    [1e305, 1e-6, 5.0],          # huge numerator
    [1.0, 1e-6, 1e200],          # (x - maxp)^2 overflows
    [np.inf, 1e-6, 5.0], [1.0, np.inf, 5.0], [1.0, 1e-6, np.nan], [np.nan, 1e-6, 5.0],
])
def test_division_fast_and_ieee_paths_agree_with_oracle(bad):
    """A tile whose parameters leave the fast domain must run the IEEE division loop; tiles before
    and after it keep the fast path.  Either way the result is the oracle's, bit for bit."""
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    rng = np.random.default_rng(7)
    lor = _random_lorentzians(rng, 1300)  # three tiles
    x = np.concatenate([rng.uniform(-2.2, 11.8, 4000), [5.0, 4.0, 0.0]])
    for pos in (0, 700, 1299):
        cur = lor.copy()
        cur[pos] = bad
        got = superposition_vec_array(x, cur)
        want = O.superposition_vec(x, cur)
        both_nan = np.isnan(got) & np.isnan(want)  # NaN payloads are not part of the contract
        assert_bit_equal(np.where(both_nan, 0.0, got), np.where(both_nan, 0.0, want), f"bad={bad} at {pos}")


def test_division_fast_path_out_of_domain_x():
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    rng = np.random.default_rng(8)
    lor = _random_lorentzians(rng, 600)
    x = np.array([1e300, -1e155, 1e-320, 0.0, np.inf, np.nan, 3.3] + list(rng.uniform(0, 10, 300)))
    got = superposition_vec_array(x, lor)
    want = O.superposition_vec(x, lor)
    both_nan = np.isnan(got) & np.isnan(want)
    assert_bit_equal(np.where(both_nan, 0.0, got), np.where(both_nan, 0.0, want), "out-of-domain x")


def test_superposition_device_memory_and_unaligned_parameters():
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    rng = np.random.default_rng(9)
    n, p = 5000, 777
    x = rng.uniform(-2.2, 11.8, n)
    lor = _random_lorentzians(rng, p)
    want = O.superposition_vec(x, lor)
    dx = torch.from_numpy(x).cuda()
    flat = torch.zeros(3 * p + 1, dtype=torch.float64, device="cuda")
    for shift in (0, 1):  # shift 1: parameter block only 8-byte aligned
        flat[shift:shift + 3 * p] = torch.from_numpy(lor.reshape(-1)).cuda()
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        st = lib.mdb_superposition_vec(dx.data_ptr(), n, flat.data_ptr() + 8 * shift, p, out.data_ptr(), _lib.MDB_MEM_DEVICE)
        assert st == 0, _lib.last_error()
        torch.cuda.synchronize()
        assert_bit_equal(out.cpu().numpy(), want, f"device memory, shift {shift}")


# ------------------------------------------------------------------------------ smoothing (K1)
@pytest.mark.parametrize("iterations,window", [(3, 3), (1, 3), (2, 5), (10, 7), (3, 4), (2, 2), (4, 9), (1, 65)])
def test_smoothing_bit_exact_blood(blood_arrays, iterations, window):
    _, y = blood_arrays
    y = y[:40000]
    assert_bit_equal(gpu_smooth(y, iterations, window), O.smooth_values(y, iterations, window),
                     f"smooth ({iterations},{window})")


def test_smoothing_bit_exact_full_blood_and_synthetic(blood_arrays):
    _, y = blood_arrays
    assert_bit_equal(gpu_smooth(y, 3, 3), O.smooth_values(y, 3, 3), "blood_01 (3,3)")
    ys = synth.config3(1, n=32768)
    assert_bit_equal(gpu_smooth(ys, 3, 3), O.smooth_values(ys, 3, 3), "synthetic f64 (3,3)")
    yi = synth.config3(1, n=32768, integer=True)
    assert_bit_equal(gpu_smooth(yi, 3, 3), O.smooth_values(yi, 3, 3), "synthetic int (3,3)")


@pytest.mark.parametrize("n,iterations,window", [(960, 3, 3), (9600, 3, 3), (96 * 7 + 1, 3, 3), (96 * 7 - 1, 2, 3),
                                                 (800, 2, 5), (80 * 11, 9, 5), (840, 10, 7), (84 * 5, 4, 7),
                                                 (192, 3, 3), (97, 3, 3), (288, 10, 3)])
def test_smoothing_tile_boundaries(n, iterations, window):
    # lengths that are exact multiples of the kernel's tile (96 / 80 / 84 points) or one off
    rng = np.random.default_rng(n + iterations)
    y = np.rint(rng.normal(0, 1e5, n))
    assert_bit_equal(gpu_smooth(y, iterations, window), O.smooth_values(y, iterations, window),
                     f"n={n} ({iterations},{window})")
    y = rng.normal(0, 1e5, n)
    assert_bit_equal(gpu_smooth(y, iterations, window), O.smooth_values(y, iterations, window),
                     f"n={n} ({iterations},{window}) f64")


@pytest.mark.parametrize("n", [4096, 4097, 6143, 8191, 8192, 8193, 10000, 2048 * 9 + 5, 50001])
def test_smoothing_stream_kernel_ring_edges(n, monkeypatch):
    """The latency form of K1 (smooth_stream.cuh: rings of 2 048 points in shared memory, taken by
    calls of a few long spectra): lengths around multiples of the ring, every window parity, up to
    12 passes; identical bits from the lane-packed kernel (MDB_SMOOTH_STREAM=0) and the oracle."""
    rng = np.random.default_rng(n)
    y = rng.normal(0, 1e4, n)
    yi = np.rint(y)
    # windows 3, 5 and 7 take the interior loop specialised on the window (register-resident blocks, stores
    # one round late); MDB_STREAM_GENERIC=1 sends them through the any-window loop as well
    # and with at most six passes they run one chain warp per pass with the multiply taken off the chain
    # (smooth_split.cuh) unless MDB_SMOOTH_SPLIT=0
    for it, w in [(3, 3), (1, 2), (2, 2), (12, 3), (5, 8), (4, 9), (2, 33), (3, 64), (1, 7), (2, 5), (12, 5), (7, 7), (1, 5),
                  (6, 5), (4, 7), (6, 3)]:
        want = O.smooth_values(y, it, w)
        assert_bit_equal(gpu_smooth(y, it, w), want, f"stream n={n} ({it},{w})")
        assert_bit_equal(gpu_smooth(yi, it, w), O.smooth_values(yi, it, w), f"stream n={n} ({it},{w}) integer")
        if w in (3, 5, 7):
            monkeypatch.setenv("MDB_SMOOTH_SPLIT", "0")
            assert_bit_equal(gpu_smooth(y, it, w), want, f"stream, one warp for all passes n={n} ({it},{w})")
            monkeypatch.delenv("MDB_SMOOTH_SPLIT")
            monkeypatch.setenv("MDB_STREAM_GENERIC", "1")
            assert_bit_equal(gpu_smooth(y, it, w), want, f"stream, any-window loop n={n} ({it},{w})")
            monkeypatch.delenv("MDB_STREAM_GENERIC")
        monkeypatch.setenv("MDB_SMOOTH_STREAM", "0")
        assert_bit_equal(gpu_smooth(y, it, w), want, f"lanes n={n} ({it},{w})")
        monkeypatch.delenv("MDB_SMOOTH_STREAM")


@pytest.mark.parametrize("n", [5, 6, 7, 8, 17, 63, 64, 65])
def test_smoothing_short_inputs(n):
    rng = np.random.default_rng(n)
    y = rng.normal(0, 1000, n)
    for it, w in [(3, 3), (2, 5), (3, 7), (1, 4)]:
        assert_bit_equal(gpu_smooth(y, it, w), O.smooth_values(y, it, w), f"n={n} ({it},{w})")


def test_smoothing_batch_entry_small_and_large_tile_variants(blood_arrays):
    """mdb_stage_smooth_batch (device memory, one launch): launches of more warps than SMs take the
    small-tile kernel, smaller ones the 224-point tile; both must reproduce the oracle."""
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    _, y = blood_arrays
    n = 20000
    rows = np.stack([np.roll(y, 997 * s)[:n] for s in range(7)])
    want = np.stack([O.smooth_values(r, 3, 3) for r in rows])
    for count in (7, 2000, 4500):  # 1 warp (224-point tiles), 200 warps (112), 450 warps (56)
        src = torch.from_numpy(rows).cuda().repeat((count + 6) // 7, 1)[:count].contiguous()
        dst = torch.zeros_like(src)
        ms = C.c_double()
        st = lib.mdb_stage_smooth_batch(src.data_ptr(), n, count, n, 3, 3, dst.data_ptr(), C.byref(ms))
        assert st == 0, _lib.last_error()
        got = dst.cpu().numpy()
        for s in range(count):
            assert_bit_equal(got[s], want[s % 7], f"batch smoothing count={count} row {s}")
        assert ms.value > 0.0


# ------------------------------------------------------------------------------ detection (K2/K3)
def _check_detect(sm, what):
    pk, sc = gpu_detect(sm)
    d2 = O.second_derivative(sm)
    want = O.detect_peaks(d2)
    assert pk.shape[0] == want.shape[0], f"{what}: {pk.shape[0]} triplets vs {want.shape[0]}"
    assert np.array_equal(pk.astype(np.int64), want.astype(np.int64)), f"{what}: triplets differ"
    a = np.abs(d2)
    want_sc = np.array([O.score_peak(a, int(l), int(c), int(r)) for l, c, r in want])
    assert_bit_equal(sc, want_sc, f"{what}: scores")
    return pk


def test_detection_reference_kats():
    # common.rs:50-58 / detector.rs:172-208, through the full detector on tiny inputs.
    # [1,2,3,2,1] -> d2 = [0,-2,0]; centre 2; no borders inside -> NoPeaksDetected-style empty list
    pk, _ = gpu_detect(np.array([1.0, 2.0, 3.0, 2.0, 1.0]))
    assert pk.shape[0] == 0
    # a bump wide enough to have both borders
    y = np.array([0, 0, 1, 3, 7, 12, 15, 12, 7, 3, 1, 0, 0, 0.0])
    _check_detect(y, "bump")


def test_detection_bit_exact_blood(blood_arrays):
    _, y = blood_arrays
    sm = O.smooth_values(y, 3, 3)
    pk = _check_detect(sm, "blood_01")
    assert pk.shape[0] == 16100


@pytest.mark.parametrize("seed,integer,n", [(0, False, 131072), (1, True, 65536), (2, False, 4099), (3, True, 1000)])
def test_detection_bit_exact_synthetic(seed, integer, n):
    y = synth.config3(seed, n=n, integer=integer)
    _check_detect(O.smooth_values(y, 3, 3), f"synthetic seed={seed}")
    _check_detect(y, f"synthetic unsmoothed seed={seed}")


def test_detection_edge_inputs():
    rng = np.random.default_rng(7)
    for n in (5, 6, 7, 9, 31, 64, 65, 127, 129):
        _check_detect(rng.normal(0, 1, n), f"noise n={n}")
    _check_detect(np.zeros(1000), "zeros")
    _check_detect(np.arange(1000, dtype=np.float64) ** 2, "parabola")
    t = np.linspace(0, 40 * np.pi, 5000)
    _check_detect(np.sin(t) * 1e6, "sine")          # smooth: long monotone d2 runs, far borders
    _check_detect(np.rint(np.sin(t) * 50), "ties")  # exact ties in d2
    y = rng.normal(0, 1, 3000)
    y[100] = np.nan
    _check_detect(y, "nan")


def _signal_from_d2(d2):
    """Integrate a prescribed second difference twice: y[j+2] = d2[j] - y[j] + 2 y[j+1]."""
    y = np.zeros(len(d2) + 2)
    for j, v in enumerate(d2):
        y[j + 2] = v - y[j] + 2.0 * y[j + 1]
    return y


@pytest.mark.parametrize("centre", [200, 450, 511, 512, 513, 575, 1023, 1024, 1500])
@pytest.mark.parametrize("run", [3, 70, 200, 700])
def test_detection_walks_leaving_the_shared_memory_window(centre, run):
    """Borders further away than the kernel's 64-point halo (and across one or more 512-point
    tiles) take the global-memory continuation of the walks and of the left score sum; the result
    must not depend on the tiling.  d2 is prescribed: a long strictly decreasing run into the
    centre and a long strictly increasing run out of it, placed at and around tile edges."""
    n = 2600
    rng = np.random.default_rng(centre * 1000 + run)
    d2 = rng.integers(-3, 4, n - 2).astype(np.float64)           # small integers: exact arithmetic, many ties
    c = centre                                                   # centre index in intensity space: minimum at d2[c-1]
    lo, hi = max(0, c - 1 - run), min(n - 3, c - 1 + run)
    d2[lo:c - 1] = 1000.0 + np.arange(c - 1 - lo, 0, -1)         # ... 1003, 1002, 1001 (decreasing)
    d2[c - 1] = -50.0
    d2[c:hi + 1] = 2000.0 + np.arange(1, hi - c + 2)             # 2001, 2002, ... (increasing)
    y = _signal_from_d2(d2)
    assert np.array_equal(O.second_derivative(y), d2)            # integers stay exact
    pk = _check_detect(y, f"centre={centre} run={run}")
    if lo > 0 and hi < n - 3:  # both runs end inside the spectrum: the planted peak has its borders at the run ends
        hit = pk[pk[:, 1] == c]
        assert hit.shape[0] == 1 and hit[0, 0] == lo + 1 and hit[0, 2] == hi + 1


# ------------------------------------------------------------------------------ selection (K4)
def _check_select(dec, okind, thr, sm, sb_idx, ig, what):
    st, pk, (mean, sd) = gpu_select(dec, sm, sb_idx, ig)
    want = O.select_peaks(sm, okind, thr, sb_idx, ig)
    assert st == want.status or (want.status == O.PANIC and st == 100), f"{what}: status {st} vs {want.status}"
    if want.status != O.OK:
        return
    assert np.array_equal(pk.astype(np.int64), want.peaks.astype(np.int64)), f"{what}: selected peaks differ"
    if okind == O.SELECT_NOISE_SCORE_FILTER:
        assert_bit_equal([mean, sd], [want.mean, want.sd], f"{what}: mean/sd")


def test_selection_blood(blood_arrays):
    x, y = blood_arrays
    sm = O.smooth_values(y, 3, 3)
    sb_idx = O.signal_boundaries_indices(x, (11.8, -2.2))
    ig = O.ignore_region_indices(x, (11.8, -2.2), [(4.7, 4.9)])
    dec = Deconvoluter()
    _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, sb_idx, None, "blood no-ignore")
    _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, sb_idx, ig, "blood water ignored")
    st, pk, (mean, sd) = gpu_select(dec, sm, sb_idx, ig)
    assert pk.shape[0] == 981 and pk[0].tolist() == [41583, 41585, 41588] and pk[-1].tolist() == [97895, 97896, 97898]
    assert float(mean).hex() == "0x1.1ac34b169a537p+8" and float(sd).hex() == "0x1.6a11671179e58p+7"  # Appendix B
    for thr in (0.5, 6.4, 8.0, 1e6):
        dec.set_noise_score_selector(thr)
        _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, thr, sm, sb_idx, ig, f"blood thr={thr}")
    dec.set_detector_only()
    _check_select(dec, O.SELECT_DETECTOR_ONLY, 0.0, sm, sb_idx, None, "blood detector-only")
    _check_select(dec, O.SELECT_DETECTOR_ONLY, 0.0, sm, sb_idx, ig, "blood detector-only ignore")


def test_selection_region_edge_cases(blood_arrays):
    x, y = blood_arrays
    sm = O.smooth_values(y, 3, 3)[:20000]
    dec = Deconvoluter()
    n = sm.size
    cases = [(0, n), (0, 10), (n - 10, n), (5000, 5001), (5000, 5000), (19990, 20000), (3, 7), (100, 19000),
             (19000, 100), (n + 5, n + 9)]
    for sb in cases:
        _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, sb, None, f"sb={sb}")
        _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, sb, [(6000, 9000), (12000, 12001)], f"sb={sb} ig")
    # everything ignored -> the reference panics
    _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, (100, 19000), [(0, n)], "all ignored")
    # no peaks at all
    st, _, _ = gpu_select(dec, np.zeros(500), (10, 400), None)
    assert st == _lib.MDB_ERR_NO_PEAKS_DETECTED


def test_selection_synthetic():
    dec = Deconvoluter()
    for seed, integer in [(0, False), (1, True), (2, False)]:
        y = synth.config3(seed, n=65536, integer=integer)
        x = synth.axis(65536)
        sm = O.smooth_values(y, 3, 3)
        sb_idx = O.signal_boundaries_indices(x, synth.SIGNAL_BOUNDARIES)
        _check_select(dec, O.SELECT_NOISE_SCORE_FILTER, 5.0, sm, sb_idx, None, f"synthetic {seed}")


# ------------------------------------------------------------------------------ fit (K5/K6)
def test_fit_reference_kat():
    # fitter_analytical.rs:188-196: stencil (4,5),(8,10),(12,5) -> maxp 8, hw 4, sfhw/hw 40.
    # One peak, one iteration would rescale; use the trace's initial solve.
    x = np.arange(0.0, 16.0, 1.0)
    y = np.zeros(16)
    y[4], y[8], y[12] = 5.0, 10.0, 5.0
    _, tr = gpu_fit(x, y, [[4, 8, 12]], 1)
    sfhw, hw2, maxp = tr[0, 0]
    assert abs(maxp - 8.0) < 1e-12 and abs(np.sqrt(hw2) - 4.0) < 1e-12 and abs(sfhw / np.sqrt(hw2) - 40.0) < 1e-10
    want, wtr = O.fit_lorentzian(x, y, [[4, 8, 12]], 1, trace=True)
    assert_bit_equal(tr, wtr, "single-stencil trace")


def test_fit_bit_exact_blood_every_iteration(blood_arrays):
    x, y = blood_arrays
    r = O.deconvolute_spectrum(O.Settings(ignore_regions=[(4.7, 4.9)]), x, y, (11.8, -2.2))
    want, wtr = O.fit_lorentzian(x, y, r.peaks, 10, trace=True)
    got, tr = gpu_fit(x, y, r.peaks, 10)
    for it in range(11):
        assert_bit_equal(tr[it], wtr[it], f"blood fit, state after pass {it}")
    assert got.shape[0] == 760  # Appendix B
    assert_bit_equal(got, want, "blood retained lorentzians")
    assert float(got[0, 0]).hex() == "0x1.084fd4b50b8dfp-4" and float(got[0, 2]).hex() == "0x1.0eac0d0cd9679p+3"


@pytest.mark.parametrize("iterations", [1, 2, 5, 15])
def test_fit_bit_exact_synthetic(iterations):
    n = 32768
    x = synth.axis(n)
    y = synth.config3(4, n=n)
    r = O.deconvolute_spectrum(O.Settings(), x, y, synth.SIGNAL_BOUNDARIES)
    assert r.status == O.OK and len(r.peaks) > 100
    want, wtr = O.fit_lorentzian(x, y, r.peaks, iterations, trace=True)
    got, tr = gpu_fit(x, y, r.peaks, iterations)
    assert_bit_equal(tr, wtr, f"synthetic fit trace it={iterations}")
    assert_bit_equal(got, want, "synthetic retained")


def test_fit_wide_and_per_peak_forms_agree(blood_arrays, monkeypatch):
    """A call of a few spectra runs a refinement pass as superposition per (stencil point, peak) +
    solve per peak (fit_wide.cuh); MDB_FIT_WIDE=0 keeps fit_iter_kernel (one thread per peak).  Same
    traces, pass by pass, and both equal to the oracle's -- also with values outside div_fast's domain."""
    x, y = blood_arrays
    r = O.deconvolute_spectrum(O.Settings(ignore_regions=[(4.7, 4.9)]), x, y, (11.8, -2.2))
    want, wtr = O.fit_lorentzian(x, y, r.peaks, 4, trace=True)
    wide, tr_wide = gpu_fit(x, y, r.peaks, 4)          # producers / accumulators (fit_wide2_superpose_kernel)
    monkeypatch.setenv("MDB_FIT_WIDE", "1")            # 8 Lorentzians per thread (fit_wide_superpose_kernel)
    ilp, tr_ilp = gpu_fit(x, y, r.peaks, 4)
    monkeypatch.setenv("MDB_FIT_WIDE", "0")
    narrow, tr_narrow = gpu_fit(x, y, r.peaks, 4)
    monkeypatch.delenv("MDB_FIT_WIDE")
    assert_bit_equal(tr_ilp, wtr, "wide form (instruction-parallel), trace")
    assert_bit_equal(ilp, want, "wide form (instruction-parallel), retained")
    assert_bit_equal(tr_wide, wtr, "wide form, trace")
    assert_bit_equal(tr_narrow, wtr, "per-peak form, trace")
    assert_bit_equal(wide, want, "wide form, retained")
    assert_bit_equal(narrow, want, "per-peak form, retained")
    # a few peaks, one of them on a huge intensity (quotients beyond 2^300: the IEEE division loop)
    n = 4096
    xs = synth.axis(n)
    ys = synth.spectrum(77, n=n, k=30, hw_range=(8e-3, 5e-2), x=xs)
    ys[2000:2003] = [1e200, 3e200, 2e200]
    peaks = np.array([[1999, 2001, 2003], [999, 1001, 1003], [3000, 3002, 3004]], dtype=np.int64)
    want, wtr = O.fit_lorentzian(xs, ys, peaks, 3, trace=True)
    got, tr = gpu_fit(xs, ys, peaks, 3)
    nan = np.isnan(wtr)
    assert_bit_equal(np.where(nan, 0.0, tr), np.where(nan, 0.0, wtr), "wide form, out-of-domain values")
    # peak counts around the 56-Lorentzian tile and the 64-chain CTA of the producer form
    xb, yb = blood_arrays
    for count in (1, 7, 8, 9, 55, 56, 57, 63, 64, 65, 112, 113, 129):
        sub = r.peaks[100:100 + count]
        want_c, wtr_c = O.fit_lorentzian(xb, yb, sub, 2, trace=True)
        got_c, tr_c = gpu_fit(xb, yb, sub, 2)
        assert_bit_equal(tr_c, wtr_c, f"wide form, {count} peaks, trace")
        assert_bit_equal(got_c, want_c, f"wide form, {count} peaks")


def test_fit_many_peaks_tiles():
    # more peaks than one shared-memory tile (LOR_TILE = 512) and a ragged last block
    n = 131072
    x = synth.axis(n)
    y = synth.config5(0, n=n)
    r = O.deconvolute_spectrum(O.Settings(), x, y, synth.SIGNAL_BOUNDARIES)
    assert len(r.peaks) > 2000
    want, wtr = O.fit_lorentzian(x, y, r.peaks, 3, trace=True)
    got, tr = gpu_fit(x, y, r.peaks, 3)
    assert_bit_equal(tr, wtr, "config-5 fit trace")
    assert_bit_equal(got, want, "config-5 retained")


# ------------------------------------------------------------------------------ end to end (a1)
def _check_e2e(dec, osettings, spectra, what):
    outs = dec.deconvolute_spectra(spectra)
    for i, (sp, out) in enumerate(zip(spectra, outs)):
        r = O.deconvolute_spectrum(osettings, sp.chemical_shifts, sp.intensities, sp.signal_boundaries)
        assert r.status == O.OK
        assert np.array_equal(out.peaks.astype(np.int64), r.peaks.astype(np.int64)), f"{what}[{i}]: peak set differs"
        assert_bit_equal(out.parameters, r.lorentzians, f"{what}[{i}]: lorentzians")
        assert_bit_equal([out.mse], [r.mse], f"{what}[{i}]: mse")
    return outs


def test_config1_blood_default_deconvoluter(golden_dir):
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
    dec = Deconvoluter()
    dec.add_ignore_region((4.7, 4.9))
    out = _check_e2e(dec, O.Settings(ignore_regions=[(4.7, 4.9)]), [sp], "blood_01 water ignored")[0]
    assert len(out.lorentzians) == 760 and float(out.mse).hex() == "0x1.0808a64fe177ep+35"  # Appendix B
    dec.clear_ignore_regions()
    out = _check_e2e(dec, O.Settings(), [sp], "blood_01")[0]
    assert out.peaks.shape[0] == 992 and len(out.lorentzians) == 766
    # single-spectrum entry points agree with the batch one
    one = dec.par_deconvolute_spectrum(sp)
    assert_bit_equal(one.parameters, out.parameters, "par_deconvolute_spectrum")


def test_config2_jcampdx_blood_default_deconvoluter(golden_dir):
    """BASELINE config 2: the bundled JCAMP-DX spectrum (XYDATA, DIFDUP) with default settings.
    Its axis differs from the Bruker one at the 1e-4 ppm level (Hz grid + .SHIFT REFERENCE), so the
    result is compared with the oracle run on the same decoded arrays."""
    sp = Spectrum.read_jcampdx(os.path.join(golden_dir, "jcampdx", "blood_01.dx"), (-2.2, 11.8))
    assert len(sp) == 131072
    out = _check_e2e(Deconvoluter(), O.Settings(), [sp], "blood_01.dx")[0]
    assert out.peaks.shape[0] > 900 and len(out.lorentzians) > 700
    # same intensities as config 1; the axes drift apart by up to one grid step over the range, so
    # the signal-boundary indices (and with them the noise threshold) may differ by one point
    bru = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
    ref = Deconvoluter().deconvolute_spectrum(bru)
    assert abs(out.peaks.shape[0] - ref.peaks.shape[0]) <= 0.02 * ref.peaks.shape[0]


def test_sim_spectrum_recovers_generating_parameters(golden_dir):
    # reference integration test `sim` (tests/deconvoluter.rs:7-21): sim_01, signal region 3.35..3.55
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    dec = Deconvoluter()
    out = _check_e2e(dec, O.Settings(), [sp], "sim_01")[0]
    truth = np.loadtxt(os.path.join(golden_dir, "bruker", "sim_01", "lorentzians.csv"), delimiter=",", skiprows=1)
    # approximate known answer: every fitted maxp sits within two grid steps of a generating peak
    step = abs(sp.chemical_shifts[1] - sp.chemical_shifts[0])
    dist = np.min(np.abs(out.parameters[:, 2][:, None] - truth[:, 2][None, :]), axis=1)
    assert np.median(dist) < 2 * step


def test_batch_mixed_lengths_settings_and_order():
    specs = []
    for s, (n, integer) in enumerate([(32768, False), (16384, True), (32768, False), (20000, False), (8192, True)]):
        x = synth.axis(n)
        specs.append(Spectrum(x, synth.config3(10 + s, n=n, integer=integer, x=x), (-2.2, 11.8)))
    dec = Deconvoluter()
    _check_e2e(dec, O.Settings(), specs, "mixed batch default")
    dec.set_moving_average_smoother(2, 5)
    dec.set_noise_score_selector(6.5)
    dec.set_analytical_fitter(5)
    dec.add_ignore_region((4.7, 4.9))
    _check_e2e(dec, O.Settings(smoothing_iterations=2, smoothing_window=5, threshold=6.5, fitting_iterations=5,
                               ignore_regions=[(4.7, 4.9)]), specs, "mixed batch custom")
    dec = Deconvoluter()
    dec.set_identity_smoother()
    _check_e2e(dec, O.Settings(smoothing_kind=O.SMOOTH_IDENTITY), specs[:2], "identity smoother")
    dec = Deconvoluter()
    dec.set_detector_only()
    dec.set_analytical_fitter(2)
    _check_e2e(dec, O.Settings(selection_kind=O.SELECT_DETECTOR_ONLY, fitting_iterations=2), specs[1:2], "detector only")


def test_increasing_axis_and_ignore_regions():
    """The reference accepts either axis direction (meta/monotonicity.rs); Spectrum::new orders the
    signal boundaries to the axis (spectrum.rs:854-863) and the index helpers use the signed step.
    An increasing axis is the decreasing one read backwards."""
    n = 32768
    xd = synth.axis(n)
    yd = synth.config3(21, n=n, x=xd)
    xi, yi = xd[::-1].copy(), yd[::-1].copy()
    sp = Spectrum(xi, yi, (11.8, -2.2))          # given in the "wrong" order on purpose
    assert sp.signal_boundaries == (-2.2, 11.8)
    dec = Deconvoluter()
    _check_e2e(dec, O.Settings(), [sp], "increasing axis")
    dec.add_ignore_region((4.9, 4.7))
    dec.add_ignore_region((7.0, 7.5))
    dec.add_ignore_region((7.4, 8.0))            # merges with the previous one (deconvoluter.rs:438-472)
    assert dec.ignore_regions() == [(4.7, 4.9), (7.0, 8.0)]
    settings = O.Settings(ignore_regions=[(4.7, 4.9), (7.0, 8.0)])
    _check_e2e(dec, settings, [sp], "increasing axis with two ignore regions")
    # The same two regions on a DECREASING axis: the reference builds its MSE ranges from the regions in
    # ppm order (deconvoluter.rs:828-845), whose indices then run backwards, and the slice
    # `superpositions[start..end]` panics.  The library reports exactly that instead of inventing a result.
    spd = Spectrum(xd, yd, (-2.2, 11.8))
    r = O.deconvolute_spectrum(settings, spd.chemical_shifts, spd.intensities, spd.signal_boundaries)
    assert r.status == O.PANIC
    with pytest.raises(exceptions.UnexpectedError, match="reference implementation panics"):
        dec.deconvolute_spectrum(spd)
    dec.clear_ignore_regions()
    dec.add_ignore_region((4.7, 4.9))            # a single region is fine in either direction
    _check_e2e(dec, O.Settings(ignore_regions=[(4.7, 4.9)]), [spd, sp], "one ignore region, both directions")


def test_batch_spanning_several_chunks(monkeypatch):
    monkeypatch.setenv("MDB_CHUNK_SPECTRA", "3")
    n = 8192
    x = synth.axis(n)
    specs = [Spectrum(x, synth.config3(100 + s, n=n, x=x), (-2.2, 11.8)) for s in range(11)]
    _check_e2e(Deconvoluter(), O.Settings(), specs, "11 spectra in chunks of 3")


def test_error_semantics_first_failure_wins():
    n = 4096
    x = synth.axis(n)
    good = Spectrum(x, synth.config3(5, n=n, x=x), (-2.2, 11.8))
    flat = Spectrum(x, np.zeros(n), (-2.2, 11.8))                       # NoPeaksDetected
    narrow = Spectrum(x, synth.config3(6, n=n, x=x), (5.0, 5.0 + 3 * abs(x[1] - x[0])))  # nothing selectable
    dec = Deconvoluter()
    with pytest.raises(exceptions.NoPeaksDetected):
        dec.deconvolute_spectra([good, flat, narrow])
    with pytest.raises(exceptions.EmptySignalRegion):
        dec.deconvolute_spectra([good, narrow, flat])
    r = O.deconvolute_spectrum(O.Settings(), narrow.chemical_shifts, narrow.intensities, narrow.signal_boundaries)
    assert r.status == O.EMPTY_SIGNAL_REGION
    assert len(dec.deconvolute_spectra([])) == 0


def test_device_resident_inputs_match_host_inputs():
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    n = 16384
    x = synth.axis(n)
    ys = np.stack([synth.config3(40 + s, n=n, x=x) for s in range(4)])
    dec = Deconvoluter()
    host = dec.deconvolute_spectra([Spectrum(x, y, (-2.2, 11.8)) for y in ys])
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(ys).cuda()
    views = (_lib.SpectrumView * 4)()
    for i in range(4):
        views[i].chemical_shifts = xd.data_ptr()
        views[i].intensities = yd[i].data_ptr()
        views[i].len = n
        views[i].signal_boundaries[0], views[i].signal_boundaries[1] = 11.8, -2.2
    batch = C.c_void_p()
    assert lib.mdb_deconvolute_spectra(dec._h, views, 4, _lib.MDB_MEM_DEVICE, C.byref(batch)) == 0, _lib.last_error()
    try:
        for i in range(4):
            k = lib.mdb_batch_n_lorentzians(batch, i)
            got = np.ctypeslib.as_array(C.cast(lib.mdb_batch_lorentzians(batch, i), C.POINTER(C.c_double)), (k, 3)).copy()
            assert_bit_equal(got, host[i].parameters, f"device input {i}")
            assert lib.mdb_batch_mse(batch, i) == host[i].mse
    finally:
        lib.mdb_batch_free(batch)


def test_persistent_fit_kernel_matches_multi_launch(monkeypatch, golden_dir):
    """MDB_FIT_PERSISTENT=1: all refinement passes in one work-queue launch (an experiment kept
    behind an environment variable); results must be the same bits, including per-spectrum
    iteration counts (optimize_settings) and spectra of different peak counts in one chunk."""
    n = 16384
    x = synth.axis(n)
    specs = [Spectrum(x, synth.config3(500 + s, n=n, x=x), (-2.2, 11.8)) for s in range(6)]
    specs.append(Spectrum(synth.axis(4096), synth.config3(9, n=4096), (-2.2, 11.8)))
    dec = Deconvoluter()
    base = dec.deconvolute_spectra(specs)
    monkeypatch.setenv("MDB_FIT_PERSISTENT", "1")
    pers = dec.deconvolute_spectra(specs)
    for a, b in zip(base, pers):
        assert_bit_equal(b.parameters, a.parameters, "persistent fit")
        assert a.mse == b.mse
    sim = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.339, 3.553))
    d1, d2 = Deconvoluter(), Deconvoluter()
    m2 = d2.optimize_settings(sim)
    monkeypatch.delenv("MDB_FIT_PERSISTENT")
    m1 = d1.optimize_settings(sim)
    assert m1 == m2 and d1.fitting_settings() == d2.fitting_settings() and d1.smoothing_settings() == d2.smoothing_settings()


# ------------------------------------------------------------------------------ optimize_settings
def test_optimize_settings_sim_matches_oracle(golden_dir):
    """deconvoluter.rs:761-825 on the reference's own example spectrum (sim_01, 3.339..3.553):
    same optimum, same MSE bits, and the deconvoluter is left holding the optimal settings."""
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.339, 3.553))
    status, best, want_mse, all_mse = O.optimize_settings(O.Settings(), sp.chemical_shifts, sp.intensities,
                                                          sp.signal_boundaries)
    assert status == O.OK
    dec = Deconvoluter()
    mse = dec.optimize_settings(sp)
    assert_bit_equal([mse], [want_mse], "optimal mse")
    assert dec.smoothing_settings() == {"method": "MovingAverage", "iterations": best[0], "windowSize": best[1]}
    assert dec.selection_settings()["threshold"] == best[2]
    assert dec.fitting_settings() == {"method": "Analytical", "iterations": best[3]}
    # deconvoluting with the optimised settings reproduces that MSE
    out = dec.deconvolute_spectrum(sp)
    assert_bit_equal([out.mse], [want_mse], "mse with the optimal settings")


def test_optimize_settings_blood_with_ignore_region(blood_arrays):
    x, y = blood_arrays
    sp = Spectrum(x, y, (-2.2, 11.8))
    dec = Deconvoluter()
    dec.add_ignore_region((4.7, 4.9))
    mse = dec.optimize_settings(sp)
    status, best, want_mse, _ = O.optimize_settings(O.Settings(ignore_regions=[(4.7, 4.9)]), x, y, sp.signal_boundaries)
    assert status == O.OK
    assert_bit_equal([mse], [want_mse], "blood optimal mse")
    assert dec.smoothing_settings() == {"method": "MovingAverage", "iterations": best[0], "windowSize": best[1]}
    assert dec.selection_settings()["threshold"] == best[2] and dec.fitting_settings()["iterations"] == best[3]
    assert dec.ignore_regions() == [(4.7, 4.9)]


def test_optimize_settings_error_leaves_settings_unchanged():
    n = 4096
    x = synth.axis(n)
    flat = Spectrum(x, np.zeros(n), (-2.2, 11.8))
    dec = Deconvoluter()
    before = (dec.smoothing_settings(), dec.selection_settings(), dec.fitting_settings())
    with pytest.raises(exceptions.NoPeaksDetected):
        dec.optimize_settings(flat)
    assert before == (dec.smoothing_settings(), dec.selection_settings(), dec.fitting_settings())


# ------------------------------------------------------------------------------ in-process multi-GPU
def test_in_process_device_sharding_matches_single_device():
    """mdb_set_device_count: the batch is cut into contiguous shards, one pipeline per GPU.  On a
    one-GPU box the request degrades to one device; either way results equal the default path."""
    import metabodecon_rust_b200 as M
    n = 8192
    x = synth.axis(n)
    specs = [Spectrum(x, synth.config3(300 + s, n=n, x=x), (-2.2, 11.8)) for s in range(9)]
    dec = Deconvoluter()
    base = dec.deconvolute_spectra(specs)
    M.set_devices(0)  # all visible devices
    try:
        multi = dec.deconvolute_spectra(specs)
        bad = specs[:5] + [Spectrum(x, np.zeros(n), (-2.2, 11.8))] + specs[5:]
        with pytest.raises(exceptions.NoPeaksDetected):
            dec.deconvolute_spectra(bad)
    finally:
        M.set_devices(1)
    for i, (a, b) in enumerate(zip(base, multi)):
        assert np.array_equal(a.peaks, b.peaks)
        assert_bit_equal(a.parameters, b.parameters, f"shard result {i}")
        assert a.mse == b.mse
    with pytest.raises(Exception):
        M.set_devices(-1)


# ------------------------------------------------------------------------------ full-size properties
# BASELINE.json sizes, where the oracle would take minutes: size-independent exact properties.
def _config5_batch(n_spec, n=131072):
    x = synth.axis(n)
    base = synth.config5(11, n=n, x=x)
    rng = np.random.default_rng(5)
    return x, [base + rng.normal(0.0, 40.0, n) for _ in range(n_spec)]


def test_full_size_power_of_two_scaling_is_exact():
    """Multiplying the intensities by 2^k is exact in binary floating point and commutes with every
    operation of the path (sums, the 1/len products, comparisons, the three-point solve, the
    division): peaks must be identical, sfhw scale by 2^k, hw2 and maxp stay, the MSE scales by 4^k
    -- bit for bit, at 2^17 points and ~2,100 peaks per spectrum."""
    x, ys = _config5_batch(3)
    dec = Deconvoluter()
    base = dec.deconvolute_spectra([Spectrum(x, y, (-2.2, 11.8)) for y in ys])
    for k in (3, -7):
        f = 2.0 ** k
        scaled = dec.deconvolute_spectra([Spectrum(x, y * f, (-2.2, 11.8)) for y in ys])
        for a, b in zip(base, scaled):
            assert a.peaks.shape[0] > 1500
            assert np.array_equal(a.peaks, b.peaks)
            assert_bit_equal(b.parameters[:, 0], a.parameters[:, 0] * f, f"sfhw x 2^{k}")
            assert_bit_equal(b.parameters[:, 1:], a.parameters[:, 1:], f"hw2, maxp under 2^{k}")
            assert_bit_equal([b.mse], [a.mse * f * f], f"mse x 4^{k}")


def test_full_size_results_do_not_depend_on_batch_position():
    """The same spectrum placed at different batch indices (different chunks, streams, CTAs and
    tile offsets; 150 spectra = several adaptive chunks) must give identical bits every time."""
    x, ys = _config5_batch(2)
    order = [0, 1] * 75
    dec = Deconvoluter()
    outs = dec.deconvolute_spectra([Spectrum(x, ys[i], (-2.2, 11.8)) for i in order])
    ref = {0: outs[0], 1: outs[1]}
    for i, out in zip(order, outs):
        assert np.array_equal(out.peaks, ref[i].peaks)
        assert_bit_equal(out.parameters, ref[i].parameters, "batch position")
        assert out.mse == ref[i].mse
    # and a sample of it agrees with the oracle
    r = O.deconvolute_spectrum(O.Settings(), x, ys[1], synth.SIGNAL_BOUNDARIES)
    assert_bit_equal(outs[1].parameters, r.lorentzians, "oracle spot check")


def test_config4_size_superposition_grid_split_and_scaling():
    """Config 4 scale (2^24 grid points x 20,000 Lorentzians would take the oracle ~10 minutes; a
    2^22-point slice of it is used here): grid points are independent, so evaluating the grid in
    slices (the multi-GPU sharding) must reproduce the one-shot result bit for bit, scaling sfhw by
    a power of two scales every value exactly, and 4,096 sampled points agree with the oracle."""
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    rng = np.random.Generator(np.random.PCG64(20260004))
    p = 20000
    maxp = rng.uniform(0.0, 10.0, p)
    hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
    sf = np.exp(rng.uniform(0.0, np.log(1e4), p))
    lor = np.ascontiguousarray(np.stack([sf * hw, hw * hw, maxp], axis=1))
    x = np.linspace(-2.2, 11.8, 1 << 24)[: 1 << 22]
    whole = superposition_vec_array(x, lor)
    parts = np.concatenate([superposition_vec_array(x[lo:hi], lor) for lo, hi in ((0, 1000001), (1000001, 3 << 20), (3 << 20, 1 << 22))])
    assert_bit_equal(parts, whole, "grid slices")
    lor8 = lor.copy()
    lor8[:, 0] *= 8.0
    assert_bit_equal(superposition_vec_array(x, lor8), whole * 8.0, "sfhw x 8")
    idx = rng.integers(0, x.size, 4096)
    assert_bit_equal(whole[idx], O.superposition_vec(x[idx], lor, parallel=True), "oracle sample")


# ------------------------------------------------------------------------------ the ABI from plain C
def test_c_host_program_through_the_abi(golden_dir, tmp_path):
    """examples/deconvolute.c: a compiled host linking libmdb200.so directly (no Python, no torch
    types) reproduces BASELINE config 1 -- the Appendix-B checkpoints, bit for bit."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = str(tmp_path / "deconvolute")
    libdir = os.path.join(root, "metabodecon_rust_b200")
    subprocess.run([gcc, "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "deconvolute.c"),
                    "-o", exe, "-L" + libdir, "-lmdb200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    out = subprocess.run([exe, os.path.join(golden_dir, "bruker", "blood_01", "10", "pdata", "10", "1r")],
                         check=True, capture_output=True, text=True).stdout
    first = out.splitlines()[0].split()
    assert first[:6] == ["points", "131072", "selected_peaks", "981", "lorentzians", "760"]
    assert float.fromhex(first[7]) == float.fromhex("0x1.0808a64fe177ep+35")
    l0 = out.splitlines()[1].split()
    assert float.fromhex(l0[2]) == float.fromhex("0x1.084fd4b50b8dfp-4") and float.fromhex(l0[6]) == float.fromhex("0x1.0eac0d0cd9679p+3")


# ------------------------------------------------------------------------------ randomised differential
def test_randomised_differential_gpu_vs_oracle():
    """120 random spectra (integer / float values, flat stretches, 200..6000 points) under random
    settings, in random-size batches: the GPU result must equal the oracle's in bits, and error
    statuses must agree."""
    from test_oracle_kats import _random_spectrum
    rng = np.random.default_rng(777)
    done = 0
    while done < 120:
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
        first_bad = next((w.status for w in wants if w.status != O.OK), O.OK)
        if first_bad != O.OK:
            with pytest.raises(Exception):
                dec.deconvolute_spectra(specs)
            done += batch
            continue
        outs = dec.deconvolute_spectra(specs)
        for i, (out, w) in enumerate(zip(outs, wants)):
            what = f"random case {done + i} (iters={iters}, window={window}, thr={thr:.3f}, fit={fit})"
            assert np.array_equal(out.peaks.astype(np.int64), w.peaks.astype(np.int64)), what
            nan = np.isnan(w.lorentzians)
            assert_bit_equal(np.where(nan, 0.0, out.parameters), np.where(nan, 0.0, w.lorentzians), what)
            assert out.mse == w.mse or (np.isnan(out.mse) and np.isnan(w.mse)), what
        done += batch


# ------------------------------------------------------------------------------ small-spectrum path
def _launches_of(fn):
    lib = _lib.load()
    lib.mdb_reset_kernel_launch_count()
    out = fn()
    return out, int(lib.mdb_kernel_launch_count())


def _same_results(a, b, what):
    assert len(a) == len(b)
    for i, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u.peaks, v.peaks), f"{what}[{i}]: peaks"
        assert np.array_equal(bits(u.parameters), bits(v.parameters)), f"{what}[{i}]: lorentzians"
        assert bits([u.mse])[0] == bits([v.mse])[0], f"{what}[{i}]: mse"


def test_small_path_is_one_launch_and_matches_general_path_and_oracle(monkeypatch, golden_dir):
    """Spectra of <= 4 096 points run as ONE fused launch (small_fused.cuh), smoothing included; the
    same kernel behind the batch smoothing kernel (MDB_SMALL_SMOOTH=separate), the general chunk
    pipeline (MDB_SMALL_PATH=0) and the oracle must all give the same bits."""
    sim = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    dec = Deconvoluter()
    small, launches = _launches_of(lambda: _check_e2e(dec, O.Settings(), [sim], "sim_01 small path"))
    assert launches == 1, launches
    monkeypatch.setenv("MDB_SMALL_SMOOTH", "separate")
    two, launches_two = _launches_of(lambda: dec.deconvolute_spectra([sim]))
    monkeypatch.delenv("MDB_SMALL_SMOOTH")
    assert launches_two == 2, launches_two
    _same_results(small, two, "sim_01, smoothing as its own launch")
    monkeypatch.setenv("MDB_SMALL_PATH", "0")
    general, launches_general = _launches_of(lambda: dec.deconvolute_spectra([sim]))
    monkeypatch.delenv("MDB_SMALL_PATH")
    assert launches_general > 10
    _same_results(small, general, "sim_01")
    # a batch: mixed lengths up to the limit, integer and float values, every spectrum its own axis object
    specs = []
    for s, n in enumerate([4096, 2048, 4095, 517, 64, 3000, 4096, 1025]):
        x = synth.axis(n)
        specs.append(Spectrum(x, synth.spectrum(300 + s, n=n, k=max(3, n // 60), hw_range=(8e-3, 5e-2), integer=bool(s & 1), x=x),
                              (-2.2, 11.8)))
    for setup, settings in [
        (lambda d: None, O.Settings()),
        (lambda d: (d.set_moving_average_smoother(4, 5), d.set_noise_score_selector(2.5), d.set_analytical_fitter(3),
                    d.add_ignore_region((4.7, 4.9))),
         O.Settings(smoothing_iterations=4, smoothing_window=5, threshold=2.5, fitting_iterations=3, ignore_regions=[(4.7, 4.9)])),
        (lambda d: (d.set_identity_smoother(), d.set_noise_score_selector(3.0)),
         O.Settings(smoothing_kind=O.SMOOTH_IDENTITY, threshold=3.0)),
        (lambda d: (d.set_detector_only(), d.set_analytical_fitter(2)),
         O.Settings(selection_kind=O.SELECT_DETECTOR_ONLY, fitting_iterations=2)),
        (lambda d: (d.set_moving_average_smoother(2, 11), d.set_noise_score_selector(1.0)),
         O.Settings(smoothing_iterations=2, smoothing_window=11, threshold=1.0)),
        (lambda d: (d.set_moving_average_smoother(7, 2), d.set_noise_score_selector(1.5)),
         O.Settings(smoothing_iterations=7, smoothing_window=2, threshold=1.5)),
        (lambda d: (d.set_moving_average_smoother(1, 40), d.set_noise_score_selector(1.5)),
         O.Settings(smoothing_iterations=1, smoothing_window=40, threshold=1.5)),
        (lambda d: (d.set_moving_average_smoother(33, 3), d.set_noise_score_selector(1.5)),  # > 32 passes: smoothing launched separately
         O.Settings(smoothing_iterations=33, smoothing_window=3, threshold=1.5)),
        # windows 3, 5, 7: the interior loop specialised on the window (register-resident blocks), many passes
        (lambda d: (d.set_moving_average_smoother(12, 3), d.set_noise_score_selector(1.5)),
         O.Settings(smoothing_iterations=12, smoothing_window=3, threshold=1.5)),
        (lambda d: (d.set_moving_average_smoother(9, 7), d.set_noise_score_selector(1.5)),
         O.Settings(smoothing_iterations=9, smoothing_window=7, threshold=1.5)),
        (lambda d: (d.set_moving_average_smoother(32, 5), d.set_noise_score_selector(1.5)),
         O.Settings(smoothing_iterations=32, smoothing_window=5, threshold=1.5)),
        (lambda d: (d.set_moving_average_smoother(1, 5), d.set_noise_score_selector(4.0)),
         O.Settings(smoothing_iterations=1, smoothing_window=5, threshold=4.0)),
    ]:
        dec = Deconvoluter()
        setup(dec)
        ok = [sp for sp in specs
              if O.deconvolute_spectrum(settings, sp.chemical_shifts, sp.intensities, sp.signal_boundaries).status == O.OK]
        assert len(ok) >= 4
        small, launches = _launches_of(lambda: _check_e2e(dec, settings, ok, "small batch"))
        assert launches == (1 if settings.smoothing_iterations <= 32 else 34), launches
        monkeypatch.setenv("MDB_SMALL_PATH", "0")
        general = dec.deconvolute_spectra(ok)
        monkeypatch.delenv("MDB_SMALL_PATH")
        _same_results(small, general, "small batch vs general")
        if settings.smoothing_kind != O.SMOOTH_IDENTITY and settings.smoothing_window in (3, 5, 7):
            monkeypatch.setenv("MDB_STREAM_GENERIC", "1")  # the any-window interior loop on the same inputs
            generic = dec.deconvolute_spectra(ok)
            monkeypatch.delenv("MDB_STREAM_GENERIC")
            _same_results(small, generic, "small batch, specialised vs any-window smoothing loop")


def test_small_path_dense_peaks_short_inputs_and_errors():
    """Worst-case peak density (a centre every other point fills the per-centre arrays), more
    selected peaks than the CTA has threads, the shortest inputs, and every error status."""
    rng = np.random.default_rng(4096)
    # zig-zag: the second difference alternates in sign, so every other interior point is a centre
    n = 4096
    x = synth.axis(n)
    zig = 1000.0 + 50.0 * (np.arange(n) % 2) + rng.uniform(0.0, 5.0, n) + 4000.0 / (1.0 + ((x - 5.0) / 0.5) ** 2)
    sp = Spectrum(x, zig, (-2.2, 11.8))
    dec = Deconvoluter()
    dec.set_identity_smoother()
    dec.set_detector_only()
    dec.set_analytical_fitter(3)
    settings = O.Settings(smoothing_kind=O.SMOOTH_IDENTITY, selection_kind=O.SELECT_DETECTOR_ONLY, fitting_iterations=3)
    r = O.deconvolute_spectrum(settings, x, zig, sp.signal_boundaries)
    assert r.status == O.OK and len(r.peaks) > 1200, len(r.peaks)   # > 256 peaks: several peaks per thread
    _check_e2e(dec, settings, [sp], "zig-zag, detector only")
    dec = Deconvoluter()
    dec.set_identity_smoother()
    dec.set_noise_score_selector(0.1)
    settings = O.Settings(smoothing_kind=O.SMOOTH_IDENTITY, threshold=0.1)
    r = O.deconvolute_spectrum(settings, x, zig, sp.signal_boundaries)
    assert r.status == O.OK and len(r.peaks) > 256, len(r.peaks)
    _check_e2e(dec, settings, [sp], "zig-zag, noise score filter")
    # shortest inputs the API accepts, default and loose settings: statuses must agree with the oracle
    from metabodecon_rust_b200 import exceptions as E
    status_to_exc = {O.NO_PEAKS_DETECTED: E.NoPeaksDetected, O.EMPTY_SIGNAL_REGION: E.EmptySignalRegion,
                     O.EMPTY_SIGNAL_FREE_REGION: E.EmptySignalFreeRegion, O.PANIC: E.UnexpectedError}
    seen = set()
    for n in list(range(5, 40)) + [63, 64, 65, 255, 256, 257]:
        for trial in range(4):
            xs = 10.0 - np.arange(n) * (10.0 / (n - 1))
            y = np.rint(rng.normal(0.0, 100.0, n)) if trial & 1 else rng.normal(0.0, 100.0, n)
            lo, hi = sorted(rng.uniform(0.5, 9.5, 2))
            sp = Spectrum(xs, y, (lo, hi))
            for thr, (iters, window) in ((0.01, (1, 3)), (5.0, (1, 3)), (0.01, (2, 9)), (0.01, (3, 41)), (0.5, (5, 2))):
                if n < window // 2:
                    continue  # the reference panics (moving_average.rs:62); covered by test_smoothing_short_inputs
                dec = Deconvoluter()
                dec.set_moving_average_smoother(iters, window)
                dec.set_noise_score_selector(thr)
                settings = O.Settings(smoothing_iterations=iters, smoothing_window=window, threshold=thr)
                r = O.deconvolute_spectrum(settings, xs, y, sp.signal_boundaries)
                seen.add(r.status)
                if r.status == O.OK:
                    _check_e2e(dec, settings, [sp], f"n={n}")
                else:
                    with pytest.raises(status_to_exc[r.status]):
                        dec.deconvolute_spectrum(sp)
    assert O.OK in seen and len(seen) >= 3, seen


def test_small_path_many_spectra_device_memory_and_chunking():
    """2 000 small spectra in one call (several fused launches: the result slots are capped at
    64 MiB per chunk), from host memory and from device memory; results must not depend on either."""
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    n = 4096
    x = synth.axis(n)
    base = [synth.spectrum(500 + s, n=n, k=40, hw_range=(8e-3, 5e-2), x=x) for s in range(8)]
    want = [O.deconvolute_spectrum(O.Settings(), x, y, synth.SIGNAL_BOUNDARIES) for y in base]
    assert all(w.status == O.OK for w in want)
    total = 2000
    specs = [Spectrum(x, base[s % 8], (-2.2, 11.8)) for s in range(total)]
    dec = Deconvoluter()
    outs, launches = _launches_of(lambda: dec.deconvolute_spectra(specs))
    assert launches == 3, launches   # 909 result slots of 73 792 bytes fit in 64 MiB -> chunks of 909, 909, 182
    for s in range(total):
        w = want[s % 8]
        assert np.array_equal(outs[s].peaks.astype(np.int64), w.peaks.astype(np.int64)), s
        assert np.array_equal(bits(outs[s].parameters), bits(w.lorentzians)), s
        assert outs[s].mse == w.mse, s
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(np.stack(base)).cuda()
    views = (_lib.SpectrumView * 8)()
    for i in range(8):
        views[i].chemical_shifts = xd.data_ptr()
        views[i].intensities = yd[i].data_ptr()
        views[i].len = n
        views[i].signal_boundaries[0], views[i].signal_boundaries[1] = 11.8, -2.2
    batch = C.c_void_p()
    assert lib.mdb_deconvolute_spectra(dec._h, views, 8, _lib.MDB_MEM_DEVICE, C.byref(batch)) == 0, _lib.last_error()
    try:
        for i in range(8):
            k = lib.mdb_batch_n_lorentzians(batch, i)
            got = np.ctypeslib.as_array(C.cast(lib.mdb_batch_lorentzians(batch, i), C.POINTER(C.c_double)), (k, 3)).copy()
            assert np.array_equal(bits(got), bits(want[i].lorentzians)), i
            assert lib.mdb_batch_mse(batch, i) == want[i].mse
    finally:
        lib.mdb_batch_free(batch)


# ------------------------------------------------------------------------------ re-entrancy
def test_concurrent_callers_share_one_deconvoluter():
    """The reference's Deconvoluter is Send + Sync (deconvoluter.rs:913-917) and rayon calls it from
    many workers at once; the C ABI must be re-entrant: four threads, one shared handle, different
    batches, every result identical to the single-threaded one."""
    import threading
    n = 8192
    x = synth.axis(n)
    batches = [[Spectrum(x, synth.config3(900 + 10 * t + s, n=n, x=x), (-2.2, 11.8)) for s in range(5)] for t in range(4)]
    dec = Deconvoluter()
    dec.add_ignore_region((4.7, 4.9))
    want = [dec.deconvolute_spectra(b) for b in batches]
    got, errors = [None] * 4, []

    def work(t):
        try:
            for _ in range(3):
                got[t] = dec.deconvolute_spectra(batches[t])
        except Exception as err:  # noqa: BLE001
            errors.append(err)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(4):
        for a, b in zip(want[t], got[t]):
            assert np.array_equal(a.peaks, b.peaks)
            assert_bit_equal(b.parameters, a.parameters, f"thread {t}")
            assert a.mse == b.mse
    # the same through the small-spectrum path (one fused launch per call, result slots in pinned memory)
    xs = synth.axis(2048)
    small = [[Spectrum(xs, synth.spectrum(1200 + 10 * t + s, n=2048, k=30, hw_range=(8e-3, 5e-2), x=xs), (-2.2, 11.8))
              for s in range(1 + t)] for t in range(4)]
    want = [dec.deconvolute_spectra(b) for b in small]
    got, errors = [None] * 4, []

    def work_small(t):
        try:
            for _ in range(25):
                got[t] = dec.deconvolute_spectra(small[t])
        except Exception as err:  # noqa: BLE001
            errors.append(err)

    threads = [threading.Thread(target=work_small, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(4):
        for a, b in zip(want[t], got[t]):
            assert np.array_equal(a.peaks, b.peaks)
            assert_bit_equal(b.parameters, a.parameters, f"small path, thread {t}")
            assert a.mse == b.mse
