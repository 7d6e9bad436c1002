#!/usr/bin/env python
"""The reference's own criterion benchmarks (benches/deconvoluter.rs:8-52), re-run here: the GPU
library through its Python mirror against the oracle port (serial, and OpenMP where the reference
uses rayon).  Fixtures: sim_01 and blood_01 from tests/golden (the reference's 16-spectrum sets are
replaced by 16 copies of the one fixture that ships with this repo -- same sizes, same settings).
Development aid; run on the GPU box.  Prints one JSON line per benchmark."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "bruker")
sim = Spectrum.read_bruker(os.path.join(G, "sim_01"), 10, 10, (3.34, 3.56))
blood = Spectrum.read_bruker(os.path.join(G, "blood_01"), 10, 10, (-2.2, 11.8))
dec = Deconvoluter()
cores = O.use_all_cores()


def best_of(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


def oracle_one(sp, parallel):
    r = O.deconvolute_spectrum(O.Settings(), sp.chemical_shifts, sp.intensities, sp.signal_boundaries, parallel=parallel)
    assert r.status == O.OK
    return r


def oracle_many(sp, count):
    ys = np.tile(sp.intensities, (count, 1))
    st, *_ = O.par_deconvolute_spectra(O.Settings(), sp.chemical_shifts, ys, sp.signal_boundaries)
    assert st == O.OK


rows = [
    ("deconvolute_sim_spectrum", lambda: dec.deconvolute_spectrum(sim), lambda: oracle_one(sim, False), 1),
    ("deconvolute_blood_spectrum", lambda: dec.deconvolute_spectrum(blood), lambda: oracle_one(blood, False), 1),
    ("parallel_deconvolute_sim_spectrum", lambda: dec.par_deconvolute_spectrum(sim), lambda: oracle_one(sim, True), cores),
    ("parallel_deconvolute_blood_spectrum", lambda: dec.par_deconvolute_spectrum(blood), lambda: oracle_one(blood, True), cores),
    ("parallel_deconvolute_sim_spectra(16)", lambda: dec.par_deconvolute_spectra([sim] * 16), lambda: oracle_many(sim, 16), cores),
    ("parallel_deconvolute_blood_spectra(16)", lambda: dec.par_deconvolute_spectra([blood] * 16), lambda: oracle_many(blood, 16), cores),
]
for name, gpu, cpu, threads in rows:
    g_min, g_med = best_of(gpu, 20)
    c_min, c_med = best_of(cpu, 3)
    print(json.dumps({"bench": name, "gpu_ms_median": g_med * 1e3, "gpu_ms_min": g_min * 1e3, "oracle_ms_median": c_med * 1e3,
                      "oracle_threads": threads, "speedup_median": c_med / g_med}), flush=True)

# the same single-spectrum call without the Python wrappers on either side: the C ABI entry of the
# library against the C entry of the oracle, buffers prepared once (what a Rust caller would see)
import ctypes as C  # noqa: E402
from metabodecon_rust_b200 import _lib  # noqa: E402
lib = _lib.load()


def raw_gpu(sp):
    views = (_lib.SpectrumView * 1)()
    views[0].chemical_shifts = sp.chemical_shifts.ctypes.data
    views[0].intensities = sp.intensities.ctypes.data
    views[0].len = sp.chemical_shifts.size
    views[0].signal_boundaries[0], views[0].signal_boundaries[1] = sp.signal_boundaries
    batch = C.c_void_p()

    def call():
        assert lib.mdb_deconvolute_spectra(dec._h, views, 1, _lib.MDB_MEM_HOST, C.byref(batch)) == 0
        lib.mdb_batch_free(batch)
    return call


def raw_cpu(sp, parallel):
    n = sp.intensities.size
    cap = n // 2 + 1
    st, _keep = O.Settings()._c()
    lor = np.zeros((cap, 3))
    sel = [np.zeros(cap, dtype=np.uintp) for _ in range(3)]
    sm = np.empty(n)
    res = O._Result()
    x, y = np.ascontiguousarray(sp.chemical_shifts), np.ascontiguousarray(sp.intensities)
    args = (C.byref(st), O._dp(x), O._dp(y), C.c_size_t(n), C.c_double(sp.signal_boundaries[0]),
            C.c_double(sp.signal_boundaries[1]), C.c_int(parallel), O._dp(lor), O._sp(sel[0]), O._sp(sel[1]), O._sp(sel[2]),
            O._dp(sm), C.byref(res))
    fn = O.lib().orc_deconvolute_spectrum
    keep = (st, _keep, lor, sel, sm, res, x, y)

    def call():
        fn(*args)
        assert res.status == O.OK and keep
    return call


for name, sp, parallel in [("deconvolute_sim_spectrum [C ABI]", sim, False), ("parallel_deconvolute_sim_spectrum [C ABI]", sim, True),
                           ("deconvolute_blood_spectrum [C ABI]", blood, False)]:
    g_min, g_med = best_of(raw_gpu(sp), 50)
    c_min, c_med = best_of(raw_cpu(sp, parallel), 5)
    print(json.dumps({"bench": name, "gpu_ms_median": g_med * 1e3, "gpu_ms_min": g_min * 1e3, "oracle_ms_median": c_med * 1e3,
                      "oracle_threads": cores if parallel else 1, "speedup_median": c_med / g_med}), flush=True)

# where a single small spectrum spends its time on the device (CUDA events inside the library)
lib.mdb_profile_enable(1)
lib.mdb_profile_reset()
for _ in range(20):
    dec.deconvolute_spectrum(sim)
prof = {}
for kid, name in enumerate(_lib.KERNEL_NAMES):
    ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
    lib.mdb_profile_read(kid, C.byref(ms), C.byref(n), C.byref(work))
    if n.value:
        prof[name] = {"us_per_launch": 1e3 * ms.value / n.value, "launches": int(n.value)}
lib.mdb_profile_enable(0)
print(json.dumps({"bench": "sim_spectrum_device_time", "kernels": prof}), flush=True)
