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

# where a single small spectrum spends its time on the device (CUDA events inside the library)
import ctypes as C  # noqa: E402
from metabodecon_rust_b200 import _lib  # noqa: E402
lib = _lib.load()
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
