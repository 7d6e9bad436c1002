#!/usr/bin/env python
"""Times Deconvoluter.optimize_settings (810 deconvolutions of blood_01) on the GPU against the
oracle port on the host cores.  Development aid; run on the GPU box."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

sp = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
dec = Deconvoluter()
dec.optimize_settings(sp)  # warm-up (workspace allocation)
t0 = time.perf_counter()
mse = Deconvoluter().optimize_settings(sp)
t_gpu = time.perf_counter() - t0
cores = O.use_all_cores()
t0 = time.perf_counter()
status, best, want, _ = O.optimize_settings(O.Settings(), sp.chemical_shifts, sp.intensities, sp.signal_boundaries)
t_cpu = time.perf_counter() - t0
print(f"optimize_settings(blood_01): GPU {t_gpu * 1e3:.1f} ms, oracle {t_cpu:.2f} s on {cores} threads, "
      f"speed-up {t_cpu / t_gpu:.1f}x, best {best}, mse equal: {mse == want}")
