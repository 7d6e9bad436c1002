#!/usr/bin/env python
"""Per-kernel device time of one default blood_01 deconvolution (CUDA events inside the library)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metabodecon_rust_b200 import Deconvoluter, Spectrum, _lib  # noqa: E402

blood = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
dec = Deconvoluter()
lib = _lib.load()
for _ in range(3):
    dec.deconvolute_spectrum(blood)
lib.mdb_profile_enable(1)
lib.mdb_profile_reset()
reps = 10
for _ in range(reps):
    dec.deconvolute_spectrum(blood)
out = {}
for kid, name in enumerate(_lib.KERNEL_NAMES):
    ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
    lib.mdb_profile_read(kid, C.byref(ms), C.byref(n), C.byref(work))
    if n.value:
        out[name] = {"ms_per_call": ms.value / reps, "launches_per_call": n.value / reps}
print(json.dumps(out))
