#!/usr/bin/env python
"""Two default deconvolutions of blood_01 (single-spectrum call: latency forms of K1 and K6; ncu target)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

blood = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
dec = Deconvoluter()
dec.add_ignore_region((4.7, 4.9))
for _ in range(2):
    out = dec.deconvolute_spectrum(blood)
print(len(out.lorentzians), out.mse)
