#!/usr/bin/env python
"""One default deconvolution of sim_01 through the small-spectrum path (ncu target)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

sim = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "sim_01"), 10, 10, (3.34, 3.56))
dec = Deconvoluter()
for _ in range(3):
    out = dec.deconvolute_spectrum(sim)
print(len(out.lorentzians), out.mse)
