#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of
the pipeline, both tile classes of the smoothing kernel's small shapes, the IEEE fallback tile, odd
peak counts, two chunks, the superposition kernels, optimize_settings on the 2,048-point fixture."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402
from metabodecon_rust_b200.lorentzian import superposition_vec_array  # noqa: E402

os.environ["MDB_CHUNK_SPECTRA"] = "3"
n = 8192
x = synth.axis(n)
specs = [Spectrum(x, synth.config3(700 + s, n=n, x=x, integer=bool(s & 1)), (-2.2, 11.8)) for s in range(5)]
specs.append(Spectrum(synth.axis(3001), synth.config3(9, n=3001), (-2.2, 11.8)))
dec = Deconvoluter()
dec.add_ignore_region((4.7, 4.9))
outs = dec.deconvolute_spectra(specs)
print("deconvoluted", [len(o.parameters) for o in outs])
dec2 = Deconvoluter()
dec2.set_moving_average_smoother(5, 7)
dec2.set_detector_only()
dec2.set_analytical_fitter(3)
print("detector-only", len(dec2.deconvolute_spectrum(specs[0]).parameters))
rng = np.random.default_rng(1)
lor = np.stack([rng.uniform(1e-3, 1.0, 777), rng.uniform(1e-7, 1e-5, 777), rng.uniform(0, 10, 777)], axis=1)
lor[5] = [0.0, 1e-6, 5.0]  # leaves the fast-division domain: IEEE tile
print("superposition", float(superposition_vec_array(np.linspace(-2, 12, 10001), lor).sum()))
# the few-ulp form with 16 points per thread (default mode, grids of at least 1 184 x 1 024 points)
big = superposition_vec_array(np.linspace(-2, 12, 1184 * 1024 + 77), lor[:40])
print("superposition, 16 points per thread", float(big[::4099].sum()))
sim = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "sim_01"), 10, 10, (3.339, 3.553))
print("optimize", Deconvoluter().optimize_settings(sim))
