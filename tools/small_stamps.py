#!/usr/bin/env python
"""Development aid: phase clock of the fused small-spectrum kernel (MDB_SMALL_STAMPS=1) on sim_01
for a few smoothing settings.  Run on the GPU box; the library prints the cycle counts to stderr."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MDB_SMALL_STAMPS"] = "1"
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "bruker")
sim = Spectrum.read_bruker(os.path.join(G, "sim_01"), 10, 10, (3.34, 3.56))
for iters, window in [(2, 5), (1, 5), (2, 3), (3, 7)]:
    dec = Deconvoluter()
    dec.set_moving_average_smoother(iters, window)
    print(f"iterations={iters} window={window}", file=sys.stderr, flush=True)
    for _ in range(2):
        try:
            dec.deconvolute_spectrum(sim)
        except Exception as err:  # noqa: BLE001
            print("  ", type(err).__name__, file=sys.stderr, flush=True)
