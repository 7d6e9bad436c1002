#!/usr/bin/env python
"""Host- and GPU-clock stamps of ONE default blood_01 deconvolution (MDB_TIMELINE), to see where a 2.3 ms
single-spectrum call spends its time outside the kernels.  Run on the GPU box."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metabodecon_rust_b200 import Deconvoluter, Spectrum  # noqa: E402

blood = Spectrum.read_bruker(os.path.join(ROOT, "tests", "golden", "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
dec = Deconvoluter()
for _ in range(5):
    dec.deconvolute_spectrum(blood)
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    dec.deconvolute_spectrum(blood)
    ts.append(time.perf_counter() - t0)
print("wall ms (python mirror): median %.3f min %.3f" % (1e3 * sorted(ts)[len(ts) // 2], 1e3 * min(ts)))
path = tempfile.mktemp(suffix=".jsonl")
os.environ["MDB_TIMELINE"] = path
for _ in range(3):
    dec.deconvolute_spectrum(blood)
del os.environ["MDB_TIMELINE"]
with open(path) as fh:
    last = json.loads(fh.readlines()[-1])
print(json.dumps(last)[:3000])
