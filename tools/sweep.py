#!/usr/bin/env python
"""In-process sweeps of the pipeline's tuning knobs on device-resident synthetic batches (one data
generation, many settings): chunk size, pipeline depth, stream priorities, refinement arithmetic.
Prints one line per setting: spectra/s (median of the timed steps).
  python tools/sweep.py [--workload config5] [--spectra 2000] [--steps 3] [--set NAME=V,NAME=V ...]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

DEFAULT_SETS = {
    "config5": ["", "MDB_STREAM_PRIORITIES=0", "MDB_CHUNK_SPECTRA=48", "MDB_CHUNK_SPECTRA=64", "MDB_CHUNK_SPECTRA=78",
                "MDB_CHUNK_SPECTRA=96", "MDB_CHUNK_SPECTRA=128", "MDB_CHUNK_SPECTRA=156", "MDB_PIPELINE_DEPTH=4", "MDB_PIPELINE_DEPTH=12",
                "FIT=corrected", "FIT=ulp", "SUP=exact", "SUP=exact,FIT=corrected"],
    "config3": ["", "MDB_STREAM_PRIORITIES=0", "MDB_TARGET_EVALS=0.9e10", "MDB_TARGET_EVALS=3.6e10", "MDB_PIPELINE_DEPTH=4",
                "MDB_PIPELINE_DEPTH=12", "FIT=corrected", "SUP=exact"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--spectra", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--set", action="append", default=None)
    args = ap.parse_args()
    import torch
    from metabodecon_rust_b200 import _lib
    lib = _lib.load()
    k, hw, _ = B.WORKLOADS[args.workload]
    x = B.axis(B.N_POINTS)
    S = args.spectra
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x).to(dev)
    y = torch.empty((S, B.N_POINTS), dtype=torch.float64, device=dev)
    for s in range(S):
        p, noise = B.draw_spectrum(s, k, hw)
        pd = torch.from_numpy(p).to(dev)
        assert lib.mdb_superposition_vec_mode(xd.data_ptr(), B.N_POINTS, pd.data_ptr(), k, y[s].data_ptr(), 1, 0) == 0
        y[s] += torch.from_numpy(noise).to(dev)
    views = (_lib.SpectrumView * S)()
    for s in range(S):
        views[s].chemical_shifts = xd.data_ptr()
        views[s].intensities = y.data_ptr() + s * B.N_POINTS * 8
        views[s].len = B.N_POINTS
        views[s].signal_boundaries[0], views[s].signal_boundaries[1] = B.SB
    dec = C.c_void_p()
    assert lib.mdb_deconvoluter_default(C.byref(dec)) == 0

    def step():
        b = C.c_void_p()
        assert lib.mdb_deconvolute_spectra(dec, views, S, 1, C.byref(b)) == 0, _lib.last_error()
        lib.mdb_batch_free(b)

    for setting in (args.set or DEFAULT_SETS[args.workload]):
        env = dict(kv.split("=") for kv in setting.split(",") if kv)
        fit = {"exact": 0, "corrected": 1, "ulp": 2}[env.pop("FIT", "exact")]
        sup = {"exact": 0, "fast": 1}[env.pop("SUP", "fast")]
        assert lib.mdb_deconvoluter_set_fit_arithmetic(dec, fit) == 0
        assert lib.mdb_deconvoluter_set_superposition_mode(dec, sup) == 0
        for kk, v in env.items():
            os.environ[kk] = v
        recreate = any(k in env for k in ("MDB_STREAM_PRIORITIES", "MDB_FIT_PRIORITY"))
        if recreate:
            lib.mdb_release_workspaces()  # streams are created with the pooled stream sets
        step()
        ts = []
        for _ in range(args.steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        for kk in env:
            del os.environ[kk]
        if recreate:
            lib.mdb_release_workspaces()
        print(f"{args.workload} {S} spectra | {setting or 'default':40s} | {S / (np.median(ts) / 1e3):9.0f} spectra/s | ms {[round(t, 1) for t in ts]}", flush=True)


if __name__ == "__main__":
    main()
