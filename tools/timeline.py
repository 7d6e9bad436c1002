#!/usr/bin/env python
"""Host-observed timeline of one mdb_deconvolute_spectra call (MDB_TIMELINE): per chunk, when stage A was
queued, when its selected-peak counts had arrived on the host, when stage B was queued and when its
results had arrived.  Shows what a strong-scaled shard (1,250 spectra per GPU) loses to pipeline fill
and drain.  Usage: python tools/timeline.py [--spectra 1250] [--workload config5] [--memory device|pinned|pageable]"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spectra", type=int, default=1250)
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--memory", default="device", choices=["device", "pinned", "pageable"])
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import torch
    from metabodecon_rust_b200 import _lib
    from metabodecon_rust_b200.lorentzian import superposition_vec_array
    lib = _lib.load()
    k, hw, _ = B.WORKLOADS[args.workload]
    x = B.axis(B.N_POINTS)
    S = args.spectra
    ys = np.empty((S, B.N_POINTS))
    for s in range(S):
        p, noise = B.draw_spectrum(s, k, hw)
        ys[s] = superposition_vec_array(x, p, mode="exact") + noise
    if args.memory == "device":
        yt, xt = torch.from_numpy(ys).cuda(), torch.from_numpy(x).cuda()
        xp, yp, mem = xt.data_ptr(), yt.data_ptr(), _lib.MDB_MEM_DEVICE
    elif args.memory == "pinned":
        yt, xt = torch.from_numpy(ys).pin_memory(), torch.from_numpy(x.copy()).pin_memory()
        xp, yp, mem = xt.data_ptr(), yt.data_ptr(), _lib.MDB_MEM_HOST
    else:
        xp, yp, mem = x.ctypes.data, ys.ctypes.data, _lib.MDB_MEM_HOST
    views = (_lib.SpectrumView * S)()
    for s in range(S):
        views[s].chemical_shifts = xp
        views[s].intensities = yp + s * B.N_POINTS * 8
        views[s].len = B.N_POINTS
        views[s].signal_boundaries[0], views[s].signal_boundaries[1] = B.SB
    dec = C.c_void_p()
    assert lib.mdb_deconvoluter_default(C.byref(dec)) == 0

    def call():
        b = C.c_void_p()
        t0 = time.perf_counter()
        assert lib.mdb_deconvolute_spectra(dec, views, S, mem, C.byref(b)) == 0, _lib.last_error()
        dt = time.perf_counter() - t0
        lib.mdb_batch_free(b)
        return dt

    call(); call()
    path = tempfile.mktemp(suffix=".jsonl")
    os.environ["MDB_TIMELINE"] = path
    walls = [call() for _ in range(args.repeat)]
    del os.environ["MDB_TIMELINE"]
    runs = [json.loads(line) for line in open(path)]
    os.unlink(path)
    best = min(range(len(runs)), key=lambda i: runs[i]["total_ms"])
    r = runs[best]
    print(f"# {args.workload}, {S} spectra, {args.memory} inputs: call {1e3 * walls[best]:.2f} ms wall, pipeline {r['total_ms']:.2f} ms "
          f"= {S / (walls[best]):.0f} spectra/s; all walls (ms): {[round(1e3 * w, 2) for w in walls]}")
    print("# host clock: stage A queued | counts on host | stage B queued | results on host;  GPU clock: inputs landed | smoothed | stage A done | stage B starts | stage B done   (ms since the call started)")
    print("# chunk first count |  A queued |    counts |  B queued |   results ||    inputs |  smoothed |    A done |   B start |    B done")
    for i, row in enumerate(r["chunks"]):
        first, count, ta, tc, tb, td = row[:6]
        g = row[6:] if len(row) > 6 else [-1] * 5
        print(f"{i:4d} {first:6d} {count:5d} | {ta:9.2f} | {tc:9.2f} | {tb:9.2f} | {td:9.2f} || " + " | ".join(f"{v:9.2f}" for v in g))
    ch = [row[:6] for row in r["chunks"]]
    if r.get("kernels"):
        from metabodecon_rust_b200._lib import KERNEL_NAMES
        # GPU busy profile: per millisecond, how many kernels of each family were running (events on the launching streams)
        end = max(k[2] for k in r["kernels"])
        print("# kernels on the GPU clock: for every 2 ms window, the summed run time (ms) of the launches of each family that overlap it")
        names = [n for i, n in enumerate(KERNEL_NAMES) if any(k[0] == i for k in r["kernels"])]
        print("#   window  " + " ".join(f"{n[:12]:>12s}" for n in names))
        t = 0.0
        while t < end:
            row = []
            for n in names:
                i = KERNEL_NAMES.index(n)
                row.append(sum(max(0.0, min(k[2], t + 2.0) - max(k[1], t)) for k in r["kernels"] if k[0] == i))
            print(f"# {t:5.0f}-{t + 2:<4.0f} " + " ".join(f"{v:12.2f}" for v in row))
            t += 2.0
        tot = {n: sum(k[2] - k[1] for k in r["kernels"] if k[0] == KERNEL_NAMES.index(n)) for n in names}
        print("# summed launch durations (ms): " + ", ".join(f"{n} {v:.1f}" for n, v in tot.items()))
    steady = [(ch[i + 1][5] - ch[i][5]) / ch[i + 1][1] for i in range(len(ch) // 3, len(ch) - 2) if ch[i + 1][1]]
    if steady:
        per = float(np.median(steady))
        ideal = per * S
        print(f"# steady state: {per * 1e3:.1f} us per spectrum = {1e3 / per:.0f} spectra/s; the call at that rate would take {ideal:.2f} ms; "
              f"fill + drain + host = {r['total_ms'] - ideal:.2f} ms ({100 * (r['total_ms'] - ideal) / r['total_ms']:.1f} %); "
              f"first counts after {ch[0][3]:.2f} ms, first results after {ch[0][5]:.2f} ms")


if __name__ == "__main__":
    main()
