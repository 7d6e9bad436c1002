#!/usr/bin/env python
"""bench.py -- spectra/s deconvolved (2^17 points) on B200, BASELINE.json's headline metric.

One step = one pass of the whole deconvolution hot path (smooth -> detect -> select -> fit -> MSE)
over one batch of synthetic spectra (SURVEY.md §8d, config 5 by default: ~2,000 selected peaks
per spectrum).  Spectra are independent, so N GPUs = N processes, each with its own shard, no
collective on the data path (torch.distributed is used only for the barrier and the max-over-ranks
of the step time).

  value : whole-job spectra/s with the inputs already resident in HBM (C ABI, MDB_MEM_DEVICE)
  e2e   : the same through the C ABI with HOST buffers (pinned), H2D + compute + D2H all timed
  roofline / roofline_hbm : dominant FP64 kernel and the streaming detection kernel, timed live
          with CUDA events on the launching stream (mdb_profile_*), see DESIGN.md
  cpu_baseline : the oracle (C port of the reference path, OpenMP over spectra) on a bounded
          sample of the same workload, rank 0, N=1 only

`--impl reference` times the reference's CPU algorithm (the same oracle port; the Rust reference
cannot be built in this image) on the host cores and prints the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_POINTS = 131072
X_MAX, X_WIDTH = 14.81146, 20.0236139622347
SB = (11.8, -2.2)
WORKLOADS = {
    # name: (true lorentzians K, hw range, description)
    "config5": (3000, (3e-4, 1.5e-3), "synthetic 2^17-pt spectra, K=3000 lorentzians + N(0,300) noise (~2,100 selected peaks)"),
    "config3": (500, (5e-4, 3e-3), "synthetic 2^17-pt spectra, K=500 lorentzians + N(0,300) noise (~510 selected peaks)"),
}
FP64_LANES_PER_SM, N_SM = 64, 148
FP64_INSTR_PER_EVAL = 12  # counted from SASS: sub, mul, add, 8 for the IEEE division, accumulate
SUP_MODES = {"exact": 0, "fast": 1}  # include/mdb200.h MDB_SUPERPOSITION_*
FP64_INSTR_PER_EVAL_FAST = 6  # MDB_SUPERPOSITION_FAST (K7 / K8 only): sub, fma, 3 fma for the reciprocal, fma accumulate
FLOPS_PER_EVAL = 5        # algorithmic: sub, mul, add, div, accumulate


def axis(n):
    i = np.arange(n, dtype=np.float64)
    return X_MAX - i * X_WIDTH / (float(n) - 1.0)


def draw_params(global_index, k, hw_range):
    """(K,3) array of (sfhw, hw2, maxp) with sfhw = A*hw^2, i.e. height A at the maximum."""
    rng = np.random.Generator(np.random.PCG64(20260000 + global_index))
    maxp = rng.uniform(-2.0, 11.6, k)
    hw = np.exp(rng.uniform(np.log(hw_range[0]), np.log(hw_range[1]), k))
    amp = np.exp(rng.uniform(np.log(1e4), np.log(1e7), k))
    return np.ascontiguousarray(np.stack([amp * hw * hw, hw * hw, maxp], axis=1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            try:
                pw.append(float(parts[2]))
            except ValueError:
                pass
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm), "reasons": sorted(reasons),
                "power_w_median": float(np.median(pw)) if pw else None}


def run_reference(args, rank, world):
    """The reference's CPU algorithm (oracle port, OpenMP over spectra) on the host cores."""
    if rank != 0:
        return
    import oracle as O
    k_true, hw_range, desc = WORKLOADS[args.workload]
    cores = O.use_all_cores()
    n_sample = args.cpu_sample or 4 * cores
    x = axis(N_POINTS)
    ys = np.empty((n_sample, N_POINTS))
    for s in range(n_sample):
        p = draw_params(s, k_true, hw_range)
        rng = np.random.Generator(np.random.PCG64(7_000_000 + s))
        ys[s] = O.superposition_vec(x, p, parallel=True) + rng.normal(0.0, 300.0, N_POINTS)
    settings = O.Settings()
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        status, lors, mse, nsel = O.par_deconvolute_spectra(settings, x, ys, SB)
        t1 = time.perf_counter()
        assert status == O.OK
        if it >= args.warmup:
            times.append(t1 - t0)
    ms = 1e3 * sum(times) / len(times)
    value = n_sample / (ms / 1e3)
    sample = f"{n_sample} of the workload's spectra per step, OpenMP over spectra, {cores} threads"
    line = {
        "impl": "reference", "metric": "spectra/s deconvolved (2^17 pts)", "value": value, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "points": N_POINTS, "spectra_per_step": n_sample,
                   "mean_selected_peaks": float(np.mean(nsel)), "settings": "Deconvoluter::default()"},
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "clocks": None,
        "note": "oracle/ C port of the reference path (the Rust reference cannot be built here: no cargo/rustc)",
    }
    emit(line)


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the one JSON line and nothing else: libraries that print there (NCCL's version
    # banner does) are sent to stderr by swapping file descriptor 1 before anything is loaded
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5", choices=sorted(WORKLOADS))
    ap.add_argument("--spectra", type=int, default=2000, help="spectra per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="spectra in the CPU baseline sample (0 = 16 x cores)")
    ap.add_argument("--no-superposition", action="store_true", help="skip the config-4 superposition_vec measurement")
    ap.add_argument("--no-smooth-saturation", action="store_true",
                    help="skip the K1 launch-size sweep (smoothing GB/s at 64 ... 5920 spectra per launch)")
    ap.add_argument("--sup-points", type=int, default=1 << 24, help="config 4: grid points (whole job, sharded over the GPUs)")
    ap.add_argument("--sup-lorentzians", type=int, default=20000, help="config 4: Lorentzians")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small-spectra", action="store_true", help="skip the 2 048-point small-spectrum measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--superposition", default="fast", choices=["fast", "exact"],
                    help="arithmetic of the MSE superposition and superposition_vec (mdb_set_superposition_mode); "
                         "fast is the library's default, exact replays the reference's operators bit for bit")
    ap.add_argument("--host-memory", default="pinned", choices=["pinned", "pageable"],
                    help="host buffers of the e2e measurement (pageable = what NumPy callers hand over)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from metabodecon_rust_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k_true, hw_range, desc = WORKLOADS[args.workload]
    S = args.spectra
    x_np = axis(N_POINTS)
    x_dev = torch.from_numpy(x_np).to(dev)
    y_dev = torch.empty((S, N_POINTS), dtype=torch.float64, device=dev)
    # ---- synthetic batch, generated on the device: clean signal through the library's own
    # superposition kernel, plus N(0, 300) noise
    gen = torch.Generator(device=dev)
    assert lib.mdb_set_superposition_mode(SUP_MODES["exact"]) == 0  # the inputs do not depend on the mode measured
    for s in range(S):
        gi = rank * S + s
        p = torch.from_numpy(draw_params(gi, k_true, hw_range)).to(dev)
        st = lib.mdb_superposition_vec(x_dev.data_ptr(), N_POINTS, p.data_ptr(), k_true, y_dev[s].data_ptr(), _lib.MDB_MEM_DEVICE)
        assert st == 0, _lib.last_error()
        gen.manual_seed(7_000_000 + gi)
        y_dev[s] += torch.randn(N_POINTS, generator=gen, dtype=torch.float64, device=dev) * 300.0
    torch.cuda.synchronize()
    assert lib.mdb_set_superposition_mode(SUP_MODES[args.superposition]) == 0

    dec = C.c_void_p()
    assert lib.mdb_deconvoluter_default(C.byref(dec)) == 0

    def make_views(xp, y_rows_ptr):
        views = (_lib.SpectrumView * S)()
        for s in range(S):
            views[s].chemical_shifts = xp
            views[s].intensities = y_rows_ptr + s * N_POINTS * 8
            views[s].len = N_POINTS
            views[s].signal_boundaries[0], views[s].signal_boundaries[1] = SB
        return views

    dev_views = make_views(x_dev.data_ptr(), y_dev.data_ptr())
    stats = {}
    chunk_max = max(16, min(512, (1 << 23) // N_POINTS))  # chunk_size_for() in csrc/api.cu

    def step(views, memory):
        batch = C.c_void_p()
        st = lib.mdb_deconvolute_spectra(dec, views, S, memory, C.byref(batch))
        assert st == 0, _lib.last_error()
        n_lor = sum(lib.mdb_batch_n_lorentzians(batch, i) for i in range(S))
        n_pk = sum(lib.mdb_batch_n_peaks(batch, i) for i in range(S))
        stats["lorentzians"], stats["peaks"] = n_lor, n_pk
        lib.mdb_batch_free(batch)

    def timed(views, memory, warmup, steps, profile=False):
        for _ in range(warmup):
            step(views, memory)
        if profile:
            lib.mdb_profile_reset()
            lib.mdb_profile_enable(1)
        lib.mdb_reset_kernel_launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(views, memory)
        e1.record()
        barrier()
        if profile:
            lib.mdb_profile_enable(0)
        ms = e0.elapsed_time(e1) / steps
        launches = lib.mdb_kernel_launch_count()
        h2d_b, d2h_b = C.c_uint64(), C.c_uint64()
        lib.mdb_transfer_bytes(C.byref(h2d_b), C.byref(d2h_b))
        stats["h2d_per_step"], stats["d2h_per_step"] = h2d_b.value // steps, d2h_b.value // steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- value: inputs resident in HBM, three chunk pipelines in flight (the product's default)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches = timed(dev_views, _lib.MDB_MEM_DEVICE, args.warmup, args.steps)
    clocks = sampler.stop()
    value = world * S / (ms_dev / 1e3)
    # the same steps in the other arithmetic mode of K7 (reported beside the headline, not as it)
    other_mode = "exact" if args.superposition == "fast" else "fast"
    assert lib.mdb_set_superposition_mode(SUP_MODES[other_mode]) == 0
    ms_other, _ = timed(dev_views, _lib.MDB_MEM_DEVICE, 1, max(2, args.steps // 2))
    assert lib.mdb_set_superposition_mode(SUP_MODES[args.superposition]) == 0

    # ---- per-kernel rooflines: one extra step with the chunk pipeline forced serial
    # (MDB_PIPELINE_DEPTH=1), so that every kernel is alone on the GPU while its CUDA events
    # (recorded on the launching stream inside the library, mdb_profile_*) bracket it
    # and MDB_CHUNK_SPECTRA=256 so that every launch is large enough to fill the GPU on its own
    # (the product's own chunks are smaller because eight of them overlap)
    ROOFLINE_CHUNK = 256
    os.environ["MDB_PIPELINE_DEPTH"] = "1"
    os.environ["MDB_CHUNK_SPECTRA"] = str(ROOFLINE_CHUNK)
    ms_serial, _ = timed(dev_views, _lib.MDB_MEM_DEVICE, 1, 1, profile=True)
    del os.environ["MDB_PIPELINE_DEPTH"]
    del os.environ["MDB_CHUNK_SPECTRA"]
    prof = {}
    for kid, name in enumerate(_lib.KERNEL_NAMES):
        ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
        lib.mdb_profile_read(kid, C.byref(ms), C.byref(n), C.byref(work))
        prof[name] = {"ms": ms.value, "launches": int(n.value), "work": work.value}

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = args.host_memory == "pinned"
        y_host = torch.empty((S, N_POINTS), dtype=torch.float64, pin_memory=pin)
        y_host.copy_(y_dev)
        x_host = torch.from_numpy(x_np.copy())
        if pin:
            x_host = x_host.pin_memory()
        torch.cuda.synchronize()
        host_views = make_views(x_host.data_ptr(), y_host.data_ptr())
        ms_host, _ = timed(host_views, _lib.MDB_MEM_HOST, max(1, args.warmup), args.steps)
        # bytes as counted by the library around its own copies (mdb_transfer_bytes): intensities, the
        # axis once per chunk and descriptors going in; counts, peaks, Lorentzians and MSEs coming out
        e2e = {"value": world * S / (ms_host / 1e3), "unit": "spectra/s", "h2d_bytes_per_step": int(stats["h2d_per_step"]),
               "d2h_bytes_per_step": int(stats["d2h_per_step"]), "ms_per_step": ms_host, "host_memory": args.host_memory}

    # ---- config 4: one superposition_vec over a 2^24-point grid x 20,000 Lorentzians, the grid
    # sharded contiguously over the ranks (strong scaling), parameters replicated, no exchange
    sup = None
    if not args.no_superposition:
        n_all, p4 = args.sup_points, args.sup_lorentzians
        lo, hi = rank * n_all // world, (rank + 1) * n_all // world
        rng4 = np.random.Generator(np.random.PCG64(20260004))
        maxp4 = rng4.uniform(0.0, 10.0, p4)
        hw4 = np.exp(rng4.uniform(np.log(5e-4), np.log(3e-3), p4))
        sf4 = np.exp(rng4.uniform(0.0, np.log(1e4), p4))
        lor4 = torch.from_numpy(np.ascontiguousarray(np.stack([sf4 * hw4, hw4 * hw4, maxp4], axis=1))).to(dev)
        x4 = torch.linspace(-2.2, 11.8, n_all, dtype=torch.float64)[lo:hi].contiguous().to(dev)
        out4 = torch.empty_like(x4)

        def sup_step():
            st = lib.mdb_superposition_vec(x4.data_ptr(), hi - lo, lor4.data_ptr(), p4, out4.data_ptr(), _lib.MDB_MEM_DEVICE)
            assert st == 0, _lib.last_error()

        sup_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            sup_step()
        e1.record()
        barrier()
        ms4 = e0.elapsed_time(e1) / 2
        if world > 1:
            t = torch.tensor([ms4], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms4 = float(t.item())
        sup = {"workload": f"config4: superposition_vec, {n_all} grid points x {p4} Lorentzians, grid sharded over {world} GPU(s)",
               "mode": args.superposition,
               "evals_per_s": n_all * p4 / (ms4 / 1e3), "ms": ms4, "scaling": "strong",
               "checksum": float(out4[:: max(1, (hi - lo) // 1024)].sum().item())}
        # the other arithmetic mode on the same grid: time, and the largest relative difference between the two
        fast_out = out4.clone()
        assert lib.mdb_set_superposition_mode(SUP_MODES[other_mode]) == 0
        sup_step()
        barrier()
        e0.record()
        sup_step()
        e1.record()
        barrier()
        ms4o = e0.elapsed_time(e1)
        assert lib.mdb_set_superposition_mode(SUP_MODES[args.superposition]) == 0
        rel4 = float(((fast_out - out4).abs() / out4.abs().clamp_min(1e-300)).max().item())
        if world > 1:
            t = torch.tensor([ms4o, rel4], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms4o, rel4 = float(t[0].item()), float(t[1].item())
        sup["other_mode"] = {"mode": other_mode, "evals_per_s": n_all * p4 / (ms4o / 1e3), "ms": ms4o,
                             "max_rel_difference_between_modes": rel4}
        del fast_out
        del x4, out4

    # ---- K1 alone at growing launch sizes: the exact-recurrence smoothing runs at chain latency, so
    # its GB/s is set by how many spectra share a launch; it turns bandwidth bound only with
    # thousands of them.  Reuses the batch as input, rank 0 at N = 1 only.
    smooth_sat = None
    if not args.no_smooth_saturation and world == 1:
        smooth_sat = []
        for count in (64, 512, 1480, 2960, 5920, 11840):
            reps = (count + S - 1) // S
            src = y_dev if reps == 1 else y_dev.repeat(reps, 1)
            src = src[:count].contiguous()
            dst = torch.empty_like(src)
            ms_k = C.c_double()
            best = None
            for _ in range(3):
                st = lib.mdb_stage_smooth_batch(src.data_ptr(), N_POINTS, count, N_POINTS, 3, 3, dst.data_ptr(), C.byref(ms_k))
                assert st == 0, _lib.last_error()
                best = ms_k.value if best is None else min(best, ms_k.value)
            smooth_sat.append({"spectra_per_launch": count, "ms": best, "gb_per_s": 16.0 * N_POINTS * count / (best / 1e3) / 1e9})
            del src, dst
        torch.cuda.empty_cache()

    # ---- the small-spectrum path (one fused launch per call, csrc/small_fused.cuh): 2 048-point spectra,
    # the size of the reference's `sim` benchmark set (benches/deconvoluter.rs:8-52).  Latency of one
    # call from host memory through the C ABI, and spectra/s for 2 000 of them in one call, each beside
    # the oracle port on the host cores.  Rank 0 at N = 1 only; not the headline metric.
    small = None
    if not args.no_small_spectra and world == 1:
        import oracle as O   # the checker beside the product, as in the cpu_baseline leg
        import synth
        n_s, count_s = 2048, 2000
        xs = synth.axis(n_s)
        ys_small = np.stack([synth.spectrum(7000 + s, n=n_s, k=30, hw_range=(8e-3, 5e-2), x=xs) for s in range(16)])
        ys_all = np.ascontiguousarray(np.tile(ys_small, (count_s // 16, 1)))
        views_s = (_lib.SpectrumView * count_s)()
        for i in range(count_s):
            views_s[i].chemical_shifts = xs.ctypes.data
            views_s[i].intensities = ys_all[i].ctypes.data
            views_s[i].len = n_s
            views_s[i].signal_boundaries[0], views_s[i].signal_boundaries[1] = SB

        def small_call(k):
            b = C.c_void_p()
            assert lib.mdb_deconvolute_spectra(dec, views_s, k, _lib.MDB_MEM_HOST, C.byref(b)) == 0, _lib.last_error()
            lib.mdb_batch_free(b)

        def median_ms(fn, reps):
            fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return 1e3 * float(np.median(ts))

        lat_ms = median_ms(lambda: small_call(1), 200)
        batch_ms = median_ms(lambda: small_call(count_s), 10)
        O.use_all_cores()
        cpu1_ms = median_ms(lambda: O.deconvolute_spectrum(O.Settings(), xs, ys_all[0], SB), 20)
        t0 = time.perf_counter()
        st_s, *_ = O.par_deconvolute_spectra(O.Settings(), xs, ys_all, SB)
        cpu_batch_ms = 1e3 * (time.perf_counter() - t0)
        assert st_s == O.OK
        small = {"workload": f"{n_s}-point spectra, default settings, host memory through the C ABI",
                 "single_call_ms": lat_ms, "single_call_oracle_ms": cpu1_ms,
                 "spectra_per_s": count_s / (batch_ms / 1e3), "spectra_per_call": count_s,
                 "oracle_spectra_per_s": count_s / (cpu_batch_ms / 1e3)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines from the live per-kernel timings
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    sm_mhz = clocks["sm_mhz"] if clocks else 1965.0
    sm_max = clocks["sm_max_mhz"] if clocks else 1965.0
    fp64_peak_tflops = N_SM * FP64_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12  # FMA = 2 flops, at max clock

    traffic_db = {}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            traffic_db = json.load(fh).get("bytes_per_spectrum", {})
    spectra_per_launch = S / ((S + ROOFLINE_CHUNK - 1) // ROOFLINE_CHUNK)  # launches of the serial roofline pass

    def traffic(name):
        """DRAM bytes per launch: ncu-measured bytes per spectrum (profiles/traffic.json) x spectra per launch."""
        return traffic_db[name] * spectra_per_launch if name in traffic_db else None

    fast = args.superposition == "fast"
    instr_of = {"fit_iter": FP64_INSTR_PER_EVAL, "mse_superposition": FP64_INSTR_PER_EVAL_FAST if fast else FP64_INSTR_PER_EVAL}

    def fp64_roofline(name):
        p = prof[name]
        if p["launches"] == 0 or p["ms"] <= 0:
            return None
        FP64_INSTR_PER_EVAL = instr_of[name]
        evals_per_s = p["work"] / (p["ms"] / 1e3)
        achieved = evals_per_s * FLOPS_PER_EVAL / 1e12
        pipe_rate = N_SM * FP64_LANES_PER_SM * sm_mhz * 1e6
        return {"bound": "fp64", "kernel": name, "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak_tflops, "traffic": traffic(name),
                "peak_source": f"computed: {N_SM} SMs x {FP64_LANES_PER_SM} FP64 lanes x 2 x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP64 entry; "
                               "tools/kbench.cu measures 98.5-99 % of the corresponding instruction rate with DADD/DMUL/DFMA streams, profiles/kbench_r1.txt)",
                "evals_per_s": evals_per_s, "evals_per_launch": p["work"] / p["launches"],
                "ms_per_launch": p["ms"] / p["launches"], "launches": p["launches"],
                "fp64_pipe_util": evals_per_s * FP64_INSTR_PER_EVAL / pipe_rate,
                "fp64_pipe_util_note": f"{FP64_INSTR_PER_EVAL} FP64-pipe instructions per evaluation (SASS) at the median SM clock under load ({sm_mhz:.0f} MHz)"}

    def hbm_roofline(name):
        p = prof[name]
        if p["launches"] == 0 or p["ms"] <= 0:
            return None
        achieved = p["work"] / (p["ms"] / 1e3) / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic(name), "peak_source": hbm_src,
                "bytes_per_launch": p["work"] / p["launches"], "ms_per_launch": p["ms"] / p["launches"],
                "launches": p["launches"]}

    if sup is not None:
        sup["fp64_instr_per_eval"] = instr_of["mse_superposition"]
        sup["fp64_pipe_util"] = sup["evals_per_s"] * instr_of["mse_superposition"] / (world * N_SM * FP64_LANES_PER_SM * sm_mhz * 1e6)
        sup["frac_of_fp64_peak"] = sup["evals_per_s"] * FLOPS_PER_EVAL / 1e12 / (world * fp64_peak_tflops)
    # whole-step view: Lorentzian evaluations of one step (from the library's own work counters of the
    # serial pass) over the pipelined step time -- how close the full pipeline runs to the FP64 pipe
    evals_per_step = prof["fit_iter"]["work"] + prof["mse_superposition"]["work"]
    instr_per_step = sum(prof[k]["work"] * instr_of[k] for k in instr_of)
    overall = {"evals_per_step": evals_per_step, "evals_per_s": world * evals_per_step / (ms_dev / 1e3),
               "fp64_instr_per_step": instr_per_step,
               "fp64_pipe_util": instr_per_step / (ms_dev / 1e3) / (N_SM * FP64_LANES_PER_SM * sm_mhz * 1e6)}
    dominant = max(("mse_superposition", "fit_iter"), key=lambda k: prof[k]["ms"])
    roofline = fp64_roofline(dominant)
    other = fp64_roofline("fit_iter" if dominant == "mse_superposition" else "mse_superposition")

    # ---- CPU baseline on a bounded sample + parity of that sample
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        import oracle as O
        cores = O.use_all_cores()
        n_sample = min(S, args.cpu_sample or 16 * cores)
        ys = y_dev[:n_sample].cpu().numpy()
        O.par_deconvolute_spectra(O.Settings(), x_np, ys[:min(n_sample, cores)], SB)  # warm-up
        t0 = time.perf_counter()
        status, lors, mse, nsel = O.par_deconvolute_spectra(O.Settings(), x_np, ys, SB)
        dt = time.perf_counter() - t0
        cpu = {"value": n_sample / dt, "unit": "spectra/s", "cores": cores, "kind": "port",
               "sample": f"first {n_sample} spectra of the batch, oracle port with OpenMP over spectra, {dt:.1f} s"}
        # parity of the sample: GPU (device-resident path) vs oracle, bit patterns
        sub = (_lib.SpectrumView * n_sample)(*[dev_views[i] for i in range(n_sample)])
        batch = C.c_void_p()
        assert lib.mdb_deconvolute_spectra(dec, sub, n_sample, _lib.MDB_MEM_DEVICE, C.byref(batch)) == 0
        ok = status == O.OK
        mse_rel = 0.0
        for i in range(n_sample):
            k = lib.mdb_batch_n_lorentzians(batch, i)
            got = np.ctypeslib.as_array(C.cast(lib.mdb_batch_lorentzians(batch, i), C.POINTER(C.c_double)), (max(k, 1), 3))[:k]
            ok = ok and k == len(lors[i]) and np.array_equal(got.view(np.uint64), lors[i].view(np.uint64))
            ok = ok and lib.mdb_batch_n_peaks(batch, i) == nsel[i]
            mse_rel = max(mse_rel, abs(lib.mdb_batch_mse(batch, i) - mse[i]) / abs(mse[i]))
        lib.mdb_batch_free(batch)
        # peak counts and Lorentzian parameters: identical bit patterns in both modes; the MSE is
        # bit-identical in exact mode and within 1e-9 relative (north_star) in fast mode
        parity = {"spectra": n_sample, "peak_sets_and_lorentzians_bit_exact_vs_oracle": bool(ok),
                  "mse_max_rel_err_vs_oracle": mse_rel, "mse_tolerance": 1e-9 if fast else 0.0,
                  "superposition_mode": args.superposition, "pass": bool(ok and mse_rel <= (1e-9 if fast else 0.0))}

    line = {
        "metric": "spectra/s deconvolved (2^17 pts)", "value": value, "unit": "spectra/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "points": N_POINTS, "spectra_per_gpu_per_step": S,
                   "mean_selected_peaks": stats["peaks"] / S, "mean_lorentzians": stats["lorentzians"] / S,
                   "settings": "Deconvoluter::default()", "superposition_mode": args.superposition, "parallelism": f"spectra sharded over {world} GPU(s), no collective",
                   "l2": f"inputs {S * N_POINTS * 8 / 2**20:.0f} MiB per GPU per step, larger than the 126 MB L2"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "roofline_other_fp64": other, "pipeline_fp64": overall,
        "other_superposition_mode": {"mode": other_mode, "value": world * S / (ms_other / 1e3), "unit": "spectra/s", "ms_per_step": ms_other},
        "roofline_hbm": {"detect": hbm_roofline("detect"), "smooth": hbm_roofline("smooth")},
        "kernel_ms_serial_step": {k: v["ms"] for k, v in prof.items() if v["launches"]},
        "serial_step_ms": ms_serial, "roofline_pass": f"one extra step, chunks of {ROOFLINE_CHUNK} spectra, one chunk at a time",
        "superposition_vec": sup,
        "small_spectra": small,
        "smooth_launch_size_sweep": None if smooth_sat is None else
        [dict(e, frac_of_hbm_peak=e["gb_per_s"] / hbm_peak) for e in smooth_sat],
        "cpu_baseline": cpu, "parity_sample": parity,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
