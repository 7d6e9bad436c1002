#!/usr/bin/env python
"""bench.py -- spectra/s deconvolved (2^17 points) on B200, BASELINE.json's headline metric.

Headline workload = BASELINE.json config 5 / the north-star batch: 10,000 synthetic 2^17-point
spectra (~2,100 selected peaks each), default settings.  One step = one pass of the whole
deconvolution hot path (smooth -> detect -> select -> fit -> MSE) over that batch.  With N GPUs the
SAME 10,000 spectra are cut into N contiguous shards, one process per GPU (strong scaling); spectra
are independent, so there is no collective on the data path (torch.distributed carries only the
barriers, the max-over-ranks of the step time and the gather that feeds the one-call leg).

  value        whole-job spectra/s, inputs resident in HBM (C ABI, MDB_MEM_DEVICE)
  e2e          the same through the C ABI with pinned HOST buffers: H2D + compute + D2H timed
  e2e_pageable the same with pageable host rows (what a drop-in caller holds)
  e2e_one_call ONE mdb_deconvolute_spectra call from ONE process over all 10,000 host spectra with
               mdb_set_device_count(N): the in-process sharder a par_deconvolute_spectra caller gets
  weak_scaling 1,250 spectra per GPU at every N (extra key)
  config3 / superposition_vec  BASELINE.json configs 3 and 4 (extra keys), device-resident and e2e
  roofline*    per-kernel rooflines from CUDA events on the launching streams (mdb_profile_*)
  cpu_baseline the oracle (C port of the reference path, OpenMP over spectra) on a bounded sample
  parity_sample / parity_per_rank   GPU vs oracle, bit patterns, float AND integer-valued inputs

Inputs: parameters and noise of spectrum s come from NumPy Generator(PCG64(20260000 + s)) on the
HOST (draw order maxp[K], hw[K], A[K], noise[N]), the clean signal from an exact
ordered superposition (the library's exact kernel here, the oracle's superposition_vec in
`--impl reference`: bit-identical, tests/test_gpu_parity.py), y = signal + noise: both arms see the
same arrays.

`--impl reference` times the reference's CPU algorithm (the oracle port; the Rust reference cannot
be built in this image) on the host cores and prints the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

# read by CUDA when the context is created (i.e. before torch touches the GPU): the library's
# pipeline wants its ~10 streams on separate hardware queues (metabodecon_rust_b200/_lib.py)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
FP64_INSTR_PER_EVAL_FAST = 5.25  # MDB_SUPERPOSITION_FAST (K7 / K8 only), four Lorentzians behind one reciprocal (lorentz_quad_ulp):
# 4 sub, 4 fma, 3 + 6 mul / fma for the common denominator and numerator, 3 fma for the reciprocal, 1 fma accumulate = 21 per 4 (SASS)
FLOPS_PER_EVAL = 5        # algorithmic: sub, mul, add, div, accumulate


def axis(n):
    i = np.arange(n, dtype=np.float64)
    return X_MAX - i * X_WIDTH / (float(n) - 1.0)


def draw_spectrum(global_index, k, hw_range, n=N_POINTS):
    """Parameters ((K,3): sfhw = A*hw^2, hw2, maxp) and noise of synthetic spectrum `global_index`:
    one PCG64 stream per spectrum, draw order maxp[K], hw[K], A[K], noise[N]."""
    rng = np.random.Generator(np.random.PCG64(20260000 + global_index))
    maxp = rng.uniform(-2.0, 11.6, k)
    hw = np.exp(rng.uniform(np.log(hw_range[0]), np.log(hw_range[1]), k))
    amp = np.exp(rng.uniform(np.log(1e4), np.log(1e7), k))
    noise = rng.normal(0.0, 300.0, n)
    return np.ascontiguousarray(np.stack([amp * hw * hw, hw * hw, maxp], axis=1)), noise


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


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


def workload_name(args):
    k_true, hw_range, desc = WORKLOADS[args.workload]
    return (f"{args.workload}: batch of {args.total_spectra} {desc}, sharded over the GPUs (strong scaling)")


def workload_config(args, world):
    """The `config` object of the JSON line: a pure description of the workload, IDENTICAL for the GPU arm and for
    --impl reference at the same --gpus (measured statistics live in `workload_stats`, what a step of the
    reference arm actually timed in its `cpu_baseline.sample`)."""
    total = args.total_spectra
    per_gpu = total // world  # rank 0's shard: [0, total // world)
    return {"workload": workload_name(args), "points": N_POINTS, "total_spectra": total,
            "spectra_per_gpu_per_step": per_gpu, "settings": "Deconvoluter::default()",
            "superposition_mode": args.superposition,
            "parallelism": f"the {total} spectra cut into {world} contiguous shard(s), one process per GPU, no collective",
            "inputs": "host PCG64 parameters + noise per spectrum, exact ordered superposition for the signal (the same arrays for the GPU arm, its cpu_baseline and --impl reference)",
            "l2": f"inputs {per_gpu * N_POINTS * 8 / 2**20:.0f} MiB per GPU per step, larger than the 126 MB L2"}


def sample_indices(total, count):
    """`count` spectrum indices spread over [0, total): the sample is drawn across the batch."""
    count = max(1, min(count, total))
    return sorted({int(i) for i in np.linspace(0, total - 1, count)})


def run_reference(args, rank, world):
    """The reference's CPU algorithm (oracle port, OpenMP over spectra) on the host cores, on a
    bounded sample of the SAME workload: spectra drawn across the 10,000-spectra batch, same PCG64
    parameters and noise, signal from the oracle's ordered superposition (bit-identical to the
    exact GPU kernel that makes the GPU arm's inputs)."""
    if rank != 0:
        return
    import oracle as O
    k_true, hw_range, desc = WORKLOADS[args.workload]
    cores = O.use_all_cores()
    n_sample = args.cpu_sample or 4 * cores
    x = axis(N_POINTS)
    idx = sample_indices(args.total_spectra, n_sample)
    ys = np.empty((len(idx), N_POINTS))
    for j, gi in enumerate(idx):
        p, noise = draw_spectrum(gi, k_true, hw_range)
        ys[j] = O.superposition_vec(x, p, parallel=True) + noise
    settings = O.Settings()
    times = []
    nsel = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        status, lors, mse, nsel = O.par_deconvolute_spectra(settings, x, ys, SB)
        t1 = time.perf_counter()
        assert status == O.OK
        if it >= args.warmup:
            times.append(t1 - t0)
    ms = 1e3 * sum(times) / len(times)
    value = len(idx) / (ms / 1e3)
    sample = (f"{len(idx)} spectra drawn across the {args.total_spectra}-spectra batch per step, "
              f"OpenMP over spectra, {cores} threads, gcc -O3 -ffp-contract=off")
    line = {
        "impl": "reference", "metric": "spectra/s deconvolved (2^17 pts)", "value": value, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "workload_stats": {"spectra_per_step_timed": len(idx), "mean_selected_peaks": float(np.mean(nsel))},
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
    ap.add_argument("--total-spectra", type=int, default=10000, help="spectra in the batch (the whole job, sharded over the GPUs)")
    ap.add_argument("--weak-spectra", type=int, default=1250, help="spectra per GPU of the weak-scaling extra key")
    ap.add_argument("--cpu-sample", type=int, default=0, help="spectra in the CPU baseline sample (0 = 8 x cores)")
    ap.add_argument("--no-superposition", action="store_true", help="skip the config-4 superposition_vec measurement")
    ap.add_argument("--no-smooth-saturation", action="store_true",
                    help="skip the K1 launch-size sweep (smoothing GB/s at 64 ... 11840 spectra per launch)")
    ap.add_argument("--sup-points", type=int, default=1 << 24, help="config 4: grid points (whole job, sharded over the GPUs)")
    ap.add_argument("--sup-lorentzians", type=int, default=20000, help="config 4: Lorentzians")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small-spectra", action="store_true", help="skip the 2 048-point small-spectrum measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-one-call", action="store_true")
    ap.add_argument("--superposition", default="fast", choices=["fast", "exact"],
                    help="arithmetic of the MSE superposition and superposition_vec (mdb_deconvoluter_set_superposition_mode); "
                         "fast is the library's default, exact replays the reference's operators bit for bit")
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
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # host-side waits that must not occupy the GPUs
    lib = _lib.load()
    t_start = time.perf_counter()

    def log(msg):
        if rank == 0:
            print(f"[bench {time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

    def max_over_ranks(values):
        if world == 1:
            return [float(v) for v in values]
        t = torch.tensor(list(values), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def min_over_ranks(values):
        return [-v for v in max_over_ranks([-float(v) for v in values])]

    # ---- measured FP64 instruction rate of this GPU (the denominator of every fp64_pipe_util below)
    dfma_rate, dadd_rate = C.c_double(), C.c_double()
    assert lib.mdb_measure_fp64_rate(C.byref(dfma_rate), C.byref(dadd_rate)) == 0, _lib.last_error()
    fp64_rate = dfma_rate.value

    x_np = axis(N_POINTS)
    x_dev = torch.from_numpy(x_np).to(dev)
    pool = ThreadPoolExecutor(max_workers=max(2, min(32, host_threads() // max(1, world))))

    def generate(workload, first, count):
        """Device tensor (count, N) of synthetic spectra [first, first + count) of `workload`."""
        k_true, hw_range, _ = WORKLOADS[workload]
        y = torch.empty((count, N_POINTS), dtype=torch.float64, device=dev)
        block = 256
        for b0 in range(0, count, block):
            b1 = min(count, b0 + block)
            drawn = list(pool.map(lambda gi: draw_spectrum(gi, k_true, hw_range), range(first + b0, first + b1)))
            noise = torch.from_numpy(np.stack([d[1] for d in drawn])).to(dev)
            for j, (p, _) in enumerate(drawn):
                pd = torch.from_numpy(p).to(dev)
                st = lib.mdb_superposition_vec_mode(x_dev.data_ptr(), N_POINTS, pd.data_ptr(), k_true, y[b0 + j].data_ptr(),
                                                    _lib.MDB_MEM_DEVICE, SUP_MODES["exact"])
                assert st == 0, _lib.last_error()
            y[b0:b1] += noise
        torch.cuda.synchronize()
        return y

    def new_deconvoluter(mode):
        d = C.c_void_p()
        assert lib.mdb_deconvoluter_default(C.byref(d)) == 0, _lib.last_error()
        assert lib.mdb_deconvoluter_set_superposition_mode(d, SUP_MODES[mode]) == 0
        return d

    other_mode = "exact" if args.superposition == "fast" else "fast"
    dec = new_deconvoluter(args.superposition)
    dec_other = new_deconvoluter(other_mode)

    def make_views(xp, y_ptr, count, stride_bytes=N_POINTS * 8):
        views = (_lib.SpectrumView * count)()
        for s in range(count):
            views[s].chemical_shifts = xp
            views[s].intensities = y_ptr + s * stride_bytes
            views[s].len = N_POINTS
            views[s].signal_boundaries[0], views[s].signal_boundaries[1] = SB
        return views

    stats = {}

    def run_batch(d, views, count, memory, export=None):
        """One mdb_deconvolute_spectra call; returns (n_lorentzians, n_peaks) totals.  export: indices
        whose Lorentzians / peak counts / MSE are copied out (parity checks)."""
        batch = C.c_void_p()
        st = lib.mdb_deconvolute_spectra(d, views, count, memory, C.byref(batch))
        assert st == 0, _lib.last_error()
        tl, tp = C.c_size_t(), C.c_size_t()
        lib.mdb_batch_totals(batch, C.byref(tl), C.byref(tp))
        out = None
        if export is not None:
            out = []
            for i in export:
                k = lib.mdb_batch_n_lorentzians(batch, i)
                got = np.zeros((0, 3))
                if k:
                    got = np.ctypeslib.as_array(C.cast(lib.mdb_batch_lorentzians(batch, i), C.POINTER(C.c_double)), (k, 3)).copy()
                out.append((got, int(lib.mdb_batch_n_peaks(batch, i)), float(lib.mdb_batch_mse(batch, i))))
        lib.mdb_batch_free(batch)
        return tl.value, tp.value, out

    def timed(d, views, count, memory, warmup, steps, profile=False, all_ranks=True):
        """ms per step (max over ranks when all_ranks), kernel launches of the timed region; byte counts in stats."""
        for _ in range(warmup):
            run_batch(d, views, count, memory)
        if profile:
            lib.mdb_profile_reset()
            lib.mdb_profile_enable(1)
        lib.mdb_reset_kernel_launch_count()
        if all_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            stats["lorentzians"], stats["peaks"], _ = run_batch(d, views, count, memory)
        e1.record()
        if all_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        if profile:
            lib.mdb_profile_enable(0)
        ms = e0.elapsed_time(e1) / steps
        launches = lib.mdb_kernel_launch_count()  # kernels launched inside the timed region (all `steps` steps)
        h2d_b, d2h_b = C.c_uint64(), C.c_uint64()
        lib.mdb_transfer_bytes(C.byref(h2d_b), C.byref(d2h_b))
        stats["h2d_per_step"], stats["d2h_per_step"] = h2d_b.value // steps, d2h_b.value // steps
        if all_ranks:
            ms = max_over_ranks([ms])[0]
        return ms, launches

    def sum_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # =================================================================================== headline
    TOTAL = args.total_spectra
    lo, hi = rank * TOTAL // world, (rank + 1) * TOTAL // world
    S = hi - lo
    log(f"generating {S} of {TOTAL} {args.workload} spectra on rank 0 (and likewise on the other ranks)")
    y_dev = generate(args.workload, lo, S)
    dev_views = make_views(x_dev.data_ptr(), y_dev.data_ptr(), S)

    # ---- value: inputs resident in HBM
    log("value leg (device-resident)")
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches = timed(dec, dev_views, S, _lib.MDB_MEM_DEVICE, args.warmup, args.steps)
    clocks = sampler.stop()
    value = TOTAL / (ms_dev / 1e3)
    total_peaks = sum_over_ranks(stats["peaks"])
    total_lor = sum_over_ranks(stats["lorentzians"])
    launches_all = sum_over_ranks(launches)
    # the same steps in the other arithmetic mode of K7 (reported beside the headline, not as it)
    ms_other, _ = timed(dec_other, dev_views, S, _lib.MDB_MEM_DEVICE, 1, 2)

    # ---- per-kernel rooflines: one extra step over (at most) 2,000 spectra with the chunk pipeline
    # forced serial (MDB_PIPELINE_DEPTH=1) and 256-spectra chunks, so that every kernel is alone on
    # the GPU and fills it while its CUDA events (recorded on the launching stream inside the library,
    # mdb_profile_*) bracket it
    ROOFLINE_CHUNK = 256
    n_roof = min(S, 2000)
    os.environ["MDB_PIPELINE_DEPTH"] = "1"
    os.environ["MDB_CHUNK_SPECTRA"] = str(ROOFLINE_CHUNK)
    ms_serial, _ = timed(dec, dev_views, n_roof, _lib.MDB_MEM_DEVICE, 1, 1, profile=True)
    del os.environ["MDB_PIPELINE_DEPTH"]
    del os.environ["MDB_CHUNK_SPECTRA"]
    prof = {}
    for kid, name in enumerate(_lib.KERNEL_NAMES):
        ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
        lib.mdb_profile_read(kid, C.byref(ms), C.byref(n), C.byref(work))
        prof[name] = {"ms": ms.value, "launches": int(n.value), "work": work.value}
    # the pipelined (product) run over the same n_roof spectra, for the whole-step FP64 view
    ms_roof_pipelined, _ = timed(dec, dev_views, n_roof, _lib.MDB_MEM_DEVICE, 1, 3)

    # ---- weak-scaling extra key: a fixed number of spectra per GPU at every N
    W = min(S, args.weak_spectra)
    ms_weak, _ = timed(dec, dev_views, W, _lib.MDB_MEM_DEVICE, 1, 5)
    weak = {"spectra_per_gpu": W, "value": world * W / (ms_weak / 1e3), "unit": "spectra/s", "ms_per_step": ms_weak, "scaling": "weak"}

    # ---- variant B of SURVEY 8d: the same spectra rounded to integers (Bruker-like; ties decide peaks)
    n_int = min(S, args.weak_spectra)
    y_int = torch.round(y_dev[:n_int])
    int_views = make_views(x_dev.data_ptr(), y_int.data_ptr(), n_int)
    ms_int, _ = timed(dec, int_views, n_int, _lib.MDB_MEM_DEVICE, 1, 3)
    integer_variant = {"spectra_per_gpu": n_int, "value": world * n_int / (ms_int / 1e3), "unit": "spectra/s", "ms_per_step": ms_int,
                       "mean_selected_peaks": stats["peaks"] / n_int}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = e2e_pageable = one_call = None
    y_host = None
    if not args.no_e2e:
        log("e2e leg (pinned host buffers)")
        y_host = torch.empty((S, N_POINTS), dtype=torch.float64, pin_memory=True)
        y_host.copy_(y_dev)
        x_host = torch.from_numpy(x_np.copy()).pin_memory()
        torch.cuda.synchronize()
        host_views = make_views(x_host.data_ptr(), y_host.data_ptr(), S)
        ms_host, _ = timed(dec, host_views, S, _lib.MDB_MEM_HOST, min(2, max(1, args.warmup)), args.steps)
        # bytes as counted by the library around its own copies (mdb_transfer_bytes): intensities, the
        # axis once per chunk and descriptors going in; counts, peaks, Lorentzians and MSEs coming out
        e2e = {"value": TOTAL / (ms_host / 1e3), "unit": "spectra/s",
               "h2d_bytes_per_step": int(sum_over_ranks(stats["h2d_per_step"])),
               "d2h_bytes_per_step": int(sum_over_ranks(stats["d2h_per_step"])), "ms_per_step": ms_host, "host_memory": "pinned"}
        # pageable rows: what a drop-in caller holds (Spectrum owns Arc<[f64]>, spectrum/spectrum.rs:101-116)
        log("e2e leg (pageable host buffers)")
        y_page = np.empty((S, N_POINTS), dtype=np.float64)
        torch.from_numpy(y_page).copy_(y_host)
        page_views = make_views(x_np.ctypes.data, y_page.ctypes.data, S)
        ms_page, _ = timed(dec, page_views, S, _lib.MDB_MEM_HOST, 1, 3)
        e2e_pageable = {"value": TOTAL / (ms_page / 1e3), "unit": "spectra/s", "ms_per_step": ms_page, "host_memory": "pageable",
                        "steps": 3, "fraction_of_pinned": ms_host / ms_page}
        del y_page, page_views

    # ---- one call, all GPUs: rank 0 holds the whole batch in pinned host memory and makes ONE
    # mdb_deconvolute_spectra call with mdb_set_device_count(world); the other ranks wait on the host
    one_call_parity = None
    if not args.no_e2e and not args.no_one_call and world > 1:
        log("one-call leg: gathering the batch on rank 0")
        sizes = [(r + 1) * TOTAL // world - r * TOTAL // world for r in range(world)]
        if len(set(sizes)) == 1:
            parts = [torch.empty((sizes[r], N_POINTS), dtype=torch.float64, device=dev) for r in range(world)] if rank == 0 else None
            dist.gather(y_dev, parts, dst=0)
        else:  # unequal shards: one broadcast per rank
            parts = []
            for r in range(world):
                t = y_dev if r == rank else torch.empty((sizes[r], N_POINTS), dtype=torch.float64, device=dev)
                dist.broadcast(t, src=r)
                if rank == 0:
                    parts.append(t)
        torch.cuda.synchronize()
        if rank == 0:
            y_all = torch.empty((TOTAL, N_POINTS), dtype=torch.float64, pin_memory=True)
            off = 0
            for r in range(world):
                y_all[off:off + sizes[r]].copy_(parts[r])
                off += sizes[r]
            torch.cuda.synchronize()
            del parts
            torch.cuda.empty_cache()
        host_barrier()  # the other ranks now sit in the next host barrier, their GPUs idle
        if rank == 0:
            all_views = make_views(x_host.data_ptr(), y_all.data_ptr(), TOTAL)
            assert lib.mdb_set_device_count(world) == 0
            try:
                for _ in range(2):
                    run_batch(dec, all_views, TOTAL, _lib.MDB_MEM_HOST)  # creates contexts / workspaces on every GPU
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    run_batch(dec, all_views, TOTAL, _lib.MDB_MEM_HOST)
                    ts.append(time.perf_counter() - t0)
                ms_one = 1e3 * float(np.median(ts))
                # parity of the in-process sharder: 16 spectra of EVERY device's shard against the oracle
                import oracle as O
                O.use_all_cores()
                idx = sorted({r * TOTAL // world + int(j) for r in range(world)
                              for j in np.linspace(0, sizes[r] - 1, min(16, sizes[r]))})
                _, _, got = run_batch(dec, all_views, TOTAL, _lib.MDB_MEM_HOST, export=idx)
                ys = y_all[idx].numpy()
                status, lors, mse, nsel = O.par_deconvolute_spectra(O.Settings(), x_np, ys, SB)
                ok = status == O.OK
                mse_rel = 0.0
                for (g, npk, m), want, wn, wm in zip(got, lors, nsel, mse):
                    ok = ok and g.shape == want.shape and np.array_equal(g.view(np.uint64), want.view(np.uint64)) and npk == wn
                    mse_rel = max(mse_rel, abs(m - wm) / abs(wm))
                one_call_parity = {"spectra": len(idx), "per_device": 16, "peak_sets_and_lorentzians_bit_exact_vs_oracle": bool(ok),
                                   "mse_max_rel_err_vs_oracle": mse_rel}
                one_call = {"value": TOTAL / (ms_one / 1e3), "unit": "spectra/s", "ms_per_call": ms_one, "devices": world,
                            "timing": "host wall clock around ONE C-ABI call (median of 3), pinned host rows, all copies inside",
                            "fraction_of_torchrun_e2e": (TOTAL / (ms_one / 1e3)) / e2e["value"], "parity": one_call_parity}
            finally:
                assert lib.mdb_set_device_count(1) == 0
            del y_all, all_views
        host_barrier()
    elif not args.no_e2e and world == 1:
        one_call = {"value": e2e["value"], "unit": "spectra/s", "devices": 1,
                    "note": "with one GPU the one-call leg IS the e2e leg (one process, one C-ABI call per step)"}
    del y_host
    torch.cuda.empty_cache()

    # ---- per-rank parity: 16 spectra of this rank's shard (8 float + the same 8 rounded to integers)
    # against the oracle, bit patterns; every rank checks its own GPU's results
    parity_rank = None
    if not args.no_cpu_baseline:
        import oracle as O
        O.use_cores(max(1, host_threads() // world))
        idx = sample_indices(min(S, n_int), 8)  # spread over the part of the shard both variants cover
        fast = args.superposition == "fast"

        def check(y_rows, views, count):
            _, _, got = run_batch(dec, views, count, _lib.MDB_MEM_DEVICE, export=idx)
            status, lors, mse, nsel = O.par_deconvolute_spectra(O.Settings(), x_np, y_rows[idx].cpu().numpy(), SB)
            ok, rel = status == O.OK, 0.0
            for (g, npk, m), want, wn, wm in zip(got, lors, nsel, mse):
                ok = ok and g.shape == want.shape and np.array_equal(g.view(np.uint64), want.view(np.uint64)) and npk == wn
                rel = max(rel, abs(m - wm) / abs(wm))
            return ok, rel

        ok_f, rel_f = check(y_dev, dev_views, S)
        ok_i, rel_i = check(y_int, int_views, n_int)
        ok_all = min_over_ranks([1.0 if (ok_f and ok_i) else 0.0])[0] == 1.0
        rel_all = max_over_ranks([max(rel_f, rel_i)])[0]
        parity_rank = {"spectra_per_rank": 2 * len(idx), "variants": "float and integer-rounded", "ranks": world,
                       "peak_sets_and_lorentzians_bit_exact_vs_oracle_on_every_rank": bool(ok_all),
                       "mse_max_rel_err_vs_oracle": rel_all, "mse_tolerance": 1e-9 if fast else 0.0,
                       "pass": bool(ok_all and rel_all <= (1e-9 if fast else 0.0))}

    # ================================================================== config 3 (extra key)
    config3 = None
    if not args.no_config3 and args.workload == "config5":
        log("config 3")
        config3 = {"workload": "config3: " + WORKLOADS["config3"][2] + ", sharded over the GPUs (strong scaling)", "batches": []}
        tot3 = 4000
        lo3, hi3 = rank * tot3 // world, (rank + 1) * tot3 // world
        y3 = generate("config3", lo3, hi3 - lo3)
        v3 = make_views(x_dev.data_ptr(), y3.data_ptr(), hi3 - lo3)
        y3_host = torch.empty((hi3 - lo3, N_POINTS), dtype=torch.float64, pin_memory=True)
        y3_host.copy_(y3)
        x3_host = torch.from_numpy(x_np.copy()).pin_memory()
        torch.cuda.synchronize()
        v3h = make_views(x3_host.data_ptr(), y3_host.data_ptr(), hi3 - lo3)
        y3_page = y3_host.numpy().copy()
        v3p = make_views(x_np.ctypes.data, y3_page.ctypes.data, hi3 - lo3)
        for total in (1000, 4000):
            cnt = (rank + 1) * total // world - rank * total // world  # the first `cnt` spectra of this rank's block
            ms3, _ = timed(dec, v3, cnt, _lib.MDB_MEM_DEVICE, 2, 5)
            peaks3 = sum_over_ranks(stats["peaks"])
            ms3h, _ = timed(dec, v3h, cnt, _lib.MDB_MEM_HOST, 1, 5)
            ms3p, _ = timed(dec, v3p, cnt, _lib.MDB_MEM_HOST, 1, 3)
            config3["batches"].append({"total_spectra": total, "value": total / (ms3 / 1e3), "ms_per_step": ms3,
                                       "e2e": total / (ms3h / 1e3), "e2e_ms_per_step": ms3h, "e2e_fraction_of_value": ms3 / ms3h,
                                       "e2e_pageable": total / (ms3p / 1e3), "e2e_pageable_fraction_of_pinned": ms3h / ms3p,
                                       "mean_selected_peaks": peaks3 / total, "unit": "spectra/s"})
        # per-kernel view of config 3 (serial pass over this rank's 4,000 / N spectra) and the pipelined FP64 view
        os.environ["MDB_PIPELINE_DEPTH"] = "1"
        os.environ["MDB_CHUNK_SPECTRA"] = str(ROOFLINE_CHUNK)
        timed(dec, v3, hi3 - lo3, _lib.MDB_MEM_DEVICE, 0, 1, profile=True)
        del os.environ["MDB_PIPELINE_DEPTH"]
        del os.environ["MDB_CHUNK_SPECTRA"]
        prof3 = {}
        for kid, name in enumerate(_lib.KERNEL_NAMES):
            ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
            lib.mdb_profile_read(kid, C.byref(ms), C.byref(n), C.byref(work))
            prof3[name] = {"ms": ms.value, "launches": int(n.value), "work": work.value}
        config3["_prof"] = prof3
        config3["_ms_full"] = config3["batches"][-1]["ms_per_step"]
        del y3, v3, y3_host, v3h, y3_page, v3p
        torch.cuda.empty_cache()

    # ================================================================== config 4 (extra key)
    # one superposition_vec over a 2^24-point grid x 20,000 Lorentzians, the grid sharded contiguously
    # over the ranks (strong scaling), parameters replicated, no exchange
    sup = None
    if not args.no_superposition:
        log("config 4")
        n_all, p4 = args.sup_points, args.sup_lorentzians
        lo4, hi4 = rank * n_all // world, (rank + 1) * n_all // world
        rng4 = np.random.Generator(np.random.PCG64(20260004))
        maxp4 = rng4.uniform(0.0, 10.0, p4)
        hw4 = np.exp(rng4.uniform(np.log(5e-4), np.log(3e-3), p4))
        sf4 = np.exp(rng4.uniform(0.0, np.log(1e4), p4))
        lor4_np = np.ascontiguousarray(np.stack([sf4 * hw4, hw4 * hw4, maxp4], axis=1))
        lor4 = torch.from_numpy(lor4_np).to(dev)
        x4_all = torch.linspace(-2.2, 11.8, n_all, dtype=torch.float64)
        x4 = x4_all[lo4:hi4].contiguous().to(dev)
        out4 = torch.empty_like(x4)
        mode_id = SUP_MODES[args.superposition]

        def sup_step(mode=mode_id):
            st = lib.mdb_superposition_vec_mode(x4.data_ptr(), hi4 - lo4, lor4.data_ptr(), p4, out4.data_ptr(), _lib.MDB_MEM_DEVICE, mode)
            assert st == 0, _lib.last_error()

        def time_fn(fn, reps):
            fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            barrier()
            return max_over_ranks([e0.elapsed_time(e1) / reps])[0]

        ms4 = time_fn(sup_step, 3)
        sup = {"workload": f"config4: superposition_vec, {n_all} grid points x {p4} Lorentzians, grid sharded over {world} GPU(s)",
               "mode": args.superposition, "evals_per_s": n_all * p4 / (ms4 / 1e3), "ms": ms4, "scaling": "strong",
               "checksum": float(out4[:: max(1, (hi4 - lo4) // 1024)].sum().item())}
        # the other arithmetic mode on the same grid: time, and the largest relative difference between the two
        main_out = out4.clone()
        ms4o = time_fn(lambda: sup_step(SUP_MODES[other_mode]), 1)
        rel4 = float(((main_out - out4).abs() / out4.abs().clamp_min(1e-300)).max().item())
        rel4 = max_over_ranks([rel4])[0]
        sup["other_mode"] = {"mode": other_mode, "evals_per_s": n_all * p4 / (ms4o / 1e3), "ms": ms4o,
                             "max_rel_difference_between_modes": rel4}
        # e2e: the same slice from pinned HOST memory through the C ABI (chunked H2D / kernel / D2H overlap)
        x4_host = x4_all[lo4:hi4].contiguous().pin_memory()
        out4_host = torch.empty_like(x4_host).pin_memory()

        def sup_host_step():
            st = lib.mdb_superposition_vec_mode(x4_host.data_ptr(), hi4 - lo4, lor4_np.ctypes.data, p4, out4_host.data_ptr(),
                                                _lib.MDB_MEM_HOST, mode_id)
            assert st == 0, _lib.last_error()

        ms4h = time_fn(sup_host_step, 3)
        sup_step()
        torch.cuda.synchronize()
        same = bool(torch.equal(out4_host.to(dev), out4))
        same = min_over_ranks([1.0 if same else 0.0])[0] == 1.0
        sup["e2e"] = {"evals_per_s": n_all * p4 / (ms4h / 1e3), "ms": ms4h, "fraction_of_device_resident": ms4 / ms4h,
                      "h2d_bytes": 8 * n_all + 24 * p4 * world, "d2h_bytes": 8 * n_all, "host_memory": "pinned",
                      "chunked_result_bit_identical_to_one_shot": same}
        # one call, all GPUs, from rank 0 (host wall clock; the other ranks wait on the host)
        if world > 1 and not args.no_one_call:
            host_barrier()
            if rank == 0:
                xa = x4_all.contiguous().pin_memory()
                oa = torch.empty_like(xa).pin_memory()
                assert lib.mdb_set_device_count(world) == 0
                try:
                    def one():
                        st = lib.mdb_superposition_vec_mode(xa.data_ptr(), n_all, lor4_np.ctypes.data, p4, oa.data_ptr(), _lib.MDB_MEM_HOST, mode_id)
                        assert st == 0, _lib.last_error()
                    one(); one()
                    ts = []
                    for _ in range(5):
                        t0 = time.perf_counter()
                        one()
                        ts.append(time.perf_counter() - t0)
                    assert lib.mdb_set_device_count(1) == 0
                    t0 = time.perf_counter()
                    one()
                    t_single = time.perf_counter() - t0
                    ref = oa.clone()
                    assert lib.mdb_set_device_count(world) == 0
                    one()
                    sup["one_call"] = {"devices": world, "ms": 1e3 * float(np.median(ts)), "evals_per_s": n_all * p4 / float(np.median(ts)),
                                       "one_device_same_call_ms": 1e3 * t_single, "speedup_over_one_device": t_single / float(np.median(ts)),
                                       "slices_bit_identical_to_one_device": bool(torch.equal(ref, oa)),
                                       "timing": "host wall clock around ONE C-ABI call from one process, pinned host memory"}
                finally:
                    assert lib.mdb_set_device_count(1) == 0
                del xa, oa
            host_barrier()
        del main_out, x4, out4, x4_host, out4_host, x4_all

    # ---- K1 alone at growing launch sizes: the exact-recurrence smoothing runs at chain latency, so
    # its GB/s is set by how many spectra share a launch; it turns bandwidth bound only with
    # thousands of them.  Reuses the batch as input, rank 0 at N = 1 only.
    smooth_sat = None
    if not args.no_smooth_saturation and world == 1:
        log("smoothing launch-size sweep")
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
        log("small spectra")
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

    # ---- CPU baseline on a bounded sample (rank 0 at N = 1) + parity of that sample, both value variants
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        log("cpu baseline + parity sample")
        import oracle as O
        cores = O.use_all_cores()
        n_sample = min(S, args.cpu_sample or 8 * cores)
        idx = sample_indices(S, n_sample)
        ys = y_dev[idx].cpu().numpy()
        O.par_deconvolute_spectra(O.Settings(), x_np, ys[:min(len(idx), cores)], SB)  # warm-up
        t0 = time.perf_counter()
        status, lors, mse, nsel = O.par_deconvolute_spectra(O.Settings(), x_np, ys, SB)
        dt = time.perf_counter() - t0
        cpu = {"value": len(idx) / dt, "unit": "spectra/s", "cores": cores, "kind": "port",
               "sample": f"{len(idx)} spectra drawn across the batch, oracle port (gcc -O3 -ffp-contract=off) with OpenMP over spectra, {dt:.1f} s"}
        fast = args.superposition == "fast"

        def compare(got, status, lors, mse, nsel):
            ok, rel = status == O.OK, 0.0
            for (g, npk, m), want, wn, wm in zip(got, lors, nsel, mse):
                ok = ok and g.shape == want.shape and np.array_equal(g.view(np.uint64), want.view(np.uint64)) and npk == wn
                rel = max(rel, abs(m - wm) / abs(wm))
            return bool(ok), rel

        _, _, got = run_batch(dec, dev_views, S, _lib.MDB_MEM_DEVICE, export=idx)
        ok_f, rel_f = compare(got, status, lors, mse, nsel)
        idx_i = sample_indices(n_int, n_sample)
        ys_i = y_int[idx_i].cpu().numpy()
        status_i, lors_i, mse_i, nsel_i = O.par_deconvolute_spectra(O.Settings(), x_np, ys_i, SB)
        _, _, got_i = run_batch(dec, int_views, n_int, _lib.MDB_MEM_DEVICE, export=idx_i)
        ok_i, rel_i = compare(got_i, status_i, lors_i, mse_i, nsel_i)
        # peak counts and Lorentzian parameters: identical bit patterns in both modes; the MSE is
        # bit-identical in exact mode and within 1e-9 relative (north_star) in fast mode
        tol = 1e-9 if fast else 0.0
        parity = {"float_valued": {"spectra": len(idx), "drawn": "across the batch", "peak_sets_and_lorentzians_bit_exact_vs_oracle": ok_f,
                                   "mse_max_rel_err_vs_oracle": rel_f},
                  "integer_valued": {"spectra": len(idx_i), "drawn": f"across the first {n_int} spectra, rounded to integers",
                                     "peak_sets_and_lorentzians_bit_exact_vs_oracle": ok_i, "mse_max_rel_err_vs_oracle": rel_i},
                  "mse_tolerance": tol, "superposition_mode": args.superposition,
                  "pass": bool(ok_f and ok_i and max(rel_f, rel_i) <= tol)}

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
    fp64_measured = {"dfma_per_s": dfma_rate.value, "dadd_per_s": dadd_rate.value,
                     "fraction_of_computed_peak": dfma_rate.value / (N_SM * FP64_LANES_PER_SM * sm_max * 1e6),
                     "how": "mdb_measure_fp64_rate: 16 independent DFMA (DADD) chains per thread on every SM, CUDA events, best of 3, run at the start of this bench"}

    traffic_db = {}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            traffic_db = json.load(fh).get("bytes_per_spectrum", {})
    spectra_per_launch = n_roof / ((n_roof + ROOFLINE_CHUNK - 1) // ROOFLINE_CHUNK)  # launches of the serial roofline pass

    def traffic(name):
        """DRAM bytes per launch: ncu-measured bytes per spectrum (profiles/traffic.json) x spectra per launch."""
        return traffic_db[name] * spectra_per_launch if name in traffic_db else None

    fast = args.superposition == "fast"
    instr_of = {"fit_iter": FP64_INSTR_PER_EVAL, "mse_superposition": FP64_INSTR_PER_EVAL_FAST if fast else FP64_INSTR_PER_EVAL}

    def fp64_roofline(p, name):
        if p["launches"] == 0 or p["ms"] <= 0:
            return None
        per_eval = instr_of[name]
        evals_per_s = p["work"] / (p["ms"] / 1e3)
        achieved = evals_per_s * FLOPS_PER_EVAL / 1e12
        return {"bound": "fp64", "kernel": name, "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak_tflops, "traffic": traffic(name),
                "peak_source": f"computed FMA peak: {N_SM} SMs x {FP64_LANES_PER_SM} FP64 lanes x 2 x {sm_max:.0f} MHz; the instruction rate "
                               f"behind it is MEASURED in this run (fp64_measured: {fp64_rate / 1e12:.2f} T DFMA/s)",
                "evals_per_s": evals_per_s, "evals_per_launch": p["work"] / p["launches"],
                "ms_per_launch": p["ms"] / p["launches"], "launches": p["launches"],
                "fp64_pipe_util": evals_per_s * per_eval / fp64_rate,
                "fp64_pipe_util_note": f"{per_eval} FP64-pipe instructions per evaluation (SASS) against the measured DFMA rate of this GPU"}

    def hbm_roofline(p, name):
        if p["launches"] == 0 or p["ms"] <= 0:
            return None
        achieved = p["work"] / (p["ms"] / 1e3) / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic(name), "peak_source": hbm_src,
                "bytes_per_launch": p["work"] / p["launches"], "ms_per_launch": p["ms"] / p["launches"],
                "launches": p["launches"]}

    def pipeline_view(p, ms_pipelined):
        evals = p["fit_iter"]["work"] + p["mse_superposition"]["work"]
        instr = sum(p[k]["work"] * instr_of[k] for k in instr_of)
        return {"evals_per_step": evals, "evals_per_s": evals / (ms_pipelined / 1e3), "fp64_instr_per_step": instr,
                "ms_pipelined": ms_pipelined, "fp64_pipe_util": instr / (ms_pipelined / 1e3) / fp64_rate,
                "note": "per GPU: evaluations of the serial roofline pass over the pipelined time of the same spectra, against the measured DFMA rate"}

    if sup is not None:
        sup["fp64_instr_per_eval"] = instr_of["mse_superposition"]
        sup["fp64_pipe_util"] = sup["evals_per_s"] * instr_of["mse_superposition"] / (world * fp64_rate)
        sup["frac_of_fp64_peak"] = sup["evals_per_s"] * FLOPS_PER_EVAL / 1e12 / (world * fp64_peak_tflops)
    overall = pipeline_view(prof, ms_roof_pipelined)
    dominant = max(("mse_superposition", "fit_iter"), key=lambda k: prof[k]["ms"])
    roofline = fp64_roofline(prof[dominant], dominant)
    other_name = "fit_iter" if dominant == "mse_superposition" else "mse_superposition"
    other = fp64_roofline(prof[other_name], other_name)
    if config3 is not None:
        p3 = config3.pop("_prof")
        config3["pipeline_fp64"] = pipeline_view(p3, config3.pop("_ms_full"))
        config3["roofline_fit_iter"] = fp64_roofline(p3["fit_iter"], "fit_iter")
        config3["roofline_mse_superposition"] = fp64_roofline(p3["mse_superposition"], "mse_superposition")
        config3["kernel_ms_serial_step"] = {k: v["ms"] for k, v in p3.items() if v["launches"]}
        for r in (config3["roofline_fit_iter"], config3["roofline_mse_superposition"]):
            if r:
                r["traffic"] = None

    line = {
        "metric": "spectra/s deconvolved (2^17 pts)", "value": value, "unit": "spectra/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "workload_stats": {"mean_selected_peaks": total_peaks / TOTAL, "mean_lorentzians": total_lor / TOTAL},
        "e2e": e2e, "e2e_pageable": e2e_pageable, "e2e_one_call": one_call,
        "gpu_launches": int(launches_all), "clocks": clocks, "fp64_measured": fp64_measured,
        "roofline": roofline, "roofline_other_fp64": other, "pipeline_fp64": overall,
        "other_superposition_mode": {"mode": other_mode, "value": TOTAL / (ms_other / 1e3), "unit": "spectra/s", "ms_per_step": ms_other},
        "weak_scaling": weak, "integer_valued_inputs": integer_variant,
        "roofline_hbm": {"detect": hbm_roofline(prof["detect"], "detect"), "smooth": hbm_roofline(prof["smooth"], "smooth")},
        "kernel_ms_serial_step": {k: v["ms"] for k, v in prof.items() if v["launches"]},
        "serial_step_ms": ms_serial,
        "roofline_pass": f"one extra step over {n_roof} spectra, chunks of {ROOFLINE_CHUNK}, one chunk at a time",
        "config3": config3, "superposition_vec": sup,
        "small_spectra": small,
        "smooth_launch_size_sweep": None if smooth_sat is None else
        [dict(e, frac_of_hbm_peak=e["gb_per_s"] / hbm_peak) for e in smooth_sat],
        "cpu_baseline": cpu, "parity_sample": parity, "parity_per_rank": parity_rank,
        "bench_wall_s": time.perf_counter() - t_start,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
