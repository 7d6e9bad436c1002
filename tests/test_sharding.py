"""N > 1 host logic on CPU: world_size-2 `gloo` process groups exercise the sharding layer that
`bench.py --gpus N` and multi-GPU callers use.  The per-rank compute is the CPU oracle here (test
infrastructure standing in for the GPU library, which needs a device); what is under test is the
partitioning, result ordering and first-error-wins semantics (deconvoluter.rs:655-658)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from metabodecon_rust_b200 import sharding  # noqa: E402


def test_shard_bounds_partition_exactly():
    for n in (0, 1, 2, 7, 8, 1000, 10000, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            bounds = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = sharding.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spectra(n_spec, n=4096):
    import synth
    x = synth.axis(n)
    return x, [synth.spectrum(100 + s, n=n, k=40, hw_range=(2e-3, 8e-3), x=x) for s in range(n_spec)]


class _Fail(Exception):
    pass


def _oracle_compute(x, bad=()):
    import oracle as O

    def compute(shard):
        out = []
        for tag, y in shard:
            if tag in bad:
                raise _Fail(f"spectrum {tag}")
            r = O.deconvolute_spectrum(O.Settings(), x, y, (11.8, -2.2))
            assert r.status == O.OK
            out.append((tag, np.ascontiguousarray(r.lorentzians), r.mse))
        return out
    return compute


def _worker(rank, world, port, n_spec, bad, queue):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, ys = _spectra(n_spec)
        items = list(enumerate(ys))
        try:
            res = sharding.deconvolute_spectra_sharded(items, compute=_oracle_compute(x, bad))
            payload = ("ok", [(t, l.tobytes(), m) for t, l, m in res])
        except _Fail as err:
            payload = ("fail", str(err))
        # grid sharding: every rank evaluates its slice, all ranks end with the full vector
        import oracle as O
        grid = np.linspace(-2.2, 11.8, 10001)
        lor = np.array([[0.03, 9e-4, 4.8], [0.02, 4e-4, 5.0], [0.03, 9e-4, 5.2]])
        full = sharding.superposition_vec_sharded(grid, lor, compute=lambda xs, p: O.superposition_vec(xs, p))
        queue.put((rank, payload, full.tobytes()))
    finally:
        dist.destroy_process_group()


def _run(world, n_spec, bad):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_spec, bad, queue)) for r in range(world)]
    for p in procs:
        p.start()
    out = [queue.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.timeout(600)
def test_two_rank_gloo_sharded_deconvolution_matches_unsharded():
    import oracle as O
    n_spec = 5
    x, ys = _spectra(n_spec)
    want = _oracle_compute(x)(list(enumerate(ys)))
    out = _run(2, n_spec, ())
    grid = np.linspace(-2.2, 11.8, 10001)
    lor = np.array([[0.03, 9e-4, 4.8], [0.02, 4e-4, 5.0], [0.03, 9e-4, 5.2]])
    full_want = O.superposition_vec(grid, lor).tobytes()
    for rank, (status, res), full in out:
        assert status == "ok"
        assert [t for t, _, _ in res] == list(range(n_spec))  # global order preserved
        for (t, lb, m), (wt, wl, wm) in zip(res, want):
            assert lb == wl.tobytes() and m == wm
        assert full == full_want


@pytest.mark.timeout(600)
def test_two_rank_gloo_first_failure_in_global_order_wins():
    # spectra 1 (rank 0's shard) and 3 (rank 1's shard) fail: every rank must report spectrum 1
    out = _run(2, 5, (3, 1))
    for rank, (status, msg), _ in out:
        assert status == "fail" and msg == "spectrum 1"
    # only rank 1's shard fails
    out = _run(2, 5, (4,))
    for rank, (status, msg), _ in out:
        assert status == "fail" and msg == "spectrum 4"
