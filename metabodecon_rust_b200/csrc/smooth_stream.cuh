// smooth_stream.cuh -- K1 for a FEW long spectra: the latency form of the moving average.
//
// smooth_lanes_kernel packs floor(32 / passes) spectra into a warp and is the right shape when
// thousands of spectra are in flight (67 % of HBM at 11 840 per launch).  For one spectrum, or a
// few dozen, what matters is the per-point cost of the sequential running-sum chain, and there the
// shuffle hand-off of that kernel costs 43 cycles per point against 16.4 for the two dependent
// additions.  This kernel is the interior loop of small_fused.cuh (29 cycles per point) stretched
// to any length: one CTA per spectrum; warp 0 runs the chain, lane p = pass p, reading ring p and
// writing ring p+1 in shared memory; the other warps feed ring 0 from global memory and drain the
// last ring to global memory.  The chain warp meets the movers only every STREAM_G rounds
// (256 points) through three counters in shared memory, so flow control costs nothing per point.
//
//   ring p (p = 0..passes): STREAM_R points as a circular buffer, element i at slot i & (R-1),
//   plus a mirror of slots [0, 8) at [R, R+8): every block of 8 consecutive elements can then be
//   read from its first slot without wrapping (blocks are written slot-aligned, block 0 twice).
//
// Arithmetic and operation order are those of smooth_pass_generic_kernel / moving_average.rs:53-83
// (see the notes on small_smooth_warp in small_fused.cuh for the +0.0 / -0.0 edge identities and
// the hazard analysis of the pass delay `db`); tests/test_gpu_parity.py::test_smoothing_* compare
// it bit for bit with the oracle and with the other two smoothing kernels.
#pragma once
#include "kernels.cuh"
#include "small_fused.cuh"

namespace mdb {

constexpr int STREAM_THREADS = 128;   // warp 0: chain; warps 1-3: movers
constexpr int STREAM_MOVERS = STREAM_THREADS - 32;
constexpr int STREAM_R = 2048;        // ring length in points (power of two)
constexpr int STREAM_U = 8;           // points per round
constexpr int STREAM_G = 32;          // rounds between flow-control checks
constexpr int STREAM_MAX_ITERS = 12;  // (passes + 1) rings of 16 KB must fit into shared memory

__host__ __device__ inline size_t smooth_stream_smem_bytes(int iters) { return (size_t)(iters + 1) * (STREAM_R + 8) * 8; }

struct StreamFlow {
    int fed;       // input points in ring 0 (movers -> chain)
    int taken;     // rounds completed by the chain warp (chain -> movers)
    int drained;   // output points copied out of the last ring (movers -> chain)
    int pad_;      // the movers' shared sample of `taken`
};

__device__ __forceinline__ int ld_volatile_s32(const int *p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s32(int *p, int v)
{
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}

__global__ void __launch_bounds__(STREAM_THREADS)
smooth_stream_kernel(const SpecDesc *__restrict__ sd, int iters, int w)
{
    extern __shared__ __align__(16) unsigned char stream_smem[];
    __shared__ StreamFlow flow;
    __shared__ __align__(16) double sink[STREAM_U];
    constexpr int U = STREAM_U, R = STREAM_R, G = STREAM_G;
    const SpecDesc d = sd[blockIdx.x];
    const int n = d.n;
    const int t = threadIdx.x, lane = t & 31;
    const int r = w / 2;
    const int db = 2 + (max(r, w - r) + U - 1) / U;   // blocks between consecutive passes
    const int n_blocks = (n + U - 1) / U;
    const int rounds = n_blocks + (iters - 1) * db;
    double *rings = reinterpret_cast<double *>(stream_smem);
    constexpr int RS = R + 8;                          // ring stride in doubles
    if (t == 0) { flow.fed = 0; flow.taken = 0; flow.drained = 0; }
    __syncthreads();

    if (t >= 32) {
        // ---------------------------------------------------------------- movers
        const int m = t - 32;
        const double *__restrict__ y = d.y;
        double *__restrict__ ys = d.ys;
        double *ring_in = rings, *ring_out = rings + (size_t)iters * RS;
        int fed = 0, drained = 0;
        for (;;) {
            // one mover samples the chain's progress for all (every mover must take the same branches:
            // they meet at named barrier 1)
            if (m == 0) st_volatile_s32(&flow.pad_, ld_volatile_s32(&flow.taken));
            asm volatile("bar.sync 1, %0;" ::"n"(STREAM_MOVERS) : "memory");
            const int taken = ld_volatile_s32(&flow.pad_);
            __threadfence_block();  // the chain's ring stores behind `taken` are visible from here
            bool worked = false;
            // feed: lane 0 of the chain is at block `taken`, its lowest live index is taken*U + r - w
            const int feed_limit = min(n, taken * U + r - w + R - U);
            if (fed < feed_limit) {
                const int hi = min(feed_limit, fed + 8 * STREAM_MOVERS);
                for (int i = fed + m; i < hi; i += STREAM_MOVERS) {
                    const double v = y[i];
                    const int s = i & (R - 1);
                    ring_in[s] = v;
                    if (s < 8) ring_in[R + s] = v;
                }
                fed = hi;
                worked = true;
            }
            // drain: the last pass has finished blocks below taken - (iters-1)*db
            const int done_pts = (taken >= rounds) ? n : min(n, max(0, taken - (iters - 1) * db) * U);
            if (drained < done_pts) {
                const int hi = min(done_pts, drained + 8 * STREAM_MOVERS);
                for (int i = drained + m; i < hi; i += STREAM_MOVERS) ys[i] = ring_out[i & (R - 1)];
                drained = hi;
                worked = true;
            }
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"n"(STREAM_MOVERS) : "memory");
            if (m == 0 && worked) {
                st_volatile_s32(&flow.fed, fed);
                st_volatile_s32(&flow.drained, drained);
            }
            if (drained >= n) break;
            if (!worked) __nanosleep(200);
        }
        return;
    }

    // -------------------------------------------------------------------- chain warp
    // Lanes without a pass shadow lane 0 (same reads, stores to a sink): the warp stays converged.
    const bool ghost = lane >= iters;
    const int pl = ghost ? 0 : lane;
    const uint32_t in_s = smem_addr(rings + (size_t)pl * RS), out_s = ghost ? smem_addr(sink) : smem_addr(rings + (size_t)(pl + 1) * RS);
    const uint32_t out_mask = ghost ? 0u : (uint32_t)(R - 1);   // ghosts: every store lands in the sink
    int blk0 = ghost ? 0 : -lane * db;                          // this lane's block in round b is b + blk0
    double sum = 0.0, div = 1.0;

    // wait until the movers are far enough for rounds [b, b_end): lane 0's inputs, the last lane's room
    auto flow_wait = [&](int b_end) {
        const int need_fed = min(n, (b_end + 1) * U + r);
        const int need_drained = (b_end - (iters - 1) * db) * U - R;
        while (ld_volatile_s32(&flow.fed) < need_fed || ld_volatile_s32(&flow.drained) < need_drained) __nanosleep(100);
        __threadfence_block();  // the movers' ring stores behind `fed` are visible from here
    };
    auto publish = [&](int b_done) {
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_volatile_s32(&flow.taken, b_done);
    };

    // any block, edges included (callers guarantee n >= w); element addresses are masked one by one
    auto edge_round = [&](int b) {
        const int mb = b + blk0;
        if (mb >= 0 && mb < n_blocks) {
            const int i0 = mb * U;
            if (mb == 0)
                for (int k = 0; k < r; ++k) sum = __dadd_rn(sum, lds_f64(in_s + 8u * (uint32_t)(k & (R - 1))));
            double a[U], q[U], dv[U];
            bool redo[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ai = i0 + u + r, qi = ai - w;
                a[u] = (ai < n) ? lds_f64(in_s + 8u * (uint32_t)(ai & (R - 1))) : -0.0;
                q[u] = (qi >= 0) ? lds_f64(in_s + 8u * (uint32_t)(qi & (R - 1))) : 0.0;
                redo[u] = qi < 0 || ai >= n;  // the window grows or shrinks in this step
                dv[u] = redo[u] ? __ddiv_rn(1.0, (double)(min(ai, n - 1) - max(qi, -1))) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                sum = __dadd_rn(sum, a[u]);
                sum = __dsub_rn(sum, q[u]);
                if (redo[u]) div = dv[u];
                a[u] = __dmul_rn(sum, div);
            }
            const uint32_t s0 = (uint32_t)i0 & out_mask;  // block-aligned: no wrap inside the block
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u < n) {
                    sts_f64(out_s + 8u * (s0 + u), a[u]);
                    if (s0 == 0 && !ghost) sts_f64(out_s + 8u * (R + u), a[u]);  // the mirror of slots [0, 8)
                }
        }
        __syncwarp();
    };

    // interior blocks: window full (i0 >= w - r) and no tail (i0 + U <= n - r)
    const int f0 = (w - r + U - 1) / U, f1 = (n - r) / U - 1;
    const int b_lo = f0 + (iters - 1) * db, b_hi = f1;  // rounds in which EVERY pass is on an interior block
    int b = 0;
    {
        const int first = min(rounds, b_lo);
        for (; b < first; b += G) {
            const int e = min(first, b + G);
            flow_wait(e);
            for (int bb = b; bb < e; ++bb) edge_round(bb);
            publish(e);
        }
        b = first;
    }
    if (b_hi - b_lo >= 3 && b == b_lo) {
        double a0[U], q0[U], a1[U], q1[U];
        const uint32_t w8 = 8u * (uint32_t)w;
        int i_in = (b + blk0) * U + r;        // logical index of the first input of the block in registers
        uint32_t so = ((uint32_t)((b + blk0) * U)) & out_mask;  // output slot of that block
        // one interior round: chain over the block held in (ca, cq), next block fetched into (na, nq);
        // a block of 8 inputs is read from the slot of its first element (mirror: no wrap inside)
        auto lean_round = [&](double (&ca)[U], double (&cq)[U], double (&na)[U], double (&nq)[U], bool prefetch) {
            if (prefetch) {
                const uint32_t pa = in_s + 8u * ((uint32_t)(i_in + U) & (uint32_t)(R - 1));
                const uint32_t pq = in_s + 8u * ((uint32_t)(i_in + U - w) & (uint32_t)(R - 1));
#pragma unroll
                for (int u = 0; u < U; ++u) { na[u] = lds_f64(pa + 8u * u); nq[u] = lds_f64(pq + 8u * u); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                sum = __dadd_rn(sum, ca[u]);
                sum = __dsub_rn(sum, cq[u]);
                ca[u] = __dmul_rn(sum, div);
            }
            const uint32_t po = out_s + 8u * so;
#pragma unroll
            for (int u = 0; u < U; ++u) sts_f64(po + 8u * u, ca[u]);
            if (so == 0 && !ghost) {
#pragma unroll
                for (int u = 0; u < U; ++u) sts_f64(out_s + 8u * (R + u), ca[u]);
            }
            i_in += U;
            so = (so + U) & out_mask;
            __syncwarp();
        };
        (void)w8;
        flow_wait(min(b_hi + 1, b + G));
        {
            const uint32_t pa = in_s + 8u * ((uint32_t)i_in & (uint32_t)(R - 1));
            const uint32_t pq = in_s + 8u * ((uint32_t)(i_in - w) & (uint32_t)(R - 1));
#pragma unroll
            for (int u = 0; u < U; ++u) { a0[u] = lds_f64(pa + 8u * u); q0[u] = lds_f64(pq + 8u * u); }
        }
        while (b <= b_hi) {
            const int e = min(b_hi + 1, b + G);      // this group: rounds [b, e)
            // the prefetch of the group's last round reaches one block into the next group
            flow_wait(min(b_hi + 1, e + 1));
            int left = e - b;
            for (; left >= 2; left -= 2) {
                lean_round(a0, q0, a1, q1, true);
                lean_round(a1, q1, a0, q0, true);
            }
            if (left == 1) {
                lean_round(a0, q0, a1, q1, true);
#pragma unroll
                for (int u = 0; u < U; ++u) { a0[u] = a1[u]; q0[u] = q1[u]; }
            }
            b = e;
            publish(b);
        }
    }
    for (; b < rounds; b += G) {
        const int e = min(rounds, b + G);
        flow_wait(e);
        for (int bb = b; bb < e; ++bb) edge_round(bb);
        publish(e);
    }
    publish(rounds);
}

}  // namespace mdb
