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
constexpr int STREAM_G = 64;          // rounds between flow-control checks
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

__device__ __forceinline__ void lds_f64x2(uint32_t a, double &v0, double &v1)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double v0, double v1)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}

// WS: the window size when it is one the interior loop is specialised for (3, 5, 7: the default and the
// values optimize_settings tries), 0 = any window (runtime `w`, sixteen 8-byte loads per round).
template <int WS>
__global__ void __launch_bounds__(STREAM_THREADS)
smooth_stream_kernel(const SpecDesc *__restrict__ sd, int iters, int w_arg)
{
    extern __shared__ __align__(16) unsigned char stream_smem[];
    __shared__ StreamFlow flow;
    __shared__ __align__(16) double sink[STREAM_U];
    constexpr int U = STREAM_U, R = STREAM_R, G = STREAM_G;
    const SpecDesc d = sd[blockIdx.x];
    const int n = d.n;
    const int t = threadIdx.x, lane = t & 31;
    const int w = WS > 0 ? WS : w_arg;
    const int r = w / 2;
    // blocks between consecutive passes.  Any-window loop: the consumer's reach into the producer's stream
    // (ceil(max(r, w - r) / U) blocks), one round until a store is visible (__syncwarp), one of slack.
    // Specialised loop: it fetches ALIGNED block m+2 in round m (no reach beyond the block), the producer stores
    // a block one round after computing it and the store is visible one round later: m + 2 computed in round
    // m + 2 - db of the consumer's clock, stored in m + 3 - db, readable from m + 4 - db <= m.
    const int db = WS > 0 ? 4 : 2 + (max(r, w - r) + U - 1) / U;
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
                // all eight loads of a thread in flight at once: a mover pass costs one global-memory round trip,
                // not eight (the movers, not the chain, bounded the kernel once the chain loop had been tightened)
                const int hi = min(feed_limit, fed + 8 * STREAM_MOVERS);
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = fed + m + k * STREAM_MOVERS;
                    v[k] = (i < hi) ? __ldg(y + i) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = fed + m + k * STREAM_MOVERS;
                    if (i < hi) {
                        const int s = i & (R - 1);
                        ring_in[s] = v[k];
                        if (s < 8) ring_in[R + s] = v[k];
                    }
                }
                fed = hi;
                worked = true;
            }
            // drain: the last pass has finished blocks below taken - (iters-1)*db
            const int done_pts = (taken >= rounds) ? n : min(n, max(0, taken - (iters - 1) * db) * U);
            if (drained < done_pts) {
                const int hi = min(done_pts, drained + 8 * STREAM_MOVERS);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = drained + m + k * STREAM_MOVERS;
                    if (i < hi) ys[i] = ring_out[i & (R - 1)];
                }
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
    int ghost_i = lane >= iters ? 1 : 0;
    asm volatile("mov.b32 %0, %0;" : "+r"(ghost_i));  // opaque, see below
    const bool ghost = ghost_i != 0;
    const int pl = ghost ? 0 : lane;
    uint32_t in_s = smem_addr(rings + (size_t)pl * RS), out_s = ghost ? smem_addr(sink) : smem_addr(rings + (size_t)(pl + 1) * RS);
    uint32_t out_mask = ghost ? 0u : (uint32_t)(R - 1);   // ghosts: every store lands in the sink
    // opaque to ptxas from here on: it otherwise re-derives these inside the chain loop (S2R SR_TID / SR_CgaCtaId
    // and the address arithmetic behind them, with the in-order warp waiting for the special-register reads)
    asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;\n\tmov.b32 %2, %2;" : "+r"(in_s), "+r"(out_s), "+r"(out_mask));
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
    if constexpr (WS > 0) {
    if (b_hi - b_lo >= 8 && b == b_lo) {
        // Window size known at compile time (WS = w): the chain warp keeps the ALIGNED 8-point blocks
        // m-1, m, m+1 of its input stream in registers, so the incoming value x[i + r] and the outgoing
        // x[i + r - w] of every step are register operands picked at compile time, and a round costs
        // four 16-byte shared loads (block m+2) instead of sixteen 8-byte ones.  Software pipeline over
        // rounds: in round k the warp (1) fetches block k+2, (2) stores the outputs of block k-1, which
        // have been sitting in registers since the last round, (3) runs the chain over block k.  Nothing
        // issued in a round depends on that round's chain, so the in-order warp never waits for a
        // product or a load before it may start the next chain.  The store delay and the deeper
        // prefetch are what the larger pass distance `db` pays for.
        constexpr int RR = WS / 2, QB = WS - RR;          // x[i + RR] comes in, x[i - QB] goes out
        static_assert(RR <= U && QB <= U, "window wider than two blocks");
        double B0[U], B1[U], B2[U], B3[U];                // aligned blocks m-1, m, m+1, m+2 (rotating)
        // outputs of block m-1, not yet stored: always in prB between rounds.  Two sets, so that the products of
        // a round never overwrite registers a store issued in the same round has yet to read
        double prA[U], prB[U];
        int m = b + blk0;                                 // this lane's block in the coming round
        uint32_t so = ((uint32_t)(m * U)) & out_mask;     // output slot of block m
        uint32_t so_prev = 0;
        auto load_aligned = [&](double (&x)[U], int blk) {
            const uint32_t pa = in_s + 8u * ((uint32_t)(blk * U) & (uint32_t)(R - 1));
#pragma unroll
            for (int u = 0; u < U; u += 2) lds_f64x2(pa + 8u * u, x[u], x[u + 1]);
        };
        auto store_pending = [&](const double (&ps)[U]) {
            const uint32_t po = out_s + 8u * so_prev;
#pragma unroll
            for (int u = 0; u < U; u += 2) sts_f64x2(po + 8u * u, ps[u], ps[u + 1]);
            if (so_prev == 0 && !ghost) {                 // the mirror of slots [0, 8) (read by the edge rounds)
#pragma unroll
                for (int u = 0; u < U; u += 2) sts_f64x2(out_s + 8u * (R + u), ps[u], ps[u + 1]);
            }
        };
        // one round: chain over block m (inputs from pm / p0 / p1 = blocks m-1 / m / m+1), products into pw;
        // block m+2 fetched into ld; the products of the round before (ps) stored
        auto lean_round = [&](const double (&pm)[U], const double (&p0)[U], const double (&p1)[U], double (&ld)[U],
                              double (&pw)[U], const double (&ps)[U], const bool pend) {
            load_aligned(ld, m + 2);
            if (pend) store_pending(ps);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double a = (u + RR < U) ? p0[(u + RR) % U] : p1[(u + RR) % U];
                const double q = (u >= QB) ? p0[(u + U - QB) % U] : pm[(u + U - QB) % U];
                sum = __dadd_rn(sum, a);
                sum = __dsub_rn(sum, q);
                pw[u] = __dmul_rn(sum, div);
            }
            so_prev = so;
            ++m;
            so = (so + U) & out_mask;
            __syncwarp();
        };
        flow_wait(min(b_hi + 1, b + G));
        load_aligned(B0, m - 1);
        load_aligned(B1, m);
        load_aligned(B2, m + 1);
        bool pend = false;
        while (b <= b_hi) {
            const int e = min(b_hi + 1, b + G);      // this group: rounds [b, e)
            // the prefetch of the group's last round reaches two blocks into the next group
            flow_wait(min(b_hi + 1, e + 2));
            int left = e - b;
            if (!pend && left >= 1) {                // first lean round of all: nothing to store yet
                lean_round(B0, B1, B2, B3, prA, prB, false);
#pragma unroll
                for (int u = 0; u < U; ++u) { B0[u] = B1[u]; B1[u] = B2[u]; B2[u] = B3[u]; prB[u] = prA[u]; }
                pend = true;
                --left;
            }
            for (; left >= 4; left -= 4) {
                lean_round(B0, B1, B2, B3, prA, prB, true);
                lean_round(B1, B2, B3, B0, prB, prA, true);
                lean_round(B2, B3, B0, B1, prA, prB, true);
                lean_round(B3, B0, B1, B2, prB, prA, true);
            }
            for (; left >= 1; --left) {
                lean_round(B0, B1, B2, B3, prA, prB, true);
#pragma unroll
                for (int u = 0; u < U; ++u) { B0[u] = B1[u]; B1[u] = B2[u]; B2[u] = B3[u]; prB[u] = prA[u]; }
            }
            b = e;
            // block b-1 of every pass is still in registers: the movers are told about one round less
            if (b <= b_hi) publish(b - 1);
        }
        if (pend) { store_pending(prB); }
        publish(b);
    }
    } else {
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
