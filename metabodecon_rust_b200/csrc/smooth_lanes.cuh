// smooth_lanes.cuh -- K1, the exact-recurrence moving average with one LANE per (spectrum, pass).
//
// smoothing/moving_average.rs:53-83 keeps one running sum per pass,
//     sum = fl(fl(sum + v[i+r]) - popped);  v[i] = fl(sum * div),
// and peak parity depends on that exact rounding sequence, so every pass is a
// strictly sequential chain per spectrum.  What CAN run in parallel:
//   * the I passes of one spectrum: pass p consumes the output stream of pass p-1, delayed by
//     (W + r) steps.  Lane p of a group of I lanes owns pass p; its input is what lane p-1 emitted
//     W steps earlier (one __shfl_up of a W-deep output history, so the shuffle never waits for
//     the step before it).  Every lane executes the same 3 FP64 instructions per step and the only
//     loop-carried dependence is each lane's own running sum (2 dependent adds): the kernel runs
//     at chain latency instead of issuing all I passes from one lane;
//   * spectra: floor(32 / I) spectra per warp, one warp per CTA, several CTAs per SM;
//   * memory: the [G spectra x T points] input tile of a warp is staged through shared memory with
//     one TMA bulk copy per row (cp.async.bulk + mbarrier, 4 stages in flight) and the smoothed
//     tile goes back with one bulk store per row, so HBM traffic is exactly 8N read + 8N written
//     per spectrum with full lines although each group walks its own row.
//
// State machine of one pass (circular_buffer.rs:34-59 semantics), input index s = 0, 1, ...:
//   s < r              preload: push, sum += v                                  (:58-61)
//   r <= s < n         push; if the FIFO was full pop the oldest and subtract it, else
//                      div = 1/len; emit out[s-r] = sum * div                   (:62-70)
//   n <= s < n + r     tail: pop, sum -= popped, div = 1/len, emit out[s-r]     (:71-79)
// The ring slot written at global step tau is (tau mod W) for every lane, so once a FIFO is full
// its oldest element sits exactly in the slot about to be overwritten and, with the step loop
// unrolled by W, ring slots are addressed statically (registers).
#pragma once
#include <cstdlib>
#include "kernels.cuh"

namespace mdb {

// ---- PTX wrappers for the store direction (loads: tma_bulk_g2s / mbarrier_* in kernels.cuh)
__device__ __forceinline__ bool mbarrier_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int SL_OSTAGES = 3;   // output tiles in the ring (two being written, one being stored)
constexpr int SL_THREADS = 32;  // one warp per CTA

// Tile length: a multiple of W (static ring slots) and even (16-byte bulk copies).  Three classes:
//   CLS 0  ~224 points, 4 input stages: amortises the per-tile costs (mbarrier wait, one bulk copy
//          per row each way, proxy fence) when a warp holds <= 10 spectra and the launch has at most
//          one warp per SM (its ~126 KB of shared memory allow no more);
//   CLS 1  ~112 points, 4 input stages: the 32-spectra-per-warp case (1 pass) and mid-size launches;
//   CLS 2  ~56 points, 3 input stages (~28 KB per warp): launches with thousands of spectra, where
//          eight warps per SM turn the kernel from latency bound to bandwidth bound.
template <int W, int CLS> struct SmoothTile {
    static constexpr int Q = (CLS == 0 ? 224 : CLS == 1 ? 112 : 56) / W;
    static constexpr int T = W * ((W % 2) ? (Q & ~1) : Q);
    static constexpr int IN_STAGES = CLS == 2 ? 3 : 4;  // input tiles in flight per warp
    // row stride in doubles: even (16-byte rows for the bulk copies) with an odd half, so that the
    // rows of the spectra sharing a warp start in different shared-memory banks
    static constexpr int STRIDE = ((T + 2) / 2) % 2 ? T + 2 : T + 4;
};

inline size_t smooth_lanes_smem_bytes(int stride, int in_stages, int groups)
{
    return (size_t)(in_stages + SL_OSTAGES) * groups * stride * 8 + in_stages * 8 + 64;
}

// State of one moving-average pass held by one lane: FIFO ring (slot written at global step tau is
// tau mod W), running sum, reciprocal length, and a W-deep history of its own outputs.
template <int W>
struct PassState {
    double f[W];   // FIFO contents
    double o[W];   // outputs of the last W steps: slot (tau mod W) is read by the next lane W steps later
    double sum, div, out;
    int len, head;

    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int k = 0; k < W; ++k) { f[k] = 0.0; o[k] = 0.0; }
        sum = 0.0; div = 1.0; out = 0.0; len = 0; head = 0;
    }
    __device__ __forceinline__ static double pick(const double (&a)[W], int slot)
    {
        double v = a[0];
#pragma unroll
        for (int k = 1; k < W; ++k) v = (slot == k) ? a[k] : v;
        return v;
    }
    __device__ __forceinline__ static void put(double (&a)[W], int slot, double v)
    {
#pragma unroll
        for (int k = 0; k < W; ++k) a[k] = (slot == k) ? v : a[k];
    }
    // One step of the state machine for any phase; sidx = input index of this pass at this step,
    // slot = tau mod W, n = spectrum length.  Returns true when an output (index sidx - W/2) was
    // produced into `out` (and recorded in the history).
    __device__ __forceinline__ bool step(bool valid, int sidx, int slot, double vin, int n)
    {
        constexpr int R = W / 2;
        if (!valid || sidx < 0 || sidx >= n + R) return false;
        if (sidx < n) {
            sum = __dadd_rn(sum, vin);
            if (len == W) {  // circular_buffer.rs:35-40: pop the oldest, push into its slot
                const double popped = pick(f, head);
                put(f, slot, vin);
                head = (head + 1 == W) ? 0 : head + 1;
                sum = __dsub_rn(sum, popped);
            } else {
                if (len == 0) head = slot;
                put(f, slot, vin);
                ++len;
                if (sidx >= R) div = __ddiv_rn(1.0, (double)len);  // main loop only (:66-68), not the preload
            }
            if (sidx < R) return false;
        } else {
            if (len <= 0) return false;  // tail (:71-79)
            const double popped = pick(f, head);
            head = (head + 1 == W) ? 0 : head + 1;
            --len;
            sum = __dsub_rn(sum, popped);
            div = __ddiv_rn(1.0, (double)len);
        }
        out = __dmul_rn(sum, div);
        put(o, slot, out);
        return true;
    }
};

template <int W, int CLS>
__global__ void __launch_bounds__(SL_THREADS)
smooth_lanes_kernel(const SpecDesc *__restrict__ sd, int n_spec, int iters)
{
    constexpr int T = SmoothTile<W, CLS>::T;
    constexpr int STRIDE = SmoothTile<W, CLS>::STRIDE;
    constexpr int SL_STAGES = SmoothTile<W, CLS>::IN_STAGES;
    constexpr int R = W / 2;
    static_assert(T % W == 0 && T % 2 == 0, "tile must hold whole ring rotations and 16-byte rows");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int G = 32 / iters;                 // spectra per warp
    const int grp = lane / iters, p = lane - grp * iters;
    const int s = blockIdx.x * G + grp;
    const bool valid = grp < G && s < n_spec;
    const bool first = valid && p == 0, last = valid && p == iters - 1;
    const int delay = p * (W + R);            // input index of this lane at global step tau: tau - delay
    const int L = R + (iters - 1) * (W + R);  // output index written at global step tau: tau - L

    double *in_buf = reinterpret_cast<double *>(smem_raw);                 // [SL_STAGES][G][STRIDE]
    double *out_buf = in_buf + SL_STAGES * G * STRIDE;                     // [SL_OSTAGES][G][STRIDE]
    uint64_t *full = reinterpret_cast<uint64_t *>(out_buf + SL_OSTAGES * G * STRIDE);

    const double *__restrict__ y = valid ? sd[s].y : nullptr;
    double *__restrict__ ys = valid ? sd[s].ys : nullptr;
    const int n = valid ? sd[s].n : 0;
    const int n_max = __reduce_max_sync(0xffffffffu, n);
    const int n_min = __reduce_min_sync(0xffffffffu, valid ? n : 0x7fffffff);
    const int steps = n_max + L + 1;          // every lane has drained by then
    const int tiles = (steps + T - 1) / T;
    const int in_tiles = (n_max + T - 1) / T;

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < SL_STAGES; ++q) mbarrier_init(&full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();

    auto issue_load = [&](int k) {  // rows of input tile k, one bulk copy per spectrum of the warp
        const int stage = k % SL_STAGES;
        int cnt = first ? n - k * T : 0;
        cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
        const uint32_t bytes = (uint32_t)(cnt & ~1) * 8u;  // bulk copies move multiples of 16 bytes
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbarrier_expect_tx(&full[stage], total);
        __syncwarp();
        if (bytes) tma_bulk_g2s(in_buf + (stage * G + grp) * STRIDE, y + (size_t)k * T, bytes, &full[stage]);
    };
    auto orow = [&](int tile) -> double * { return out_buf + ((tile % SL_OSTAGES) * G + grp) * STRIDE; };
    auto flush_tile = [&](int j) {  // output tile j of every spectrum: generic-proxy writes -> bulk store
        fence_proxy_async_smem();
        __syncwarp();
        int cnt = last ? n - j * T : 0;
        cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
        const uint32_t bytes = (uint32_t)(cnt & ~1) * 8u;
        if (bytes) tma_bulk_s2g(ys + (size_t)j * T, orow(j), bytes);
        if (cnt & 1) ys[(size_t)j * T + cnt - 1] = orow(j)[cnt - 1];
        tma_bulk_commit();
    };

    for (int k = 0; k < SL_STAGES && k < in_tiles; ++k) issue_load(k);

    // ---- per-lane pass state
    PassState<W> ps;
    ps.reset();

    for (int k = 0; k < tiles; ++k) {
        const int stage = k % SL_STAGES;
        const bool has_input = k < in_tiles;
        if (has_input) {
            while (!mbarrier_try_wait(&full[stage], (uint32_t)((k / SL_STAGES) & 1))) {}
        }
        double *irow = in_buf + (stage * G + grp) * STRIDE;
        if (has_input && first) {  // odd trailing element of a row: not part of the 16-byte-granular bulk copy
            int cnt = n - k * T;
            cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
            if (cnt & 1) irow[cnt - 1] = y[(size_t)k * T + cnt - 1];
        }
        // the ring slot that output tile k will use was flushed SL_OSTAGES tiles ago: make sure it was read
        tma_bulk_wait_read<1>();
        __syncwarp();

        const int tau0 = k * T;
        // outputs of this tile's steps have indices tau - L: the first L of them land in output tile k-1
        double *__restrict__ pa = orow(k >= 1 ? k - 1 : 0) + (T - L);  // only dereferenced when k >= 1 (fast tiles)
        double *__restrict__ pb = orow(k) - L;
        const bool fast = (k >= 1) && (tau0 >= W + (iters - 1) * (W + R)) && (tau0 + T <= n_min);
        if (fast) {
            // steady state for every valid lane: FIFO full, static ring slots, no phase logic.
            // The W inputs of the next rotation are fetched from shared memory one rotation ahead,
            // so the only latency left on the critical path is the running sum's two adds.
            double cur[W], nxt[W];
#pragma unroll
            for (int u = 0; u < W; ++u) { cur[u] = irow[u]; nxt[u] = 0.0; }
#pragma unroll 4
            for (int g = 0; g < T / W; ++g) {
                if (g + 1 < T / W) {
#pragma unroll
                    for (int u = 0; u < W; ++u) nxt[u] = irow[(g + 1) * W + u];
                }
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int j = g * W + u;
                    const double up = __shfl_up_sync(0xffffffffu, ps.o[u], 1);  // what lane p-1 emitted W steps ago
                    const double vin = first ? cur[u] : up;
                    ps.sum = __dsub_rn(__dadd_rn(ps.sum, vin), ps.f[u]);
                    ps.f[u] = vin;
                    ps.o[u] = __dmul_rn(ps.sum, ps.div);
                    if (last) { double *q = (j < L) ? pa : pb; q[j] = ps.o[u]; }
                }
#pragma unroll
                for (int u = 0; u < W; ++u) cur[u] = nxt[u];
            }
        } else {
            for (int j = 0; j < T; ++j) {
                const int tau = tau0 + j;
                const int slot = tau % W;
                double vin = __shfl_up_sync(0xffffffffu, PassState<W>::pick(ps.o, slot), 1);
                if (first && tau < n) vin = irow[j];
                const int sidx = tau - delay;
                const bool emitted = ps.step(valid, sidx, slot, vin, n);
                if (last && emitted) {
                    const int oi = sidx - R;  // == tau - L
                    orow(oi / T)[oi % T] = ps.out;
                }
            }
            // a fast tile may follow: its static addressing needs head == tau mod W, which holds for
            // every full FIFO (see the header); nothing to normalise
        }
        __syncwarp();
        if (k + SL_STAGES < in_tiles) issue_load(k + SL_STAGES);
        if (k >= 1) flush_tile(k - 1);
    }
    if (tiles >= 1) flush_tile(tiles - 1);
    tma_bulk_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------
// All passes of ONE spectrum at once (Deconvoluter::optimize_settings, deconvoluter.rs:761-825):
// the output of pass k of an I-pass smoothing IS the result of smoothing with k iterations, so one
// run with I = `iters` lanes yields every iteration count 1..iters for this window.  Lane p writes
// pass p's output stream to out + p * stride.  One warp; blockIdx.x selects the job.
// ---------------------------------------------------------------------------------------------
struct SmoothAllJob {
    const double *y;   // raw intensities
    double *out;       // iters rows of `stride` doubles
    long long stride;
    int n, iters, window;
};

template <int W>
__device__ void smooth_all_passes_body(const SmoothAllJob &job)
{
    constexpr int R = W / 2;
    const int lane = threadIdx.x;
    const int n = job.n, iters = job.iters;
    const bool valid = lane < iters;
    const int delay = lane * (W + R);
    double *__restrict__ row = job.out + (long long)lane * job.stride;
    PassState<W> ps;
    ps.reset();
    const int steps = n + R + (iters - 1) * (W + R) + 1;
    double chunk = (lane < n) ? job.y[lane] : 0.0;  // inputs tau0 .. tau0+31, one per lane (coalesced)
    for (int tau0 = 0; tau0 < steps; tau0 += 32) {
        const int nxt_i = tau0 + 32 + lane;
        const double nxt = (nxt_i < n) ? job.y[nxt_i] : 0.0;  // prefetch the next 32 inputs
#pragma unroll 1
        for (int u = 0; u < 32; ++u) {
            const int tau = tau0 + u;
            const int slot = tau % W;
            const double raw = __shfl_sync(0xffffffffu, chunk, u);
            const double up = __shfl_up_sync(0xffffffffu, PassState<W>::pick(ps.o, slot), 1);
            const double vin = (lane == 0) ? raw : up;
            const int sidx = tau - delay;
            if (ps.step(valid, sidx, slot, vin, n)) row[sidx - R] = ps.out;
        }
        chunk = nxt;
    }
}

__global__ void __launch_bounds__(32)
smooth_all_passes_kernel(const SmoothAllJob *__restrict__ jobs)
{
    const SmoothAllJob job = jobs[blockIdx.x];
    switch (job.window) {
    case 2: smooth_all_passes_body<2>(job); break;
    case 3: smooth_all_passes_body<3>(job); break;
    case 4: smooth_all_passes_body<4>(job); break;
    case 5: smooth_all_passes_body<5>(job); break;
    case 6: smooth_all_passes_body<6>(job); break;
    case 7: smooth_all_passes_body<7>(job); break;
    case 8: smooth_all_passes_body<8>(job); break;
    case 9: smooth_all_passes_body<9>(job); break;
    default: break;  // the host only submits windows 2..9
    }
}

using SmoothLanesFn = void (*)(const SpecDesc *, int, int);

// Returns the kernel for `window` (2..9) and the tile length, or nullptr when the settings need
// the generic path (window > 9, more than 32 iterations, or a pipeline lag longer than a tile).
inline SmoothLanesFn smooth_lanes_lookup(int window, int iterations, size_t n_spectra, int sm_count, int *stride,
                                         int *in_stages)
{
    if (iterations < 1 || iterations > 32) return nullptr;
    const size_t warps = (n_spectra + (size_t)(32 / iterations) - 1) / (size_t)(32 / iterations);
    const int lag = (window / 2) + (iterations - 1) * (window + window / 2);  // L = r + (I-1)(W+r) must fit a tile
    int cls = (iterations >= 3 && warps <= (size_t)sm_count) ? 0 : (iterations >= 3 && warps > (size_t)3 * sm_count) ? 2 : 1;
    if (const char *env = std::getenv("MDB_SMOOTH_CLASS")) {  // sweeps: force a tile class (0: 224 points / ~126 KB per warp, 1: 112, 2: 56 / ~28 KB)
        if (env[0] >= '0' && env[0] <= '2' && iterations >= 3) cls = env[0] - '0';
    }
#define MDB_SL_CASE(Wv) \
    if (window == Wv) { \
        if (cls == 2 && lag > SmoothTile<Wv, 2>::T) cls = 1; \
        if (cls == 1 && lag > SmoothTile<Wv, 1>::T) cls = iterations >= 3 ? 0 : -1; \
        if (cls == 0 && lag > SmoothTile<Wv, 0>::T) cls = -1; \
        if (cls < 0) return nullptr; \
        *stride = cls == 0 ? SmoothTile<Wv, 0>::STRIDE : cls == 1 ? SmoothTile<Wv, 1>::STRIDE : SmoothTile<Wv, 2>::STRIDE; \
        *in_stages = cls == 2 ? SmoothTile<Wv, 2>::IN_STAGES : SmoothTile<Wv, 0>::IN_STAGES; \
        return cls == 0 ? smooth_lanes_kernel<Wv, 0> : cls == 1 ? smooth_lanes_kernel<Wv, 1> : smooth_lanes_kernel<Wv, 2>; }
    MDB_SL_CASE(2) MDB_SL_CASE(3) MDB_SL_CASE(4) MDB_SL_CASE(5) MDB_SL_CASE(6) MDB_SL_CASE(7) MDB_SL_CASE(8) MDB_SL_CASE(9)
#undef MDB_SL_CASE
    return nullptr;
}

}  // namespace mdb
