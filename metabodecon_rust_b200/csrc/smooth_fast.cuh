// smooth_fast.cuh -- K1, the exact-recurrence moving average as ONE streaming kernel.
//
// smoothing/moving_average.rs:53-83 keeps one running sum per pass,
//     sum = fl(fl(sum + v[i+r]) - popped);  v[i] = fl(sum * div),
// and peak parity depends on that exact rounding sequence, so every pass is a
// strictly sequential chain per spectrum.  What CAN be parallel:
//   * spectra: one lane per spectrum, 32 spectra per warp;
//   * the I passes: pass p+1 consumes pass p's output stream as it is produced, so all passes run
//     software-pipelined inside the same thread (I independent dependency chains = ILP), with the
//     FIFOs held in registers (ring slots addressed statically in the steady state);
//   * memory: the [32 spectra x T points] input tile of a warp is staged through shared memory
//     with one TMA bulk copy per row (cp.async.bulk + mbarrier, 4 stages in flight), and the
//     smoothed tile goes back with one bulk store per row, so HBM traffic is exactly 8N read +
//     8N written per spectrum with full 128-byte lines although each lane walks its own row.
// Rows are padded to an odd multiple of 16 bytes in shared memory (bulk-copy alignment + at most
// 2-way bank conflicts for the per-lane 64-bit accesses).
#pragma once
#include "kernels.cuh"

namespace mdb {

// ---- PTX wrappers (TMA bulk copies, mbarrier, proxy fence)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- one moving-average pass as a push/pop state machine with a register-resident ring
template <int W>
struct MAPass {
    double f[W];   // ring buffer (circular_buffer.rs): oldest at slot `head`
    double sum, div;
    int len, head, cnt;  // items held, slot of the oldest, inputs consumed so far

    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = 0.0;
        sum = 0.0; div = 1.0; len = 0; head = 0; cnt = 0;
    }
    __device__ __forceinline__ double get(int slot) const
    {
        double r = f[0];
#pragma unroll
        for (int k = 1; k < W; ++k) r = (slot == k) ? f[k] : r;
        return r;
    }
    __device__ __forceinline__ void set(int slot, double v)
    {
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = (slot == k) ? v : f[k];
    }
    // Consume one input; returns true and the smoothed value once cnt > r.  moving_average.rs:58-70.
    __device__ __forceinline__ bool push(double v, double &out)
    {
        constexpr int R = W / 2;
        const bool emits = cnt >= R;
        ++cnt;
        sum = __dadd_rn(sum, v);
        if (!emits) {  // preload (:58-61): push, no division update, no output
            int slot = head + len; slot = slot >= W ? slot - W : slot;
            set(slot, v);
            ++len;
            return false;
        }
        if (len == W) {  // circular_buffer.rs:35-40: pop the oldest, push the new value into its slot
            const double popped = get(head);
            set(head, v);
            head = (head + 1 == W) ? 0 : head + 1;
            sum = __dsub_rn(sum, popped);
        } else {
            int slot = head + len; slot = slot >= W ? slot - W : slot;
            set(slot, v);
            ++len;
            div = __ddiv_rn(1.0, (double)len);
        }
        out = __dmul_rn(sum, div);
        return true;
    }
    // One step of the shrinking tail (:71-79).
    __device__ __forceinline__ bool tail_pop(double &out)
    {
        if (len <= 0) return false;
        const double popped = get(head);
        head = (head + 1 == W) ? 0 : head + 1;
        --len;
        sum = __dsub_rn(sum, popped);
        div = __ddiv_rn(1.0, (double)len);
        out = __dmul_rn(sum, div);
        return true;
    }
    // Rotate the ring so that the oldest element sits in slot 0 (only meaningful when full).
    __device__ __forceinline__ void normalize()
    {
        double g[W];
#pragma unroll
        for (int k = 0; k < W; ++k) { int s = head + k; s = s >= W ? s - W : s; g[k] = get(s); }
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = g[k];
        head = 0;
    }
};

constexpr int SM_STAGES = 4;    // input tiles in flight per warp
constexpr int SM_OSTAGES = 3;   // output tiles in the ring
constexpr int SM_THREADS = 32;  // one warp per CTA

template <int W> struct SmoothTile { static constexpr int T = (W == 3) ? 96 : (W == 5) ? 80 : 84; };

template <int W>
constexpr size_t smooth_fast_smem_bytes()
{
    return (size_t)(SM_STAGES + SM_OSTAGES) * 32 * (SmoothTile<W>::T + 2) * 8 + SM_STAGES * 8 + 64;
}

template <int W, int I>
__global__ void __launch_bounds__(SM_THREADS, 1)
smooth_fast_kernel(const SpecDesc *__restrict__ sd, int n_spec)
{
    constexpr int T = SmoothTile<W>::T;
    constexpr int STRIDE = T + 2;            // doubles; (T+2)*8 bytes is an odd multiple of 16
    constexpr int R = W / 2;
    constexpr int L = I * R;                 // output lag of the pass pipeline
    static_assert(T % W == 0 && L <= T, "tile must hold whole ring rotations and the pipeline lag");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *in_buf = reinterpret_cast<double *>(smem_raw);                  // [SM_STAGES][32][STRIDE]
    double *out_buf = in_buf + SM_STAGES * 32 * STRIDE;                     // [32][SM_OSTAGES][STRIDE]... see orow
    uint64_t *full = reinterpret_cast<uint64_t *>(out_buf + SM_OSTAGES * 32 * STRIDE);

    const int lane = threadIdx.x;
    const int s = blockIdx.x * 32 + lane;
    const bool valid = s < n_spec;
    const double *__restrict__ y = valid ? sd[s].y : nullptr;
    double *__restrict__ ys = valid ? sd[s].ys : nullptr;
    const int n = valid ? sd[s].n : 0;
    const int n_max = __reduce_max_sync(0xffffffffu, n);
    const int n_min = __reduce_min_sync(0xffffffffu, valid ? n : 0x7fffffff);
    const int tiles = (n_max + T - 1) / T;

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < SM_STAGES; ++q) mbar_init(&full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();

    auto issue_load = [&](int k) {
        const int stage = k % SM_STAGES;
        int cnt = n - k * T;
        cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
        const uint32_t bytes = (uint32_t)(cnt & ~1) * 8u;  // bulk copies move multiples of 16 bytes
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_expect_tx(&full[stage], total);
        __syncwarp();
        if (bytes) bulk_g2s(in_buf + (stage * 32 + lane) * STRIDE, y + (size_t)k * T, bytes, &full[stage]);
    };
    auto orow = [&](int tile) -> double * { return out_buf + ((tile % SM_OSTAGES) * 32 + lane) * STRIDE; };
    auto flush_tile = [&](int j) {
        // generic-proxy writes -> async-proxy reads
        fence_proxy_async_smem();
        __syncwarp();
        int cnt = n - j * T;
        cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
        const uint32_t bytes = (uint32_t)(cnt & ~1) * 8u;
        if (bytes) bulk_s2g(ys + (size_t)j * T, orow(j), bytes);
        if (cnt & 1) ys[(size_t)j * T + cnt - 1] = orow(j)[cnt - 1];
        bulk_commit();
    };

    for (int k = 0; k < SM_STAGES && k < tiles; ++k) issue_load(k);

    MAPass<W> pass[I];
#pragma unroll
    for (int p = 0; p < I; ++p) pass[p].reset();
    bool done = !valid;
    int o = 0;  // next output index of this lane

    // feed v into pass p0 and push whatever comes out down the remaining passes
    auto feed = [&](int p0, double v) {
        bool have = true;
#pragma unroll
        for (int p = 0; p < I; ++p) {
            if (p < p0 || !have) continue;
            double out;
            have = pass[p].push(v, out);
            v = out;
        }
        if (have) {
            orow(o / T)[o % T] = v;
            ++o;
        }
    };

    for (int k = 0; k < tiles; ++k) {
        const int stage = k % SM_STAGES;
        while (!mbar_try_wait(&full[stage], (uint32_t)((k / SM_STAGES) & 1))) {}
        const double *__restrict__ irow = in_buf + (stage * 32 + lane) * STRIDE;
        {   // odd trailing element of a row: not part of the 16-byte-granular bulk copy
            int cnt = n - k * T;
            cnt = cnt < 0 ? 0 : (cnt > T ? T : cnt);
            if (cnt & 1) const_cast<double *>(irow)[cnt - 1] = y[(size_t)k * T + cnt - 1];
        }
        // the ring slot that out tile k+1 will use was flushed two tiles ago: make sure it was read
        bulk_wait_read<1>();
        __syncwarp();

        const bool fast = (k >= 1) && ((k + 1) * T <= n_min) && (k * T - (I - 1) * R >= W);
        if (fast) {
            // Steady state: every pass is full, so after normalisation ring slots are static.
            // The passes are SKEWED by one step: at step tau pass p consumes what pass p-1 emitted
            // at step tau-1 (kept in carry[p]).  Within a step the I passes are then independent
            // dependency chains (the only loop-carried dependence is each pass's own running sum),
            // which is what hides the FP64 latency.  Pass p is active for p <= tau < T + p, so
            // every pass still consumes exactly T inputs per tile and the skew never leaves the tile.
#pragma unroll
            for (int p = 0; p < I; ++p) pass[p].normalize();
            double carry[I + 1];
#pragma unroll
            for (int p = 0; p <= I; ++p) carry[p] = 0.0;
            double *__restrict__ pa = orow(k - 1) + (T - L);  // outputs j' <  L land in the previous out tile
            double *__restrict__ pb = orow(k) - L;            // outputs j' >= L land in this one
            // one skewed step; TAU_MOD = tau mod W and the active range [p_lo, p_hi] are static
#define MDB_SKEW_STEP(TAU_MOD, P_LO, P_HI, VIN)                                                   \
            {                                                                                     \
                _Pragma("unroll") for (int p = I - 1; p >= 0; --p) {                               \
                    if (p < (P_LO) || p > (P_HI)) continue;                                        \
                    const int slot = (((TAU_MOD) - p) % W + W) % W;                                \
                    const double vin_ = (p == 0) ? (VIN) : carry[p];                               \
                    pass[p].sum = __dsub_rn(__dadd_rn(pass[p].sum, vin_), pass[p].f[slot]);        \
                    pass[p].f[slot] = vin_;                                                        \
                    carry[p + 1] = __dmul_rn(pass[p].sum, pass[p].div);                            \
                }                                                                                  \
            }
            // prologue: tau = 0 .. I-2, passes 0..tau active, nothing leaves the last pass yet
#pragma unroll
            for (int tau = 0; tau < I - 1; ++tau) MDB_SKEW_STEP(tau % W, 0, tau, irow[tau])
            // main: tau = I-1 .. T-1, all passes active; out index within the tile j' = tau - (I-1)
            constexpr int TAU0 = I - 1;
            constexpr int M = T - TAU0;
            constexpr int G = M / W;
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const double *__restrict__ ig = irow + TAU0 + g * W;
                const int jg = g * W;
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    MDB_SKEW_STEP((TAU0 + u) % W, 0, I - 1, ig[u])
                    const int jj = jg + u;
                    double *q = (jj < L) ? pa : pb;
                    q[jj] = carry[I];
                }
            }
#pragma unroll
            for (int u = 0; u < M % W; ++u) {  // main steps that do not fill a whole ring rotation
                MDB_SKEW_STEP((TAU0 + G * W + u) % W, 0, I - 1, irow[TAU0 + G * W + u])
                const int jj = G * W + u;
                double *q = (jj < L) ? pa : pb;
                q[jj] = carry[I];
            }
            // epilogue: tau = T .. T+I-2, passes tau-T+1 .. I-1 still have one input pending
#pragma unroll
            for (int e = 0; e < I - 1; ++e) {
                MDB_SKEW_STEP((T + e) % W, e + 1, I - 1, 0.0)
                const int jj = T + e - TAU0;
                double *q = (jj < L) ? pa : pb;
                q[jj] = carry[I];
            }
#undef MDB_SKEW_STEP
#pragma unroll
            for (int p = 0; p < I; ++p) pass[p].cnt += T;
            o += T;
        } else {
            for (int j = 0; j < T; ++j) {
                const int t = k * T + j;
                if (t < n) feed(0, irow[j]);
            }
        }
        if (!done && (k + 1) * T >= n) {
            // shrinking tails (:71-79): pass p drains completely before pass p+1 starts its own.
            // Runs after whichever path consumed the lane's last input (a fast tile leaves the
            // rings normalised with head = 0, which tail_pop handles like any other state).
#pragma unroll
            for (int p = 0; p < I; ++p) {
                for (int q = 0; q < R; ++q) {
                    double out;
                    if (pass[p].tail_pop(out)) {
                        if (p + 1 < I) feed(p + 1, out);
                        else { orow(o / T)[o % T] = out; ++o; }
                    }
                }
            }
            done = true;
        }
        __syncwarp();
        if (k + SM_STAGES < tiles) issue_load(k + SM_STAGES);
        if (k >= 1) flush_tile(k - 1);
    }
    if (tiles >= 1) flush_tile(tiles - 1);
    bulk_wait_read<0>();
}

using SmoothFastFn = void (*)(const SpecDesc *, int);

template <int W, int I> struct SmoothFastEntry {
    static SmoothFastFn fn() { return smooth_fast_kernel<W, I>; }
};

// Returns the kernel for (window, iterations) or nullptr when the settings need the generic path.
inline SmoothFastFn smooth_fast_lookup(int window, int iterations, size_t *smem_bytes)
{
#define MDB_SF_CASE(Wv, Iv) \
    if (window == Wv && iterations == Iv) { *smem_bytes = smooth_fast_smem_bytes<Wv>(); return smooth_fast_kernel<Wv, Iv>; }
#define MDB_SF_ROW(Wv) \
    MDB_SF_CASE(Wv, 1) MDB_SF_CASE(Wv, 2) MDB_SF_CASE(Wv, 3) MDB_SF_CASE(Wv, 4) MDB_SF_CASE(Wv, 5) \
    MDB_SF_CASE(Wv, 6) MDB_SF_CASE(Wv, 7) MDB_SF_CASE(Wv, 8) MDB_SF_CASE(Wv, 9) MDB_SF_CASE(Wv, 10)
    MDB_SF_ROW(3)
    MDB_SF_ROW(5)
    MDB_SF_ROW(7)
#undef MDB_SF_ROW
#undef MDB_SF_CASE
    return nullptr;
}

}  // namespace mdb
