// smooth_split.cuh -- K1 for a FEW long spectra, second latency form: one chain warp PER PASS and the
// multiply taken off the chain.
//
// smooth_stream.cuh runs every pass in one warp (lane p = pass p); its interior loop is down to 23.8 cycles
// per point, of which 16.4 are the two dependent additions of moving_average.rs:62-70 and about 4 the multiply
// `sum * div`, which shares the FP64 pipe -- and the in-order instruction stream -- with the chain.  Here the
// output of a pass leaves its chain warp as the RAW running sum:
//
//   global y --movers--> X[0] --chain 0--> S[0] --movers: * div--> X[1] --chain 1--> S[1] ... --movers: * div--> global ys
//
//   X[p], S[p]: rings of SPLIT_R points in shared memory (element i at slot i & (R-1)).
//   chain warp p: ONE thread (its other lanes exit): 2 dependent DADD per point, four 16-byte loads and four
//   16-byte stores per 8 points, aligned 8-point blocks m-1, m, m+1 of X[p] held in registers (the window
//   size is a template parameter, so x[i+r] and x[i+r-w] are register operands chosen at compile time).
//   Blocks in which the window is still filling or already shrinking (`div` changes there) are done element
//   by element and stored as finished products; the movers know which blocks those are.
//   mover warps: 96 threads that feed X[0], turn S[p] into X[p+1] (interior elements: one rounded multiply
//   by div = RN(1/w), exactly the chain's own `sum * div`) and drain S[last] to global memory, 768 elements
//   per step and stage, all lanes busy.
// Stages meet through monotone counters in shared memory (one writer each), polled once per group of rounds,
// with a fence on either side.  Arithmetic and operation order per pass are those of smooth_stream.cuh /
// moving_average.rs:53-83; tests/test_gpu_parity.py compares the three forms and the oracle bit for bit.
#pragma once
#include "kernels.cuh"
#include "smooth_stream.cuh"

namespace mdb {

constexpr int SPLIT_R = 2048;          // ring length in points
constexpr int SPLIT_U = 8;             // points per block
constexpr int SPLIT_G = 64;            // blocks between flow-control checks of a chain warp
constexpr int SPLIT_MOVERS = 96;       // mover threads
constexpr int SPLIT_MAX_ITERS = 6;     // 2 * iters rings of 16 KB
constexpr int SPLIT_STEP = 8 * SPLIT_MOVERS;  // elements per mover step and stage

__host__ __device__ inline size_t smooth_split_smem_bytes(int iters) { return (size_t)2 * iters * SPLIT_R * 8; }
inline int smooth_split_threads(int iters) { return 32 * iters + SPLIT_MOVERS; }

struct SplitFlow {
    int x_done[SPLIT_MAX_ITERS + 1];   // points present in X[p] (x_done[0]: fed by the movers; x_done[iters]: drained to global)
    int c_done[SPLIT_MAX_ITERS];       // points of S[p] written by chain warp p
};

template <int WS>
__global__ void __launch_bounds__(32 * SPLIT_MAX_ITERS + SPLIT_MOVERS)
smooth_split_kernel(const SpecDesc *__restrict__ sd, int iters)
{
    extern __shared__ __align__(16) unsigned char split_smem[];
    __shared__ SplitFlow flow;
    __shared__ int mover_sample[SPLIT_MAX_ITERS];
    constexpr int U = SPLIT_U, R = SPLIT_R, G = SPLIT_G;
    constexpr int W = WS, RR = WS / 2, QB = WS - RR;   // x[i + RR] comes in, x[i - QB] goes out
    static_assert(RR <= U && QB <= U && WS >= 2, "window wider than two blocks");
    const SpecDesc d = sd[blockIdx.x];
    const int n = d.n;
    const int t = threadIdx.x;
    double *xr = reinterpret_cast<double *>(split_smem);             // X[p] at xr + p * R
    double *sr = xr + (size_t)iters * R;                             // S[p] at sr + p * R
    const int n_blocks = (n + U - 1) / U;
    // interior blocks: window full at the first element (i0 >= w - r) and no tail inside (i0 + U <= n - r)
    const int f0 = (W - RR + U - 1) / U, f1 = (n - RR) / U - 1;
    if (t < SPLIT_MAX_ITERS + 1) flow.x_done[t] = 0;
    if (t < SPLIT_MAX_ITERS) flow.c_done[t] = 0;
    __syncthreads();

    const int warp = t >> 5;
    if (warp < iters) {
        // ------------------------------------------------------------------ chain warp of pass `warp`
        if ((t & 31) != 0) return;
        const int p = warp;
        const uint32_t in_s = smem_addr(xr + (size_t)p * R), out_s = smem_addr(sr + (size_t)p * R);
        const int *avail_p = &flow.x_done[p];          // points of X[p] present
        const int *taken_p = &flow.x_done[p + 1];      // points of S[p] consumed (converted / drained) by the movers
        int *done_p = &flow.c_done[p];
        // rounds [.., b_end) may run: inputs up to `need_in` present, S[p] has room up to block b_end
        auto wait_for = [&](int need_in, int b_end) {
            need_in = min(need_in, n);
            const int need_taken = b_end * U - R;
            while (ld_volatile_s32(avail_p) < need_in || ld_volatile_s32(taken_p) < need_taken) __nanosleep(100);
            __threadfence_block();
        };
        auto publish = [&](int blocks_done) {
            __threadfence_block();
            st_volatile_s32(done_p, min(n, blocks_done * U));
        };
        double sum = 0.0, div = 1.0;
        // any block, element by element (edges included); products are stored
        auto edge_block = [&](int mb) {
            const int i0 = mb * U;
            if (mb == 0)
                for (int k = 0; k < RR; ++k) sum = __dadd_rn(sum, lds_f64(in_s + 8u * (uint32_t)(k & (R - 1))));
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u;
                if (i >= n) break;
                const int ai = i + RR, qi = ai - W;
                const double a = (ai < n) ? lds_f64(in_s + 8u * (uint32_t)(ai & (R - 1))) : -0.0;
                const double q = (qi >= 0) ? lds_f64(in_s + 8u * (uint32_t)(qi & (R - 1))) : 0.0;
                sum = __dadd_rn(sum, a);
                sum = __dsub_rn(sum, q);
                if (qi < 0 || ai >= n) div = __ddiv_rn(1.0, (double)(min(ai, n - 1) - max(qi, -1)));  // the window grows or shrinks
                sts_f64(out_s + 8u * (uint32_t)(i & (R - 1)), __dmul_rn(sum, div));
            }
        };
        int b = 0;
        if (f1 - f0 >= 8) {
            wait_for(f0 * U + RR, f0);
            for (; b < f0; ++b) edge_block(b);
            publish(b);
            // ---- interior: blocks f0 .. f1, raw sums stored
            // blocks m-1, m, m+1 feed round m; block m+2 is fetched during it.  (Fetching two rounds ahead with a fifth
            // register set was measured slower: 1.45 against 1.42 ms on blood_01.)
            double B0[U], B1[U], B2[U], B3[U];
            auto load_aligned = [&](double (&x)[U], int blk) {
                const uint32_t pa = in_s + 8u * ((uint32_t)(blk * U) & (uint32_t)(R - 1));
#pragma unroll
                for (int u = 0; u < U; u += 2) lds_f64x2(pa + 8u * u, x[u], x[u + 1]);
            };
            auto round = [&](const double (&pm)[U], const double (&p0)[U], const double (&p1)[U], double (&ld)[U], int m) {
                load_aligned(ld, m + 2);
                double o[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double a = (u + RR < U) ? p0[(u + RR) % U] : p1[(u + RR) % U];
                    const double q = (u >= QB) ? p0[(u + U - QB) % U] : pm[(u + U - QB) % U];
                    sum = __dadd_rn(sum, a);
                    sum = __dsub_rn(sum, q);
                    o[u] = sum;
                }
                const uint32_t po = out_s + 8u * ((uint32_t)(m * U) & (uint32_t)(R - 1));
#pragma unroll
                for (int u = 0; u < U; u += 2) sts_f64x2(po + 8u * u, o[u], o[u + 1]);
            };
            wait_for((min(f1 + 1, b + G) + 2) * U, min(f1 + 1, b + G));
            load_aligned(B0, b - 1);
            load_aligned(B1, b);
            load_aligned(B2, b + 1);
            while (b <= f1) {
                const int e = min(f1 + 1, b + G);
                wait_for((e + 2) * U, e);
                int left = e - b;
                for (; left >= 4; left -= 4, b += 4) {
                    round(B0, B1, B2, B3, b);
                    round(B1, B2, B3, B0, b + 1);
                    round(B2, B3, B0, B1, b + 2);
                    round(B3, B0, B1, B2, b + 3);
                }
                for (; left >= 1; --left, ++b) {
                    round(B0, B1, B2, B3, b);
#pragma unroll
                    for (int u = 0; u < U; ++u) { B0[u] = B1[u]; B1[u] = B2[u]; B2[u] = B3[u]; }
                }
                publish(b);
            }
        }
        // ---- the tail (or everything, for inputs too short to have an interior)
        for (; b < n_blocks; b += G) {
            const int e = min(n_blocks, b + G);
            wait_for(n, e);
            for (int bb = b; bb < e; ++bb) edge_block(bb);
            publish(e);
        }
        publish(n_blocks);
        return;
    }

    // ---------------------------------------------------------------------- movers
    const int m = t - 32 * iters;
    if (m >= SPLIT_MOVERS) return;
    const double *__restrict__ y = d.y;
    double *__restrict__ ys = d.ys;
    const double div_full = __ddiv_rn(1.0, (double)W);
    // element i of a pass output is a raw sum (to be multiplied) inside the interior blocks, a finished product elsewhere
    const bool has_interior = f1 - f0 >= 8;
    const int int_lo = has_interior ? f0 * U : n, int_hi = has_interior ? (f1 + 1) * U : n;
    int fed = 0, drained = 0;
    int conv[SPLIT_MAX_ITERS];      // conv[p]: elements of S[p-1] turned into X[p], p = 1 .. iters-1
#pragma unroll
    for (int p = 0; p < SPLIT_MAX_ITERS; ++p) conv[p] = 0;
    for (;;) {
        // one mover samples the chains' progress for all (every mover must take the same branches)
        if (m < iters) mover_sample[m] = ld_volatile_s32(&flow.c_done[m]);
        asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_MOVERS) : "memory");
        __threadfence_block();  // the chains' ring stores behind c_done are visible from here
        bool worked = false;
        // feed X[0]: chain 0 has finished mover_sample[0] points, its oldest live block is the one before
        {
            const int lo = max(0, mover_sample[0] / U - 1) * U;
            const int limit = min(n, lo + R);
            if (fed < limit) {
                const int hi = min(limit, fed + SPLIT_STEP);
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { const int i = fed + m + k * SPLIT_MOVERS; v[k] = (i < hi) ? __ldg(y + i) : 0.0; }
#pragma unroll
                for (int k = 0; k < 8; ++k) { const int i = fed + m + k * SPLIT_MOVERS; if (i < hi) xr[i & (R - 1)] = v[k]; }
                fed = hi;
                worked = true;
            }
        }
        // S[p-1] -> X[p]
        for (int p = 1; p < iters; ++p) {
            const int lo = max(0, mover_sample[p] / U - 1) * U;       // chain p's oldest live block
            const int limit = min(mover_sample[p - 1], lo + R);
            if (conv[p] < limit) {
                const int hi = min(limit, conv[p] + SPLIT_STEP);
                const double *src = sr + (size_t)(p - 1) * R;
                double *dst = xr + (size_t)p * R;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = conv[p] + m + k * SPLIT_MOVERS;
                    if (i < hi) {
                        const double s = src[i & (R - 1)];
                        dst[i & (R - 1)] = (i >= int_lo && i < int_hi) ? __dmul_rn(s, div_full) : s;
                    }
                }
                conv[p] = hi;
                worked = true;
            }
        }
        // drain S[last] to global memory
        {
            const int limit = mover_sample[iters - 1];
            if (drained < limit) {
                const int hi = min(limit, drained + SPLIT_STEP);
                const double *src = sr + (size_t)(iters - 1) * R;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = drained + m + k * SPLIT_MOVERS;
                    if (i < hi) {
                        const double s = src[i & (R - 1)];
                        ys[i] = (i >= int_lo && i < int_hi) ? __dmul_rn(s, div_full) : s;
                    }
                }
                drained = hi;
                worked = true;
            }
        }
        __threadfence_block();
        asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_MOVERS) : "memory");
        if (m == 0 && worked) {
            st_volatile_s32(&flow.x_done[0], fed);
            for (int p = 1; p < iters; ++p) st_volatile_s32(&flow.x_done[p], conv[p]);
            st_volatile_s32(&flow.x_done[iters], drained);
        }
        if (drained >= n) break;
        if (!worked) __nanosleep(200);
    }
}

}  // namespace mdb
