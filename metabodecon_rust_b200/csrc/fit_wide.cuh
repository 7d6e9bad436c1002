// fit_wide.cuh -- K6 for a FEW spectra: one refinement pass as two launches.
//
// fit_iter_kernel gives every peak one thread that walks all P Lorentzians at its three stencil
// points; a single spectrum of ~1 000 peaks is then 8 CTAs, each thread a chain of 1 000 ordered
// evaluations, and a pass takes 82 us however idle the other 140 SMs are.  Here a pass is split:
//   fit_wide_superpose_kernel  one thread per (stencil point, peak): three times the CTAs, and each
//                              thread interleaves 8 consecutive Lorentzians at its ONE point
//                              (lorentz_multi_q) instead of 3 points at one Lorentzian; it ends with
//                              the ratio and the rescaled stencil value (fitter_analytical.rs:42-54);
//   fit_wide_solve_kernel      one thread per peak: mirror and re-solve (:55-64).
// The sum over j stays one ordered chain per (point, peak); results are bit-identical to
// fit_iter_kernel (tests: the per-pass traces of mdb_stage_fit run through this form, batches
// through the other, both against the oracle).
#pragma once
#include "kernels.cuh"
#include "small_fused.cuh"

namespace mdb {

// Ordered superposition of Lorentzians [0, p) at ONE point per thread; tiles, barriers and the
// fast-domain test as in superpose_tiles.  All threads of the CTA must call.
template <int T>
__device__ __forceinline__ void superpose_tiles_point(unsigned char *smem, double *qs, const double *__restrict__ src, int p,
                                                      const double x, double &acc, uint32_t &tc)
{
    double(*tile)[3 * LOR_TILE] = reinterpret_cast<double(*)[3 * LOR_TILE]>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * 3 * LOR_TILE * sizeof(double));
    const int tid = threadIdx.x;
    if (tc == 0) {
        if (tid == 0) {
            mbarrier_init(&bar[0], 1);
            mbarrier_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const bool x_ok = x_fast_domain(x);
    const int ntiles = (p + LOR_TILE - 1) / LOR_TILE;
    auto issue = [&](int t) {
        const int cnt = min(LOR_TILE, p - t * LOR_TILE);
        const double *g = src + 3ll * t * LOR_TILE;
        const uint32_t slot = (tc + (uint32_t)t) & 1u;
        double *buf = tile[slot];
        const uint32_t bytes = (uint32_t)cnt * 24u, bulk = bytes & ~15u;
        if (bytes & 8u) buf[3 * cnt - 1] = __ldcg(g + 3 * cnt - 1);
        mbarrier_expect_tx(&bar[slot], bulk);
        if (bulk) tma_bulk_g2s(buf, g, bulk, &bar[slot]);
    };
    if (tid == 0 && ntiles > 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
        const int cnt = min(LOR_TILE, p - t * LOR_TILE);
        const uint32_t seq = tc + (uint32_t)t, slot = seq & 1u;
        if (tid == 0 && t + 1 < ntiles) issue(t + 1);
        mbarrier_wait(&bar[slot], (seq >> 1) & 1u);
        const double *s = tile[slot];
        bool ok = x_ok;
        for (int j = tid; j < cnt; j += T) ok = ok && params_fast_domain(s[3 * j], s[3 * j + 1], s[3 * j + 2]);
        if (__syncthreads_and(ok)) {
            int j = 0;
            if (cnt >= 8) {
                // Groups of 8 Lorentzians.  The quotients of a group go through the thread's own column
                // of shared memory and join the ordered sum one group later: with the eight division
                // chains all ending in a store, ptxas gives them equal priority and interleaves them
                // (ending in the ordered adds they get descending priority and are emitted one after
                // the other -- measured: 8 MUFU.RCP64H spread over 340 instructions, 20 % issue rate),
                // and the adds of the previous group fill their stalls.
                double q[8];
                int b = 0;
                lorentz_multi_q<8>(s, x, q);
#pragma unroll
                for (int u = 0; u < 8; ++u) qs[(0 * 8 + u) * T + tid] = q[u];
#pragma unroll 1
                for (j = 8; j + 8 <= cnt; j += 8) {
                    double pq[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) pq[u] = qs[(b * 8 + u) * T + tid];
                    lorentz_multi_q<8>(s + 3 * j, x, q);
#pragma unroll
                    for (int u = 0; u < 8; ++u) qs[((b ^ 1) * 8 + u) * T + tid] = q[u];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, pq[u]);
                    b ^= 1;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, qs[(b * 8 + u) * T + tid]);
            }
            if (j + 4 <= cnt) { lorentz_multi<4>(s + 3 * j, x, acc); j += 4; }
            if (j + 2 <= cnt) { lorentz_multi<2>(s + 3 * j, x, acc); j += 2; }
            if (j < cnt) lorentz_multi<1>(s + 3 * j, x, acc);
        } else {
            const double xs[1] = {x};
            double as[1] = {acc};
#pragma unroll 1
            for (int j = 0; j < cnt; ++j) lorentz_step<1, false>(s[3 * j], s[3 * j + 1], s[3 * j + 2], xs, as);
            acc = as[0];
        }
        __syncthreads();
    }
    tc += (uint32_t)ntiles;
}

// grid (ceil(max_peaks / FIT_THREADS), spectra, 3): blockIdx.z is the stencil point
__global__ void __launch_bounds__(FIT_THREADS)
fit_wide_superpose_kernel(const FitDesc *__restrict__ fd, FitState st, double *__restrict__ yn, long long yn_stride, int it)
{
    extern __shared__ __align__(128) unsigned char lor_smem[];
    __shared__ double quot_smem[2 * 8 * FIT_THREADS];  // [buffer][slot][thread]: a column per thread, conflict-free
    const FitDesc f = fd[blockIdx.y];
    if (blockIdx.x * FIT_THREADS >= f.n_peaks || it >= f.n_iters) return;
    const double *__restrict__ pin = (it & 1) ? st.pb : st.pa;
    const int q = blockIdx.z;
    const double *__restrict__ ox = q == 0 ? st.ox1 : (q == 1 ? st.ox2 : st.ox3);
    const double *__restrict__ oy = q == 0 ? st.oy1 : (q == 1 ? st.oy2 : st.oy3);
    const double *__restrict__ sy = q == 0 ? st.sy1 : (q == 1 ? st.sy2 : st.sy3);
    const int k = blockIdx.x * FIT_THREADS + threadIdx.x;
    const bool active = k < f.n_peaks;
    const long long g = f.off + (active ? k : 0);
    const double x = ox[g];
    double acc = 0.0;
    uint32_t tc = 0;
    superpose_tiles_point<FIT_THREADS>(lor_smem, quot_smem, pin + 3 * f.off, f.n_peaks, x, acc, tc);
    if (!active) return;
    // ratio = y_orig / superposition (:42-47); y_k = y_k * ratio_k (:52-54)
    yn[(long long)q * yn_stride + g] = __dmul_rn(sy[g], __ddiv_rn(oy[g], acc));
}

// ---------------------------------------------------------------------------------------------
// The superposition step once more, with thread-level instead of instruction-level parallelism.
// A chain (stencil point, peak) is an ordered sum over all P Lorentzians, but only the SUM is
// ordered: the quotients are independent.  A CTA owns WIDE2_CHAINS chains; seven producer threads
// per chain evaluate 8 quotients each of a 56-Lorentzian tile into shared memory, and one
// accumulator thread per chain adds the finished tile in index order while the producers are
// already on the next one (two tile buffers, named barriers full[b] / empty[b]).  3 P chains x P
// quotients then keep ~50 SMs x 14 producer warps busy instead of 24 SMs x 4 warps of 8-deep ILP.
// Every quotient and every addition is the same operation as in fit_iter_kernel.
// ---------------------------------------------------------------------------------------------
// Round 2 shape: 32 chains per CTA (one warp wide), eight warps.  Warp 0 holds the 32 accumulator threads, warps
// 1-3 and 5-7 are the producers of phases 0-5 (48 Lorentzians per tile), warp 4 exits: warps are dealt to the
// SM's four schedulers by warp index modulo 4, so the accumulator warp has scheduler 0 -- and its FP64 issue
// slots -- to itself, and its chain of dependent additions (the critical path: P additions of 8.2 cycles each)
// is not queued behind the producers' divisions.  Half the chains per CTA also spreads one blood spectrum over
// 93 SMs instead of 48.  (64 chains x 7 producer phases in 16 warps, the round-1 shape: 30 us per pass; with the
// parameters in shared memory 22 us; this shape: see DESIGN.md section 4.)
constexpr int WIDE2_CHAINS = 32;
constexpr int WIDE2_PHASES = 6;
constexpr int WIDE2_TILE = 8 * WIDE2_PHASES;                          // Lorentzians per tile
constexpr int WIDE2_THREADS = 8 * 32;                                 // warp 0: accumulators, 1-3 / 5-7: producers, 4: idle
constexpr int WIDE2_BAR_THREADS = (1 + WIDE2_PHASES) * 32;            // threads that take part in the named barriers
constexpr size_t WIDE2_SMEM = (size_t)2 * WIDE2_TILE * WIDE2_CHAINS * sizeof(double);

// PARAMS_IN_SMEM: the spectrum's whole parameter set (24 bytes per Lorentzian) is copied into shared memory once,
// behind the quotient buffers, and the producers read their eight Lorentzians per tile from there.  With the
// parameters fetched from L2 at the start of every tile the producers -- which move in lockstep with the tile
// barriers -- spent more time waiting for those loads than computing (ncu, blood_01: long-scoreboard stalls on
// top, FP64 pipe 31 % of active cycles, 30 us per pass).
// grid (ceil(max_peaks / WIDE2_CHAINS), spectra, 3): blockIdx.z is the stencil point
// `done` (may be null): one counter per (spectrum, peak block).  The three CTAs of a peak block (one per stencil
// point) each bump it once per pass when their rescaled stencil values are written; the CTA that draws the
// third ticket finds all three values of its 32 peaks in memory and does the mirror + solve step itself
// (fit_wide_solve_kernel's body), which saves a launch boundary per pass.  Counters only grow: ticket % 3.
template <bool PARAMS_IN_SMEM>
__global__ void __launch_bounds__(WIDE2_THREADS)
fit_wide2_superpose_kernel(const FitDesc *__restrict__ fd, FitState st, double *__restrict__ yn, long long yn_stride, int it,
                           int *__restrict__ done)
{
    extern __shared__ __align__(16) unsigned char wide2_smem[];
    double *quot = reinterpret_cast<double *>(wide2_smem);            // [buffer][slot in tile][chain]
    const FitDesc f = fd[blockIdx.y];
    if (blockIdx.x * WIDE2_CHAINS >= f.n_peaks || it >= f.n_iters) return;
    const double *__restrict__ pin = ((it & 1) ? st.pb : st.pa) + 3 * f.off;
    const double *sp = reinterpret_cast<const double *>(wide2_smem + WIDE2_SMEM);  // the parameter set (PARAMS_IN_SMEM)
    if (PARAMS_IN_SMEM) {
        double *w = reinterpret_cast<double *>(wide2_smem + WIDE2_SMEM);
        for (int i = threadIdx.x; i < 3 * f.n_peaks; i += WIDE2_THREADS) w[i] = __ldcg(pin + i);  // L2: written by the last solve launch
        __syncthreads();
    }
    const int q = blockIdx.z;
    const double *__restrict__ ox = q == 0 ? st.ox1 : (q == 1 ? st.ox2 : st.ox3);
    const int c = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 4) return;                                            // leaves scheduler 0 to the accumulator warp
    const int role = warp == 0 ? 0 : (warp < 4 ? warp : warp - 1);    // 0: accumulator, 1..6: producer of phase role - 1
    const int k = blockIdx.x * WIDE2_CHAINS + c;
    const bool active = k < f.n_peaks;
    const long long g = f.off + (active ? k : 0);
    const int P = f.n_peaks;
    const int n_tiles = (P + WIDE2_TILE - 1) / WIDE2_TILE;
    // named barriers 1, 2: full[0], full[1]; 3, 4: empty[0], empty[1]; every use counts the seven working warps
    auto bar_sync = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(WIDE2_BAR_THREADS) : "memory"); };
    auto bar_arrive = [](int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(WIDE2_BAR_THREADS) : "memory"); };
    if (role == 0) {
        double acc = 0.0;
        for (int t = 0; t < n_tiles; ++t) {
            const int b = t & 1, cnt = min(WIDE2_TILE, P - t * WIDE2_TILE);
            bar_sync(1 + b);                                          // tile t is in buffer b
            const double *src = quot + ((size_t)b * WIDE2_TILE) * WIDE2_CHAINS + c;
            if (cnt == WIDE2_TILE) {
                // a full tile: all 48 quotients are fetched up front (independent loads, one latency), so the chain of
                // additions never waits for shared memory
                double v[WIDE2_TILE];
#pragma unroll
                for (int u = 0; u < WIDE2_TILE; ++u) v[u] = src[(size_t)u * WIDE2_CHAINS];
                bar_arrive(3 + b);                                    // buffer b may be refilled: its values are in registers
#pragma unroll
                for (int u = 0; u < WIDE2_TILE; ++u) acc = __dadd_rn(acc, v[u]);
                continue;
            }
            int jj = 0;
            for (; jj + 8 <= cnt; jj += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(jj + u) * WIDE2_CHAINS];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, v[u]);
            }
            for (; jj < cnt; ++jj) acc = __dadd_rn(acc, src[(size_t)jj * WIDE2_CHAINS]);
            bar_arrive(3 + b);                                        // buffer b may be refilled
        }
        if (active) {
            const double *__restrict__ oy = q == 0 ? st.oy1 : (q == 1 ? st.oy2 : st.oy3);
            const double *__restrict__ sy = q == 0 ? st.sy1 : (q == 1 ? st.sy2 : st.sy3);
            yn[(long long)q * yn_stride + g] = __dmul_rn(sy[g], __ddiv_rn(oy[g], acc));  // :42-54
        }
        if (done) {  // the whole accumulator warp is here (role 0 = warp 0)
            __threadfence();  // this CTA's stencil values are visible before its ticket is
            int ticket = 0;
            if (c == 0) ticket = atomicAdd(&done[(size_t)blockIdx.y * gridDim.x + blockIdx.x], 1);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket % 3 == 2) {
                __threadfence();
                if (active) {  // fit_wide_solve_kernel, for this block's peaks
                    double *__restrict__ pout = (it & 1) ? st.pa : st.pb;
                    Stencil p;
                    p.x1 = st.sx1[g]; p.x2 = st.ox2[g]; p.x3 = st.sx3[g];
                    p.y1 = __ldcg(yn + g); p.y2 = __ldcg(yn + yn_stride + g); p.y3 = __ldcg(yn + 2 * yn_stride + g);
                    mirror_shoulder(p);
                    st.sx1[g] = p.x1; st.sx3[g] = p.x3;
                    st.sy1[g] = p.y1; st.sy2[g] = p.y2; st.sy3[g] = p.y3;
                    double sfhw, hw2, maxp;
                    solve_stencil(p, sfhw, hw2, maxp);  // :61-64
                    pout[3 * g] = sfhw; pout[3 * g + 1] = hw2; pout[3 * g + 2] = maxp;
                }
            }
        }
    } else {
        const int ph = role - 1;
        const double x = ox[g];
        const bool x_ok = x_fast_domain(x);
        for (int t = 0; t < n_tiles; ++t) {
            const int b = t & 1;
            if (t >= 2) bar_sync(3 + b);                              // the accumulators are done with tile t-2
            const int j0 = t * WIDE2_TILE + ph * 8;                   // this thread's 8 Lorentzians
            double *dst = quot + ((size_t)b * WIDE2_TILE + ph * 8) * WIDE2_CHAINS + c;
            if (j0 + 8 <= P) {
                double prm[24];
                if (PARAMS_IN_SMEM) {
#pragma unroll
                    for (int i = 0; i < 24; ++i) prm[i] = sp[3 * j0 + i];
                } else {
#pragma unroll
                    for (int i = 0; i < 24; ++i) prm[i] = __ldg(pin + 3 * j0 + i);
                }
                bool ok = x_ok;
#pragma unroll
                for (int u = 0; u < 8; ++u) ok = ok && params_fast_domain(prm[3 * u], prm[3 * u + 1], prm[3 * u + 2]);
                double qv[8];
                if (ok) {
                    lorentz_multi_q<8>(prm, x, qv);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {  // unrolled: prm / qv stay in registers
                        const double d0 = __dsub_rn(x, prm[3 * u + 2]);
                        qv[u] = __ddiv_rn(prm[3 * u], __dadd_rn(prm[3 * u + 1], __dmul_rn(d0, d0)));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) dst[(size_t)u * WIDE2_CHAINS] = qv[u];
            } else {
                for (int u = 0; j0 + u < P; ++u) {                    // the ragged end: one at a time, IEEE division
                    const double a = __ldg(pin + 3 * (j0 + u)), h = __ldg(pin + 3 * (j0 + u) + 1), m = __ldg(pin + 3 * (j0 + u) + 2);
                    const double d0 = __dsub_rn(x, m);
                    dst[(size_t)u * WIDE2_CHAINS] = __ddiv_rn(a, __dadd_rn(h, __dmul_rn(d0, d0)));
                }
            }
            bar_arrive(1 + b);
        }
        // the accumulators' last arrivals on empty[] need their partners
        for (int t = max(0, n_tiles - 2); t < n_tiles; ++t) bar_sync(3 + (t & 1));
    }
}

// grid (ceil(max_peaks / FIT_THREADS), spectra)
__global__ void __launch_bounds__(FIT_THREADS)
fit_wide_solve_kernel(const FitDesc *__restrict__ fd, FitState st, const double *__restrict__ yn, long long yn_stride, int it)
{
    const FitDesc f = fd[blockIdx.y];
    const int k = blockIdx.x * FIT_THREADS + threadIdx.x;
    if (k >= f.n_peaks || it >= f.n_iters) return;
    double *__restrict__ pout = (it & 1) ? st.pa : st.pb;
    const long long g = f.off + k;
    Stencil p;
    p.x1 = st.sx1[g]; p.x2 = st.ox2[g]; p.x3 = st.sx3[g];
    p.y1 = yn[g]; p.y2 = yn[yn_stride + g]; p.y3 = yn[2 * yn_stride + g];
    mirror_shoulder(p);
    st.sx1[g] = p.x1; st.sx3[g] = p.x3;
    st.sy1[g] = p.y1; st.sy2[g] = p.y2; st.sy3[g] = p.y3;
    double sfhw, hw2, maxp;
    solve_stencil(p, sfhw, hw2, maxp);  // :61-64
    pout[3 * g] = sfhw; pout[3 * g + 1] = hw2; pout[3 * g + 2] = maxp;
}

}  // namespace mdb
