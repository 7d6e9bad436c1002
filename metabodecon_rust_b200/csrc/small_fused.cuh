// small_fused.cuh -- the whole deconvolution of a SMALL spectrum in ONE launch.
//
// The chunked pipeline of api.cu (smooth -> detect -> select -> host round trip -> fit_init ->
// fit_iter x I_f -> retain -> superposition -> mse_reduce) is built for batches of 2^17-point spectra: ~20 launches
// and two host synchronisations per chunk, which is all a 2 048-point spectrum (the reference's
// `sim` bench set, benches/deconvoluter.rs:8-52) ever pays for.  Here one CTA owns one spectrum and
// walks the same stages back to back -- the moving average included, one warp with a lane per pass --
// with the whole smoothed row and second difference resident in shared memory:
// no tiles, no halos, no cold fallbacks, no counts travelling to the host between stages; results
// go straight into a page-locked host slot (zero-copy stores), so the call is one H2D copy, this
// launch and one stream synchronisation.
//
// Arithmetic is the same as the big kernels', operation for operation (same helpers where they
// exist), and every ordered sum is still a left fold: results are bit-identical to the general
// path and to the oracle (tests/test_gpu_parity.py::test_small_*).
//
// Reference: peak_selection/common.rs:5-40, detector.rs:99-164, scorer.rs:65-74,
// noise_score_filter.rs:32-54, 91-138, detector_only.rs:16-39, fitter_analytical.rs:19-72, 147-172,
// peak_stencil.rs:113-131, lorentzian.rs:546-548, 606-635, deconvoluter.rs:828-862.
#pragma once
#include "kernels.cuh"

namespace mdb {

constexpr int SMALL_MAX_N = 4096;     // longest spectrum the fused kernel takes
constexpr int SMALL_THREADS = 256;
constexpr int SMALL_WARPS = SMALL_THREADS / 32;
constexpr int SMALL_R = 4;            // points per thread in the MSE superposition

// Shared-memory map, in bytes, N = the launch's longest spectrum rounded up to a multiple of 8
// (cap = N/2 bounds the number of centres: two adjacent points can never both be centres):
//   selection phase                                  fit / MSE phase
//   [ 0, 6N)  sel   selected (l,c,r) triples          sel (kept for nothing but the order of writes)
//   [6N,14N)  d2    second difference                 [6N,18N)  parameter buffer A (cap x 24 B)
//   [14N,20N) pk    per-centre (l,c,r)                [18N,30N) parameter buffer B
//   [20N,24N) sc    per-centre score                  residuals of the MSE overlay the dead buffer
//   [24N,28N) sfr   dense ordered SFR scores
//   [28N,30N) cen   centre list, then rank/flag per centre
// and before any of that, while the spectrum is smoothed: [14N,22N) and [22N,30N) are the two
// ping-pong rows of the moving average (the last pass leaves its row for the second difference).
__host__ __device__ inline size_t small_smem_bytes(int n_al) { return (size_t)30 * n_al; }

struct SmallDesc {
    const int *ranges;      // MSE ranges of this spectrum, (start, end) pairs (deconvoluter.rs:829-845)
    double *fit_state;      // 14 * cap doubles of scratch: stencils and superposition values per peak
    unsigned char *out;     // result slot in mapped host memory: SmallOut, lor[cap], peaks[3 * cap]
    int n_ranges, n_iters, cap, skip;
};

struct SmallOut {           // header of a result slot (64 bytes)
    SelectOut info;
    int n_kept, pad_;
    double mse;
};

__host__ __device__ inline size_t small_slot_bytes(int cap) { return ((size_t)64 + (size_t)36 * cap + 63) & ~(size_t)63; }

// Order-preserving compaction of the indices of [begin, end) for which pred holds: emit(i, rank)
// is called once per kept index with its rank in ascending index order.  Every warp owns one
// contiguous run; ranks come from ballot/popc inside the run plus the counts of the runs before
// it.  pred is evaluated twice per index and must not depend on what emit writes for OTHER
// indices.  All threads of the CTA must call; returns the number kept.
template <class Pred, class Emit>
__device__ __forceinline__ int small_ordered_compact(int begin, int end, int *warp_cnt, Pred pred, Emit emit)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int seg = (end - begin + SMALL_WARPS - 1) / SMALL_WARPS;
    seg = (seg + 31) & ~31;
    const int b = begin + wid * seg, e = min(end, b + seg);
    int cnt = 0;
    for (int i0 = b; i0 < e; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < e && pred(i);
        cnt += __popc(__ballot_sync(0xffffffffu, f));
    }
    if (lane == 0) warp_cnt[wid] = cnt;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SMALL_WARPS; ++w) {
        if (w < wid) base += warp_cnt[w];
        tot += warp_cnt[w];
    }
    for (int i0 = b; i0 < e; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < e && pred(i);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (f) emit(i, base + __popc(bal & lt_mask));
        base += __popc(bal);
    }
    __syncthreads();
    return tot;
}

__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// U consecutive Lorentzians (AoS triples at p) at ONE point, fast-domain division: the U
// quotients are independent and interleave stage by stage (what lorentz_step does across points),
// then join the running sum in index order -- the same operations per element as lorentz_step.
template <int U>
__device__ __forceinline__ void lorentz_multi_q(const double *p, const double x, double (&out)[U])
{
    double a[U], den[U], r[U], e[U], q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a[u] = p[3 * u]; den[u] = __dsub_rn(x, p[3 * u + 2]); }
#pragma unroll
    for (int u = 0; u < U; ++u) den[u] = __dmul_rn(den[u], den[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) den[u] = __dadd_rn(p[3 * u + 1], den[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = rcp_seed(den[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-den[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(e[u], e[u], e[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(r[u], e[u], r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-den[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(r[u], e[u], r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = __dmul_rn(a[u], r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-den[u], q[u], a[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = fma(r[u], e[u], q[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = q[u];
}
template <int U>
__device__ __forceinline__ void lorentz_multi(const double *p, const double x, double &acc)
{
    double q[U];
    lorentz_multi_q<U>(p, x, q);
#pragma unroll
    for (int u = 0; u < U; ++u) acc = __dadd_rn(acc, q[u]);
}

// Squared residuals of R points per thread, points i0 + t + r * SMALL_THREADS below e0
// (deconvoluter.rs:846-855): ordered superposition of the kept Lorentzians, minus y, squared.
// ULP: the few-ulp evaluation of MDB_SUPERPOSITION_FAST (lorentz_step_ulp) inside the fast domain.
template <int R, bool ULP>
__device__ __forceinline__ void small_mse_block(const double *__restrict__ x, const double *__restrict__ y, const double *kept,
                                                int n_kept, bool kfast, int i0, int e0, double *resid)
{
    double xv[R], acc[R];
    int idx[R];
    bool xok = true;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        idx[r] = i0 + (int)threadIdx.x + r * SMALL_THREADS;
        xv[r] = (idx[r] < e0) ? x[idx[r]] : 0.0;
        acc[r] = 0.0;
        xok = xok && x_fast_domain(xv[r]);
    }
    if (idx[0] >= e0) return;
    if (kfast && xok) {
#pragma unroll 2
        for (int j = 0; j < n_kept; ++j) lorentz_step<R, ULP ? 2 : 1>(kept[3 * j], kept[3 * j + 1], kept[3 * j + 2], xv, acc);
    } else {
#pragma unroll 1
        for (int j = 0; j < n_kept; ++j) lorentz_step<R, 0>(kept[3 * j], kept[3 * j + 1], kept[3 * j + 2], xv, acc);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (idx[r] < e0) {
            const double dd = __dsub_rn(acc[r], y[idx[r]]);  // :852
            resid[idx[r]] = __dmul_rn(dd, dd);
        }
}

// Moving average of one spectrum by ONE warp (smoothing/moving_average.rs:53-83, the recurrence of
// smooth_pass_generic_kernel): lane p runs pass p, `db` blocks of U points behind lane p-1, reading
// the row that lane writes and writing the other row (pass p reads row p&1).  Every pass is still
// the reference's sequential running sum -- two dependent additions per point, 16.4 cycles -- but
// the passes overlap.  A single warp has nobody to hide behind, so everything that is not on the
// chain is kept out of its way:
//  * one code shape for every block.  A step is always `sum = sum + a; sum = sum - q; out = sum * div`
//    with a = in[i+r] or -0.0 when the window has run off the end (x + -0.0 == x for every x) and
//    q = in[i+r-w] or +0.0 while the window is still filling (x - 0.0 == x for every x), so the
//    edges cost a few predicated loads and divisions instead of a branchy per-point loop;
//  * interior blocks run in a lean loop: no edge tests, the inputs of block mb+1 are fetched into a
//    second register set while the chain of block mb runs (two rounds per iteration, sets swapped);
//  * lanes without a pass shadow lane 0 (stores to a sink), so the warp stays converged and the
//    per-round __syncwarp() is its cheap converged form;
//  * shared memory is addressed through 32-bit shared-space addresses held in registers (left to
//    itself the compiler re-derives the window base, S2UR SR_CgaCtaId, inside the loop and the
//    in-order warp waits for it).
// `div` follows the reference: recomputed as 1/len only in steps where the window grows or
// shrinks, carried otherwise.  Inputs shorter than the window (n < w) take the per-point form.
// Hazards: with the prefetch a lane reads, in round b, indices up to (mb+2)U-1+r of its input row,
// all written by its producer before round b when db >= 3 + (r-1)/U, and none written in round b
// itself; the lane behind it overwrites indices below (mb-db+1)U of that row, which this lane no
// longer reads when db*U >= w-r.  db = 2 + ceil(max(r, w-r)/U) satisfies both.
constexpr int SMALL_SMOOTH_U = 8;
constexpr int SMALL_SMOOTH_MAX_ITERS = 32;

// WS: the window size when the interior loop is specialised for it (3, 5, 7), 0 = any window (runtime w_arg).
// The specialised loop keeps the aligned 8-point blocks m-1, m, m+1 of the lane's input row in registers
// (incoming and outgoing values of every step are register operands chosen at compile time; four 16-byte
// loads per round instead of sixteen 8-byte ones), fetches block m+2 two rounds early and stores the
// outputs of block m-1 one round late, so that nothing issued in a round depends on that round's chain;
// that costs two more blocks of distance between consecutive passes (see smooth_stream.cuh).
template <int WS>
__device__ __forceinline__ void small_smooth_warp(double *rows, int row_stride, int n, int iters, int w_arg, double *sink)
{
    constexpr int U = SMALL_SMOOTH_U;
    const int lane = threadIdx.x & 31;
    const int w = WS > 0 ? WS : w_arg;
    const int r = w / 2;
    const int db = WS > 0 ? 4 : 2 + (max(r, w - r) + U - 1) / U;
    const int n_blocks = (n + U - 1) / U;
    // pass p reads row p&1 and writes the other one
    const int in_o = (lane & 1) * row_stride, out_o = row_stride - in_o;
    double sum = 0.0, div = 1.0;
    const int rounds = n_blocks + (iters - 1) * db;

    if (n < w) {  // the window never fills: per-point form, as smooth_pass_generic_kernel
        if (lane >= iters) return;  // the caller's CTA barrier collects the other lanes
        const unsigned mask = iters >= 32 ? 0xffffffffu : ((1u << iters) - 1u);
        int len = 0;
        for (int b = 0; b < rounds; ++b) {
            const int mb = b - lane * db;
            if (mb >= 0 && mb < n_blocks) {
                if (mb == 0) {
                    for (int k = 0; k < r && k < n; ++k) sum = __dadd_rn(sum, rows[in_o + k]);
                    len = n;
                }
                for (int u = 0; u < U; ++u) {
                    const int i = mb * U + u;
                    if (i >= n) break;
                    if (i < n - r) {
                        sum = __dadd_rn(sum, rows[in_o + i + r]);
                        div = __ddiv_rn(1.0, (double)(i + r + 1));
                        rows[out_o + i] = __dmul_rn(sum, div);
                    } else if (len > 0) {
                        sum = __dsub_rn(sum, rows[in_o + n - len]);
                        --len;
                        div = __ddiv_rn(1.0, (double)len);
                        rows[out_o + i] = __dmul_rn(sum, div);
                    }
                }
            }
            __syncwarp(mask);
        }
        return;
    }

    // Lanes without a pass run along as ghosts of lane 0 (same reads, stores dropped or sent to a
    // sink), so the warp never diverges for long and the per-round barrier is the converged
    // __syncwarp() fast path instead of a masked one (MATCH / REDUX / VOTE every round).
    const bool ghost = lane >= iters;
    const uint32_t rows_s = smem_addr(rows);
    uint32_t in_s = rows_s + (ghost ? 0u : 8u * (uint32_t)in_o), out_s = rows_s + 8u * (uint32_t)out_o;
    int blk0 = ghost ? 0 : -lane * db;  // this lane's block in round b is b + blk0
    asm volatile("" : "+r"(in_s), "+r"(out_s), "+r"(blk0));  // opaque: held in registers, never re-derived

    // any block, edges included (n >= w)
    auto edge_round = [&](int b) {
        const int mb = b + blk0;
        if (mb >= 0 && mb < n_blocks) {
            const int i0 = mb * U;
            if (mb == 0)
                for (int k = 0; k < r; ++k) sum = __dadd_rn(sum, lds_f64(in_s + 8u * k));
            double a[U], q[U], dv[U];
            bool redo[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ai = i0 + u + r, qi = ai - w;
                a[u] = (ai < n) ? lds_f64(in_s + 8u * ai) : -0.0;
                q[u] = (qi >= 0) ? lds_f64(in_s + 8u * qi) : 0.0;
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
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u < n && !ghost) sts_f64(out_s + 8u * (i0 + u), a[u]);
        }
        __syncwarp();
    };
    // one interior round: chain over the block held in (ca, cq), next block fetched into (na, nq);
    // pa_s: address of this lane's in[i0 + r], po_s: of its out[i0]
    uint32_t pa_s = 0, po_s = 0;
    const uint32_t w8 = 8u * (uint32_t)w, po_step = ghost ? 0u : 8u * U;
    auto lean_round = [&](double (&ca)[U], double (&cq)[U], double (&na)[U], double (&nq)[U], bool prefetch) {
        if (prefetch) {
#pragma unroll
            for (int u = 0; u < U; ++u) { na[u] = lds_f64(pa_s + 8u * (U + u)); nq[u] = lds_f64(pa_s + 8u * (U + u) - w8); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sum = __dadd_rn(sum, ca[u]);
            sum = __dsub_rn(sum, cq[u]);
            ca[u] = __dmul_rn(sum, div);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) sts_f64(po_s + 8u * u, ca[u]);
        pa_s += 8u * U;
        po_s += po_step;
        __syncwarp();
    };

    // interior blocks: window full (i0 >= w - r) and no tail (i0 + U <= n - r)
    const int f0 = (w - r + U - 1) / U, f1 = (n - r) / U - 1;
    const int b_lo = f0 + (iters - 1) * db, b_hi = f1;  // rounds in which EVERY pass is on an interior block
    int b = 0;
    if constexpr (WS > 0) {
    if (b_hi - b_lo >= 8) {
        for (; b < b_lo; ++b) edge_round(b);
        constexpr int RR = WS / 2, QB = WS - RR;          // x[i + RR] comes in, x[i - QB] goes out
        static_assert(RR <= U && QB <= U, "window wider than two blocks");
        double B0[U], B1[U], B2[U], B3[U];                // aligned blocks m-1, m, m+1, m+2 of the input row (rotating)
        double prA[U], prB[U];                            // outputs of block m-1, not yet stored: in prB between rounds
        int m = b + blk0;                                 // this lane's block in the coming round
        const int last_blk = row_stride / U - 1;          // rows are a multiple of U long; the prefetch never leaves the row
        const uint32_t sink_s = smem_addr(sink);
        uint32_t po_prev = 0;
        uint32_t po_s2 = ghost ? sink_s : out_s + 8u * (uint32_t)(m * U);
        const uint32_t po_step2 = ghost ? 0u : 8u * U;
        asm volatile("" : "+r"(po_s2));
        auto load_aligned = [&](double (&x)[U], int blk) {
            const uint32_t pa = in_s + 8u * (uint32_t)(min(blk, last_blk) * U);
#pragma unroll
            for (int u = 0; u < U; u += 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[u]), "=d"(x[u + 1]) : "r"(pa + 8u * u) : "memory");
        };
        auto store_pending = [&](const double (&ps)[U]) {
#pragma unroll
            for (int u = 0; u < U; u += 2) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(po_prev + 8u * u), "d"(ps[u]), "d"(ps[u + 1]) : "memory");
        };
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
            po_prev = po_s2;
            po_s2 += po_step2;
            ++m;
            __syncwarp();
        };
        load_aligned(B0, m - 1);
        load_aligned(B1, m);
        load_aligned(B2, m + 1);
        int left = b_hi - b + 1;  // interior rounds to go
        b = b_hi + 1;
        lean_round(B0, B1, B2, B3, prA, prB, false);  // first lean round: nothing to store yet
#pragma unroll
        for (int u = 0; u < U; ++u) { B0[u] = B1[u]; B1[u] = B2[u]; B2[u] = B3[u]; prB[u] = prA[u]; }
        --left;
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
        store_pending(prB);
        __syncwarp();
    }
    } else {
    if (b_hi - b_lo >= 3) {
        for (; b < b_lo; ++b) edge_round(b);
        double a0[U], q0[U], a1[U], q1[U];
        {
            const int i0 = (b + blk0) * U;
            pa_s = in_s + 8u * (uint32_t)(i0 + r);
            po_s = ghost ? smem_addr(sink) : out_s + 8u * (uint32_t)i0;
            asm volatile("" : "+r"(pa_s), "+r"(po_s));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { a0[u] = lds_f64(pa_s + 8u * u); q0[u] = lds_f64(pa_s + 8u * u - w8); }
        int left = b_hi - b + 1;  // interior rounds to go
        b = b_hi + 1;
        for (; left >= 2; left -= 2) {
            lean_round(a0, q0, a1, q1, true);
            lean_round(a1, q1, a0, q0, left > 2);
        }
        if (left == 1) lean_round(a0, q0, a1, q1, false);
    }
    }
    for (; b < rounds; ++b) edge_round(b);
}

__global__ void __launch_bounds__(SMALL_THREADS)
small_fused_kernel(const SpecDesc *__restrict__ sd, const SmallDesc *__restrict__ xd, int n_al, int selector_kind,
                   int smooth_iters, int smooth_window, int flags, long long *__restrict__ stamps)
{
    const int fast_mse = flags & 1;        // MDB_SUPERPOSITION_FAST tail
    const int generic_smooth = flags & 2;  // MDB_STREAM_GENERIC=1 (tests): the any-window smoothing loop
    // optional phase clock of CTA 0 (MDB_SMALL_STAMPS=1, development aid): SM cycles at phase ends
    int stamp_k = 0;
    auto stamp = [&]() { if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[stamp_k++] = clock64(); };
    stamp();
    extern __shared__ __align__(16) unsigned char small_smem[];
    __shared__ int warp_cnt[SMALL_WARPS];
    __shared__ __align__(16) double smooth_sink[SMALL_SMOOTH_U];
    __shared__ double mse_part[SMALL_WARPS];
    __shared__ int s_raw, s_np, s_c0, s_c1;
    __shared__ double s_thr, s_mean, s_sd;

    const SmallDesc e = xd[blockIdx.x];
    if (e.skip) return;
    const SpecDesc d = sd[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31;
    const int n = d.n;
    const size_t N = (size_t)n_al;

    int *sel = reinterpret_cast<int *>(small_smem);
    double *d2 = reinterpret_cast<double *>(small_smem + 6 * N);
    int *pk = reinterpret_cast<int *>(small_smem + 14 * N);
    double *sc = reinterpret_cast<double *>(small_smem + 20 * N);
    double *sfr = reinterpret_cast<double *>(small_smem + 24 * N);
    int *cen = reinterpret_cast<int *>(small_smem + 28 * N);  // centre list, later rank / flag per centre
    double *par_a = reinterpret_cast<double *>(small_smem + 6 * N);
    double *par_b = reinterpret_cast<double *>(small_smem + 18 * N);

    SmallOut *hdr = reinterpret_cast<SmallOut *>(e.out);
    double *lor_out = reinterpret_cast<double *>(e.out + 64);
    int *pk_out = reinterpret_cast<int *>(e.out + 64 + (size_t)24 * e.cap);

    SmallOut o;
    o.info.status = ST_OK; o.info.n_detected = 0; o.info.n_after_ignore = 0; o.info.n_selected = 0;
    o.info.region_left = 0; o.info.region_right = 0; o.info.n_sfr = 0; o.info.pad_ = 0;
    o.info.mean = 0.0; o.info.sd = 0.0;
    o.n_kept = 0; o.pad_ = 0; o.mse = 0.0;

    if (t == 0) { s_raw = 0; s_np = 0; s_c0 = 0; s_c1 = 0; }

    // ---- smoothing in shared memory (smooth_iters == 0: d.ys already holds the smoothed row)
    const double *ys = d.ys;
    if (smooth_iters > 0) {
        double *row0 = reinterpret_cast<double *>(small_smem + 14 * N);
        double *row1 = reinterpret_cast<double *>(small_smem + 22 * N);
        const double *__restrict__ y = d.y;
        for (int j = t; j < n; j += SMALL_THREADS) row0[j] = y[j];
        __syncthreads();
        if (t < 32) {
            if (smooth_window == 5 && generic_smooth == 0) small_smooth_warp<5>(row0, n_al, n, smooth_iters, 5, smooth_sink);
            else if (smooth_window == 3 && generic_smooth == 0) small_smooth_warp<3>(row0, n_al, n, smooth_iters, 3, smooth_sink);
            else if (smooth_window == 7 && generic_smooth == 0) small_smooth_warp<7>(row0, n_al, n, smooth_iters, 7, smooth_sink);
            else small_smooth_warp<0>(row0, n_al, n, smooth_iters, smooth_window, smooth_sink);
        }
        __syncthreads();
        ys = (smooth_iters & 1) ? row1 : row0;
    }
    stamp();  // 1: smoothed
    // ---- second difference (common.rs:8), all of it resident
    for (int j = t; j < n - 2; j += SMALL_THREADS) d2[j] = d2_at(ys, j);
    __syncthreads();

    // ---- centres (detector.rs:124), ascending
    const int n_cen = small_ordered_compact(
        2, n - 2, warp_cnt,
        [&](int c) { const double b1 = d2[c - 1]; return b1 < 0.0 && b1 < d2[c - 2] && b1 < d2[c]; },
        [&](int c, int rank) { cen[rank] = c; });
    stamp();  // 2: centres

    // ---- one thread per centre: borders (detector.rs:150-164), MinimumSum score (scorer.rs:65-74),
    // first retain (noise_score_filter.rs:41-48 / detector_only.rs:26-38) and the region-split counts
    {
        int raw = 0, np = 0, c0 = 0, c1 = 0;
        for (int j0 = 0; j0 < n_cen; j0 += SMALL_THREADS) {
            const int j = j0 + t;
            bool found = false, kept = false;
            int c = 0;
            if (j < n_cen) {
                c = cen[j];
                int ll = 0, rr = 0;
                double pa = d2[c - 1], pb = d2[c];
                double rsum = fabs(pa);
                for (int q = c + 1; q <= n - 3; ++q) {
                    const double pc = d2[q];
                    rsum = __dadd_rn(rsum, fabs(pb));
                    if (pb > pa && (pb >= pc || (pb < 0.0 && pc >= 0.0))) { rr = q; break; }
                    pa = pb; pb = pc;
                }
                if (rr != 0) {
                    double qc = d2[c - 1], qb = d2[c - 2];
                    for (int q = c - 1; q >= 2; --q) {
                        const double qa = d2[q - 2];
                        if (qb > qc && (qb >= qa || (qb < 0.0 && qa >= 0.0))) { ll = q; break; }
                        qc = qb; qb = qa;
                    }
                }
                found = (ll != 0 && rr != 0);  // detector.rs:105
                double score = 0.0;
                if (found) {
                    double lsum = 0.0;
                    for (int u = ll - 1; u <= c - 1; ++u) lsum = __dadd_rn(lsum, fabs(d2[u]));
                    score = fmin(lsum, rsum);
                    kept = candidate_kept(d, selector_kind, ll, rr);
                }
                pk[3 * j] = ll; pk[3 * j + 1] = c; pk[3 * j + 2] = rr;
                sc[j] = score;
                cen[j] = found ? (kept ? 1 : 0) : -1;  // only this thread ever reads cen[j] before the barrier
            }
            raw += __popc(__ballot_sync(0xffffffffu, found));
            np += __popc(__ballot_sync(0xffffffffu, kept));
            c0 += __popc(__ballot_sync(0xffffffffu, kept && c <= d.sb0));
            c1 += __popc(__ballot_sync(0xffffffffu, kept && c <= d.sb1));
        }
        if (lane == 0) {
            if (raw) atomicAdd(&s_raw, raw);
            if (np) atomicAdd(&s_np, np);
            if (c0) atomicAdd(&s_c0, c0);
            if (c1) atomicAdd(&s_c1, c1);
        }
    }
    __syncthreads();
    stamp();  // 3: borders + scores
    const int n_raw = s_raw, np = s_np;
    o.info.n_detected = n_raw;
    o.info.n_after_ignore = np;
    if (n_raw == 0) {  // detector.rs:107-109
        if (t == 0) { o.info.status = ST_NO_PEAKS; *hdr = o; }
        return;
    }

    int n_sel = 0;
    if (selector_kind == 0) {  // DetectorOnly: everything that survived the retains, in order
        n_sel = small_ordered_compact(
            0, n_cen, warp_cnt, [&](int j) { return cen[j] == 1; },
            [&](int j, int rank) {
                const int l = pk[3 * j], c = pk[3 * j + 1], r = pk[3 * j + 2];
                sel[3 * rank] = l; sel[3 * rank + 1] = c; sel[3 * rank + 2] = r;
                pk_out[3 * rank] = l; pk_out[3 * rank + 1] = c; pk_out[3 * rank + 2] = r;
            });
    } else {
        if (np == 0) {  // `peaks.len() - 1` underflows and the slicing panics (common.rs:37)
            if (t == 0) { o.info.status = ST_PANIC; *hdr = o; }
            return;
        }
        // region boundaries in the filtered list (common.rs:26-40)
        const int cnt0 = s_c0, cnt1 = s_c1;
        const int bl = (cnt0 < np) ? cnt0 : 0;
        const int cand_r = (cnt1 > bl) ? cnt1 : bl;
        const int br = (cand_r < np) ? cand_r : np - 1;
        o.info.region_left = bl; o.info.region_right = br;
        if (bl == 0 && br >= np) {  // noise_score_filter.rs:102-104
            if (t == 0) { o.info.status = ST_EMPTY_SFR; *hdr = o; }
            return;
        }
        if (bl == br) {             // noise_score_filter.rs:105-107
            if (t == 0) { o.info.status = ST_EMPTY_SIGNAL; *hdr = o; }
            return;
        }
        const int n_sfr = bl + (np - br);
        o.info.n_sfr = n_sfr;
        // ranks in the filtered list; SFR scores gathered densely: peaks[0..bl] then peaks[br..] (:109-113)
        small_ordered_compact(
            0, n_cen, warp_cnt, [&](int j) { return cen[j] >= 1; },
            [&](int j, int rank) {
                cen[j] = 2 + rank;
                if (rank < bl) sfr[rank] = sc[j];
                else if (rank >= br) sfr[bl + rank - br] = sc[j];
            });
        // ordered mean / sd (noise_score_filter.rs:129-138)
        if (t < 32) {
            const double total = warp_ordered_sum<0>(sfr, n_sfr, 0.0);
            const double mean = __ddiv_rn(total, (double)n_sfr);
            const double vs = warp_ordered_sum<1>(sfr, n_sfr, mean);
            const double sdv = __dsqrt_rn(__ddiv_rn(vs, (double)n_sfr));
            if (t == 0) {
                s_mean = mean; s_sd = sdv;
                s_thr = __dadd_rn(mean, __dmul_rn(d.threshold, sdv));  // :118, no FMA
            }
        }
        __syncthreads();
        const double thr = s_thr;
        o.info.mean = s_mean; o.info.sd = s_sd;
        n_sel = small_ordered_compact(
            0, n_cen, warp_cnt,
            [&](int j) { const int rk = cen[j] - 2; return rk >= bl && rk < br && sc[j] >= thr; },
            [&](int j, int rank) {
                const int l = pk[3 * j], c = pk[3 * j + 1], r = pk[3 * j + 2];
                sel[3 * rank] = l; sel[3 * rank + 1] = c; sel[3 * rank + 2] = r;
                pk_out[3 * rank] = l; pk_out[3 * rank + 1] = c; pk_out[3 * rank + 2] = r;
            });
        if (n_sel == 0) {  // :121-123
            if (t == 0) { o.info.status = ST_EMPTY_SIGNAL; *hdr = o; }
            return;
        }
    }
    o.info.n_selected = n_sel;
    stamp();  // 4: selected
    // (the compaction ended with a barrier: d2 / pk / sc / sfr / cen are dead from here on)

    // ---- fit (fitter_analytical.rs:19-72).  Parameters ping-pong in shared memory; per-peak stencil
    // state and the 3P superposition values live in the CTA's scratch rows (fit_state, 14 x cap).
    // A refinement pass has two steps: one thread per (stencil point, peak) evaluates the ordered
    // superposition there (3P independent chains instead of P), then one thread per peak forms the
    // ratios, mirrors and re-solves.
    const int P = n_sel;
    const size_t cap = (size_t)e.cap;
    double *ox = e.fit_state, *oy = ox + 3 * cap;           // original stencils: [point][peak]
    double *sx1 = oy + 3 * cap, *sx3 = sx1 + cap;           // current stencil x (x2 never changes)
    double *sy = sx3 + cap;                                 // current stencil y: [point][peak]
    double *sup = sy + 3 * cap;                             // rescaled stencil y of the running pass: [point][peak]
    bool ok = true;   // this thread's positions / parameters are inside div_fast's domain
    bool pok = true;
    for (int k = t; k < P; k += SMALL_THREADS) {
        const int l = sel[3 * k], c = sel[3 * k + 1], r = sel[3 * k + 2];
        Stencil p;
        p.x1 = d.x[l]; p.x2 = d.x[c]; p.x3 = d.x[r];
        p.y1 = d.y[l]; p.y2 = d.y[c]; p.y3 = d.y[r];
        ox[k] = p.x1; ox[cap + k] = p.x2; ox[2 * cap + k] = p.x3;
        oy[k] = p.y1; oy[cap + k] = p.y2; oy[2 * cap + k] = p.y3;
        ok = ok && x_fast_domain(p.x1) && x_fast_domain(p.x2) && x_fast_domain(p.x3);
        mirror_shoulder(p);
        sx1[k] = p.x1; sx3[k] = p.x3;
        sy[k] = p.y1; sy[cap + k] = p.y2; sy[2 * cap + k] = p.y3;
        double sfhw, hw2, maxp;
        solve_stencil(p, sfhw, hw2, maxp);
        par_a[3 * k] = sfhw; par_a[3 * k + 1] = hw2; par_a[3 * k + 2] = maxp;
        pok = pok && params_fast_domain(sfhw, hw2, maxp);
    }
    stamp();  // 5: stencils + first solve
    const bool x_ok = __syncthreads_and(ok);  // positions never change; parameters are re-checked every pass
    bool fast = __syncthreads_and(pok) && x_ok;
    for (int it = 0; it < e.n_iters; ++it) {
        const double *pin = (it & 1) ? par_b : par_a;
        double *pout = (it & 1) ? par_a : par_b;
        for (int w = t; w < 3 * P; w += SMALL_THREADS) {
            const int q = w / P, k = w - q * P;
            const double x = ox[q * cap + k];
            double acc = 0.0;
            if (fast) {
                int j = 0;
                if (P >= 8) {  // groups of 8: the ordered adds of one group run under the divisions of the next
                    double qa[8], qb[8];
                    lorentz_multi_q<8>(pin, x, qa);
                    for (j = 8; j + 8 <= P; j += 8) {
                        lorentz_multi_q<8>(pin + 3 * j, x, qb);
#pragma unroll
                        for (int u = 0; u < 8; ++u) { acc = __dadd_rn(acc, qa[u]); qa[u] = qb[u]; }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, qa[u]);
                }
                if (j + 4 <= P) { lorentz_multi<4>(pin + 3 * j, x, acc); j += 4; }
                if (j + 2 <= P) { lorentz_multi<2>(pin + 3 * j, x, acc); j += 2; }
                if (j < P) lorentz_multi<1>(pin + 3 * j, x, acc);
            } else {
                const double xs[1] = {x};
                double as[1] = {0.0};
#pragma unroll 1
                for (int j = 0; j < P; ++j) lorentz_step<1, false>(pin[3 * j], pin[3 * j + 1], pin[3 * j + 2], xs, as);
                acc = as[0];
            }
            // ratio = y_orig / superposition (:42-47), y_k = y_k * ratio_k (:52-54), here where 3P threads share the divisions
            sup[q * cap + k] = __dmul_rn(sy[q * cap + k], __ddiv_rn(oy[q * cap + k], acc));
        }
        __syncthreads();
        pok = true;
        for (int k = t; k < P; k += SMALL_THREADS) {
            Stencil p;
            p.x1 = sx1[k]; p.x2 = ox[cap + k]; p.x3 = sx3[k];
            p.y1 = sup[k]; p.y2 = sup[cap + k]; p.y3 = sup[2 * cap + k];
            mirror_shoulder(p);
            sx1[k] = p.x1; sx3[k] = p.x3;
            sy[k] = p.y1; sy[cap + k] = p.y2; sy[2 * cap + k] = p.y3;
            double sfhw, hw2, maxp;
            solve_stencil(p, sfhw, hw2, maxp);
            pout[3 * k] = sfhw; pout[3 * k + 1] = hw2; pout[3 * k + 2] = maxp;
            pok = pok && params_fast_domain(sfhw, hw2, maxp);
        }
        fast = __syncthreads_and(pok) && x_ok;
    }

    stamp();  // 6: refinement passes
    // ---- retain (fitter_analytical.rs:67-69): compact into the other buffer and into the result slot
    const double *__restrict__ fin = (e.n_iters & 1) ? par_b : par_a;
    double *__restrict__ kept = (e.n_iters & 1) ? par_a : par_b;
    double *__restrict__ resid = (e.n_iters & 1) ? par_b : par_a;  // overlays `fin` once the compaction is done
    const double CP = 1.0e+3 * 2.220446049250313e-16;  // lib.rs:277
    const int n_kept = small_ordered_compact(
        0, P, warp_cnt, [&](int k) { return fin[3 * k] > CP && fin[3 * k + 1] > CP; },
        [&](int k, int rank) {
            const double a = fin[3 * k], h = fin[3 * k + 1], m = fin[3 * k + 2];
            kept[3 * rank] = a; kept[3 * rank + 1] = h; kept[3 * rank + 2] = m;
            lor_out[3 * rank] = a; lor_out[3 * rank + 1] = h; lor_out[3 * rank + 2] = m;
        });
    o.n_kept = n_kept;

    // ---- MSE (deconvoluter.rs:828-862): superposition of the kept Lorentzians on every range point
    bool kok = true;
    for (int k = t; k < n_kept; k += SMALL_THREADS) kok = kok && params_fast_domain(kept[3 * k], kept[3 * k + 1], kept[3 * k + 2]);
    const bool kfast = __syncthreads_and(kok);
    auto mse_ranges = [&](auto ulp) {
        constexpr bool ULP = decltype(ulp)::value;
        int pos = 0;
        for (int q = 0; q < e.n_ranges; ++q) {
            const int s0 = e.ranges[2 * q], e0 = e.ranges[2 * q + 1];
            int i0 = s0;
            for (; e0 - i0 >= SMALL_THREADS * SMALL_R; i0 += SMALL_THREADS * SMALL_R)
                small_mse_block<SMALL_R, ULP>(d.x, d.y, kept, n_kept, kfast, i0, e0, resid + pos - s0);
            // the rest of the range with just enough points per thread (no thread evaluates padding)
            const int rest = (e0 - i0 + SMALL_THREADS - 1) / SMALL_THREADS;
            if (rest == 4) small_mse_block<4, ULP>(d.x, d.y, kept, n_kept, kfast, i0, e0, resid + pos - s0);
            else if (rest == 3) small_mse_block<3, ULP>(d.x, d.y, kept, n_kept, kfast, i0, e0, resid + pos - s0);
            else if (rest == 2) small_mse_block<2, ULP>(d.x, d.y, kept, n_kept, kfast, i0, e0, resid + pos - s0);
            else if (rest == 1) small_mse_block<1, ULP>(d.x, d.y, kept, n_kept, kfast, i0, e0, resid + pos - s0);
            pos += e0 - s0;
        }
    };
    if (fast_mse) mse_ranges(std::true_type{});
    else mse_ranges(std::false_type{});
    __syncthreads();
    stamp();  // 7: MSE superposition
    if (fast_mse) {
        // default mode: every range summed by the whole CTA through a fixed tree (as mse_reduce_fast_kernel),
        // the range sums folded in range order
        double residuals = 0.0;
        long long length = 0;
        int pos = 0;
        for (int q = 0; q < e.n_ranges; ++q) {
            const int len = e.ranges[2 * q + 1] - e.ranges[2 * q];
            const double *__restrict__ src = resid + pos;
            double v = 0.0;
            for (int i = t; i < len; i += SMALL_THREADS) v = __dadd_rn(v, src[i]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
            if ((t & 31) == 0) mse_part[t >> 5] = v;
            __syncthreads();
            if (t == 0) {
                double part = 0.0;
#pragma unroll
                for (int w = 0; w < SMALL_WARPS; ++w) part = __dadd_rn(part, mse_part[w]);
                residuals = __dadd_rn(residuals, part);
            }
            __syncthreads();
            length += len;
            pos += len;
        }
        if (t == 0) {
            o.mse = __ddiv_rn(residuals, (double)length);
            *hdr = o;
        }
    } else if (t == 0) {
        // each range is a left fold from 0.0, the range sums are folded in range order (:846-861)
        double residuals = 0.0;
        long long length = 0;
        int pos = 0;
        for (int q = 0; q < e.n_ranges; ++q) {
            const int len = e.ranges[2 * q + 1] - e.ranges[2 * q];
            const double *__restrict__ src = resid + pos;
            double part = 0.0;
            int i = 0;
            for (; i + 8 <= len; i += 8) {
                const double r0 = src[i], r1 = src[i + 1], r2 = src[i + 2], r3 = src[i + 3];
                const double r4 = src[i + 4], r5 = src[i + 5], r6 = src[i + 6], r7 = src[i + 7];
                part = __dadd_rn(part, r0); part = __dadd_rn(part, r1);
                part = __dadd_rn(part, r2); part = __dadd_rn(part, r3);
                part = __dadd_rn(part, r4); part = __dadd_rn(part, r5);
                part = __dadd_rn(part, r6); part = __dadd_rn(part, r7);
            }
            for (; i < len; ++i) part = __dadd_rn(part, src[i]);
            residuals = __dadd_rn(residuals, part);
            length += len;
            pos += len;
        }
        o.mse = __ddiv_rn(residuals, (double)length);
        *hdr = o;
    }
    stamp();  // 8: ordered fold, header written
}

}  // namespace mdb
