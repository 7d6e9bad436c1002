// fp64_warp_bench.cu -- what ONE warp can do with the FP64 pipe on B200 (development aid).
// The smoothing recurrence is two dependent DADDs per point plus one DMUL that nothing on the
// chain depends on; this measures whether that DMUL (or any independent FP64 work of the same
// warp) is free, i.e. whether a single warp gets instruction-level parallelism on the FP64 pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_warp_bench fp64_warp_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int K>
__global__ void chains_kernel(double *out, long long *cyc, int iters, double seed)
{
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = seed + k;
    const double b = 1.0 + seed * 1e-9;
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] = __dadd_rn(a[k], b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += a[k];
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = s; }
}

// MODE 0: sum += a; sum -= q            (the chain alone)
// MODE 1: ... and o = sum * div         (DMUL hanging off the chain, same warp)
// MODE 2: as 1, values through shared memory like the kernel (LDS a, LDS q, STS o)
template <int MODE>
__global__ void smooth_like_kernel(double *out, long long *cyc, int iters, double seed, int active_lanes)
{
    __shared__ double in[32][80], res[32][80];
    const int lane = threadIdx.x;
    for (int i = 0; i < 80; ++i) { in[lane][i] = seed + i * 1e-3 + lane; res[lane][i] = 0.0; }
    __syncwarp();
    double sum = seed, div = 1.0 / 3.0, acc = 0.0;
    const long long t0 = clock64();
    if (lane < active_lanes) {
        for (int i = 0; i < iters; ++i) {
            double a[8], q[8], o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (MODE == 2) { a[u] = in[lane][(i * 8 + u) % 64 + 3]; q[u] = in[lane][(i * 8 + u) % 64]; }
                else { a[u] = seed + u; q[u] = seed - u; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                sum = __dadd_rn(sum, a[u]);
                sum = __dsub_rn(sum, q[u]);
                if (MODE >= 1) o[u] = __dmul_rn(sum, div);
            }
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = (o[u] > acc) ? o[u] : acc;   // keep the products alive, off the FP64 pipe
            }
            if (MODE == 2) {
#pragma unroll
                for (int u = 0; u < 8; ++u) res[lane][(i * 8 + u) % 64] = o[u];
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = sum + acc + res[0][5]; }
}

int main()
{
    double *d; long long *c, h;
    CK(cudaMalloc(&d, 8)); CK(cudaMalloc(&c, 8));
    const int iters = 4096;
#define RUN_CH(K) chains_kernel<K><<<1, 32>>>(d, c, iters, 1.0000001); CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost)); \
    printf("one warp, %d independent DADD chains: %.2f cycles per DADD, %.2f per step\n", K, (double)h / iters / K, (double)h / iters);
    RUN_CH(1) RUN_CH(2) RUN_CH(3) RUN_CH(4) RUN_CH(8)
    const char *names[] = {"chain only (2 DADD/pt)", "chain + DMUL", "chain + DMUL + LDS/STS"};
#define RUN_SM(M, L) smooth_like_kernel<M><<<1, 32>>>(d, c, 512, 1.0000001, L); CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost)); \
    printf("%-26s %2d lanes: %.2f cycles per point\n", names[M], L, (double)h / 512 / 8);
    RUN_SM(0, 1) RUN_SM(0, 32) RUN_SM(1, 1) RUN_SM(1, 3) RUN_SM(1, 32) RUN_SM(2, 1) RUN_SM(2, 3) RUN_SM(2, 32)
    return 0;
}
