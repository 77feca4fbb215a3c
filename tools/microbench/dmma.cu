// FP64 tensor-core (mma.sync.m8n8k4.f64) characteristics on sm_100a: latency, throughput, and a syrk tile
// C(8x8) -= A^T A with operands loaded from shared memory.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_chains(double* out, long long* cyc, int iters) {
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++) dmma(c[i][0], c[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (s == 1.2345) out[0] = s;
}

// syrk tile: every warp owns 8x8 tiles of C = W^T W (W: 24 x 40 in shared memory), 6 DMMA per tile
__global__ void dmma_syrk(double* out, long long* cyc, int iters) {
    __shared__ double W[24 * 40];
    for (int i = threadIdx.x; i < 24 * 40; i += blockDim.x) W[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = warp % 5, tj = (warp / 5 + warp) % 5;
    double acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double c0 = 0, c1 = 0;
#pragma unroll
        for (int k0 = 0; k0 < 24; k0 += 4) {
            const double a = W[(k0 + (lane & 3)) * 40 + 8 * ti + (lane >> 2)];     // A[row=lane/4][col=lane%4] = W[k][i]
            const double b = W[(k0 + (lane & 3)) * 40 + 8 * tj + (lane >> 2)];     // B[row=lane%4][col=lane/4] = W[k][j]
            dmma(c0, c1, a, b);
        }
        acc += c0 + c1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 1.2345) out[0] = acc;
}

int main() {
    double* d; long long* c; long long h;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 2000;
#define RUN(K, NAME, PER) do { K; cudaDeviceSynchronize(); K; cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
        printf("%-52s %8.2f cycles per DMMA (256 FMA = 8 warp-DFMA)\n", NAME, (double)h / (PER)); } while (0)
    RUN((dmma_chains<1><<<1, 32>>>(d, c, iters)), "1 warp, dependent chain", 8.0 * iters);
    RUN((dmma_chains<2><<<1, 32>>>(d, c, iters)), "1 warp, 2 chains", 16.0 * iters);
    RUN((dmma_chains<4><<<1, 32>>>(d, c, iters)), "1 warp, 4 chains", 32.0 * iters);
    RUN((dmma_chains<4><<<1, 128>>>(d, c, iters)), "4 warps x 4 chains", 32.0 * iters);
    RUN((dmma_chains<4><<<1, 512>>>(d, c, iters)), "16 warps x 4 chains", 32.0 * iters);
    RUN((dmma_syrk<<<1, 32>>>(d, c, iters)), "syrk tile from smem (2 LDS.64 + DMMA), 1 warp", 6.0 * iters);
    RUN((dmma_syrk<<<1, 128>>>(d, c, iters)), "syrk tile from smem, 4 warps", 6.0 * iters);
    RUN((dmma_syrk<<<1, 512>>>(d, c, iters)), "syrk tile from smem, 16 warps", 6.0 * iters);
    return 0;
}
