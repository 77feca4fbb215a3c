// Single-warp FP64 characteristics on sm_100a (developer microbenchmark):
//   dependent DFMA chain latency, independent DFMA issue rate (1 warp, 1 CTA), LDS.64 -> DFMA patterns.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chains(double* out, long long* cyc, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++) v[i] = fma(v[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += v[i];
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
    if (s == 1.2345) out[0] = s;
}

__global__ void lds_dfma(double* out, long long* cyc, int iters) {
    __shared__ double sm[24 * 40];
    for (int i = threadIdx.x; i < 24 * 40; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int ti = threadIdx.x % 13, tj = (threadIdx.x / 13) % 13;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 4
        for (int l = 0; l < 24; l++) {
            const double* r = sm + l * 39;
            double u0 = r[3 * ti], u1 = r[3 * ti + 1], u2 = r[3 * ti + 2];
            double v0 = r[3 * tj], v1 = r[3 * tj + 1], v2 = r[3 * tj + 2];
            acc[0] += u0 * v0; acc[1] += u0 * v1; acc[2] += u0 * v2;
            acc[3] += u1 * v0; acc[4] += u1 * v1; acc[5] += u1 * v2;
            acc[6] += u2 * v0; acc[7] += u2 * v1; acc[8] += u2 * v2;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 9; i++) s += acc[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (s == 1.2345) out[0] = s;
}

__global__ void lds128_dfma(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[24 * 40];
    for (int i = threadIdx.x; i < 24 * 40; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int ti = threadIdx.x % 10, tj = (threadIdx.x / 10) % 10;
    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[p][q] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 2
        for (int l = 0; l < 24; l++) {
            const double2* r = reinterpret_cast<const double2*>(sm + l * 40);
            const double2 ua = r[2 * ti], ub = r[2 * ti + 1], va = r[2 * tj], vb = r[2 * tj + 1];
            const double u[4] = {ua.x, ua.y, ub.x, ub.y}, v[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] += u[p] * v[q];
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) s += acc[p][q];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (s == 1.2345) out[0] = s;
}

// broadcast LDS.128 + 2 DFMA (the substitution pattern): all lanes read the same address
__global__ void bcast_dfma(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[24 * 24];
    for (int i = threadIdx.x; i < 24 * 24; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double a[24];
#pragma unroll
    for (int i = 0; i < 24; i++) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 23; j++) {
            const double sj = a[j] * 1e-3;
#pragma unroll
            for (int i = j + 1; i < 24; i++) a[i] -= sm[j * 24 + i] * sj;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 24; i++) s += a[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (s == 1.2345) out[0] = s;
}

int main() {
    double* d; long long* c; long long h[8];
    cudaMalloc(&d, 8); cudaMalloc(&c, 64);
    const int iters = 1000;
#define RUN(K, NAME, WARPS, WORK) do { K; cudaDeviceSynchronize(); K; cudaDeviceSynchronize(); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); \
        printf("%-44s warps=%d  %8.2f cycles per %s\n", NAME, WARPS, (double)h[0] / (WORK), "warp-DFMA"); } while (0)
    RUN((dfma_chains<1><<<1, 32>>>(d, c, iters, 1.0000001, 1e-9)), "dependent chain (1 chain)", 1, 16.0 * iters);
    RUN((dfma_chains<2><<<1, 32>>>(d, c, iters, 1.0000001, 1e-9)), "2 independent chains", 1, 32.0 * iters);
    RUN((dfma_chains<4><<<1, 32>>>(d, c, iters, 1.0000001, 1e-9)), "4 independent chains", 1, 64.0 * iters);
    RUN((dfma_chains<8><<<1, 32>>>(d, c, iters, 1.0000001, 1e-9)), "8 independent chains", 1, 128.0 * iters);
    RUN((dfma_chains<16><<<1, 32>>>(d, c, iters, 1.0000001, 1e-9)), "16 independent chains", 1, 256.0 * iters);
    RUN((dfma_chains<8><<<1, 128>>>(d, c, iters, 1.0000001, 1e-9)), "8 chains, 4 warps (1 per sub-partition)", 4, 128.0 * iters);
    RUN((dfma_chains<8><<<1, 256>>>(d, c, iters, 1.0000001, 1e-9)), "8 chains, 8 warps (2 per sub-partition)", 8, 128.0 * iters);
    RUN((dfma_chains<8><<<1, 512>>>(d, c, iters, 1.0000001, 1e-9)), "8 chains, 16 warps (4 per sub-partition)", 16, 128.0 * iters);
    RUN((lds_dfma<<<1, 32>>>(d, c, iters)), "syrk tile loop (6 LDS.64 + 9 DFMA), 1 warp", 1, 216.0 * iters);
    RUN((lds_dfma<<<1, 128>>>(d, c, iters)), "syrk tile loop, 4 warps", 4, 216.0 * iters);
    RUN((lds_dfma<<<1, 512>>>(d, c, iters)), "syrk tile loop, 16 warps", 16, 216.0 * iters);
    RUN((lds128_dfma<<<1, 32>>>(d, c, iters)), "syrk 4x4 tile loop (4 LDS.128 + 16 DFMA), 1 warp", 1, 384.0 * iters);
    RUN((lds128_dfma<<<1, 128>>>(d, c, iters)), "syrk 4x4 tile loop, 4 warps", 4, 384.0 * iters);
    RUN((lds128_dfma<<<1, 512>>>(d, c, iters)), "syrk 4x4 tile loop, 16 warps", 16, 384.0 * iters);
    RUN((bcast_dfma<<<1, 32>>>(d, c, iters)), "substitution (bcast LDS.128 + 2 DFMA), 1 warp", 1, 276.0 * iters);
    RUN((bcast_dfma<<<1, 128>>>(d, c, iters)), "substitution, 4 warps", 4, 276.0 * iters);
    RUN((bcast_dfma<<<1, 512>>>(d, c, iters)), "substitution, 16 warps", 16, 276.0 * iters);
    return 0;
}
