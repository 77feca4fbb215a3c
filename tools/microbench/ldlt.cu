// Warp-level LDL^T of a 24 x 24 SPD matrix (lane = column) with the inverse factor built in the freed lanes:
// variants of the step structure, timed in isolation (1 warp per SM) and with four such warps on one SM
// sub-partition (4 CTAs per SM, as in the solver).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   V0: columns published through shared memory (128-bit), lane j skips its own step (solver build)
//   V1: same with 64-bit shared accesses
//   V2: columns broadcast by warp shuffles (no shared-memory column traffic, no divergent publish)
//   V3: V0 without the E trick (lanes <= j update dead values; reference for the cost of the chain alone)
//   V4: V0 with the pivot chain through shuffles (1/pivot and col_j[j+1]) and the column through shared memory
//   V10: by symmetry col_j[t] is lane t's own a[j]: every lane stores one entry, nobody publishes a column
//   V20: V15 (rolled, R = 2) with the two pivot steps of a trip as ONE 2 x 2 block step: lanes j and j+1 publish their
//        columns together (j+1 raw), every lane forms the Schur pivot p1 = p_{j+1} - c^2 / p_j itself and applies both
//        updates at once: one publish -> sync -> load round trip per two steps.  V20: a[i] -= col_j[i] (s1 - l s2) + col_{j+1}[i] s2
//        (two FMAs per entry; NOT what the solver uses: the two large terms cancel in the accumulator);
//   V21: a[i] -= col_j[i] s1; a[i] -= (col_{j+1}[i] - col_j[i] l) s2 with l from lane j+1's own copy of the symmetric entry: the
//        elimination of V15 bit for bit, three FMAs per entry (the solver's fp32 form, sddp_backward_srbd.cuh SDDP_D1BLOCK)
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
constexpr int NU = 24;

__device__ __forceinline__ double fast_rcp(double p) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(p));
    double e = fma(-p, x, 1.0);
    x = fma(x, e, x);
    e = fma(-p, x, 1.0);
    return fma(x, e, x);
}
template <bool W128>
__device__ __forceinline__ void axpy_row(double* a, const double* row, int lo, double s) {
    if (W128) {
#pragma unroll
        for (int i0 = 0; i0 < NU; i0 += 8) {
            double2 c[4];
#pragma unroll
            for (int q = 0; q < 4; q++) if (i0 + 2 * q + 1 >= lo) c[q] = *reinterpret_cast<const double2*>(row + i0 + 2 * q);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = i0 + 2 * q;
                if (i >= lo) a[i] -= c[q].x * s;
                if (i + 1 >= lo) a[i + 1] -= c[q].y * s;
            }
        }
    } else {
#pragma unroll
        for (int i0 = lo; i0 < NU; i0 += 8) {
            double col[8];
#pragma unroll
            for (int q = 0; q < 8; q++) if (i0 + q < NU) col[q] = row[i0 + q];
#pragma unroll
            for (int q = 0; q < 8; q++) if (i0 + q < NU) a[i0 + q] -= col[q] * s;
        }
    }
}
template <bool W128>
__device__ __forceinline__ void store_row(double* row, const double* a, int lo) {
    if (W128) {
#pragma unroll
        for (int i = 0; i < NU; i += 2) {
            if (i >= lo) *reinterpret_cast<double2*>(row + i) = make_double2(a[i], a[i + 1]);
            else if (i + 1 >= lo) row[i + 1] = a[i + 1];
        }
    } else {
#pragma unroll
        for (int i = 0; i < NU; i++) if (i >= lo) row[i] = a[i];
    }
}

struct alignas(16) Sm { double Q0[NU * NU]; double Quu[NU * NU]; double invp[NU]; double rs[NU]; double prow[4][NU]; };
__device__ int g_zero;
__device__ __forceinline__ double tie(double s, double v, int zero) { return __hiloint2double(__double2hiint(s) + (__double2hiint(v) & zero), __double2loint(s)); }
template <int I0, int I1>
__device__ __forceinline__ void axpy_part(double* a, const double* row, int lo, double s) {
#pragma unroll
    for (int i0 = I0; i0 < I1; i0 += 8) {
        double2 c[4];
#pragma unroll
        for (int q = 0; q < 4; q++) if (i0 + 2 * q + 1 >= lo) c[q] = *reinterpret_cast<const double2*>(row + i0 + 2 * q);
#pragma unroll
        for (int q = 0; q < 4; q++) { const int i = i0 + 2 * q; if (i >= lo) a[i] -= c[q].x * s; if (i + 1 >= lo) a[i + 1] -= c[q].y * s; }
    }
}

__device__ long long g_split[3];
template <int V>
__device__ __forceinline__ void ldlt(Sm& S, int lane) {
    const long long c0 = clock64();
    double a[NU];
    const int t = lane < NU ? lane : NU - 1;
#pragma unroll
    for (int i = 0; i < NU; i++) a[i] = S.Q0[i * NU + t];
    __syncwarp();
    double myinv = fast_rcp(a[0]), pinv = 0.0;
    const long long c1 = clock64();
    if (V == 0 || V == 1 || V == 3 || V == 9) {
        constexpr bool W = (V != 1);
#pragma unroll
        for (int j = 0; j < NU; j++) {
            if (lane == j) { pinv = myinv; S.invp[j] = myinv; store_row<W>(S.Quu + j * NU, a, j + 1); }
            __syncwarp();
            if (j + 1 < NU) {
                const double sj = (V != 3 && lane == j) ? 0.0 : S.invp[j] * a[j];
                a[j + 1] -= S.Quu[j * NU + j + 1] * sj;
                myinv = fast_rcp(a[j + 1]);
                if (j + 2 < NU) axpy_row<W>(a, S.Quu + j * NU, j + 2, sj);
            }
        }
    } else if (V == 10) {      // every lane publishes its own entry of row j (= col_j[lane] by symmetry): no divergent publish
#pragma unroll
        for (int j = 0; j < NU; j++) {
            if (lane < NU) S.Quu[j * NU + lane] = a[j];
            if (lane == j) S.invp[j] = myinv;
            pinv = (lane == j) ? myinv : pinv;
            __syncwarp();
            if (j + 1 < NU) {
                const double sj = (lane == j) ? 0.0 : S.invp[j] * a[j];
                a[j + 1] -= S.Quu[j * NU + j + 1] * sj;
                myinv = fast_rcp(a[j + 1]);
                if (j + 2 < NU) axpy_row<true>(a, S.Quu + j * NU, j + 2, sj);
            }
        }
    } else if (V >= 5 && V <= 8) {      // timing probes of V0 (wrong results): 5 no reciprocal, 6 no warp sync, 7 chain only, 8 no chain
#pragma unroll
        for (int j = 0; j < NU; j++) {
            if (V != 8) { if (lane == j) { pinv = myinv; S.invp[j] = myinv; if (V != 7) store_row<true>(S.Quu + j * NU, a, j + 1); else S.Quu[j * NU + j + 1] = a[j + 1]; } }
            if (V != 6) __syncwarp();
            if (j + 1 < NU) {
                const double sj = (V == 8) ? 1e-3 * a[j] : ((lane == j) ? 0.0 : S.invp[j] * a[j]);
                if (V != 8) { a[j + 1] -= S.Quu[j * NU + j + 1] * sj; myinv = (V == 5) ? a[j + 1] * 1e-3 : fast_rcp(a[j + 1]); }
                if (V != 7 && j + 2 < NU) axpy_row<true>(a, (V == 8 ? S.Q0 : S.Quu) + j * NU, j + 2, sj);
            }
        }
    } else if (V >= 11 && V <= 17) {
        // rolled over the pivot steps, rotating register frame, Et rows stored in the loop.
        // 11: plain  12: first 16 tied to the seed, last 8 to the result  13: three ties  14: R = 8 plain  15: R = 2 plain  16: R = 4, unrolled loop (no roll)
        constexpr int R = (V == 14) ? 8 : (V == 15 ? 2 : 4);
        const int zero = g_zero;
        double npinv = 0.0;
#pragma unroll 1
        for (int jb = 0; jb < NU; jb += R) {
#pragma unroll
            for (int s_ = 0; s_ < R; s_++) {
                const int j = jb + s_;
                double* pr = S.prow[s_ & 1];
                const double m = a[s_];
                if (V == 16 || V == 17) {      // symmetric publish: col_j[lane] = this lane's own a[j]
                    if (lane >= jb && lane < NU) pr[lane - jb] = m;
                    if (lane == j) { npinv = -myinv; pinv = myinv; S.invp[j] = myinv; }
                } else if (lane == j) {
                    npinv = -myinv; pinv = myinv;
                    S.invp[j] = myinv;
                    store_row<true>(pr, a, s_ + 1);
                }
                __syncwarp();
                const double sj = (lane == j) ? 0.0 : S.invp[j] * m;
                a[s_ + 1] -= pr[s_ + 1] * sj;
                if (V == 11 || V >= 14) { myinv = fast_rcp(a[s_ + 1]); axpy_part<0, NU>(a, pr, s_ + 2, sj); if (V == 17) __syncwarp(); }
                else {
                    const double pn = a[s_ + 1];
                    double x0;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(pn));
                    if (V == 12) {
                        axpy_part<0, 16>(a, pr, s_ + 2, tie(sj, x0, zero));
                        double e_ = fma(-pn, x0, 1.0); x0 = fma(x0, e_, x0); e_ = fma(-pn, x0, 1.0); myinv = fma(x0, e_, x0);
                        axpy_part<16, NU>(a, pr, s_ + 2, tie(sj, myinv, zero));
                    } else {
                        axpy_part<0, 8>(a, pr, s_ + 2, tie(sj, x0, zero));
                        double e_ = fma(-pn, x0, 1.0); x0 = fma(x0, e_, x0);
                        axpy_part<8, 16>(a, pr, s_ + 2, tie(sj, x0, zero));
                        e_ = fma(-pn, x0, 1.0); myinv = fma(x0, e_, x0);
                        axpy_part<16, NU>(a, pr, s_ + 2, tie(sj, myinv, zero));
                    }
                }
                const double ev = (lane < j) ? npinv * m : (lane == j ? 1.0 : 0.0);
                if (lane < NU) S.Quu[j * NU + lane] = ev;
            }
#pragma unroll
            for (int q = 0; q < NU; q++) a[q] = (q + R < NU) ? a[q + R] : 0.0;
        }
    } else if (V == 21) {
        double npinv = 0.0;
        int par = 0;
#pragma unroll 1
        for (int jb = 0; jb < NU; jb += 2, par ^= 2) {
            double* pr0 = S.prow[par];
            double* pr1 = S.prow[par + 1];
            const double m0 = a[0];
            if (lane == jb || lane == jb + 1) {
                if (lane == jb) { npinv = -myinv; pinv = myinv; S.invp[jb] = myinv; }
                store_row<true>(S.prow[par + (lane - jb)], a, 0);
            }
            __syncwarp();
            const double p0inv = S.invp[jb], c = pr0[1], p1raw = pr1[1];
            const double l = pr1[0] * p0inv;
            const double p1 = fma(-c, l, p1raw);
            const double p1inv = fast_rcp(p1);
            const double s1 = (lane == jb) ? 0.0 : m0 * p0inv;
            const double m1 = fma(-c, s1, a[1]);
            if (lane == jb + 1) { npinv = -p1inv; pinv = p1inv; }
            const double s2 = (lane == jb + 1) ? 0.0 : m1 * p1inv;
            a[2] -= pr0[2] * s1; a[2] -= fma(-pr0[2], l, pr1[2]) * s2;
            myinv = fast_rcp(a[2]);
#pragma unroll
            for (int i0 = 0; i0 < NU; i0 += 8) {
                double2 c0[4], c1[4];
#pragma unroll
                for (int q = 0; q < 4; q++) if (i0 + 2 * q + 1 >= 3) { c0[q] = *reinterpret_cast<const double2*>(pr0 + i0 + 2 * q); c1[q] = *reinterpret_cast<const double2*>(pr1 + i0 + 2 * q); }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = i0 + 2 * q;
                    if (i >= 3) { a[i] -= c0[q].x * s1; a[i] -= fma(-c0[q].x, l, c1[q].x) * s2; }
                    if (i + 1 >= 3) { a[i + 1] -= c0[q].y * s1; a[i + 1] -= fma(-c0[q].y, l, c1[q].y) * s2; }
                }
            }
            const double e0 = (lane < jb) ? npinv * m0 : (lane == jb ? 1.0 : 0.0);
            const double e1 = (lane < jb + 1) ? npinv * m1 : (lane == jb + 1 ? 1.0 : 0.0);
            if (lane < NU) { S.Quu[jb * NU + lane] = e0; S.Quu[(jb + 1) * NU + lane] = e1; }
#pragma unroll
            for (int q = 0; q < NU; q++) a[q] = (q + 2 < NU) ? a[q + 2] : 0.0;
        }
    } else if (V == 20) {
        double npinv = 0.0;
        int par = 0;
#pragma unroll 1
        for (int jb = 0; jb < NU; jb += 2, par ^= 2) {
            double* pr0 = S.prow[par];
            double* pr1 = S.prow[par + 1];
            const double m0 = a[0];
            if (lane == jb || lane == jb + 1) {
                if (lane == jb) { npinv = -myinv; pinv = myinv; S.invp[jb] = myinv; }
                store_row<true>(S.prow[par + (lane - jb)], a, 1);
            }
            __syncwarp();
            const double p0inv = S.invp[jb], c = pr0[1], p1raw = pr1[1];
            const double l = c * p0inv;
            const double p1 = fma(-c, l, p1raw);
            const double p1inv = fast_rcp(p1);
            const double s1 = (lane == jb) ? 0.0 : m0 * p0inv;
            const double m1 = fma(-c, s1, a[1]);
            if (lane == jb + 1) { npinv = -p1inv; pinv = p1inv; }
            const double s2 = (lane == jb + 1) ? 0.0 : m1 * p1inv;
            const double g1 = fma(-l, s2, s1);
            // the next pivot first, its reciprocal started at once
            a[2] -= pr0[2] * g1; a[2] -= pr1[2] * s2;
            myinv = fast_rcp(a[2]);
            axpy_part<0, NU>(a, pr0, 3, g1);
            axpy_part<0, NU>(a, pr1, 3, s2);
            const double e0 = (lane < jb) ? npinv * m0 : (lane == jb ? 1.0 : 0.0);
            const double e1 = (lane < jb + 1) ? npinv * m1 : (lane == jb + 1 ? 1.0 : 0.0);
            if (lane < NU) { S.Quu[jb * NU + lane] = e0; S.Quu[(jb + 1) * NU + lane] = e1; }
#pragma unroll
            for (int q = 0; q < NU; q++) a[q] = (q + 2 < NU) ? a[q + 2] : 0.0;
        }
    } else if (V == 2) {
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const double ij = __shfl_sync(FULL, myinv, j);
            if (lane == j) pinv = myinv;
            if (j + 1 < NU) {
                const double sj = (lane == j) ? 0.0 : ij * a[j];
                a[j + 1] -= __shfl_sync(FULL, a[j + 1], j) * sj;
                myinv = fast_rcp(a[j + 1]);
#pragma unroll
                for (int i = j + 2; i < NU; i++) a[i] -= __shfl_sync(FULL, a[i], j) * sj;
            }
        }
    } else if (V == 4) {
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const double ij = __shfl_sync(FULL, myinv, j);
            const double cj1 = (j + 1 < NU) ? __shfl_sync(FULL, a[j + 1], j) : 0.0;
            if (lane == j) { pinv = myinv; if (j + 2 < NU) store_row<true>(S.Quu + j * NU, a, j + 2); }
            const double sj = (lane == j) ? 0.0 : ij * a[j];
            if (j + 1 < NU) { a[j + 1] -= cj1 * sj; myinv = fast_rcp(a[j + 1]); }
            __syncwarp();
            if (j + 2 < NU) axpy_row<true>(a, S.Quu + j * NU, j + 2, sj);
        }
    }
    const long long c2 = clock64();
    if (lane < NU) S.rs[lane] = sqrt(pinv);
    __syncwarp();
    if (V >= 11) {         // rows of Et are in place; scale them here only so that the result can be compared
        if (lane < NU) for (int i = 0; i < NU; i++) S.Quu[i * NU + lane] *= S.rs[i];
    } else if (V == 9) {          // the divergent-branch epilogue the solver had first (kept as a warning)
        if (lane < NU) {
#pragma unroll
            for (int i = 0; i < NU; i++) {
                if (i > lane) S.Quu[i * NU + lane] = -pinv * a[i] * S.rs[i];
                else if (i == lane) S.Quu[i * NU + lane] = S.rs[i];
            }
        }
    } else {
        const double np = -pinv;
#pragma unroll
        for (int i = 0; i < NU; i += 2) {
            const double2 r = *reinterpret_cast<const double2*>(S.rs + i);
            const double v0 = (i > t) ? np * a[i] * r.x : (i == t ? r.x : 0.0);
            const double v1 = (i + 1 > t) ? np * a[i + 1] * r.y : (i + 1 == t ? r.y : 0.0);
            if (lane < NU) { S.Quu[i * NU + lane] = v0; S.Quu[(i + 1) * NU + lane] = v1; }
        }
    }
    __syncwarp();
    const long long c3 = clock64();
    if (lane == 0 && blockIdx.x == 0) { g_split[0] += c1 - c0; g_split[1] += c2 - c1; g_split[2] += c3 - c2; }
}

template <int V>
__global__ void __launch_bounds__(128, 4) bench(const double* Q, double* out, long long* cyc, int iters) {
    __shared__ Sm S;
    for (int i = threadIdx.x; i < NU * NU; i += blockDim.x) S.Q0[i] = Q[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) ldlt<V>(S, threadIdx.x);
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < NU * NU; i += blockDim.x) out[i] = S.Quu[i];
}

template <int V>
static void run(const char* name, const double* dQ, double* dout, long long* dcyc, const double* ref) {
    const int iters = 200;
    for (int ctas = 1; ctas <= 4; ctas *= 4) {
        const int grid = 148 * ctas;
        bench<V><<<grid, 128>>>(dQ, dout, dcyc, iters);
        bench<V><<<grid, 128>>>(dQ, dout, dcyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        static long long h[148 * 4];
        cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < grid; i++) s += h[i];
        static double o[NU * NU];
        cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < NU; i++) for (int j = 0; j <= i; j++) { double d = fabs(o[i * NU + j] - ref[i * NU + j]); if (d > err) err = d; }
        long long sp[3], z[3] = {0, 0, 0};
        cudaMemcpyFromSymbol(sp, g_split, sizeof(sp)); cudaMemcpyToSymbol(g_split, z, sizeof(z));
        printf("%-44s %d CTA/SM: %8.0f cycles per factorisation (load %lld, steps %lld, Es %lld)  max |Es - ref| = %.2e\n", name, ctas, s / grid / iters,
               sp[0] / (2 * iters), sp[1] / (2 * iters), sp[2] / (2 * iters), err);
    }
}

int main() {
    static double Q[NU * NU], L[NU * NU], ref[NU * NU];
    // SPD test matrix with the solver's scale spread: A = G G^T + diag
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
    static double G[NU * NU];
    for (int i = 0; i < NU * NU; i++) G[i] = rnd();
    for (int i = 0; i < NU; i++) for (int j = 0; j < NU; j++) {
        double v = 0;
        for (int k = 0; k < NU; k++) v += G[i * NU + k] * G[j * NU + k];
        Q[i * NU + j] = v * ((i % 6 < 3) ? 1.0 : 30.0) * ((j % 6 < 3) ? 1.0 : 30.0) + (i == j ? (i % 6 < 3 ? 2.0 : 2e8) : 0.0);
    }
    // reference Es = L^-1 of the Cholesky factor (Es = D^-1/2 Lt^-1)
    for (int i = 0; i < NU * NU; i++) L[i] = 0;
    for (int j = 0; j < NU; j++) {
        double d = Q[j * NU + j];
        for (int k = 0; k < j; k++) d -= L[j * NU + k] * L[j * NU + k];
        L[j * NU + j] = sqrt(d);
        for (int i = j + 1; i < NU; i++) {
            double v = Q[i * NU + j];
            for (int k = 0; k < j; k++) v -= L[i * NU + k] * L[j * NU + k];
            L[i * NU + j] = v / L[j * NU + j];
        }
    }
    for (int c = 0; c < NU; c++) {      // forward substitution on identity columns
        for (int i = 0; i < NU; i++) {
            double v = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) v -= L[i * NU + k] * ref[k * NU + c];
            ref[i * NU + c] = v / L[i * NU + i];
        }
    }
    double *dQ, *dout; long long* dcyc;
    cudaMalloc(&dQ, sizeof(Q)); cudaMalloc(&dout, sizeof(Q)); cudaMalloc(&dcyc, sizeof(long long) * 148 * 4);
    cudaMemcpy(dQ, Q, sizeof(Q), cudaMemcpyHostToDevice);
    run<0>("V0 smem 128-bit, E in freed lanes", dQ, dout, dcyc, ref);
    run<1>("V1 smem 64-bit, E in freed lanes", dQ, dout, dcyc, ref);
    run<2>("V2 shuffle broadcast, E in freed lanes", dQ, dout, dcyc, ref);
    run<3>("V3 smem 128-bit, no E (wrong Es, timing only)", dQ, dout, dcyc, ref);
    run<4>("V4 shuffle pivot chain + smem column", dQ, dout, dcyc, ref);
    run<5>("V5 = V0 without the reciprocal (timing)", dQ, dout, dcyc, ref);
    run<6>("V6 = V0 without __syncwarp (timing)", dQ, dout, dcyc, ref);
    run<7>("V7 = V0 pivot chain only (timing)", dQ, dout, dcyc, ref);
    run<8>("V8 = V0 column updates only (timing)", dQ, dout, dcyc, ref);
    run<9>("V9 = V0 with a branch per entry in the Es write", dQ, dout, dcyc, ref);
    run<10>("V10 every lane publishes its row-j entry", dQ, dout, dcyc, ref);
    run<11>("V11 rolled R=4, rotating frame, Et rows in loop", dQ, dout, dcyc, ref);
    run<12>("V12 = V11, update tied behind seed / result", dQ, dout, dcyc, ref);
    run<13>("V13 = V11, update in thirds tied to the rcp stages", dQ, dout, dcyc, ref);
    run<14>("V14 = V11 with R=8", dQ, dout, dcyc, ref);
    run<15>("V15 = V11 with R=2", dQ, dout, dcyc, ref);
    run<16>("V16 = V11 with the symmetric publish of V10", dQ, dout, dcyc, ref);
    run<17>("V17 = V16 + a second warp sync per step", dQ, dout, dcyc, ref);
    run<20>("V20 = V15 with one 2 x 2 block step per trip", dQ, dout, dcyc, ref);
    run<21>("V21 = V20, bit-identical elimination (3 FMAs)", dQ, dout, dcyc, ref);
    return 0;
}
