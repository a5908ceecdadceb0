// Development aid: SM-level micro-benchmarks that decide the blind-rotation schedule
// (fp64 latency / issue cadence, co-issue with integer and shared-memory work, double->int64
// conversion rate, fp64 MMA vs DFMA).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o build/ubench tools/ubench.cu ; run on the GPU box.  One CTA, cycles from clock64().
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// mode 0: 1 dependent DFMA chain; 1: 8 chains; 2: 8 chains + 8 IADD/LOP per 8 DFMA; 3: 8 chains + 16 int;
// 4: 8 DFMA + 2 LDS.128; 5: F2I.S64.F64 x8 ; 6: DMMA x4 chains; 7: DMMA x4 + 8 DFMA; 8: int only (16);
// 9: 8 DFMA + 4 LDS.128 ; 10: 16 chains DFMA; 11: 8 DFMA + 8 F2I
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *out, long long *cyc, double a, double b, int ia) {
    __shared__ double4 sm[1024];
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    int n[16];
#pragma unroll
    for (int i = 0; i < 16; i++) n[i] = threadIdx.x * 3 + i;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double4(i, 1, 2, 3);
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0, d7 = 0;
    long long acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 8; u++) x[0] = fma(x[0], a, b);
        }
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 4 || MODE == 7 || MODE == 9 || MODE == 11) {
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = fma(x[u], a, b);
        }
        if (MODE == 10) {
#pragma unroll
            for (int u = 0; u < 16; u++) x[u] = fma(x[u], a, b);
        }
        if (MODE == 2 || MODE == 3 || MODE == 8) {
#pragma unroll
            for (int u = 0; u < (MODE == 2 ? 8 : 16); u++) n[u] = ((n[u] ^ ia) + ia) ^ (n[u] >> 3);
        }
        if (MODE == 4 || MODE == 9) {
#pragma unroll
            for (int u = 0; u < (MODE == 4 ? 2 : 4); u++) {
                const double4 v = sm[(threadIdx.x + 32 * u + it) & 1023];
                x[8 + u] += v.x;  // 1 extra fp64 per load
                acc += __double_as_longlong(v.z);
            }
        }
        if (MODE == 5 || MODE == 11) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                x[8 + u] += a;
                acc += (long long) x[8 + u];
            }
        }
        if (MODE == 6 || MODE == 7) {
            dmma(d0, d1, a, b);
            dmma(d2, d3, a, b);
            dmma(d4, d5, a, b);
            dmma(d6, d7, a, b);
        }
    }
    const long long t1 = clock64();
    double s = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i] + n[i];
    if (s == 123.456 || acc == 77) out[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
static void run(const char *name, double *d, long long *c) {
    const int warps_per_smsp[] = {1, 2, 3, 4};
    printf("%-44s", name);
    for (int w : warps_per_smsp) {
        k<MODE><<<1, 128 * w>>>(d, c, 1.0000001, 1e-9, 12345);
        k<MODE><<<1, 128 * w>>>(d, c, 1.0000001, 1e-9, 12345);
        long long h = 0;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("  %dw/smsp: %8.2f cyc/iter", w, (double) h / ITERS);
    }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    double *d;
    long long *c;
    cudaMalloc(&d, 64);
    cudaMalloc(&c, 64);
    run<0>("8 dependent DFMA (1 chain)", d, c);
    run<1>("8 DFMA (8 chains)", d, c);
    run<10>("16 DFMA (16 chains)", d, c);
    run<8>("16 x (LOP,IADD,SHF,LOP) int only", d, c);
    run<2>("8 DFMA + 8 x 4 int", d, c);
    run<3>("8 DFMA + 16 x 4 int", d, c);
    run<4>("8 DFMA + 2 LDS.128 (+2 DADD)", d, c);
    run<9>("8 DFMA + 4 LDS.128 (+4 DADD)", d, c);
    run<5>("8 x (DADD + F2I.S64.F64 + IADD64)", d, c);
    run<11>("8 DFMA + 8 x (DADD + F2I.S64.F64)", d, c);
    run<6>("4 DMMA m8n8k4 (4 chains)", d, c);
    run<7>("8 DFMA + 4 DMMA", d, c);
    return 0;
}
