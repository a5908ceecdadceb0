// Roofline denominator that MEASURED_PEAKS.json does not carry: the fp64 vector
// (DFMA) peak of the device, measured live (SURVEY.md §8d asks the builder to
// measure it with an FMA micro-benchmark on the box).
#include <cuda_runtime.h>

#include "../../include/tfhe_b200.h"

namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

}  // namespace

// burst: best single launch (about 3 ms); sustained: mean over ~0.4 s of back-to-back launches
// (what a kernel timed inside a long step can hope for under the power cap).
extern "C" int tfhe_b200_measure_fp64_peak(int device, double *tflops_out, double *sustained_out) {
    if (!tflops_out) return 1;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
    double *d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return 1;
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (sustained_out) {
        const int reps = (int) (400.0f / best) + 1;
        cudaEventRecord(e0);
        for (int rep = 0; rep < reps; rep++) dfma_peak_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *sustained_out = flops * reps / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}
