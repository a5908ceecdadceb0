// Development aid: does a key chunk read from TENSOR MEMORY (tcgen05.cp smem -> TMEM once per CTA, tcgen05.ld by
// every warp) beat reading it from shared memory (LDS.128 by every warp)?  Checks the copy / load semantics
// (32x128b.warpx4: 32 rows of 16 B broadcast to the four sub-partitions) and times both read paths under a
// DFMA load like the Fourier multiply's.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_probe tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

struct cpx { double x, y; };

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((addr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

constexpr int kChunkCplx = 16 * 32;   // [pos][m1]
constexpr int kChunks = 8;

struct __align__(128) Smem {
    cpx chunk[kChunks][kChunkCplx];   // 64 KiB
    unsigned long long bar;
    uint32_t tmem_base;
};

// mode 0: LDS path, mode 1: TMEM path.  Each of 8 warps multiplies z[16] by all 8 chunks, `iters` times.
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(const cpx *key, cpx *out, long long *cyc, int iters, int *bad) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem &S = *reinterpret_cast<Smem *>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kChunks * kChunkCplx; i += blockDim.x) (&S.chunk[0][0])[i] = key[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S.tmem_base)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    if (threadIdx.x == 0) {
        // generic-proxy writes of the chunks -> visible to the async proxy (tcgen05.cp reads shared memory)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const long long c0 = clock64();
        for (int c = 0; c < kChunks; c++)
            for (int pos = 0; pos < 16; pos++) {
                // 32 rows (m1) of 16 B, contiguous: core matrices of 8 rows = 128 B, SBO = 128 B
                const uint64_t d = smem_desc(smem_addr(&S.chunk[c][pos * 32]), 0, 128);
                const uint32_t dst = tmem + (uint32_t) (c * 64 + pos * 4);
                asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(dst), "l"(d) : "memory");
            }
        const long long c1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&S.bar)) : "memory");
        asm volatile(
            "{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_addr(&S.bar)) : "memory");
        const long long c2 = clock64();
        if (blockIdx.x == 0 && MODE == 1) printf("128 tcgen05.cp: issue %lld cycles (%.1f each), until the commit has arrived %lld more\n", c1 - c0, (double) (c1 - c0) / 128.0, c2 - c1);
    }
    {   // wait for the copies
        asm volatile(
            "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_addr(&S.bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    // ---- semantics check: TMEM lane (warp%4)*32 + lane, columns c*64 + pos*4 .. +3 == chunk[c][pos*32 + lane]
    if (MODE == 1) {
        for (int c = 0; c < kChunks; c++)
            for (int g = 0; g < 4; g++) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + (uint32_t) (c * 64 + g * 16), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int p = 0; p < 4; p++) {
                    const cpx want = S.chunk[c][(g * 4 + p) * 32 + lane];
                    const double gx = __hiloint2double((int) v[4 * p + 1], (int) v[4 * p]);
                    const double gy = __hiloint2double((int) v[4 * p + 3], (int) v[4 * p + 2]);
                    if (gx != want.x || gy != want.y) atomicAdd(bad, 1);
                }
            }
    }
    __syncthreads();
    cpx z[16], acc[16];
    for (int i = 0; i < 16; i++) { z[i].x = 1.0 + lane + i; z[i].y = 0.5 * i - warp; acc[i].x = acc[i].y = 0.0; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int c = 0; c < kChunks; c++) {
            if (MODE == 0) {
                const cpx *part = S.chunk[c];
#pragma unroll
                for (int pos = 0; pos < 16; pos++) {
                    const cpx w = part[pos * 32 + lane];
                    acc[pos].x = fma(z[pos].x, w.x, acc[pos].x);
                    acc[pos].y = fma(z[pos].x, w.y, acc[pos].y);
                    acc[pos].x = fma(-z[pos].y, w.y, acc[pos].x);
                    acc[pos].y = fma(z[pos].y, w.x, acc[pos].y);
                }
            } else {
                uint32_t v[2][16];
                tmem_ld16(tmem + lane_base + (uint32_t) (c * 64), v[0]);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (g < 3) tmem_ld16(tmem + lane_base + (uint32_t) (c * 64 + (g + 1) * 16), v[(g + 1) & 1]);
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const int pos = g * 4 + p;
                        const double wx = __hiloint2double((int) v[g & 1][4 * p + 1], (int) v[g & 1][4 * p]);
                        const double wy = __hiloint2double((int) v[g & 1][4 * p + 3], (int) v[g & 1][4 * p + 2]);
                        acc[pos].x = fma(z[pos].x, wx, acc[pos].x);
                        acc[pos].y = fma(z[pos].x, wy, acc[pos].y);
                        acc[pos].x = fma(-z[pos].y, wy, acc[pos].x);
                        acc[pos].y = fma(z[pos].y, wx, acc[pos].y);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    for (int i = 0; i < 16; i++) out[(blockIdx.x * 256 + threadIdx.x) * 16 + i] = acc[i];
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
    const int n = kChunks * kChunkCplx, iters = 200, grid = 148;
    cpx *h = (cpx *) malloc(n * sizeof(cpx));
    for (int i = 0; i < n; i++) { h[i].x = 1e-3 * i + 0.25; h[i].y = -2e-3 * i + 0.5; }
    cpx *d_key, *d_out[2]; long long *d_cyc; int *d_bad;
    cudaMalloc(&d_key, n * sizeof(cpx)); cudaMemcpy(d_key, h, n * sizeof(cpx), cudaMemcpyHostToDevice);
    for (int m = 0; m < 2; m++) cudaMalloc(&d_out[m], (size_t) grid * 256 * 16 * sizeof(cpx));
    cudaMalloc(&d_cyc, grid * sizeof(long long)); cudaMalloc(&d_bad, sizeof(int)); cudaMemset(d_bad, 0, sizeof(int));
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem));
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem));
    for (int rep = 0; rep < 2; rep++)
        for (int m = 0; m < 2; m++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (m == 0) probe<0><<<grid, 256, sizeof(Smem)>>>(d_key, d_out[0], d_cyc, iters, d_bad);
            else probe<1><<<grid, 256, sizeof(Smem)>>>(d_key, d_out[1], d_cyc, iters, d_bad);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long c0; cudaMemcpy(&c0, d_cyc, sizeof(c0), cudaMemcpyDeviceToHost);
            printf("mode %d (%s): %s, %.3f ms, %.0f cycles per 8-chunk pass of a warp (8 warps per SM)\n", m, m ? "TMEM" : "LDS",
                   cudaGetErrorString(e), ms, (double) c0 / iters);
        }
    int bad; cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    cpx *o0 = (cpx *) malloc(256 * 16 * sizeof(cpx)), *o1 = (cpx *) malloc(256 * 16 * sizeof(cpx));
    cudaMemcpy(o0, d_out[0], 256 * 16 * sizeof(cpx), cudaMemcpyDeviceToHost);
    cudaMemcpy(o1, d_out[1], 256 * 16 * sizeof(cpx), cudaMemcpyDeviceToHost);
    int diff = 0;
    for (int i = 0; i < 256 * 16; i++) diff += (o0[i].x != o1[i].x) || (o0[i].y != o1[i].y);
    printf("copy/load mismatches: %d, result words differing between the two paths: %d\n", bad, diff);
    return 0;
}
