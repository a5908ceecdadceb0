// LWE key switch for a batch of extracted samples, plus the small linear
// kernels of the gate API (NOT / COPY / CONSTANT / linear combinations).
//
// lweKeySwitch (lwe-keyswitch-functions.cu:955-987) and
// lweKeySwitchTranslate_fromArray (:101-127):
//   res = (0, u.b);  for i < N, j < t:
//     aij = (((uint32)u.a[i] + 2^(31 - basebit*t)) >> (32 - (j+1)*basebit)) & (base-1)
//     if aij != 0: res -= ks[i][j][aij]
// Integer arithmetic mod 2^32: bit-exact whatever the summation order.
//
// One CTA handles a tile of 32 gates for a range of i.  The digits of the tile
// are precomputed into shared memory (2 bits per gate, 16 gates per word, two
// words per (i, j)), then every thread owns four output columns for 16 gates in
// registers and streams the table rows (coalesced 128-bit loads, L2 resident)
// exactly once per tile instead of once per gate as the reference does
// (lweKeySwitchVectorSubstraction_gpu_testing_coalesce_n_Bit, boot-gates.cu:2382-2421).
// Small batches split the i range over several CTAs and combine with integer
// atomics so that a single gate still uses the whole chip.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace tfhe_b200 {

namespace {

constexpr int kTile = kKsTile;  // gates per CTA: two groups of 16
constexpr int kGrp = 16;        // gates per thread
constexpr int kKsThreads = 256; // 128 column quads x 2 gate groups
constexpr int kMaxT = 16;

template <bool kAtomic>
__global__ void __launch_bounds__(kKsThreads, 2) keyswitch_kernel(const KsLaunch L, int nsplit) {
    extern __shared__ uint32_t dig[];  // [i_per * t][2]: digits of gate group 0 / 1
    const int g0 = blockIdx.x * kTile;
    const int ng = min(kTile, L.count - g0);
    const int i_per = L.N / nsplit;
    const int i0 = blockIdx.y * i_per;
    const int t = L.t;
    const size_t ustride = (size_t) L.N + 1;
    const uint32_t prec_offset = 1u << (32 - (1 + L.basebit * t));

    // ---- digits of this tile: dig[((i - i0) * t + j) * 2 + grp], gate g of the group in bits [2g, 2g+2) ----
    for (int w0 = (int) threadIdx.x; w0 < 2 * i_per; w0 += kKsThreads) {
        const int i = i0 + (w0 >> 1), grp = w0 & 1;
        uint32_t w[kMaxT];
#pragma unroll
        for (int j = 0; j < kMaxT; j++) w[j] = 0;
        for (int g = 0; g < kGrp; g++) {
            const int gg = grp * kGrp + g;
            if (gg >= ng) break;
            uint32_t a = (uint32_t) __ldg(L.u + (size_t) (g0 + gg) * ustride + i);
            if (L.nsrc == 2) a += (uint32_t) __ldg(L.u + (size_t) (g0 + gg + L.count) * ustride + i);
            const uint32_t aibar = a + prec_offset;
#pragma unroll
            for (int j = 0; j < kMaxT; j++)
                if (j < t) w[j] |= ((aibar >> (32 - (j + 1) * 2)) & 3u) << (2 * g);
        }
#pragma unroll
        for (int j = 0; j < kMaxT; j++)
            if (j < t) dig[((i - i0) * t + j) * 2 + grp] = w[j];
    }
    __syncthreads();

    // ---- accumulate table rows: 4 columns x 16 gates per thread ----------------
    const int cq = (threadIdx.x & 127) * 4, grp = threadIdx.x >> 7;
    uint32_t acc[4][kGrp];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int g = 0; g < kGrp; g++) acc[c][g] = 0;
    const int32_t *tbl = L.ks + (size_t) i0 * t * 3 * kKsRowWords + cq;
    const int steps = i_per * t;
#pragma unroll 2
    for (int idx = 0; idx < steps; idx++) {
        const uint32_t w = dig[idx * 2 + grp];
        const int32_t *r = tbl + (size_t) idx * 3 * kKsRowWords;
        const uint4 r1 = __ldg(reinterpret_cast<const uint4 *>(r));
        const uint4 r2 = __ldg(reinterpret_cast<const uint4 *>(r + kKsRowWords));
        const uint4 r3 = __ldg(reinterpret_cast<const uint4 *>(r + 2 * kKsRowWords));
        // value(d) = lo*r1 + hi*r2 + (lo&hi)*(r3 - r1 - r2) for d = lo + 2*hi: three integer
        // multiply-adds per cell; the digit bits are extracted once per gate for four columns
        const uint32_t a1[4] = {r1.x, r1.y, r1.z, r1.w}, a2[4] = {r2.x, r2.y, r2.z, r2.w};
        const uint32_t a3[4] = {r3.x - r1.x - r2.x, r3.y - r1.y - r2.y, r3.z - r1.z - r2.z, r3.w - r1.w - r2.w};
#pragma unroll
        for (int g = 0; g < kGrp; g++) {
            const uint32_t lo = (w >> (2 * g)) & 1u, hi = (w >> (2 * g + 1)) & 1u, lh = lo & hi;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c][g]) : "r"(lo), "r"(a1[c]));
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c][g]) : "r"(hi), "r"(a2[c]));
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c][g]) : "r"(lh), "r"(a3[c]));
            }
        }
    }

    // ---- result = (0, u.b + cst) - acc -----------------------------------------
    const int n = L.n;
#pragma unroll
    for (int g = 0; g < kGrp; g++) {
        const int gg = grp * kGrp + g;
        if (gg < ng) {
            int local = g0 + gg, di = 0;
            while (di + 1 < L.ndst && local >= L.dst[di].count) {
                local -= L.dst[di].count;
                di++;
            }
            const long long orow = L.dst[di].idx ? (long long) __ldg(L.dst[di].idx + local) : (long long) local;
            int32_t *row = L.dst[di].out + orow * L.dst[di].stride;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int col = cq + c;
                if (col > n) continue;
                uint32_t v = 0u - acc[c][g];
                if (col == n && blockIdx.y == 0) {
                    v += (uint32_t) __ldg(L.u + (size_t) (g0 + gg) * ustride + L.N) + (uint32_t) L.cst;
                    if (L.nsrc == 2) v += (uint32_t) __ldg(L.u + (size_t) (g0 + gg + L.count) * ustride + L.N);
                }
                if (kAtomic) atomicAdd(reinterpret_cast<unsigned int *>(row + col), v);
                else row[col] = (int32_t) v;
            }
        }
    }
}

__global__ void ks_zero_kernel(const KsLaunch L, int words) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long) L.count * words) return;
    int local = (int) (t / words), di = 0;
    while (di + 1 < L.ndst && local >= L.dst[di].count) {
        local -= L.dst[di].count;
        di++;
    }
    const long long orow = L.dst[di].idx ? (long long) L.dst[di].idx[local] : (long long) local;
    L.dst[di].out[orow * L.dst[di].stride + (t % words)] = 0;
}

// src [N][t][base][n+1] -> dst [N][t][base-1][512] (row h = 0 is the noiseless zero sample,
// lwe-keyswitch-functions.cu:919, and is dropped; columns > n are zero padding)
__global__ void ks_relayout_kernel(const int32_t *__restrict__ src, int32_t *__restrict__ dst, long long rows_out,
                                   int base, int n) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_out * kKsRowWords) return;
    const long long row = t / kKsRowWords;
    const int c = (int) (t % kKsRowWords);
    const long long ij = row / (base - 1);
    const int h = (int) (row % (base - 1)) + 1;
    dst[t] = (c <= n) ? src[(ij * base + h) * (n + 1) + c] : 0;
}

// out = c0*in0 + c1*in1 + (0, cst) on rows of n+1 words
__global__ void lwe_linear_kernel(int32_t *out, long long so, const int32_t *in0, long long s0, int c0,
                                  const int32_t *in1, long long s1, int c1, int32_t cst, int count, int n) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int words = n + 1;
    if (t >= (long long) count * words) return;
    const long long g = t / words;
    const int c = (int) (t % words);
    uint32_t v = 0;
    if (in0) v += (uint32_t) c0 * (uint32_t) in0[g * s0 + c];
    if (in1) v += (uint32_t) c1 * (uint32_t) in1[g * s1 + c];
    if (c == n) v += (uint32_t) cst;
    out[g * so + c] = (int32_t) v;
}

// gather / scatter form: out row idx_out[g] = c0 * (in row idx_in[g]) + (0, cst)
__global__ void lwe_linear_idx_kernel(int32_t *out, const int32_t *in, long long stride, const int32_t *idx_out,
                                      const int32_t *idx_in, int c0, int32_t cst, int count, int n) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int words = n + 1;
    if (t >= (long long) count * words) return;
    const int g = (int) (t / words);
    const int c = (int) (t % words);
    uint32_t v = 0;
    if (c0 != 0) v = (uint32_t) c0 * (uint32_t) in[(long long) __ldg(idx_in + g) * stride + c];
    if (c == n) v += (uint32_t) cst;
    out[(long long) __ldg(idx_out + g) * stride + c] = (int32_t) v;
}

}  // namespace

cudaError_t launch_lwe_linear_idx(int32_t *out, const int32_t *in, long long stride, const int32_t *idx_out,
                                  const int32_t *idx_in, int c0, int32_t cst, int count, int n, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const long long total = (long long) count * (n + 1);
    lwe_linear_idx_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(out, in, stride, idx_out, idx_in, c0,
                                                                                cst, count, n);
    return cudaGetLastError();
}

// zeroes the output rows of a key switch (split launches accumulate into them with atomics)
cudaError_t launch_ks_zero(const KsLaunch &L, cudaStream_t stream) {
    const long long total = (long long) L.count * (L.n + 1);
    if (total <= 0) return cudaSuccess;
    ks_zero_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(L, L.n + 1);
    return cudaGetLastError();
}

cudaError_t launch_keyswitch(const KsLaunch &L, int sm_count, cudaStream_t stream) {
    if (L.count <= 0) return cudaSuccess;
    if (L.basebit != 2 || L.t > kMaxT || L.n + 1 > 2 * kKsThreads) return cudaErrorInvalidValue;
    const int tiles = (L.count + kTile - 1) / kTile;
    int nsplit = 1;
    while (tiles * nsplit < 2 * sm_count && nsplit < 64 && (L.N / (nsplit * 2)) >= 8) nsplit *= 2;
    const size_t smem = (size_t) (L.N / nsplit) * L.t * 2 * sizeof(uint32_t);
    dim3 grid(tiles, nsplit);
    if (smem > 48 * 1024) {  // up to 64 KiB of digits per CTA (attribute is per device: set on every launch)
        cudaError_t e = nsplit > 1
                            ? cudaFuncSetAttribute(keyswitch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)
                            : cudaFuncSetAttribute(keyswitch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        if (e != cudaSuccess) return e;
    }
    if (nsplit > 1) {
        const long long total = (long long) L.count * (L.n + 1);
        ks_zero_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(L, L.n + 1);
        keyswitch_kernel<true><<<grid, kKsThreads, smem, stream>>>(L, nsplit);
    } else {
        keyswitch_kernel<false><<<grid, kKsThreads, smem, stream>>>(L, nsplit);
    }
    return cudaGetLastError();
}

cudaError_t launch_ks_relayout(const int32_t *src, int32_t *dst, int N, int t, int base, int n, cudaStream_t stream) {
    const long long rows_out = (long long) N * t * (base - 1);
    const long long total = rows_out * kKsRowWords;
    ks_relayout_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(src, dst, rows_out, base, n);
    return cudaGetLastError();
}

cudaError_t launch_lwe_linear(int32_t *out, long long out_stride, const int32_t *in0, long long s0, int c0,
                              const int32_t *in1, long long s1, int c1, int32_t cst, int count, int n,
                              cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const long long total = (long long) count * (n + 1);
    lwe_linear_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(out, out_stride, in0, s0, c0, in1, s1, c1,
                                                                            cst, count, n);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
