// LWE key switch on the 5th-generation tensor cores (tcgen05, sm_100a) for large batches.
//
// lweKeySwitch (lwe-keyswitch-functions.cu:955-987) is, for a batch of gates, the one dense
// matrix product of the path:   out[g][c] = b_g [c == n]  -  sum_kappa A[g][kappa] * KS[kappa][c]
// with kappa = (i, j, h), h = 1..3, A[g][kappa] = [digit_{g,i,j} == h] in {0,1} and KS the key-switch
// table (int32 mod 2^32).  The table is split into its four bytes (unsigned 8-bit limbs); one
// `tcgen05.mma.kind::i8` per limb pair accumulates exact int32 sums in tensor memory
// (at most 8192 * 255 < 2^31 per entry); the epilogue recombines  sum_l D_l << 8l  mod 2^32.
// Bit-exact with the SIMT kernel (keyswitch.cu): integer arithmetic, any order.
//
// Why: the SIMT form needs three integer multiply-adds per (gate, i, j, column) and sits at
// that roofline (51 ms per 65536 gates).  Here the arithmetic is 3.3e12 int8 MACs (1.5 ms at
// the dense int8 rate) and the kernel is bound by streaming the 50 MB table from L2 once per
// 128-gate tile (25.8 GB per 65536 gates).
//
// CTA = 128 gates x 128 output columns (x 4 limbs = 512 accumulator columns = all of tensor
// memory), 6 warps:
//   warps 0-3  build the one-hot A tiles in shared memory (lookup of 4 digits at a time), then
//              run the epilogue (warp w owns tensor-memory lanes 32w..32w+31 = gates);
//   warp 4     streams the pre-tiled table with 1-D TMA bulk copies (64 KB per stage);
//   warp 5     allocates tensor memory and issues the MMAs (one elected lane).
// Operand layout in shared memory: K-major, no swizzle, core matrices of 8 rows x 16 bytes
// (UMMA "interleave" canonical layout): byte (row r, k) of a K = 32 slice lives at
//   (k / 16) * LBO + (r / 8) * 128 + (r % 8) * 16 + k % 16,   LBO = 16 * rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace tfhe_b200 {

namespace {

constexpr int kMmaM = 128;            // gates per CTA
constexpr int kColsPerCta = 128;      // output columns per CTA
constexpr int kNTiles = kKsRowWords / kColsPerCta;  // 4
constexpr int kMmaN = 256;            // per instruction; two per K step (4 limbs x 128 columns)
constexpr int kStepK = 32;            // K per instruction (8-bit operands)
// K order: groups of four coefficients; inside a group the digit value h = 1..3 is the slow index:
//   kappa = (group * 3 + (h - 1)) * 32 + (i % 4) * 8 + j,
// so one pipeline stage = one group = three K steps built from FOUR loads of the extracted sample
// (each coefficient is read once, not once per digit value).
constexpr int kStepsPerStage = 3;
#ifndef TFHE_B200_KSM_SLOTS
#define TFHE_B200_KSM_SLOTS 3
#endif
constexpr int kSlots = TFHE_B200_KSM_SLOTS;          // stages in flight
constexpr int kCoefStage = 4;                        // coefficients (i) per stage
#ifndef TFHE_B200_KSM_PREFETCH
#define TFHE_B200_KSM_PREFETCH 8
#endif
constexpr int kPrefetch = TFHE_B200_KSM_PREFETCH;    // stages of sample words kept in flight per thread
constexpr int kABytesStep = kMmaM * kStepK;               // 4 KiB
constexpr int kBBytesStep = 2 * kMmaN * kStepK;           // 16 KiB
constexpr int kABytesStage = kStepsPerStage * kABytesStep;  // 16 KiB
constexpr int kBBytesStage = kStepsPerStage * kBBytesStep;  // 64 KiB
constexpr int kMmaThreads = 192;
constexpr int kIJ = 1024 * 8;                             // (i, j) pairs: N * t
constexpr int kKTotal = 3 * kIJ;                          // 24576
constexpr int kSteps = kKTotal / kStepK;                  // 768
constexpr int kStages = kSteps / kStepsPerStage;          // 256
static_assert(kStages % kPrefetch == 0, "prefetch ring must divide the stage count");

struct __align__(1024) MmaSmem {
    uint8_t a[kSlots][kABytesStage];
    uint8_t b[kSlots][kBBytesStage];
    uint32_t lut[3][256];  // [h-1][byte of four digits] -> four 0/1 bytes
    unsigned long long full_a[kSlots], full_b[kSlots], empty[kSlots], done;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KSM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KSM_DONE;\n"
        "bra KSM_WAIT;\n"
        "KSM_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((addr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;  // descriptor version for sm_100
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = unsigned 8 bit, K-major, N, M
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t) (kMmaN >> 3) << 17) |
                            ((uint32_t) (kMmaM >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// nsplit > 1 (small batches): blockIdx.z takes the stages [z * kStages / nsplit, (z + 1) * kStages / nsplit)
// of the contraction and ADDS its partial result to the (pre-zeroed) output with integer atomics — exact
// whatever the order — so that a few hundred gates do not wait for one CTA per column tile to stream its
// whole quarter of the table (a flat 0.285 ms): the table pass is spread over up to 148 / (4 * tiles) CTAs.
__global__ void __launch_bounds__(kMmaThreads, 1) keyswitch_mma_kernel(const KsLaunch L, const uint8_t *__restrict__ tbl,
                                                                       const int nsplit) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MmaSmem &S = *reinterpret_cast<MmaSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g0 = blockIdx.x * kMmaM;
    const int nt = blockIdx.y;
    const int nst = kStages / nsplit;            // stages of this CTA (a multiple of kPrefetch)
    const int st_first = (int) blockIdx.z * nst;

    // lookup: four digits (one byte of the 16 digit bits of a coefficient) -> 0/1 bytes for h
    for (int e = threadIdx.x; e < 3 * 256; e += kMmaThreads) {
        const int h = e / 256 + 1, b = e % 256;
        uint32_t w = 0;
        for (int j = 0; j < 4; j++)
            if (((b >> (6 - 2 * j)) & 3) == h) w |= 1u << (8 * j);
        S.lut[h - 1][b] = w;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; s++) {
            mbar_init(&S.full_a[s], 128);
            mbar_init(&S.full_b[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        mbar_init(&S.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S.tmem_base)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;

    if (warp < 4) {
        // ================= A producer: thread = gate row =================
        const int m = threadIdx.x;
        const int g = g0 + m;
        const bool valid = g < L.count;
        const size_t ustride = (size_t) L.N + 1;
        const int32_t *u0 = L.u + (size_t) (valid ? g : 0) * ustride;
        const int32_t *u1 = (L.nsrc == 2) ? L.u + (size_t) ((valid ? g : 0) + L.count) * ustride : nullptr;
        const uint32_t prec_offset = 1u << (32 - (1 + 2 * 8));  // lwe-keyswitch-functions.cu:106
        const uint32_t row_off = (uint32_t) (m >> 3) * 128 + (uint32_t) (m & 7) * 16;
        // The sample words of stage st + kPrefetch are requested before the tile of stage st is
        // built (register ring, fully unrolled), so that their L2 / DRAM latency is off the path.
        auto load_digits = [&](int st, uint32_t (&dg)[kCoefStage]) {
#pragma unroll
            for (int k = 0; k < kCoefStage; k++) {
                uint32_t a = 0;
                if (valid) {
                    a = (uint32_t) __ldg(u0 + st * kCoefStage + k);
                    if (u1) a += (uint32_t) __ldg(u1 + st * kCoefStage + k);
                }
                dg[k] = (a + prec_offset) >> 16;  // the eight 2-bit digits, j = 0 on top
            }
        };
        uint32_t ring[kPrefetch][kCoefStage];
#pragma unroll
        for (int p = 0; p < kPrefetch; p++) load_digits(st_first + p, ring[p]);
        for (int st0 = 0; st0 < nst; st0 += kPrefetch) {
#pragma unroll
            for (int p = 0; p < kPrefetch; p++) {
                const int st = st0 + p;   // local stage number: slot and phase arithmetic
                const int slot = st % kSlots;
                const uint32_t phase = (uint32_t) (st / kSlots) & 1u;
                uint32_t dg[kCoefStage];
#pragma unroll
                for (int k = 0; k < kCoefStage; k++) dg[k] = ring[p][k];
                if (st + kPrefetch < nst) load_digits(st_first + st + kPrefetch, ring[p]);
                mbar_wait(&S.empty[slot], phase ^ 1u);
                uint8_t *abase = S.a[slot] + row_off;
#pragma unroll
                for (int h = 0; h < 3; h++) {      // K step h of the stage: digit value h + 1
                    const uint32_t *lut = S.lut[h];
#pragma unroll
                    for (int k2 = 0; k2 < 2; k2++) {  // 16-byte chunk: coefficients 2*k2, 2*k2 + 1 of the group
                        uint4 v;
                        v.x = valid ? lut[dg[2 * k2] >> 8] : 0u;
                        v.y = valid ? lut[dg[2 * k2] & 255u] : 0u;
                        v.z = valid ? lut[dg[2 * k2 + 1] >> 8] : 0u;
                        v.w = valid ? lut[dg[2 * k2 + 1] & 255u] : 0u;
                        *reinterpret_cast<uint4 *>(abase + h * kABytesStep + k2 * 2048) = v;  // LBO = 16 * 128 rows
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
                mbar_arrive(&S.full_a[slot]);
            }
        }
        // ================= epilogue =================
        mbar_wait(&S.done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int32_t *row = nullptr;
        uint32_t bval = 0;
        if (valid) {
            int local = g, di = 0;
            while (di + 1 < L.ndst && local >= L.dst[di].count) {
                local -= L.dst[di].count;
                di++;
            }
            const long long orow = L.dst[di].idx ? (long long) __ldg(L.dst[di].idx + local) : (long long) local;
            row = L.dst[di].out + orow * L.dst[di].stride;
            if (blockIdx.z == 0) {  // the body (and the constant) are added once
                bval = (uint32_t) __ldg(u0 + L.N) + (uint32_t) L.cst;
                if (u1) bval += (uint32_t) __ldg(u1 + L.N);
            }
        }
        const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
        for (int c0 = 0; c0 < kColsPerCta; c0 += 16) {
            uint32_t d0[16], d1[16], d2[16], d3[16];
            tmem_ld16(tmem + lane_base + (uint32_t) (0 * 128 + c0), d0);
            tmem_ld16(tmem + lane_base + (uint32_t) (1 * 128 + c0), d1);
            tmem_ld16(tmem + lane_base + (uint32_t) (2 * 128 + c0), d2);
            tmem_ld16(tmem + lane_base + (uint32_t) (3 * 128 + c0), d3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int col = nt * kColsPerCta + c0 + k;
                    if (col > L.n) continue;
                    uint32_t v = 0u - (d0[k] + (d1[k] << 8) + (d2[k] << 16) + (d3[k] << 24));
                    if (col == L.n) v += bval;
                    if (nsplit > 1) atomicAdd(reinterpret_cast<unsigned int *>(row + col), v);
                    else row[col] = (int32_t) v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ================= table loader =================
        if (lane == 0) {
            const uint8_t *src = tbl + ((size_t) nt * kSteps + (size_t) st_first * kStepsPerStage) * kBBytesStep;
            for (int st = 0; st < nst; st++) {
                const int slot = st % kSlots;
                const uint32_t phase = (uint32_t) (st / kSlots) & 1u;
                mbar_wait(&S.empty[slot], phase ^ 1u);
                mbar_expect_tx(&S.full_b[slot], kBBytesStage);
#pragma unroll
                for (int q = 0; q < kStepsPerStage; q++)
                    tma_load_1d(S.b[slot] + q * kBBytesStep, src + ((size_t) st * kStepsPerStage + q) * kBBytesStep,
                                kBBytesStep, &S.full_b[slot]);
            }
        }
    } else {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int st = 0; st < nst; st++) {
                const int slot = st % kSlots;
                const uint32_t phase = (uint32_t) (st / kSlots) & 1u;
                mbar_wait(&S.full_a[slot], phase);
                mbar_wait(&S.full_b[slot], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_addr(S.a[slot]), b0 = smem_addr(S.b[slot]);
#pragma unroll
                for (int q = 0; q < kStepsPerStage; q++) {
                    const uint64_t ad = smem_desc(a0 + q * kABytesStep, 16 * kMmaM, 128);
#pragma unroll
                    for (int n2 = 0; n2 < 2; n2++) {
                        const uint64_t bd = smem_desc(b0 + q * kBBytesStep + n2 * (kMmaN * kStepK), 16 * kMmaN, 128);
                        mma_i8(tmem + (uint32_t) (n2 * kMmaN), ad, bd, (st | q) != 0 ? 1u : 0u);
                    }
                }
                mma_commit(&S.empty[slot]);  // frees the slot when these MMAs have read it
            }
            mma_commit(&S.done);
        }
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// src [N][t][base][n+1] int32 -> tiled unsigned bytes
//   [nt 4][K step 768][n2 2][k chunk 2][row group 32][row 8][16 bytes],
// row n (0..255) of half n2 = limb (2*n2 + n/128), output column nt*128 + n%128;
// K step = 3 * (i / 4) + (h - 1), byte of the step = (i % 4) * 8 + j.
__global__ void ks_mma_relayout_kernel(const int32_t *__restrict__ src, uint8_t *__restrict__ dst, int base, int n) {
    const long long tid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long) kNTiles * kSteps * kBBytesStep;
    if (tid >= total) return;
    const int kt = (int) (tid & 15);
    const int r8 = (int) ((tid >> 4) & 7);
    const int ng = (int) ((tid >> 7) & 31);
    const int k2 = (int) ((tid >> 12) & 1);
    const int n2 = (int) ((tid >> 13) & 1);
    const long long rest = tid >> 14;
    const int step = (int) (rest % kSteps);
    const int nt = (int) (rest / kSteps);
    const int nrow = ng * 8 + r8;
    const int limb = 2 * n2 + nrow / 128;
    const int col = nt * kColsPerCta + nrow % 128;
    const int h = step % 3 + 1;
    const int i = 4 * (step / 3) + (k2 * 16 + kt) / 8, j = kt % 8;
    const int ij = i * 8 + j;
    uint32_t v = 0;
    if (col <= n) v = (uint32_t) src[((long long) ij * base + h) * (n + 1) + col];
    dst[tid] = (uint8_t) ((v >> (8 * limb)) & 255u);
}

}  // namespace

size_t ks_mma_table_bytes() { return (size_t) kNTiles * kSteps * kBBytesStep; }

bool ks_mma_supported(int N, int t, int basebit, int n) { return N == 1024 && t == 8 && basebit == 2 && n <= 511; }

cudaError_t launch_ks_mma_relayout(const int32_t *src, uint8_t *dst, int base, int n, cudaStream_t stream) {
    const long long total = (long long) ks_mma_table_bytes();
    ks_mma_relayout_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(src, dst, base, n);
    return cudaGetLastError();
}

cudaError_t launch_keyswitch_mma(const KsLaunch &L, const uint8_t *tbl, int sm_count, cudaStream_t stream) {
    if (L.count <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(keyswitch_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) sizeof(MmaSmem) + 1024);
    if (e != cudaSuccess) return e;
    const int tiles = (L.count + kMmaM - 1) / kMmaM;
    // split the contraction while all CTAs still fit one wave (one CTA per SM: 240 KB of shared memory)
    int nsplit = 1;
    while (nsplit < kStages / kPrefetch && tiles * kNTiles * nsplit * 2 <= sm_count) nsplit *= 2;
    if (nsplit > 1) {
        e = launch_ks_zero(L, stream);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(tiles, kNTiles, nsplit);
    keyswitch_mma_kernel<<<grid, kMmaThreads, sizeof(MmaSmem) + 1024, stream>>>(L, tbl, nsplit);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
