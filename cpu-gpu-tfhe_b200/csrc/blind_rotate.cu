// Persistent blind-rotation + sample-extraction kernel for sm_100a.
//
// Replaces the reference's host-driven loop of 2500 launches + 1000 cuFFT calls
// per gate batch (bootstrapAndKeySwitch_n_Bit, boot-gates.cu:2481-2629, kernels
// :2127-2362) and implements tfhe_blindRotateAndExtract_FFT
// (lwe-bootstrapping-functions-fft.cu:1408-1456) for a whole batch:
//
//   * a pair of warps owns one ciphertext; its TLWE accumulator (8 KB) and the FFT
//     exchange buffers stay in shared memory for all n iterations; the
//     transforms run in registers (br_core.cuh) and only pair-level named
//     barriers are needed between phases; both warps do identical work between
//     barriers (two warps per SM sub-partition hide each other's latencies: ncu of
//     the one-warp-per-ciphertext version showed 43 % fp64 pipe use);
//   * the rotation X^a * ACC is read at immediate offsets from a negacyclically extended
//     copy of the accumulator (br_core.cuh phase_f1_decomp), sign / subtraction /
//     decomposition offset / digit shift are two integer multiply-adds per coefficient;
//   * 4 ciphertexts per CTA share each 16 KiB row of the bootstrapping key,
//     streamed from L2/HBM with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     expect_tx) into a 3-stage ring per role; a stage is released right behind the
//     last load of its chunk and the last warp to release it issues the refill, so
//     there is no producer warp (small batches: the idle last slot of the CTA refills);
//   * the gate's linear prologue and the mod-switch are computed on the fly
//     from the input samples (no temporaries in global memory);
//   * grid = min(#groups, #SMs) CTAs, each looping over groups of 4 ciphertexts;
//   * the two resources that bind it are the fp64 pipe and the shared-memory pipe: the gadget digits become
//     doubles on the conversion pipe (I2F), and a lane's pass-2 multipliers live in tensor memory (lane-private
//     constants: tcgen05.st once, tcgen05.ld per stage) instead of shared memory + derivations;
//   * batches of at most one / two ciphertexts per SM run on two latency kernels (below) that spread ONE
//     ciphertext over eight / four warps.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "br_core.cuh"
#include "kernels.h"

namespace tfhe_b200 {

namespace {

constexpr int kCtWarps = 4;                       // ciphertexts per CTA
// Two warps per ciphertext (8 warps per CTA, 2 per SM sub-partition, 255 registers each).
// A one-warp-per-ciphertext schedule (4 warps per CTA) was measured too: it moves fewer
// shared-memory bytes per iteration but leaves every sub-partition with a single warp that
// cannot hide its own latencies (profiles/README.md: 503 ms vs 455 ms per 65536 gates).
constexpr int kThreads = 2 * kCtWarps * 32;
// Key ring.  Role r (= accumulator polynomial r) multiplies the decomposed rows 2r and 2r+1; per
// iteration it streams four 8 KiB chunks (one result polynomial of one TGSW row each) in the order
//   (row 2r, half r), (row 2r, half 1-r), (row 2r+1, half r), (row 2r+1, half 1-r)
// i.e. always the half this role KEEPS first, then the half it GIVES to its partner.  Each role has
// its own 3-stage ring so that every consumer of a ring takes every chunk in order (a warp
// skipping chunks could get two mbarrier phases ahead on a stage: parity aliasing).
// Measured and dropped (profiles/README.md): 4 KiB / 2 KiB chunks (every extra wait + release per
// iteration costs about 500 cycles), one key stream per ciphertext pair with the pairs forced half
// an iteration out of phase, release by mbarrier arrive with a lazily refilling designated warp.
constexpr int kChunkPos = 16;                          // positions per chunk
constexpr int kChunkCplx = kChunkPos * 32;
constexpr uint32_t kStageBytes = kChunkCplx * sizeof(cpx);
constexpr uint32_t kRingStages = 3;                    // per role
constexpr uint32_t kChunksPerIter = 4;                 // per role
constexpr int kStages = 2 * (int) kRingStages;
// positions of a chunk multiplied AFTER its stage release has been issued (mac_consume)
#ifndef TFHE_B200_NO_REFILL_FENCE
#define TFHE_B200_NO_REFILL_FENCE 0   // 1: experiment, no proxy fence in front of a ring refill
#endif
#ifndef TFHE_B200_BR_MAC_TAIL
#define TFHE_B200_BR_MAC_TAIL 6
#endif
// TFHE_B200_RING_STRICT=1: the stage release carries a true data dependency on the last loads of the
// chunk (+ an acquire fence in front of the refill), so the write-after-read order "all reads of a
// stage, then its TMA refill" no longer rests on the shared-memory pipe serving the loads and the
// atomic of a warp in issue order.  Measured: +3.2 % kernel time (395.9 vs 383.7 ms per 65536 gates;
// an atom.release instead: +2.0 %), because the release then leaves a load latency later, four times
// per iteration.  The default (0) keeps the early release: it is HARDWARE DEPENDENT (in-order LDS /
// ATOMS per warp, true on sm_100a), and guarded by tests/test_gpu_parity.py::test_key_ring_stress_*
// (exact results under maximal skip ratios and mixed idle slots); build with =1 for the strict form.
#ifndef TFHE_B200_RING_STRICT
#define TFHE_B200_RING_STRICT 0
#endif
// TFHE_B200_E2_TMEM=1: the full-batch kernel keeps a lane's eight pass-2 multipliers (the four base values of its
// frequency class and the four derived ones) in TENSOR MEMORY: 32 words per lane, written once with tcgen05.st,
// read per stage with tcgen05.ld — instead of 4 LDS.128 and 16 fp64 instructions per transform (three transforms
// per warp and iteration).  Tensor memory is lane private, which is exactly what these constants are; its load
// path uses neither the shared-memory pipe nor the fp64 pipe, the two that bind this kernel.  (Not in the
// small-batch instantiation: with three ciphertexts per CTA the load latency is not hidden, 444 gates 3.13 -> 3.29 ms.)
#ifndef TFHE_B200_E2_TMEM
#define TFHE_B200_E2_TMEM 1
#endif

struct __align__(128) CtaSmem {
    WarpSmem w[kCtWarps];
    cpx e2[32 * kE2Row];
    cpx ring[kStages][kChunkCplx];
    unsigned long long full[kStages];   // mbarriers: TMA completion of a ring stage
    unsigned int drained[kStages];      // warps that have finished with the stage's current chunk
    uint32_t tmem_base;                 // tensor-memory allocation (TFHE_B200_E2_TMEM)
};

static_assert(sizeof(WarpSmem) % 16 == 0, "warp working set must keep 16 B alignment");
static_assert(offsetof(CtaSmem, ring) % 128 == 0, "TMA destination alignment");
static_assert(sizeof(CtaSmem) <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Non-blocking probe: has the phase with this parity completed?  (acquire, like the wait)
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA unit).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Position of one warp in its role's ring: ring-local chunk number, stage and phase parity.
struct RingPos {
    uint32_t stage = 0, phase = 0, chunk = 0;
    __device__ __forceinline__ void advance(uint32_t nstages) {
        chunk++;
        if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
        }
    }
};

__device__ __forceinline__ void build_e2(cpx *e2) {
    for (int t = threadIdx.x; t < 32 * 4; t += blockDim.x) {
        const int m1 = t / 4, idx = t % 4;
        double s, c;
        sincospi(e2_shift(m1, idx), &s, &c);
        e2[m1 * kE2Row + idx].x = c;
        e2[m1 * kE2Row + idx].y = s;
    }
}

// modSwitchFromTorus32(x, 2N), numeric-functions.cu:60-66  ==  ((uint32)x + 2^20) >> 21
__device__ __forceinline__ int modswitch_2N(uint32_t x) { return (int) ((x + (1u << 20)) >> 21); }

#ifndef TFHE_B200_PHASE_TIMING
#define TFHE_B200_PHASE_TIMING 0
#endif
#if TFHE_B200_PHASE_TIMING
// Development aid: cycles spent per phase by warp 0 of CTA 0 (tools/phase_timing.py).
__device__ long long g_phase_cycles[16];
#define PHASE_T0() long long pt_ = clock64()
#define PHASE_MARK(i)                                              \
    do {                                                           \
        const long long now_ = clock64();                          \
        if (blockIdx.x == 0 && threadIdx.x == TFHE_B200_PHASE_THREAD) atomicAdd((unsigned long long *) &g_phase_cycles[i], (unsigned long long) (now_ - pt_));   \
        pt_ = now_;                                                \
    } while (0)
#ifndef TFHE_B200_PHASE_THREAD
#define TFHE_B200_PHASE_THREAD 0
#endif
#else
#define PHASE_T0() do {} while (0)
#define PHASE_MARK(i) do {} while (0)
#endif

// Issue the TMA copy of chunk `sub` (0..kChunksPerIter-1) of blind-rotation iteration `it` of role
// `role`'s stream into absolute stage `stage` (one elected lane).
__device__ __forceinline__ void ring_fill_at(CtaSmem &S, const BrLaunch &L, int role, uint32_t it, uint32_t sub,
                                             uint32_t stage) {
    const uint32_t row = 2u * (uint32_t) role + (sub >> 1);
    const uint32_t out = (sub & 1u) ? 1u - (uint32_t) role : (uint32_t) role;
    const cpx *src = L.bk + ((size_t) (L.bk_first + it) * kKpl + row) * kBkRowCplx + out * kBkHalfCplx;
    mbar_arrive_expect_tx(&S.full[stage], kStageBytes);
    tma_load_1d(S.ring[stage], src, kStageBytes, &S.full[stage]);
}

// The same by running chunk number (priming only: the modulo is a division by a run-time value).
__device__ __forceinline__ void ring_fill(CtaSmem &S, const BrLaunch &L, int role, uint32_t chunk, uint32_t stage) {
    ring_fill_at(S, L, role, (chunk / kChunksPerIter) % (uint32_t) L.n_iter, chunk % kChunksPerIter, stage);
}

// Where a consumer is in the key stream: ring position plus (iteration, chunk of the iteration),
// kept incrementally so that a refill needs no division.
struct StreamPos {
    RingPos rp;
    uint32_t it = 0, sub = 0;
    __device__ __forceinline__ void advance(uint32_t n_iter) {
        rp.advance(kRingStages);
        if (++sub == kChunksPerIter) {
            sub = 0;
            if (++it == n_iter) it = 0;
        }
    }
};

// Has the chunk `ahead` positions after the current one already landed?  Probed EARLY (before the
// transform that precedes a multiply) so that the barrier round trip is off the critical path.
__device__ __forceinline__ bool ring_probe(CtaSmem &S, const StreamPos &sp, uint32_t ring_base, int ahead) {
    RingPos r = sp.rp;
    for (int i = 0; i < ahead; i++) r.advance(kRingStages);
    return mbar_test(&S.full[ring_base + r.stage], r.phase);
}

// Release of the current stage by one warp (lane 0 only; `seen` = result of the drained-counter
// atomic): the last of the four consumers refills the stage at once with the chunk kRingStages ahead.
// (A release by mbarrier arrive with a designated, lazily refilling warp was 3.5 % slower: the
// sooner the refill goes out, the better.)
__device__ __forceinline__ void ring_refill_if_last(CtaSmem &S, const BrLaunch &L, int role, const StreamPos &sp,
                                                    uint32_t st, unsigned int seen, uint32_t ring_chunks) {
    if ((seen % kCtWarps) == kCtWarps - 1 && sp.rp.chunk + kRingStages < ring_chunks) {
        static_assert(kRingStages < kChunksPerIter, "refill target is at most one iteration ahead");
        uint32_t fsub = sp.sub + kRingStages, fit = sp.it;
        if (fsub >= kChunksPerIter) {
            fsub -= kChunksPerIter;
            if (++fit >= (uint32_t) L.n_iter) fit = 0;
        }
#if TFHE_B200_RING_STRICT
        asm volatile("fence.acq_rel.cta;" ::: "memory");  // acquire: the other consumers' releases
#endif
#if !TFHE_B200_NO_REFILL_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        ring_fill_at(S, L, role, fit, fsub, st);
    }
}

// A warp with nothing to compute keeps its place in the stream.
// HELPER (small batches, cts_per_group < 4: the last ciphertext slot of the CTA is always idle): the
// idle warp of that slot does ALL refills of its role's ring — it releases, waits until the other
// three consumers have released too, and issues the copy — so that the computing warps never pay
// the proxy fence and the TMA issue (a lone ciphertext is always the last to release: 4 refills
// per iteration on its critical path otherwise; single gate 2.36 -> 2.16 ms).  Lending that slot's
// working set to the ring as two more stages per role was measured on top: no further gain.
// In the HELPER instantiation the slots between the last used one and the helper slot leave the kernel
// at once (they would only spin next to the computing warps of their sub-partition), so a stage has
// `ncons` = cts_per_group + 1 consumers there.
template <bool HELPER>
__device__ __forceinline__ void ring_skip(CtaSmem &S, const BrLaunch &L, int role, int lane, bool designated,
                                          StreamPos &sp, uint32_t ring_base, uint32_t ring_chunks,
                                          unsigned int ncons) {
    const uint32_t st = ring_base + sp.rp.stage;
    mbar_wait(&S.full[st], sp.rp.phase);
    __syncwarp();
    if (lane == 0) {
        const unsigned int seen = atomicAdd(&S.drained[st], 1u);  // an idle warp reads nothing from the stage
        if (!HELPER) {
            ring_refill_if_last(S, L, role, sp, st, seen, ring_chunks);
        } else if (designated) {
            const unsigned int target = seen - (seen % ncons) + ncons;  // all releases of this chunk
            auto poll = [&]() {
                unsigned int v;
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&S.drained[st])) : "memory");
                return v;
            };
            while ((int) (poll() - target) < 0) __nanosleep(40);
            ring_refill_if_last(S, L, role, sp, st, (unsigned int) (kCtWarps - 1), ring_chunks);
        }
    }
    sp.advance((uint32_t) L.n_iter);
}

// ---- pass-2 multipliers in tensor memory (full-batch kernel) ---------------------------------------------
// columns of a lane: 0 g0, 4 g1, 8 g2, 12 g2 e^{i pi/4}, 16 g3, 20 g3 e^{i pi/4}, 24 g3 e^{i pi/8}, 28 g3 e^{3 i pi/8}
// (one complex double = 4 columns); the derived values are computed exactly as pass2_stage derives them.
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tm_ld_cpx(uint32_t taddr, cpx &a) {
    uint32_t v[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
    a.x = __hiloint2double((int) v[1], (int) v[0]);
    a.y = __hiloint2double((int) v[3], (int) v[2]);
}

// raw forms: the registers are only valid after tm_wait_ld()
struct TmWords4 { uint32_t v[4]; };
struct TmWords8 { uint32_t v[8]; };
struct TmWords16 { uint32_t v[16]; };
__device__ __forceinline__ void tm_ld4(uint32_t taddr, TmWords4 &w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w.v[0]), "=r"(w.v[1]), "=r"(w.v[2]), "=r"(w.v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, TmWords8 &w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w.v[0]), "=r"(w.v[1]), "=r"(w.v[2]), "=r"(w.v[3]), "=r"(w.v[4]), "=r"(w.v[5]), "=r"(w.v[6]), "=r"(w.v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_ld16w(uint32_t taddr, TmWords16 &w) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(w.v[0]), "=r"(w.v[1]), "=r"(w.v[2]), "=r"(w.v[3]), "=r"(w.v[4]), "=r"(w.v[5]), "=r"(w.v[6]), "=r"(w.v[7]),
          "=r"(w.v[8]), "=r"(w.v[9]), "=r"(w.v[10]), "=r"(w.v[11]), "=r"(w.v[12]), "=r"(w.v[13]), "=r"(w.v[14]), "=r"(w.v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ cpx tm_cpx(const uint32_t *v) {
    cpx a;
    a.x = __hiloint2double((int) v[1], (int) v[0]);
    a.y = __hiloint2double((int) v[3], (int) v[2]);
    return a;
}

// every lane writes the eight multipliers of its frequency class (= lane) to its 32 columns
__device__ __forceinline__ void tm_store_consts(uint32_t taddr, const cpx *e) {
    const cpx g2 = e[2], g3 = e[3];
    cpx c[8];
    c[0] = e[0];
    c[1] = e[1];
    c[2] = g2;
    c[3].x = (g2.x - g2.y) * kSqrtHalf;
    c[3].y = (g2.x + g2.y) * kSqrtHalf;
    c[4] = g3;
    c[5].x = (g3.x - g3.y) * kSqrtHalf;
    c[5].y = (g3.x + g3.y) * kSqrtHalf;
    c[6] = cmul_const(g3, kCosPi8, kSinPi8);
    c[7] = cmul_const(g3, kSinPi8, kCosPi8);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[4 * k + 0] = (uint32_t) __double2loint(c[4 * h + k].x);
            v[4 * k + 1] = (uint32_t) __double2hiint(c[4 * h + k].x);
            v[4 * k + 2] = (uint32_t) __double2loint(c[4 * h + k].y);
            v[4 * k + 3] = (uint32_t) __double2hiint(c[4 * h + k].y);
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                taddr + 16u * (uint32_t) h),
            "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
            "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
            : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// pass 2 forward / inverse with the multipliers streamed from tensor memory one stage ahead of their use
// (w0: the first stage's multiplier, loaded by the caller ahead of the shared-memory reads of z)
__device__ __forceinline__ void fwd16_tm(cpx (&z)[16], uint32_t ta, const TmWords4 &w0) {
    TmWords4 w1;
    TmWords8 w2;
    TmWords16 w3;
    tm_wait_ld();
    tm_ld4(ta + 4, w1);
    {
        const cpx g = tm_cpx(w0.v);
        Pass2Stage<0, false>::run(z, g, g, g, g);
    }
    tm_wait_ld();
    tm_ld8(ta + 8, w2);
    {
        const cpx g = tm_cpx(w1.v);
        Pass2Stage<1, false>::run(z, g, g, g, g);
    }
    tm_wait_ld();
    tm_ld16w(ta + 16, w3);
    {
        const cpx g = tm_cpx(w2.v), h4 = tm_cpx(w2.v + 4);
        Pass2Stage<2, false>::run(z, g, h4, g, g);
    }
    tm_wait_ld();
    Pass2Stage<3, false>::run(z, tm_cpx(w3.v), tm_cpx(w3.v + 4), tm_cpx(w3.v + 8), tm_cpx(w3.v + 12));
}

// (w3: the first stage's four multipliers, loaded by the caller ahead of the hand-over additions)
__device__ __forceinline__ void inv16_tm(cpx (&z)[16], uint32_t ta, const TmWords16 &w3) {
    TmWords4 w0, w1;
    TmWords8 w2;
    tm_wait_ld();
    tm_ld8(ta + 8, w2);
    Pass2Stage<3, true>::run(z, tm_cpx(w3.v), tm_cpx(w3.v + 4), tm_cpx(w3.v + 8), tm_cpx(w3.v + 12));
    tm_wait_ld();
    tm_ld4(ta + 4, w1);
    {
        const cpx g = tm_cpx(w2.v), h4 = tm_cpx(w2.v + 4);
        Pass2Stage<2, true>::run(z, g, h4, g, g);
    }
    tm_wait_ld();
    tm_ld4(ta, w0);
    {
        const cpx g = tm_cpx(w1.v);
        Pass2Stage<1, true>::run(z, g, g, g, g);
    }
    tm_wait_ld();
    {
        const cpx g = tm_cpx(w0.v);
        Pass2Stage<0, true>::run(z, g, g, g, g);
    }
}

// Operands of the gate prologue of bootstrap g (x = (0,cst) + sa*in0 + sb*in1 [+ sc*in2 [+ sd*in3]]).
struct GateIn {
    const int32_t *in0 = nullptr, *in1 = nullptr, *in2 = nullptr, *in3 = nullptr;
    uint32_t sa = 0, sb = 0, sc = 0, sd = 0, cst = 0;
};

// word `idx` of the linear combination (idx = n: the body, incl. the gate's constant)
__device__ __forceinline__ uint32_t prologue_word(const GateIn &I, int idx, uint32_t cst) {
    uint32_t x = cst + I.sa * (uint32_t) __ldg(I.in0 + idx) + I.sb * (uint32_t) __ldg(I.in1 + idx);
    if (I.in2 != nullptr) x += I.sc * (uint32_t) __ldg(I.in2 + idx);
    if (I.in3 != nullptr) x += I.sd * (uint32_t) __ldg(I.in3 + idx);
    return x;
}

__device__ __forceinline__ GateIn resolve_inputs(const BrLaunch &L, int g) {
    GateIn I;
    if (L.explicit_inputs != 0) return I;
    int local = g;
    int si = 0;
    while (si + 1 < L.nseg && local >= L.seg[si].count) {
        local -= L.seg[si].count;
        si++;
    }
    const long long r0 = L.seg[si].idx0 ? (long long) __ldg(L.seg[si].idx0 + local) : (long long) local;
    const long long r1 = L.seg[si].idx1 ? (long long) __ldg(L.seg[si].idx1 + local) : (long long) local;
    I.in0 = L.seg[si].in0 + r0 * L.seg[si].stride0;
    I.in1 = L.seg[si].in1 + r1 * L.seg[si].stride1;
    I.sa = (uint32_t) L.seg[si].sa;
    I.sb = (uint32_t) L.seg[si].sb;
    if (L.seg[si].in2 != nullptr) {
        const long long r2 = L.seg[si].idx2 ? (long long) __ldg(L.seg[si].idx2 + local) : (long long) local;
        I.in2 = L.seg[si].in2 + r2 * L.seg[si].stride2;
        I.sc = (uint32_t) L.seg[si].sc;
    }
    if (L.seg[si].in3 != nullptr) {
        const long long r3 = L.seg[si].idx3 ? (long long) __ldg(L.seg[si].idx3 + local) : (long long) local;
        I.in3 = L.seg[si].in3 + r3 * L.seg[si].stride3;
        I.sd = (uint32_t) L.seg[si].sd;
    }
    I.cst = (uint32_t) L.seg[si].cst;
    return I;
}

// bara of iteration idx (mod-switched mask word of the prologue's linear combination)
__device__ __forceinline__ int load_bara(const BrLaunch &L, const GateIn &I, int g, int idx, int n_iter, bool rotate) {
    if (idx >= n_iter || !rotate) return 0;
    if (L.explicit_inputs != 0) return __ldg(L.bara + (size_t) g * n_iter + idx) & (2 * kN - 1);
    return modswitch_2N(prologue_word(I, idx, 0u));
}

// TFHE_B200_EXP_NOBAR: timing-only experiment (garbage results): the pair barriers cost nothing
#ifndef TFHE_B200_EXP_NOBAR
#define TFHE_B200_EXP_NOBAR 0
#endif
__device__ __forceinline__ void named_sync(int id, int nthreads) {
    if (!TFHE_B200_EXP_NOBAR) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void named_arrive(int id, int nthreads) {
    if (!TFHE_B200_EXP_NOBAR) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Fourier MAC of z against the current ring chunk, with the release of the stage issued EARLY: the
// drained-counter atomic goes out right behind the last load of the chunk (the shared-memory pipe
// serves one warp's requests in order, so every lane's reads precede it, and the refill is at
// least three more atomics, a proxy fence and an L2 round trip away), and its result is only
// looked at after the remaining multiply-adds.  With the release after the arithmetic every chunk
// paid the atomic's round trip and the refill left later (4.4 % of the kernel).
// `ready`: the chunk was seen complete by an earlier probe.
template <bool HELPER>
__device__ __forceinline__ void mac_consume(CtaSmem &S, const BrLaunch &L, int role, int lane, StreamPos &sp,
                                            uint32_t ring_base, uint32_t ring_chunks, bool ready,
                                            const cpx (&z)[16], cpx (&acc)[16]) {
    const uint32_t st = ring_base + sp.rp.stage;
    if (!ready) mbar_wait(&S.full[st], sp.rp.phase);
    const cpx *part = S.ring[st];
    constexpr int kTail = TFHE_B200_BR_MAC_TAIL, kHead = kChunkPos - kTail;
    phase_mac_part<0, kHead>(lane, z, part, acc);
    cpx w[kTail];
#pragma unroll
    for (int p = 0; p < kTail; p++) w[p] = part[(kHead + p) * 32 + lane];
    unsigned int seen = 0;
#if TFHE_B200_RING_STRICT
    // The release must not overtake the reads of the stage.  The atomic is predicated on a value
    // computed from the LAST loads of the chunk (their high words, never all-ones for finite doubles:
    // the predicate is always true for lane 0, but the hardware cannot issue the atomic before those
    // loads have returned, and a warp's shared-memory loads return in order), and the refilling warp
    // executes an acquire fence before the proxy fence and the TMA copy.  The write-after-read order
    // "every consumer's reads, then the refill" therefore rests on a true data dependency, not on the
    // shared-memory pipe serving loads and atomics of a warp in issue order (which the first version
    // relied on).  An atom.release here (a membar in front of every release) cost 2 % of the kernel.
    unsigned int live = 0xffffffffu;
#pragma unroll
    for (int p = 0; p < kTail; p++) live &= (unsigned int) __double2hiint(w[p].y);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %2, 0;\n"
        "setp.ne.and.u32 p, %3, 0xffffffff, p;\n"
        "@p atom.shared.add.u32 %0, [%1], 1;\n"
        "}\n"
        : "+r"(seen)
        : "r"(smem_u32(&S.drained[st])), "r"(lane), "r"(live)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %2, 0;\n"
        "@p atom.shared.add.u32 %0, [%1], 1;\n"
        "}\n"
        : "+r"(seen)
        : "r"(smem_u32(&S.drained[st])), "r"(lane)
        : "memory");
#endif
#pragma unroll
    for (int p = 0; p < kTail; p++) cmac(acc[kHead + p], z[kHead + p], w[p]);
    // (one condition, written out here rather than through ring_refill_if_last: the compiler
    // schedules the call form 1.7 % slower)
    if (!HELPER && lane == 0 && (seen % kCtWarps) == kCtWarps - 1 && sp.rp.chunk + kRingStages < ring_chunks) {
        uint32_t fsub = sp.sub + kRingStages, fit = sp.it;
        if (fsub >= kChunksPerIter) {
            fsub -= kChunksPerIter;
            if (++fit >= (uint32_t) L.n_iter) fit = 0;
        }
#if TFHE_B200_RING_STRICT
        asm volatile("fence.acq_rel.cta;" ::: "memory");  // acquire: the other consumers' releases
#endif
#if !TFHE_B200_NO_REFILL_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        ring_fill_at(S, L, role, fit, fsub, st);
    }
    sp.advance((uint32_t) L.n_iter);
}

// HELPER: instantiation for small batches (cts_per_group < 4), see ring_skip.
template <bool HELPER>
__global__ void __launch_bounds__(kThreads, 1) blind_rotate_kernel(const BrLaunch L) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaSmem &S = *reinterpret_cast<CtaSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Small batches are spread over the SMs: a group holds `cpg` <= 4 ciphertexts (the other
    // warp pairs of the CTA only keep the key ring moving), so that a single gate or a narrow
    // circuit level does not share its SM with copies of itself.
    const int cpg = L.cts_per_group;
    const int ngroups = (L.total + cpg - 1) / cpg;
    const int n_iter = L.n_iter;
    const int my_groups = (ngroups - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;
    const uint32_t iters_total = (uint32_t) my_groups * (uint32_t) n_iter;

    build_e2(S.e2);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&S.full[s], 1);
            S.drained[s] = 0;
        }

        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // prime the rings; afterwards the consumers refill them (no producer warp: a ninth warp
        // would cap the kernel at 168 registers per thread)
        for (int r = 0; r < 2; r++)
            for (uint32_t c = 0; c < kRingStages && c < kChunksPerIter * iters_total; c++)
                ring_fill(S, L, r, c, r * kRingStages + c);
    }
    constexpr bool kE2Tm = !HELPER && TFHE_B200_E2_TMEM;   // pass-2 multipliers in tensor memory (full batches)
    if (kE2Tm) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(32)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tm_fence_before();
    }
    __syncthreads();
    uint32_t tm_e2 = 0;   // this lane's multipliers: its sub-partition's quarter of tensor memory, columns 0..31
    if (kE2Tm) {
        tm_fence_after();
        tm_e2 = S.tmem_base + ((uint32_t) ((warp & 3) * 32) << 16);
        tm_store_consts(tm_e2, S.e2 + lane * kE2Row);   // (both warps of a sub-partition write the same values)
        __syncwarp();
    }
    const bool rotate = (L.extern_only == 0);

    // -------------------- ciphertext warps ------------------------------------------
#ifndef TFHE_B200_ROLE_SWIZZLE
#define TFHE_B200_ROLE_SWIZZLE 0
#endif
    // sub-partition w % 4 holds the same role of two ciphertexts (SWIZZLE: the two roles, experiment)
    const int ct = warp >> 1, role = TFHE_B200_ROLE_SWIZZLE ? ((warp & 1) ^ ((warp >> 2) & 1)) : (warp & 1);
    // small batches: slots that never hold a ciphertext and are not the helper slot have nothing to do
    if (HELPER && ct >= cpg && ct != kCtWarps - 1) return;
    const unsigned int ncons = HELPER ? (unsigned int) cpg + 1u : (unsigned int) kCtWarps;
    WarpSmem &W = S.w[ct];
    // Named barriers of the pair: X = the hand-over of the partial sums (the one blocking barrier of an
    // iteration); Y[r] = "the partner has read role r's buffer B" (arrive by the reader, sync by the
    // owner before it overwrites B: a whole phase later, so the sync practically never waits).
    const int bar_x = 1 + 3 * ct, bar_mine = 2 + 3 * ct + role, bar_partner = 2 + 3 * ct + (1 - role);
    auto pair_sync = [bar_x]() { named_sync(bar_x, 64); };
    const uint32_t ring_base = (uint32_t) role * kRingStages;
    const uint32_t ring_chunks = kChunksPerIter * iters_total;
    StreamPos sp;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        int g = grp * cpg + ct;
        const bool valid = ct < cpg && g < L.total;
        if (!valid) g = L.total - 1;
        const GateIn I = resolve_inputs(L, g);

        if (L.acc_in != nullptr || L.testvect != nullptr) {
            if (role == 0) {
                if (L.acc_in != nullptr) {
                    phase_load_acc(lane, W, L.acc_in + (size_t) g * (kK + 1) * kN);
                } else {
                    // ACC = (0, X^{2N-barb} * testvect)
                    const int barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
                    for (int j = lane; j < kN; j += 32) {
                        const int s = (j + barb) & (2 * kN - 1);
                        const uint32_t v = (uint32_t) __ldg(L.testvect + (s & (kN - 1)));
                        W.acc[0][(j & 15) * kAccRow + (j >> 4)] = 0;
                        W.acc[kK][(j & 15) * kAccRow + (j >> 4)] = (int32_t) (s < kN ? v : 0u - v);
                    }
                }
            }
            pair_sync();
        } else {
            int barb;  // each role initialises its own polynomial
            if (L.explicit_inputs != 0) barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
            else {
                barb = modswitch_2N(prologue_word(I, L.n, I.cst));
            }
            phase_init(lane, W, role, barb, L.mu);
        }
        __syncwarp();
        phase_ext_build(lane, W, role);
        __syncwarp();
        named_arrive(bar_partner, 64);  // prime: nobody is reading the partner's buffer B

        int a_blk = 0;  // lane l holds bara of iteration (it & ~31) + l
        for (int it = 0; it < n_iter; it++) {
            if ((it & 31) == 0) a_blk = load_bara(L, I, g, it + lane, n_iter, rotate);
            const int a = __shfl_sync(0xffffffffu, a_blk, it & 31);
            // tfhe_blindRotate_FFT :705 skips barai == 0 (no-op); idle slots do no arithmetic either
            const bool active = valid && ((a != 0) || !rotate);

            PHASE_T0();
            if (!active) {
                // nothing to compute (bara = 0, or an idle slot of a small batch): only keep this
                // warp's place in the key stream
#pragma unroll 1
                for (uint32_t c = 0; c < kChunksPerIter; c++)
                    ring_skip<HELPER>(S, L, role, lane, ct == kCtWarps - 1, sp, ring_base, ring_chunks, ncons);
                continue;
            }
            // ---- one MuxRotate step: straight-line code, no per-phase conditionals -----------
            {
                cpx x[32];
                phase_f1_decomp(lane, W, role, a, rotate, x);
                phase_f1_fft(x);
                __syncwarp();  // the stores overwrite the extended copy the whole warp has just read
                if (!rotate) phase_acc_clear(lane, W, role);  // external product only: result replaces ACC
                named_sync(bar_mine, 64);  // the partner has finished with last iteration's give in B
                phase_f1_store(lane, W, role, x);
                __syncwarp();  // a warp multiplies exactly the rows it has just transformed
            }
            PHASE_MARK(0);
            // keep / give: partial sums of the result polynomial this warp finishes / hands over
            cpx keep[16], give[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                keep[i].x = 0.0; keep[i].y = 0.0;
                give[i].x = 0.0; give[i].y = 0.0;
            }
#pragma unroll 1
            for (int row = 2 * role; row < 2 * role + 2; row++) {
                cpx z[16];
                const bool rdy_keep = ring_probe(S, sp, ring_base, 0);
                const bool rdy_give = ring_probe(S, sp, ring_base, 1);
                if (kE2Tm) {
                    TmWords4 w0;
                    tm_ld4(tm_e2, w0);
                    const cpx *src = W.exch[row] + lane * kExchRow;
#pragma unroll
                    for (int j2 = 0; j2 < 16; j2++) z[j2] = src[j2];
                    fwd16_tm(z, tm_e2, w0);
                } else {
                    phase_f2_fft(lane, W, S.e2, row, z);
                }
                PHASE_MARK(1);
                mac_consume<HELPER>(S, L, role, lane, sp, ring_base, ring_chunks, rdy_keep, z, keep);
                PHASE_MARK(2);
                mac_consume<HELPER>(S, L, role, lane, sp, ring_base, ring_chunks, rdy_give, z, give);
                PHASE_MARK(3);
            }
            __syncwarp();  // every lane has read its pass-1 output in B
            phase_xchg_store(lane, W, role, give);
            PHASE_MARK(4);
            pair_sync();
            PHASE_MARK(5);
            TmWords16 w3;
            if (kE2Tm) tm_ld16w(tm_e2 + 16, w3);
            phase_xchg_load(lane, W, role, keep);
            named_arrive(bar_partner, 64);  // done with the partner's B
            if (kE2Tm) {
                inv16_tm(keep, tm_e2, w3);
                cpx *d = W.exch[2 * role] + lane * kExchRow;
#pragma unroll
                for (int j2 = 0; j2 < 16; j2++) d[j2] = keep[j2];
            } else {
                phase_inv16_store(lane, W, S.e2, role, keep);
            }
            __syncwarp();
            PHASE_MARK(6);
            {
                cpx x[16], send[8], recv[8];
                phase_i2_inner(lane, W, role, x);
                PHASE_MARK(7);
                phase_i2_send(lane, x, send);
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    recv[b].x = __shfl_xor_sync(0xffffffffu, send[b].x, 16);
                    recv[b].y = __shfl_xor_sync(0xffffffffu, send[b].y, 16);
                }
                PHASE_MARK(8);
                __syncwarp();  // all reads of A (inverse pass-2 output) precede the extended-copy stores
                phase_i2_final(lane, W, role, x, recv);
            }
            __syncwarp();
            PHASE_MARK(9);
        }
        named_sync(bar_mine, 64);  // drain the last arrive (keeps arrive / sync balanced)
        pair_sync();               // the partner's polynomial is final
        if (valid && role == 0) {
            if (L.u_out != nullptr) phase_extract(lane, W, L.u_out + (size_t) g * (kN + 1));
            if (L.acc_out != nullptr) phase_dump_acc(lane, W, L.acc_out + (size_t) g * (kK + 1) * kN);
        }
        pair_sync();
    }
    if (kE2Tm) {
        tm_fence_before();
        __syncthreads();
        if (warp == 0) {
            tm_fence_after();
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(S.tmem_base), "n"(32) : "memory");
        }
    }
}

// =================================================================================
// LATENCY kernel: one ciphertext per CTA, eight warps (batches of at most one ciphertext per SM: a single
// gate, the narrow levels of an adder or multiplier).
//
// The throughput kernel gives a ciphertext two warps because eight 255-register warps fill an SM with
// four ciphertexts.  A lone ciphertext on an SM leaves six warps idle, and its iteration is a chain of
// dependent phases; a lone warp on a sub-partition issues an fp64 instruction every ~3.4 cycles and ~0.45
// instructions per cycle over all, so the length of the per-lane instruction stream is what counts.  Here:
//   - warp (o, q) (warps 0..3) decomposes accumulator polynomial o at digit level q only and runs pass 1 of
//     that ONE row, split over lane pairs (br_core.cuh phase_f1h_*);
//   - the key of an iteration (64 KiB) is double buffered whole in shared memory: one thread of warp 4 (idle
//     until the next barrier) re-arms the buffer of iteration it for iteration it + 2 right behind the barrier
//     that ends the Fourier section (one bulk copy per iteration; no producer warps, counters or polling);
//   - the Fourier section: pass 2 by warp cq (class octet cq, all four rows, in place), multiply (by position
//     pairs, both result polynomials per lane) and inverse pass 2 by the warps cq (position half 0) and
//     cq + 4 (half 1) (br_core.cuh "Fourier section");
//   - the inverse pass 1 + conversion + accumulator update of result polynomial o is shared by the warps
//     (o, 0) and (o, 1) by halves of the slices (8 positions per lane, the last two stages through lane ^ 8
//     and lane ^ 16); warps 4..7 wait at the next barrier during pass 1 and this part.
// Two CTA barriers and three 64-thread barriers per iteration.  History (single gate, blind rotation):
// two-warp kernel 2.30 ms; four warps by decomposed row with partial sums in shared memory and four more
// warps feeding per-row key rings 1.60 ms; sums over rows in registers 1.55 ms; eight warps with the key
// double buffered 1.50 ms; pass 2 without redundant reads 1.44 ms; multiply without redundant reads 1.42 ms
// (profiles/README.md).
// Results are the same Torus32 words as the throughput kernel's: both return the exact integer product
// (the fp64 sums are taken in a different order, far inside the rounding margin; tests/test_gpu_parity.py
// compares the two kernels word for word).
struct __align__(128) OctoCtaSmem {
    LatencySmem w;
    cpx e2[32 * kE2Row];
    cpx key[2][kBkIterCplx];
    unsigned long long full[2];
};
static_assert(offsetof(OctoCtaSmem, key) % 128 == 0, "TMA destination alignment");
static_assert(sizeof(OctoCtaSmem) <= 227 * 1024, "shared memory budget");

// key of iteration `it` into buffer it & 1 (one 64 KiB bulk copy)
__device__ __forceinline__ void octo_fill(OctoCtaSmem &S, const BrLaunch &L, int it) {
    unsigned long long *bar = &S.full[it & 1];
    const cpx *src = L.bk + (size_t) (L.bk_first + it) * kBkIterCplx;
    mbar_arrive_expect_tx(bar, (uint32_t) (kBkIterCplx * sizeof(cpx)));
    tma_load_1d(S.key[it & 1], src, (uint32_t) (kBkIterCplx * sizeof(cpx)), bar);
}

__global__ void __launch_bounds__(kThreads, 1) blind_rotate_octo_kernel(const BrLaunch L) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OctoCtaSmem &S = *reinterpret_cast<OctoCtaSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = L.n_iter;
    const int g = blockIdx.x;  // one ciphertext per CTA
    const bool rotate = (L.extern_only == 0);

    build_e2(S.e2);
    if (threadIdx.x == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int it = 0; it < 2 && it < n_iter; it++) octo_fill(S, L, it);
    }
    __syncthreads();

    LatencySmem &W = S.w;
    const int tid = threadIdx.x;
    auto sync_all = [](int id) { named_sync(id, kThreads); };
    // pass 1 / inverse pass 1 (warps 0..3): decomposed row r = 2o + q
    const int o = (warp >> 1) & 1, q = warp & 1, r = warp & 3;
    const bool edge = warp < 4;
    // Fourier section: class octet cq, position half ph; lane (rr, c) / (g4, c) / (kk, oo, c)
    const int cq = warp & 3, ph = warp >> 2;   // the octet's warps (cq, cq + 4) sit on the same sub-partition
    const int rr = lane >> 3, m1 = 8 * cq + (lane & 7);
    const int oo = (lane >> 3) & 1, kb4 = 2 * ph + (lane >> 4);   // inverse stages 1, 0: positions {kb4 + 4 m} of polynomial oo
    const cpx ig1 = S.e2[m1 * kE2Row + 1], ig0 = S.e2[m1 * kE2Row];
    const int g4 = lane >> 3, p0 = 8 * ph + 2 * g4;   // multiply: positions p0, p0 + 1
    cpx pc3, pk2;
    phase_p_inv_consts(ph, g4, S.e2 + m1 * kE2Row, pc3, pk2);

    const GateIn I = resolve_inputs(L, g);
    // ---- accumulator initialisation
    if (L.acc_in != nullptr) {
        phase_load_acc_p(tid, kThreads, W.acc, L.acc_in + (size_t) g * (kK + 1) * kN);
    } else if (L.testvect != nullptr) {
        const int barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
        for (int j = tid; j < kN; j += kThreads) {
            const int s = (j + barb) & (2 * kN - 1);
            const uint32_t v = (uint32_t) __ldg(L.testvect + (s & (kN - 1)));
            W.acc[0][(j & 15) * kAccRow + (j >> 4)] = 0;
            W.acc[kK][(j & 15) * kAccRow + (j >> 4)] = (int32_t) (s < kN ? v : 0u - v);
        }
    } else if (edge && q == 0) {
        int barb;
        if (L.explicit_inputs != 0) barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
        else barb = modswitch_2N(prologue_word(I, L.n, I.cst));
        phase_init_p(lane, W.acc[o], o, barb, L.mu);
    }
    sync_all(1);
    if (edge && q == 0) phase_ext_build_p(lane, W.acc[o], W.ext[o]);
    sync_all(2);

    int a_blk = 0;  // lane l holds bara of iteration (it & ~31) + l
    for (int it = 0; it < n_iter; it++) {
        if ((it & 31) == 0) a_blk = load_bara(L, I, g, it + lane, n_iter, rotate);
        const int a = __shfl_sync(0xffffffffu, a_blk, it & 31);
        unsigned long long *full = &S.full[it & 1];
        const uint32_t parity = (uint32_t) (it >> 1) & 1u;
        if (a == 0 && rotate) {  // tfhe_blindRotate_FFT :705: nothing to do; the key buffer moves on
            if (threadIdx.x == 4 * 32) {
                mbar_wait(full, parity);
                if (it + 2 < n_iter) octo_fill(S, L, it + 2);
            }
            continue;
        }
        PHASE_T0();
        if (edge) {
            // pass 1 of row (o, q) split over lane pairs (br_core.cuh phase_f1h_*)
            const int hh = lane >> 4, j2 = lane & 15;
            cpx x[16], w[16], recv[16];
            phase_f1h_decomp_p(hh, j2, q, W.acc[o], W.ext[o], a, rotate, x);
            phase_f1h_cross_send(hh, x, w);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, w[i].x, 16);
                recv[i].y = __shfl_xor_sync(0xffffffffu, w[i].y, 16);
            }
            phase_f1h_finish(hh, w, recv, x);
            phase_f1h_store_p(hh, j2, W.exch[r], x);
        }
        PHASE_MARK(0);
        sync_all(1);  // the pass-1 output of all four rows is in place
        PHASE_MARK(1);
        // pass 2 of the octet's 32 (row, class) transforms by warp cq alone, in place (16 positions per lane; an
        // eight-warp version that produced 8 positions per lane from all 16 inputs read every input twice: 256
        // more shared-memory wavefronts, +3 % time); warp cq + 4 waits for it
        if (edge) phase_c_f2_inplace(rr, m1, W.exch, S.e2);
        named_sync(3 + cq, 64);
        PHASE_MARK(2);
        {
            // multiply by position pairs (br_core.cuh): lane (g4, c): positions p0, p0 + 1, both result polynomials
            cpx zr[kKpl][2], acc[kK + 1][2];
            phase_p_load_rows(p0, m1, W.exch, zr);
#pragma unroll
            for (int u = 0; u <= kK; u++) acc[u][0].x = 0.0, acc[u][0].y = 0.0, acc[u][1].x = 0.0, acc[u][1].y = 0.0;
            mbar_wait(full, parity);
            const cpx *kb = S.key[it & 1] + p0 * 32 + m1;
#pragma unroll
            for (int row = 0; row < kKpl; row++)
#pragma unroll
                for (int u = 0; u <= kK; u++)
#pragma unroll
                    for (int e = 0; e < 2; e++) cmac(acc[u][e], zr[row][e], kb[row * kBkRowCplx + u * kBkHalfCplx + e * 32]);
            PHASE_MARK(3);
#pragma unroll
            for (int u = 0; u <= kK; u++) {
                bf_inv(acc[u][0], acc[u][1], pc3.x, pc3.y);   // inverse stage 3
#pragma unroll
                for (int e = 0; e < 2; e++) {                  // inverse stage 2 through lane ^ 8
                    cpx recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, acc[u][e].x, 8);
                    recv.y = __shfl_xor_sync(0xffffffffu, acc[u][e].y, 8);
                    phase_p_inv_cross(g4 & 1, pk2, recv, acc[u][e]);
                }
                phase_p_inv_store(p0, m1, W.inv[u], acc[u]);
            }
        }
        PHASE_MARK(4);
        named_sync(3 + cq, 64);  // stages 3, 2 of both position halves of this class octet are in place
        PHASE_MARK(5);
        phase_o_inv_b_inplace(kb4, m1, W.inv[oo], ig1, ig0);
        PHASE_MARK(6);
        sync_all(2);  // the inverse pass-2 output is complete; nobody reads this iteration's key any more
        PHASE_MARK(7);
        // this iteration's key buffer is free: re-arm it for iteration it + 2 from a warp that is idle until
        // the next barrier 1 (on warp 0 the proxy fence and the bulk-copy issue were ~175 cycles of its
        // inverse pass 1: measured with the phase timers)
        if (threadIdx.x == 4 * 32 && it + 2 < n_iter) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            octo_fill(S, L, it + 2);
        }
        if (edge) {
            // inverse pass 1 + conversion + update of result polynomial o, shared by the warps (o, 0) and (o, 1)
            const int qq = lane >> 3;
            cpx x[8], recv[8];
            phase_q_i2_local(lane, q, W.inv[o], x);
            PHASE_MARK(8);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, x[i].x, 8);
                recv[i].y = __shfl_xor_sync(0xffffffffu, x[i].y, 8);
            }
            phase_q_i2_cross((qq & 1) != 0, 1 + (qq >> 1), recv, x);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, x[i].x, 16);
                recv[i].y = __shfl_xor_sync(0xffffffffu, x[i].y, 16);
            }
            phase_q_i2_cross((qq >> 1) != 0, 0, recv, x);
            if (!rotate) phase_q_acc_clear(lane, q, W.acc[o]);  // external product only: result replaces ACC
            phase_q_final(lane, q, W.acc[o], W.ext[o], x);
            PHASE_MARK(9);
            named_sync(7 + o, 64);  // polynomial o and its extended copy are final for its two warps
            PHASE_MARK(10);
        }
    }
    sync_all(1);  // both polynomials are final for everybody
    if (L.u_out != nullptr) phase_extract_p(tid, kThreads, W.acc, L.u_out + (size_t) g * (kN + 1));
    if (L.acc_out != nullptr) phase_dump_acc_p(tid, kThreads, W.acc, L.acc_out + (size_t) g * (kK + 1) * kN);
}

// =================================================================================
// LATENCY kernel for TWO ciphertexts per SM (batches of 149 .. 2 x #SMs gates): one ciphertext per CTA on FOUR
// warps, two CTAs resident per SM (109 KB of shared memory and 128 x <= 255 registers each).
// Same phases as the eight-warp kernel, except that (i) the whole Fourier section of a class octet is run by
// ONE warp (br_core.cuh "by ONE warp per class octet": measured 0.5 % slower than the eight-warp split for a
// lone ciphertext, and it needs no pair barriers), (ii) the key of an iteration has ONE buffer, re-armed by
// thread 0 right behind the barrier that ends the Fourier section (the copy has the inverse pass 1, pass 1 and
// pass 2 of the next iteration to land), (iii) the inverse pass-2 output and the extended accumulator copies
// alias the exchange buffers (QuadSmem), which costs one more CTA barrier per iteration (between the
// decomposition reads and the pass-1 stores).  Two such CTAs put two warps on every sub-partition: 296 gates
// in ~2.0 ms instead of 2.62 ms on the two-warp kernel.
struct __align__(128) QuadCtaSmem {
    QuadSmem w;
    cpx e2[32 * kE2Row];
    cpx key[kBkIterCplx];
    unsigned long long full;
};
static_assert(offsetof(QuadCtaSmem, key) % 128 == 0, "TMA destination alignment");
static_assert(2 * (sizeof(QuadCtaSmem) + 1024) <= 227 * 1024, "two CTAs per SM");
constexpr int kQuadThreads = 128;

__device__ __forceinline__ void quad_fill(QuadCtaSmem &S, const BrLaunch &L, int it) {
    const cpx *src = L.bk + (size_t) (L.bk_first + it) * kBkIterCplx;
    mbar_arrive_expect_tx(&S.full, (uint32_t) (kBkIterCplx * sizeof(cpx)));
    tma_load_1d(S.key, src, (uint32_t) (kBkIterCplx * sizeof(cpx)), &S.full);
}

__global__ void __launch_bounds__(kQuadThreads, 2) blind_rotate_quad_kernel(const BrLaunch L) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    QuadCtaSmem &S = *reinterpret_cast<QuadCtaSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = L.n_iter;
    const int g = blockIdx.x;  // one ciphertext per CTA
    const bool rotate = (L.extern_only == 0);

    build_e2(S.e2);
    if (threadIdx.x == 0) {
        mbar_init(&S.full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (n_iter > 0) quad_fill(S, L, 0);
    }
    __syncthreads();

    QuadSmem &W = S.w;
    const int tid = threadIdx.x;
    auto sync_all = [](int id) { named_sync(id, kQuadThreads); };
    const int o = warp >> 1, q = warp & 1, r = warp;                   // pass 1 / inverse pass 1: row r = 2o + q
    int32_t *ext_o = reinterpret_cast<int32_t *>(W.exch[2 + o]);      // extended copy of polynomial o
    cpx *inv_o = W.exch[o];                                            // inverse pass-2 output of polynomial o
    const int rr = lane >> 3, m1 = 8 * warp + (lane & 7);              // Fourier section: class octet = warp
    cpx wc3, wc2;
    phase_w_inv_consts(rr, S.e2 + m1 * kE2Row, wc3, wc2);
    const cpx ig1 = S.e2[m1 * kE2Row + 1], ig0 = S.e2[m1 * kE2Row];

    const GateIn I = resolve_inputs(L, g);
    if (L.acc_in != nullptr) {
        phase_load_acc_p(tid, kQuadThreads, W.acc, L.acc_in + (size_t) g * (kK + 1) * kN);
    } else if (L.testvect != nullptr) {
        const int barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
        for (int j = tid; j < kN; j += kQuadThreads) {
            const int s = (j + barb) & (2 * kN - 1);
            const uint32_t v = (uint32_t) __ldg(L.testvect + (s & (kN - 1)));
            W.acc[0][(j & 15) * kAccRow + (j >> 4)] = 0;
            W.acc[kK][(j & 15) * kAccRow + (j >> 4)] = (int32_t) (s < kN ? v : 0u - v);
        }
    } else if (q == 0) {
        int barb;
        if (L.explicit_inputs != 0) barb = L.barb ? (__ldg(L.barb + g) & (2 * kN - 1)) : 0;
        else barb = modswitch_2N(prologue_word(I, L.n, I.cst));
        phase_init_p(lane, W.acc[o], o, barb, L.mu);
    }
    sync_all(1);
    if (q == 0) phase_ext_build_p(lane, W.acc[o], ext_o);
    sync_all(2);

    int a_blk = 0;  // lane l holds bara of iteration (it & ~31) + l
    for (int it = 0; it < n_iter; it++) {
        if ((it & 31) == 0) a_blk = load_bara(L, I, g, it + lane, n_iter, rotate);
        const int a = __shfl_sync(0xffffffffu, a_blk, it & 31);
        const uint32_t parity = (uint32_t) it & 1u;
        if (a == 0 && rotate) {  // tfhe_blindRotate_FFT :705: nothing to do; the key buffer moves on
            if (threadIdx.x == 0) {
                mbar_wait(&S.full, parity);
                if (it + 1 < n_iter) quad_fill(S, L, it + 1);
            }
            continue;
        }
        {
            // pass 1 of row (o, q) split over lane pairs (br_core.cuh phase_f1h_*)
            const int hh = lane >> 4, j2 = lane & 15;
            cpx x[16], w[16], recv[16];
            phase_f1h_decomp_p<true>(hh, j2, q, W.acc[o], ext_o, a, rotate, x);  // two CTAs per SM: the fp64 pipe binds
            sync_all(3);  // every warp has read the extended copies: their buffers take the pass-1 output of rows 2, 3
            phase_f1h_cross_send(hh, x, w);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, w[i].x, 16);
                recv[i].y = __shfl_xor_sync(0xffffffffu, w[i].y, 16);
            }
            phase_f1h_finish(hh, w, recv, x);
            phase_f1h_store_p(hh, j2, W.exch[r], x);
        }
        sync_all(1);  // the pass-1 output of all four rows is in place
        {
            // the whole Fourier section of class octet `warp` (nothing leaves the warp)
            phase_c_f2_inplace(rr, m1, W.exch, S.e2);
            __syncwarp();
            cpx acc2[kK + 1][4];
#pragma unroll
            for (int u = 0; u <= kK; u++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc2[u][i].x = 0.0, acc2[u][i].y = 0.0;
            {
                cpx zr[kKpl][4];
                phase_w_load_rows(rr, m1, W.exch, zr);
                mbar_wait(&S.full, parity);
                const cpx *kb = S.key + (4 * rr) * 32 + m1;
#pragma unroll
                for (int u = 0; u <= kK; u++)
#pragma unroll
                    for (int row = 0; row < kKpl; row++)
#pragma unroll
                        for (int i = 0; i < 4; i++) cmac(acc2[u][i], zr[row][i], kb[row * kBkRowCplx + u * kBkHalfCplx + i * 32]);
            }
            __syncwarp();  // every lane has read its rows: rows 0, 1 of this octet take the inverse pass-2 values
#pragma unroll
            for (int u = 0; u <= kK; u++) phase_w_inv_a_store(rr, m1, W.exch[u], wc3, wc2, acc2[u]);
            __syncwarp();
            phase_w_inv_b_inplace(rr & 1, m1, W.exch[rr >> 1], ig1, ig0);
        }
        sync_all(2);  // the inverse pass-2 output is complete; nobody reads this iteration's key any more
        if (threadIdx.x == 0 && it + 1 < n_iter) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            quad_fill(S, L, it + 1);
        }
        {
            // inverse pass 1 + conversion + update of result polynomial o, shared by the warps (o, 0) and (o, 1)
            const int qq = lane >> 3;
            cpx x[8], recv[8];
            phase_q_i2_local(lane, q, inv_o, x);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, x[i].x, 8);
                recv[i].y = __shfl_xor_sync(0xffffffffu, x[i].y, 8);
            }
            phase_q_i2_cross((qq & 1) != 0, 1 + (qq >> 1), recv, x);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                recv[i].x = __shfl_xor_sync(0xffffffffu, x[i].x, 16);
                recv[i].y = __shfl_xor_sync(0xffffffffu, x[i].y, 16);
            }
            phase_q_i2_cross((qq >> 1) != 0, 0, recv, x);
            if (!rotate) phase_q_acc_clear(lane, q, W.acc[o]);  // external product only: result replaces ACC
            phase_q_final(lane, q, W.acc[o], ext_o, x);
            named_sync(4 + o, 64);  // polynomial o and its extended copy are final for its two warps
        }
    }
    sync_all(1);  // both polynomials are final for everybody
    if (L.u_out != nullptr) phase_extract_p(tid, kQuadThreads, W.acc, L.u_out + (size_t) g * (kN + 1));
    if (L.acc_out != nullptr) phase_dump_acc_p(tid, kQuadThreads, W.acc, L.acc_out + (size_t) g * (kK + 1) * kN);
}

// ------------------------------------------------------------ key conversion

struct __align__(128) FwdSmem {
    WarpSmem w[kCtWarps];
    cpx e2[32 * kE2Row];
};

__global__ void __launch_bounds__(kCtWarps * 32, 1)
forward_polys_kernel(const int32_t *__restrict__ coef, cpx *__restrict__ out, int ngroups, double scale) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem &S = *reinterpret_cast<FwdSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    build_e2(S.e2);
    __syncthreads();
    WarpSmem &W = S.w[warp];
    for (int grp = blockIdx.x * kCtWarps + warp; grp < ngroups; grp += gridDim.x * kCtWarps) {
        const int32_t *src = coef + (size_t) grp * 4 * kN;
        fwd4_pass1(lane, W, [&](int p, int j) { return (double) __ldg(src + p * kN + j) * scale; });
        __syncwarp();
        fwd4_pass2(lane, W, S.e2, out + (size_t) grp * 4 * kM);
        __syncwarp();
    }
}

}  // namespace

size_t blind_rotate_smem_bytes() { return sizeof(CtaSmem); }

// 0: the fp64 products are rounded to the nearest integer (default: the accumulator equals the
// exact negacyclic product); 1: built with -DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1, the conversion
// truncates like Torus32(int64_t(x)) of fft_processor_fftw.cu:177 (libtfhe_b200_trunc.so)
extern "C" int tfhe_b200_conversion_mode(void) { return TFHE_B200_TRUNCATE_LIKE_REFERENCE ? 1 : 0; }

#if TFHE_B200_PHASE_TIMING
extern "C" int tfhe_b200_debug_phase_cycles(long long *out, int reset) {
    if (cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles)) != cudaSuccess) return 1;
    if (reset) {
        long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

cudaError_t blind_rotate_configure() {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) sizeof(CtaSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) sizeof(CtaSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_octo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) sizeof(OctoCtaSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) sizeof(QuadCtaSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(forward_polys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int) sizeof(FwdSmem));
}

cudaError_t launch_blind_rotate(const BrLaunch &L_in, int sm_count, cudaStream_t stream) {
    if (L_in.total <= 0) return cudaSuccess;
    BrLaunch L = L_in;
    int cpg = (L.total + sm_count - 1) / sm_count;
    L.cts_per_group = cpg < 1 ? 1 : (cpg > kCtWarps ? kCtWarps : cpg);
    const int ngroups = (L.total + L.cts_per_group - 1) / L.cts_per_group;
    const int grid = ngroups < sm_count ? ngroups : sm_count;
    // at most one ciphertext per SM: the latency kernel (one ciphertext per CTA, eight warps).
    // TFHE_B200_BR_LATENCY=0 keeps such batches on the two-warp kernel (A/B measurements)
    static const bool use_latency = [] {
        const char *v = getenv("TFHE_B200_BR_LATENCY");
        return v == nullptr || atoi(v) != 0;
    }();
    if (use_latency && L.total <= sm_count) {
        blind_rotate_octo_kernel<<<L.total, kThreads, sizeof(OctoCtaSmem), stream>>>(L);
        return cudaGetLastError();
    }
    // at most two ciphertexts per SM: the four-warp latency kernel, two CTAs per SM
    if (use_latency && L.total <= 2 * sm_count) {
        blind_rotate_quad_kernel<<<L.total, kQuadThreads, sizeof(QuadCtaSmem), stream>>>(L);
        return cudaGetLastError();
    }
    // small batches: the always-idle last slot of every CTA refills the key rings (ring_skip<true>)
    if (L.cts_per_group < kCtWarps) blind_rotate_kernel<true><<<grid, kThreads, sizeof(CtaSmem), stream>>>(L);
    else blind_rotate_kernel<false><<<grid, kThreads, sizeof(CtaSmem), stream>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_forward_polys(const int32_t *coef, cpx *out, int npolys, double scale, cudaStream_t stream) {
    if (npolys <= 0) return cudaSuccess;
    if (npolys % 4 != 0) return cudaErrorInvalidValue;
    const int ngroups = npolys / 4;
    int grid = (ngroups + kCtWarps - 1) / kCtWarps;
    if (grid > 4 * 148) grid = 4 * 148;
    forward_polys_kernel<<<grid, kCtWarps * 32, sizeof(FwdSmem), stream>>>(coef, out, ngroups, scale);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
