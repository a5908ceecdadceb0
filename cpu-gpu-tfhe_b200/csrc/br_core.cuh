// Blind-rotation core: per-lane phase functions of the warp-per-ciphertext
// persistent kernel (blind_rotate.cu).  Everything here is __host__ __device__
// so that tests/host_emul.cu can run the exact same index math lane by lane on
// the CPU (there is no GPU in the build container).
//
// A pair of warps owns one ciphertext for all n iterations of
//   ACC <- ACC + BK_i (.) ((X^{a_i} - 1) * ACC)
// (reference: tfhe_MuxRotate_FFT, lwe-bootstrapping-functions-fft.cu:105-185;
//  tGswFFTExternMulToTLwe, tgsw-fft-operations.cu:124-264).
//
// Number representation.  A real polynomial P mod X^N+1 (N = 1024) is held in
// the Fourier domain as the M = N/2 = 512 values
//     V_m = P(zeta^(4m+1)),  zeta = exp(i*pi/N),  m = 0..511
//         = sum_{j<512} (p_j + i*p_{j+512}) * exp(2*pi*i * j*(m + 1/4)/512)
// i.e. the reference's "lagrangehalfc" values (fft_processor_fftw.cu:148-181,
// lagrangehalfc_impl.h:45-52) at the conjugate half of the odd roots and in a
// different order.  Any consistent pair of transforms gives the same negacyclic
// product; the key is converted with the same forward transform, so ordering
// never needs to be undone.
//
// Transform.  9 radix-2 stages with the 1/4 frequency shift carried through the
// recursion (a size-L block with shift d: u = a + e*b -> shift d/2,
// v = a - e*b -> shift (d+1)/2, e = exp(i*pi*d)), so there is no separate twist
// or twiddle pass and every butterfly is 6 FMA-class fp64 instructions.
// Index split j = 16*j1 + j2, m = m1 + 32*m2:
//   pass 1 (stages 0-4): lane (o, j2) holds the 32 elements j1 = 0..31 of
//           polynomial o's slice j2; multipliers are compile-time constants.
//   exchange through shared memory (rows of 17 complex -> conflict free).
//   pass 2 (stages 5-8): lane m1 holds the 16 elements j2 = 0..15 of frequency
//           class m1; multipliers depend on the lane (table e2).
// The inverse is the exact conjugate transpose (pass 2 then pass 1 reversed),
// unnormalised (factor 512 is folded into the key's scale 2^-9).
#pragma once

#include <stdint.h>

#include "fft_consts.h"

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define TFHE_HD __host__ __device__ __forceinline__

namespace tfhe_b200 {

constexpr int kN = 1024;          // ring degree
constexpr int kM = 512;           // complex points per polynomial
constexpr int kK = 1;             // TLWE mask polynomials
constexpr int kL = 2;             // gadget length
constexpr int kBgbit = 10;        // log2 gadget base
constexpr int kKpl = (kK + 1) * kL;
// offset = Bg/2 * sum_i 2^(32-(i+1)*Bgbit) = 512*(2^22+2^12)   (tgsw.cu:19-27)
constexpr uint32_t kDecompOffset = 0x80200000u;
constexpr int kExchRow = 17;      // complex per exchange row (16 used)
constexpr int kExchPoly = 32 * kExchRow;
constexpr int kAccRow = 68;       // words per accumulator row (64 used): 16 B aligned rows whose 128-bit
                                  // accesses by 8 lanes (rows j2..j2+7) cover the 32 banks exactly once
constexpr int kAccPoly = 16 * kAccRow;
constexpr int kE2Row = 5;         // complex per pass-2 constant row (4 used; 80 B stride is conflict free)
constexpr int kBkHalfCplx = 16 * 32;                // one result polynomial of a TGSW row: [pos][m1]
constexpr int kBkRowCplx = 2 * kBkHalfCplx;         // one TGSW row: [o][pos][m1]
constexpr int kBkIterCplx = kKpl * kBkRowCplx;     // one BK_i: 4096 complex = 64 KiB

struct cpx {
    double x, y;
};

// Extended accumulator copy (see phase_f1_decomp): row of 127 words per (o, j2),
//   E[63 + k] = +row[k] (k = 0..63),  E[63 + k] = -row[64 + k] (k = -63..-1)
// i.e. the negacyclic continuation of the 64-coefficient row to the left, so that a rotated read
// is `base + e` with a per-lane base and NO per-element index or sign arithmetic.
// Rows are 127 words (odd: the 16 rows of a polynomial start in 16 different banks, so the
// single-word rotated reads are conflict free inside a polynomial; 16 B aligned rows, which
// 128-bit stores would need, put 32 single-word readers on 8 banks: measured slower).
constexpr int kExtRow = 127;
constexpr int kExtOrg = 63;       // word of row[0]; -row[i] is at word i - 1 (row[0] has no left image)
// Bank placement of the 16 rows of a polynomial.  With consecutive rows (GAP = 0) row j2 starts in bank
// -j2: the rotated single-word reads of 16 rows are conflict free, but the final stage stores, in ONE
// instruction, words w (lanes (0, j2)) and w + 8 (lanes (1, j2)) of all 16 rows: the two 16-bank windows
// overlap by 8 banks (2-way conflict on every one of the 127 stores: 19 % of all store wavefronts of the
// round-2 kernel, ncu l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st).  GAP = 24 words in front of
// row 8 puts rows 0..7 in banks {0, 31 .. 25} and rows 8..15 in banks {16 .. 9}: that set and the same
// set + 8 are disjoint, and the rotated reads stay conflict free (the lanes whose rotation wraps read one
// word lower: still distinct banks).  The latency kernel reads words w and w + 16 of a row in one
// instruction (lane-pair split of pass 1) and keeps GAP = 0.
#ifndef TFHE_B200_EXT_GAP
#define TFHE_B200_EXT_GAP 24
#endif
constexpr int kExtGapThroughput = TFHE_B200_EXT_GAP;
template <int GAP>
TFHE_HD int ext_row_off(int j2) { return j2 * kExtRow + GAP * (j2 >> 3); }

struct alignas(16) word4 {
    int32_t v[4];
};

// Per-ciphertext shared-memory working set.
struct WarpSmem {
    union {
        cpx exch[kKpl][kExchPoly];          // 4 * 8704 B: transform exchange buffers
        int32_t exw[kKpl][kExchPoly * 4];   // the same as words: buffer 2o holds the extended copy of
                                            // accumulator polynomial o between iterations
    };
    int32_t acc[kK + 1][kAccPoly]; // 2 * 4160 B: the accumulator (master copy)
};
static_assert(16 * kExtRow + kExtGapThroughput <= kExchPoly * 4, "extended copy must fit an exchange buffer");

// Working set of the LATENCY kernel (one ciphertext per CTA, eight warps: blind_rotate.cu
// blind_rotate_octo_kernel): the extended accumulator copies have buffers of their own (both warps of
// a polynomial read them while the exchange buffers are being rewritten), and two more buffers take
// the inverse pass-2 output.
constexpr int kExtPolyWords = 2048;   // 16 rows of 127 words, padded
struct LatencySmem {
    cpx exch[kKpl][kExchPoly];          // row (o, q): pass-1 output of warp (o, q), read by all eight warps
    cpx inv[kK + 1][kExchPoly];         // polynomial o: inverse pass-2 output (written by all eight warps)
    int32_t acc[kK + 1][kAccPoly];
    int32_t ext[kK + 1][kExtPolyWords];
};
static_assert(16 * kExtRow <= kExtPolyWords, "extended copy must fit its buffer");

// Working set of the FOUR-warp latency kernel (two CTAs per SM, blind_rotate_quad_kernel): the inverse pass-2
// output of polynomial o lives in exchange buffer o, the extended copy of polynomial o in exchange buffer 2 + o
// (8 KiB of its 8.5 KiB); the kernel's barriers separate the life times.
struct QuadSmem {
    cpx exch[kKpl][kExchPoly];
    int32_t acc[kK + 1][kAccPoly];
};
static_assert(kExtPolyWords * 4 <= kExchPoly * 16, "extended copy must fit an exchange buffer");


TFHE_HD int32_t *ext_poly(WarpSmem &ws, int o) { return ws.exw[2 * o]; }

TFHE_HD constexpr int bitrev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// (a, b) <- (a + e*b, a - e*b)
TFHE_HD void bf_fwd(cpx &a, cpx &b, double er, double ei) {
    double ur = fma(er, b.x, a.x);
    double ui = fma(er, b.y, a.y);
    ur = fma(-ei, b.y, ur);
    ui = fma(ei, b.x, ui);
    b.x = fma(2.0, a.x, -ur);
    b.y = fma(2.0, a.y, -ui);
    a.x = ur;
    a.y = ui;
}

// (u, v) <- (u + v, conj(e)*(u - v))
TFHE_HD void bf_inv(cpx &a, cpx &b, double er, double ei) {
    const double tr = a.x - b.x, ti = a.y - b.y;
    a.x += b.x;
    a.y += b.y;
    b.x = fma(er, tr, ei * ti);
    b.y = fma(er, ti, -(ei * tr));
}

// Gadget digit in offset form (0..1023) -> double(digit - 512), two ways with the same value:
//   kI2F = true : an integer subtraction and I2F.F64.S32.  The conversion runs on its own pipe and leaves the
//                 fp64 pipe alone: the 64 conversions per warp and iteration are 3.4 % of the throughput kernel's
//                 fp64 instructions, and that kernel IS bound by the fp64 pipe (measured, round 2: 386.4 ->
//                 371.3 ms per 65536 gates, -3.9 %; two ciphertexts per SM: 2.32 -> 2.29 ms).
//   kI2F = false: the word 0x43300000:dig is the double 2^52 + dig, and one exact DADD subtracts 2^52 + 512.
//                 For a LONE warp on a sub-partition (eight-warp latency kernel) the fp64 pipe has room and the
//                 quarter-rate conversion pipe is the longer path: 1.418 ms per gate against 1.451 with I2F.
template <bool kI2F>
TFHE_HD double digit_to_double(uint32_t dig) {
#ifdef __CUDA_ARCH__
    if (kI2F) return __int2double_rn((int) dig - 512);
    return __hiloint2double(0x43300000, (int) dig) - 4503599627371008.0;  // 2^52 + 512
#else
    const uint64_t bits = (UINT64_C(0x43300000) << 32) | dig;
    double d;
    __builtin_memcpy(&d, &bits, sizeof(d));
    return d - 4503599627371008.0;
#endif
}

// double -> Torus32.  The reference does Torus32(int64_t(x)) (fft_processor_fftw.cu:177): truncation
// toward zero of a value that lies within fp64 rounding error (about 0.1) of the exact integer n
// of the negacyclic product, i.e. n or n -+ 1 at random.  Default here: round to nearest through the
// 2^52 + 2^51 magic constant (one DADD on the fp64 pipe; the low mantissa word is the two's
// complement result for |x| < 2^51), which returns n itself; F2I.S64.F64 runs on the quarter-rate
// XU pipe (8 cycles per warp instruction, measured) with a long latency in front of the
// accumulator update.  -DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1 restores the cast.
#ifndef TFHE_B200_TRUNCATE_LIKE_REFERENCE
#define TFHE_B200_TRUNCATE_LIKE_REFERENCE 0
#endif
#ifndef TFHE_B200_OUT_F2I
#define TFHE_B200_OUT_F2I 0
#endif
TFHE_HD uint32_t double_to_torus32(double x) {
#if defined(TFHE_B200_CONV_PROBE) && !defined(__CUDA_ARCH__)
    TFHE_B200_CONV_PROBE(x);  // host emulation only: tests record the distance of x from the integers
#endif
#if TFHE_B200_TRUNCATE_LIKE_REFERENCE
    return (uint32_t) (int32_t) (long long) x;
#else
#if defined(__CUDA_ARCH__) && TFHE_B200_OUT_F2I
    return (uint32_t) __double2ll_rn(x);   // F2I.S64.F64.RN on the conversion pipe: same result, -0.3 % measured
#endif
    const double y = x + 6755399441055744.0;  // 2^52 + 2^51
#ifdef __CUDA_ARCH__
    return (uint32_t) __double2loint(y);
#else
    uint64_t bits;
    __builtin_memcpy(&bits, &y, sizeof(bits));
    return (uint32_t) bits;
#endif
#endif
}

// acc += z * w
TFHE_HD void cmac(cpx &acc, const cpx &z, const cpx &w) {
    acc.x = fma(z.x, w.x, acc.x);
    acc.y = fma(z.x, w.y, acc.y);
    acc.x = fma(-z.y, w.y, acc.x);
    acc.y = fma(z.y, w.x, acc.y);
}

// Butterfly multipliers of pass 1 read from the __constant__ table: with a compile-time index
// they become constant-bank operands of the DFMAs (as immediates every one of them costs two
// UMOVs per use: 259 extra instructions per iteration in the first version).
TFHE_HD double c1_re_rt(int i) {
#ifdef __CUDA_ARCH__
    return d_c1_tab_re[i];
#else
    return h_c1_tab_re[i];
#endif
}

TFHE_HD double c1_im_rt(int i) {
#ifdef __CUDA_ARCH__
    return d_c1_tab_im[i];
#else
    return h_c1_tab_im[i];
#endif
}

// Stages 0-4 on the 32 in-register elements (natural j1 in, bit-reversed m1 out).
TFHE_HD void fwd32(cpx (&x)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int half = 16 >> s;
#pragma unroll
        for (int b = 0; b < (1 << s); b++) {
            const int ci = (1 << s) - 1 + b;
#pragma unroll
            for (int i = 0; i < half; i++)
                bf_fwd(x[b * 2 * half + i], x[b * 2 * half + i + half], c1_re_rt(ci), c1_im_rt(ci));
        }
    }
}

TFHE_HD void inv32(cpx (&x)[32]) {
#pragma unroll
    for (int s = 4; s >= 0; s--) {
        const int half = 16 >> s;
#pragma unroll
        for (int b = 0; b < (1 << s); b++) {
            const int ci = (1 << s) - 1 + b;
#pragma unroll
            for (int i = 0; i < half; i++)
                bf_inv(x[b * 2 * half + i], x[b * 2 * half + i + half], c1_re_rt(ci), c1_im_rt(ci));
        }
    }
}

// Stages 5-8 on the 16 in-register elements of frequency class m1; e points at
// this lane's 15 multipliers (natural j2 in, bit-reversed m2 out).
// Stage constants of pass 2.  Block b of stage s has shift (d0 + rev_s(b)) / 2^s with
// d0 = (m1 + 1/4)/32, so its multiplier is g_s * exp(i*pi*rev_s(b)/2^s) with the
// lane-dependent g_s = exp(i*pi*d0/2^s).  Only g_0..g_3 are kept per lane (table e, 4 complex
// instead of 15: shared-memory wavefronts were the busiest resource of the kernel); the other
// factors are exp(i*pi/2) = i (free), exp(i*pi/4), exp(i*pi/8), exp(3i*pi/8) (16 fp64
// instructions per transform).
struct Stage3Consts {
    cpx g, h4, h8, h38;  // g_3, g_3*e^{i pi/4}, g_3*e^{i pi/8}, g_3*e^{3 i pi/8}
};

TFHE_HD cpx cmul_const(const cpx &g, double cr, double ci) {
    cpx r;
    r.x = fma(g.x, cr, -(g.y * ci));
    r.y = fma(g.x, ci, g.y * cr);
    return r;
}

constexpr double kSqrtHalf = 0.70710678118654752440;
constexpr double kCosPi8 = 0.92387953251128675613, kSinPi8 = 0.38268343236508977173;

// constant of (stage s, block b) given the stage's base values; odd blocks are i * (even block)
template <int S, int B>
TFHE_HD cpx pass2_const(const cpx &g, const cpx &h4, const cpx &h8, const cpx &h38) {
    constexpr int base = B >> 1;
    cpx c = (S <= 1 || base == 0) ? g : (base == 1 ? h4 : (base == 2 ? h8 : h38));
    if (B & 1) {
        const double t = c.x;
        c.x = -c.y;
        c.y = t;
    }
    return c;
}

template <int S, bool kInv, int B = 0>
struct Pass2Stage {
    TFHE_HD static void run(cpx (&z)[16], const cpx &g, const cpx &h4, const cpx &h8, const cpx &h38) {
        constexpr int half = 8 >> S;
        const cpx c = pass2_const<S, B>(g, h4, h8, h38);
#pragma unroll
        for (int i = 0; i < half; i++) {
            if (kInv) bf_inv(z[B * 2 * half + i], z[B * 2 * half + i + half], c.x, c.y);
            else bf_fwd(z[B * 2 * half + i], z[B * 2 * half + i + half], c.x, c.y);
        }
        if constexpr (B + 1 < (1 << S)) Pass2Stage<S, kInv, B + 1>::run(z, g, h4, h8, h38);
    }
};

template <int S, bool kInv>
TFHE_HD void pass2_stage(cpx (&z)[16], const cpx *e) {
    const cpx g = e[S];
    cpx h4 = g, h8 = g, h38 = g;
    // (derived, not tabulated: a table of all eight multipliers per class — 4 more 128-bit shared-memory loads
    // per transform instead of 16 fp64 instructions — measured 3.5 % SLOWER: shared memory binds as well)
    if (S >= 2) {
        h4.x = (g.x - g.y) * kSqrtHalf;
        h4.y = (g.x + g.y) * kSqrtHalf;
    }
    if (S >= 3) {
        h8 = cmul_const(g, kCosPi8, kSinPi8);
        h38 = cmul_const(g, kSinPi8, kCosPi8);
    }
    Pass2Stage<S, kInv>::run(z, g, h4, h8, h38);
}

// Stages 5-8 on the 16 in-register elements of frequency class m1; e points at
// this lane's 4 base multipliers (natural j2 in, bit-reversed m2 out).
TFHE_HD void fwd16(cpx (&z)[16], const cpx *e) {
    pass2_stage<0, false>(z, e);
    pass2_stage<1, false>(z, e);
    pass2_stage<2, false>(z, e);
    pass2_stage<3, false>(z, e);
}

TFHE_HD void inv16(cpx (&z)[16], const cpx *e) {
    pass2_stage<3, true>(z, e);
    pass2_stage<2, true>(z, e);
    pass2_stage<1, true>(z, e);
    pass2_stage<0, true>(z, e);
}

// Shift of the lane-dependent base multiplier g_s of pass-2 stage s for frequency class m1.
TFHE_HD double e2_shift(int m1, int s) {
    double d = ((double) m1 + 0.25) / 32.0;
    for (int i = 0; i < s; i++) d *= 0.5;
    return d;
}

// ---------------------------------------------------------------- phases ---
//
// Two warps per ciphertext, split BY ACCUMULATOR POLYNOMIAL: warp `o` (o = 0: the mask polynomial a,
// o = 1: the body b) owns polynomial o of the accumulator, its master copy acc[o], its extended copy
// and the two exchange buffers exch[2o] ("A") and exch[2o+1] ("B").  One MuxRotate iteration of warp o:
//   decompose (X^a - 1) * ACC_o into BOTH digit levels (lane (q, j2): digit level q, slice j2) and run
//     pass 1 of the two forward transforms (decomposed rows 2o, 2o+1)          480 fp64 instr + 64 conversions
//   pass 2 + MAC of those rows against BK rows 2o, 2o+1 -> partial sums of result polynomial o
//     ("keep") and of result polynomial 1-o ("give")                            2 x 320
//   park `give` in B | PAIR BARRIER | keep += partner's give                   the ONLY pair barrier
//   inverse pass 2 of result polynomial o -> A; inverse pass 1 (lane (hh, j2): positions 16hh..16hh+15,
//     last stage through lane ^ 16 by warp shuffles); to Torus32; ACC_o += ...; extended copy rewritten
// Nothing else crosses the pair: a warp reads and writes only its own polynomial and buffers, so the
// five pair barriers of the first version (split by digit level, both warps touching both polynomials
// in the rotation and in the final stage) are one, plus a non-blocking arrive/sync pair that orders the
// partner's read of B before B is overwritten by the next iteration's pass-1 output.
// Buffer life times of warp o:  A: extended copy -> pass-1 output (q = 0) -> inverse pass-2 output ->
// extended copy;  B: pass-1 output (q = 1) -> give (read by the partner).

// ACC = (0, X^{2N-barb} * (mu, ..., mu))   tfhe_blindRotateAndExtract_FFT,
// lwe-bootstrapping-functions-fft.cu:1425-1431 ; torusPolynomialMulByXai :492-519
// Warp o initialises polynomial o: lane (h, j2) writes coefficients [32h, 32h+32) of row j2.
TFHE_HD void phase_init_p(int lane, int32_t *acc_o, int o, int barb, int32_t mu) {
    const int h = lane >> 4, j2 = lane & 15;
    int32_t *row = acc_o + j2 * kAccRow + 32 * h;
#pragma unroll 8
    for (int e = 0; e < 32; e++) {
        const int j = 16 * (e + 32 * h) + j2;
        int32_t v = 0;
        if (o == kK) v = (((j + barb) & (2 * kN - 1)) < kN) ? mu : (int32_t) (0u - (uint32_t) mu);
        row[e] = v;
    }
}

TFHE_HD void phase_init(int lane, WarpSmem &ws, int o, int barb, int32_t mu) { phase_init_p(lane, ws.acc[o], o, barb, mu); }

// Build the extended copy of accumulator polynomial o from its master copy (start of a ciphertext;
// afterwards phase_i2_final keeps it up to date).  Lane (h, j2) copies coefficients [32h, 32h+32) of row j2.
template <int GAP = 0>
TFHE_HD void phase_ext_build_p(int lane, const int32_t *acc_o, int32_t *ext_o) {
    const int h = lane >> 4, j2 = lane & 15;
    const int32_t *row = acc_o + j2 * kAccRow + 32 * h;
    int32_t *ext = ext_o + ext_row_off<GAP>(j2) + 32 * h;
#pragma unroll 2
    for (int b = 0; b < 32; b += 4) {
        const word4 v = *reinterpret_cast<const word4 *>(row + b);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            ext[kExtOrg + b + i] = v.v[i];
            if (b + i > 0 || h != 0) ext[b + i - 1] = (int32_t) (0u - (uint32_t) v.v[i]);
        }
    }
}

TFHE_HD void phase_ext_build(int lane, WarpSmem &ws, int o) {
    phase_ext_build_p<kExtGapThroughput>(lane, ws.acc[o], ext_poly(ws, o));
}

// Pass 1 of the two forward transforms of accumulator polynomial o (decomposed rows (o, q), q = 0..l-1),
// fused with the rotation (torusPolynomialMulByXaiMinusOne, toruspolynomial-functions.cu:191-213)
// and the gadget decomposition (tGswTorus32PolynomialDecompH, tgsw-functions.cu:301-352).
// Lane (q, j2): digit level q of slice j2; the two lanes of a slice read the same words (broadcast).
// rotate == false: plain decomposition of ACC (stand-alone external product).
//
// Coefficient j = 16 e + j2 of X^a * ACC is (-1)^f * ACC[16 (e - sh) + j2p] continued negacyclically,
// with a = 16 a_hi + a_lo, j2p = (j2 - a_lo) mod 16, sh = a_hi + [j2 < a_lo] = 64 f + h: in the
// extended row of (o, j2p) that is word kExtOrg - h + e, so the 64 rotated values of a lane are read
// at immediate offsets from ONE base (conflict free: the 16 rows of a polynomial start in 16 different
// banks and h differs by at most one between the lanes of a warp).  Sign, the subtraction of ACC, the
// decomposition offset and the digit shift are two integer multiply-adds (t * 2^(10 q) = v * (+-2^(10 q)) +
// (ACC * -2^(10 q) + offset * 2^(10 q)); digit = top 10 bits), on the FMA pipe: the first version
// spent 19 instructions per coefficient here, most of them on the half-rate ALU pipe.
// The outputs are NOT stored here: the extended copy lives in buffer A, which the stores overwrite
// (the caller separates the two with a __syncwarp).
// (pointer form: digit level q of slice j2 of the polynomial whose master / extended copy are given)
template <int GAP = 0>
TFHE_HD void phase_f1_decomp_p(int j2, int q, const int32_t *acc_o, const int32_t *ext_o, int a, bool rotate,
                               cpx (&x)[32]) {
    const int a_lo = a & 15, a_hi = a >> 4;
    const int j2p = (j2 - a_lo) & 15;
    const int sh = a_hi + (j2 < a_lo ? 1 : 0);
    const int h = sh & 63;
    const uint32_t m = 1u << (q * kBgbit);
    const uint32_t sm = rotate ? (((sh >> 6) & 1) ? 0u - m : m) : 0u;  // multiplier of the rotated value
    const uint32_t cm = rotate ? 0u - m : m;                            // multiplier of ACC itself
    const uint32_t offm = kDecompOffset * m;
    const int32_t *own = acc_o + j2 * kAccRow;
    const int32_t *rot = ext_o + ext_row_off<GAP>(j2p) + (kExtOrg - h);
#pragma unroll
    for (int blk = 0; blk < 64; blk += 16) {
        uint32_t vr[16], vo[16];
#pragma unroll
        for (int i = 0; i < 16; i++) vr[i] = (uint32_t) rot[blk + i];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const word4 w = *reinterpret_cast<const word4 *>(own + blk + i);
#pragma unroll
            for (int k = 0; k < 4; k++) vo[i + k] = (uint32_t) w.v[k];
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int e = blk + i;
            const uint32_t t = vr[i] * sm + (vo[i] * cm + offm);
            const double d = digit_to_double<true>(t >> (32 - kBgbit));
            if (e < 32) x[e & 31].x = d;
            else x[e & 31].y = d;
        }
    }
}

TFHE_HD void phase_f1_decomp(int lane, WarpSmem &ws, int o, int a, bool rotate, cpx (&x)[32]) {
    phase_f1_decomp_p<kExtGapThroughput>(lane & 15, lane >> 4, ws.acc[o], ext_poly(ws, o), a, rotate, x);
}

// Stages 0-4 of the two transforms on the decomposed digits (no shared-memory access).
TFHE_HD void phase_f1_fft(cpx (&x)[32]) { fwd32(x); }

TFHE_HD void phase_f1_store_p(int j2, cpx *buf, const cpx (&x)[32]) {
    cpx *dst = buf + j2;
#pragma unroll
    for (int pos = 0; pos < 32; pos++) dst[bitrev5(pos) * kExchRow] = x[pos];
}

TFHE_HD void phase_f1_store(int lane, WarpSmem &ws, int o, const cpx (&x)[32]) {
    phase_f1_store_p(lane & 15, ws.exch[o * kL + (lane >> 4)], x);
}

// ---- pass 1 split over lane PAIRS (latency kernel) --------------------------------------------
// A decomposed row has 16 slices of 32 points, so a warp that transforms ONE row keeps half its lanes
// idle if a lane takes a whole slice.  Here lane (hh, j2) takes the points j1 = 16 hh .. 16 hh + 15 of
// slice j2: stage 0 pairs point j1 of lane (0, j2) with point j1 + 16 of lane (1, j2) — one exchange
// through lane ^ 16 — and stages 1-4 are local (the multipliers of block hh: lane dependent).
// 288 instead of 480 fp64 instructions per lane, 32 instead of 64 digit conversions.
// The sums are taken in a different order than in fwd32, so the Fourier values differ in the last
// bits from the throughput kernel's; the accumulator words do not (both are the exact product).

// decomposition of the 32 coefficients lane (hh, j2) needs: e = 16 hh + i (real) and 32 + 16 hh + i (imag)
template <bool kI2F = false>
TFHE_HD void phase_f1h_decomp_p(int hh, int j2, int q, const int32_t *acc_o, const int32_t *ext_o, int a, bool rotate,
                                cpx (&x)[16]) {
    const int a_lo = a & 15, a_hi = a >> 4;
    const int j2p = (j2 - a_lo) & 15;
    const int sh = a_hi + (j2 < a_lo ? 1 : 0);
    const int h = sh & 63;
    const uint32_t m = 1u << (q * kBgbit);
    const uint32_t sm = rotate ? (((sh >> 6) & 1) ? 0u - m : m) : 0u;
    const uint32_t cm = rotate ? 0u - m : m;
    const uint32_t offm = kDecompOffset * m;
    const int32_t *own = acc_o + j2 * kAccRow + 16 * hh;
    const int32_t *rot = ext_o + j2p * kExtRow + (kExtOrg - h) + 16 * hh;
#pragma unroll
    for (int part = 0; part < 2; part++) {  // real parts (e < 32), imaginary parts (e >= 32)
        uint32_t vr[16], vo[16];
#pragma unroll
        for (int i = 0; i < 16; i++) vr[i] = (uint32_t) rot[32 * part + i];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const word4 w = *reinterpret_cast<const word4 *>(own + 32 * part + i);
#pragma unroll
            for (int k = 0; k < 4; k++) vo[i + k] = (uint32_t) w.v[k];
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t t = vr[i] * sm + (vo[i] * cm + offm);
            const double d = digit_to_double<kI2F>(t >> (32 - kBgbit));
            if (part == 0) x[i].x = d;
            else x[i].y = d;
        }
    }
}

// stage 0, first half: w = kappa * x with kappa = 1 (hh = 0: the lane holds the a's) or e0 (hh = 1: the b's)
TFHE_HD void phase_f1h_cross_send(int hh, const cpx (&x)[16], cpx (&w)[16]) {
    const double kr = hh ? c1_re_rt(0) : 1.0, ki = hh ? c1_im_rt(0) : 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        w[i].x = fma(kr, x[i].x, -(ki * x[i].y));
        w[i].y = fma(kr, x[i].y, ki * x[i].x);
    }
}

// stage 0, second half (recv = the partner lane's w): hh = 0: a + e0 b, hh = 1: a - e0 b; then stages 1-4
TFHE_HD void phase_f1h_finish(int hh, const cpx (&w)[16], const cpx (&recv)[16], cpx (&x)[16]) {
    const double sg = hh ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        x[i].x = fma(sg, w[i].x, recv[i].x);
        x[i].y = fma(sg, w[i].y, recv[i].y);
    }
#pragma unroll
    for (int s = 1; s < 5; s++) {
        const int half = 16 >> s;
#pragma unroll
        for (int b = 0; b < (1 << (s - 1)); b++) {
            const int c0 = (1 << s) - 1 + b, c1 = c0 + (1 << (s - 1));
            const double er = hh ? c1_re_rt(c1) : c1_re_rt(c0), ei = hh ? c1_im_rt(c1) : c1_im_rt(c0);
#pragma unroll
            for (int i = 0; i < half; i++) bf_fwd(x[b * 2 * half + i], x[b * 2 * half + i + half], er, ei);
        }
    }
}

// x[i] is output position 16 hh + i of slice j2 (bit-reversed frequency class, as in phase_f1_store_p)
TFHE_HD void phase_f1h_store_p(int hh, int j2, cpx *buf, const cpx (&x)[16]) {
    cpx *dst = buf + j2 + hh * kExchRow;
#pragma unroll
    for (int i = 0; i < 16; i++) dst[bitrev5(i) * kExchRow] = x[i];  // bitrev5(16 hh + i) = bitrev5(i) + hh
}

// Pass 2 of the forward transform of a decomposed polynomial from its exchange buffer (lane m1: 16 values).
TFHE_HD void phase_f2_fft_p(int lane, const cpx *buf, const cpx *e2, cpx (&z)[16]) {
    const cpx *src = buf + lane * kExchRow;
#pragma unroll
    for (int j2 = 0; j2 < 16; j2++) z[j2] = src[j2];
    fwd16(z, e2 + lane * kE2Row);
}

TFHE_HD void phase_f2_fft(int lane, WarpSmem &ws, const cpx *e2, int row, cpx (&z)[16]) {
    phase_f2_fft_p(lane, ws.exch[row], e2, z);
}

// Fourier MAC against one result-polynomial half of a TGSW row (tLweFFTAddMulRTo,
// tlwe-fft-operations.cu:286 -> LagrangeHalfCPolynomialAddMul, lagrangehalfc_impl.cu:95-117).
// half: [pos][m1] complex, 8 KiB.
TFHE_HD void phase_mac_half(int lane, const cpx (&z)[16], const cpx *half, cpx (&acc)[16]) {
#pragma unroll
    for (int pos = 0; pos < 16; pos++) cmac(acc[pos], z[pos], half[pos * 32 + lane]);
}

// The same over positions [POS0, POS0 + NPOS) only; `part` points at the ring chunk that holds
// exactly those positions ([pos - POS0][m1]).
template <int POS0, int NPOS>
TFHE_HD void phase_mac_part(int lane, const cpx (&z)[16], const cpx *part, cpx (&acc)[16]) {
#pragma unroll
    for (int p = 0; p < NPOS; p++) cmac(acc[POS0 + p], z[POS0 + p], part[p * 32 + lane]);
}

// Hand the partial sum of the result polynomial the OTHER warp finishes to that warp: warp o parks it
// in its own buffer B (rows this lane owns; B's pass-1 output has been consumed by then).
// (pointer forms: a lane's 16 partial sums parked in / added from rows it owns of an exchange buffer)
TFHE_HD void phase_part_store(int lane, cpx *buf, const cpx (&v)[16]) {
    cpx *d = buf + lane * kExchRow;
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = v[i];
}

TFHE_HD void phase_part_add(int lane, const cpx *buf, cpx (&acc)[16]) {
    const cpx *s = buf + lane * kExchRow;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const cpx v = s[i];
        acc[i].x += v.x;
        acc[i].y += v.y;
    }
}

TFHE_HD void phase_xchg_store(int lane, WarpSmem &ws, int o, const cpx (&give)[16]) {
    phase_part_store(lane, ws.exch[2 * o + 1], give);
}

// keep += partner's partial sum (parked by its phase_xchg_store in ITS buffer B).
TFHE_HD void phase_xchg_load(int lane, const WarpSmem &ws, int o, cpx (&keep)[16]) {
    phase_part_add(lane, ws.exch[2 * (1 - o) + 1], keep);
}

// Inverse pass 2 of the finished Fourier sum of a result polynomial; the 16 outputs go to `buf`.
TFHE_HD void phase_inv16_store_p(int lane, cpx *buf, const cpx *e2, cpx (&keep)[16]) {
    inv16(keep, e2 + lane * kE2Row);
    cpx *d = buf + lane * kExchRow;
#pragma unroll
    for (int j2 = 0; j2 < 16; j2++) d[j2] = keep[j2];
}

// ... of result polynomial o: to buffer A.
TFHE_HD void phase_inv16_store(int lane, WarpSmem &ws, const cpx *e2, int o, cpx (&keep)[16]) {
    phase_inv16_store_p(lane, ws.exch[2 * o], e2, keep);
}

// Inverse pass 1 of result polynomial o, first part: lane (hh, j2) runs the four inner stages on
// positions [16 hh, 16 hh + 16) of slice j2 (the multipliers of block hh: lane dependent).
TFHE_HD void phase_i2_inner_p(int lane, const cpx *buf, cpx (&x)[16]) {
    const int hh = lane >> 4, j2 = lane & 15;
    const cpx *src = buf + j2 + hh * kExchRow;
#pragma unroll
    for (int p = 0; p < 16; p++) {
        // position 16*hh + p holds frequency class m1 = bitrev5(16*hh + p) = 2*bitrev4(p) + hh
        x[p] = src[bitrev5(p) * kExchRow];  // bitrev5(p) for p < 16 is even
    }
#pragma unroll
    for (int s = 4; s >= 1; s--) {
        const int half = 16 >> s;
#pragma unroll
        for (int b = 0; b < (1 << (s - 1)); b++) {
            const int c0 = (1 << s) - 1 + b, c1 = c0 + (1 << (s - 1));
            const double er = hh ? c1_re_rt(c1) : c1_re_rt(c0), ei = hh ? c1_im_rt(c1) : c1_im_rt(c0);
#pragma unroll
            for (int i = 0; i < half; i++) bf_inv(x[b * 2 * half + i], x[b * 2 * half + i + half], er, ei);
        }
    }
}

TFHE_HD void phase_i2_inner(int lane, const WarpSmem &ws, int o, cpx (&x)[16]) {
    phase_i2_inner_p(lane, ws.exch[2 * o], x);
}

// The last stage pairs position i of lane (0, j2) (u) with position i of lane (1, j2) (v).  Lane hh
// finishes the 8 pairs i = 8 hh .. 8 hh + 7 completely, so it hands its partner lane (lane ^ 16) the 8
// values of the OTHER pairs: warp shuffles on the device, an array in the host emulation.
TFHE_HD void phase_i2_send(int lane, const cpx (&x)[16], cpx (&send)[8]) {
    const int hh = lane >> 4;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        send[b].x = hh ? x[b].x : x[8 + b].x;
        send[b].y = hh ? x[b].y : x[8 + b].y;
    }
}

// Inverse pass 1, last stage, conversion to Torus32 (execute_direct_Torus32,
// fft_processor_fftw.cu:168-181; rounding: double_to_torus32) and the tLweAddTo of MuxRotate
// (tlwe-functions.cu:170).  p: the partner lane's 8 values (its phase_i2_send).  Lane hh updates
// coefficients j1 = i and j1 = i + 16 (i = 8 hh + b) of row j2, and the matching upper-half
// coefficients (+32) from the imaginary parts: four groups of 8 consecutive words.
// With mine / theirs = the value of this lane / of the partner lane, u + v = mine + theirs and
// conj(e) (u - v) = (+-conj(e)) (mine - theirs), sign by lane: identical arithmetic in both lanes.
template <int GAP = 0>
TFHE_HD void phase_i2_final_p(int lane, int32_t *acc_o, int32_t *ext_o, const cpx (&x)[16], const cpx (&p)[8]) {
    const int hh = lane >> 4, j2 = lane & 15;
    int32_t *row = acc_o + j2 * kAccRow + 8 * hh;
    int32_t *ext = ext_o + ext_row_off<GAP>(j2) + 8 * hh;  // extended copy, same coefficients
    const double er = hh ? -c1_re_rt(0) : c1_re_rt(0), ei = hh ? -c1_im_rt(0) : c1_im_rt(0);
    // all loads first (128-bit accesses to the master copy: the 8 consecutive coefficients of a group
    // are two aligned quads), then the butterflies and conversions, then the updates
    word4 acc4[4][2];  // group g: coefficients 16 g + 8 hh + (0..7) of the row
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int h = 0; h < 2; h++) acc4[g][h] = *reinterpret_cast<const word4 *>(row + 16 * g + 4 * h);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        cpx mine;
        mine.x = hh ? x[8 + b].x : x[b].x;
        mine.y = hh ? x[8 + b].y : x[b].y;
        const double sre = mine.x + p[b].x, sim = mine.y + p[b].y;  // u + v             -> coefficient j1 = i
        const double tr = mine.x - p[b].x, ti = mine.y - p[b].y;    // conj(e) * (u - v) -> coefficient j1 = i + 16
        const double dre = fma(er, tr, ei * ti), dim = fma(er, ti, -(ei * tr));
        int32_t &c0 = acc4[0][b >> 2].v[b & 3], &c1 = acc4[1][b >> 2].v[b & 3];
        int32_t &c2 = acc4[2][b >> 2].v[b & 3], &c3 = acc4[3][b >> 2].v[b & 3];
        c0 = (int32_t) ((uint32_t) c0 + double_to_torus32(sre));
        c1 = (int32_t) ((uint32_t) c1 + double_to_torus32(dre));
        c2 = (int32_t) ((uint32_t) c2 + double_to_torus32(sim));   // imaginary parts: coefficients + 512
        c3 = (int32_t) ((uint32_t) c3 + double_to_torus32(dim));
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {
#pragma unroll
        for (int h = 0; h < 2; h++) *reinterpret_cast<word4 *>(row + 16 * g + 4 * h) = acc4[g][h];
        // extended copy: +value at kExtOrg + index, -value at index - 1 (index 0 has no left image)
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t nv = (uint32_t) acc4[g][b >> 2].v[b & 3];
            ext[kExtOrg + 16 * g + b] = (int32_t) nv;
            if (16 * g + b > 0 || hh != 0) ext[16 * g + b - 1] = (int32_t) (0u - nv);
        }
    }
}

TFHE_HD void phase_i2_final(int lane, WarpSmem &ws, int o, const cpx (&x)[16], const cpx (&p)[8]) {
    phase_i2_final_p<kExtGapThroughput>(lane, ws.acc[o], ext_poly(ws, o), x, p);
}

// ---- Fourier section of the latency kernel: pass 2, multiply, inverse pass 2 by CLASS OCTETS ---------
// After pass 1 (warp r = decomposed row r) and a CTA barrier, the warps cq and cq + 4 take the frequency
// classes 8 cq .. 8 cq + 7 of ALL FOUR rows and both result polynomials; they meet at two 64-thread
// barriers, nothing else leaves the pair:
//   pass 2:   warp cq, lane (rr, c): the 16-point transform of (row rr, class m1 = 8 cq + c) IN PLACE in the
//             exchange buffer;
//   multiply: warp cq + 4 ph (position half ph), lane (g4, c) re-reads positions 8 ph + 2 g4, + 1 of all four
//             rows of its class (a transposition through shared memory) and accumulates BOTH result
//             polynomials over the four key rows: the sum over rows is a sum in registers ("multiply by
//             position pairs" below);
//   inverse:  stage 3 on the lane's pair, stage 2 through lane ^ 8, values to the inverse buffer; (pair
//             barrier); stages 1, 0 on the four positions {k, k + 4, k + 8, k + 12}, k = 2 ph + kk, lane
//             (kk, oo, c), in place.
// Earlier versions: one warp per decomposed row all the way, the partial sums parked in shared memory and
// three quarters of them read back (18 % of the iteration); then a sum over rows by warp shuffles (a 64-bit
// shuffle pair moves 8 bytes per lane where a 128-bit shared-memory access moves 16).
// pass 2 of (row rr, class m1) in place in the exchange buffer (all 16 positions by one lane)
TFHE_HD void phase_c_f2_inplace(int rr, int m1, cpx (*exch)[kExchPoly], const cpx *e2) {
    cpx *row = exch[rr] + m1 * kExchRow;
    cpx z[16];
#pragma unroll
    for (int j2 = 0; j2 < 16; j2++) z[j2] = row[j2];
    fwd16(z, e2 + m1 * kE2Row);
#pragma unroll
    for (int pos = 0; pos < 16; pos++) row[pos] = z[pos];
}

// ---- the Fourier section by ONE warp per class octet (four-warp latency kernel, blind_rotate_quad_kernel) ----
// Warp cq alone: pass 2 in place (phase_c_f2_inplace), then lane (g, c) re-reads positions 4 g .. 4 g + 3 of the
// four rows of class m1 = 8 cq + c, accumulates BOTH result polynomials, runs inverse stages 3, 2 inside its block
// of four positions and stores to the inverse buffer; after a __syncwarp lane (oo, k2, c) runs stages 1, 0 on
// positions {2 k2, 2 k2 + 1} + {0, 4, 8, 12} of result polynomial oo in place.  No barrier inside the section.
TFHE_HD void phase_w_load_rows(int g, int m1, const cpx (*exch)[kExchPoly], cpx (&zr)[kKpl][4]) {
#pragma unroll
    for (int row = 0; row < kKpl; row++) {
        const cpx *src = exch[row] + m1 * kExchRow + 4 * g;
#pragma unroll
        for (int i = 0; i < 4; i++) zr[row][i] = src[i];
    }
}

// lane constants of the inverse stages 3 and 2 for the block of four positions g (pass2_const)
TFHE_HD void phase_w_inv_consts(int g, const cpx *e, cpx &c3, cpx &c2) {
    const cpx g3 = e[3], g2 = e[2];
    cpx h4, h8, h38;
    h4.x = (g3.x - g3.y) * kSqrtHalf;
    h4.y = (g3.x + g3.y) * kSqrtHalf;
    h8 = cmul_const(g3, kCosPi8, kSinPi8);
    h38 = cmul_const(g3, kSinPi8, kCosPi8);
    c3 = g == 0 ? g3 : (g == 1 ? h4 : (g == 2 ? h8 : h38));  // stage 3, blocks 2g (c3) and 2g + 1 (i * c3)
    cpx b;
    b.x = (g >> 1) ? (g2.x - g2.y) * kSqrtHalf : g2.x;       // stage 2, block g: base g >> 1, odd block times i
    b.y = (g >> 1) ? (g2.x + g2.y) * kSqrtHalf : g2.y;
    c2.x = (g & 1) ? -b.y : b.x;
    c2.y = (g & 1) ? b.x : b.y;
}

TFHE_HD void phase_w_inv_a_store(int g, int m1, cpx *inv_o, const cpx &c3, const cpx &c2, cpx (&z)[4]) {
    bf_inv(z[0], z[1], c3.x, c3.y);
    bf_inv(z[2], z[3], -c3.y, c3.x);
    bf_inv(z[0], z[2], c2.x, c2.y);
    bf_inv(z[1], z[3], c2.x, c2.y);
    cpx *d = inv_o + m1 * kExchRow + 4 * g;
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = z[i];
}

TFHE_HD void phase_w_inv_b_inplace(int k2, int m1, cpx *inv_o, const cpx &g1, const cpx &g0) {
    cpx *d = inv_o + m1 * kExchRow + 2 * k2;
    cpx z[2][4];
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
        for (int m = 0; m < 4; m++) z[e][m] = d[e + 4 * m];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        bf_inv(z[e][0], z[e][1], g1.x, g1.y);    // stage 1, block 0
        bf_inv(z[e][2], z[e][3], -g1.y, g1.x);   // stage 1, block 1: times i
        bf_inv(z[e][0], z[e][2], g0.x, g0.y);    // stage 0
        bf_inv(z[e][1], z[e][3], g0.x, g0.y);
    }
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
        for (int m = 0; m < 4; m++) d[e + 4 * m] = z[e][m];
}

// ---- multiply by POSITION PAIRS (latency kernel) ------------------------------------------------------
// Lane (g4, c) of warp (cq, ph) takes the two positions p0 = 8 ph + 2 g4, p0 + 1 of class m1 for BOTH result
// polynomials: every Fourier value is read once (the (g2, oo, c) layout read each one twice, once per result
// polynomial: 256 of the 1,024 shared-memory wavefronts of the multiply phase, which saturates that pipe).
// Inverse stage 3 pairs the lane's own two positions; stage 2 pairs position i of lane g4 = 2b with position i
// of lane 2b + 1: one exchange through lane ^ 8 (u' = u + v on the even lane, v' = conj(c2) (u - v) on the
// odd one: identical arithmetic in both lanes, out = kappa * (recv +- mine)).
TFHE_HD cpx times_i(const cpx &c) {
    cpx r;
    r.x = -c.y;
    r.y = c.x;
    return r;
}

TFHE_HD void phase_p_load_rows(int p0, int m1, const cpx (*exch)[kExchPoly], cpx (&zr)[kKpl][2]) {
#pragma unroll
    for (int row = 0; row < kKpl; row++) {
        const cpx *src = exch[row] + m1 * kExchRow + p0;
        zr[row][0] = src[0];
        zr[row][1] = src[1];
    }
}

// lane constants: c3 = multiplier of inverse stage 3 for the pair p0 / 2; k2 = conj-multiplier of stage 2 for
// the odd lane of a block of four positions, 1 for the even lane
TFHE_HD void phase_p_inv_consts(int ph, int g4, const cpx *e, cpx &c3, cpx &k2) {
    const cpx g3 = e[3], g2 = e[2];
    cpx h4, h8, h38;
    h4.x = (g3.x - g3.y) * kSqrtHalf;
    h4.y = (g3.x + g3.y) * kSqrtHalf;
    h8 = cmul_const(g3, kCosPi8, kSinPi8);
    h38 = cmul_const(g3, kSinPi8, kCosPi8);
    const int base3 = 2 * ph + (g4 >> 1);                 // stage 3, block 4 ph + g4: base (4 ph + g4) >> 1
    cpx b3 = base3 == 0 ? g3 : (base3 == 1 ? h4 : (base3 == 2 ? h8 : h38));
    if (g4 & 1) b3 = times_i(b3);
    c3 = b3;
    cpx b2;                                                // stage 2, block 2 ph + (g4 >> 1): base ph
    b2.x = ph ? (g2.x - g2.y) * kSqrtHalf : g2.x;
    b2.y = ph ? (g2.x + g2.y) * kSqrtHalf : g2.y;
    if ((g4 >> 1) & 1) b2 = times_i(b2);
    k2.x = (g4 & 1) ? b2.x : 1.0;
    k2.y = (g4 & 1) ? b2.y : 0.0;
}

// stage 2, second half: recv = the partner lane's value of the same element
TFHE_HD void phase_p_inv_cross(int odd, const cpx &k2, const cpx &recv, cpx &z) {
    const double sg = odd ? -1.0 : 1.0;
    const double dx = fma(sg, z.x, recv.x), dy = fma(sg, z.y, recv.y);
    z.x = fma(k2.x, dx, k2.y * dy);
    z.y = fma(k2.x, dy, -(k2.y * dx));
}

TFHE_HD void phase_p_inv_store(int p0, int m1, cpx *inv_o, const cpx (&z)[2]) {
    cpx *d = inv_o + m1 * kExchRow + p0;
    d[0] = z[0];
    d[1] = z[1];
}

// inverse stages 1 and 0 on positions {k, k + 4, k + 8, k + 12} of class m1, in place
TFHE_HD void phase_o_inv_b_inplace(int k, int m1, cpx *inv_o, const cpx &g1, const cpx &g0) {
    cpx *d = inv_o + m1 * kExchRow + k;
    cpx z[4];
#pragma unroll
    for (int m = 0; m < 4; m++) z[m] = d[4 * m];
    bf_inv(z[0], z[1], g1.x, g1.y);    // stage 1, block 0
    bf_inv(z[2], z[3], -g1.y, g1.x);   // stage 1, block 1: times i
    bf_inv(z[0], z[2], g0.x, g0.y);    // stage 0
    bf_inv(z[1], z[3], g0.x, g0.y);
#pragma unroll
    for (int m = 0; m < 4; m++) d[4 * m] = z[m];
}

// Inverse pass 1 (32 points per slice j2): warp h takes the slices j2 = 8 h .. 8 h + 7, lane (qq, j2') the
// positions 8 qq .. 8 qq + 7 of slice 8 h + j2'.  Stages 4, 3, 2 are local (multipliers of the lane's
// blocks: run-time index into the pass-1 table), stage 1 pairs qq with qq ^ 1 (lane ^ 8), stage 0 qq with
// qq ^ 2 (lane ^ 16).  After stage 0 the lane holds the coefficients j1 = 8 qq .. 8 qq + 7 of the slice.
TFHE_HD void phase_q_i2_local(int lane, int h, const cpx *buf, cpx (&x)[8]) {
    const int qq = lane >> 3, j2 = 8 * h + (lane & 7);
    // position 8 qq + k holds frequency class bitrev5(8 qq + k) = 4 bitrev3(k) + bitrev2(qq)
    const cpx *src = buf + j2 + (((qq & 1) << 1) | (qq >> 1)) * kExchRow;
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = src[4 * (((k & 1) << 2) | (k & 2) | (k >> 2)) * kExchRow];
#pragma unroll
    for (int b = 0; b < 4; b++) {  // stage 4: blocks 4 qq + b
        const int ci = 15 + 4 * qq + b;
        bf_inv(x[2 * b], x[2 * b + 1], c1_re_rt(ci), c1_im_rt(ci));
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {  // stage 3: blocks 2 qq + b
        const int ci = 7 + 2 * qq + b;
        const double er = c1_re_rt(ci), ei = c1_im_rt(ci);
        bf_inv(x[4 * b], x[4 * b + 2], er, ei);
        bf_inv(x[4 * b + 1], x[4 * b + 3], er, ei);
    }
    {   // stage 2: block qq
        const double er = c1_re_rt(3 + qq), ei = c1_im_rt(3 + qq);
#pragma unroll
        for (int i = 0; i < 4; i++) bf_inv(x[i], x[i + 4], er, ei);
    }
}

// one cross-lane stage: is_v = this lane holds the v's of the pairs; ci = the stage's multiplier
TFHE_HD void phase_q_i2_cross(bool is_v, int ci, const cpx (&recv)[8], cpx (&x)[8]) {
    const double sg = is_v ? -1.0 : 1.0;
    const double kx = is_v ? c1_re_rt(ci) : 1.0, ky = is_v ? c1_im_rt(ci) : 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double dx = fma(sg, x[i].x, recv[i].x), dy = fma(sg, x[i].y, recv[i].y);
        x[i].x = fma(kx, dx, ky * dy);
        x[i].y = fma(kx, dy, -(ky * dx));
    }
}

// conversion to Torus32 and accumulator update of the 16 coefficients of lane (qq, j2'):
// words 8 qq .. 8 qq + 7 (real parts) and 32 + 8 qq .. (imaginary parts) of row j2; extended copy (GAP = 0)
TFHE_HD void phase_q_final(int lane, int h, int32_t *acc_o, int32_t *ext_o, const cpx (&x)[8]) {
    const int qq = lane >> 3, j2 = 8 * h + (lane & 7);
    int32_t *row = acc_o + j2 * kAccRow + 8 * qq;
    int32_t *ext = ext_o + j2 * kExtRow + 8 * qq;
    word4 w[2][2];  // [real / imaginary][quad]
#pragma unroll
    for (int part = 0; part < 2; part++)
#pragma unroll
        for (int v = 0; v < 2; v++) w[part][v] = *reinterpret_cast<const word4 *>(row + 32 * part + 4 * v);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int32_t &re = w[0][k >> 2].v[k & 3], &im = w[1][k >> 2].v[k & 3];
        re = (int32_t) ((uint32_t) re + double_to_torus32(x[k].x));
        im = (int32_t) ((uint32_t) im + double_to_torus32(x[k].y));
    }
#pragma unroll
    for (int part = 0; part < 2; part++) {
#pragma unroll
        for (int v = 0; v < 2; v++) *reinterpret_cast<word4 *>(row + 32 * part + 4 * v) = w[part][v];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t nv = (uint32_t) w[part][k >> 2].v[k & 3];
            ext[kExtOrg + 32 * part + k] = (int32_t) nv;
            if (part != 0 || k > 0 || qq != 0) ext[32 * part + k - 1] = (int32_t) (0u - nv);
        }
    }
}

// external product only: clear the 16 coefficients this lane will set
TFHE_HD void phase_q_acc_clear(int lane, int h, int32_t *acc_o) {
    const int qq = lane >> 3, j2 = 8 * h + (lane & 7);
    int32_t *row = acc_o + j2 * kAccRow + 8 * qq;
    word4 z;
#pragma unroll
    for (int k = 0; k < 4; k++) z.v[k] = 0;
#pragma unroll
    for (int part = 0; part < 2; part++)
#pragma unroll
        for (int v = 0; v < 2; v++) *reinterpret_cast<word4 *>(row + 32 * part + 4 * v) = z;
}

// Stand-alone external product: the result REPLACES the accumulator, so the master copy is
// cleared once its decomposition has been read (lane (hh, j2) clears the coefficients it will update).
TFHE_HD void phase_acc_clear_p(int lane, int32_t *acc_o) {
    const int hh = lane >> 4, j2 = lane & 15;
    int32_t *row = acc_o + j2 * kAccRow + 8 * hh;
    word4 z;
#pragma unroll
    for (int k = 0; k < 4; k++) z.v[k] = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        *reinterpret_cast<word4 *>(row + 16 * g) = z;
        *reinterpret_cast<word4 *>(row + 16 * g + 4) = z;
    }
}

TFHE_HD void phase_acc_clear(int lane, WarpSmem &ws, int o) { phase_acc_clear_p(lane, ws.acc[o]); }

// coefficient j of polynomial o of a master copy int32[k+1][kAccPoly]
TFHE_HD int32_t acc_coef_p(const int32_t (*acc)[kAccPoly], int o, int j) {
    return acc[o][(j & 15) * kAccRow + (j >> 4)];
}

// Sample extraction at index 0 (tLweExtractLweSampleIndex, lwe.cu:41-56):
// u.a[0] = ACC.a[0], u.a[j] = -ACC.a[N-j], u.b = ACC.b[0].  u: int32[N+1].
// `step` lanes share the work (32: one warp; 256: the eight warps of the latency kernel).
TFHE_HD void phase_extract_p(int lane, int step, const int32_t (*acc)[kAccPoly], int32_t *u) {
    for (int j = lane; j < kN; j += step) {
        const int32_t v = (j == 0) ? acc_coef_p(acc, 0, 0) : (int32_t) (0u - (uint32_t) acc_coef_p(acc, 0, kN - j));
        u[j] = v;
    }
    if (lane == 0) u[kN] = acc_coef_p(acc, kK, 0);
}

// Raw accumulator dump, natural coefficient order: int32[2][N].
TFHE_HD void phase_dump_acc_p(int lane, int step, const int32_t (*acc)[kAccPoly], int32_t *out) {
    for (int j = lane; j < (kK + 1) * kN; j += step) out[j] = acc_coef_p(acc, j >> 10, j & (kN - 1));
}

TFHE_HD void phase_load_acc_p(int lane, int step, int32_t (*acc)[kAccPoly], const int32_t *in) {
    for (int j = lane; j < (kK + 1) * kN; j += step) {
        const int o = j >> 10, c = j & (kN - 1);
        acc[o][(c & 15) * kAccRow + (c >> 4)] = in[j];
    }
}

TFHE_HD int32_t acc_coef(const WarpSmem &ws, int o, int j) { return acc_coef_p(ws.acc, o, j); }
TFHE_HD void phase_extract(int lane, const WarpSmem &ws, int32_t *u) { phase_extract_p(lane, 32, ws.acc, u); }
TFHE_HD void phase_dump_acc(int lane, const WarpSmem &ws, int32_t *out) { phase_dump_acc_p(lane, 32, ws.acc, out); }
TFHE_HD void phase_load_acc(int lane, WarpSmem &ws, const int32_t *in) { phase_load_acc_p(lane, 32, ws.acc, in); }

// ------------------------------------------------ generic forward transform -
// Forward transform of 4 polynomials given as doubles via `fetch(p, j)`; used
// for the key (TorusPolynomial_ifft in tGswToFFTConvert, tgsw-fft-operations.cu:84)
// and for the standalone IntPolynomial_ifft entry point.  Pass 1: lane (o, j2)
// handles polynomials 2o and 2o+1.
template <typename Fetch>
TFHE_HD void fwd4_pass1(int lane, WarpSmem &ws, Fetch fetch) {
    const int o = lane >> 4, j2 = lane & 15;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int p = o * 2 + q;
        cpx x[32];
#pragma unroll
        for (int j1 = 0; j1 < 32; j1++) {
            x[j1].x = fetch(p, 16 * j1 + j2);
            x[j1].y = fetch(p, 16 * j1 + j2 + kM);
        }
        fwd32(x);
        cpx *dst = ws.exch[p] + j2;
#pragma unroll
        for (int pos = 0; pos < 32; pos++) dst[bitrev5(pos) * kExchRow] = x[pos];
    }
}

// Pass 2: writes polynomial p's 512 values to out[p*512 + pos*32 + m1].
TFHE_HD void fwd4_pass2(int lane, WarpSmem &ws, const cpx *e2, cpx *out) {
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
        cpx z[16];
        const cpx *src = ws.exch[p] + lane * kExchRow;
#pragma unroll
        for (int j2 = 0; j2 < 16; j2++) z[j2] = src[j2];
        fwd16(z, e2 + lane * kE2Row);
#pragma unroll
        for (int pos = 0; pos < 16; pos++) out[p * kM + pos * 32 + lane] = z[pos];
    }
}

// Frequency index m held at (pos, m1) of the device layout.
TFHE_HD constexpr int freq_of(int pos, int m1) {
    return m1 + 32 * (((pos & 1) << 3) | ((pos & 2) << 1) | ((pos & 4) >> 1) | ((pos & 8) >> 3));
}

}  // namespace tfhe_b200
