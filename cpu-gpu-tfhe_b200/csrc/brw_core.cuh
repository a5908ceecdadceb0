// Blind-rotation core, ONE warp per ciphertext (blind_rotate.cu, kernel variant "warp").
//
// Why: the warp-pair kernel (br_core.cuh) is bound by the LATENCY of one ciphertext's
// iteration, not by any SM resource: ncu/gpurun measurements show the same 8.6 us per
// iteration with 1, 2 or 4 ciphertexts per SM (fp64 pipe 47 %, issue 48 %, shared memory
// 61 % at 4).  Throughput therefore scales with the number of ciphertexts resident on an SM,
// and that number was capped at 4 by shared memory (43 KB each) and registers (2 warps x 255).
// Here a ciphertext needs 24 KB and ONE warp, so 8 are resident.
//
// What changes with respect to br_core.cuh (the transform itself, the frequency order of
// the results and therefore the key layout are IDENTICAL):
//   * index split j = 32*j1 + j2 (j1 < 16, j2 < 32): pass 1 = stages 0-3 on the 16
//     register-resident elements of lane j2 (all 32 lanes work on ONE polynomial, so a
//     transform needs one exchange buffer and half the registers); stage 4 moves into pass 2:
//     lane m1 = (p, h) reads both inputs of its 16 stage-4 butterflies and computes only the
//     output it keeps (u for h = 0, v for h = 1: 4 FMA per value), then runs stages 5-8 as
//     before.  The inverse mirrors this.
//   * the accumulator is stored in natural coefficient order (lanes = consecutive
//     coefficients: conflict free without padding); polynomial 1 is stored with bit 4 of the
//     index flipped so that the final update (16 lanes per polynomial) is conflict free too.
//   * exchange rows are 32 complex, 16-byte units XOR-swizzled by (row >> 1) & 7.
//   * no cross-warp synchronisation at all: __syncwarp() between passes, and the key ring.
#pragma once

#include "br_core.cuh"

namespace tfhe_b200 {

struct CtSmem {
    int32_t acc[kK + 1][kN];  // 2 x 4 KiB, word j of polynomial o at acc[o][j ^ (16 * o)]
    cpx x[2][kM];             // 2 x 8 KiB exchange buffers
};

// load batch sizes (register pressure: the two Fourier-domain sums, 128 registers, are live
// across pass 1 and pass 2 of the second digit level)
#ifndef TFHE_B200_F1_BATCH
#define TFHE_B200_F1_BATCH 4
#endif
#ifndef TFHE_B200_F2_BATCH
#define TFHE_B200_F2_BATCH 4
#endif
constexpr int kF1Batch = TFHE_B200_F1_BATCH, kF2Batch = TFHE_B200_F2_BATCH;

TFHE_HD constexpr int bitrev4(int v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3); }

TFHE_HD int w_acc_index(int o, int j) { return j ^ (o << 4); }

// 16-byte unit of element (row p, column col) of an exchange buffer
TFHE_HD int w_xunit(int p, int col) { return p * 32 + (col ^ ((p >> 1) & 7)); }

// Stages 0-3 on the 16 in-register elements of one lane (natural j1 in, position p out;
// position p is the block whose stage-4 multiplier is c1[15 + p]).
TFHE_HD void w_fwd16a(cpx (&x)[16]) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int half = 8 >> s;
#pragma unroll
        for (int b = 0; b < (1 << s); b++) {
            const int ci = (1 << s) - 1 + b;
#pragma unroll
            for (int i = 0; i < half; i++)
                bf_fwd(x[b * 2 * half + i], x[b * 2 * half + i + half], c1_re_rt(ci), c1_im_rt(ci));
        }
    }
}

TFHE_HD void w_inv16a(cpx (&x)[16]) {
#pragma unroll
    for (int s = 3; s >= 0; s--) {
        const int half = 8 >> s;
#pragma unroll
        for (int b = 0; b < (1 << s); b++) {
            const int ci = (1 << s) - 1 + b;
#pragma unroll
            for (int i = 0; i < half; i++)
                bf_inv(x[b * 2 * half + i], x[b * 2 * half + i + half], c1_re_rt(ci), c1_im_rt(ci));
        }
    }
}

// Entry 4 of a lane's row of the pass-2 constant table: the stage-4 multiplier of lane
// m1 = (p, h), +exp(i*pi*d_p) for h = 0 (u = a + e*b) and -exp(i*pi*d_p) for h = 1.
// d_p = shift of block p after four stages = (bitrev4(p) + 1/4) / 16.
TFHE_HD double w_stage4_shift(int m1) { return ((double) (m1 & 15) + 0.25) / 16.0; }

// ACC = (0, X^{2N-barb} * (mu, ..., mu))   (tfhe_blindRotateAndExtract_FFT :1425-1431)
TFHE_HD void w_phase_init(int lane, CtSmem &S, int barb, int32_t mu) {
    for (int j = lane; j < kN; j += 32) {
        S.acc[0][w_acc_index(0, j)] = 0;
        S.acc[kK][w_acc_index(kK, j)] = (((j + barb) & (2 * kN - 1)) < kN) ? mu : (int32_t) (0u - (uint32_t) mu);
    }
}

TFHE_HD void w_phase_init_testvect(int lane, CtSmem &S, int barb, const int32_t *tv) {
    for (int j = lane; j < kN; j += 32) {
        const int s = (j + barb) & (2 * kN - 1);
        const uint32_t v = (uint32_t) tv[s & (kN - 1)];
        S.acc[0][w_acc_index(0, j)] = 0;
        S.acc[kK][w_acc_index(kK, j)] = (int32_t) (s < kN ? v : 0u - v);
    }
}

TFHE_HD void w_phase_load_acc(int lane, CtSmem &S, const int32_t *in) {
    for (int j = lane; j < (kK + 1) * kN; j += 32) S.acc[j >> 10][w_acc_index(j >> 10, j & (kN - 1))] = in[j];
}

TFHE_HD void w_phase_dump_acc(int lane, const CtSmem &S, int32_t *out) {
    for (int j = lane; j < (kK + 1) * kN; j += 32) out[j] = S.acc[j >> 10][w_acc_index(j >> 10, j & (kN - 1))];
}

// tLweExtractLweSampleIndex(.., 0), lwe.cu:41-56
TFHE_HD void w_phase_extract(int lane, const CtSmem &S, int32_t *u) {
    for (int j = lane; j < kN; j += 32) {
        const int32_t v = (j == 0) ? S.acc[0][0] : (int32_t) (0u - (uint32_t) S.acc[0][kN - j]);
        u[j] = v;
    }
    if (lane == 0) u[kN] = S.acc[kK][w_acc_index(kK, 0)];
}

// Pass 1 of the forward transform of decomposed polynomial (o, q), fused with the rotation
// (torusPolynomialMulByXaiMinusOne, toruspolynomial-functions.cu:191-213) and the gadget
// decomposition (tGswTorus32PolynomialDecompH, tgsw-functions.cu:301-352).  Lane = j2.
// Output: exchange buffer S.x[o].  rotate == false: plain decomposition of ACC.
TFHE_HD void w_phase_f1(int lane, CtSmem &S, int a, int o, int q, bool rotate = true) {
    const int flip = o << 4;
    const int32_t *acc = S.acc[o];
    const int32_t *own = acc + (lane ^ flip);  // word j = 32*j1 + lane of polynomial o
    const int shift = 32 - (q + 1) * kBgbit;
    const uint32_t rmask = rotate ? 0xffffffffu : 0u;
    const int base = lane - a;  // (j - a) for j1 = 0
    cpx x[16];
#pragma unroll
    for (int blk = 0; blk < 16; blk += kF1Batch) {
        uint32_t vr[kF1Batch], vi[kF1Batch], wr[kF1Batch], wi[kF1Batch];
        // all loads of a block first so that their latencies overlap
#pragma unroll
        for (int i = 0; i < kF1Batch; i++) {
            const int w = ((base + 32 * (blk + i)) & (kN - 1)) ^ flip;
            vr[i] = (uint32_t) acc[w];
            vi[i] = (uint32_t) acc[w ^ kM];  // (idx + N/2) mod N
            wr[i] = (uint32_t) own[32 * (blk + i)];
            wi[i] = (uint32_t) own[32 * (blk + i) + kM];
        }
#pragma unroll
        for (int i = 0; i < kF1Batch; i++) {
            const int idx = base + 32 * (blk + i);                         // (j - a), any sign
            const uint32_t neg_r = 0u - (uint32_t) ((idx >> 10) & 1);         // X^N = -1
            const uint32_t neg_i = 0u - (uint32_t) (((idx + kM) >> 10) & 1);
            const uint32_t tr = ((vr[i] ^ neg_r) - neg_r) & rmask;
            const uint32_t ti = ((vi[i] ^ neg_i) - neg_i) & rmask;
            const uint32_t ur = (rotate ? tr - wr[i] : wr[i]) + kDecompOffset;
            const uint32_t ui = (rotate ? ti - wi[i] : wi[i]) + kDecompOffset;
            x[blk + i].x = digit_to_double((ur >> shift) & 1023u);
            x[blk + i].y = digit_to_double((ui >> shift) & 1023u);
        }
    }
    w_fwd16a(x);
    cpx *dst = S.x[o];
#pragma unroll
    for (int p = 0; p < 16; p++) dst[w_xunit(p, lane)] = x[p];
}

// Pass 2 of the forward transform from exchange buffer `buf` (lane = frequency class m1):
// stage 4 on the fly, then stages 5-8.  z[pos] = value at frequency m1 + 32*bitrev4(pos).
TFHE_HD void w_phase_f2(int lane, const cpx *buf, const cpx *e2, cpx (&z)[16]) {
    const int p = bitrev4(lane & 15);
    const cpx E = e2[lane * kE2Row + 4];
    const cpx *row = buf + p * 32;
    const int s = (p >> 1) & 7;  // column c lives at unit c ^ s: only the low three bits move
#pragma unroll
    for (int blk = 0; blk < 16; blk += kF2Batch) {
        cpx va[kF2Batch], vb[kF2Batch];
#pragma unroll
        for (int i = 0; i < kF2Batch; i++) {
            const cpx *src = row + (((blk + i) & 7) ^ s);
            va[i] = src[(blk + i) & 8];
            vb[i] = src[16 + ((blk + i) & 8)];
        }
#pragma unroll
        for (int i = 0; i < kF2Batch; i++) {
            double zr = fma(E.x, vb[i].x, va[i].x);
            double zi = fma(E.x, vb[i].y, va[i].y);
            z[blk + i].x = fma(-E.y, vb[i].y, zr);
            z[blk + i].y = fma(E.y, vb[i].x, zi);
        }
    }
    fwd16(z, e2 + lane * kE2Row);
}

// Inverse pass 2 of one result polynomial (lane = m1 = (p, h)): stages 8..5, values parked
// at row p, columns 16*h + i of exchange buffer `buf` (stage 4 is undone by the reader).
TFHE_HD void w_phase_i1(int lane, cpx *buf, const cpx *e2, cpx (&acc)[16]) {
    const int p = bitrev4(lane & 15), h = lane >> 4;
    inv16(acc, e2 + lane * kE2Row);
    cpx *row = buf + p * 32 + 16 * h;
    const int s = (p >> 1) & 7;
#pragma unroll
    for (int i = 0; i < 16; i++) row[((i & 7) ^ s) + (i & 8)] = acc[i];
}

// Inverse pass 1, coefficient half g (j2 = 16*g + i), both result polynomials at once:
// lane = (o, i).  Undoes stage 4 (u + v for g = 0, conj(e_p)*(u - v) for g = 1: g is uniform
// over the warp), stages 3..0, conversion to Torus32 (fft_processor_fftw.cu:168-181) and the
// tLweAddTo of MuxRotate.  accumulate == false: the result replaces ACC.
template <int G>
TFHE_HD void w_phase_i2(int lane, CtSmem &S, bool accumulate = true) {
    const int o = lane >> 4, i = lane & 15;
    const cpx *buf = S.x[o];
    cpx y[16];
#pragma unroll
    for (int blk = 0; blk < 16; blk += 8) {
        cpx u[8], v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u[k] = buf[w_xunit(blk + k, i)];
            v[k] = buf[w_xunit(blk + k, 16 + i)];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (G == 0) {
                y[blk + k].x = u[k].x + v[k].x;
                y[blk + k].y = u[k].y + v[k].y;
            } else {
                const double er = c1_re_rt(15 + blk + k), ei = c1_im_rt(15 + blk + k);
                const double tr = u[k].x - v[k].x, ti = u[k].y - v[k].y;
                y[blk + k].x = fma(er, tr, ei * ti);
                y[blk + k].y = fma(er, ti, -(ei * tr));
            }
        }
    }
    w_inv16a(y);
    int32_t *own = S.acc[o] + ((16 * G + i) ^ (o << 4));  // word j = 32*j1 + 16*G + i
#pragma unroll
    for (int blk = 0; blk < 16; blk += 8) {
        uint32_t old_r[8], old_i[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            old_r[k] = accumulate ? (uint32_t) own[32 * (blk + k)] : 0u;
            old_i[k] = accumulate ? (uint32_t) own[32 * (blk + k) + kM] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            own[32 * (blk + k)] = (int32_t) (old_r[k] + double_to_torus32(y[blk + k].x));
            own[32 * (blk + k) + kM] = (int32_t) (old_i[k] + double_to_torus32(y[blk + k].y));
        }
    }
}

}  // namespace tfhe_b200
