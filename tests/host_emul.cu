// CPU lane-by-lane emulation of the blind-rotation kernel's phase functions
// (cpu-gpu-tfhe_b200/csrc/br_core.cuh are __host__ __device__).  Checks, with no
// GPU: the twisted FFT against the defining sum, the key layout, and one
// MuxRotate step (rotation + decomposition + external product + add) against
// the oracle's exact-integer path, plus init and extraction.
// Built and run by tests/test_host_emulation.py.  Links oracle/liboracle.so
// (allowed: this is a test).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// distance of every converted fp64 value from the nearest integer (round-to-nearest margin)
static double g_max_frac = 0.0, g_sum_frac2 = 0.0;
static long g_conversions = 0;
#define TFHE_B200_CONV_PROBE(x)                          \
    do {                                                 \
        const double f_ = fabs((x) - nearbyint(x));      \
        if (f_ > g_max_frac) g_max_frac = f_;            \
        g_sum_frac2 += f_ * f_;                          \
        g_conversions++;                                 \
    } while (0)
#include "../cpu-gpu-tfhe_b200/csrc/br_core.cuh"
#include "../oracle/tfhe_oracle.h"

using namespace tfhe_b200;

static std::vector<cpx> make_e2() {
    std::vector<cpx> e2(32 * kE2Row);
    for (int m1 = 0; m1 < 32; m1++)
        for (int idx = 0; idx < 4; idx++) {
            const double d = e2_shift(m1, idx);
            e2[m1 * kE2Row + idx].x = cos(M_PI * d);
            e2[m1 * kE2Row + idx].y = sin(M_PI * d);
        }
    return e2;
}

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("\n"); fails++; } } while (0)

int main() {
    std::vector<cpx> e2 = make_e2();
    WarpSmem *ws = new WarpSmem();
    srand(12345);

    // ---- A. forward transform vs the defining sum -------------------------
    {
        std::vector<double> polys(4 * kN);
        for (auto &v : polys) v = (double) ((rand() % 1024) - 512);
        auto fetch = [&](int p, int j) { return polys[p * kN + j]; };
        for (int lane = 0; lane < 32; lane++) fwd4_pass1(lane, *ws, fetch);
        std::vector<cpx> out(4 * kM);
        for (int lane = 0; lane < 32; lane++) fwd4_pass2(lane, *ws, e2.data(), out.data());
        double maxerr = 0;
        for (int p = 0; p < 4; p++)
            for (int pos = 0; pos < 16; pos++)
                for (int m1 = 0; m1 < 32; m1 += 7) {
                    const int m = freq_of(pos, m1);
                    long double re = 0, im = 0;
                    for (int j = 0; j < kM; j++) {
                        const long double ang = 2.0L * M_PIl * (long double) j * ((long double) m + 0.25L) / 512.0L;
                        const long double c = cosl(ang), s = sinl(ang);
                        const long double a = polys[p * kN + j], b = polys[p * kN + j + kM];
                        re += a * c - b * s;
                        im += a * s + b * c;
                    }
                    const cpx g = out[p * kM + pos * 32 + m1];
                    maxerr = fmax(maxerr, fmax(fabs((double) (re - g.x)), fabs((double) (im - g.y))));
                }
        printf("A. forward transform max abs err %.3e\n", maxerr);
        CHECK(maxerr < 1e-8, "forward transform mismatch");
    }

    // ---- B. one MuxRotate step vs exact integer arithmetic -----------------
    OracleParams P;
    oracle_default_params(&P);
    P.n = 6;
    std::vector<int32_t> lwe_key(P.n), tlwe_key(P.N), bk(oracle_bk_words(&P)), ks(oracle_ks_words(&P));
    oracle_keygen(&P, 99, lwe_key.data(), tlwe_key.data(), bk.data(), ks.data());

    // key -> device layout [i][row][o][pos][m1], scale 2^-9 (= 2^-32 * 2^32 / 512)
    std::vector<cpx> bkdev((size_t) P.n * kBkIterCplx);
    for (int g = 0; g < P.n * kKpl * 2 / 4; g++) {
        const int32_t *src = bk.data() + (size_t) g * 4 * kN;
        auto fetch = [&](int p, int j) { return (double) src[p * kN + j] * (1.0 / 512.0); };
        for (int lane = 0; lane < 32; lane++) fwd4_pass1(lane, *ws, fetch);
        for (int lane = 0; lane < 32; lane++) fwd4_pass2(lane, *ws, e2.data(), bkdev.data() + (size_t) g * 4 * kM);
    }

    // init
    const int32_t mu = oracle_modswitch_to(1, 8);
    for (int barb : {0, 1, 17, 1023, 1024, 1500, 2047}) {
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_init(lane, *ws, o, barb, mu);
        std::vector<int32_t> got(2 * kN), tv(kN, mu), exp_b(kN);
        for (int lane = 0; lane < 32; lane++) phase_dump_acc(lane, *ws, got.data());
        if (barb) oracle_mul_by_xai(2 * kN - barb, kN, tv.data(), exp_b.data());
        else exp_b = tv;
        bool ok = true;
        for (int j = 0; j < kN; j++) ok = ok && got[j] == 0 && got[kN + j] == exp_b[j];
        CHECK(ok, "phase_init barb=%d", barb);
    }

    // ---- B0. rotation + gadget decomposition, EVERY rotation amount, against the definition ----
    // (toruspolynomial-functions.cu:191-213 torusPolynomialMulByXaiMinusOne, tgsw-functions.cu:301-352)
    {
        std::vector<int32_t> acc0(2 * kN);
        for (auto &v : acc0) v = (int32_t) ((((uint32_t) rand()) << 16) ^ (uint32_t) rand() ^ (((uint32_t) rand()) << 31));
        for (int lane = 0; lane < 32; lane++) phase_load_acc(lane, *ws, acc0.data());
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_ext_build(lane, *ws, o);
        long bad = 0;
        for (int a = 0; a < 2 * kN; a++) {
            for (int rot = 1; rot >= 0; rot--) {
                if (!rot && a > 3) continue;  // rotate == false ignores a: a few values are enough
                for (int o = 0; o < 2; o++)
                    for (int lane = 0; lane < 32; lane++) {
                        cpx x[32];
                        phase_f1_decomp(lane, *ws, o, a, rot != 0, x);
                        const int q = lane >> 4, j2 = lane & 15;
                        for (int e = 0; e < 64; e++) {
                            const int j = 16 * e + j2;
                            const uint32_t own = (uint32_t) acc0[o * kN + j];
                            const int src = ((j - a) % (2 * kN) + 2 * kN) % (2 * kN);  // X^a * P: coefficient j comes from j - a
                            uint32_t r = (uint32_t) acc0[o * kN + (src & (kN - 1))];
                            if (src >= kN) r = 0u - r;
                            const uint32_t t = (rot ? r - own : own) + kDecompOffset;
                            const double want = (double) (int) ((t >> (32 - (q + 1) * kBgbit)) & 1023u) - 512.0;
                            const double got = e < 32 ? x[e].x : x[e - 32].y;
                            if (got != want) bad++;
                        }
                    }
            }
        }
        printf("B0. rotation + decomposition, all 2048 rotations x 2 levels: %ld wrong digits\n", bad);
        CHECK(bad == 0, "rotation/decomposition through the extended accumulator copy");
    }

    // random accumulator, several rotations, teacher forcing against the exact path
    std::vector<int32_t> acc(2 * kN);
    for (auto &v : acc) v = (int32_t) ((((uint32_t) rand()) << 16) ^ (uint32_t) rand() ^ (((uint32_t) rand()) << 31));
    int maxdiff = 0, ndiff = 0;
    const int rots[] = {1, 2047, 1024, 15, 16, 17, 1023, 1025, 777, 1300, 31, 2032};
    const std::vector<int32_t> acc0 = acc;
    for (int it = 0; it < 12; it++) {
        const int i = it % P.n;
        const int a = rots[it];
        for (int lane = 0; lane < 32; lane++) phase_load_acc(lane, *ws, acc.data());
        // warp pair: warp o owns accumulator polynomial o (rotation + decomposition of both digit levels)
        if (it == 0)  // afterwards phase_i2_final keeps the extended copy up to date
            for (int o = 0; o < 2; o++)
                for (int lane = 0; lane < 32; lane++) phase_ext_build(lane, *ws, o);
        static cpx x1[2][32][32];
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) {
                phase_f1_decomp(lane, *ws, o, a, true, x1[o][lane]);
                phase_f1_fft(x1[o][lane]);
            }
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_f1_store(lane, *ws, o, x1[o][lane]);
        cpx keep[2][32][16], give[2][32][16];
        memset(keep, 0, sizeof(keep));
        memset(give, 0, sizeof(give));
        for (int row = 0; row < kKpl; row++) {
            const int o = row >> 1;  // warp o owns rows 2o, 2o+1; keep = result polynomial o
            const cpx *bkrow = bkdev.data() + ((size_t) i * kKpl + row) * kBkRowCplx;
            for (int lane = 0; lane < 32; lane++) {
                cpx z[16];
                phase_f2_fft(lane, *ws, e2.data(), row, z);
                phase_mac_half(lane, z, bkrow + o * kBkHalfCplx, keep[o][lane]);
                phase_mac_half(lane, z, bkrow + (1 - o) * kBkHalfCplx, give[o][lane]);
            }
        }
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_xchg_store(lane, *ws, o, give[o][lane]);
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_xchg_load(lane, *ws, o, keep[o][lane]);
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_inv16_store(lane, *ws, e2.data(), o, keep[o][lane]);
        static cpx xh[2][32][16], snd[2][32][8];
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) {
                phase_i2_inner(lane, *ws, o, xh[o][lane]);
                phase_i2_send(lane, xh[o][lane], snd[o][lane]);
            }
        for (int o = 0; o < 2; o++)
            for (int lane = 0; lane < 32; lane++) phase_i2_final(lane, *ws, o, xh[o][lane], snd[o][lane ^ 16]);
        std::vector<int32_t> got(2 * kN);
        for (int lane = 0; lane < 32; lane++) phase_dump_acc(lane, *ws, got.data());
        // expected: acc + BK_i (.) ((X^a - 1) acc), exact
        std::vector<int32_t> tmp(2 * kN);
        for (int o = 0; o < 2; o++) oracle_mul_by_xai_minus_one(a, kN, acc.data() + o * kN, tmp.data() + o * kN);
        oracle_extern_mul_exact(&P, bk.data() + (size_t) i * kKpl * 2 * kN, tmp.data());
        for (int j = 0; j < 2 * kN; j++) {
            const int32_t e = (int32_t) ((uint32_t) acc[j] + (uint32_t) tmp[j]);
            const int d = (int) ((uint32_t) got[j] - (uint32_t) e);
            if (d != 0) ndiff++;
            if (abs(d) > maxdiff) maxdiff = abs(d);
        }
        acc = got;  // keep going from the emulator's own state
    }
    printf("B. MuxRotate vs exact: max |diff| = %d LSB, %d of %d words differ\n", maxdiff, ndiff, 12 * 2 * kN);
    printf("B. fp64 value vs nearest integer at the conversion: max %.4f, rms %.4f over %ld conversions\n", g_max_frac,
           sqrt(g_sum_frac2 / (double) (g_conversions ? g_conversions : 1)), g_conversions);
#if TFHE_B200_TRUNCATE_LIKE_REFERENCE
    CHECK(maxdiff <= 1, "MuxRotate step differs from exact result by more than truncation");
#else
    // round to nearest: the step IS the exact integer product (margin: 0.5 would flip a rounding)
    CHECK(maxdiff == 0, "MuxRotate step differs from the exact integer product");
    CHECK(g_max_frac < 0.25, "fp64 rounding error too close to 1/2");
#endif

    // ---- B3. the same steps through the LATENCY kernel's split (eight warps per ciphertext) ---------
    // warp (o, q): decomposition + pass 1 of row (o, q); pass 2, multiply and inverse pass 2 on eight warps by
    // class octets and position halves; inverse pass 1 by halves of the slices on warps (o, 0), (o, 1).
    for (int variant = 0; variant < 2; variant++) {   // 0: eight-warp kernel, 1: four-warp kernel (Fourier section by one warp)
        LatencySmem *qs = new LatencySmem();
        std::vector<int32_t> accq = acc0;
        int maxdiffq = 0;
        for (int it = 0; it < 12; it++) {
            const int i = it % P.n;
            const int a = rots[it];
            for (int lane = 0; lane < 32; lane++) phase_load_acc_p(lane, 32, qs->acc, accq.data());
            if (it == 0)
                for (int o = 0; o < 2; o++)
                    for (int lane = 0; lane < 32; lane++) phase_ext_build_p(lane, qs->acc[o], qs->ext[o]);
            // pass 1 split over lane pairs: lane (hh, j2) transforms half a slice, stage 0 through lane ^ 16
            static cpx xq[4][32][16], wq[4][32][16];
            for (int w = 0; w < 4; w++)
                for (int lane = 0; lane < 32; lane++) {
                    phase_f1h_decomp_p(lane >> 4, lane & 15, w & 1, qs->acc[w >> 1], qs->ext[w >> 1], a, true, xq[w][lane]);
                    phase_f1h_cross_send(lane >> 4, xq[w][lane], wq[w][lane]);
                }
            for (int w = 0; w < 4; w++)
                for (int lane = 0; lane < 32; lane++) {
                    phase_f1h_finish(lane >> 4, wq[w][lane], wq[w][lane ^ 16], xq[w][lane]);
                    phase_f1h_store_p(lane >> 4, lane & 15, qs->exch[w], xq[w][lane]);
                }
            // pass 2 by warp cq; Fourier multiply + inverse pass 2 on eight warps: warp v = (class octet v & 3,
            // position half v >> 2); transpositions through the exchange / inverse buffers
            static cpx x8[4][32][8];
            if (variant == 1) {
                for (int w = 0; w < 4; w++) {
                    for (int lane = 0; lane < 32; lane++) phase_c_f2_inplace(lane >> 3, 8 * w + (lane & 7), qs->exch, e2.data());
                    for (int lane = 0; lane < 32; lane++) {
                        const int g = lane >> 3, m1 = 8 * w + (lane & 7);
                        cpx zr[kKpl][4], accv[2][4];
                        memset(accv, 0, sizeof(accv));
                        phase_w_load_rows(g, m1, qs->exch, zr);
                        for (int row = 0; row < kKpl; row++) {
                            const cpx *bkrow = bkdev.data() + ((size_t) i * kKpl + row) * kBkRowCplx;
                            for (int oo = 0; oo < 2; oo++)
                                for (int p4 = 0; p4 < 4; p4++)
                                    cmac(accv[oo][p4], zr[row][p4], bkrow[oo * kBkHalfCplx + (4 * g + p4) * 32 + m1]);
                        }
                        cpx c3, c2;
                        phase_w_inv_consts(g, e2.data() + m1 * kE2Row, c3, c2);
                        for (int oo = 0; oo < 2; oo++) phase_w_inv_a_store(g, m1, qs->inv[oo], c3, c2, accv[oo]);
                    }
                    for (int lane = 0; lane < 32; lane++) {
                        const int rr = lane >> 3, m1 = 8 * w + (lane & 7);
                        phase_w_inv_b_inplace(rr & 1, m1, qs->inv[rr >> 1], e2[m1 * kE2Row + 1], e2[m1 * kE2Row]);
                    }
                }
            } else
            {
                for (int w = 0; w < 4; w++)   // pass 2 by warp w = class octet w, in place
                    for (int lane = 0; lane < 32; lane++) phase_c_f2_inplace(lane >> 3, 8 * w + (lane & 7), qs->exch, e2.data());
                for (int v = 0; v < 8; v++) {   // multiply by position pairs: lane (g4, c), both result polynomials
                    static cpx accp[32][2][2], snd[32][2][2];
                    for (int lane = 0; lane < 32; lane++) {
                        const int g4 = lane >> 3, ph = v >> 2, m1 = 8 * (v & 3) + (lane & 7), p0 = 8 * ph + 2 * g4;
                        cpx zr[kKpl][2];
                        memset(accp[lane], 0, sizeof(accp[lane]));
                        phase_p_load_rows(p0, m1, qs->exch, zr);
                        for (int row = 0; row < kKpl; row++) {
                            const cpx *bkrow = bkdev.data() + ((size_t) i * kKpl + row) * kBkRowCplx;
                            for (int oo = 0; oo < 2; oo++)
                                for (int e = 0; e < 2; e++)
                                    cmac(accp[lane][oo][e], zr[row][e], bkrow[oo * kBkHalfCplx + (p0 + e) * 32 + m1]);
                        }
                        cpx c3, k2;
                        phase_p_inv_consts(ph, g4, e2.data() + m1 * kE2Row, c3, k2);
                        for (int oo = 0; oo < 2; oo++) bf_inv(accp[lane][oo][0], accp[lane][oo][1], c3.x, c3.y);
                        memcpy(snd[lane], accp[lane], sizeof(snd[lane]));
                    }
                    for (int lane = 0; lane < 32; lane++) {
                        const int g4 = lane >> 3, ph = v >> 2, m1 = 8 * (v & 3) + (lane & 7), p0 = 8 * ph + 2 * g4;
                        cpx c3, k2;
                        phase_p_inv_consts(ph, g4, e2.data() + m1 * kE2Row, c3, k2);
                        for (int oo = 0; oo < 2; oo++) {
                            for (int e = 0; e < 2; e++) phase_p_inv_cross(g4 & 1, k2, snd[lane ^ 8][oo][e], accp[lane][oo][e]);
                            phase_p_inv_store(p0, m1, qs->inv[oo], accp[lane][oo]);
                        }
                    }
                }
                for (int v = 0; v < 8; v++)   // (pair barrier)
                    for (int lane = 0; lane < 32; lane++) {
                        const int kk = lane >> 4, oo = (lane >> 3) & 1, m1 = 8 * (v & 3) + (lane & 7);
                        phase_o_inv_b_inplace(2 * (v >> 2) + kk, m1, qs->inv[oo], e2[m1 * kE2Row + 1], e2[m1 * kE2Row]);
                    }
            }
            for (int w = 0; w < 4; w++) {
                const int o = w >> 1, h = w & 1;
                for (int lane = 0; lane < 32; lane++) phase_q_i2_local(lane, h, qs->inv[o], x8[w][lane]);
                static cpx xc[32][8];
                memcpy(xc, x8[w], sizeof(xc));
                for (int lane = 0; lane < 32; lane++)   // stage 1: lane ^ 8, blocks of 16: multiplier 1 + (qq >> 1)
                    phase_q_i2_cross(((lane >> 3) & 1) != 0, 1 + (lane >> 4), xc[lane ^ 8], x8[w][lane]);
                memcpy(xc, x8[w], sizeof(xc));
                for (int lane = 0; lane < 32; lane++)   // stage 0: lane ^ 16, multiplier 0
                    phase_q_i2_cross((lane >> 4) != 0, 0, xc[lane ^ 16], x8[w][lane]);
                for (int lane = 0; lane < 32; lane++) phase_q_final(lane, h, qs->acc[o], qs->ext[o], x8[w][lane]);
            }
            std::vector<int32_t> got(2 * kN);
            for (int lane = 0; lane < 32; lane++) phase_dump_acc_p(lane, 32, qs->acc, got.data());
            std::vector<int32_t> tmp(2 * kN);
            for (int o = 0; o < 2; o++) oracle_mul_by_xai_minus_one(a, kN, accq.data() + o * kN, tmp.data() + o * kN);
            oracle_extern_mul_exact(&P, bk.data() + (size_t) i * kKpl * 2 * kN, tmp.data());
            for (int j = 0; j < 2 * kN; j++) {
                const int32_t e = (int32_t) ((uint32_t) accq[j] + (uint32_t) tmp[j]);
                const int d = abs((int) ((uint32_t) got[j] - (uint32_t) e));
                if (d > maxdiffq) maxdiffq = d;
            }
            accq = got;
        }
        printf("B3.%d latency-kernel split (%s warps per ciphertext) vs exact: max |diff| = %d LSB\n", variant,
               variant ? "four" : "eight", maxdiffq);
#if TFHE_B200_TRUNCATE_LIKE_REFERENCE
        CHECK(maxdiffq <= 1, "latency-kernel split differs from exact result by more than truncation");
#else
        CHECK(maxdiffq == 0, "latency-kernel split differs from the exact integer product");
#endif
        std::vector<int32_t> u(kN + 1), u2(kN + 1);
        for (int lane = 0; lane < 256; lane++) phase_extract_p(lane, 256, qs->acc, u.data());
        for (int lane = 0; lane < 32; lane++) phase_load_acc(lane, *ws, accq.data());
        for (int lane = 0; lane < 32; lane++) phase_extract(lane, *ws, u2.data());
        CHECK(u == u2, "extraction by 256 lanes");
        delete qs;
    }

    // ---- C. extraction ----------------------------------------------------
    {
        for (int lane = 0; lane < 32; lane++) phase_load_acc(lane, *ws, acc.data());
        std::vector<int32_t> u(kN + 1);
        for (int lane = 0; lane < 32; lane++) phase_extract(lane, *ws, u.data());
        bool ok = u[0] == acc[0] && u[kN] == acc[kN];
        for (int j = 1; j < kN; j++) ok = ok && u[j] == (int32_t) (0u - (uint32_t) acc[kN - j]);
        CHECK(ok, "extraction");
        printf("C. extraction %s\n", ok ? "ok" : "BAD");
    }

    delete ws;
    printf(fails ? "HOST EMULATION FAILED (%d)\n" : "HOST EMULATION OK\n", fails);
    return fails ? 1 : 0;
}
