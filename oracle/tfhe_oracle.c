/*
 * TEST INFRASTRUCTURE — not part of the product.  See tfhe_oracle.h.
 *
 * Plain-C restatement of the reference's bootstrapped-gate path.  Compiled
 * with -ffp-contract=off so that, in ORACLE_FFT_REF mode, every floating-point
 * operation happens in the same order and with the same rounding as the
 * reference's host code built in oracle/_ref (both call the same transform in
 * oracle/fftw_shim).  Paths in the comments are relative to
 * /root/reference/gpuParallel.
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "fftw_shim/fftw3.h"

/* ------------------------------------------------------------------ params */

/* tfhe_gate_bootstrapping.cu:25-49 */
void oracle_default_params(OracleParams *p) {
    p->N = 1024;
    p->k = 1;
    p->n = 500;
    p->l = 2;
    p->Bgbit = 10;
    p->ks_basebit = 2;
    p->ks_t = 8;
    p->alpha_lwe = pow(2., -15) * sqrt(2. / M_PI);
    p->alpha_bk = 9.e-9 * sqrt(2. / M_PI);
}

size_t oracle_bk_words(const OracleParams *p) {
    return (size_t) p->n * (p->k + 1) * p->l * (p->k + 1) * p->N;
}

size_t oracle_ks_words(const OracleParams *p) {
    return (size_t) p->N * p->k * p->ks_t * (1 << p->ks_basebit) * (p->n + 1);
}

/* --------------------------------------------------------------------- rng */
/* The reference draws from std::default_random_engine (numeric-functions.cu:11-13),
 * which is implementation-defined; the oracle uses xoshiro256** so that its
 * fixtures are reproducible anywhere. */

static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static uint64_t rng_next(OracleRng *r) {
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

void oracle_rng_seed(OracleRng *r, uint64_t seed) {
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
    r->has_spare = 0;
    r->spare = 0.;
}

int32_t oracle_rng_torus(OracleRng *r) { return (int32_t) (rng_next(r) >> 32); }

static int rng_bit(OracleRng *r) { return (int) (rng_next(r) >> 63); }

double oracle_rng_gauss(OracleRng *r, double sigma) {
    if (r->has_spare) {
        r->has_spare = 0;
        return r->spare * sigma;
    }
    double u1, u2;
    do {
        u1 = (double) (rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
    } while (u1 <= 0.);
    u2 = (double) (rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2. * log(u1));
    r->spare = rad * sin(2. * M_PI * u2);
    r->has_spare = 1;
    return rad * cos(2. * M_PI * u2) * sigma;
}

/* ----------------------------------------------------------------- numeric */

/* numeric-functions.cu:33-35 */
int32_t oracle_dtot32(double d) {
    return (int32_t) (int64_t) ((d - (double) (int64_t) d) * 4294967296.);
}

/* numeric-functions.cu:60-66 */
int oracle_modswitch_from(int32_t phase, int msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t) msize) * 2;
    uint64_t half_interval = interv / 2;
    uint64_t phase64 = (((uint64_t) (int64_t) phase) << 32) + half_interval;
    return (int) (phase64 / interv);
}

/* numeric-functions.cu:72-77 */
int32_t oracle_modswitch_to(int mu, int msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t) msize) * 2;
    uint64_t phase64 = (uint64_t) (int64_t) mu * interv;
    return (int32_t) (phase64 >> 32);
}

/* -------------------------------------------------------------- LWE basics */

/* lwe-functions.cu:36-47 */
void oracle_lwe_encrypt(OracleRng *r, const int32_t *key, int n, int32_t mu, double alpha,
                        int32_t *out) {
    uint32_t b = (uint32_t) mu + (uint32_t) oracle_dtot32(oracle_rng_gauss(r, alpha));
    for (int i = 0; i < n; i++) {
        out[i] = oracle_rng_torus(r);
        b += (uint32_t) out[i] * (uint32_t) key[i];
    }
    out[n] = (int32_t) b;
}

/* lwe-functions.cu:72-81 */
int32_t oracle_lwe_phase(const int32_t *sample, const int32_t *key, int n) {
    uint32_t axs = 0;
    for (int i = 0; i < n; i++) axs += (uint32_t) sample[i] * (uint32_t) key[i];
    return (int32_t) ((uint32_t) sample[n] - axs);
}

/* tfhe_gate_bootstrapping.cu:114-119 */
void oracle_encrypt_bit(const OracleParams *p, OracleRng *r, const int32_t *lwe_key, int bit,
                        int32_t *out) {
    const int32_t mu = oracle_modswitch_to(1, 8);
    oracle_lwe_encrypt(r, lwe_key, p->n, bit ? mu : -mu, p->alpha_lwe, out);
}

/* tfhe_gate_bootstrapping.cu:122-125 */
int oracle_decrypt_bit(const OracleParams *p, const int32_t *lwe_key, const int32_t *sample) {
    return oracle_lwe_phase(sample, lwe_key, p->n) > 0 ? 1 : 0;
}

/* -------------------------------------------------------------- polynomials */

/* toruspolynomial-functions.cu:492-519 */
void oracle_mul_by_xai(int a, int N, const int32_t *in, int32_t *out) {
    const uint32_t *u = (const uint32_t *) in;
    uint32_t *o = (uint32_t *) out;
    if (a < N) {
        for (int i = 0; i < a; i++) o[i] = 0u - u[i - a + N];
        for (int i = a; i < N; i++) o[i] = u[i - a];
    } else {
        const int aa = a - N;
        for (int i = 0; i < aa; i++) o[i] = u[i - aa + N];
        for (int i = aa; i < N; i++) o[i] = 0u - u[i - aa];
    }
}

/* toruspolynomial-functions.cu:191-213 */
void oracle_mul_by_xai_minus_one(int a, int N, const int32_t *in, int32_t *out) {
    const uint32_t *u = (const uint32_t *) in;
    uint32_t *o = (uint32_t *) out;
    if (a < N) {
        for (int i = 0; i < a; i++) o[i] = 0u - u[i - a + N] - u[i];
        for (int i = a; i < N; i++) o[i] = u[i - a] - u[i];
    } else {
        const int aa = a - N;
        for (int i = 0; i < aa; i++) o[i] = u[i - aa + N] - u[i];
        for (int i = aa; i < N; i++) o[i] = 0u - u[i - aa] - u[i];
    }
}

/* tgsw.cu:19-27 : offset = Bg/2 * sum_i 2^(32-(i+1)Bgbit) */
static uint32_t decomp_offset(const OracleParams *p) {
    uint32_t t = 0;
    for (int i = 0; i < p->l; i++) t += 1u << (32 - (i + 1) * p->Bgbit);
    return t * (uint32_t) ((1 << p->Bgbit) / 2);
}

/* tgsw-functions.cu:301-352 (scalar path) */
void oracle_decomp(const OracleParams *p, const int32_t *poly, int32_t *out) {
    const int N = p->N;
    const uint32_t mask = (uint32_t) ((1 << p->Bgbit) - 1);
    const int32_t half = (1 << p->Bgbit) / 2;
    const uint32_t off = decomp_offset(p);
    for (int q = 0; q < p->l; q++) {
        const int decal = 32 - (q + 1) * p->Bgbit;
        for (int j = 0; j < N; j++) {
            const uint32_t v = (uint32_t) poly[j] + off;
            out[q * N + j] = (int32_t) ((v >> decal) & mask) - half;
        }
    }
}

/* ------------------------------------------------------------------ keygen */

/* b += key * a  (negacyclic, key binary) — exact integer version of the
 * torusPolynomialAddMulR call in tlwe-functions.cu:26-39 */
static void addmul_binary_key(uint32_t *b, const int32_t *key, const uint32_t *a, int N) {
    for (int s = 0; s < N; s++) {
        if (!key[s]) continue;
        for (int i = 0; i < s; i++) b[i] -= a[i - s + N];
        for (int i = s; i < N; i++) b[i] += a[i - s];
    }
}

void oracle_keygen(const OracleParams *p, uint64_t seed, int32_t *lwe_key, int32_t *tlwe_key,
                   int32_t *bk, int32_t *ks) {
    const int n = p->n, N = p->N, k = p->k, l = p->l, kpl = (k + 1) * l;
    const int t = p->ks_t, basebit = p->ks_basebit, base = 1 << basebit;
    OracleRng rng;
    oracle_rng_seed(&rng, seed);

    /* lwe-functions.cu:21-27, tlwe-functions.cu:15-23 */
    for (int i = 0; i < n; i++) lwe_key[i] = rng_bit(&rng);
    for (int i = 0; i < k * N; i++) tlwe_key[i] = rng_bit(&rng);

    /* key-switch key, extracted key (lwe.cu:287-296) -> lwe key:
     * lwe-keyswitch-functions.cu:890-942 */
    {
        const int nin = k * N;
        const int sizeks = nin * t * (base - 1);
        double *noise = (double *) malloc(sizeof(double) * (size_t) sizeks);
        double err = 0;
        for (int i = 0; i < sizeks; i++) {
            noise[i] = oracle_rng_gauss(&rng, p->alpha_lwe);
            err += noise[i];
        }
        err = err / sizeks;
        for (int i = 0; i < sizeks; i++) noise[i] -= err;
        int index = 0;
        for (int i = 0; i < nin; i++)
            for (int j = 0; j < t; j++) {
                int32_t *row0 = ks + (((size_t) i * t + j) * base + 0) * (n + 1);
                memset(row0, 0, sizeof(int32_t) * (size_t) (n + 1));
                for (int h = 1; h < base; h++) {
                    int32_t *row = ks + (((size_t) i * t + j) * base + h) * (n + 1);
                    const uint32_t mess = (uint32_t) (tlwe_key[i] * h) * (1u << (32 - (j + 1) * basebit));
                    uint32_t b = mess + (uint32_t) oracle_dtot32(noise[index]);
                    for (int c = 0; c < n; c++) {
                        row[c] = oracle_rng_torus(&rng);
                        b += (uint32_t) row[c] * (uint32_t) lwe_key[c];
                    }
                    row[n] = (int32_t) b;
                    index++;
                }
            }
        free(noise);
    }

    /* bootstrapping key: lwe-bootstrapping-functions.cu:206-215 ->
     * tGswSymEncryptInt (tgsw-functions.cu:191-194) = tGswEncryptZero (:129-136,
     * each row tLweSymEncryptZero tlwe-functions.cu:26-39) + tGswAddMuIntH (:113-124) */
    for (int i = 0; i < n; i++) {
        for (int r = 0; r < kpl; r++) {
            uint32_t *row = (uint32_t *) bk + ((size_t) i * kpl + r) * (k + 1) * N;
            uint32_t *b = row + (size_t) k * N;
            for (int j = 0; j < N; j++) b[j] = (uint32_t) oracle_dtot32(oracle_rng_gauss(&rng, p->alpha_bk));
            for (int m = 0; m < k; m++) {
                uint32_t *a = row + (size_t) m * N;
                for (int j = 0; j < N; j++) a[j] = (uint32_t) oracle_rng_torus(&rng);
                addmul_binary_key(b, tlwe_key + (size_t) m * N, a, N);
            }
        }
        for (int bloc = 0; bloc <= k; bloc++)
            for (int q = 0; q < l; q++) {
                uint32_t *row = (uint32_t *) bk + ((size_t) i * kpl + bloc * l + q) * (k + 1) * N;
                row[(size_t) bloc * N] += (uint32_t) lwe_key[i] * (1u << (32 - (q + 1) * p->Bgbit));
            }
    }
}

/* ----------------------------------------------------------------- context */

typedef struct { double re, im; } cpx;

struct OracleCtx {
    OracleParams p;
    int fft_mode;
    int Ns2;
    cpx *bkfft;          /* [n][kpl][k+1][N/2] */
    const int32_t *ks;   /* borrowed */
    /* folded-mode tables */
    cpx *twist;          /* exp(+i*pi*j/N), j < N/2          */
    cpx *tw_pos;         /* exp(+2*pi*i*k/(N/2))             */
    cpx *tw_neg;         /* exp(-2*pi*i*k/(N/2))             */
};

/* per-call scratch (one per thread) */
typedef struct {
    const OracleCtx *c;
    fftw_plan rev_p, dir_p;      /* REF mode: fft_processor_fftw.cu:140-141 */
    double *rev_in, *out;        /* 2N reals                               */
    cpx *rev_out, *in;           /* N+1 complex                            */
    cpx *w0, *w1;                /* folded work arrays                      */
    int32_t *deca;               /* [kpl][N]                               */
    cpx *decafft;                /* [kpl][N/2]                             */
    cpx *tmpa;                   /* [k+1][N/2]                             */
    int32_t *tmp_acc;            /* [k+1][N]                               */
} Work;

static Work *work_new(const OracleCtx *c) {
    const int N = c->p.N, k = c->p.k, kpl = (k + 1) * c->p.l;
    Work *w = (Work *) calloc(1, sizeof(Work));
    w->c = c;
    w->rev_in = (double *) fftw_malloc(sizeof(double) * 2 * N);
    w->out = (double *) fftw_malloc(sizeof(double) * 2 * N);
    w->rev_out = (cpx *) fftw_malloc(sizeof(cpx) * (N + 1));
    w->in = (cpx *) fftw_malloc(sizeof(cpx) * (N + 1));
    if (c->fft_mode == ORACLE_FFT_REF) {
        w->rev_p = fftw_plan_dft_r2c_1d(2 * N, w->rev_in, (fftw_complex *) w->rev_out, FFTW_ESTIMATE);
        w->dir_p = fftw_plan_dft_c2r_1d(2 * N, (fftw_complex *) w->in, w->out, FFTW_ESTIMATE);
    }
    w->w0 = (cpx *) fftw_malloc(sizeof(cpx) * N);
    w->w1 = (cpx *) fftw_malloc(sizeof(cpx) * N);
    w->deca = (int32_t *) malloc(sizeof(int32_t) * (size_t) kpl * N);
    w->decafft = (cpx *) fftw_malloc(sizeof(cpx) * (size_t) kpl * (N / 2));
    w->tmpa = (cpx *) fftw_malloc(sizeof(cpx) * (size_t) (k + 1) * (N / 2));
    w->tmp_acc = (int32_t *) malloc(sizeof(int32_t) * (size_t) (k + 1) * N);
    return w;
}

static void work_free(Work *w) {
    if (w->rev_p) fftw_destroy_plan(w->rev_p);
    if (w->dir_p) fftw_destroy_plan(w->dir_p);
    fftw_free(w->rev_in); fftw_free(w->out); fftw_free(w->rev_out); fftw_free(w->in);
    fftw_free(w->w0); fftw_free(w->w1);
    free(w->deca); fftw_free(w->decafft); fftw_free(w->tmpa); free(w->tmp_acc);
    free(w);
}

/* coefficients (as doubles, already scaled) -> Lagrange half-complex */
static void to_lagrange(Work *w, const int32_t *a, double scale_ref, double scale_folded, cpx *res) {
    const OracleCtx *c = w->c;
    const int N = c->p.N, Ns2 = N / 2;
    if (c->fft_mode == ORACLE_FFT_REF) {
        /* fft_processor_fftw.cu:148-167 */
        if (scale_ref == 0.5) for (int i = 0; i < N; i++) w->rev_in[i] = a[i] / 2.;
        else for (int i = 0; i < N; i++) w->rev_in[i] = a[i] * scale_ref;
        for (int i = 0; i < N; i++) w->rev_in[N + i] = -w->rev_in[i];
        fftw_execute(w->rev_p);
        for (int i = 0; i < Ns2; i++) res[i] = w->rev_out[2 * i + 1];
    } else {
        /* P(zeta^(4m+1)), zeta = exp(i*pi/N): fold, twist, N/2-point DFT (+ sign) */
        for (int j = 0; j < Ns2; j++) {
            const double re = a[j] * scale_folded, im = a[j + Ns2] * scale_folded;
            const cpx z = c->twist[j];
            w->w0[j].re = re * z.re - im * z.im;
            w->w0[j].im = re * z.im + im * z.re;
        }
        cpx *r = (cpx *) fftw_shim_cfft(Ns2, (const double *) c->tw_pos, 1, (double *) w->w0, (double *) w->w1);
        memcpy(res, r, sizeof(cpx) * (size_t) Ns2);
    }
}

/* Lagrange half-complex -> torus coefficients; fft_processor_fftw.cu:168-181 */
static void from_lagrange(Work *w, const cpx *a, int32_t *res) {
    const OracleCtx *c = w->c;
    const int N = c->p.N, Ns2 = N / 2;
    static const double _2p32 = 4294967296.;
    if (c->fft_mode == ORACLE_FFT_REF) {
        const double _1sN = (double) 1 / (double) N;
        for (int i = 0; i <= Ns2; i++) { w->in[2 * i].re = 0; w->in[2 * i].im = 0; }
        for (int i = 0; i < Ns2; i++) w->in[2 * i + 1] = a[i];
        fftw_execute(w->dir_p);
        for (int i = 0; i < N; i++) res[i] = (int32_t) (int64_t) (w->out[i] * _1sN * _2p32);
    } else {
        memcpy(w->w0, a, sizeof(cpx) * (size_t) Ns2);
        cpx *r = (cpx *) fftw_shim_cfft(Ns2, (const double *) c->tw_neg, 0, (double *) w->w0, (double *) w->w1);
        const double sc = _2p32 / (double) Ns2;
        for (int j = 0; j < Ns2; j++) {
            const cpx z = c->twist[j];
            const double re = r[j].re * z.re + r[j].im * z.im;   /* times conj(twist) */
            const double im = r[j].im * z.re - r[j].re * z.im;
            res[j] = (int32_t) (int64_t) (re * sc);
            res[j + Ns2] = (int32_t) (int64_t) (im * sc);
        }
    }
}

OracleCtx *oracle_ctx_new(const OracleParams *p, const int32_t *bk, const int32_t *ks, int fft_mode) {
    OracleCtx *c = (OracleCtx *) calloc(1, sizeof(OracleCtx));
    c->p = *p;
    c->fft_mode = fft_mode;
    c->Ns2 = p->N / 2;
    c->ks = ks;
    const int N = p->N, Ns2 = N / 2, k = p->k, kpl = (k + 1) * p->l;
    c->twist = (cpx *) fftw_malloc(sizeof(cpx) * Ns2);
    c->tw_pos = (cpx *) fftw_malloc(sizeof(cpx) * Ns2);
    c->tw_neg = (cpx *) fftw_malloc(sizeof(cpx) * Ns2);
    for (int j = 0; j < Ns2; j++) {
        c->twist[j].re = cos(M_PI * j / N);
        c->twist[j].im = sin(M_PI * j / N);
        c->tw_pos[j].re = cos(2. * M_PI * j / Ns2);
        c->tw_pos[j].im = sin(2. * M_PI * j / Ns2);
        c->tw_neg[j].re = c->tw_pos[j].re;
        c->tw_neg[j].im = -c->tw_pos[j].im;
    }
    /* init_LweBootstrappingKeyFFT: lwe-bootstrapping-functions-fft.cu:60-89 ->
     * tGswToFFTConvert tgsw-fft-operations.cu:84 -> TorusPolynomial_ifft (scale 2^-33, doubled) */
    const size_t npoly = (size_t) p->n * kpl * (k + 1);
    c->bkfft = (cpx *) fftw_malloc(sizeof(cpx) * npoly * Ns2);
    if (bk) {
        Work *w = work_new(c);
        const double _2pm33 = 1. / (double) (INT64_C(1) << 33);
        const double _2pm32 = 1. / (double) (INT64_C(1) << 32);
        for (size_t q = 0; q < npoly; q++)
            to_lagrange(w, bk + q * N, _2pm33, _2pm32, c->bkfft + q * Ns2);
        work_free(w);
    }
    return c;
}

void oracle_ctx_free(OracleCtx *c) {
    if (!c) return;
    fftw_free(c->twist); fftw_free(c->tw_pos); fftw_free(c->tw_neg); fftw_free(c->bkfft);
    free(c);
}

const OracleParams *oracle_ctx_params(const OracleCtx *c) { return &c->p; }
const double *oracle_ctx_bkfft(const OracleCtx *c) { return (const double *) c->bkfft; }

void oracle_ifft_int(const OracleCtx *c, const int32_t *poly, double *out) {
    Work *w = work_new(c);
    to_lagrange(w, poly, 0.5, 1.0, (cpx *) out);
    work_free(w);
}

void oracle_ifft_torus(const OracleCtx *c, const int32_t *poly, double *out) {
    Work *w = work_new(c);
    to_lagrange(w, poly, 1. / (double) (INT64_C(1) << 33), 1. / (double) (INT64_C(1) << 32), (cpx *) out);
    work_free(w);
}

void oracle_fft_torus(const OracleCtx *c, const double *in, int32_t *poly_out) {
    Work *w = work_new(c);
    from_lagrange(w, (const cpx *) in, poly_out);
    work_free(w);
}

/* --------------------------------------------------------- external product */

/* tgsw-fft-operations.cu:124-264 */
static void extern_mul_w(Work *w, int bk_index, int32_t *accum) {
    const OracleCtx *c = w->c;
    const OracleParams *p = &c->p;
    const int N = p->N, Ns2 = N / 2, k = p->k, l = p->l, kpl = (k + 1) * l;
    /* :148-150 decomposition of each accumulator polynomial */
    for (int i = 0; i <= k; i++) oracle_decomp(p, accum + (size_t) i * N, w->deca + (size_t) i * l * N);
    /* :151-153 */
    for (int q = 0; q < kpl; q++) to_lagrange(w, w->deca + (size_t) q * N, 0.5, 1.0, w->decafft + (size_t) q * Ns2);
    /* :154 tLweFFTClear */
    for (int i = 0; i < (k + 1) * Ns2; i++) { w->tmpa[i].re = 0; w->tmpa[i].im = 0; }
    /* :156-158 tLweFFTAddMulRTo (tlwe-fft-operations.cu:286) -> LagrangeHalfCPolynomialAddMul
     * (lagrangehalfc_impl.cu:95-117): rr[i] += aa[i]*bb[i] */
    const cpx *gsw = c->bkfft + (size_t) bk_index * kpl * (k + 1) * Ns2;
    for (int q = 0; q < kpl; q++) {
        const cpx *aa = w->decafft + (size_t) q * Ns2;
        for (int i = 0; i <= k; i++) {
            const cpx *bb = gsw + ((size_t) q * (k + 1) + i) * Ns2;
            cpx *rr = w->tmpa + (size_t) i * Ns2;
            for (int j = 0; j < Ns2; j++) {
                const double tr = aa[j].re * bb[j].re - aa[j].im * bb[j].im;
                const double ti = aa[j].re * bb[j].im + aa[j].im * bb[j].re;
                rr[j].re += tr;
                rr[j].im += ti;
            }
        }
    }
    /* :163 tLweFromFFTConvert (tlwe-fft-operations.cu:72) */
    for (int i = 0; i <= k; i++) from_lagrange(w, w->tmpa + (size_t) i * Ns2, accum + (size_t) i * N);
}

void oracle_extern_mul(const OracleCtx *c, int bk_index, int32_t *accum) {
    Work *w = work_new(c);
    extern_mul_w(w, bk_index, accum);
    work_free(w);
}

/* Exact integer external product: decomposition + naive negacyclic products
 * (multiplication.cu:53-77 torusPolynomialMultNaive; tgsw-functions.cu:156-170) */
void oracle_extern_mul_exact(const OracleParams *p, const int32_t *bk_i, int32_t *accum) {
    const int N = p->N, k = p->k, l = p->l, kpl = (k + 1) * l;
    int32_t *deca = (int32_t *) malloc(sizeof(int32_t) * (size_t) kpl * N);
    uint32_t *res = (uint32_t *) calloc((size_t) (k + 1) * N, sizeof(uint32_t));
    for (int i = 0; i <= k; i++) oracle_decomp(p, accum + (size_t) i * N, deca + (size_t) i * l * N);
    for (int q = 0; q < kpl; q++)
        for (int i = 0; i <= k; i++) {
            const uint32_t *b = (const uint32_t *) bk_i + ((size_t) q * (k + 1) + i) * N;
            const int32_t *d = deca + (size_t) q * N;
            uint32_t *r = res + (size_t) i * N;
            for (int x = 0; x < N; x++) {
                const uint32_t dx = (uint32_t) d[x];
                if (!dx) continue;
                for (int y = 0; y < N - x; y++) r[x + y] += dx * b[y];
                for (int y = N - x; y < N; y++) r[x + y - N] -= dx * b[y];
            }
        }
    memcpy(accum, res, sizeof(uint32_t) * (size_t) (k + 1) * N);
    free(deca);
    free(res);
}

/* ------------------------------------------------------------ blind rotation */

/* lwe-bootstrapping-functions-fft.cu:105-185 (MuxRotate) and :676-737 (loop) */
static void blind_rotate_w(Work *w, int32_t *accum, const int32_t *bara, int n_iter) {
    const OracleParams *p = &w->c->p;
    const int N = p->N, k = p->k;
    for (int i = 0; i < n_iter; i++) {
        const int barai = bara[i];
        if (barai == 0) continue;                                  /* :705 */
        for (int j = 0; j <= k; j++)                               /* tLweMulByXaiMinusOne tlwe-functions.cu:334 */
            oracle_mul_by_xai_minus_one(barai, N, accum + (size_t) j * N, w->tmp_acc + (size_t) j * N);
        extern_mul_w(w, i, w->tmp_acc);                            /* tGswFFTExternMulToTLwe */
        for (int j = 0; j < (k + 1) * N; j++)                      /* tLweAddTo tlwe-functions.cu:170 */
            accum[j] = (int32_t) ((uint32_t) accum[j] + (uint32_t) w->tmp_acc[j]);
    }
}

void oracle_blind_rotate(const OracleCtx *c, int32_t *accum, const int32_t *bara, int n_iter) {
    Work *w = work_new(c);
    blind_rotate_w(w, accum, bara, n_iter);
    work_free(w);
}

/* lwe-bootstrapping-functions-fft.cu:1408-1456 ; extraction lwe.cu:41-56 (index 0) */
static void blind_rotate_and_extract_w(Work *w, const int32_t *testvect, int barb, const int32_t *bara,
                                       int n_iter, int32_t *u) {
    const OracleParams *p = &w->c->p;
    const int N = p->N, k = p->k;
    int32_t *acc = (int32_t *) calloc((size_t) (k + 1) * N, sizeof(int32_t));
    if (barb != 0) oracle_mul_by_xai(2 * N - barb, N, testvect, acc + (size_t) k * N);
    else memcpy(acc + (size_t) k * N, testvect, sizeof(int32_t) * (size_t) N);
    blind_rotate_w(w, acc, bara, n_iter);
    for (int i = 0; i < k; i++) {
        u[i * N] = acc[(size_t) i * N];
        for (int j = 1; j < N; j++) u[i * N + j] = (int32_t) (0u - (uint32_t) acc[(size_t) i * N + N - j]);
    }
    u[k * N] = acc[(size_t) k * N];
    free(acc);
}

void oracle_blind_rotate_and_extract(const OracleCtx *c, const int32_t *testvect, int barb,
                                     const int32_t *bara, int n_iter, int32_t *u_out) {
    Work *w = work_new(c);
    blind_rotate_and_extract_w(w, testvect, barb, bara, n_iter, u_out);
    work_free(w);
}

/* lwe-bootstrapping-functions-fft.cu:1834-1870 */
static void bootstrap_woks_w(Work *w, int32_t mu, const int32_t *x, int32_t *u) {
    const OracleParams *p = &w->c->p;
    const int N = p->N, n = p->n;
    int32_t *testvect = (int32_t *) malloc(sizeof(int32_t) * (size_t) N);
    int32_t *bara = (int32_t *) malloc(sizeof(int32_t) * (size_t) n);
    const int barb = oracle_modswitch_from(x[n], 2 * N);
    for (int i = 0; i < n; i++) bara[i] = oracle_modswitch_from(x[i], 2 * N);
    for (int i = 0; i < N; i++) testvect[i] = mu;
    blind_rotate_and_extract_w(w, testvect, barb, bara, n, u);
    free(bara);
    free(testvect);
}

void oracle_bootstrap_woks(const OracleCtx *c, int32_t mu, const int32_t *x, int32_t *u_out) {
    Work *w = work_new(c);
    bootstrap_woks_w(w, mu, x, u_out);
    work_free(w);
}

/* ------------------------------------------------- exact (FFT-free) bootstrap
 * The reference's non-FFT path (tfhe_MuxRotate / tfhe_blindRotate / tfhe_blindRotateAndExtract /
 * tfhe_bootstrap_woKS, lwe-bootstrapping-functions.cu:34-179; tGswExternMulToTLwe
 * tgsw-functions.cu:156-170) with every polynomial product taken by definition
 * (torusPolynomialMultNaive, multiplication.cu:53-77): integer arithmetic mod 2^32 only, no
 * rounding anywhere.  The CUDA path rounds its fp64 products to the nearest integer, so a complete
 * blind rotation must reproduce these words exactly (tests/test_gpu_exact.py). */
void oracle_blind_rotate_exact(const OracleParams *p, const int32_t *bk, int32_t *accum, const int32_t *bara,
                               int n_iter) {
    const int N = p->N, k = p->k, kpl = (k + 1) * p->l;
    int32_t *tmp = (int32_t *) malloc(sizeof(int32_t) * (size_t) (k + 1) * N);
    for (int i = 0; i < n_iter; i++) {
        const int barai = bara[i];
        if (barai == 0) continue;                                  /* lwe-bootstrapping-functions.cu:66 */
        for (int j = 0; j <= k; j++)                               /* tLweMulByXaiMinusOne */
            oracle_mul_by_xai_minus_one(barai, N, accum + (size_t) j * N, tmp + (size_t) j * N);
        oracle_extern_mul_exact(p, bk + (size_t) i * kpl * (k + 1) * N, tmp);   /* tGswExternMulToTLwe */
        for (int j = 0; j < (k + 1) * N; j++)                      /* tLweAddTo */
            accum[j] = (int32_t) ((uint32_t) accum[j] + (uint32_t) tmp[j]);
    }
    free(tmp);
}

/* tfhe_blindRotateAndExtract (lwe-bootstrapping-functions.cu:95-126) + tfhe_bootstrap_woKS (:137-160) */
void oracle_bootstrap_woks_exact(const OracleParams *p, const int32_t *bk, int32_t mu, const int32_t *x,
                                 int32_t *u) {
    const int N = p->N, n = p->n, k = p->k;
    int32_t *testvect = (int32_t *) malloc(sizeof(int32_t) * (size_t) N);
    int32_t *bara = (int32_t *) malloc(sizeof(int32_t) * (size_t) n);
    int32_t *acc = (int32_t *) calloc((size_t) (k + 1) * N, sizeof(int32_t));
    const int barb = oracle_modswitch_from(x[n], 2 * N);
    for (int i = 0; i < n; i++) bara[i] = oracle_modswitch_from(x[i], 2 * N);
    for (int i = 0; i < N; i++) testvect[i] = mu;
    if (barb != 0) oracle_mul_by_xai(2 * N - barb, N, testvect, acc + (size_t) k * N);
    else memcpy(acc + (size_t) k * N, testvect, sizeof(int32_t) * (size_t) N);
    oracle_blind_rotate_exact(p, bk, acc, bara, n);
    for (int i = 0; i < k; i++) {                                  /* tLweExtractLweSample, lwe.cu:41-56 */
        u[i * N] = acc[(size_t) i * N];
        for (int j = 1; j < N; j++) u[i * N + j] = (int32_t) (0u - (uint32_t) acc[(size_t) i * N + N - j]);
    }
    u[k * N] = acc[(size_t) k * N];
    free(acc);
    free(bara);
    free(testvect);
}

void oracle_bootstrap_woks_exact_batch(const OracleParams *p, const int32_t *bk, int32_t mu, const int32_t *x,
                                       int32_t *u, int count, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int g = 0; g < count; g++)
        oracle_bootstrap_woks_exact(p, bk, mu, x + (size_t) g * (p->n + 1), u + (size_t) g * (p->N * p->k + 1));
}

/* --------------------------------------------------------------- key switch */

/* lwe-keyswitch-functions.cu:955-987 and :101-127 */
void oracle_keyswitch(const OracleCtx *c, const int32_t *u, int32_t *out) {
    const OracleParams *p = &c->p;
    const int n = p->n, nin = p->N * p->k, t = p->ks_t, basebit = p->ks_basebit, base = 1 << basebit;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    const uint32_t mask = (uint32_t) base - 1;
    uint32_t *r = (uint32_t *) out;
    for (int i = 0; i < n; i++) r[i] = 0;
    r[n] = (uint32_t) u[nin];
    for (int i = 0; i < nin; i++) {
        const uint32_t aibar = (uint32_t) u[i] + prec_offset;
        for (int j = 0; j < t; j++) {
            const uint32_t aij = (aibar >> (32 - (j + 1) * basebit)) & mask;
            if (aij != 0) {
                const uint32_t *row = (const uint32_t *) c->ks + (((size_t) i * t + j) * base + aij) * (n + 1);
                for (int q = 0; q <= n; q++) r[q] -= row[q];
            }
        }
    }
}

/* lwe-bootstrapping-functions-fft.cu:1884-1910 */
static void bootstrap_w(Work *w, int32_t mu, const int32_t *x, int32_t *out) {
    const OracleParams *p = &w->c->p;
    int32_t *u = (int32_t *) malloc(sizeof(int32_t) * (size_t) (p->N * p->k + 1));
    bootstrap_woks_w(w, mu, x, u);
    oracle_keyswitch(w->c, u, out);
    free(u);
}

void oracle_bootstrap(const OracleCtx *c, int32_t mu, const int32_t *x, int32_t *out) {
    Work *w = work_new(c);
    bootstrap_w(w, mu, x, out);
    work_free(w);
}

/* -------------------------------------------------------------------- gates */

/* boot-gates.cu: NAND :98-116, OR :124-142, AND :150-168, XOR :190-208, XNOR :216-234,
 * NOR :275-293, ANDNY :301-319, ANDYN :327-345, ORNY :353-371, ORYN :379-397.
 * x = (0, c/8 or c/4) + sa*ca + sb*cb */
static const struct { int cnum, cden, sa, sb; } GATE_TAB[ORACLE_NUM_GATES] = {
    /* NAND  */ { 1, 8, -1, -1},
    /* OR    */ { 1, 8,  1,  1},
    /* AND   */ {-1, 8,  1,  1},
    /* XOR   */ { 1, 4,  2,  2},
    /* XNOR  */ {-1, 4, -2, -2},
    /* NOR   */ {-1, 8, -1, -1},
    /* ANDNY */ {-1, 8, -1,  1},
    /* ANDYN */ {-1, 8,  1, -1},
    /* ORNY  */ { 1, 8, -1,  1},
    /* ORYN  */ { 1, 8,  1, -1},
};

void oracle_gate_prologue(const OracleParams *p, int gate, const int32_t *ca, const int32_t *cb,
                          int32_t *x) {
    const int n = p->n;
    const uint32_t cst = (uint32_t) oracle_modswitch_to(GATE_TAB[gate].cnum, GATE_TAB[gate].cden);
    const uint32_t sa = (uint32_t) GATE_TAB[gate].sa, sb = (uint32_t) GATE_TAB[gate].sb;
    for (int i = 0; i < n; i++) x[i] = (int32_t) (sa * (uint32_t) ca[i] + sb * (uint32_t) cb[i]);
    x[n] = (int32_t) (cst + sa * (uint32_t) ca[n] + sb * (uint32_t) cb[n]);
}

static void gate_w(Work *w, int gate, const int32_t *ca, const int32_t *cb, int32_t *out) {
    const OracleParams *p = &w->c->p;
    int32_t *x = (int32_t *) malloc(sizeof(int32_t) * (size_t) (p->n + 1));
    oracle_gate_prologue(p, gate, ca, cb, x);
    bootstrap_w(w, oracle_modswitch_to(1, 8), x, out);
    free(x);
}

void oracle_gate(const OracleCtx *c, int gate, const int32_t *ca, const int32_t *cb, int32_t *out) {
    Work *w = work_new(c);
    gate_w(w, gate, ca, cb, out);
    work_free(w);
}

/* boot-gates.cu:407-448 */
static void mux_w(Work *w, const int32_t *a, const int32_t *b, const int32_t *cc, int32_t *out) {
    const OracleParams *p = &w->c->p;
    const int n = p->n, nin = p->N * p->k;
    const int32_t MU = oracle_modswitch_to(1, 8);
    const uint32_t AndConst = (uint32_t) oracle_modswitch_to(-1, 8);
    int32_t *x = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *u1 = (int32_t *) malloc(sizeof(int32_t) * (size_t) (nin + 1));
    int32_t *u2 = (int32_t *) malloc(sizeof(int32_t) * (size_t) (nin + 1));
    for (int i = 0; i < n; i++) x[i] = (int32_t) ((uint32_t) a[i] + (uint32_t) b[i]);
    x[n] = (int32_t) (AndConst + (uint32_t) a[n] + (uint32_t) b[n]);
    bootstrap_woks_w(w, MU, x, u1);
    for (int i = 0; i < n; i++) x[i] = (int32_t) ((uint32_t) cc[i] - (uint32_t) a[i]);
    x[n] = (int32_t) (AndConst - (uint32_t) a[n] + (uint32_t) cc[n]);
    bootstrap_woks_w(w, MU, x, u2);
    for (int i = 0; i < nin; i++) u1[i] = (int32_t) ((uint32_t) u1[i] + (uint32_t) u2[i]);
    u1[nin] = (int32_t) ((uint32_t) MU + (uint32_t) u1[nin] + (uint32_t) u2[nin]);
    oracle_keyswitch(w->c, u1, out);
    free(x); free(u1); free(u2);
}

void oracle_mux(const OracleCtx *c, const int32_t *a, const int32_t *b, const int32_t *cc, int32_t *out) {
    Work *w = work_new(c);
    mux_w(w, a, b, cc, out);
    work_free(w);
}

/* boot-gates.cu:242-250 */
void oracle_not(const OracleParams *p, const int32_t *ca, int32_t *out) {
    for (int i = 0; i <= p->n; i++) out[i] = (int32_t) (0u - (uint32_t) ca[i]);
}

/* boot-gates.cu:263-267 */
void oracle_constant(const OracleParams *p, int value, int32_t *out) {
    const int32_t MU = oracle_modswitch_to(1, 8);
    for (int i = 0; i < p->n; i++) out[i] = 0;
    out[p->n] = value ? MU : -MU;
}

void oracle_gate_batch(const OracleCtx *c, int gate, const int32_t *ca, const int32_t *cb,
                       int32_t *out, int count, int threads) {
    const int stride = c->p.n + 1;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        Work *w = work_new(c);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int g = 0; g < count; g++)
            gate_w(w, gate, ca + (size_t) g * stride, cb + (size_t) g * stride, out + (size_t) g * stride);
        work_free(w);
    }
    (void) threads;
}

/* ----------------------------------------------------------------- circuits */

/* Cipher::addBits, Cipher.cu:367-378: t1=a^c, t2=b^c, sum=a^t2, t1=t1&t2, cout=c^t1 */
static void add_bits_w(Work *w, const int32_t *a, const int32_t *b, const int32_t *carry,
                       int32_t *sum, int32_t *cout) {
    const int stride = w->c->p.n + 1;
    int32_t *t1 = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * 2);
    int32_t *t2 = t1 + stride;
    gate_w(w, ORACLE_XOR, a, carry, t1);
    gate_w(w, ORACLE_XOR, b, carry, t2);
    gate_w(w, ORACLE_XOR, a, t2, sum);
    gate_w(w, ORACLE_AND, t1, t2, t1);
    gate_w(w, ORACLE_XOR, carry, t1, cout);
    free(t1);
}

/* Cipher operator+, Cipher.cu:334-352 (result truncated to nbits, :348) */
void oracle_add(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int32_t *out) {
    const int stride = c->p.n + 1;
    Work *w = work_new(c);
    int32_t *carry = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * 2);
    int32_t *cnext = carry + stride;
    oracle_constant(&c->p, 0, carry);
    for (int i = 0; i < nbits; i++) {
        add_bits_w(w, a + (size_t) i * stride, b + (size_t) i * stride, carry, out + (size_t) i * stride, cnext);
        memcpy(carry, cnext, sizeof(int32_t) * (size_t) stride);
    }
    free(carry);
    work_free(w);
}

/* Shift-add multiplier, truncated to nbits: Cipher operator*, Cipher.cu:83-108
 * (mulBinary = AND of every bit of a with b_i, innerLeftShift(i), out += sum);
 * the GPU schedule main.cu:1483-1579 computes the same value. */
void oracle_mul(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int32_t *out) {
    const int stride = c->p.n + 1;
    Work *w = work_new(c);
    int32_t *row = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nbits);
    int32_t *acc = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nbits);
    int32_t *tmp = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nbits);
    for (int j = 0; j < nbits; j++) oracle_constant(&c->p, 0, acc + (size_t) j * stride);
    for (int i = 0; i < nbits; i++) {
        for (int j = 0; j < nbits; j++) {
            if (j < i) oracle_constant(&c->p, 0, row + (size_t) j * stride);
            else gate_w(w, ORACLE_AND, a + (size_t) (j - i) * stride, b + (size_t) i * stride, row + (size_t) j * stride);
        }
        work_free(w);
        oracle_add(c, acc, row, nbits, tmp);
        w = work_new(c);
        memcpy(acc, tmp, sizeof(int32_t) * (size_t) stride * nbits);
    }
    memcpy(out, acc, sizeof(int32_t) * (size_t) stride * nbits);
    free(row); free(acc); free(tmp);
    work_free(w);
}

/* ---- the CPU reference's Cipher schedules (cpuParallel/Cipher.cpp), for the CPU baseline ----------
 * addNumberToSelf (cpuParallel/Cipher.cpp:198-226): ripple carry over all bits, 5 gates per bit. */
static void add_to_self_w(Work *w, int32_t *self, const int32_t *a, int nbits) {
    const int stride = w->c->p.n + 1;
    int32_t *carry = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * 3);
    int32_t *cnext = carry + stride, *sum = carry + 2 * stride;
    oracle_constant(&w->c->p, 0, carry);
    for (int i = 0; i < nbits; i++) {
        add_bits_w(w, self + (size_t) i * stride, a + (size_t) i * stride, carry, sum, cnext);
        memcpy(self + (size_t) i * stride, sum, sizeof(int32_t) * (size_t) stride);
        memcpy(carry, cnext, sizeof(int32_t) * (size_t) stride);
    }
    free(carry);
}

/* Cipher operator* (cpuParallel/Cipher.cpp:83-112): 2*nbits-bit product; for every bit i of b:
 * mulBinary (:244-249, nbits ANDs), innerLeftShift(i), out += sum (2*nbits-bit ripple add), under
 * `#pragma omp parallel for schedule(static) reduction(OMP_CIPHER_SUM:out)` (:90-94): every thread
 * accumulates its static chunk of i into a private zero-initialised sum, the private sums are then
 * added into `out` one after the other (libgomp combines reductions under a lock).
 * threads = 1 is the sequential form the file compiles to with PARALLEL undefined (:13). */
void oracle_cipher_mul(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int threads,
                       int32_t *out) {
    const int stride = c->p.n + 1, nb = 2 * nbits;
    if (threads < 1) threads = 1;
    if (threads > nbits) threads = nbits;
    for (int j = 0; j < nb; j++) oracle_constant(&c->p, 0, out + (size_t) j * stride);
#pragma omp parallel num_threads(threads)
    {
        Work *w = work_new(c);
        int32_t *priv = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nb);
        int32_t *sum = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nb);
        for (int j = 0; j < nb; j++) oracle_constant(&c->p, 0, priv + (size_t) j * stride);
#pragma omp for schedule(static)
        for (int i = 0; i < nbits; i++) {
            for (int j = 0; j < nb; j++) oracle_constant(&c->p, 0, sum + (size_t) j * stride);
            for (int j = 0; j < nbits; j++)  /* mulBinary + innerLeftShift(i) */
                gate_w(w, ORACLE_AND, b + (size_t) i * stride, a + (size_t) j * stride, sum + (size_t) (i + j) * stride);
            add_to_self_w(w, priv, sum, nb);
        }
#pragma omp critical
        add_to_self_w(w, out, priv, nb);
        free(sum);
        free(priv);
        work_free(w);
    }
}

/* `units` independent inner-loop bodies of the CPU matrix multiply (cpuParallel/cloud.cpp:390-408):
 * temp = A[i][k] * B[k][j] (operator*, sequential inside: nested inside `omp parallel for` over i, j)
 * and C[i][j] = C[i][j] + temp (2*nbits-bit add), one unit per thread at a time.
 * a, b: [units][nbits] samples; cacc: [units][2*nbits] samples, updated in place. */
void oracle_matmul_units(const OracleCtx *c, const int32_t *a, const int32_t *b, int32_t *cacc, int nbits,
                         int units, int threads) {
    const int stride = c->p.n + 1, nb = 2 * nbits;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int u = 0; u < units; u++) {
        int32_t *temp = (int32_t *) malloc(sizeof(int32_t) * (size_t) stride * nb);
        oracle_cipher_mul(c, a + (size_t) u * nbits * stride, b + (size_t) u * nbits * stride, nbits, 1, temp);
        Work *w = work_new(c);
        add_to_self_w(w, cacc + (size_t) u * nb * stride, temp, nb);
        work_free(w);
        free(temp);
    }
}
