/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * Thin extern "C" adapter over the UNMODIFIED reference host path
 * (/root/reference/gpuParallel/*.cu, compiled where they lie by
 * oracle/build_ref.py into oracle/_ref/libtfhe_ref.so).  It only calls the
 * reference's public C API (gpuParallel/tfhe.h, tfhe_gate_bootstrapping_functions.h)
 * and flattens its pointer-rich structs into plain arrays so that the same
 * keys / ciphertexts can be handed to the C restatement (oracle/tfhe_oracle.c)
 * and to the CUDA engine.
 *
 * Flat formats (all little-endian int32 unless noted):
 *   lwe_key  [n]
 *   tlwe_key [k][N]
 *   bk       [n][kpl][k+1][N]        coefficient domain (TGswSample rows)
 *   ks       [N][t][base][n+1]       a[0..n) then b
 *   sample   [n+1]                   a[0..n) then b
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <new>

#include "tfhe.h"
#include "lwekey.h"
#include "tlwe.h"
#include "tgsw.h"
#include "polynomials.h"
#include "lagrangehalfc_impl.h"

extern "C" {

struct RefHandle {
    TFheGateBootstrappingParameterSet *params;
    TFheGateBootstrappingSecretKeySet *sk;      /* owns everything when keygen'd here  */
    /* pieces kept alive when the keyset was imported from flat arrays */
    LweKey *lwe_key;
    TGswKey *tgsw_key;
    LweBootstrappingKey *bk;
    LweBootstrappingKeyFFT *bkFFT;
    TFheGateBootstrappingCloudKeySet *cloud;
};

void ref_dims(int *out7) {
    TFheGateBootstrappingParameterSet *p = new_default_gate_bootstrapping_parameters(110);
    out7[0] = p->in_out_params->n;
    out7[1] = p->tgsw_params->tlwe_params->N;
    out7[2] = p->tgsw_params->tlwe_params->k;
    out7[3] = p->tgsw_params->l;
    out7[4] = p->tgsw_params->Bgbit;
    out7[5] = p->ks_t;
    out7[6] = p->ks_basebit;
    delete_gate_bootstrapping_parameters(p);
}

void ref_alphas(double *out3) {
    TFheGateBootstrappingParameterSet *p = new_default_gate_bootstrapping_parameters(110);
    out3[0] = p->in_out_params->alpha_min;
    out3[1] = p->tgsw_params->tlwe_params->alpha_min;
    out3[2] = p->in_out_params->alpha_max;
    delete_gate_bootstrapping_parameters(p);
}

RefHandle *ref_keygen(const uint32_t *seed, int nseed) {
    RefHandle *h = new RefHandle();
    memset(h, 0, sizeof(*h));
    tfhe_random_generator_setSeed(const_cast<uint32_t *>(seed), nseed);
    h->params = new_default_gate_bootstrapping_parameters(110);
    h->sk = new_random_gate_bootstrapping_secret_keyset(h->params);
    h->lwe_key = const_cast<LweKey *>(h->sk->lwe_key);
    h->tgsw_key = const_cast<TGswKey *>(h->sk->tgsw_key);
    h->bk = const_cast<LweBootstrappingKey *>(h->sk->cloud.bk);
    h->bkFFT = const_cast<LweBootstrappingKeyFFT *>(h->sk->cloud.bkFFT);
    h->cloud = const_cast<TFheGateBootstrappingCloudKeySet *>(&h->sk->cloud);
    return h;
}

void ref_reseed(const uint32_t *seed, int nseed) {
    tfhe_random_generator_setSeed(const_cast<uint32_t *>(seed), nseed);
}

/* Build a reference keyset from flat arrays (keys made elsewhere, e.g. by the
 * C restatement's portable keygen). */
RefHandle *ref_import(const int32_t *lwe_key, const int32_t *tlwe_key, const int32_t *bk,
                      const int32_t *ks) {
    RefHandle *h = new RefHandle();
    memset(h, 0, sizeof(*h));
    h->params = new_default_gate_bootstrapping_parameters(110);
    const LweParams *lp = h->params->in_out_params;
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    const int n = lp->n, N = tp->N, k = tp->k, kpl = gp->kpl;
    const int t = h->params->ks_t, basebit = h->params->ks_basebit, base = 1 << basebit;

    h->lwe_key = new_LweKey(lp);
    for (int i = 0; i < n; i++) h->lwe_key->key[i] = lwe_key[i];
    h->tgsw_key = new_TGswKey(gp);
    for (int j = 0; j < k; j++)
        for (int c = 0; c < N; c++) h->tgsw_key->key[j].coefs[c] = tlwe_key[j * N + c];

    h->bk = new_LweBootstrappingKey(t, basebit, lp, gp);
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++) {
                Torus32 *dst = h->bk->bk[i].all_sample[r].a[j].coefsT;
                memcpy(dst, bk + (((size_t) i * kpl + r) * (k + 1) + j) * N, sizeof(int32_t) * N);
            }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < t; j++)
            for (int v = 0; v < base; v++) {
                const int32_t *src = ks + (((size_t) i * t + j) * base + v) * (n + 1);
                LweSample *s = &h->bk->ks->ks[i][j][v];
                memcpy(s->a, src, sizeof(int32_t) * n);
                s->b = src[n];
                s->current_variance = 0.;
            }
    h->bkFFT = new_LweBootstrappingKeyFFT(h->bk);
    h->cloud = new TFheGateBootstrappingCloudKeySet(h->params, h->bk, h->bkFFT);
    return h;
}

void ref_export_lwe_key(const RefHandle *h, int32_t *out) {
    const int n = h->params->in_out_params->n;
    for (int i = 0; i < n; i++) out[i] = h->lwe_key->key[i];
}

void ref_export_tlwe_key(const RefHandle *h, int32_t *out) {
    const TLweParams *tp = h->params->tgsw_params->tlwe_params;
    for (int j = 0; j < tp->k; j++)
        for (int c = 0; c < tp->N; c++) out[j * tp->N + c] = h->tgsw_key->key[j].coefs[c];
}

void ref_export_bk(const RefHandle *h, int32_t *out) {
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    const int n = h->params->in_out_params->n, N = tp->N, k = tp->k, kpl = gp->kpl;
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++)
                memcpy(out + (((size_t) i * kpl + r) * (k + 1) + j) * N,
                       h->bk->bk[i].all_sample[r].a[j].coefsT, sizeof(int32_t) * N);
}

/* Fourier-domain BK exactly as the reference holds it: complex<double>
 * [n][kpl][k+1][N/2], value j = P(exp(-i*pi*(2j+1)/N)) (fft_processor_fftw.cu:158-167). */
void ref_export_bkfft(const RefHandle *h, double *out) {
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    const int n = h->params->in_out_params->n, N = tp->N, k = tp->k, kpl = gp->kpl;
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++) {
                const LagrangeHalfCPolynomial_IMPL *lp =
                    (const LagrangeHalfCPolynomial_IMPL *) &h->bkFFT->bkFFT[i].all_samples[r].a[j];
                memcpy(out + (((size_t) i * kpl + r) * (k + 1) + j) * N, lp->coefsC,
                       sizeof(double) * N);
            }
}

void ref_export_ks(const RefHandle *h, int32_t *out) {
    const int n = h->params->in_out_params->n;
    const int N = h->params->tgsw_params->tlwe_params->N;
    const int t = h->params->ks_t, base = 1 << h->params->ks_basebit;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < t; j++)
            for (int v = 0; v < base; v++) {
                int32_t *dst = out + (((size_t) i * t + j) * base + v) * (n + 1);
                const LweSample *s = &h->bkFFT->ks->ks[i][j][v];
                memcpy(dst, s->a, sizeof(int32_t) * n);
                dst[n] = s->b;
            }
}

static LweSample *mk(const LweParams *p, const int32_t *flat) {
    LweSample *s = new_LweSample(p);
    memcpy(s->a, flat, sizeof(int32_t) * p->n);
    s->b = flat[p->n];
    s->current_variance = 0.;
    return s;
}

static void put(int32_t *flat, const LweSample *s, int n) {
    memcpy(flat, s->a, sizeof(int32_t) * n);
    flat[n] = s->b;
}

void ref_encrypt(const RefHandle *h, int message, int32_t *out) {
    const LweParams *p = h->params->in_out_params;
    LweSample *s = new_LweSample(p);
    if (h->sk) bootsSymEncrypt(s, message, h->sk);
    else {
        Torus32 mu = modSwitchToTorus32(1, 8);
        lweSymEncrypt(s, message ? mu : -mu, p->alpha_min, h->lwe_key);
    }
    put(out, s, p->n);
    delete_LweSample(s);
}

int32_t ref_phase(const RefHandle *h, const int32_t *sample) {
    const LweParams *p = h->params->in_out_params;
    LweSample *s = mk(p, sample);
    Torus32 ph = lwePhase(s, h->lwe_key);
    delete_LweSample(s);
    return ph;
}

/* gate ids shared with include/tfhe_b200.h */
enum { G_NAND = 0, G_OR, G_AND, G_XOR, G_XNOR, G_NOR, G_ANDNY, G_ANDYN, G_ORNY, G_ORYN };

void ref_gate(const RefHandle *h, int gate, const int32_t *ca, const int32_t *cb, int32_t *out) {
    const LweParams *p = h->params->in_out_params;
    LweSample *a = mk(p, ca), *b = mk(p, cb), *r = new_LweSample(p);
    switch (gate) {
        case G_NAND: bootsNAND(r, a, b, h->cloud); break;
        case G_OR: bootsOR(r, a, b, h->cloud); break;
        case G_AND: bootsAND(r, a, b, h->cloud); break;
        case G_XOR: bootsXOR(r, a, b, h->cloud); break;
        case G_XNOR: bootsXNOR(r, a, b, h->cloud); break;
        case G_NOR: bootsNOR(r, a, b, h->cloud); break;
        case G_ANDNY: bootsANDNY(r, a, b, h->cloud); break;
        case G_ANDYN: bootsANDYN(r, a, b, h->cloud); break;
        case G_ORNY: bootsORNY(r, a, b, h->cloud); break;
        case G_ORYN: bootsORYN(r, a, b, h->cloud); break;
        default: fprintf(stderr, "ref_gate: bad gate %d\n", gate); abort();
    }
    put(out, r, p->n);
    delete_LweSample(a); delete_LweSample(b); delete_LweSample(r);
}

void ref_mux(const RefHandle *h, const int32_t *ca, const int32_t *cb, const int32_t *cc, int32_t *out) {
    const LweParams *p = h->params->in_out_params;
    LweSample *a = mk(p, ca), *b = mk(p, cb), *c = mk(p, cc), *r = new_LweSample(p);
    bootsMUX(r, a, b, c, h->cloud);
    put(out, r, p->n);
    delete_LweSample(a); delete_LweSample(b); delete_LweSample(c); delete_LweSample(r);
}

/* x (dim n) -> u (dim N), no key switch: tfhe_bootstrap_woKS_FFT */
void ref_bootstrap_woks(const RefHandle *h, int32_t mu, const int32_t *x, int32_t *u_out) {
    const LweParams *p = h->params->in_out_params;
    const LweParams *ep = &h->params->tgsw_params->tlwe_params->extracted_lweparams;
    LweSample *xs = mk(p, x), *u = new_LweSample(ep);
    tfhe_bootstrap_woKS_FFT(u, h->bkFFT, mu, xs);
    put(u_out, u, ep->n);
    delete_LweSample(xs); delete_LweSample(u);
}

/* u (dim N) -> result (dim n): lweKeySwitch */
void ref_keyswitch(const RefHandle *h, const int32_t *u, int32_t *out) {
    const LweParams *p = h->params->in_out_params;
    const LweParams *ep = &h->params->tgsw_params->tlwe_params->extracted_lweparams;
    LweSample *us = mk(ep, u), *r = new_LweSample(p);
    lweKeySwitch(r, h->bkFFT->ks, us);
    put(out, r, p->n);
    delete_LweSample(us); delete_LweSample(r);
}

/* One external product accum <- BK_i (.) accum, accum = int32[k+1][N]: tGswFFTExternMulToTLwe */
void ref_extern_mul(const RefHandle *h, int bk_index, int32_t *accum) {
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    TLweSample *acc = new_TLweSample(tp);
    for (int j = 0; j <= tp->k; j++) memcpy(acc->a[j].coefsT, accum + j * tp->N, sizeof(int32_t) * tp->N);
    tGswFFTExternMulToTLwe(acc, &h->bkFFT->bkFFT[bk_index], gp);
    for (int j = 0; j <= tp->k; j++) memcpy(accum + j * tp->N, acc->a[j].coefsT, sizeof(int32_t) * tp->N);
    delete_TLweSample(acc);
}

/* Blind rotation only, on a caller-provided accumulator, over the first n_iter
 * key elements: tfhe_blindRotate_FFT */
void ref_blind_rotate(const RefHandle *h, int32_t *accum, const int32_t *bara, int n_iter) {
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    TLweSample *acc = new_TLweSample(tp);
    for (int j = 0; j <= tp->k; j++) memcpy(acc->a[j].coefsT, accum + j * tp->N, sizeof(int32_t) * tp->N);
    tfhe_blindRotate_FFT(acc, h->bkFFT->bkFFT, bara, n_iter, gp);
    for (int j = 0; j <= tp->k; j++) memcpy(accum + j * tp->N, acc->a[j].coefsT, sizeof(int32_t) * tp->N);
    delete_TLweSample(acc);
}

/* The reference's NON-FFT blind rotation (tfhe_blindRotate / tfhe_MuxRotate,
 * lwe-bootstrapping-functions.cu:34-79; tGswExternMulToTLwe, tgsw-functions.cu:156-170) with the
 * polynomial products taken by the reference's own torusPolynomialMultNaive (multiplication.cu:72).
 * As shipped, polynomials_arithmetic.h:112-114 routes torusPolynomialAddMulR to the FFT, so the
 * sequence is spelled out here from the reference's functions; everything is integer arithmetic. */
void ref_blind_rotate_naive(const RefHandle *h, int32_t *accum, const int32_t *bara, int n_iter) {
    const TGswParams *gp = h->params->tgsw_params;
    const TLweParams *tp = gp->tlwe_params;
    const int N = tp->N, k = tp->k, kpl = gp->kpl;
    TLweSample *acc = new_TLweSample(tp), *tmp = new_TLweSample(tp), *res = new_TLweSample(tp);
    IntPolynomial *dec = new_IntPolynomial_array(kpl, N);
    TorusPolynomial *prod = new_TorusPolynomial(N);
    for (int j = 0; j <= k; j++) memcpy(acc->a[j].coefsT, accum + j * N, sizeof(int32_t) * N);
    for (int i = 0; i < n_iter; i++) {
        if (bara[i] == 0) continue;
        tLweMulByXaiMinusOne(tmp, bara[i], acc, tp);
        tGswTLweDecompH(dec, tmp, gp);
        tLweClear(res, tp);
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++) {
                torusPolynomialMultNaive(prod, &dec[r], &h->bk->bk[i].all_sample[r].a[j]);
                torusPolynomialAddTo(&res->a[j], prod);
            }
        tLweAddTo(res, acc, tp);
        for (int j = 0; j <= k; j++) memcpy(acc->a[j].coefsT, res->a[j].coefsT, sizeof(int32_t) * N);
    }
    for (int j = 0; j <= k; j++) memcpy(accum + j * N, acc->a[j].coefsT, sizeof(int32_t) * N);
    delete_TorusPolynomial(prod);
    delete_IntPolynomial_array(kpl, dec);
    delete_TLweSample(res); delete_TLweSample(tmp); delete_TLweSample(acc);
}

/* `count` calls of a two-input gate function from OpenMP worker threads, one sample per call, the way
 * the CPU reference's Cipher operators issue them (cpuParallel/Cipher.cpp:114-121 cipherAND,
 * :75-76).  fn is any function with the bootsXXX signature (the reference's own or the drop-in
 * library's, found by the test through dlsym); returns the wall time of the parallel loop. */
typedef void (*ref_gate2_fn)(LweSample *, const LweSample *, const LweSample *, const TFheGateBootstrappingCloudKeySet *);
double ref_omp_gate_calls(void *fn, void **results, void **as, void **bs, const void *cloud, int count, int threads) {
    ref_gate2_fn f = (ref_gate2_fn) fn;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int i = 0; i < count; i++)
        f((LweSample *) results[i], (const LweSample *) as[i], (const LweSample *) bs[i],
          (const TFheGateBootstrappingCloudKeySet *) cloud);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* Gadget decomposition of one polynomial: tGswTorus32PolynomialDecompH */
void ref_decomp(const RefHandle *h, const int32_t *poly, int32_t *out_l_by_N) {
    const TGswParams *gp = h->params->tgsw_params;
    const int N = gp->tlwe_params->N, l = gp->l;
    TorusPolynomial *p = new_TorusPolynomial(N);
    IntPolynomial *d = new_IntPolynomial_array(l, N);
    memcpy(p->coefsT, poly, sizeof(int32_t) * N);
    tGswTorus32PolynomialDecompH(d, p, gp);
    for (int q = 0; q < l; q++) memcpy(out_l_by_N + q * N, d[q].coefs, sizeof(int32_t) * N);
    delete_IntPolynomial_array(l, d);
    delete_TorusPolynomial(p);
}

/* (X^a - 1) * poly : torusPolynomialMulByXaiMinusOne ; X^a * poly : torusPolynomialMulByXai */
void ref_mul_by_xai(int a, int minus_one, int N, const int32_t *poly, int32_t *out) {
    TorusPolynomial *p = new_TorusPolynomial(N), *r = new_TorusPolynomial(N);
    memcpy(p->coefsT, poly, sizeof(int32_t) * N);
    if (minus_one) torusPolynomialMulByXaiMinusOne(r, a, p);
    else torusPolynomialMulByXai(r, a, p);
    memcpy(out, r->coefsT, sizeof(int32_t) * N);
    delete_TorusPolynomial(p); delete_TorusPolynomial(r);
}

int ref_modswitch_from(int32_t phase, int msize) { return modSwitchFromTorus32(phase, msize); }
int32_t ref_modswitch_to(int mu, int msize) { return modSwitchToTorus32(mu, msize); }

/* Reference Fourier transforms on one polynomial (fft_processor_fftw.cu:148-181) */
void ref_ifft_int(const int32_t *poly, double *out_cplx_Ns2) {
    fp1024_fftw.execute_reverse_int((cplx *) out_cplx_Ns2, poly);
}
void ref_ifft_torus(const int32_t *poly, double *out_cplx_Ns2) {
    fp1024_fftw.execute_reverse_torus32((cplx *) out_cplx_Ns2, poly);
}
void ref_fft_torus(const double *in_cplx_Ns2, int32_t *poly_out) {
    fp1024_fftw.execute_direct_Torus32(poly_out, (const cplx *) in_cplx_Ns2);
}

/* Timing helper for the CPU baseline: runs `count` bootsNAND gates on the given
 * inputs (flat samples, cycled) and returns elapsed seconds. */
double ref_time_nand(const RefHandle *h, const int32_t *ca, const int32_t *cb, int n_inputs,
                     int count, int32_t *last_out) {
    const LweParams *p = h->params->in_out_params;
    LweSample **A = new LweSample *[n_inputs], **B = new LweSample *[n_inputs];
    for (int i = 0; i < n_inputs; i++) {
        A[i] = mk(p, ca + (size_t) i * (p->n + 1));
        B[i] = mk(p, cb + (size_t) i * (p->n + 1));
    }
    LweSample *r = new_LweSample(p);
    auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < count; g++) bootsNAND(r, A[g % n_inputs], B[g % n_inputs], h->cloud);
    auto t1 = std::chrono::steady_clock::now();
    if (last_out) put(last_out, r, p->n);
    for (int i = 0; i < n_inputs; i++) { delete_LweSample(A[i]); delete_LweSample(B[i]); }
    delete[] A; delete[] B;
    delete_LweSample(r);
    return std::chrono::duration<double>(t1 - t0).count();
}

/* Raw reference objects, for ABI tests of the drop-in library: the pointers are handed to
 * libtfhe_b200.so's bootsNAND / tfhe_bootstrap_FFT / lweKeySwitch ... unchanged. */
const void *ref_cloud_keyset(const RefHandle *h) { return h->cloud; }
const void *ref_bkfft(const RefHandle *h) { return h->bkFFT; }
const void *ref_tgsw_fft_array(const RefHandle *h) { return h->bkFFT->bkFFT; }
const void *ref_tgsw_params(const RefHandle *h) { return h->params->tgsw_params; }
const void *ref_ks_key(const RefHandle *h) { return h->bkFFT->ks; }
void *ref_sample_new(const RefHandle *h, int extracted) {
    return new_LweSample(extracted ? &h->params->tgsw_params->tlwe_params->extracted_lweparams
                                   : h->params->in_out_params);
}
void ref_sample_set(void *s, const int32_t *flat, int n) {
    LweSample *p = (LweSample *) s;
    memcpy(p->a, flat, sizeof(int32_t) * n);
    p->b = flat[n];
}
void ref_sample_get(const void *s, int32_t *flat, int n) {
    const LweSample *p = (const LweSample *) s;
    memcpy(flat, p->a, sizeof(int32_t) * n);
    flat[n] = p->b;
}
void ref_sample_free(void *s) { delete_LweSample((LweSample *) s); }
void *ref_tlwe_new(const RefHandle *h) { return new_TLweSample(h->params->tgsw_params->tlwe_params); }
void ref_tlwe_set(void *s, const int32_t *flat, int N, int k) {
    TLweSample *p = (TLweSample *) s;
    for (int j = 0; j <= k; j++) memcpy(p->a[j].coefsT, flat + j * N, sizeof(int32_t) * N);
}
void ref_tlwe_get(const void *s, int32_t *flat, int N, int k) {
    const TLweSample *p = (const TLweSample *) s;
    for (int j = 0; j <= k; j++) memcpy(flat + j * N, p->a[j].coefsT, sizeof(int32_t) * N);
}
void ref_tlwe_free(void *s) { delete_TLweSample((TLweSample *) s); }
void *ref_torus_poly_new(int N, const int32_t *coefs) {
    TorusPolynomial *p = new_TorusPolynomial(N);
    memcpy(p->coefsT, coefs, sizeof(int32_t) * N);
    return p;
}
void ref_torus_poly_free(void *p) { delete_TorusPolynomial((TorusPolynomial *) p); }

/* ---- the reference's own file formats (tfhe_io.cu) ------------------------------------ */
int ref_write_cloud_key(const RefHandle *h, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    export_tfheGateBootstrappingCloudKeySet_toFile(f, h->cloud);
    return fclose(f);
}

int ref_write_secret_key(const RefHandle *h, const char *path) {
    if (!h->sk) return 2;
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    export_tfheGateBootstrappingSecretKeySet_toFile(f, h->sk);
    return fclose(f);
}

RefHandle *ref_read_cloud_key(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return nullptr;
    TFheGateBootstrappingCloudKeySet *ck = new_tfheGateBootstrappingCloudKeySet_fromFile(f);
    fclose(f);
    RefHandle *h = new RefHandle();
    memset(h, 0, sizeof(*h));
    h->params = const_cast<TFheGateBootstrappingParameterSet *>(ck->params);
    h->bk = const_cast<LweBootstrappingKey *>(ck->bk);
    h->bkFFT = const_cast<LweBootstrappingKeyFFT *>(ck->bkFFT);
    h->cloud = ck;
    return h;
}

RefHandle *ref_read_secret_key(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return nullptr;
    TFheGateBootstrappingSecretKeySet *sk = new_tfheGateBootstrappingSecretKeySet_fromFile(f);
    fclose(f);
    RefHandle *h = new RefHandle();
    memset(h, 0, sizeof(*h));
    h->params = const_cast<TFheGateBootstrappingParameterSet *>(sk->params);
    h->sk = sk;
    h->lwe_key = const_cast<LweKey *>(sk->lwe_key);
    h->tgsw_key = const_cast<TGswKey *>(sk->tgsw_key);
    h->bk = const_cast<LweBootstrappingKey *>(sk->cloud.bk);
    h->bkFFT = const_cast<LweBootstrappingKeyFFT *>(sk->cloud.bkFFT);
    h->cloud = const_cast<TFheGateBootstrappingCloudKeySet *>(&sk->cloud);
    return h;
}

int ref_write_ciphertexts(const RefHandle *h, const char *path, const int32_t *flat, const double *variances,
                          int count) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    const int n = h->params->in_out_params->n;
    for (int i = 0; i < count; i++) {
        LweSample *s = mk(h->params->in_out_params, flat + (size_t) i * (n + 1));
        s->current_variance = variances ? variances[i] : 0.;
        export_gate_bootstrapping_ciphertext_toFile(f, s, h->params);
        delete_LweSample(s);
    }
    return fclose(f);
}

int ref_read_ciphertexts(const RefHandle *h, const char *path, int32_t *flat, double *variances, int count) {
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    const int n = h->params->in_out_params->n;
    LweSample *s = new_LweSample(h->params->in_out_params);
    for (int i = 0; i < count; i++) {
        import_gate_bootstrapping_ciphertext_fromFile(f, s, h->params);
        put(flat + (size_t) i * (n + 1), s, n);
        if (variances) variances[i] = s->current_variance;
    }
    delete_LweSample(s);
    fclose(f);
    return 0;
}

void ref_free(RefHandle *h) {
    if (!h) return;
    if (h->sk) {
        delete_gate_bootstrapping_secret_keyset(h->sk);
    } else {
        if (h->bkFFT) delete_LweBootstrappingKeyFFT(h->bkFFT);
        if (h->bk) delete_LweBootstrappingKey(h->bk);
        if (h->tgsw_key) delete_TGswKey(h->tgsw_key);
        if (h->lwe_key) delete_LweKey(h->lwe_key);
        /* cloud keyset struct itself holds only pointers */
        ::operator delete((void *) h->cloud);
    }
    delete h;
}

} /* extern "C" */
