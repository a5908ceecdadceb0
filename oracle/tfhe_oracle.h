/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * CPU restatement (plain C) of the reference's bootstrapped-gate hot path:
 * gate prologue -> mod-switch -> blind rotation (X^a rotation, gadget
 * decomposition, negacyclic FFT, Fourier MAC against the bootstrapping key,
 * inverse FFT) -> sample extraction -> LWE key switch, plus the key
 * generation / encryption / decryption needed to build test inputs, and the
 * Cipher-level circuits (ripple add, shift-add multiply, matrix multiply).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/gpuParallel unless noted).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use this code.
 *
 * Parity status: PINNED against the reference itself — oracle/_ref is the
 * reference's own host code compiled here; tests/test_oracle_vs_ref.py checks
 * this restatement bit-for-bit against it (FFT mode ORACLE_FFT_REF) and the
 * committed tests/golden/ vectors were produced by it.  The reference ships
 * no golden vectors of its own.
 *
 * Flat data formats (shared with oracle/ref_adapter.cpp and the CUDA engine):
 *   lwe_key  int32[n]                      binary
 *   tlwe_key int32[k][N]                   binary
 *   bk       int32[n][kpl][k+1][N]         TGSW rows, coefficient domain
 *   ks       int32[N*k][t][base][n+1]      a[0..n) then b
 *   sample   int32[n+1]                    a[0..n) then b
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int n;           /* LWE dimension                          (500)  */
    int N;           /* ring degree                            (1024) */
    int k;           /* TLWE mask polynomials                  (1)    */
    int l;           /* gadget length                          (2)    */
    int Bgbit;       /* log2 gadget base                       (10)   */
    int ks_t;        /* key-switch length                      (8)    */
    int ks_basebit;  /* key-switch log2 base                   (2)    */
    double alpha_lwe;/* LWE / KS noise stdev                          */
    double alpha_bk; /* TLWE / BK noise stdev                         */
} OracleParams;

enum {
    ORACLE_FFT_REF = 0,    /* 2N-point real transform, same arithmetic as oracle/_ref */
    ORACLE_FFT_FOLDED = 1  /* N/2-point twisted complex transform (fast port)        */
};

enum {
    ORACLE_NAND = 0, ORACLE_OR, ORACLE_AND, ORACLE_XOR, ORACLE_XNOR, ORACLE_NOR,
    ORACLE_ANDNY, ORACLE_ANDYN, ORACLE_ORNY, ORACLE_ORYN, ORACLE_NUM_GATES
};

typedef struct { uint64_t s[4]; int has_spare; double spare; } OracleRng;

typedef struct OracleCtx OracleCtx;

/* tfhe_gate_bootstrapping.cu:25-49 */
void oracle_default_params(OracleParams *p);
size_t oracle_bk_words(const OracleParams *p);
size_t oracle_ks_words(const OracleParams *p);

void oracle_rng_seed(OracleRng *r, uint64_t seed);
int32_t oracle_rng_torus(OracleRng *r);
double oracle_rng_gauss(OracleRng *r, double sigma);

/* numeric-functions.cu:33-77 */
int32_t oracle_dtot32(double d);
int oracle_modswitch_from(int32_t phase, int msize);
int32_t oracle_modswitch_to(int mu, int msize);

/* key generation: tfhe_gate_bootstrapping.cu:57-68, lwe-bootstrapping-functions.cu:185-217,
 * lwe-keyswitch-functions.cu:890-942, tgsw-functions.cu:129-194, tlwe-functions.cu:15-39 */
void oracle_keygen(const OracleParams *p, uint64_t seed, int32_t *lwe_key, int32_t *tlwe_key,
                   int32_t *bk, int32_t *ks);

/* lwe-functions.cu:36-81 ; tfhe_gate_bootstrapping.cu:114-125 */
void oracle_lwe_encrypt(OracleRng *r, const int32_t *key, int n, int32_t mu, double alpha,
                        int32_t *sample_out);
int32_t oracle_lwe_phase(const int32_t *sample, const int32_t *key, int n);
void oracle_encrypt_bit(const OracleParams *p, OracleRng *r, const int32_t *lwe_key, int bit,
                        int32_t *sample_out);
int oracle_decrypt_bit(const OracleParams *p, const int32_t *lwe_key, const int32_t *sample);

/* toruspolynomial-functions.cu:191-235, 492-519 */
void oracle_mul_by_xai(int a, int N, const int32_t *in, int32_t *out);
void oracle_mul_by_xai_minus_one(int a, int N, const int32_t *in, int32_t *out);
/* tgsw-functions.cu:301-352, tgsw.cu:7-29 */
void oracle_decomp(const OracleParams *p, const int32_t *poly, int32_t *out_l_by_N);

/* Context: parameters + Fourier-domain bootstrapping key + key-switch key.
 * bk / ks are the flat coefficient-domain arrays above; ks is borrowed (must
 * outlive the context).  lwe-bootstrapping-functions-fft.cu:60-89 */
OracleCtx *oracle_ctx_new(const OracleParams *p, const int32_t *bk, const int32_t *ks, int fft_mode);
void oracle_ctx_free(OracleCtx *c);
const OracleParams *oracle_ctx_params(const OracleCtx *c);
/* Fourier BK as held by the context: complex double [n][kpl][k+1][N/2] */
const double *oracle_ctx_bkfft(const OracleCtx *c);

/* fft_processor_fftw.cu:148-181 (mode REF) / folded equivalent (mode FOLDED) */
void oracle_ifft_int(const OracleCtx *c, const int32_t *poly, double *out_cplx);
void oracle_ifft_torus(const OracleCtx *c, const int32_t *poly, double *out_cplx);
void oracle_fft_torus(const OracleCtx *c, const double *in_cplx, int32_t *poly_out);

/* tgsw-fft-operations.cu:124-264 */
void oracle_extern_mul(const OracleCtx *c, int bk_index, int32_t *accum);
/* lwe-bootstrapping-functions-fft.cu:105-185, 676-737 */
void oracle_blind_rotate(const OracleCtx *c, int32_t *accum, const int32_t *bara, int n_iter);
/* lwe-bootstrapping-functions-fft.cu:1408-1456 ; lwe.cu:41-56 */
void oracle_blind_rotate_and_extract(const OracleCtx *c, const int32_t *testvect, int barb,
                                     const int32_t *bara, int n_iter, int32_t *u_out);
/* lwe-bootstrapping-functions-fft.cu:1834-1910 */
void oracle_bootstrap_woks(const OracleCtx *c, int32_t mu, const int32_t *x, int32_t *u_out);
void oracle_bootstrap(const OracleCtx *c, int32_t mu, const int32_t *x, int32_t *out);
/* lwe-keyswitch-functions.cu:101-127, 955-987 */
void oracle_keyswitch(const OracleCtx *c, const int32_t *u, int32_t *out);
/* boot-gates.cu:98-448 */
void oracle_gate_prologue(const OracleParams *p, int gate, const int32_t *ca, const int32_t *cb,
                          int32_t *x_out);
void oracle_gate(const OracleCtx *c, int gate, const int32_t *ca, const int32_t *cb, int32_t *out);
void oracle_mux(const OracleCtx *c, const int32_t *a, const int32_t *b, const int32_t *cc, int32_t *out);
void oracle_not(const OracleParams *p, const int32_t *ca, int32_t *out);
void oracle_constant(const OracleParams *p, int value, int32_t *out);

/* Exact (FFT-free) external product on coefficient-domain BK_i: int32[kpl][k+1][N]
 * lwe-bootstrapping-functions.cu:34-179 + multiplication.cu:53-77 */
void oracle_extern_mul_exact(const OracleParams *p, const int32_t *bk_i, int32_t *accum);

/* The whole non-FFT bootstrap with exact products (lwe-bootstrapping-functions.cu:34-179): bk is the
 * flat coefficient-domain key; no floating point anywhere, so the result is THE integer answer. */
void oracle_blind_rotate_exact(const OracleParams *p, const int32_t *bk, int32_t *accum, const int32_t *bara,
                               int n_iter);
void oracle_bootstrap_woks_exact(const OracleParams *p, const int32_t *bk, int32_t mu, const int32_t *x,
                                 int32_t *u_out);
void oracle_bootstrap_woks_exact_batch(const OracleParams *p, const int32_t *bk, int32_t mu, const int32_t *x,
                                       int32_t *u_out, int count, int threads);

/* Batch helpers (OpenMP over independent gates; used for the CPU baseline) */
void oracle_gate_batch(const OracleCtx *c, int gate, const int32_t *ca, const int32_t *cb,
                       int32_t *out, int count, int threads);

/* Cipher-level circuits (gate schedules only).  Operands are arrays of
 * nbits samples, LSB first, two's complement (Cipher.cu:5-7).
 * add: ripple carry, Cipher.cu:334-378 (addBits = 2 XOR + 2 AND + 1 OR per bit... see .c)
 * mul: shift-add, Cipher.cu:83-150 ; gpu schedule main.cu:1483-1579 */
void oracle_add(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int32_t *out);
void oracle_mul(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int32_t *out);

/* The CPU reference's own schedules (cpuParallel/Cipher.cpp, cpuParallel/cloud.cpp), timed by
 * bench.py beside the GPU circuits: operator* with its OpenMP reduction (2*nbits-bit product,
 * :83-112) and the inner-loop body of the matrix multiply (cloud.cpp:390-408). */
void oracle_cipher_mul(const OracleCtx *c, const int32_t *a, const int32_t *b, int nbits, int threads,
                       int32_t *out_2nbits);
void oracle_matmul_units(const OracleCtx *c, const int32_t *a, const int32_t *b, int32_t *cacc, int nbits,
                         int units, int threads);

#ifdef __cplusplus
}
#endif

#endif
