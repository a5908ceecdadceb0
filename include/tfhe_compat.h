/*
 * tfhe_compat.h — drop-in replacements for the reference's own entry points on the
 * bootstrapped-gate path, with the reference's struct layouts.
 *
 * A program written against the reference API (gpuParallel/tfhe.h, <tfhe/tfhe.h> for the
 * cpuParallel programs) can link libtfhe_b200.so for these symbols: same names, same
 * argument meaning, same ownership rules (the caller allocates every result; results may
 * alias inputs), same error behaviour (failures abort the process, as die_dramatically()
 * does: gpuParallel/tfhe_gate_bootstrapping.cu:11-15).  The arithmetic runs on the GPU; there
 * is no CPU fallback.
 *
 * Struct definitions below are plain-C mirrors (same member order and types, hence the same
 * layout) of the reference's C++ structs; each cites its origin.  Include EITHER this header
 * OR the reference's headers, not both.
 */
#ifndef TFHE_COMPAT_H
#define TFHE_COMPAT_H

#include <stdint.h>

#include "tfhe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TFHE_CORE_H /* the reference's own headers are not in use */

typedef int32_t Torus32; /* tfhe_core.h:28 */

typedef struct LweParams { /* lweparams.h:13-16 */
    int n;
    double alpha_min;
    double alpha_max;
} LweParams;

typedef struct LweSample { /* lwesamples.h:18-29 */
    Torus32 *a;
    Torus32 b;
    double current_variance;
} LweSample;

typedef struct LweSample_16 { /* lwesamples.h:9-13: a on the DEVICE [nBits][n], b / variance on the HOST */
    int *a;
    int *b;
    double *current_variance;
} LweSample_16;

typedef struct TLweParams { /* tlwe.h:10-16 */
    int N;
    int k;
    double alpha_min;
    double alpha_max;
    LweParams extracted_lweparams;
} TLweParams;

typedef struct IntPolynomial { /* polynomials.h:11-14 */
    int N;
    int *coefs;
} IntPolynomial;

typedef struct TorusPolynomial { /* polynomials.h:24-27 */
    int N;
    Torus32 *coefsT;
} TorusPolynomial;

typedef struct LagrangeHalfCPolynomial { /* polynomials.h:43-47; data -> N/2 complex<double> */
    void *data;
    void *precomp;
} LagrangeHalfCPolynomial;

typedef struct TLweSample { /* tlwe.h:47-52 */
    TorusPolynomial *a;
    TorusPolynomial *b;
    double current_variance;
    int k;
} TLweSample;

typedef struct TLweSampleFFT { /* tlwe.h:65-70 */
    LagrangeHalfCPolynomial *a;
    LagrangeHalfCPolynomial *b;
    double current_variance;
    int k;
} TLweSampleFFT;

typedef struct TGswParams { /* tgsw.h:10-20 */
    int l;
    int Bgbit;
    int Bg;
    int32_t halfBg;
    uint32_t maskMod;
    const TLweParams *tlwe_params;
    int kpl;
    Torus32 *h;
    uint32_t offset;
} TGswParams;

typedef struct TGswSample { /* tgsw.h:60-65 */
    TLweSample *all_sample;
    TLweSample **bloc_sample;
    int k;
    int l;
} TGswSample;

typedef struct TGswSampleFFT { /* tgsw.h:78-84 */
    TLweSampleFFT *all_samples;
    TLweSampleFFT **sample;
    int k;
    int l;
} TGswSampleFFT;

typedef struct LweKeySwitchKey { /* lwekeyswitch.h:11-20 */
    int n;
    int t;
    int basebit;
    int base;
    const LweParams *out_params;
    LweSample *ks0_raw;
    LweSample **ks1_raw;
    LweSample ***ks;
} LweKeySwitchKey;

typedef struct LweBootstrappingKey { /* lwebootstrappingkey.h:10-16 */
    const LweParams *in_out_params;
    const TGswParams *bk_params;
    const TLweParams *accum_params;
    const LweParams *extract_params;
    TGswSample *bk;
    LweKeySwitchKey *ks;
} LweBootstrappingKey;

typedef struct LweBootstrappingKeyFFT { /* lwebootstrappingkey.h:36-43 */
    const LweParams *in_out_params;
    const TGswParams *bk_params;
    const TLweParams *accum_params;
    const LweParams *extract_params;
    const TGswSampleFFT *bkFFT;
    const LweKeySwitchKey *ks;
} LweBootstrappingKeyFFT;

typedef struct TFheGateBootstrappingParameterSet { /* tfhe_gate_bootstrapping_structures.h:8-12 */
    int ks_t;
    int ks_basebit;
    const LweParams *in_out_params;
    const TGswParams *tgsw_params;
} TFheGateBootstrappingParameterSet;

typedef struct TFheGateBootstrappingCloudKeySet { /* tfhe_gate_bootstrapping_structures.h:26-29 */
    const TFheGateBootstrappingParameterSet *params;
    const LweBootstrappingKey *bk;
    const LweBootstrappingKeyFFT *bkFFT;
} TFheGateBootstrappingCloudKeySet;

#endif /* TFHE_CORE_H */

/* ---- files and result buffers (tfhe_io.h, tfhe_gate_bootstrapping_functions.h) ----
 * What cpu/cloud.cpp:138-161 needs besides the gates: read cloud.key, allocate / read / write
 * ciphertexts.  A key set read here holds the coefficient-domain key (bkFFT == NULL; the gates
 * convert it on the GPU at first use) and is released with tfhe_b200_delete_cloud_keyset_fromFile. */
#include <stdio.h>
TFheGateBootstrappingCloudKeySet *new_tfheGateBootstrappingCloudKeySet_fromFile(FILE *F); /* tfhe_io.cu:1117 */
void tfhe_b200_delete_cloud_keyset_fromFile(TFheGateBootstrappingCloudKeySet *keyset);
void export_tfheGateBootstrappingCloudKeySet_toFile(FILE *F, const TFheGateBootstrappingCloudKeySet *keyset); /* :1109 */
void export_gate_bootstrapping_ciphertext_toFile(FILE *F, const LweSample *sample,
                                                 const TFheGateBootstrappingParameterSet *params); /* :1214 */
void import_gate_bootstrapping_ciphertext_fromFile(FILE *F, LweSample *sample,
                                                   const TFheGateBootstrappingParameterSet *params); /* :1223 */
LweSample *new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet *params);
LweSample *new_gate_bootstrapping_ciphertext_array(int nbelems, const TFheGateBootstrappingParameterSet *params);
void delete_gate_bootstrapping_ciphertext(LweSample *sample);
void delete_gate_bootstrapping_ciphertext_array(int nbelems, LweSample *samples); /* tfhe_gate_bootstrapping.cu:93-108 */

/* ---- classic gate API (tfhe_gate_bootstrapping_functions.h:40-87; boot-gates.cu:98-448) ---- */
void bootsNAND(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsAND(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsXOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsXNOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsNOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsANDNY(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsANDYN(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsORNY(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsORYN(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsMUX(LweSample *result, const LweSample *a, const LweSample *b, const LweSample *c,
              const TFheGateBootstrappingCloudKeySet *bk);
void bootsNOT(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk);
void bootsCOPY(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk);
void bootsCONSTANT(LweSample *result, int value, const TFheGateBootstrappingCloudKeySet *bk);

/* ---- bootstrapping internals (tfhe.h:49-52, tgsw_functions.h, lwe-functions.h) ------------ */
/* lwe-bootstrapping-functions-fft.cu:676 */
void tfhe_blindRotate_FFT(TLweSample *accum, const TGswSampleFFT *bk, const int *bara, const int n,
                          const TGswParams *bk_params);
/* :1408 */
void tfhe_blindRotateAndExtract_FFT(LweSample *result, const TorusPolynomial *v, const TGswSampleFFT *bk,
                                    const int barb, const int *bara, const int n, const TGswParams *bk_params);
/* :1834 */
void tfhe_bootstrap_woKS_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x);
/* :1884 */
void tfhe_bootstrap_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x);
/* tgsw-fft-operations.cu:124 */
void tGswFFTExternMulToTLwe(TLweSample *accum, const TGswSampleFFT *gsw, const TGswParams *params);
/* lwe-keyswitch-functions.cu:955 */
void lweKeySwitch(LweSample *result, const LweKeySwitchKey *ks, const LweSample *sample);

/* ---- batched "fullGPU" family (tfhe_gate_bootstrapping_functions.h:171-198) ---------------
 * Same semantics and the same LweSample_16 convention (a on the device, b on the host).  The
 * three raw key pointers of the reference (made by sendBootstrappingKeyToGPUCoalesceExt /
 * sendKeySwitchKeyToGPU_extendedOnePointer / sendKeySwitchBtoGPUOnePtr, main.cu:165,364,236)
 * are replaced by ONE handle: pass tfhe_b200_keys_to_gpu(bk) as bkGPU; ksA / ksB are ignored. */
void *tfhe_b200_keys_to_gpu(const TFheGateBootstrappingCloudKeySet *bk);
void tfhe_b200_keys_free(const TFheGateBootstrappingCloudKeySet *bk);
/* GPU contexts are cached per host key object (address + a fingerprint of the key MATERIAL, so a key
 * rewritten in place or a new key at a recycled address is never served from a stale device copy;
 * at most 8 bare TGSW / key-switch contexts are kept, least recently used first out).
 * tfhe_b200_compat_invalidate drops the contexts of one key object (any of: cloud key set,
 * LweBootstrappingKeyFFT, TGswSampleFFT array, LweKeySwitchKey); release_all drops everything. */
void tfhe_b200_compat_invalidate(const void *key_object);
void tfhe_b200_compat_release_all(void);
int tfhe_b200_compat_cached_contexts(void);
/* The classic single-sample gates (bootsNAND ... bootsMUX) are thread-safe and COALESCED: calls that
 * arrive from different host threads while a batch is in flight (or within TFHE_B200_COALESCE_WINDOW_US,
 * default 40 us, of a lone call) share one GPU launch (cpuParallel/Cipher.cpp:75-76, 94 call them from
 * OpenMP workers).  Statistics: launches issued / gates served for this key set. */
void tfhe_b200_compat_coalescer_stats(const TFheGateBootstrappingCloudKeySet *bk, unsigned long long *batches,
                                      unsigned long long *gates);
void bootsAND_fullGPU_n_Bit(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb, int nBits,
                            void *bkGPU, Torus32 *ksA, Torus32 *ksB);
void bootsXOR_fullGPU_n_Bit(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb, int nBits,
                            void *bkGPU, Torus32 *ksA, Torus32 *ksB);
void bootsXNOR_fullGPU_n_Bit(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb, int nBits,
                             void *bkGPU, Torus32 *ksA, Torus32 *ksB);
void bootsMUX_fullGPU_n_Bit(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb,
                            const LweSample_16 *cc, int nBits, void *bkGPU, Torus32 *ksA, Torus32 *ksB);
/* result[0..v*nBits) = AND, result[v*nBits..2*v*nBits) = XOR (boot-gates.cu:3050-3053) */
void bootsANDXOR_fullGPU_n_Bit_vector(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb,
                                      int vLength, int nBits, void *bkGPU, Torus32 *ksA, Torus32 *ksB);
/* result = [ca1 ^ ca2, cb1 ^ cb2] (boot-gates.cu:3088-3091) */
void bootsXORXOR_fullGPU_n_Bit_vector(LweSample_16 *result, const LweSample_16 *ca1, const LweSample_16 *ca2,
                                      const LweSample_16 *cb1, const LweSample_16 *cb2, int vLength, int nBits,
                                      void *bkGPU, Torus32 *ksA, Torus32 *ksB);
void bootsNOT_16(LweSample_16 *output, LweSample_16 *input, int bitSize, int params_n);
/* boot-gates.cu:462-476 and main.cu:41 */
LweSample_16 *convertBitToNumberZero_GPU(int bitSize, const TFheGateBootstrappingCloudKeySet *bk);
/* host containers (boot-gates.cu:513-556): the caller moves `a` to the device itself, as main.cu:911-915 */
LweSample_16 *convertBitToNumber(const LweSample *input, int bitSize, const TFheGateBootstrappingCloudKeySet *bk);
LweSample *convertNumberToBits(LweSample_16 *number, int bitSize, const TFheGateBootstrappingCloudKeySet *bk);
void freeLweSample_16(LweSample_16 *input);
void freeLweSample_16_gpu(LweSample_16 *sample);

#ifdef __cplusplus
}
#endif

#endif /* TFHE_COMPAT_H */
