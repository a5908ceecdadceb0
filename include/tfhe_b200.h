/*
 * tfhe_b200.h — C ABI of the B200-native TFHE gate-bootstrapping engine.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in the signatures
 * (streams are passed as void* holding a cudaStream_t, 0 = default stream).
 * Library: cpu-gpu-tfhe_b200/libtfhe_b200.so.
 *
 * This header is the batch ("flat") interface.  The drop-in replacements for
 * the reference's own entry points (bootsNAND ... bootsMUX,
 * tfhe_blindRotateAndExtract_FFT, tGswFFTExternMulToTLwe, lweKeySwitch, with
 * the reference's struct layouts) are declared in tfhe_compat.h and are thin
 * wrappers over the functions below.
 *
 * Data formats (int32 = Torus32, reference: gpuParallel/tfhe_core.h:28):
 *   sample       int32[n+1]              a[0..n) then b     (LweSample, lwesamples.h:18-29)
 *   extracted    int32[N+1]              a[0..N) then b     (LweSample of dimension N)
 *   accumulator  int32[k+1][N]           TLweSample polynomials (tlwe.h:47-52)
 *   bk (coef)    int32[n][kpl][k+1][N]   LweBootstrappingKey->bk[i].all_sample[r].a[j]
 *                                        (lwebootstrappingkey.h:10-16, tgsw.h:60-70)
 *   ks           int32[N][t][base][n+1]  LweKeySwitchKey->ks[i][j][h] (lwekeyswitch.h:11-28)
 * Batches are contiguous rows unless a stride (in words) is given.
 *
 * Every function returns 0 on success, non-zero on failure;
 * tfhe_b200_last_error() describes the last failure of the calling thread.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Threads: the compute calls of a context (gates, MUX, bootstrap, key switch, circuit runs) may be issued
 * from several host threads at once, each on its own stream (or all on the context's default stream, where
 * they simply queue); per-call scratch is stream ordered and the error string is per thread.  Only the launch
 * counter and the optional kernel timing are then approximate.  Loading keys, destroying the context and the
 * host-buffer calls (tfhe_b200_gate_host / _mux_host, which drive the context's copy streams) must not
 * overlap other calls on the same context; tfhe_compat.h serialises and coalesces the classic gates itself.
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfhe_b200_ctx tfhe_b200_ctx;

/* reference: new_default_gate_bootstrapping_parameters, tfhe_gate_bootstrapping.cu:25-49 */
typedef struct {
    int32_t n;          /* LWE dimension (500)            */
    int32_t N;          /* ring degree (1024, fixed)      */
    int32_t k;          /* TLWE mask size (1, fixed)      */
    int32_t l;          /* gadget length (2, fixed)       */
    int32_t Bgbit;      /* log2 gadget base (10, fixed)   */
    int32_t ks_t;       /* key-switch length (8)          */
    int32_t ks_basebit; /* key-switch log2 base (2)       */
} tfhe_b200_params;

/* gate ids: x = (0, c) + sa*ca + sb*cb, then bootstrap to +-1/8 (boot-gates.cu:98-397) */
enum {
    TFHE_B200_NAND = 0, /* boot-gates.cu:98  */
    TFHE_B200_OR = 1,   /* :124 */
    TFHE_B200_AND = 2,  /* :150 */
    TFHE_B200_XOR = 3,  /* :190 */
    TFHE_B200_XNOR = 4, /* :216 */
    TFHE_B200_NOR = 5,  /* :275 */
    TFHE_B200_ANDNY = 6,/* :301 */
    TFHE_B200_ANDYN = 7,/* :327 */
    TFHE_B200_ORNY = 8, /* :353 */
    TFHE_B200_ORYN = 9, /* :379 */
    TFHE_B200_NUM_GATES = 10,
    /* extension (not in the reference): carry operator g | (p & c) for mutually exclusive g, p,
     * one bootstrap of (1/8) + 2g + p + c; operands a = g, b = p, c = c (tfhe_b200_gate_op.c) */
    TFHE_B200_GPC = 10,
    /* extensions for carry-save (3:2 compressor) arithmetic: the two outputs of a FULL ADDER, one
     * bootstrap each, both in the same level.  With bits encoded as +-1/8: a + b + c is +-1/8 or +-3/8,
     * its sign is the majority (the carry); -2(a + b + c) mod 1 is +1/4 for an odd number of ones and
     * -1/4 for an even one (the sum bit).  Margins 1/8 and 1/4, input noise 3 and 12 sigma^2. */
    TFHE_B200_XOR3 = 11, /* a ^ b ^ c */
    TFHE_B200_MAJ = 12,  /* majority(a, b, c) */
    /* extension: sum bit of a parallel-prefix adder fused with its last carry operator,
     * a ^ (b | (c & d)) for mutually exclusive b, c (a = propagate of the bit, b / c = generate / propagate
     * of the group below it, d = carry into that group): one bootstrap of 3/8 + 2a + 2b + c + d; saves the
     * last level of every addition.  Four operands (tfhe_b200_gate_op.d). */
    TFHE_B200_SUMC = 13,
    TFHE_B200_NUM_GATES_EXT = 14
};

const char *tfhe_b200_last_error(void);
void tfhe_b200_default_params(tfhe_b200_params *p);
/* number of CUDA devices visible (0 if none / no driver) */
int tfhe_b200_device_count(void);

/* ---- context and keys -------------------------------------------------- */
int tfhe_b200_ctx_create(tfhe_b200_ctx **ctx, const tfhe_b200_params *p, int device);
void tfhe_b200_ctx_destroy(tfhe_b200_ctx *ctx);
/* Upload + convert keys from HOST flat arrays.  Replaces sendBootstrappingKeyToGPUCoalesceExt /
 * sendKeySwitchKeyToGPU_extendedOnePointer / sendKeySwitchBtoGPUOnePtr (main.cu:165,364,236)
 * and init_LweBootstrappingKeyFFT (lwe-bootstrapping-functions-fft.cu:60-89). */
int tfhe_b200_load_keys(tfhe_b200_ctx *ctx, const int32_t *bk_coef, const int32_t *ks);
/* Same, from DEVICE arrays (e.g. after an NCCL broadcast). */
int tfhe_b200_load_keys_device(tfhe_b200_ctx *ctx, const int32_t *d_bk_coef, const int32_t *d_ks, void *stream);
/* Bootstrapping key given in the reference's Fourier form: complex double
 * [n][kpl][k+1][N/2], value j = P(exp(-i*pi*(2j+1)/N)) of the torus polynomial scaled to
 * [-1/2,1/2) (TGswSampleFFT, tgsw.h:78-96; fft_processor_fftw.cu:158-167). Host pointer. */
int tfhe_b200_load_bk_fourier(tfhe_b200_ctx *ctx, const double *bkfft_ref);
int tfhe_b200_load_ks(tfhe_b200_ctx *ctx, const int32_t *ks);
/* bytes of device memory held by the keys */
size_t tfhe_b200_key_bytes(const tfhe_b200_ctx *ctx);

/* ---- device buffers for hosts without the CUDA headers (cgo / JNI / plain C) ----------------
 * Sample batches for the device-buffer entry points below can be allocated and moved with any CUDA
 * runtime in the process; these five calls are the same thing through this library's own runtime,
 * so that a host needs nothing but this header.  bytes = rows * (n + 1) * 4 for samples.  stream = NULL:
 * the default stream; the copies are synchronous with respect to the host. */
int tfhe_b200_device_alloc(tfhe_b200_ctx *ctx, void **d_ptr, size_t bytes);
int tfhe_b200_device_free(tfhe_b200_ctx *ctx, void *d_ptr);
int tfhe_b200_copy_to_device(tfhe_b200_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int tfhe_b200_copy_to_host(tfhe_b200_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);
int tfhe_b200_synchronize(tfhe_b200_ctx *ctx, void *stream);

/* ---- batched gates on DEVICE buffers (asynchronous on `stream`) ----------
 * out may alias an input (the reference's callers do this, Cipher.cu:373). */
int tfhe_b200_gate(tfhe_b200_ctx *ctx, int gate, int32_t *d_out, const int32_t *d_ca, const int32_t *d_cb,
                   int count, void *stream);
/* compound gate: out[0..count) = gate0(ca,cb), out[count..2count) = gate1(ca,cb) in ONE
 * bootstrap batch (bootsANDXOR_fullGPU_n_Bit_vector, boot-gates.cu:3027-3060) */
int tfhe_b200_gate2(tfhe_b200_ctx *ctx, int gate0, int gate1, int32_t *d_out, const int32_t *d_ca,
                    const int32_t *d_cb, int count, void *stream);
/* two independent gate batches in one launch: out[0..count) = gate0(a0,b0),
 * out[count..2count) = gate1(a1,b1) (bootsXORXOR_fullGPU_n_Bit_vector, boot-gates.cu:3062-3098) */
int tfhe_b200_gate_pair(tfhe_b200_ctx *ctx, int gate0, const int32_t *d_a0, const int32_t *d_b0, int gate1,
                        const int32_t *d_a1, const int32_t *d_b1, int32_t *d_out, int count, void *stream);
/* One run of `count` gates of one type.  Operand / result row of gate g is
 * base + idx[g] * stride when the (device) index array is given, else base + g * stride
 * (strides in int32 words; a stride of 0 broadcasts one sample).  out may alias a or b. */
typedef struct {
    int32_t gate;
    int32_t count;
    const int32_t *a;
    const int32_t *b;
    int32_t *out;
    int64_t stride_a, stride_b, stride_out;
    const int32_t *idx_a, *idx_b, *idx_out; /* optional, device memory */
    const int32_t *c;                       /* third operand, only for three- / four-input gates (else NULL) */
    int64_t stride_c;
    const int32_t *idx_c;
    const int32_t *d;                       /* fourth operand, only for TFHE_B200_SUMC (else NULL) */
    int64_t stride_d;
    const int32_t *idx_d;
} tfhe_b200_gate_op;
/* Up to TFHE_B200_MAX_RUNS runs in ONE bootstrap batch (one blind-rotate launch + one key-switch launch). */
#define TFHE_B200_MAX_RUNS 16
int tfhe_b200_gate_multi(tfhe_b200_ctx *ctx, const tfhe_b200_gate_op *ops, int nops, void *stream);
/* MUX(a,b,c) = a ? b : c  (bootsMUX, boot-gates.cu:407-448; bootsMUX_fullGPU_n_Bit :2987) */
int tfhe_b200_mux(tfhe_b200_ctx *ctx, int32_t *d_out, const int32_t *d_a, const int32_t *d_b,
                  const int32_t *d_c, int count, void *stream);
/* The same on rows of ONE sample array addressed through device index tables (circuit plans):
 * row idx_out[g] = MUX(row idx_a[g], row idx_b[g], row idx_c[g]); rows are `stride` words apart */
int tfhe_b200_mux_gather(tfhe_b200_ctx *ctx, int32_t *d_rows, int64_t stride, const int32_t *idx_a,
                         const int32_t *idx_b, const int32_t *idx_c, const int32_t *idx_out, int count,
                         void *stream);
/* row idx_out[g] = coef * (row idx_in[g]) + (0, cst): coef 1 = COPY, -1 = NOT, 0 = CONSTANT (no bootstrap) */
int tfhe_b200_linear_gather(tfhe_b200_ctx *ctx, int32_t *d_rows, int64_t stride, const int32_t *idx_in,
                            const int32_t *idx_out, int coef, int32_t cst, int count, void *stream);
/* bootsNOT :242, bootsCOPY :253, bootsCONSTANT :263 — no bootstrap */
int tfhe_b200_not(tfhe_b200_ctx *ctx, int32_t *d_out, const int32_t *d_ca, int count, void *stream);
int tfhe_b200_copy(tfhe_b200_ctx *ctx, int32_t *d_out, const int32_t *d_ca, int count, void *stream);
int tfhe_b200_constant(tfhe_b200_ctx *ctx, int32_t *d_out, int value, int count, void *stream);

/* ---- building blocks (DEVICE buffers) ------------------------------------ */
/* tfhe_bootstrap_woKS_FFT (lwe-bootstrapping-functions-fft.cu:1834): x[count][n+1] -> u[count][N+1] */
int tfhe_b200_bootstrap_woks(tfhe_b200_ctx *ctx, int32_t *d_u, const int32_t *d_x, int32_t mu, int count,
                             void *stream);
/* tfhe_bootstrap_FFT (:1884) */
int tfhe_b200_bootstrap(tfhe_b200_ctx *ctx, int32_t *d_out, const int32_t *d_x, int32_t mu, int count,
                        void *stream);
/* lweKeySwitch (lwe-keyswitch-functions.cu:955): u[count][N+1] -> out[count][n+1] */
int tfhe_b200_keyswitch(tfhe_b200_ctx *ctx, int32_t *d_out, const int32_t *d_u, int count, void *stream);
/* tfhe_blindRotate_FFT (:676): acc[count][k+1][N] in place, bara[count][n_iter] in [0,2N) */
int tfhe_b200_blind_rotate(tfhe_b200_ctx *ctx, int32_t *d_acc, const int32_t *d_bara, int n_iter, int count,
                           void *stream);
/* tfhe_blindRotateAndExtract_FFT (:1408): testvect[N] shared, barb[count], bara[count][n_iter] */
int tfhe_b200_blind_rotate_and_extract(tfhe_b200_ctx *ctx, int32_t *d_u, const int32_t *d_testvect,
                                       const int32_t *d_barb, const int32_t *d_bara, int n_iter, int count,
                                       void *stream);
/* tGswFFTExternMulToTLwe (tgsw-fft-operations.cu:124) with BK_{bk_index}: acc[count][k+1][N] in place */
int tfhe_b200_extern_mul(tfhe_b200_ctx *ctx, int32_t *d_acc, int bk_index, int count, void *stream);

/* ---- Cipher-level circuits (DEVICE buffers) --------------------------------
 * Integers are arrays of nbits samples, LSB first, two's complement (Cipher.cu:5-7); vectors
 * of `count` integers are contiguous.  A circuit is compiled once into a plan (all levels,
 * index tables and workspace on the device) and can be run any number of times. */
typedef struct tfhe_b200_circuit tfhe_b200_circuit;
/* a + b mod 2^nbits for count pairs.  mode 0: bit-wise ripple carry (taskLevelParallelAdd_bitwise
 * main.cu:821, _vector_coalInput :1138, Cipher::operator+ Cipher.cu:334); mode 1: number-wise
 * (taskLevelParallelAdd main.cu:619).  Operands: a[count][nbits], b[count][nbits]. */
tfhe_b200_circuit *tfhe_b200_circuit_add(tfhe_b200_ctx *ctx, int nbits, int count, int mode);
/* Adder used inside the multiplier / matrix-multiply trees.  RIPPLE is the reference's schedule
 * (3*nbits-3 levels per addition); PREFIX is a Kogge-Stone adder whose carry operator is one
 * three-input bootstrap (TFHE_B200_GPC) and whose last level yields the sum bits themselves
 * (TFHE_B200_SUMC): 1 + ceil(log2(nbits-1)) levels per addition (SURVEY
 * 8f rank 4, not in the reference).  tfhe_b200_circuit_add mode 2 is the PREFIX adder. */
enum { TFHE_B200_ADDER_RIPPLE = 0, TFHE_B200_ADDER_PREFIX = 1,
       /* multipliers / matrix products only: all partial-product bits of a result go through a
        * carry-save (Wallace) tree of full adders — one level per 3:2 compression, two bootstraps per
        * full adder (TFHE_B200_XOR3 / TFHE_B200_MAJ) — down to two rows, then ONE parallel-prefix
        * addition.  32 bits: 15 levels and 1.7 k gates instead of 31 levels and 10 k (prefix adders)
        * or 466 levels (the reference's schedule, multiplyLweSamples main.cu:1483-1579). */
       TFHE_B200_ADDER_CARRY_SAVE = 2 };
/* a * b mod 2^nbits for count pairs (multiplyLweSamples main.cu:1483, BOOTS_vectorMultiplication
 * :1746, Cipher::operator* Cipher.cu:83) */
tfhe_b200_circuit *tfhe_b200_circuit_mul(tfhe_b200_ctx *ctx, int nbits, int count);
tfhe_b200_circuit *tfhe_b200_circuit_mul_ex(tfhe_b200_ctx *ctx, int nbits, int count, int adder);
/* full 2*nbits-bit products (isDoublePrecision of BOOTS_vectorMultiplication, main.cu:1746): schoolbook,
 * and one level of Karatsuba (karatMasterSuba, main.cu:1866; nbits even); output 2*nbits rows per pair */
tfhe_b200_circuit *tfhe_b200_circuit_mul_full(tfhe_b200_ctx *ctx, int nbits, int count, int adder);
tfhe_b200_circuit *tfhe_b200_circuit_mul_karatsuba(tfhe_b200_ctx *ctx, int nbits, int count, int adder);
/* C[rows][cols] = A[rows][inner] * B[inner][cols], nbits-bit elements mod 2^nbits
 * (BOOTS_matrixMultiplication main.cu:2342; cpu/cloud.cpp:390-408) */
tfhe_b200_circuit *tfhe_b200_circuit_matmul(tfhe_b200_ctx *ctx, int rows, int inner, int cols, int nbits);
tfhe_b200_circuit *tfhe_b200_circuit_matmul_ex(tfhe_b200_ctx *ctx, int rows, int inner, int cols, int nbits,
                                               int adder);
/* the same product of square n x n matrices by Cannon's schedule (BOOTS_CannonsAlgo, main.cu:2590) */
tfhe_b200_circuit *tfhe_b200_circuit_matmul_cannon(tfhe_b200_ctx *ctx, int n, int nbits, int adder);
/* ---- the rest of the Cipher arithmetic (Cipher.cu:237-630) ------------------
 * Same plan machinery; `adder` as above.  All operands have count numbers of nbits bits. */
/* a - b (operator-, Cipher.cu:329) and -a (twosComplement, :286) */
tfhe_b200_circuit *tfhe_b200_circuit_sub(tfhe_b200_ctx *ctx, int nbits, int count, int adder);
tfhe_b200_circuit *tfhe_b200_circuit_neg(tfhe_b200_ctx *ctx, int nbits, int count);
/* comparisons, ONE result bit per pair (operator> :561, operator<= :574, operator== :600) */
enum {
    TFHE_B200_CMP_GT = 0, TFHE_B200_CMP_LE = 1, TFHE_B200_CMP_LT = 2, TFHE_B200_CMP_GE = 3,
    TFHE_B200_CMP_EQ = 4, TFHE_B200_CMP_NE = 5
};
tfhe_b200_circuit *tfhe_b200_circuit_compare(tfhe_b200_ctx *ctx, int nbits, int count, int op, int is_signed);
/* min / max (minimum, :301) */
tfhe_b200_circuit *tfhe_b200_circuit_minmax(tfhe_b200_ctx *ctx, int nbits, int count, int want_max, int is_signed);
/* sel ? a : b; operands: sel[count] (one bit per number), a[count][nbits], b[count][nbits] */
tfhe_b200_circuit *tfhe_b200_circuit_select(tfhe_b200_ctx *ctx, int nbits, int count);
/* |a| for two's complement a (absolute, :469) */
tfhe_b200_circuit *tfhe_b200_circuit_abs(tfhe_b200_ctx *ctx, int nbits, int count, int adder);
/* shifts by a public amount (innerLeftShift :215, rightShift :237); no bootstrap */
enum { TFHE_B200_SHIFT_LEFT = 0, TFHE_B200_SHIFT_RIGHT_LOGICAL = 1, TFHE_B200_SHIFT_RIGHT_ARITH = 2 };
tfhe_b200_circuit *tfhe_b200_circuit_shift(tfhe_b200_ctx *ctx, int nbits, int count, int amount, int kind);
/* a / b (operator/ :494, divInternal :515, addSign :543): output per pair = quotient[nbits] then
 * remainder[nbits] of |a| / |b| (is_signed) or of a / b (unsigned) */
tfhe_b200_circuit *tfhe_b200_circuit_div(tfhe_b200_ctx *ctx, int nbits, int count, int is_signed, int adder);
void tfhe_b200_circuit_destroy(tfhe_b200_circuit *c);
int tfhe_b200_circuit_levels(const tfhe_b200_circuit *c);      /* sequential bootstrap batches */
/* bootstraps in batch `level` (0 .. levels - 1): the width the latency of that level depends on */
long long tfhe_b200_circuit_level_gates(const tfhe_b200_circuit *c, int level);
long long tfhe_b200_circuit_gates(const tfhe_b200_circuit *c); /* bootstrapped gates per run   */
int tfhe_b200_circuit_operands(const tfhe_b200_circuit *c);
int tfhe_b200_circuit_operand_rows(const tfhe_b200_circuit *c, int operand);
int tfhe_b200_circuit_output_rows(const tfhe_b200_circuit *c);
/* operands[o]: device array of tfhe_b200_circuit_operand_rows(c, o) samples */
int tfhe_b200_circuit_run(tfhe_b200_circuit *c, int32_t *d_out, const int32_t *const *operands, void *stream);
/* The launch sequence of a plan (all its levels) is captured into a CUDA graph at the first run on a
 * non-default stream and replayed afterwards (one graph launch instead of two kernel launches per
 * level).  enable = 0 switches the plan back to direct launches.  Returns the previous setting. */
int tfhe_b200_circuit_set_graph(tfhe_b200_circuit *c, int enable);
/* 1 if the last run of the plan replayed a captured graph */
int tfhe_b200_circuit_used_graph(const tfhe_b200_circuit *c);
/* K INDEPENDENT plans of one context run together, level by level: the gate runs of level j of all
 * plans share ONE blind-rotate + ONE key-switch launch (up to TFHE_B200_MAX_RUNS runs per launch), so
 * K narrow circuits cost about as much as one (a level of a lone 16-bit adder occupies a dozen SMs for
 * a full bootstrap latency).  The reference's counterpart is the vLength dimension of its vector
 * circuits (taskLevelParallelAdd_bitwise_vector_coalInput, main.cu:1138-1302), which only batches
 * copies of the SAME circuit.  operands[p][o]: operand o of plan p; d_outs[p]: its result rows. */
int tfhe_b200_circuit_run_many(tfhe_b200_circuit *const *plans, int nplans, int32_t *const *d_outs,
                               const int32_t *const *const *operands, void *stream);
/* Host-only check of a plan on PLAINTEXT bits (one int per sample row); needs no GPU and accepts
 * plans built with ctx == NULL.  This is schedule verification, not a compute path. */
int tfhe_b200_circuit_simulate(const tfhe_b200_circuit *c, int32_t *out_bits, const int32_t *const *operand_bits);
int tfhe_b200_ctx_words(const tfhe_b200_ctx *ctx); /* n + 1 */
int tfhe_b200_ctx_device(const tfhe_b200_ctx *ctx);

/* Key generation ON THE GPU of ctx (new_random_gate_bootstrapping_secret_keyset,
 * tfhe_gate_bootstrapping.cu:57-68): the secret bits are drawn on the host and returned, the 2000
 * TLWE encryptions of the bootstrapping key and the 24576 key-switch samples are generated in
 * device memory (Philox counter-based streams) and loaded into ctx without a host round trip.
 * bk_out / ks_out (host, flat formats) may be NULL. */
int tfhe_b200_keygen_device(tfhe_b200_ctx *ctx, const tfhe_b200_params *p, uint64_t seed, double alpha_lwe,
                            double alpha_bk, int32_t *lwe_key, int32_t *tlwe_key, int32_t *bk_out, int32_t *ks_out);

/* ---- key and ciphertext FILES of stock TFHE clients (host only) -------------
 * The serialisation of gpuParallel/tfhe_io.cu (text parameter sections + binary payloads):
 * cloud.key / secret.key as written by export_tfheGateBootstrapping{Cloud,Secret}KeySet_toFile
 * (tfhe_io.cu:1099-1103, 1160-1166; producer cpu/main.cpp:26-71, consumer cpu/cloud.cpp:138-161)
 * and ciphertext records as written by export_gate_bootstrapping_ciphertext_toFile (:90-108).
 * Payload order equals the flat formats above.  alphas4 = {lwe alpha_min, alpha_max, tlwe
 * alpha_min, alpha_max}; variances2 = {bootstrapping key, key-switch key} (the single variance
 * each section stores).  Any output pointer may be NULL (e.g. all NULL but p: read the header). */
int tfhe_b200_file_read_cloud_key(const char *path, tfhe_b200_params *p, double *alphas4, double *variances2,
                                  int32_t *bk_coef, int32_t *ks);
int tfhe_b200_file_write_cloud_key(const char *path, const tfhe_b200_params *p, const double *alphas4,
                                   const double *variances2, const int32_t *bk_coef, const int32_t *ks);
int tfhe_b200_file_read_secret_key(const char *path, tfhe_b200_params *p, double *alphas4, double *variances2,
                                   int32_t *bk_coef, int32_t *ks, int32_t *lwe_key, int32_t *tlwe_key);
int tfhe_b200_file_write_secret_key(const char *path, const tfhe_b200_params *p, const double *alphas4,
                                    const double *variances2, const int32_t *bk_coef, const int32_t *ks,
                                    const int32_t *lwe_key, const int32_t *tlwe_key);
/* ciphertext files: `count` records of dimension n into / from samples[count][n+1] */
long tfhe_b200_file_count_ciphertexts(const char *path, int n);
int tfhe_b200_file_read_ciphertexts(const char *path, int n, int32_t *samples, double *variances, int count);
int tfhe_b200_file_write_ciphertexts(const char *path, int n, const int32_t *samples, const double *variances,
                                     int count, int append);
const char *tfhe_b200_file_last_error(void);

/* ---- HOST-buffer convenience (synchronous; copies in and out) -------------
 * Batches of two or more 16-wave chunks (16 * 4 * #SMs gates) are pipelined: the host-to-device
 * copy of chunk i+1 and the device-to-host copy of chunk i-1 run under the kernels of chunk i.
 * Page-locked host buffers make the copies asynchronous; pageable ones work, without overlap. */
int tfhe_b200_gate_host(tfhe_b200_ctx *ctx, int gate, int32_t *out, const int32_t *ca, const int32_t *cb,
                        int count);
int tfhe_b200_mux_host(tfhe_b200_ctx *ctx, int32_t *out, const int32_t *a, const int32_t *b, const int32_t *c,
                       int count);

/* ---- several GPUs in ONE process (C / C++ hosts) ------------------------------
 * The reference uses device 0 only (boot-gates.cu:3344).  A multi-GPU handle owns one context per
 * device; gates are independent, so a host batch is cut into contiguous shards (sizes differ by at
 * most one), one per device, driven by one host thread each; nothing is exchanged on the data path.
 * tfhe_b200_multi_load_keys uploads the keys to the first device once and copies them device to
 * device to the others (cudaMemcpyPeerAsync, NVLink where peer access exists).
 * devices == NULL or ndevices <= 0: all visible devices.  The same device may be listed twice
 * (two contexts on one GPU).  Errors: non-zero return, tfhe_b200_multi_last_error(). */
typedef struct tfhe_b200_multi tfhe_b200_multi;
int tfhe_b200_multi_create(tfhe_b200_multi **m, const tfhe_b200_params *p, const int *devices, int ndevices);
void tfhe_b200_multi_destroy(tfhe_b200_multi *m);
int tfhe_b200_multi_devices(const tfhe_b200_multi *m);
tfhe_b200_ctx *tfhe_b200_multi_ctx(tfhe_b200_multi *m, int index); /* per-device context (borrowed) */
int tfhe_b200_multi_load_keys(tfhe_b200_multi *m, const int32_t *bk_coef, const int32_t *ks);
int tfhe_b200_multi_gate_host(tfhe_b200_multi *m, int gate, int32_t *out, const int32_t *ca, const int32_t *cb,
                              long long count);
int tfhe_b200_multi_mux_host(tfhe_b200_multi *m, int32_t *out, const int32_t *a, const int32_t *b, const int32_t *c,
                             long long count);
unsigned long long tfhe_b200_multi_launch_count(const tfhe_b200_multi *m);
const char *tfhe_b200_multi_last_error(void);

/* ---- client side: key generation, encryption, decryption (HOST, CPU) -------
 * These belong to the key owner, not to the evaluation path; the reference runs them on
 * the host as well (new_random_gate_bootstrapping_secret_keyset tfhe_gate_bootstrapping.cu:57-68,
 * bootsSymEncrypt :114, bootsSymDecrypt :122, lwePhase lwe-functions.cu:72). */
/* RANDOMNESS: every secret bit, mask word and noise term is ChaCha20 keystream (csrc/csprng.h).
 * seed == 0: the 256-bit key comes from the operating system (getrandom) — use this for real keys and
 * ciphertexts.  seed != 0: the key is derived from the seed: REPRODUCIBLE material for tests and
 * benchmarks, NOT secure (64 bits, known to the caller). */
size_t tfhe_b200_bk_words(const tfhe_b200_params *p);
size_t tfhe_b200_ks_words(const tfhe_b200_params *p);
void tfhe_b200_default_noise(double *alpha_lwe, double *alpha_bk);
int tfhe_b200_keygen(const tfhe_b200_params *p, uint64_t seed, double alpha_lwe, double alpha_bk,
                     int32_t *lwe_key, int32_t *tlwe_key, int32_t *bk, int32_t *ks);
int tfhe_b200_encrypt_bits(const tfhe_b200_params *p, const int32_t *lwe_key, uint64_t seed, double alpha,
                           const int32_t *bits, int count, int32_t *out);
int tfhe_b200_decrypt_bits(const tfhe_b200_params *p, const int32_t *lwe_key, const int32_t *samples,
                           int count, int32_t *bits_out);
int tfhe_b200_phases(const int32_t *key, int n, const int32_t *samples, int count, int32_t *phases_out);

/* ---- introspection (tests, bench) ---------------------------------------- */
/* kernels launched by this context so far (each = one launch of one of this library's kernels) */
unsigned long long tfhe_b200_launch_count(const tfhe_b200_ctx *ctx);
/* adds n to the counter (circuit plans replaying a captured CUDA graph account their launches here) */
void tfhe_b200_count_launches(tfhe_b200_ctx *ctx, unsigned long long n);
int tfhe_b200_sm_count(const tfhe_b200_ctx *ctx);
/* per-kernel device time (CUDA events on the launching stream) of the gate calls issued while
 * timing is enabled: total blind-rotate ms, total key-switch ms, number of gate calls */
int tfhe_b200_set_timing(tfhe_b200_ctx *ctx, int enable);
int tfhe_b200_get_timing(tfhe_b200_ctx *ctx, double *blind_rotate_ms, double *keyswitch_ms, int *calls);
/* How this build converts the fp64 products back to Torus32: 0 = round to nearest (default; every
 * blind-rotation step equals the exact integer negacyclic product, tests/test_gpu_exact.py),
 * 1 = truncation toward zero like the reference's Torus32(int64_t(x)), fft_processor_fftw.cu:177
 * (the libtfhe_b200_trunc.so build, -DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1) */
int tfhe_b200_conversion_mode(void);
/* fp64 FMA peak of `device` in TFLOP/s, measured live: best single launch and the mean over
 * ~0.4 s of back-to-back launches (roofline denominator of the blind-rotation kernel) */
int tfhe_b200_measure_fp64_peak(int device, double *burst_tflops, double *sustained_tflops);

#ifdef __cplusplus
}
#endif

#endif /* TFHE_B200_H */
