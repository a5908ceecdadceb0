// Drop-in replacements for the reference's entry points (include/tfhe_compat.h): they walk
// the reference's pointer-rich structs, keep one GPU context per key object (created on first
// use, thread-safe) and call the flat C ABI.  Failures abort, like die_dramatically()
// (tfhe_gate_bootstrapping.cu:11-15) and the CUDA error macros of boot-gates.cu:33-86.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/tfhe_compat.h"

namespace {

std::mutex g_mu;
std::map<const void *, tfhe_b200_ctx *> g_ctx;  // key object -> context

[[noreturn]] void die(const char *what) {
    fprintf(stderr, "tfhe_b200: %s: %s\n", what, tfhe_b200_last_error());
    abort();
}

#define OK(call, what)      \
    do {                    \
        if (call) die(what); \
    } while (0)

void cu(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        fprintf(stderr, "tfhe_b200: %s: %s\n", what, cudaGetErrorString(e));
        abort();
    }
}

tfhe_b200_params params_of(int n, const TGswParams *gp, int ks_t, int ks_basebit) {
    tfhe_b200_params p;
    p.n = n;
    p.N = gp->tlwe_params->N;
    p.k = gp->tlwe_params->k;
    p.l = gp->l;
    p.Bgbit = gp->Bgbit;
    p.ks_t = ks_t;
    p.ks_basebit = ks_basebit;
    return p;
}

// LweKeySwitchKey (lwekeyswitch.h:11-28) -> flat [N][t][base][n+1]
std::vector<int32_t> flatten_ks(const LweKeySwitchKey *ks) {
    const int n = ks->out_params->n, base = ks->base;
    std::vector<int32_t> out((size_t) ks->n * ks->t * base * (n + 1));
    for (int i = 0; i < ks->n; i++)
        for (int j = 0; j < ks->t; j++)
            for (int h = 0; h < base; h++) {
                const LweSample *s = &ks->ks[i][j][h];
                int32_t *dst = out.data() + (((size_t) i * ks->t + j) * base + h) * (n + 1);
                memcpy(dst, s->a, sizeof(int32_t) * n);
                dst[n] = s->b;
            }
    return out;
}

// TGswSample[n] coefficient domain (tgsw.h:60-65) -> flat [n][kpl][k+1][N]
std::vector<int32_t> flatten_bk(const TGswSample *bk, int n, const TGswParams *gp) {
    const int N = gp->tlwe_params->N, k = gp->tlwe_params->k, kpl = gp->kpl;
    std::vector<int32_t> out((size_t) n * kpl * (k + 1) * N);
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++)
                memcpy(out.data() + (((size_t) i * kpl + r) * (k + 1) + j) * N, bk[i].all_sample[r].a[j].coefsT,
                       sizeof(int32_t) * N);
    return out;
}

// TGswSampleFFT[n] (tgsw.h:78-84; LagrangeHalfCPolynomial_IMPL lagrangehalfc_impl.h:45-52:
// data -> N/2 complex<double>) -> flat complex [n][kpl][k+1][N/2]
std::vector<double> flatten_bkfft(const TGswSampleFFT *bk, int n, const TGswParams *gp) {
    const int Ns2 = gp->tlwe_params->N / 2, k = gp->tlwe_params->k, kpl = gp->kpl;
    std::vector<double> out((size_t) n * kpl * (k + 1) * Ns2 * 2);
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++)
                memcpy(out.data() + (((size_t) i * kpl + r) * (k + 1) + j) * Ns2 * 2, bk[i].all_samples[r].a[j].data,
                       sizeof(double) * 2 * Ns2);
    return out;
}

tfhe_b200_ctx *lookup(const void *key) {
    auto it = g_ctx.find(key);
    return it == g_ctx.end() ? nullptr : it->second;
}

int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return 0;
    return d;
}

// context for a whole cloud key set (bootstrapping + key switch)
tfhe_b200_ctx *ctx_for_cloud(const TFheGateBootstrappingCloudKeySet *ck) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (tfhe_b200_ctx *c = lookup(ck)) return c;
    const int n = ck->params->in_out_params->n;
    const tfhe_b200_params p = params_of(n, ck->params->tgsw_params, ck->params->ks_t, ck->params->ks_basebit);
    tfhe_b200_ctx *c = nullptr;
    OK(tfhe_b200_ctx_create(&c, &p, current_device()), "context creation");
    if (ck->bk) {
        const std::vector<int32_t> bk = flatten_bk(ck->bk->bk, n, ck->params->tgsw_params);
        const std::vector<int32_t> ks = flatten_ks(ck->bk->ks);
        OK(tfhe_b200_load_keys(c, bk.data(), ks.data()), "key upload");
    } else {
        const std::vector<double> bk = flatten_bkfft(ck->bkFFT->bkFFT, n, ck->params->tgsw_params);
        const std::vector<int32_t> ks = flatten_ks(ck->bkFFT->ks);
        OK(tfhe_b200_load_bk_fourier(c, bk.data()), "key upload");
        OK(tfhe_b200_load_ks(c, ks.data()), "key upload");
    }
    g_ctx[ck] = c;
    return c;
}

// context for a Fourier bootstrapping key (+ its key-switch key)
tfhe_b200_ctx *ctx_for_bkfft(const LweBootstrappingKeyFFT *bk) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (tfhe_b200_ctx *c = lookup(bk)) return c;
    const int n = bk->in_out_params->n;
    const tfhe_b200_params p = params_of(n, bk->bk_params, bk->ks->t, bk->ks->basebit);
    tfhe_b200_ctx *c = nullptr;
    OK(tfhe_b200_ctx_create(&c, &p, current_device()), "context creation");
    const std::vector<double> f = flatten_bkfft(bk->bkFFT, n, bk->bk_params);
    const std::vector<int32_t> ks = flatten_ks(bk->ks);
    OK(tfhe_b200_load_bk_fourier(c, f.data()), "key upload");
    OK(tfhe_b200_load_ks(c, ks.data()), "key upload");
    g_ctx[bk] = c;
    return c;
}

// context for a bare array of n TGSW samples in Fourier form (no key switch)
tfhe_b200_ctx *ctx_for_tgsw(const TGswSampleFFT *bk, int n, const TGswParams *gp) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (n < 1) n = 1;  // n = 0 (no iterations) still needs a context for the integer stages
    const void *key = (const char *) bk + 1;  // distinct from a LweBootstrappingKeyFFT at the same address
    if (tfhe_b200_ctx *c = lookup(key)) {
        if (tfhe_b200_ctx_words(c) == n + 1) return c;
        tfhe_b200_ctx_destroy(c);
        g_ctx.erase(key);
    }
    const tfhe_b200_params p = params_of(n, gp, 8, 2);
    tfhe_b200_ctx *c = nullptr;
    OK(tfhe_b200_ctx_create(&c, &p, current_device()), "context creation");
    const std::vector<double> f = flatten_bkfft(bk, n, gp);
    OK(tfhe_b200_load_bk_fourier(c, f.data()), "key upload");
    g_ctx[key] = c;
    return c;
}

tfhe_b200_ctx *ctx_for_ks(const LweKeySwitchKey *ks) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (tfhe_b200_ctx *c = lookup(ks)) return c;
    tfhe_b200_params p;
    tfhe_b200_default_params(&p);
    p.n = ks->out_params->n;
    p.ks_t = ks->t;
    p.ks_basebit = ks->basebit;
    if (ks->n != p.N * p.k) die("lweKeySwitch: input dimension must be N*k = 1024");
    tfhe_b200_ctx *c = nullptr;
    OK(tfhe_b200_ctx_create(&c, &p, current_device()), "context creation");
    const std::vector<int32_t> flat = flatten_ks(ks);
    OK(tfhe_b200_load_ks(c, flat.data()), "key upload");
    g_ctx[ks] = c;
    return c;
}

// small synchronous device buffer helpers for the single-sample entry points
struct DevBuf {
    int32_t *p = nullptr;
    explicit DevBuf(size_t words) { cu(cudaMalloc(&p, words * sizeof(int32_t)), "cudaMalloc"); }
    ~DevBuf() { cudaFree(p); }
    void up(const int32_t *src, size_t words, size_t off = 0) {
        cu(cudaMemcpy(p + off, src, words * sizeof(int32_t), cudaMemcpyHostToDevice), "H2D");
    }
    void down(int32_t *dst, size_t words, size_t off = 0) {
        cu(cudaMemcpy(dst, p + off, words * sizeof(int32_t), cudaMemcpyDeviceToHost), "D2H");
    }
};

void put_sample(DevBuf &d, size_t row, const LweSample *s, int n) {
    d.up(s->a, n, row * (n + 1));
    d.up(&s->b, 1, row * (n + 1) + n);
}

void get_sample(DevBuf &d, size_t row, LweSample *s, int n) {
    d.down(s->a, n, row * (n + 1));
    d.down(&s->b, 1, row * (n + 1) + n);
    s->current_variance = 0.;  // bookkeeping only; the reference's GPU path ignores it too (boot-gates.cu:2866)
}

void classic_gate(int gate, LweSample *result, const LweSample *ca, const LweSample *cb,
                  const TFheGateBootstrappingCloudKeySet *ck) {
    tfhe_b200_ctx *c = ctx_for_cloud(ck);
    const int n = ck->params->in_out_params->n;
    DevBuf d(3 * (size_t) (n + 1));
    put_sample(d, 0, ca, n);
    put_sample(d, 1, cb, n);
    OK(tfhe_b200_gate(c, gate, d.p + 2 * (n + 1), d.p, d.p + (n + 1), 1, nullptr), "gate");
    cu(cudaStreamSynchronize(nullptr), "gate");
    get_sample(d, 2, result, n);
}

// ---- LweSample_16 (a on device, b on host) <-> rows of n+1 words ---------------------------

__global__ void pack16_kernel(int32_t *rows, const int *a, const int *b, int count, int n) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long) count * (n + 1)) return;
    const int g = (int) (t / (n + 1)), c = (int) (t % (n + 1));
    rows[t] = c < n ? a[(size_t) g * n + c] : b[g];
}

__global__ void unpack16_kernel(int *a, int *b, const int32_t *rows, int count, int n) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long) count * (n + 1)) return;
    const int g = (int) (t / (n + 1)), c = (int) (t % (n + 1));
    if (c < n) a[(size_t) g * n + c] = rows[t];
    else b[g] = rows[t];
}

__global__ void negate_kernel(int32_t *rows, long long total) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) rows[t] = (int32_t) (0u - (uint32_t) rows[t]);
}

struct Rows16 {
    int32_t *rows = nullptr;
    int *db = nullptr;
    int count, n;
    Rows16(int count_, int n_) : count(count_), n(n_) {
        cu(cudaMalloc(&rows, (size_t) count * (n + 1) * sizeof(int32_t)), "cudaMalloc");
        cu(cudaMalloc(&db, (size_t) count * sizeof(int)), "cudaMalloc");
    }
    ~Rows16() {
        cudaFree(rows);
        cudaFree(db);
    }
    void pack(const LweSample_16 *s) {
        cu(cudaMemcpy(db, s->b, (size_t) count * sizeof(int), cudaMemcpyHostToDevice), "H2D");
        const long long total = (long long) count * (n + 1);
        pack16_kernel<<<(unsigned) ((total + 255) / 256), 256>>>(rows, s->a, db, count, n);
    }
    void unpack(LweSample_16 *s) {
        const long long total = (long long) count * (n + 1);
        unpack16_kernel<<<(unsigned) ((total + 255) / 256), 256>>>(s->a, db, rows, count, n);
        cu(cudaMemcpy(s->b, db, (size_t) count * sizeof(int), cudaMemcpyDeviceToHost), "D2H");
    }
};

void batched_gate2(int g0, int g1, LweSample_16 *result, const LweSample_16 *a0, const LweSample_16 *b0,
                   const LweSample_16 *a1, const LweSample_16 *b1, int count, void *handle) {
    tfhe_b200_ctx *c = (tfhe_b200_ctx *) handle;
    if (!c) die("null key handle (pass tfhe_b200_keys_to_gpu(bk) as bkGPU)");
    const int n = tfhe_b200_ctx_words(c) - 1;
    const bool two = (g1 >= 0);
    Rows16 ra0(count, n), rb0(count, n), out((two ? 2 : 1) * count, n);
    ra0.pack(a0);
    rb0.pack(b0);
    if (!two) {
        OK(tfhe_b200_gate(c, g0, out.rows, ra0.rows, rb0.rows, count, nullptr), "gate batch");
    } else if (a1 == a0 && b1 == b0) {
        OK(tfhe_b200_gate2(c, g0, g1, out.rows, ra0.rows, rb0.rows, count, nullptr), "gate batch");
    } else {
        Rows16 ra1(count, n), rb1(count, n);
        ra1.pack(a1);
        rb1.pack(b1);
        OK(tfhe_b200_gate_pair(c, g0, ra0.rows, rb0.rows, g1, ra1.rows, rb1.rows, out.rows, count, nullptr),
           "gate batch");
        cu(cudaStreamSynchronize(nullptr), "gate batch");
    }
    out.unpack(result);
}

}  // namespace

extern "C" {

void bootsNAND(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_NAND, r, a, b, bk); }
void bootsOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_OR, r, a, b, bk); }
void bootsAND(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_AND, r, a, b, bk); }
void bootsXOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_XOR, r, a, b, bk); }
void bootsXNOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_XNOR, r, a, b, bk); }
void bootsNOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_NOR, r, a, b, bk); }
void bootsANDNY(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ANDNY, r, a, b, bk); }
void bootsANDYN(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ANDYN, r, a, b, bk); }
void bootsORNY(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ORNY, r, a, b, bk); }
void bootsORYN(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ORYN, r, a, b, bk); }

void bootsMUX(LweSample *result, const LweSample *a, const LweSample *b, const LweSample *cc,
              const TFheGateBootstrappingCloudKeySet *ck) {
    tfhe_b200_ctx *c = ctx_for_cloud(ck);
    const int n = ck->params->in_out_params->n;
    DevBuf d(4 * (size_t) (n + 1));
    put_sample(d, 0, a, n);
    put_sample(d, 1, b, n);
    put_sample(d, 2, cc, n);
    OK(tfhe_b200_mux(c, d.p + 3 * (n + 1), d.p, d.p + (n + 1), d.p + 2 * (n + 1), 1, nullptr), "mux");
    cu(cudaStreamSynchronize(nullptr), "mux");
    get_sample(d, 3, result, n);
}

// bootsNOT / COPY / CONSTANT do not bootstrap (boot-gates.cu:242-267): plain host arithmetic on
// one sample, exactly as the reference does.
void bootsNOT(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    for (int i = 0; i < n; i++) result->a[i] = (Torus32) (0u - (uint32_t) ca->a[i]);
    result->b = (Torus32) (0u - (uint32_t) ca->b);
    result->current_variance = ca->current_variance;
}

void bootsCOPY(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    if (result != ca) memmove(result->a, ca->a, sizeof(Torus32) * n);
    result->b = ca->b;
    result->current_variance = ca->current_variance;
}

void bootsCONSTANT(LweSample *result, int value, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    for (int i = 0; i < n; i++) result->a[i] = 0;
    result->b = value ? 0x20000000 : -0x20000000;
    result->current_variance = 0.;
}

void tfhe_blindRotate_FFT(TLweSample *accum, const TGswSampleFFT *bk, const int *bara, const int n,
                          const TGswParams *bk_params) {
    tfhe_b200_ctx *c = ctx_for_tgsw(bk, n, bk_params);
    const int N = bk_params->tlwe_params->N, k = bk_params->tlwe_params->k;
    DevBuf d((size_t) (k + 1) * N + n);
    for (int j = 0; j <= k; j++) d.up(accum->a[j].coefsT, N, (size_t) j * N);
    d.up(bara, n, (size_t) (k + 1) * N);
    OK(tfhe_b200_blind_rotate(c, d.p, d.p + (size_t) (k + 1) * N, n, 1, nullptr), "blind rotate");
    cu(cudaStreamSynchronize(nullptr), "blind rotate");
    for (int j = 0; j <= k; j++) d.down(accum->a[j].coefsT, N, (size_t) j * N);
}

void tfhe_blindRotateAndExtract_FFT(LweSample *result, const TorusPolynomial *v, const TGswSampleFFT *bk,
                                    const int barb, const int *bara, const int n, const TGswParams *bk_params) {
    tfhe_b200_ctx *c = ctx_for_tgsw(bk, n, bk_params);
    const int N = bk_params->tlwe_params->N, k = bk_params->tlwe_params->k;
    DevBuf d((size_t) N + 1 + n + (size_t) k * N + 1);
    d.up(v->coefsT, N, 0);
    d.up(&barb, 1, N);
    d.up(bara, n, N + 1);
    int32_t *u = d.p + N + 1 + n;
    OK(tfhe_b200_blind_rotate_and_extract(c, u, d.p, d.p + N, d.p + N + 1, n, 1, nullptr), "blind rotate");
    cu(cudaStreamSynchronize(nullptr), "blind rotate");
    d.down(result->a, (size_t) k * N, N + 1 + n);
    d.down(&result->b, 1, N + 1 + n + (size_t) k * N);
    result->current_variance = 0.;
}

void tfhe_bootstrap_woKS_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x) {
    tfhe_b200_ctx *c = ctx_for_bkfft(bk);
    const int n = bk->in_out_params->n, Nk = bk->extract_params->n;
    DevBuf d((size_t) (n + 1) + Nk + 1);
    put_sample(d, 0, x, n);
    OK(tfhe_b200_bootstrap_woks(c, d.p + (n + 1), d.p, mu, 1, nullptr), "bootstrap");
    cu(cudaStreamSynchronize(nullptr), "bootstrap");
    d.down(result->a, Nk, n + 1);
    d.down(&result->b, 1, (size_t) n + 1 + Nk);
    result->current_variance = 0.;
}

void tfhe_bootstrap_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x) {
    tfhe_b200_ctx *c = ctx_for_bkfft(bk);
    const int n = bk->in_out_params->n;
    DevBuf d(2 * (size_t) (n + 1));
    put_sample(d, 0, x, n);
    OK(tfhe_b200_bootstrap(c, d.p + (n + 1), d.p, mu, 1, nullptr), "bootstrap");
    cu(cudaStreamSynchronize(nullptr), "bootstrap");
    get_sample(d, 1, result, n);
}

void tGswFFTExternMulToTLwe(TLweSample *accum, const TGswSampleFFT *gsw, const TGswParams *params) {
    tfhe_b200_ctx *c = ctx_for_tgsw(gsw, 1, params);
    const int N = params->tlwe_params->N, k = params->tlwe_params->k;
    DevBuf d((size_t) (k + 1) * N);
    for (int j = 0; j <= k; j++) d.up(accum->a[j].coefsT, N, (size_t) j * N);
    OK(tfhe_b200_extern_mul(c, d.p, 0, 1, nullptr), "external product");
    cu(cudaStreamSynchronize(nullptr), "external product");
    for (int j = 0; j <= k; j++) d.down(accum->a[j].coefsT, N, (size_t) j * N);
}

void lweKeySwitch(LweSample *result, const LweKeySwitchKey *ks, const LweSample *sample) {
    tfhe_b200_ctx *c = ctx_for_ks(ks);
    const int n = ks->out_params->n, Nin = ks->n;
    DevBuf d((size_t) Nin + 1 + n + 1);
    d.up(sample->a, Nin, 0);
    d.up(&sample->b, 1, Nin);
    OK(tfhe_b200_keyswitch(c, d.p + Nin + 1, d.p, 1, nullptr), "key switch");
    cu(cudaStreamSynchronize(nullptr), "key switch");
    d.down(result->a, n, (size_t) Nin + 1);
    d.down(&result->b, 1, (size_t) Nin + 1 + n);
    result->current_variance = 0.;
}

// ---- batched family ---------------------------------------------------------------------

void *tfhe_b200_keys_to_gpu(const TFheGateBootstrappingCloudKeySet *bk) { return ctx_for_cloud(bk); }

void tfhe_b200_keys_free(const TFheGateBootstrappingCloudKeySet *bk) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(bk);
    if (it != g_ctx.end()) {
        tfhe_b200_ctx_destroy(it->second);
        g_ctx.erase(it);
    }
}

void bootsAND_fullGPU_n_Bit(LweSample_16 *r, const LweSample_16 *a, const LweSample_16 *b, int nBits, void *h, Torus32 *, Torus32 *) {
    batched_gate2(TFHE_B200_AND, -1, r, a, b, nullptr, nullptr, nBits, h);
}
void bootsXOR_fullGPU_n_Bit(LweSample_16 *r, const LweSample_16 *a, const LweSample_16 *b, int nBits, void *h, Torus32 *, Torus32 *) {
    batched_gate2(TFHE_B200_XOR, -1, r, a, b, nullptr, nullptr, nBits, h);
}
void bootsXNOR_fullGPU_n_Bit(LweSample_16 *r, const LweSample_16 *a, const LweSample_16 *b, int nBits, void *h, Torus32 *, Torus32 *) {
    batched_gate2(TFHE_B200_XNOR, -1, r, a, b, nullptr, nullptr, nBits, h);
}

void bootsMUX_fullGPU_n_Bit(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb,
                            const LweSample_16 *cc, int nBits, void *handle, Torus32 *, Torus32 *) {
    tfhe_b200_ctx *c = (tfhe_b200_ctx *) handle;
    if (!c) die("null key handle (pass tfhe_b200_keys_to_gpu(bk) as bkGPU)");
    const int n = tfhe_b200_ctx_words(c) - 1;
    Rows16 ra(nBits, n), rb(nBits, n), rc(nBits, n), out(nBits, n);
    ra.pack(ca);
    rb.pack(cb);
    rc.pack(cc);
    OK(tfhe_b200_mux(c, out.rows, ra.rows, rb.rows, rc.rows, nBits, nullptr), "mux batch");
    out.unpack(result);
}

void bootsANDXOR_fullGPU_n_Bit_vector(LweSample_16 *result, const LweSample_16 *ca, const LweSample_16 *cb,
                                      int vLength, int nBits, void *h, Torus32 *, Torus32 *) {
    batched_gate2(TFHE_B200_AND, TFHE_B200_XOR, result, ca, cb, ca, cb, vLength * nBits, h);
}

void bootsXORXOR_fullGPU_n_Bit_vector(LweSample_16 *result, const LweSample_16 *ca1, const LweSample_16 *ca2,
                                      const LweSample_16 *cb1, const LweSample_16 *cb2, int vLength, int nBits,
                                      void *h, Torus32 *, Torus32 *) {
    batched_gate2(TFHE_B200_XOR, TFHE_B200_XOR, result, ca1, ca2, cb1, cb2, vLength * nBits, h);
}

// boot-gates.cu:1267-1292: negate a on the device, b on the host
void bootsNOT_16(LweSample_16 *output, LweSample_16 *input, int bitSize, int params_n) {
    Rows16 r(bitSize, params_n);
    r.pack(input);
    const long long total = (long long) bitSize * (params_n + 1);
    negate_kernel<<<(unsigned) ((total + 255) / 256), 256>>>(r.rows, total);
    r.unpack(output);
}

LweSample_16 *convertBitToNumberZero_GPU(int bitSize, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    LweSample_16 *s = (LweSample_16 *) malloc(sizeof(LweSample_16));
    cu(cudaMalloc(&s->a, (size_t) bitSize * n * sizeof(int)), "cudaMalloc");
    cu(cudaMemset(s->a, 0, (size_t) bitSize * n * sizeof(int)), "cudaMemset");
    s->b = (int *) calloc((size_t) bitSize, sizeof(int));
    for (int i = 0; i < bitSize; i++) s->b[i] = -0x20000000;  // boot-gates.cu:469-472
    s->current_variance = (double *) calloc((size_t) bitSize, sizeof(double));
    return s;
}

// convertBitToNumber (boot-gates.cu:513-532): `bitSize` LweSamples -> one HOST LweSample_16 (the
// caller then moves `a` to the device itself, main.cu:911-915); convertNumberToBits (:534-548): back.
LweSample_16 *convertBitToNumber(const LweSample *input, int bitSize, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    LweSample_16 *s = (LweSample_16 *) malloc(sizeof(LweSample_16));
    s->a = (int *) malloc(sizeof(int) * (size_t) bitSize * n);
    s->b = (int *) malloc(sizeof(int) * (size_t) bitSize);
    s->current_variance = (double *) malloc(sizeof(double) * (size_t) bitSize);
    for (int i = 0; i < bitSize; i++) {
        memcpy(s->a + (size_t) i * n, input[i].a, sizeof(int) * (size_t) n);
        s->b[i] = input[i].b;
        s->current_variance[i] = input[i].current_variance;
    }
    return s;
}

LweSample *convertNumberToBits(LweSample_16 *number, int bitSize, const TFheGateBootstrappingCloudKeySet *bk) {
    const int n = bk->params->in_out_params->n;
    LweSample *out = new_gate_bootstrapping_ciphertext_array(bitSize, bk->params);
    for (int i = 0; i < bitSize; i++) {
        memcpy(out[i].a, number->a + (size_t) i * n, sizeof(int) * (size_t) n);  // `a` must be a host pointer here
        out[i].b = number->b[i];
        out[i].current_variance = number->current_variance[i];
    }
    return out;
}

void freeLweSample_16(LweSample_16 *s) {  // boot-gates.cu:550-556 (host container)
    if (!s) return;
    free(s->a);
    free(s->b);
    free(s->current_variance);
    free(s);
}

void freeLweSample_16_gpu(LweSample_16 *s) {  // main.cu:41
    if (!s) return;
    cudaFree(s->a);
    free(s->b);
    free(s->current_variance);
    free(s);
}

}  // extern "C"
