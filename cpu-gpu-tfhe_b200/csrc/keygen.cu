// Cloud-key generation on the GPU.
//
// new_random_gate_bootstrapping_secret_keyset (tfhe_gate_bootstrapping.cu:57-68) spends its time
// in the 2000 TLWE encryptions of the bootstrapping key (tfhe_createLweBootstrappingKey,
// lwe-bootstrapping-functions.cu:185-217 -> tGswSymEncryptInt, tgsw-functions.cu:191: a uniform
// mask polynomial, a negacyclic product with the binary key, Gaussian noise) and the 24576
// samples of the key-switch key (lweCreateKeySwitchKey, lwe-keyswitch-functions.cu:890-942).
// Both are embarrassingly parallel: one CTA per TLWE row / per key-switch sample, counter-based
// random numbers: ChaCha20 keystream blocks (csprng.h), one stream per thread, the 256-bit key from
// the operating system (or from the caller's non-zero TEST seed).  The tiny parts stay on the host: the secret bits and the re-centred key-switch noise
// (the reference subtracts the mean of all its noise terms, :905-912).
// The flat keys are produced directly in device memory and handed to
// tfhe_b200_load_keys_device (Fourier conversion + table re-layouts), so a context can be keyed
// without the 82 MB host round trip; they can also be downloaded (to be saved with keyio.cu).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "csprng.h"

using tfhe_b200::ChaChaKey;
using tfhe_b200::ChaChaStream;

namespace {

// dtot32, numeric-functions.cu:33-35
__device__ __forceinline__ int32_t dtot32_dev(double d) {
    return (int32_t) (long long) ((d - (double) (long long) d) * 4294967296.);
}

// One CTA per TLWE row (i, r) of the bootstrapping key, N = blockDim.x * 4 coefficients.
// bk[(i*kpl + r)][k+1][N]:  a_m uniform, b = sum_m a_m (*) key_m + e, then + s_i * H on the diagonal rows.
__global__ void __launch_bounds__(256) bk_gen_kernel(int32_t *__restrict__ bk, const int32_t *__restrict__ lwe_key,
                                                     const int32_t *__restrict__ tlwe_key, int N, int k, int l,
                                                     int Bgbit, double alpha, const ChaChaKey ckey) {
    extern __shared__ uint32_t sh[];  // a[N] then key bits[N]
    uint32_t *a = sh, *key = sh + N;
    const int kpl = (k + 1) * l;
    const int row = blockIdx.x, i = row / kpl, r = row % kpl;
    uint32_t *dst = reinterpret_cast<uint32_t *>(bk) + (size_t) row * (k + 1) * N;
    ChaChaStream st(ckey, 0x100000000ull + (unsigned long long) row * blockDim.x + threadIdx.x);
    const int per = N / blockDim.x;  // coefficients per thread (4 for N = 1024)
    uint32_t b[8];
    for (int c = 0; c < per; c += 2) {  // Box-Muller: two normals from two 53-bit uniforms
        const uint32_t w0 = st.word(), w1 = st.word(), w2 = st.word(), w3 = st.word();
        const double u1 = tfhe_b200::chacha_unit(w0, w1), u2 = tfhe_b200::chacha_unit(w2, w3);
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        b[c] = (uint32_t) dtot32_dev(r * cs * alpha);
        if (c + 1 < per) b[c + 1] = (uint32_t) dtot32_dev(r * sn * alpha);
    }
    for (int m = 0; m < k; m++) {
        __syncthreads();
        for (int c = 0; c < per; c++) {
            const int j = threadIdx.x + c * blockDim.x;
            const uint32_t v = st.word();
            a[j] = v;
            key[j] = (uint32_t) tlwe_key[(size_t) m * N + j];
            dst[(size_t) m * N + j] = v;
        }
        __syncthreads();
        // b += key (*) a mod X^N + 1, exact integer arithmetic
        for (int s = 0; s < N; s++) {
            if (!key[s]) continue;  // uniform branch: the whole CTA sees the same key bit
            for (int c = 0; c < per; c++) {
                const int j = threadIdx.x + c * blockDim.x;
                b[c] += (j >= s) ? a[j - s] : 0u - a[j - s + N];
            }
        }
    }
    for (int c = 0; c < per; c++) dst[(size_t) k * N + threadIdx.x + c * blockDim.x] = b[c];
    __syncthreads();
    // + s_i * h_q on coefficient 0 of polynomial `bloc` for the row (bloc, q)   (tGswAddMuIntH, tgsw-functions.cu:129)
    if (threadIdx.x == 0) {
        const int bloc = r / l, q = r % l;
        dst[(size_t) bloc * N] += (uint32_t) lwe_key[i] * (1u << (32 - (q + 1) * Bgbit));
    }
}

// One CTA per key-switch sample (i, j, h): ks[((i*t + j)*base + h)][n+1]; h = 0 is the noiseless zero.
__global__ void __launch_bounds__(128) ks_gen_kernel(int32_t *__restrict__ ks, const int32_t *__restrict__ lwe_key,
                                                     const int32_t *__restrict__ tlwe_key,
                                                     const int32_t *__restrict__ noise, int n, int t, int basebit,
                                                     const ChaChaKey ckey) {
    __shared__ uint32_t red[128];
    const int base = 1 << basebit;
    const int sample = blockIdx.x, h = sample % base, ij = sample / base, j = ij % t, i = ij / t;
    int32_t *row = ks + (size_t) sample * (n + 1);
    if (h == 0) {
        for (int c = threadIdx.x; c <= n; c += blockDim.x) row[c] = 0;
        return;
    }
    ChaChaStream st(ckey, 0x200000000ull + (unsigned long long) sample * blockDim.x + threadIdx.x);
    uint32_t dot = 0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const uint32_t v = st.word();
        row[c] = (int32_t) v;
        dot += v * (uint32_t) lwe_key[c];
    }
    red[threadIdx.x] = dot;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if ((int) threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t msg = (uint32_t) (tlwe_key[i] * h) * (1u << (32 - (j + 1) * basebit));
        row[n] = (int32_t) (red[0] + msg + (uint32_t) noise[((size_t) ij) * (base - 1) + (h - 1)]);
    }
}

int kg_fail(const char *m) {
    fprintf(stderr, "tfhe_b200 keygen: %s\n", m);
    return 1;
}

#define KG(x)                                              \
    do {                                                   \
        cudaError_t e_ = (x);                              \
        if (e_ != cudaSuccess) {                           \
            cudaFree(d_bk); cudaFree(d_ks); cudaFree(d_lwe); cudaFree(d_tlwe); cudaFree(d_noise); \
            return kg_fail(cudaGetErrorString(e_));        \
        }                                                  \
    } while (0)

}  // namespace

extern "C" {

// Generates a fresh key set: the secret bits on the host (returned in lwe_key[n], tlwe_key[k*N]),
// the cloud keys on the GPU of `ctx`, loaded into it.  bk_out / ks_out (host, flat formats,
// tfhe_b200_bk_words / tfhe_b200_ks_words) may be NULL.
int tfhe_b200_keygen_device(tfhe_b200_ctx *ctx, const tfhe_b200_params *p, uint64_t seed, double alpha_lwe,
                            double alpha_bk, int32_t *lwe_key, int32_t *tlwe_key, int32_t *bk_out, int32_t *ks_out) {
    if (!ctx || !p || !lwe_key || !tlwe_key) return kg_fail("null argument");
    const int n = p->n, N = p->N, k = p->k, l = p->l, kpl = (k + 1) * l, t = p->ks_t, base = 1 << p->ks_basebit;
    if (N % 256 != 0 || N / 256 > 8) return kg_fail("unsupported ring degree");
    if (cudaSetDevice(tfhe_b200_ctx_device(ctx)) != cudaSuccess) return kg_fail("no CUDA device");
    ChaChaKey ckey;
    if (!tfhe_b200::chacha_key_from_seed(seed, &ckey)) return kg_fail("getrandom failed");  // seed 0: OS entropy
    tfhe_b200::ChaChaRng gen(ckey, 0);
    for (int i = 0; i < n; i++) lwe_key[i] = (int32_t) gen.bit();
    for (int i = 0; i < k * N; i++) tlwe_key[i] = (int32_t) gen.bit();
    // key-switch noise, re-centred on its mean (lwe-keyswitch-functions.cu:905-912)
    const size_t nnoise = (size_t) N * k * t * (base - 1);
    std::vector<double> noise(nnoise);
    double mean = 0;
    for (auto &v : noise) {
        v = gen.gauss(alpha_lwe);
        mean += v;
    }
    mean /= (double) nnoise;
    std::vector<int32_t> noise32(nnoise);
    for (size_t i = 0; i < nnoise; i++) {
        const double d = noise[i] - mean;
        noise32[i] = (int32_t) (int64_t) ((d - (double) (int64_t) d) * 4294967296.);
    }

    int32_t *d_bk = nullptr, *d_ks = nullptr, *d_lwe = nullptr, *d_tlwe = nullptr, *d_noise = nullptr;
    const size_t bkb = tfhe_b200_bk_words(p) * sizeof(int32_t), ksb = tfhe_b200_ks_words(p) * sizeof(int32_t);
    KG(cudaMalloc(&d_bk, bkb));
    KG(cudaMalloc(&d_ks, ksb));
    KG(cudaMalloc(&d_lwe, sizeof(int32_t) * n));
    KG(cudaMalloc(&d_tlwe, sizeof(int32_t) * k * N));
    KG(cudaMalloc(&d_noise, sizeof(int32_t) * nnoise));
    KG(cudaMemcpy(d_lwe, lwe_key, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    KG(cudaMemcpy(d_tlwe, tlwe_key, sizeof(int32_t) * k * N, cudaMemcpyHostToDevice));
    KG(cudaMemcpy(d_noise, noise32.data(), sizeof(int32_t) * nnoise, cudaMemcpyHostToDevice));
    bk_gen_kernel<<<n * kpl, 256, 2 * N * sizeof(uint32_t)>>>(d_bk, d_lwe, d_tlwe, N, k, l, p->Bgbit, alpha_bk, ckey);
    KG(cudaGetLastError());
    ks_gen_kernel<<<N * k * t * base, 128>>>(d_ks, d_lwe, d_tlwe, d_noise, n, t, p->ks_basebit, ckey);
    KG(cudaGetLastError());
    if (tfhe_b200_load_keys_device(ctx, d_bk, d_ks, nullptr)) {
        cudaFree(d_bk); cudaFree(d_ks); cudaFree(d_lwe); cudaFree(d_tlwe); cudaFree(d_noise);
        return 1;
    }
    KG(cudaDeviceSynchronize());
    if (bk_out) KG(cudaMemcpy(bk_out, d_bk, bkb, cudaMemcpyDeviceToHost));
    if (ks_out) KG(cudaMemcpy(ks_out, d_ks, ksb, cudaMemcpyDeviceToHost));
    cudaFree(d_bk);
    cudaFree(d_ks);
    cudaFree(d_lwe);
    cudaFree(d_tlwe);
    cudaFree(d_noise);
    return 0;
}

}  // extern "C"
