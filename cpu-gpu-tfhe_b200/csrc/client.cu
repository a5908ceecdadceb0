// Client-side operations of the gate API: key generation, bit encryption and
// decryption.  In TFHE these belong to the key owner, not to the evaluation hot
// path, and the reference runs them on the host too
// (new_random_gate_bootstrapping_secret_keyset, tfhe_gate_bootstrapping.cu:57-68;
//  bootsSymEncrypt :114, bootsSymDecrypt :122).  Host C++, multi-threaded for the
// 2000 TLWE encryptions of the bootstrapping key.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "csprng.h"

namespace {

// All randomness is ChaCha20 keystream (csprng.h): the key comes from the operating system unless the
// caller passes a non-zero seed (reproducible TEST material, not secure).
using Rng = tfhe_b200::ChaChaRng;

// dtot32, numeric-functions.cu:33-35
int32_t dtot32(double d) { return (int32_t) (int64_t) ((d - (double) (int64_t) d) * 4294967296.); }

// b += key * a mod X^N+1 (key binary), exact integer arithmetic
void addmul_binary(uint32_t *b, const int32_t *key, const uint32_t *a, int N) {
    for (int s = 0; s < N; s++) {
        if (!key[s]) continue;
        for (int i = 0; i < s; i++) b[i] -= a[i - s + N];
        for (int i = s; i < N; i++) b[i] += a[i - s];
    }
}

}  // namespace

extern "C" {

size_t tfhe_b200_bk_words(const tfhe_b200_params *p) {
    return (size_t) p->n * (p->k + 1) * p->l * (p->k + 1) * p->N;
}

size_t tfhe_b200_ks_words(const tfhe_b200_params *p) {
    return (size_t) p->N * p->k * p->ks_t * ((size_t) 1 << p->ks_basebit) * (p->n + 1);
}

void tfhe_b200_default_noise(double *alpha_lwe, double *alpha_bk) {
    // tfhe_gate_bootstrapping.cu:36-37 : sqrt(2/pi) * 2^-15 and sqrt(2/pi) * 9e-9
    *alpha_lwe = std::pow(2., -15) * std::sqrt(2. / M_PI);
    *alpha_bk = 9.e-9 * std::sqrt(2. / M_PI);
}

// Secret keys + cloud keys in the flat formats of tfhe_b200.h.
int tfhe_b200_keygen(const tfhe_b200_params *p, uint64_t seed, double alpha_lwe, double alpha_bk,
                     int32_t *lwe_key, int32_t *tlwe_key, int32_t *bk, int32_t *ks) {
    if (!p || !lwe_key || !tlwe_key || !bk || !ks) return 1;
    const int n = p->n, N = p->N, k = p->k, l = p->l, kpl = (k + 1) * l;
    const int t = p->ks_t, basebit = p->ks_basebit, base = 1 << basebit;
    tfhe_b200::ChaChaKey ckey;
    if (!tfhe_b200::chacha_key_from_seed(seed, &ckey)) return 1;  // seed 0: getrandom(); else test key
    Rng rng(ckey, 0);
    // lweKeyGen lwe-functions.cu:21-27 ; tLweKeyGen tlwe-functions.cu:15-23
    for (int i = 0; i < n; i++) lwe_key[i] = rng.bit();
    for (int i = 0; i < k * N; i++) tlwe_key[i] = rng.bit();

    // lweCreateKeySwitchKey lwe-keyswitch-functions.cu:890-942 (noise re-centred on its mean)
    {
        const int nin = k * N;
        const size_t sizeks = (size_t) nin * t * (base - 1);
        std::vector<double> noise(sizeks);
        double err = 0;
        for (auto &v : noise) {
            v = rng.gauss(alpha_lwe);
            err += v;
        }
        err /= (double) sizeks;
        for (auto &v : noise) v -= err;
        size_t index = 0;
        for (int i = 0; i < nin; i++)
            for (int j = 0; j < t; j++) {
                int32_t *row0 = ks + (((size_t) i * t + j) * base) * (n + 1);
                memset(row0, 0, sizeof(int32_t) * (size_t) (n + 1));
                for (int h = 1; h < base; h++) {
                    int32_t *row = row0 + (size_t) h * (n + 1);
                    uint32_t b = (uint32_t) (tlwe_key[i] * h) * (1u << (32 - (j + 1) * basebit)) +
                                 (uint32_t) dtot32(noise[index++]);
                    for (int c = 0; c < n; c++) {
                        row[c] = rng.torus();
                        b += (uint32_t) row[c] * (uint32_t) lwe_key[c];
                    }
                    row[n] = (int32_t) b;
                }
            }
    }

    // tfhe_createLweBootstrappingKey lwe-bootstrapping-functions.cu:185-217: BK_i = TGSW(s_i)
    // (tGswSymEncryptInt tgsw-functions.cu:191: kpl TLWE encryptions of 0 + s_i * H)
    {
        const unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = (int) (hw ? (hw > 16 ? 16 : hw) : 4);
        auto work = [&](int tid) {
            for (int i = tid; i < n; i += nthreads) {
                Rng r(ckey, 1 + (uint64_t) i);  // one keystream per TGSW sample
                for (int row_i = 0; row_i < kpl; row_i++) {
                    uint32_t *row = (uint32_t *) bk + ((size_t) i * kpl + row_i) * (k + 1) * N;
                    uint32_t *b = row + (size_t) k * N;
                    for (int j = 0; j < N; j++) b[j] = (uint32_t) dtot32(r.gauss(alpha_bk));
                    for (int m = 0; m < k; m++) {
                        uint32_t *a = row + (size_t) m * N;
                        for (int j = 0; j < N; j++) a[j] = (uint32_t) r.torus();
                        addmul_binary(b, tlwe_key + (size_t) m * N, a, N);
                    }
                }
                for (int bloc = 0; bloc <= k; bloc++)
                    for (int q = 0; q < l; q++) {
                        uint32_t *row = (uint32_t *) bk + ((size_t) i * kpl + bloc * l + q) * (k + 1) * N;
                        row[(size_t) bloc * N] += (uint32_t) lwe_key[i] * (1u << (32 - (q + 1) * p->Bgbit));
                    }
            }
        };
        std::vector<std::thread> th;
        for (int tid = 0; tid < nthreads; tid++) th.emplace_back(work, tid);
        for (auto &x : th) x.join();
    }
    return 0;
}

// bootsSymEncrypt (tfhe_gate_bootstrapping.cu:114-119) for a batch of bits: message +-1/8
int tfhe_b200_encrypt_bits(const tfhe_b200_params *p, const int32_t *lwe_key, uint64_t seed, double alpha,
                           const int32_t *bits, int count, int32_t *out) {
    if (!p || !lwe_key || !out || count < 0) return 1;
    const int n = p->n;
    tfhe_b200::ChaChaKey ckey;
    if (!tfhe_b200::chacha_key_from_seed(seed, &ckey)) return 1;  // seed 0: getrandom(); else test key
    Rng rng(ckey, 0x656e63ull);
    for (int g = 0; g < count; g++) {
        int32_t *s = out + (size_t) g * (n + 1);
        uint32_t b = (uint32_t) ((bits && bits[g]) ? 0x20000000 : -0x20000000) + (uint32_t) dtot32(rng.gauss(alpha));
        for (int i = 0; i < n; i++) {
            s[i] = rng.torus();
            b += (uint32_t) s[i] * (uint32_t) lwe_key[i];
        }
        s[n] = (int32_t) b;
    }
    return 0;
}

// lwePhase (lwe-functions.cu:72-81)
int tfhe_b200_phases(const int32_t *key, int n, const int32_t *samples, int count, int32_t *phases_out) {
    if (!key || !samples || !phases_out) return 1;
    for (int g = 0; g < count; g++) {
        const int32_t *s = samples + (size_t) g * (n + 1);
        uint32_t axs = 0;
        for (int i = 0; i < n; i++) axs += (uint32_t) s[i] * (uint32_t) key[i];
        phases_out[g] = (int32_t) ((uint32_t) s[n] - axs);
    }
    return 0;
}

// bootsSymDecrypt (tfhe_gate_bootstrapping.cu:122-125)
int tfhe_b200_decrypt_bits(const tfhe_b200_params *p, const int32_t *lwe_key, const int32_t *samples, int count,
                           int32_t *bits_out) {
    if (!p) return 1;
    if (tfhe_b200_phases(lwe_key, p->n, samples, count, bits_out)) return 1;
    for (int g = 0; g < count; g++) bits_out[g] = bits_out[g] > 0 ? 1 : 0;
    return 0;
}

}  // extern "C"
