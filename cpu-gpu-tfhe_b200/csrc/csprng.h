// ChaCha20 keystream generator for everything the key owner draws at random: secret key bits, the
// uniform masks of every LWE / TLWE sample, and the Gaussian noise (host: client.cu, keygen.cu; device:
// the key-generation kernels of keygen.cu).
//
// Why not a general-purpose PRNG: the mask words a[i] of every ciphertext and key row are PUBLISHED.
// With mt19937 (the first version; the reference uses std::default_random_engine,
// numeric-functions.cu:11-13) the generator state is linearly recoverable from a few hundred outputs,
// after which every later noise term is predictable and the LWE secret falls to linear algebra.
// ChaCha20 output does not reveal its key.
//
// Key: 256 bits from the operating system (getrandom) — or, ONLY when the caller passes a non-zero
// seed, derived from that 64-bit seed: reproducible key material for tests and benchmarks, NOT
// secure (the seed is known to the caller and has 64 bits).  Independent streams (threads, key rows,
// samples) differ in the 64-bit stream id that goes into the nonce words.
#pragma once

#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace tfhe_b200 {

struct ChaChaKey {
    uint32_t w[8];
};

__host__ __device__ inline uint32_t chacha_rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

#define TFHE_B200_CHACHA_QR(a, b, c, d)                \
    a += b; d ^= a; d = chacha_rotl(d, 16);            \
    c += d; b ^= c; b = chacha_rotl(b, 12);            \
    a += b; d ^= a; d = chacha_rotl(d, 8);             \
    c += d; b ^= c; b = chacha_rotl(b, 7);

// One 64-byte block of the ChaCha20 keystream (RFC 8439 block function with a 64-bit block counter and
// a 64-bit stream id in the nonce words).
__host__ __device__ inline void chacha20_block(const ChaChaKey &key, uint64_t counter, uint64_t stream, uint32_t (&out)[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key.w[0], key.w[1], key.w[2], key.w[3], key.w[4], key.w[5], key.w[6], key.w[7],
                      (uint32_t) counter, (uint32_t) (counter >> 32), (uint32_t) stream, (uint32_t) (stream >> 32)};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        TFHE_B200_CHACHA_QR(x[0], x[4], x[8], x[12])
        TFHE_B200_CHACHA_QR(x[1], x[5], x[9], x[13])
        TFHE_B200_CHACHA_QR(x[2], x[6], x[10], x[14])
        TFHE_B200_CHACHA_QR(x[3], x[7], x[11], x[15])
        TFHE_B200_CHACHA_QR(x[0], x[5], x[10], x[15])
        TFHE_B200_CHACHA_QR(x[1], x[6], x[11], x[12])
        TFHE_B200_CHACHA_QR(x[2], x[7], x[8], x[13])
        TFHE_B200_CHACHA_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// uniform double in (0, 1] from two keystream words (53 bits)
__host__ __device__ inline double chacha_unit(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t) hi << 32) | lo) >> 11;
    return ((double) v + 1.0) * (1.0 / 9007199254740992.0);
}

// A sequential reader of one stream.
struct ChaChaStream {
    ChaChaKey key;
    uint64_t stream, counter;
    uint32_t buf[16];
    int pos;
    __host__ __device__ ChaChaStream(const ChaChaKey &k, uint64_t stream_id) : key(k), stream(stream_id), counter(0), pos(16) {}
    __host__ __device__ uint32_t word() {
        if (pos == 16) {
            chacha20_block(key, counter++, stream, buf);
            pos = 0;
        }
        return buf[pos++];
    }
    __host__ __device__ int bit() { return (int) (word() >> 31); }
    __host__ __device__ uint64_t u64() {
        const uint64_t hi = word();
        return (hi << 32) | word();
    }
};

}  // namespace tfhe_b200

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>

namespace tfhe_b200 {

// seed == 0: the operating system's CSPRNG; seed != 0: deterministic TEST key (see the header comment)
inline bool chacha_key_from_seed(uint64_t seed, ChaChaKey *key) {
    if (seed == 0) {
        size_t got = 0;
        while (got < sizeof(key->w)) {
            const ssize_t r = getrandom((char *) key->w + got, sizeof(key->w) - got, 0);
            if (r <= 0) return false;
            got += (size_t) r;
        }
        return true;
    }
    // expand the 64-bit seed through the block function itself
    ChaChaKey k0 = {{0x74666865u, 0x62323030u, 0x74657374u, 0x6b657921u, (uint32_t) seed, (uint32_t) (seed >> 32),
                     ~(uint32_t) seed, ~(uint32_t) (seed >> 32)}};
    uint32_t out[16];
    chacha20_block(k0, 0, 0x5eed5eed5eed5eedull, out);
    memcpy(key->w, out, sizeof(key->w));
    return true;
}

// host-side generator with Gaussian noise (Box-Muller on two 53-bit uniforms)
struct ChaChaRng : ChaChaStream {
    bool has_spare = false;
    double spare = 0.;
    ChaChaRng(const ChaChaKey &k, uint64_t stream_id) : ChaChaStream(k, stream_id) {}
    int32_t torus() { return (int32_t) word(); }
    double gauss(double sigma) {
        if (has_spare) {
            has_spare = false;
            return spare * sigma;
        }
        const uint32_t a = word(), b = word(), c = word(), d = word();
        const double u1 = chacha_unit(a, b), u2 = chacha_unit(c, d);
        const double r = sqrt(-2.0 * log(u1)), th = 6.283185307179586476925286766559 * u2;
        spare = r * sin(th);
        has_spare = true;
        return r * cos(th) * sigma;
    }
};

}  // namespace tfhe_b200
