/*
 * examples/cloud_add.c — the "cloud" side of the reference's client / cloud hand-off
 * (cpuParallel/main.cpp writes secret.key, cloud.key and cloud.data; cpuParallel/cloud.cpp:138-161
 * reads cloud.key and cloud.data, computes, writes answer.data) as a plain C program over the flat
 * C ABI of include/tfhe_b200.h: no C++, no CUDA headers, no Python.
 *
 *   cloud_add <cloud.key> <cloud.data> <answer.data> [nbits]
 *
 * cloud.data holds two nbits-bit integers (LSB first, 2*nbits ciphertext records); answer.data
 * receives their sum mod 2^nbits (parallel-prefix adder) followed by their product mod 2^nbits
 * (carry-save multiplier).  Everything between reading the files and writing the answer runs on the GPU.
 *
 * Build:  gcc -std=c99 -O2 -I include examples/cloud_add.c -o cloud_add \
 *             -L cpu-gpu-tfhe_b200 -ltfhe_b200 -Wl,-rpath,$PWD/cpu-gpu-tfhe_b200
 */
#include <stdio.h>
#include <stdlib.h>

#include "tfhe_b200.h"

#define CHECK(call, what)                                                        \
    do {                                                                         \
        if (call) {                                                              \
            fprintf(stderr, "%s failed: %s\n", what, tfhe_b200_last_error());    \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s cloud.key cloud.data answer.data [nbits]\n", argv[0]);
        return 2;
    }
    const int nbits = argc > 4 ? atoi(argv[4]) : 16;
    tfhe_b200_params p;
    double alphas[4], variances[2];
    if (tfhe_b200_file_read_cloud_key(argv[1], &p, alphas, NULL, NULL, NULL)) {  /* header only: the sizes */
        fprintf(stderr, "cannot read %s: %s\n", argv[1], tfhe_b200_file_last_error());
        return 1;
    }
    int32_t *bk = (int32_t *) malloc(tfhe_b200_bk_words(&p) * sizeof(int32_t));
    int32_t *ks = (int32_t *) malloc(tfhe_b200_ks_words(&p) * sizeof(int32_t));
    if (!bk || !ks || tfhe_b200_file_read_cloud_key(argv[1], &p, alphas, variances, bk, ks)) {
        fprintf(stderr, "cannot read %s: %s\n", argv[1], tfhe_b200_file_last_error());
        return 1;
    }
    const int words = p.n + 1;
    const size_t row_bytes = (size_t) words * sizeof(int32_t);
    int32_t *in = (int32_t *) malloc(2 * (size_t) nbits * row_bytes);
    if (tfhe_b200_file_count_ciphertexts(argv[2], p.n) < 2 * nbits ||
        tfhe_b200_file_read_ciphertexts(argv[2], p.n, in, NULL, 2 * nbits)) {
        fprintf(stderr, "cannot read %d ciphertexts from %s: %s\n", 2 * nbits, argv[2], tfhe_b200_file_last_error());
        return 1;
    }

    tfhe_b200_ctx *ctx = NULL;
    CHECK(tfhe_b200_ctx_create(&ctx, &p, 0), "context");
    CHECK(tfhe_b200_load_keys(ctx, bk, ks), "key upload");
    free(bk);
    free(ks);

    int32_t *d_in = NULL, *d_out = NULL;
    CHECK(tfhe_b200_device_alloc(ctx, (void **) &d_in, 2 * nbits * row_bytes), "device memory");
    CHECK(tfhe_b200_device_alloc(ctx, (void **) &d_out, 2 * nbits * row_bytes), "device memory");
    CHECK(tfhe_b200_copy_to_device(ctx, d_in, in, 2 * nbits * row_bytes, NULL), "upload");
    const int32_t *operands[2] = {d_in, d_in + (size_t) nbits * words};
    tfhe_b200_circuit *add = tfhe_b200_circuit_add(ctx, nbits, 1, 2 /* parallel prefix */);
    tfhe_b200_circuit *mul = tfhe_b200_circuit_mul_ex(ctx, nbits, 1, TFHE_B200_ADDER_CARRY_SAVE);
    if (!add || !mul) {
        fprintf(stderr, "cannot build the circuit plans\n");
        return 1;
    }
    CHECK(tfhe_b200_circuit_run(add, d_out, operands, NULL), "addition");
    CHECK(tfhe_b200_circuit_run(mul, d_out + (size_t) nbits * words, operands, NULL), "multiplication");
    int32_t *out = (int32_t *) malloc(2 * (size_t) nbits * row_bytes);
    CHECK(tfhe_b200_copy_to_host(ctx, out, d_out, 2 * nbits * row_bytes, NULL), "download");  /* waits for the plans */
    if (tfhe_b200_file_write_ciphertexts(argv[3], p.n, out, NULL, 2 * nbits, 0)) {
        fprintf(stderr, "cannot write %s: %s\n", argv[3], tfhe_b200_file_last_error());
        return 1;
    }
    printf("cloud_add: %d-bit sum (%d levels, %lld gates) and product (%d levels, %lld gates) written to %s; %llu kernel launches\n",
           nbits, tfhe_b200_circuit_levels(add), tfhe_b200_circuit_gates(add), tfhe_b200_circuit_levels(mul),
           tfhe_b200_circuit_gates(mul), argv[3], tfhe_b200_launch_count(ctx));
    tfhe_b200_circuit_destroy(add);
    tfhe_b200_circuit_destroy(mul);
    tfhe_b200_device_free(ctx, d_in);
    tfhe_b200_device_free(ctx, d_out);
    tfhe_b200_ctx_destroy(ctx);
    free(in);
    free(out);
    return 0;
}
