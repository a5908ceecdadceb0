// Internal host-side declarations of the kernel launchers (engine.cu calls
// these; nothing here is part of the public C ABI in include/tfhe_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

struct cpx;

// One homogeneous run of bootstraps inside a batch: x = (0,cst) + sa*in0 + sb*in1 [+ sc*in2 [+ sd*in3]]
// (gate prologues, boot-gates.cu:98-448).  Samples are int32[n+1] rows.
struct BrSegment {
    const int32_t *in0;
    const int32_t *in1;
    long long stride0;  // in words
    long long stride1;
    // optional gather: operand row of local gate g is in + idx[g] * stride (instead of g * stride)
    const int32_t *idx0;
    const int32_t *idx1;
    // optional third operand (three-input threshold gates, e.g. the carry operator)
    const int32_t *in2;
    long long stride2;
    const int32_t *idx2;
    // optional fourth operand (the fused sum bit of the prefix adder, TFHE_B200_SUMC)
    const int32_t *in3;
    long long stride3;
    const int32_t *idx3;
    int sd;
    int sc;
    int sa, sb;
    int32_t cst;
    int count;
};

constexpr int kMaxSegments = 16;  // runs of one bootstrap batch (tfhe_b200_gate_multi, merged circuit levels)

struct BrLaunch {
    BrSegment seg[kMaxSegments];
    int nseg;
    int total;            // bootstraps in this launch
    int cts_per_group;    // set by launch_blind_rotate: ciphertexts per CTA pass (1..4)
    int n;                // LWE dimension = blind-rotation iterations available in bk
    int n_iter;           // iterations to run (<= n)
    int extern_only;      // 1: a single external product ACC <- BK_{bk_first} (.) ACC, no rotation
    int bk_first;         // first key element used (iteration it uses BK_{bk_first + it})
    int32_t mu;           // test-vector message
    const cpx *bk;        // device key: [n][4][2][16][32] complex
    // explicit inputs (API tfhe_blindRotate[AndExtract]_FFT): if explicit_inputs != 0 the
    // segments are ignored; bara is [total][n_iter], barb [total] (nullptr = 0)
    int explicit_inputs;
    const int32_t *bara;
    const int32_t *barb;
    const int32_t *acc_in;   // optional [total][2][1024]; overrides the test-vector init
    const int32_t *testvect; // optional [1024] shared test vector (instead of constant mu)
    int32_t *u_out;          // optional [total][1025] extracted samples
    int32_t *acc_out;        // optional [total][2][1024] raw accumulators
};

// persistent blind-rotate + extract kernel
cudaError_t launch_blind_rotate(const BrLaunch &L, int sm_count, cudaStream_t stream);
size_t blind_rotate_smem_bytes();
cudaError_t blind_rotate_configure();

// key conversion: int32 coefficient polynomials -> device Fourier layout
// (npolys must be a multiple of 4); out = transform(coef * scale)
cudaError_t launch_forward_polys(const int32_t *coef, cpx *out, int npolys, double scale, cudaStream_t stream);

// key switch (lweKeySwitch, lwe-keyswitch-functions.cu:955-987)
struct KsLaunch {
    const int32_t *ks;     // device table [N][t][base-1][512]
    const int32_t *u;      // [*][N+1]
    int nsrc;              // 1: u[g]; 2: u[g] + u[g + count]  (MUX)
    int32_t cst;           // added to b
    // outputs: up to kMaxSegments runs of rows (out + i*stride), in gate order
    struct Out {
        int32_t *out;
        long long stride;  // words
        int count;
        const int32_t *idx;  // optional scatter: row of local gate g is out + idx[g] * stride
    } dst[kMaxSegments];
    int ndst;
    int count;
    int n;                 // 500
    int N;                 // 1024
    int t;                 // 8
    int basebit;           // 2
};
cudaError_t launch_keyswitch(const KsLaunch &L, int sm_count, cudaStream_t stream);
constexpr int kKsRowWords = 512;
constexpr int kKsTile = 32;  // gates per key-switch CTA

// KS table re-layout on device: src [N][t][base][n+1] -> dst [N][t][base-1][512]
cudaError_t launch_ks_relayout(const int32_t *src, int32_t *dst, int N, int t, int base, int n, cudaStream_t stream);

// small linear ops on sample batches
cudaError_t launch_lwe_linear(int32_t *out, long long out_stride, const int32_t *in0, long long s0, int c0,
                              const int32_t *in1, long long s1, int c1, int32_t cst, int count, int n,
                              cudaStream_t stream);

// tensor-core key switch (keyswitch_mma.cu): same KsLaunch, table pre-tiled as unsigned byte limbs
size_t ks_mma_table_bytes();
bool ks_mma_supported(int N, int t, int basebit, int n);
cudaError_t launch_ks_mma_relayout(const int32_t *src, uint8_t *dst, int base, int n, cudaStream_t stream);
cudaError_t launch_keyswitch_mma(const KsLaunch &L, const uint8_t *tbl, int sm_count, cudaStream_t stream);
cudaError_t launch_ks_zero(const KsLaunch &L, cudaStream_t stream);

// Scratch for the extracted samples of the calling thread's NEXT bootstrap launches (engine.cu
// run_bootstrap_ks): circuit plans own their scratch so that a captured CUDA graph holds no
// allocation nodes.  nullptr restores the stream-ordered allocator.
void engine_set_thread_scratch(int32_t *d_u, size_t bytes);

cudaError_t launch_lwe_linear_idx(int32_t *out, const int32_t *in, long long stride, const int32_t *idx_out,
                                  const int32_t *idx_in, int c0, int32_t cst, int count, int n, cudaStream_t stream);

}  // namespace tfhe_b200
