// Host side of the C ABI declared in include/tfhe_b200.h: context, key upload /
// conversion, and the gate-level launch sequences (blind-rotate kernel followed
// by the key-switch kernel).  No CPU arithmetic on the data path and no
// fallback: every entry point fails loudly if CUDA is unavailable.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "br_core.cuh"
#include "kernels.h"

using namespace tfhe_b200;

struct tfhe_b200_ctx {
    tfhe_b200_params p;
    int device;
    int sm_count;
    cpx *d_bk;        // [n][4][2][16][32] complex
    int32_t *d_ks;    // [N][t][base-1][512]
    uint8_t *d_ks_mma;  // byte-limb tiles for the tensor-core key switch (large batches), or null
    size_t bk_bytes, ks_bytes;
    cudaStream_t stream;  // used by the host-buffer entry points
    cudaStream_t copy_in, copy_out;  // host-buffer entry points: copies of chunk i+1 / i-1 overlap chunk i
    std::atomic<unsigned long long> launches;
    // optional per-kernel timing (bench roofline): event triples around blind-rotate / key-switch
    bool timing;
    std::vector<cudaEvent_t> ev;
};

namespace {

thread_local char g_err[512] = "";
thread_local int32_t *tl_scratch_u = nullptr;   // engine_set_thread_scratch
thread_local size_t tl_scratch_bytes = 0;

int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                    \
    } while (0)

struct GateDef {
    int32_t cst;
    int sa, sb;
    int sc = 0;  // weight of the optional third operand
    int sd = 0;  // ... and of the fourth
};

// modSwitchToTorus32(+-1, 8) = +-0x20000000, (+-1, 4) = +-0x40000000 (numeric-functions.cu:72-77)
constexpr int32_t kMu = 0x20000000;
const GateDef kGates[TFHE_B200_NUM_GATES_EXT] = {
    /* NAND  boot-gates.cu:106-109 */ {kMu, -1, -1},
    /* OR    :132-135 */ {kMu, 1, 1},
    /* AND   :158-162 */ {-kMu, 1, 1},
    /* XOR   :198-201 */ {2 * kMu, 2, 2},
    /* XNOR  :224-227 */ {-2 * kMu, -2, -2},
    /* NOR   :283-286 */ {-kMu, -1, -1},
    /* ANDNY :309-312 */ {-kMu, -1, 1},
    /* ANDYN :335-338 */ {-kMu, 1, -1},
    /* ORNY  :361-364 */ {kMu, -1, 1},
    /* ORYN  :387-390 */ {kMu, 1, -1},
    // Carry operator g | (p & c) for MUTUALLY EXCLUSIVE g, p (generate / propagate signals of an
    // adder): with s = 2g + p + c in {0..3} the phase (s-2)/4 + 1/8 is positive iff s >= 2.
    // One bootstrap instead of AND followed by OR; not in the reference (extension used by the
    // parallel-prefix adder).  Noise weight 6 sigma^2 < XOR's 8 sigma^2.
    /* GPC   */ {kMu, 2, 1, 1},
    // Full-adder outputs (carry-save arithmetic): a + b + c in {+-1/8, +-3/8} has the sign of the
    // majority; -2 (a + b + c) = +-1/4 mod 1, positive for an odd number of ones.
    /* XOR3  */ {0, -2, -2, -2},
    /* MAJ   */ {0, 1, 1, 1},
    // Sum bit of a parallel-prefix adder fused with the last carry operator: p ^ (g | (pp & c)), g and pp
    // mutually exclusive.  v = 2p + 2g + pp + c in {0..5}; the sum is 1 exactly for v in {2, 3}; the phase
    // (v - 1.5) / 4 mod 1 puts those two values at +1/8, +3/8 and the other four at -3/8, -1/8: the
    // standard 1/8 margin, noise weight 10 sigma^2.  Operands a = p, b = g, c = pp, d = c.
    /* SUMC  */ {3 * kMu, 2, 2, 1, 1},
};

BrLaunch base_launch(const tfhe_b200_ctx *c) {
    BrLaunch L;
    memset(&L, 0, sizeof(L));
    L.n = c->p.n;
    L.n_iter = c->p.n;
    L.mu = kMu;
    L.bk = c->d_bk;
    return L;
}

void set_segment(BrSegment &s, const GateDef &g, const int32_t *a, const int32_t *b, int n, int count) {
    s.in0 = a;
    s.in1 = b;
    s.stride0 = n + 1;
    s.stride1 = n + 1;
    s.sa = g.sa;
    s.sb = g.sb;
    s.cst = g.cst;
    s.count = count;
}

int check_ctx(const tfhe_b200_ctx *c, bool need_bk, bool need_ks) {
    if (!c) return fail("null context");
    if (need_bk && !c->d_bk) return fail("bootstrapping key not loaded");
    if (need_ks && !c->d_ks) return fail("key-switch key not loaded");
    return 0;
}

// The contraction runs on the tensor cores (128-gate tiles; small batches split it over the chip, see
// keyswitch_mma.cu).  Measured in round 2 (tools/ks_bench.py), tensor-core / SIMT kernel: 1 gate 0.034 / 0.093 ms,
// 148 gates 0.041 / 0.208, 1024 gates 0.091 / 0.950, 4736 gates 0.433 / 3.93 — so the SIMT kernel only serves
// parameter sets the byte-limb table does not cover (ks_mma_supported) and TFHE_B200_KS_MMA_MIN (tests).
std::atomic<int> g_ks_mma_min{-1};
int ks_mma_min() {
    int v = g_ks_mma_min.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("TFHE_B200_KS_MMA_MIN");
        v = e ? atoi(e) : 1;
        g_ks_mma_min.store(v, std::memory_order_relaxed);
    }
    return v;
}

int run_keyswitch(tfhe_b200_ctx *c, const KsLaunch &K, cudaStream_t st) {
    const int mma_min = ks_mma_min();
    if (c->d_ks_mma && mma_min > 0 && K.count >= mma_min) {
        CU(launch_keyswitch_mma(K, c->d_ks_mma, c->sm_count, st));
        c->launches += (((K.count + 127) / 128) * 4 * 2 <= c->sm_count) ? 2 : 1;  // + the zeroing launch when split
    } else {
        CU(launch_keyswitch(K, c->sm_count, st));
        c->launches += (K.count > 0 && ((K.count + kKsTile - 1) / kKsTile) < 2 * c->sm_count) ? 2 : 1;
    }
    return 0;
}

// blind rotate (segments) -> u scratch -> key switch -> out
int run_bootstrap_ks(tfhe_b200_ctx *c, BrLaunch &L, int nsrc, int32_t ks_cst, int32_t *d_out, int out_count,
                     cudaStream_t st, const KsLaunch::Out *dst = nullptr, int ndst = 0) {
    CU(cudaSetDevice(c->device));
    int32_t *d_u = nullptr;
    const size_t ubytes = (size_t) L.total * (kN + 1) * sizeof(int32_t);
    const bool own_scratch = !(tl_scratch_u != nullptr && tl_scratch_bytes >= ubytes);
    if (own_scratch) CU(cudaMallocAsync(&d_u, ubytes, st));
    else d_u = tl_scratch_u;
    L.u_out = d_u;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != nullptr) CU(cudaStreamIsCapturing(st, &cap));
    const bool timing = c->timing && cap == cudaStreamCaptureStatusNone;  // no timing events inside a graph
    if (timing) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventCreate(&e2));
        CU(cudaEventRecord(e0, st));
    }
    CU(launch_blind_rotate(L, c->sm_count, st));
    c->launches += 1;
    if (timing) CU(cudaEventRecord(e1, st));
    KsLaunch K;
    memset(&K, 0, sizeof(K));
    K.ks = c->d_ks;
    K.u = d_u;
    K.nsrc = nsrc;
    K.cst = ks_cst;
    if (ndst > 0) {
        for (int i = 0; i < ndst; i++) K.dst[i] = dst[i];
        K.ndst = ndst;
    } else {
        K.dst[0].out = d_out;
        K.dst[0].stride = c->p.n + 1;
        K.dst[0].count = out_count;
        K.dst[0].idx = nullptr;
        K.ndst = 1;
    }
    K.count = out_count;
    K.n = c->p.n;
    K.N = c->p.N * c->p.k;
    K.t = c->p.ks_t;
    K.basebit = c->p.ks_basebit;
    if (run_keyswitch(c, K, st)) return 1;
    if (timing) {
        CU(cudaEventRecord(e2, st));
        c->ev.push_back(e0);
        c->ev.push_back(e1);
        c->ev.push_back(e2);
    }
    if (own_scratch) CU(cudaFreeAsync(d_u, st));
    return 0;
}

}  // namespace

namespace tfhe_b200 {
void engine_set_thread_scratch(int32_t *d_u, size_t bytes) {
    tl_scratch_u = d_u;
    tl_scratch_bytes = d_u ? bytes : 0;
}
}  // namespace tfhe_b200

extern "C" {

const char *tfhe_b200_last_error(void) { return g_err; }

// test hook (not in include/): batches below `min_count` take the SIMT key switch; returns the previous value
int tfhe_b200_debug_set_ks_mma_min(int min_count) {
    const int prev = ks_mma_min();
    g_ks_mma_min.store(min_count < 0 ? 0 : min_count, std::memory_order_relaxed);
    return prev;
}

void tfhe_b200_default_params(tfhe_b200_params *p) {
    p->n = 500;
    p->N = 1024;
    p->k = 1;
    p->l = 2;
    p->Bgbit = 10;
    p->ks_t = 8;
    p->ks_basebit = 2;
}

int tfhe_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int tfhe_b200_ctx_create(tfhe_b200_ctx **out, const tfhe_b200_params *p, int device) {
    if (!out || !p) return fail("null argument");
    *out = nullptr;
    if (p->N != kN || p->k != kK || p->l != kL || p->Bgbit != kBgbit)
        return fail("unsupported TGSW parameters (N=%d k=%d l=%d Bgbit=%d): this build instantiates "
                    "N=1024 k=1 l=2 Bgbit=10", p->N, p->k, p->l, p->Bgbit);
    // ks_t <= 8: the SIMT key switch holds N * t digit words per tile in shared memory (64 KiB opt-in at
    // t = 8) and the tensor-core key switch is built for t = 8
    if (p->ks_basebit != 2 || p->ks_t < 1 || p->ks_t > 8 || p->n < 1 || p->n > 511)
        return fail("unsupported LWE / key-switch parameters (n=%d t=%d basebit=%d): this build supports "
                    "basebit=2, 1<=t<=8, 1<=n<=511", p->n, p->ks_t, p->ks_basebit);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device available (%s): the engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    // the library contains sm_100a code only (tcgen05 / TMA paths are not forward compatible)
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    if ((size_t) prop.sharedMemPerBlockOptin < blind_rotate_smem_bytes())
        return fail("device offers %zu B of shared memory per block, kernel needs %zu",
                    (size_t) prop.sharedMemPerBlockOptin, blind_rotate_smem_bytes());
    CU(blind_rotate_configure());
    {
        // scratch comes from the stream-ordered allocator: keep freed blocks cached instead of
        // returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long threshold = ~0ull;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    }
    tfhe_b200_ctx *c = new (std::nothrow) tfhe_b200_ctx();
    if (!c) return fail("out of host memory");
    c->p = *p;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->d_bk = nullptr;
    c->d_ks = nullptr;
    c->d_ks_mma = nullptr;
    c->bk_bytes = c->ks_bytes = 0;
    c->launches = 0;
    c->timing = false;
    c->stream = c->copy_in = c->copy_out = nullptr;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->copy_in) cudaStreamDestroy(c->copy_in);
        delete c;
        return fail("cudaStreamCreate failed");
    }
    *out = c;
    return 0;
}

void tfhe_b200_ctx_destroy(tfhe_b200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->d_bk) cudaFree(c->d_bk);
    if (c->d_ks) cudaFree(c->d_ks);
    if (c->d_ks_mma) cudaFree(c->d_ks_mma);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_in);
    cudaStreamDestroy(c->copy_out);
    delete c;
}

// Per-kernel device time of the gate calls issued since timing was enabled: sums of the
// blind-rotate and key-switch kernel durations (CUDA events on the launching stream).
int tfhe_b200_set_timing(tfhe_b200_ctx *c, int enable) {
    if (!c) return fail("null context");
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    c->ev.clear();
    c->timing = enable != 0;
    return 0;
}

int tfhe_b200_get_timing(tfhe_b200_ctx *c, double *blind_rotate_ms, double *keyswitch_ms, int *calls) {
    if (!c) return fail("null context");
    double br = 0, ks = 0;
    const int n = (int) c->ev.size() / 3;
    for (int i = 0; i < n; i++) {
        CU(cudaEventSynchronize(c->ev[3 * i + 2]));
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, c->ev[3 * i], c->ev[3 * i + 1]));
        CU(cudaEventElapsedTime(&b, c->ev[3 * i + 1], c->ev[3 * i + 2]));
        br += a;
        ks += b;
    }
    if (blind_rotate_ms) *blind_rotate_ms = br;
    if (keyswitch_ms) *keyswitch_ms = ks;
    if (calls) *calls = n;
    return 0;
}

int tfhe_b200_ctx_words(const tfhe_b200_ctx *c) { return c ? c->p.n + 1 : 0; }

size_t tfhe_b200_key_bytes(const tfhe_b200_ctx *c) { return c ? c->bk_bytes + c->ks_bytes : 0; }
unsigned long long tfhe_b200_launch_count(const tfhe_b200_ctx *c) { return c ? c->launches.load() : 0; }
int tfhe_b200_sm_count(const tfhe_b200_ctx *c) { return c ? c->sm_count : 0; }
void tfhe_b200_count_launches(tfhe_b200_ctx *c, unsigned long long n) {
    if (c) c->launches += n;
}
int tfhe_b200_ctx_device(const tfhe_b200_ctx *c) { return c ? c->device : -1; }

int tfhe_b200_load_keys_device(tfhe_b200_ctx *c, const int32_t *d_bk_coef, const int32_t *d_ks, void *stream) {
    if (!c) return fail("null context");
    cudaStream_t st = (cudaStream_t) stream;
    CU(cudaSetDevice(c->device));
    const int n = c->p.n;
    if (d_bk_coef) {
        const size_t npolys = (size_t) n * kKpl * (kK + 1);
        if (!c->d_bk) {
            c->bk_bytes = npolys * kM * sizeof(cpx);
            CU(cudaMalloc(&c->d_bk, c->bk_bytes));
        }
        // Fourier(bk * 2^-32) * 2^32 / (N/2): the inverse transform is unnormalised
        CU(launch_forward_polys(d_bk_coef, c->d_bk, (int) npolys, 1.0 / 512.0, st));
        c->launches += 1;
    }
    if (d_ks) {
        const int N = c->p.N * c->p.k, t = c->p.ks_t, base = 1 << c->p.ks_basebit;
        if (!c->d_ks) {
            c->ks_bytes = (size_t) N * t * (base - 1) * kKsRowWords * sizeof(int32_t);
            CU(cudaMalloc(&c->d_ks, c->ks_bytes));
        }
        CU(launch_ks_relayout(d_ks, c->d_ks, N, t, base, n, st));
        c->launches += 1;
        if (ks_mma_supported(N, t, c->p.ks_basebit, n)) {
            if (!c->d_ks_mma) CU(cudaMalloc(&c->d_ks_mma, ks_mma_table_bytes()));
            CU(launch_ks_mma_relayout(d_ks, c->d_ks_mma, base, n, st));
            c->launches += 1;
        }
    }
    return 0;
}

int tfhe_b200_load_keys(tfhe_b200_ctx *c, const int32_t *bk_coef, const int32_t *ks) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    const int n = c->p.n;
    int32_t *d_tmp_bk = nullptr, *d_tmp_ks = nullptr;
    if (bk_coef) {
        const size_t bytes = (size_t) n * kKpl * (kK + 1) * kN * sizeof(int32_t);
        CU(cudaMalloc(&d_tmp_bk, bytes));
        CU(cudaMemcpyAsync(d_tmp_bk, bk_coef, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    if (ks) {
        const size_t bytes = (size_t) c->p.N * c->p.k * c->p.ks_t * (1 << c->p.ks_basebit) * (n + 1) * sizeof(int32_t);
        CU(cudaMalloc(&d_tmp_ks, bytes));
        CU(cudaMemcpyAsync(d_tmp_ks, ks, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    int rc = tfhe_b200_load_keys_device(c, d_tmp_bk, d_tmp_ks, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (d_tmp_bk) cudaFree(d_tmp_bk);
    if (d_tmp_ks) cudaFree(d_tmp_ks);
    if (rc) return rc;
    if (e != cudaSuccess) return fail("key upload failed: %s", cudaGetErrorString(e));
    return 0;
}

int tfhe_b200_load_ks(tfhe_b200_ctx *c, const int32_t *ks) { return tfhe_b200_load_keys(c, nullptr, ks); }

// Reference Fourier form -> device layout.  ref[j] = P(zeta^-(2j+1)); ours V_m = P(zeta^(4m+1)):
// m < 256: V_m = conj(ref[2m]);  m >= 256: V_m = ref[1023 - 2m].  Scale: ref is in torus
// units (int * 2^-32), ours in int * 2^-9  => factor 2^23.  Pure data re-layout.
int tfhe_b200_load_bk_fourier(tfhe_b200_ctx *c, const double *ref) {
    if (!c || !ref) return fail("null argument");
    CU(cudaSetDevice(c->device));
    const size_t npolys = (size_t) c->p.n * kKpl * (kK + 1);
    std::vector<cpx> host(npolys * kM);
    const double sc = 8388608.0;  // 2^23
    for (size_t q = 0; q < npolys; q++) {
        const double *src = ref + q * kM * 2;
        cpx *dst = host.data() + q * kM;
        for (int pos = 0; pos < 16; pos++)
            for (int m1 = 0; m1 < 32; m1++) {
                const int m = freq_of(pos, m1);
                cpx v;
                if (m < 256) {
                    v.x = src[2 * (2 * m)] * sc;
                    v.y = -src[2 * (2 * m) + 1] * sc;
                } else {
                    v.x = src[2 * (1023 - 2 * m)] * sc;
                    v.y = src[2 * (1023 - 2 * m) + 1] * sc;
                }
                dst[pos * 32 + m1] = v;
            }
    }
    if (!c->d_bk) {
        c->bk_bytes = npolys * kM * sizeof(cpx);
        CU(cudaMalloc(&c->d_bk, c->bk_bytes));
    }
    CU(cudaMemcpy(c->d_bk, host.data(), c->bk_bytes, cudaMemcpyHostToDevice));
    return 0;
}

// ---- device buffers through the library's own runtime (hosts without CUDA headers) ----------
int tfhe_b200_device_alloc(tfhe_b200_ctx *c, void **d_ptr, size_t bytes) {
    if (!c || !d_ptr) return fail("null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return 0;
}

int tfhe_b200_device_free(tfhe_b200_ctx *c, void *d_ptr) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    CU(cudaFree(d_ptr));
    return 0;
}

int tfhe_b200_copy_to_device(tfhe_b200_ctx *c, void *d_dst, const void *h_src, size_t bytes, void *stream) {
    if (!c || (!d_dst && bytes) || (!h_src && bytes)) return fail("null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
    CU(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}

int tfhe_b200_copy_to_host(tfhe_b200_ctx *c, void *h_dst, const void *d_src, size_t bytes, void *stream) {
    if (!c || (!h_dst && bytes) || (!d_src && bytes)) return fail("null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    CU(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}

int tfhe_b200_synchronize(tfhe_b200_ctx *c, void *stream) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}

// ------------------------------------------------------------------- gates --

int tfhe_b200_gate(tfhe_b200_ctx *c, int gate, int32_t *d_out, const int32_t *d_ca, const int32_t *d_cb,
                   int count, void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (gate < 0 || gate >= TFHE_B200_NUM_GATES) return fail("bad gate id %d", gate);
    if (count < 0) return fail("negative count");
    if (count == 0) return 0;
    BrLaunch L = base_launch(c);
    L.nseg = 1;
    L.total = count;
    set_segment(L.seg[0], kGates[gate], d_ca, d_cb, c->p.n, count);
    return run_bootstrap_ks(c, L, 1, 0, d_out, count, (cudaStream_t) stream);
}

int tfhe_b200_gate2(tfhe_b200_ctx *c, int gate0, int gate1, int32_t *d_out, const int32_t *d_ca,
                    const int32_t *d_cb, int count, void *stream) {
    return tfhe_b200_gate_pair(c, gate0, d_ca, d_cb, gate1, d_ca, d_cb, d_out, count, stream);
}

int tfhe_b200_gate_pair(tfhe_b200_ctx *c, int gate0, const int32_t *d_a0, const int32_t *d_b0, int gate1,
                        const int32_t *d_a1, const int32_t *d_b1, int32_t *d_out, int count, void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (gate0 < 0 || gate0 >= TFHE_B200_NUM_GATES || gate1 < 0 || gate1 >= TFHE_B200_NUM_GATES)
        return fail("bad gate id");
    if (count < 0) return fail("negative count");
    if (count == 0) return 0;
    BrLaunch L = base_launch(c);
    L.nseg = 2;
    L.total = 2 * count;
    set_segment(L.seg[0], kGates[gate0], d_a0, d_b0, c->p.n, count);
    set_segment(L.seg[1], kGates[gate1], d_a1, d_b1, c->p.n, count);
    return run_bootstrap_ks(c, L, 1, 0, d_out, 2 * count, (cudaStream_t) stream);
}

// Up to 4 independent runs of gates (each with its own gate type, strided inputs and strided
// output) in ONE bootstrap batch.  Generalises the reference's compound gates
// (bootsANDXOR / bootsXORXOR_fullGPU_n_Bit_vector, boot-gates.cu:3027-3098) and lets the
// circuit schedules address bit i of every number of a vector without copies.
int tfhe_b200_gate_multi(tfhe_b200_ctx *c, const tfhe_b200_gate_op *ops, int nops, void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (!ops || nops < 1 || nops > kMaxSegments) return fail("nops must be 1..%d", kMaxSegments);
    BrLaunch L = base_launch(c);
    KsLaunch::Out dst[kMaxSegments];
    int total = 0, nseg = 0;
    for (int i = 0; i < nops; i++) {
        const tfhe_b200_gate_op &o = ops[i];
        if (o.gate < 0 || o.gate >= TFHE_B200_NUM_GATES_EXT) return fail("bad gate id %d", o.gate);
        if (o.count < 0) return fail("negative count");
        if (o.count == 0) continue;
        if ((kGates[o.gate].sc != 0) != (o.c != nullptr)) return fail("gate %d: third operand mismatch", o.gate);
        if ((kGates[o.gate].sd != 0) != (o.d != nullptr)) return fail("gate %d: fourth operand mismatch", o.gate);
        BrSegment &sg = L.seg[nseg];
        sg.in0 = o.a;
        sg.in1 = o.b;
        sg.stride0 = o.stride_a;
        sg.stride1 = o.stride_b;
        sg.sa = kGates[o.gate].sa;
        sg.sb = kGates[o.gate].sb;
        sg.cst = kGates[o.gate].cst;
        sg.count = o.count;
        sg.idx0 = o.idx_a;
        sg.idx1 = o.idx_b;
        sg.in2 = o.c;
        sg.stride2 = o.stride_c;
        sg.idx2 = o.idx_c;
        sg.sc = kGates[o.gate].sc;
        sg.in3 = o.d;
        sg.stride3 = o.stride_d;
        sg.idx3 = o.idx_d;
        sg.sd = kGates[o.gate].sd;
        dst[nseg].out = o.out;
        dst[nseg].stride = o.stride_out;
        dst[nseg].count = o.count;
        dst[nseg].idx = o.idx_out;
        total += o.count;
        nseg++;
    }
    if (total == 0) return 0;
    L.nseg = nseg;
    L.total = total;
    return run_bootstrap_ks(c, L, 1, 0, nullptr, total, (cudaStream_t) stream, dst, nseg);
}

// bootsMUX (boot-gates.cu:407-448): u1 = BS(-1/8 + a + b), u2 = BS(-1/8 - a + c),
// result = KS((0,1/8) + u1 + u2)
int tfhe_b200_mux(tfhe_b200_ctx *c, int32_t *d_out, const int32_t *d_a, const int32_t *d_b,
                  const int32_t *d_c, int count, void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (count < 0) return fail("negative count");
    if (count == 0) return 0;
    BrLaunch L = base_launch(c);
    L.nseg = 2;
    L.total = 2 * count;
    const GateDef g1 = {-kMu, 1, 1}, g2 = {-kMu, -1, 1};
    set_segment(L.seg[0], g1, d_a, d_b, c->p.n, count);
    set_segment(L.seg[1], g2, d_a, d_c, c->p.n, count);
    return run_bootstrap_ks(c, L, 2, kMu, d_out, count, (cudaStream_t) stream);
}

// bootsMUX on rows of one sample array: out row idx_out[g] = MUX(row idx_a[g], row idx_b[g], row idx_c[g])
// (rows are `stride` words apart; index arrays in device memory).  Used by the circuit plans.
int tfhe_b200_mux_gather(tfhe_b200_ctx *c, int32_t *d_rows, int64_t stride, const int32_t *idx_a,
                         const int32_t *idx_b, const int32_t *idx_c, const int32_t *idx_out, int count,
                         void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (count < 0) return fail("negative count");
    if (count == 0) return 0;
    if (!d_rows || !idx_a || !idx_b || !idx_c || !idx_out) return fail("null argument");
    BrLaunch L = base_launch(c);
    L.nseg = 2;
    L.total = 2 * count;
    const GateDef g1 = {-kMu, 1, 1}, g2 = {-kMu, -1, 1};
    set_segment(L.seg[0], g1, d_rows, d_rows, c->p.n, count);
    set_segment(L.seg[1], g2, d_rows, d_rows, c->p.n, count);
    for (int s = 0; s < 2; s++) {
        L.seg[s].stride0 = L.seg[s].stride1 = stride;
        L.seg[s].idx0 = idx_a;
    }
    L.seg[0].idx1 = idx_b;
    L.seg[1].idx1 = idx_c;
    KsLaunch::Out dst;
    dst.out = d_rows;
    dst.stride = stride;
    dst.count = count;
    dst.idx = idx_out;
    return run_bootstrap_ks(c, L, 2, kMu, nullptr, count, (cudaStream_t) stream, &dst, 1);
}

// Bootstrap-free ops on rows of one sample array: out row idx_out[g] = coef * (row idx_in[g]) + (0, cst)
// (coef = 1: bootsCOPY, coef = -1: bootsNOT, coef = 0 and cst = +-1/8: bootsCONSTANT).
int tfhe_b200_linear_gather(tfhe_b200_ctx *c, int32_t *d_rows, int64_t stride, const int32_t *idx_in,
                            const int32_t *idx_out, int coef, int32_t cst, int count, void *stream) {
    if (check_ctx(c, false, false)) return 1;
    if (count < 0) return fail("negative count");
    if (count == 0) return 0;
    if (!d_rows || !idx_out || (coef != 0 && !idx_in)) return fail("null argument");
    CU(cudaSetDevice(c->device));
    CU(launch_lwe_linear_idx(d_rows, d_rows, stride, idx_out, idx_in, coef, cst, count, c->p.n, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

int tfhe_b200_not(tfhe_b200_ctx *c, int32_t *d_out, const int32_t *d_ca, int count, void *stream) {
    if (check_ctx(c, false, false)) return 1;
    CU(cudaSetDevice(c->device));
    const int s = c->p.n + 1;
    CU(launch_lwe_linear(d_out, s, d_ca, s, -1, nullptr, 0, 0, 0, count, c->p.n, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

int tfhe_b200_copy(tfhe_b200_ctx *c, int32_t *d_out, const int32_t *d_ca, int count, void *stream) {
    if (check_ctx(c, false, false)) return 1;
    CU(cudaSetDevice(c->device));
    if (d_out != d_ca)
        CU(cudaMemcpyAsync(d_out, d_ca, (size_t) count * (c->p.n + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                           (cudaStream_t) stream));
    return 0;
}

int tfhe_b200_constant(tfhe_b200_ctx *c, int32_t *d_out, int value, int count, void *stream) {
    if (check_ctx(c, false, false)) return 1;
    CU(cudaSetDevice(c->device));
    const int s = c->p.n + 1;
    CU(launch_lwe_linear(d_out, s, nullptr, 0, 0, nullptr, 0, 0, value ? kMu : -kMu, count, c->p.n,
                         (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

// --------------------------------------------------------- building blocks --

int tfhe_b200_bootstrap_woks(tfhe_b200_ctx *c, int32_t *d_u, const int32_t *d_x, int32_t mu, int count,
                             void *stream) {
    if (check_ctx(c, true, false)) return 1;
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    BrLaunch L = base_launch(c);
    L.nseg = 1;
    L.total = count;
    L.mu = mu;
    const GateDef id = {0, 1, 0};
    set_segment(L.seg[0], id, d_x, d_x, c->p.n, count);
    L.u_out = d_u;
    CU(launch_blind_rotate(L, c->sm_count, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

int tfhe_b200_bootstrap(tfhe_b200_ctx *c, int32_t *d_out, const int32_t *d_x, int32_t mu, int count,
                        void *stream) {
    if (check_ctx(c, true, true)) return 1;
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    BrLaunch L = base_launch(c);
    L.nseg = 1;
    L.total = count;
    L.mu = mu;
    const GateDef id = {0, 1, 0};
    set_segment(L.seg[0], id, d_x, d_x, c->p.n, count);
    return run_bootstrap_ks(c, L, 1, 0, d_out, count, (cudaStream_t) stream);
}

int tfhe_b200_keyswitch(tfhe_b200_ctx *c, int32_t *d_out, const int32_t *d_u, int count, void *stream) {
    if (check_ctx(c, false, true)) return 1;
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    KsLaunch K;
    memset(&K, 0, sizeof(K));
    K.ks = c->d_ks;
    K.u = d_u;
    K.nsrc = 1;
    K.dst[0].out = d_out;
    K.dst[0].stride = c->p.n + 1;
    K.dst[0].count = count;
    K.ndst = 1;
    K.count = count;
    K.n = c->p.n;
    K.N = c->p.N * c->p.k;
    K.t = c->p.ks_t;
    K.basebit = c->p.ks_basebit;
    return run_keyswitch(c, K, (cudaStream_t) stream);
}

int tfhe_b200_blind_rotate(tfhe_b200_ctx *c, int32_t *d_acc, const int32_t *d_bara, int n_iter, int count,
                           void *stream) {
    if (check_ctx(c, true, false)) return 1;
    if (n_iter < 0 || n_iter > c->p.n) return fail("n_iter %d out of range", n_iter);
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    BrLaunch L = base_launch(c);
    L.total = count;
    L.n_iter = n_iter;
    L.explicit_inputs = 1;
    L.bara = d_bara;
    L.acc_in = d_acc;
    L.acc_out = d_acc;
    CU(launch_blind_rotate(L, c->sm_count, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

int tfhe_b200_blind_rotate_and_extract(tfhe_b200_ctx *c, int32_t *d_u, const int32_t *d_testvect,
                                       const int32_t *d_barb, const int32_t *d_bara, int n_iter, int count,
                                       void *stream) {
    if (check_ctx(c, true, false)) return 1;
    if (n_iter < 0 || n_iter > c->p.n) return fail("n_iter %d out of range", n_iter);
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    BrLaunch L = base_launch(c);
    L.total = count;
    L.n_iter = n_iter;
    L.explicit_inputs = 1;
    L.bara = d_bara;
    L.barb = d_barb;
    L.testvect = d_testvect;
    L.u_out = d_u;
    CU(launch_blind_rotate(L, c->sm_count, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

int tfhe_b200_extern_mul(tfhe_b200_ctx *c, int32_t *d_acc, int bk_index, int count, void *stream) {
    if (check_ctx(c, true, false)) return 1;
    if (bk_index < 0 || bk_index >= c->p.n) return fail("bk_index %d out of range", bk_index);
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    BrLaunch L = base_launch(c);
    L.total = count;
    L.n_iter = 1;
    L.extern_only = 1;
    L.bk_first = bk_index;
    L.explicit_inputs = 1;
    L.acc_in = d_acc;
    L.acc_out = d_acc;
    CU(launch_blind_rotate(L, c->sm_count, (cudaStream_t) stream));
    c->launches += 1;
    return 0;
}

// ------------------------------------------------------ host-buffer variants --

// Host-buffer gate batch (the reference-facing call: inputs and outputs are host arrays).
// Batches of four waves or more are cut into chunks of whole waves (4 ciphertexts per SM) and pipelined
// over three streams: the H2D copy of chunk i+1 and the D2H copy of chunk i-1 run under the kernels of
// chunk i, so the call costs the kernels plus the copies of one wave instead of all of them.
static int host_gate_common(tfhe_b200_ctx *c, int gate, bool mux, int32_t *out, const int32_t *a, const int32_t *b,
                            const int32_t *cc, int count) {
    if (check_ctx(c, true, true)) return 1;
    if (count <= 0) return count < 0 ? fail("negative count") : 0;
    CU(cudaSetDevice(c->device));
    const size_t row = (size_t) (c->p.n + 1);
    // Memory-capped sub-batching (the reference: cudaMemGetInfo -> bootsLimit -> loop over sub-batches,
    // boot-gates.cu:2869-2907): operands, result and the extracted samples of a sub-batch must fit in 80 % of
    // the free device memory.  The query costs milliseconds (it was 4 % of a 65536-gate call when made
    // unconditionally), so it is made only when the whole batch could not be allocated, or when
    // TFHE_B200_HOST_BATCH_LIMIT (gates; tests) sets a cap.
    static const long long env_cap = [] {
        const char *v = getenv("TFHE_B200_HOST_BATCH_LIMIT");
        return v ? atoll(v) : 0ll;
    }();
    auto sub_batches = [&](size_t cap) -> int {
        const size_t wave = 4 * (size_t) c->sm_count;
        if (cap > wave) cap -= cap % wave;  // whole waves
        for (size_t g0 = 0; g0 < (size_t) count; g0 += cap) {
            const int nsub = (int) ((size_t) count - g0 < cap ? (size_t) count - g0 : cap);
            const size_t off = g0 * row;
            if (host_gate_common(c, gate, mux, out + off, a + off, b + off, mux ? cc + off : nullptr, nsub)) return 1;
        }
        return 0;
    };
    if (env_cap > 0 && (long long) count > env_cap) return sub_batches((size_t) env_cap);
    const size_t bytes = (size_t) count * row * sizeof(int32_t);
    int32_t *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_o = nullptr;
    cudaStream_t st = c->stream;
    auto release = [&]() {  // stream-ordered: safe after any error below
        if (d_a) cudaFreeAsync(d_a, st);
        if (d_b) cudaFreeAsync(d_b, st);
        if (d_o) cudaFreeAsync(d_o, st);
        if (d_c) cudaFreeAsync(d_c, st);
        return cudaStreamSynchronize(st);
    };
    if (cudaMallocAsync(&d_a, bytes, st) != cudaSuccess || cudaMallocAsync(&d_b, bytes, st) != cudaSuccess ||
        cudaMallocAsync(&d_o, bytes, st) != cudaSuccess || (mux && cudaMallocAsync(&d_c, bytes, st) != cudaSuccess)) {
        const cudaError_t e = cudaGetLastError();
        release();
        d_a = d_b = d_c = d_o = nullptr;
        // the whole batch does not fit: size sub-batches from the free memory
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_gate = ((mux ? 4 : 3) * row + (mux ? 2 : 1) * (size_t) (kN + 1)) * sizeof(int32_t);
        const size_t cap = (size_t) ((double) free_b * 0.8 / (double) per_gate);
        if (cap == 0 || cap >= (size_t) count)
            return fail("device allocation of %zu bytes per operand failed: %s", bytes, cudaGetErrorString(e));
        return sub_batches(cap);
    }
    // chunk boundaries in whole waves (4 ciphertexts per SM): 1, 3, 8, 16, 16, ..., 8, 3, 1.  Only the H2D
    // copy of the first chunk and the D2H copy of the last one are exposed, so both are small; every other
    // copy runs under a kernel at least as long as itself even at a tenth of the nominal PCIe rate (one wave
    // = 3.5 ms of kernel and 2.4 MB in / 1.2 MB out).  Batches below four waves go in one piece.
    const int wave = 4 * c->sm_count;
    std::vector<int> bounds;  // chunk i = gates [bounds[i], bounds[i + 1])
    {
        int rem = (count + wave - 1) / wave;
        std::vector<int> head, tail, waves;
        for (int s : {1, 3, 8})
            if (rem >= 2 * s + 2) {
                head.push_back(s);
                tail.push_back(s);
                rem -= 2 * s;
            }
        waves = head;
        for (; rem > 0; rem -= 16) waves.push_back(rem < 16 ? rem : 16);
        for (size_t i = tail.size(); i-- > 0;) waves.push_back(tail[i]);
        long long g = 0;
        bounds.push_back(0);
        for (int w : waves) {
            g += (long long) w * wave;
            bounds.push_back(g < count ? (int) g : count);
        }
    }
    const int nchunks = (int) bounds.size() - 1;
    int rc = 0;
    if (nchunks == 1) {
        cudaError_t e = cudaMemcpyAsync(d_a, a, bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, b, bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && mux) e = cudaMemcpyAsync(d_c, cc, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) rc = fail("H2D copy failed: %s", cudaGetErrorString(e));
        if (!rc)
            rc = mux ? tfhe_b200_mux(c, d_o, d_a, d_b, d_c, count, st) : tfhe_b200_gate(c, gate, d_o, d_a, d_b, count, st);
        if (!rc) {
            e = cudaMemcpyAsync(out, d_o, bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = fail("D2H copy failed: %s", cudaGetErrorString(e));
        }
    } else {
        std::vector<cudaEvent_t> ev(2 * (size_t) nchunks + 1, nullptr);
        for (auto &e : ev)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
        // the buffers were allocated on `st`: the copy streams may touch them only after that
        if (!rc && (cudaEventRecord(ev[2 * nchunks], st) != cudaSuccess ||
                    cudaStreamWaitEvent(c->copy_in, ev[2 * nchunks], 0) != cudaSuccess))
            rc = fail("stream setup failed");
        // Issue order: H2D(i), kernels(i), D2H(i - 1).  With PAGEABLE host memory cudaMemcpyAsync blocks the
        // calling thread (the driver stages the data itself; a D2H additionally waits for its kernels), so the
        // D2H of a chunk is issued only after the kernels of the NEXT chunk are queued: the GPU never waits for
        // the host thread.  (D2H(i) right behind kernels(i) left the GPU idle during every H2D: 470 instead
        // of 387 ms per 65536 gates.)  With pinned memory the order makes no difference.
        auto copy_out = [&](int i) {
            const int g0 = bounds[i], n = bounds[i + 1] - g0;
            const size_t off = (size_t) g0 * row, nb = (size_t) n * row * sizeof(int32_t);
            cudaError_t e = cudaStreamWaitEvent(c->copy_out, ev[2 * i + 1], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out + off, d_o + off, nb, cudaMemcpyDeviceToHost, c->copy_out);
            if (e != cudaSuccess) rc = fail("D2H copy failed: %s", cudaGetErrorString(e));
        };
        int issued = 0;  // chunks whose kernels are queued
        for (int i = 0; i < nchunks && !rc; i++) {
            const int g0 = bounds[i], n = bounds[i + 1] - g0;
            const size_t off = (size_t) g0 * row, nb = (size_t) n * row * sizeof(int32_t);
            cudaError_t e = cudaMemcpyAsync(d_a + off, a + off, nb, cudaMemcpyHostToDevice, c->copy_in);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_b + off, b + off, nb, cudaMemcpyHostToDevice, c->copy_in);
            if (e == cudaSuccess && mux) e = cudaMemcpyAsync(d_c + off, cc + off, nb, cudaMemcpyHostToDevice, c->copy_in);
            if (e == cudaSuccess) e = cudaEventRecord(ev[2 * i], c->copy_in);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev[2 * i], 0);
            if (e != cudaSuccess) {
                rc = fail("H2D copy failed: %s", cudaGetErrorString(e));
                break;
            }
            rc = mux ? tfhe_b200_mux(c, d_o + off, d_a + off, d_b + off, d_c + off, n, st)
                     : tfhe_b200_gate(c, gate, d_o + off, d_a + off, d_b + off, n, st);
            if (rc) break;
            e = cudaEventRecord(ev[2 * i + 1], st);
            if (e != cudaSuccess) {
                rc = fail("event record failed: %s", cudaGetErrorString(e));
                break;
            }
            issued = i + 1;
            if (i > 0) copy_out(i - 1);
        }
        if (!rc && issued > 0) copy_out(issued - 1);
        // everything the copy streams did must be over before the buffers go back to the pool
        cudaError_t e1 = cudaStreamSynchronize(c->copy_in), e2 = cudaStreamSynchronize(c->copy_out);
        if (!rc && (e1 != cudaSuccess || e2 != cudaSuccess))
            rc = fail("gate batch failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
    }
    const cudaError_t e = release();
    if (!rc && e != cudaSuccess) rc = fail("gate batch failed: %s", cudaGetErrorString(e));
    return rc;
}

int tfhe_b200_gate_host(tfhe_b200_ctx *c, int gate, int32_t *out, const int32_t *ca, const int32_t *cb, int count) {
    if (gate < 0 || gate >= TFHE_B200_NUM_GATES) return fail("bad gate id %d", gate);
    return host_gate_common(c, gate, false, out, ca, cb, nullptr, count);
}

int tfhe_b200_mux_host(tfhe_b200_ctx *c, int32_t *out, const int32_t *a, const int32_t *b, const int32_t *cc,
                       int count) {
    return host_gate_common(c, 0, true, out, a, b, cc, count);
}

}  // extern "C"
