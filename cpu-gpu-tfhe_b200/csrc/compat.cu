// Drop-in replacements for the reference's entry points (include/tfhe_compat.h): they walk
// the reference's pointer-rich structs, keep one GPU context per key object (created on first
// use, thread-safe) and call the flat C ABI.  Failures abort, like die_dramatically()
// (tfhe_gate_bootstrapping.cu:11-15) and the CUDA error macros of boot-gates.cu:33-86.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/tfhe_compat.h"

namespace {

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

int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return 0;
    return d;
}

// ---- content fingerprints ---------------------------------------------------------------------
// A GPU context is found by the ADDRESS of the host key object, which says nothing about its
// content: a key that is re-encrypted in place, or freed with a new one landing at the same address
// (tGswFFTExternMulToTLwe / tfhe_blindRotate_FFT in loops do exactly that), must not be served
// from a stale device copy.  Every entry therefore stores a fingerprint of the key MATERIAL
// (FNV-1a over the whole object when it is small, over an evenly spread sample of every row when it is
// tens of megabytes) which is recomputed and compared at every call; on a mismatch the context is
// rebuilt.
struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void add(const void *p, size_t bytes) {
        const unsigned char *c = (const unsigned char *) p;
        for (size_t i = 0; i < bytes; i++) h = (h ^ c[i]) * 1099511628211ull;
    }
    void add_words(const void *p, size_t bytes) {  // 8 bytes at a time (bulk data)
        const uint64_t *w = (const uint64_t *) p;
        for (size_t i = 0; i < bytes / 8; i++) h = (h ^ w[i]) * 1099511628211ull;
    }
};

// whole polynomial when `full`, else 64 bytes at a row-dependent offset
void fp_poly(Fnv &f, const void *data, size_t bytes, bool full, size_t salt) {
    if (full || bytes <= 64) f.add_words(data, bytes);
    else f.add_words((const char *) data + ((salt * 104729u) % (bytes / 64)) * 64, 64);
}

uint64_t fingerprint_tgsw_fft(const TGswSampleFFT *bk, int n, const TGswParams *gp) {
    const int Ns2 = gp->tlwe_params->N / 2, k = gp->tlwe_params->k, kpl = gp->kpl;
    Fnv f;
    f.add(&n, sizeof(n));
    const bool full = n <= 4;  // a lone TGSW sample (64 KiB) is hashed completely
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++)
                fp_poly(f, bk[i].all_samples[r].a[j].data, sizeof(double) * 2 * Ns2, full, (size_t) i * 8 + r * 2 + j);
    return f.h;
}

uint64_t fingerprint_tgsw(const TGswSample *bk, int n, const TGswParams *gp) {
    const int N = gp->tlwe_params->N, k = gp->tlwe_params->k, kpl = gp->kpl;
    Fnv f;
    f.add(&n, sizeof(n));
    for (int i = 0; i < n; i++)
        for (int r = 0; r < kpl; r++)
            for (int j = 0; j <= k; j++)
                fp_poly(f, bk[i].all_sample[r].a[j].coefsT, sizeof(int32_t) * N, n <= 4, (size_t) i * 8 + r * 2 + j);
    return f.h;
}

uint64_t fingerprint_ks(const LweKeySwitchKey *ks) {
    const int n = ks->out_params->n;
    Fnv f;
    f.add(&ks->n, sizeof(int));
    f.add(&ks->t, sizeof(int));
    f.add(&ks->basebit, sizeof(int));
    for (int i = 0; i < ks->n; i++)
        for (int j = 0; j < ks->t; j++) {
            const LweSample *s = &ks->ks[i][j][1 + (i + j) % (ks->base - 1)];
            f.add_words(s->a + ((size_t) (i * 31 + j * 7) % (size_t) (n / 16 ? n / 16 : 1)) * 16, 64 <= n * 4 ? 64 : (size_t) n * 4);
            f.add(&s->b, sizeof(int32_t));
        }
    return f.h;
}

// ---- gate coalescer ---------------------------------------------------------------------------
// The classic entry points are called from OpenMP worker threads, one sample per call
// (cpuParallel/Cipher.cpp:75-76, 94; SURVEY 8b "Threading").  Requests of all threads that arrive
// while a batch is in flight — or within a short window of the first one — are combined into ONE
// tfhe_b200_gate_multi / tfhe_b200_mux launch on a private non-blocking stream with cached pinned
// staging and device scratch: k concurrent callers cost about one gate latency instead of k, and a
// lone caller no longer pays cudaMalloc / six synchronous copies / cudaFree per gate.
// Leader / follower: the first thread to find no leader runs the batch for everybody queued.
struct GateReq {
    int gate;           // TFHE_B200_* or -1 for MUX
    const LweSample *a, *b, *c;
    LweSample *out;
    bool done = false;
};

class Coalescer {
  public:
    Coalescer(tfhe_b200_ctx *ctx, int n) : ctx_(ctx), n_(n) {
        cu(cudaSetDevice(tfhe_b200_ctx_device(ctx)), "cudaSetDevice");
        cu(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < kSideStreams; i++) {
            cu(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking), "cudaStreamCreate");
            cu(cudaEventCreateWithFlags(&side_done_[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        cu(cudaEventCreateWithFlags(&staged_, cudaEventDisableTiming), "cudaEventCreate");
        const size_t words = (size_t) kMaxBatch * 4 * (n + 1);
        cu(cudaMallocHost(&h_, words * sizeof(int32_t)), "cudaMallocHost");
        cu(cudaMalloc(&d_, words * sizeof(int32_t)), "cudaMalloc");
    }
    ~Coalescer() {
        cudaStreamSynchronize(stream_);
        cudaFree(d_);
        cudaFreeHost(h_);
        for (int i = 0; i < kSideStreams; i++) {
            cudaEventDestroy(side_done_[i]);
            cudaStreamDestroy(side_[i]);
        }
        cudaEventDestroy(staged_);
        cudaStreamDestroy(stream_);
    }
    void submit(GateReq &r) {
        std::unique_lock<std::mutex> lk(mu_);
        queue_.push_back(&r);
        while (!r.done) {
            if (!leader_) {
                leader_ = true;
                if (queue_.size() == 1 && window_us_ > 0) {  // alone: give simultaneous callers a moment to join
                    lk.unlock();
                    std::this_thread::sleep_for(std::chrono::microseconds(window_us_));
                    lk.lock();
                }
                std::vector<GateReq *> batch;
                const size_t take = queue_.size() < (size_t) kMaxBatch ? queue_.size() : (size_t) kMaxBatch;
                batch.assign(queue_.begin(), queue_.begin() + take);
                queue_.erase(queue_.begin(), queue_.begin() + take);
                lk.unlock();
                run(batch);
                lk.lock();
                for (GateReq *q : batch) q->done = true;
                batches_++;
                gates_ += batch.size();
                leader_ = false;
                cv_.notify_all();
            } else {
                cv_.wait(lk);
            }
        }
    }
    void stats(unsigned long long *batches, unsigned long long *gates) {
        std::lock_guard<std::mutex> lk(mu_);
        *batches = batches_;
        *gates = gates_;
    }
    static constexpr int kMaxBatch = 1024;

  private:
    // staging layout (host and device alike): rows of n+1 words; block 0 = first operands, block 1 =
    // second operands, block 2 = third operands (MUX), block 3 = results; requests sorted by gate type
    void run(std::vector<GateReq *> &batch) {
        std::stable_sort(batch.begin(), batch.end(), [](const GateReq *x, const GateReq *y) { return x->gate < y->gate; });
        const size_t row = (size_t) n_ + 1, blk = (size_t) kMaxBatch * row;
        const int count = (int) batch.size();
        for (int i = 0; i < count; i++) {
            const GateReq *q = batch[i];
            const LweSample *src[3] = {q->a, q->b, q->c};
            for (int o = 0; o < 3; o++) {
                if (!src[o]) continue;
                int32_t *dst = h_ + o * blk + (size_t) i * row;
                memcpy(dst, src[o]->a, sizeof(int32_t) * n_);
                dst[n_] = src[o]->b;
            }
        }
        cu(cudaSetDevice(tfhe_b200_ctx_device(ctx_)), "cudaSetDevice");
        int nmux = 0;
        while (nmux < count && batch[nmux]->gate < 0) nmux++;  // MUX requests sort first
        const size_t nb = (size_t) count * row * sizeof(int32_t);
        cu(cudaMemcpyAsync(d_, h_, nb, cudaMemcpyHostToDevice, stream_), "H2D");
        cu(cudaMemcpyAsync(d_ + blk, h_ + blk, nb, cudaMemcpyHostToDevice, stream_), "H2D");
        if (nmux)
            cu(cudaMemcpyAsync(d_ + 2 * blk, h_ + 2 * blk, (size_t) nmux * row * sizeof(int32_t), cudaMemcpyHostToDevice, stream_), "H2D");
        // A batch with MUX requests or more than four gate types needs several launches (each = one blind
        // rotation + one key switch).  They are independent and small (one ciphertext per SM), so they go
        // to side streams and run CONCURRENTLY on different SMs instead of one after the other.
        int launch = 0;
        cudaStream_t used[1 + kSideStreams];
        auto next_stream = [&]() {
            cudaStream_t st = launch == 0 ? stream_ : side_[(launch - 1) % kSideStreams];
            if (launch == 1) cu(cudaEventRecord(staged_, stream_), "event");
            if (launch >= 1 && launch <= kSideStreams) cu(cudaStreamWaitEvent(st, staged_, 0), "wait");
            used[launch <= kSideStreams ? launch : kSideStreams] = st;
            launch++;
            return st;
        };
        // runs of equal gate type, up to 4 per launch
        tfhe_b200_gate_op ops[4];
        int nops = 0;
        for (int i = nmux; i < count;) {
            int j = i;
            while (j < count && batch[j]->gate == batch[i]->gate) j++;
            tfhe_b200_gate_op &o = ops[nops++];
            memset(&o, 0, sizeof(o));
            o.gate = batch[i]->gate;
            o.count = j - i;
            o.a = d_ + (size_t) i * row;
            o.b = d_ + blk + (size_t) i * row;
            o.out = d_ + 3 * blk + (size_t) i * row;
            o.stride_a = o.stride_b = o.stride_out = (int64_t) row;
            i = j;
            if (nops == 4 || i == count) {
                OK(tfhe_b200_gate_multi(ctx_, ops, nops, next_stream()), "gate");
                nops = 0;
            }
        }
        if (nmux) OK(tfhe_b200_mux(ctx_, d_ + 3 * blk, d_, d_ + blk, d_ + 2 * blk, nmux, next_stream()), "mux");
        const int nside = launch - 1 < kSideStreams ? launch - 1 : kSideStreams;
        for (int i = 0; i < nside; i++) {  // join the side streams
            cu(cudaEventRecord(side_done_[i], side_[i]), "event");
            cu(cudaStreamWaitEvent(stream_, side_done_[i], 0), "wait");
        }
        (void) used;
        cu(cudaMemcpyAsync(h_ + 3 * blk, d_ + 3 * blk, nb, cudaMemcpyDeviceToHost, stream_), "D2H");
        cu(cudaStreamSynchronize(stream_), "gate");
        for (int i = 0; i < count; i++) {
            const int32_t *srcrow = h_ + 3 * blk + (size_t) i * row;
            LweSample *out = batch[i]->out;
            memcpy(out->a, srcrow, sizeof(int32_t) * n_);
            out->b = srcrow[n_];
            out->current_variance = 0.;  // bookkeeping only; the reference's GPU path ignores it too (boot-gates.cu:2866)
        }
    }

    tfhe_b200_ctx *ctx_;
    int n_;
    static constexpr int kSideStreams = 3;
    cudaStream_t stream_ = nullptr, side_[kSideStreams] = {};
    cudaEvent_t side_done_[kSideStreams] = {}, staged_ = nullptr;
    int32_t *h_ = nullptr, *d_ = nullptr;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<GateReq *> queue_;
    bool leader_ = false;
    unsigned long long batches_ = 0, gates_ = 0;
    const int window_us_ = [] {
        const char *v = getenv("TFHE_B200_COALESCE_WINDOW_US");
        return v ? atoi(v) : 40;
    }();
};

// ---- context cache ------------------------------------------------------------------------------
enum Kind { KIND_CLOUD = 0, KIND_BKFFT = 1, KIND_TGSW = 2, KIND_KS = 3 };

struct Entry {
    tfhe_b200_ctx *ctx = nullptr;
    Coalescer *co = nullptr;
    uint64_t fp = 0;
    unsigned long long last_use = 0;
    ~Entry() {
        delete co;
        if (ctx) tfhe_b200_ctx_destroy(ctx);
    }
};
using EntryPtr = std::shared_ptr<Entry>;  // callers keep the entry alive while they use it

std::mutex g_mu;
std::map<std::pair<const void *, int>, EntryPtr> g_ctx;  // (key object, kind) -> context
unsigned long long g_tick = 0;
constexpr size_t kMaxSmallContexts = 8;  // bare TGSW samples / key-switch keys kept at a time (LRU)

EntryPtr lookup(const void *key, int kind, uint64_t fp) {
    auto it = g_ctx.find({key, kind});
    if (it == g_ctx.end()) return nullptr;
    if (it->second->fp != fp) {  // same address, different content: the device copy is stale
        g_ctx.erase(it);
        return nullptr;
    }
    it->second->last_use = ++g_tick;
    return it->second;
}

void insert(const void *key, int kind, const EntryPtr &e) {
    e->last_use = ++g_tick;
    g_ctx[{key, kind}] = e;
    if (kind != KIND_TGSW && kind != KIND_KS) return;
    size_t same = 0;
    for (auto &kv : g_ctx) same += kv.first.second == kind;
    while (same > kMaxSmallContexts) {  // evict the least recently used of this kind
        auto victim = g_ctx.end();
        for (auto it = g_ctx.begin(); it != g_ctx.end(); ++it)
            if (it->first.second == kind && (victim == g_ctx.end() || it->second->last_use < victim->second->last_use))
                victim = it;
        g_ctx.erase(victim);
        same--;
    }
}

uint64_t fingerprint_cloud(const TFheGateBootstrappingCloudKeySet *ck) {
    const int n = ck->params->in_out_params->n;
    Fnv f;
    if (ck->bk) {
        f.h ^= fingerprint_tgsw(ck->bk->bk, n, ck->params->tgsw_params);
        f.h = f.h * 1099511628211ull ^ fingerprint_ks(ck->bk->ks);
    } else {
        f.h ^= fingerprint_tgsw_fft(ck->bkFFT->bkFFT, n, ck->params->tgsw_params);
        f.h = f.h * 1099511628211ull ^ fingerprint_ks(ck->bkFFT->ks);
    }
    return f.h;
}

// context for a whole cloud key set (bootstrapping + key switch)
EntryPtr entry_for_cloud(const TFheGateBootstrappingCloudKeySet *ck) {
    // the fingerprint of a 100 MB key set samples ~0.6 MB: computed once per pointer and thread,
    // then revalidated every 1024 calls (cloud keys are immutable in every caller of the reference)
    thread_local const void *tl_key = nullptr;
    thread_local uint64_t tl_fp = 0;
    thread_local unsigned tl_calls = 0;
    if (tl_key != ck || (++tl_calls & 1023u) == 0) {
        tl_fp = fingerprint_cloud(ck);
        tl_key = ck;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    if (EntryPtr e = lookup(ck, KIND_CLOUD, tl_fp)) return e;
    const int n = ck->params->in_out_params->n;
    const tfhe_b200_params p = params_of(n, ck->params->tgsw_params, ck->params->ks_t, ck->params->ks_basebit);
    EntryPtr e = std::make_shared<Entry>();
    OK(tfhe_b200_ctx_create(&e->ctx, &p, current_device()), "context creation");
    if (ck->bk) {
        const std::vector<int32_t> bk = flatten_bk(ck->bk->bk, n, ck->params->tgsw_params);
        const std::vector<int32_t> ks = flatten_ks(ck->bk->ks);
        OK(tfhe_b200_load_keys(e->ctx, bk.data(), ks.data()), "key upload");
    } else {
        const std::vector<double> bk = flatten_bkfft(ck->bkFFT->bkFFT, n, ck->params->tgsw_params);
        const std::vector<int32_t> ks = flatten_ks(ck->bkFFT->ks);
        OK(tfhe_b200_load_bk_fourier(e->ctx, bk.data()), "key upload");
        OK(tfhe_b200_load_ks(e->ctx, ks.data()), "key upload");
    }
    e->co = new Coalescer(e->ctx, n);
    e->fp = tl_fp;
    insert(ck, KIND_CLOUD, e);
    return e;
}

// context for a Fourier bootstrapping key (+ its key-switch key)
EntryPtr entry_for_bkfft(const LweBootstrappingKeyFFT *bk) {
    const int n = bk->in_out_params->n;
    Fnv f;
    f.h ^= fingerprint_tgsw_fft(bk->bkFFT, n, bk->bk_params);
    f.h = f.h * 1099511628211ull ^ fingerprint_ks(bk->ks);
    std::lock_guard<std::mutex> lock(g_mu);
    if (EntryPtr e = lookup(bk, KIND_BKFFT, f.h)) return e;
    const tfhe_b200_params p = params_of(n, bk->bk_params, bk->ks->t, bk->ks->basebit);
    EntryPtr e = std::make_shared<Entry>();
    OK(tfhe_b200_ctx_create(&e->ctx, &p, current_device()), "context creation");
    const std::vector<double> fl = flatten_bkfft(bk->bkFFT, n, bk->bk_params);
    const std::vector<int32_t> ks = flatten_ks(bk->ks);
    OK(tfhe_b200_load_bk_fourier(e->ctx, fl.data()), "key upload");
    OK(tfhe_b200_load_ks(e->ctx, ks.data()), "key upload");
    e->fp = f.h;
    insert(bk, KIND_BKFFT, e);
    return e;
}

// context for a bare array of n TGSW samples in Fourier form (no key switch)
EntryPtr entry_for_tgsw(const TGswSampleFFT *bk, int n, const TGswParams *gp) {
    if (n < 1) n = 1;  // n = 0 (no iterations) still needs a context for the integer stages
    const uint64_t fp = fingerprint_tgsw_fft(bk, n, gp);  // includes n
    std::lock_guard<std::mutex> lock(g_mu);
    if (EntryPtr e = lookup(bk, KIND_TGSW, fp)) return e;
    const tfhe_b200_params p = params_of(n, gp, 8, 2);
    EntryPtr e = std::make_shared<Entry>();
    OK(tfhe_b200_ctx_create(&e->ctx, &p, current_device()), "context creation");
    const std::vector<double> f = flatten_bkfft(bk, n, gp);
    OK(tfhe_b200_load_bk_fourier(e->ctx, f.data()), "key upload");
    e->fp = fp;
    insert(bk, KIND_TGSW, e);
    return e;
}

EntryPtr entry_for_ks(const LweKeySwitchKey *ks) {
    const uint64_t fp = fingerprint_ks(ks);
    std::lock_guard<std::mutex> lock(g_mu);
    if (EntryPtr e = lookup(ks, KIND_KS, fp)) return e;
    tfhe_b200_params p;
    tfhe_b200_default_params(&p);
    p.n = ks->out_params->n;
    p.ks_t = ks->t;
    p.ks_basebit = ks->basebit;
    if (ks->n != p.N * p.k) die("lweKeySwitch: input dimension must be N*k = 1024");
    EntryPtr e = std::make_shared<Entry>();
    OK(tfhe_b200_ctx_create(&e->ctx, &p, current_device()), "context creation");
    const std::vector<int32_t> flat = flatten_ks(ks);
    OK(tfhe_b200_load_ks(e->ctx, flat.data()), "key upload");
    e->fp = fp;
    insert(ks, KIND_KS, e);
    return e;
}

// Device scratch + private stream of the calling thread for the single-object entry points
// (tfhe_bootstrap_FFT, tGswFFTExternMulToTLwe, lweKeySwitch ...): grown on demand, reused by every
// later call of the thread — no cudaMalloc / cudaFree per call, no legacy default stream.
struct ThreadScratch {
    int32_t *d = nullptr, *h = nullptr;
    size_t words = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
    ~ThreadScratch() { release(); }
    void release() {
        if (device < 0) return;
        // (a worker thread may outlive the CUDA context at process exit: errors are ignored here)
        cudaSetDevice(device);
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        if (stream) cudaStreamDestroy(stream);
        d = h = nullptr;
        stream = nullptr;
        words = 0;
        device = -1;
    }
    void ensure(int dev, size_t need) {
        if (device != dev) {
            release();
            cu(cudaSetDevice(dev), "cudaSetDevice");
            cu(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
            device = dev;
        } else {
            cu(cudaSetDevice(dev), "cudaSetDevice");
        }
        if (need > words) {
            if (d) cudaFree(d);
            if (h) cudaFreeHost(h);
            words = need < 8192 ? 8192 : need;
            cu(cudaMalloc(&d, words * sizeof(int32_t)), "cudaMalloc");
            cu(cudaMallocHost(&h, words * sizeof(int32_t)), "cudaMallocHost");
        }
    }
};
thread_local ThreadScratch tl_scratch;

// staged call: `in_words` words of the pinned buffer go up, fn runs on the thread's stream, words
// [out_off, out_off + out_words) come back
template <typename Fn>
void staged(tfhe_b200_ctx *c, size_t total_words, size_t in_words, size_t out_off, size_t out_words, Fn fn) {
    ThreadScratch &t = tl_scratch;
    cu(cudaMemcpyAsync(t.d, t.h, in_words * sizeof(int32_t), cudaMemcpyHostToDevice, t.stream), "H2D");
    fn(t.d, t.stream);
    cu(cudaMemcpyAsync(t.h + out_off, t.d + out_off, out_words * sizeof(int32_t), cudaMemcpyDeviceToHost, t.stream), "D2H");
    cu(cudaStreamSynchronize(t.stream), "synchronize");
    (void) c;
    (void) total_words;
}

void classic_gate(int gate, LweSample *result, const LweSample *ca, const LweSample *cb, const LweSample *cc,
                  const TFheGateBootstrappingCloudKeySet *ck) {
    EntryPtr e = entry_for_cloud(ck);
    GateReq r;
    r.gate = gate;
    r.a = ca;
    r.b = cb;
    r.c = cc;
    r.out = result;
    e->co->submit(r);
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

void bootsNAND(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_NAND, r, a, b, nullptr, bk); }
void bootsOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_OR, r, a, b, nullptr, bk); }
void bootsAND(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_AND, r, a, b, nullptr, bk); }
void bootsXOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_XOR, r, a, b, nullptr, bk); }
void bootsXNOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_XNOR, r, a, b, nullptr, bk); }
void bootsNOR(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_NOR, r, a, b, nullptr, bk); }
void bootsANDNY(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ANDNY, r, a, b, nullptr, bk); }
void bootsANDYN(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ANDYN, r, a, b, nullptr, bk); }
void bootsORNY(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ORNY, r, a, b, nullptr, bk); }
void bootsORYN(LweSample *r, const LweSample *a, const LweSample *b, const TFheGateBootstrappingCloudKeySet *bk) { classic_gate(TFHE_B200_ORYN, r, a, b, nullptr, bk); }

void bootsMUX(LweSample *result, const LweSample *a, const LweSample *b, const LweSample *cc,
              const TFheGateBootstrappingCloudKeySet *ck) {
    classic_gate(-1, result, a, b, cc, ck);
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
    EntryPtr e = entry_for_tgsw(bk, n, bk_params);
    tfhe_b200_ctx *c = e->ctx;
    const int N = bk_params->tlwe_params->N, k = bk_params->tlwe_params->k;
    const size_t accw = (size_t) (k + 1) * N;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), accw + n);
    for (int j = 0; j <= k; j++) memcpy(t.h + (size_t) j * N, accum->a[j].coefsT, sizeof(int32_t) * N);
    memcpy(t.h + accw, bara, sizeof(int32_t) * n);
    staged(c, accw + n, accw + n, 0, accw, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_blind_rotate(c, d, d + accw, n, 1, st), "blind rotate");
    });
    for (int j = 0; j <= k; j++) memcpy(accum->a[j].coefsT, t.h + (size_t) j * N, sizeof(int32_t) * N);
}

void tfhe_blindRotateAndExtract_FFT(LweSample *result, const TorusPolynomial *v, const TGswSampleFFT *bk,
                                    const int barb, const int *bara, const int n, const TGswParams *bk_params) {
    EntryPtr e = entry_for_tgsw(bk, n, bk_params);
    tfhe_b200_ctx *c = e->ctx;
    const int N = bk_params->tlwe_params->N, k = bk_params->tlwe_params->k;
    const size_t in = (size_t) N + 1 + n, uw = (size_t) k * N + 1;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), in + uw);
    memcpy(t.h, v->coefsT, sizeof(int32_t) * N);
    t.h[N] = barb;
    memcpy(t.h + N + 1, bara, sizeof(int32_t) * n);
    staged(c, in + uw, in, in, uw, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_blind_rotate_and_extract(c, d + in, d, d + N, d + N + 1, n, 1, st), "blind rotate");
    });
    memcpy(result->a, t.h + in, sizeof(int32_t) * (size_t) k * N);
    result->b = t.h[in + (size_t) k * N];
    result->current_variance = 0.;
}

void tfhe_bootstrap_woKS_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x) {
    EntryPtr e = entry_for_bkfft(bk);
    tfhe_b200_ctx *c = e->ctx;
    const int n = bk->in_out_params->n, Nk = bk->extract_params->n;
    const size_t in = (size_t) n + 1, uw = (size_t) Nk + 1;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), in + uw);
    memcpy(t.h, x->a, sizeof(int32_t) * n);
    t.h[n] = x->b;
    staged(c, in + uw, in, in, uw, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_bootstrap_woks(c, d + in, d, mu, 1, st), "bootstrap");
    });
    memcpy(result->a, t.h + in, sizeof(int32_t) * Nk);
    result->b = t.h[in + Nk];
    result->current_variance = 0.;
}

void tfhe_bootstrap_FFT(LweSample *result, const LweBootstrappingKeyFFT *bk, Torus32 mu, const LweSample *x) {
    EntryPtr e = entry_for_bkfft(bk);
    tfhe_b200_ctx *c = e->ctx;
    const int n = bk->in_out_params->n;
    const size_t row = (size_t) n + 1;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), 2 * row);
    memcpy(t.h, x->a, sizeof(int32_t) * n);
    t.h[n] = x->b;
    staged(c, 2 * row, row, row, row, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_bootstrap(c, d + row, d, mu, 1, st), "bootstrap");
    });
    memcpy(result->a, t.h + row, sizeof(int32_t) * n);
    result->b = t.h[row + n];
    result->current_variance = 0.;
}

void tGswFFTExternMulToTLwe(TLweSample *accum, const TGswSampleFFT *gsw, const TGswParams *params) {
    EntryPtr e = entry_for_tgsw(gsw, 1, params);
    tfhe_b200_ctx *c = e->ctx;
    const int N = params->tlwe_params->N, k = params->tlwe_params->k;
    const size_t accw = (size_t) (k + 1) * N;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), accw);
    for (int j = 0; j <= k; j++) memcpy(t.h + (size_t) j * N, accum->a[j].coefsT, sizeof(int32_t) * N);
    staged(c, accw, accw, 0, accw, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_extern_mul(c, d, 0, 1, st), "external product");
    });
    for (int j = 0; j <= k; j++) memcpy(accum->a[j].coefsT, t.h + (size_t) j * N, sizeof(int32_t) * N);
}

void lweKeySwitch(LweSample *result, const LweKeySwitchKey *ks, const LweSample *sample) {
    EntryPtr e = entry_for_ks(ks);
    tfhe_b200_ctx *c = e->ctx;
    const int n = ks->out_params->n, Nin = ks->n;
    const size_t in = (size_t) Nin + 1, ow = (size_t) n + 1;
    ThreadScratch &t = tl_scratch;
    t.ensure(tfhe_b200_ctx_device(c), in + ow);
    memcpy(t.h, sample->a, sizeof(int32_t) * Nin);
    t.h[Nin] = sample->b;
    staged(c, in + ow, in, in, ow, [&](int32_t *d, cudaStream_t st) {
        OK(tfhe_b200_keyswitch(c, d + in, d, 1, st), "key switch");
    });
    memcpy(result->a, t.h + in, sizeof(int32_t) * n);
    result->b = t.h[in + n];
    result->current_variance = 0.;
}

// ---- batched family ---------------------------------------------------------------------

// The opaque handle the batched family takes in place of the reference's raw device key pointers.
// It stays valid until tfhe_b200_keys_free / tfhe_b200_compat_invalidate of the same key set.
void *tfhe_b200_keys_to_gpu(const TFheGateBootstrappingCloudKeySet *bk) { return entry_for_cloud(bk)->ctx; }

// Drop every cached GPU context of a host key object (cloud key set, LweBootstrappingKeyFFT,
// TGswSampleFFT array or LweKeySwitchKey): call it before freeing or overwriting the object if its
// address may be reused.  (The content fingerprint catches a changed key anyway; this returns the
// device memory at once.)  Contexts still in use by another thread die with their last user.
void tfhe_b200_compat_invalidate(const void *key_object) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (int kind = 0; kind < 4; kind++) g_ctx.erase({key_object, kind});
}

void tfhe_b200_compat_release_all(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_ctx.clear();
}

int tfhe_b200_compat_cached_contexts(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    return (int) g_ctx.size();
}

// how many launches the gate coalescer of a cloud key set has issued, and for how many gates
void tfhe_b200_compat_coalescer_stats(const TFheGateBootstrappingCloudKeySet *bk, unsigned long long *batches,
                                      unsigned long long *gates) {
    entry_for_cloud(bk)->co->stats(batches, gates);
}

void tfhe_b200_keys_free(const TFheGateBootstrappingCloudKeySet *bk) { tfhe_b200_compat_invalidate(bk); }

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
