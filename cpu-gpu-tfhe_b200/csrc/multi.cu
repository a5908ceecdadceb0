// In-library multi-GPU for C / C++ hosts (include/tfhe_b200.h, "several GPUs in one process").
//
// The reference runs on device 0 only (cudaGetDeviceProperties(&p, 0), boot-gates.cu:3344).  Bootstrapped
// gates are independent, so a host batch is cut into contiguous shards, one per GPU; every
// GPU holds the full key material, nothing is exchanged on the data path.  The keys are uploaded to
// the first device ONCE and copied from there to the others device to device (cudaMemcpyPeerAsync:
// NVLink where peer access exists), then every GPU converts its copy itself.  One host thread per
// device drives the (synchronous, internally pipelined) host-buffer gate call of its shard.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200.h"

struct tfhe_b200_multi {
    tfhe_b200_params p;
    std::vector<int> devices;
    std::vector<tfhe_b200_ctx *> ctx;
};

namespace {

thread_local char g_merr[512] = "";

int mfail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_merr, sizeof(g_merr), fmt, ap);
    va_end(ap);
    return 1;
}

// contiguous split, sizes differ by at most one (the same rule as dist.shard_bounds)
void shard(long long total, int world, int rank, long long *lo, long long *hi) {
    const long long base = total / world, rem = total % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

template <typename Fn>
int for_each_device(tfhe_b200_multi *m, Fn fn) {
    const int nd = (int) m->ctx.size();
    std::vector<int> rc(nd, 0);
    std::vector<std::string> err(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d]() {
            rc[d] = fn(d);
            if (rc[d]) err[d] = tfhe_b200_last_error();
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (rc[d]) return mfail("device %d: %s", m->devices[d], err[d].c_str());
    return 0;
}

}  // namespace

extern "C" {

const char *tfhe_b200_multi_last_error(void) { return g_merr; }

int tfhe_b200_multi_create(tfhe_b200_multi **out, const tfhe_b200_params *p, const int *devices, int ndevices) {
    if (!out || !p) return mfail("null argument");
    *out = nullptr;
    int avail = tfhe_b200_device_count();
    if (avail <= 0) return mfail("no CUDA device available: the engine has no CPU fallback");
    tfhe_b200_multi *m = new (std::nothrow) tfhe_b200_multi();
    if (!m) return mfail("out of host memory");
    m->p = *p;
    if (!devices || ndevices <= 0) {  // all visible devices
        for (int d = 0; d < avail; d++) m->devices.push_back(d);
    } else {
        m->devices.assign(devices, devices + ndevices);
    }
    for (int d : m->devices) {
        tfhe_b200_ctx *c = nullptr;
        if (tfhe_b200_ctx_create(&c, p, d)) {
            mfail("device %d: %s", d, tfhe_b200_last_error());
            for (tfhe_b200_ctx *x : m->ctx) tfhe_b200_ctx_destroy(x);
            delete m;
            return 1;
        }
        m->ctx.push_back(c);
    }
    // peer access between every pair of distinct devices (ignored where the topology has none:
    // cudaMemcpyPeerAsync then stages through the host)
    for (int a : m->devices)
        for (int b : m->devices) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError();  // already enabled
            }
        }
    *out = m;
    return 0;
}

void tfhe_b200_multi_destroy(tfhe_b200_multi *m) {
    if (!m) return;
    for (tfhe_b200_ctx *c : m->ctx) tfhe_b200_ctx_destroy(c);
    delete m;
}

int tfhe_b200_multi_devices(const tfhe_b200_multi *m) { return m ? (int) m->ctx.size() : 0; }

tfhe_b200_ctx *tfhe_b200_multi_ctx(tfhe_b200_multi *m, int index) {
    return (m && index >= 0 && index < (int) m->ctx.size()) ? m->ctx[index] : nullptr;
}

// Keys: host -> first device once, device -> device for the others, conversion on every GPU.
int tfhe_b200_multi_load_keys(tfhe_b200_multi *m, const int32_t *bk_coef, const int32_t *ks) {
    if (!m || !bk_coef || !ks) return mfail("null argument");
    const size_t bk_bytes = tfhe_b200_bk_words(&m->p) * sizeof(int32_t), ks_bytes = tfhe_b200_ks_words(&m->p) * sizeof(int32_t);
    const int nd = (int) m->ctx.size();
    std::vector<int32_t *> d_bk(nd, nullptr), d_ks(nd, nullptr);
    std::vector<cudaStream_t> st(nd, nullptr);
    int rc = 0;
    auto cleanup = [&]() {
        for (int d = 0; d < nd; d++) {
            cudaSetDevice(m->devices[d]);
            if (st[d]) {
                cudaStreamSynchronize(st[d]);
                cudaStreamDestroy(st[d]);
            }
            if (d_bk[d]) cudaFree(d_bk[d]);
            if (d_ks[d]) cudaFree(d_ks[d]);
        }
    };
    for (int d = 0; d < nd && !rc; d++) {
        if (cudaSetDevice(m->devices[d]) != cudaSuccess || cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&d_bk[d], bk_bytes) != cudaSuccess || cudaMalloc(&d_ks[d], ks_bytes) != cudaSuccess)
            rc = mfail("device %d: staging allocation failed: %s", m->devices[d], cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) {
        cudaSetDevice(m->devices[0]);
        cudaEvent_t up;
        if (cudaMemcpyAsync(d_bk[0], bk_coef, bk_bytes, cudaMemcpyHostToDevice, st[0]) != cudaSuccess ||
            cudaMemcpyAsync(d_ks[0], ks, ks_bytes, cudaMemcpyHostToDevice, st[0]) != cudaSuccess ||
            cudaEventCreateWithFlags(&up, cudaEventDisableTiming) != cudaSuccess) {
            rc = mfail("key upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            cudaEventRecord(up, st[0]);
            for (int d = 1; d < nd && !rc; d++) {
                cudaSetDevice(m->devices[d]);
                cudaStreamWaitEvent(st[d], up, 0);
                if (cudaMemcpyPeerAsync(d_bk[d], m->devices[d], d_bk[0], m->devices[0], bk_bytes, st[d]) != cudaSuccess ||
                    cudaMemcpyPeerAsync(d_ks[d], m->devices[d], d_ks[0], m->devices[0], ks_bytes, st[d]) != cudaSuccess)
                    rc = mfail("device %d: peer copy of the keys failed: %s", m->devices[d], cudaGetErrorString(cudaGetLastError()));
            }
            for (int d = 0; d < nd && !rc; d++) {
                if (tfhe_b200_load_keys_device(m->ctx[d], d_bk[d], d_ks[d], st[d]))
                    rc = mfail("device %d: %s", m->devices[d], tfhe_b200_last_error());
            }
            for (int d = 0; d < nd; d++) {
                cudaSetDevice(m->devices[d]);
                if (cudaStreamSynchronize(st[d]) != cudaSuccess && !rc)
                    rc = mfail("device %d: key conversion failed: %s", m->devices[d], cudaGetErrorString(cudaGetLastError()));
            }
            cudaEventDestroy(up);
        }
    }
    cleanup();
    return rc;
}

// out[g] = gate(ca[g], cb[g]) for a HOST batch, sharded contiguously over the devices.
int tfhe_b200_multi_gate_host(tfhe_b200_multi *m, int gate, int32_t *out, const int32_t *ca, const int32_t *cb,
                              long long count) {
    if (!m) return mfail("null argument");
    if (count < 0) return mfail("negative count");
    if (count == 0) return 0;
    const size_t row = (size_t) m->p.n + 1;
    const int nd = (int) m->ctx.size();
    return for_each_device(m, [&](int d) {
        long long lo, hi;
        shard(count, nd, d, &lo, &hi);
        if (hi == lo) return 0;
        return tfhe_b200_gate_host(m->ctx[d], gate, out + lo * row, ca + lo * row, cb + lo * row, (int) (hi - lo));
    });
}

int tfhe_b200_multi_mux_host(tfhe_b200_multi *m, int32_t *out, const int32_t *a, const int32_t *b, const int32_t *c,
                             long long count) {
    if (!m) return mfail("null argument");
    if (count < 0) return mfail("negative count");
    if (count == 0) return 0;
    const size_t row = (size_t) m->p.n + 1;
    const int nd = (int) m->ctx.size();
    return for_each_device(m, [&](int d) {
        long long lo, hi;
        shard(count, nd, d, &lo, &hi);
        if (hi == lo) return 0;
        return tfhe_b200_mux_host(m->ctx[d], out + lo * row, a + lo * row, b + lo * row, c + lo * row, (int) (hi - lo));
    });
}

unsigned long long tfhe_b200_multi_launch_count(const tfhe_b200_multi *m) {
    unsigned long long n = 0;
    if (m)
        for (tfhe_b200_ctx *c : m->ctx) n += tfhe_b200_launch_count(c);
    return n;
}

}  // extern "C"
