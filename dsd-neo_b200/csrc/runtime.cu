// SPDX-License-Identifier: GPL-3.0-or-later
// Library/device plumbing of the C-ABI (include/dsdneo_b200.h, "library / device" section).
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace dsdneo {

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launch_count{0};

void
set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int
cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return (e == cudaErrorMemoryAllocation) ? DSDNEO_B200_ENOMEM : DSDNEO_B200_ECUDA;
}

int
ensure_device() {
    static thread_local int checked_dev = -1;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device available (%s); libdsdneo_b200 has no CPU fallback", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return DSDNEO_B200_ENODEV;
    }
    if (dev == checked_dev) {
        return 0;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        set_error("cannot query CUDA device %d (%s)", dev, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return DSDNEO_B200_ENODEV;
    }
    if (major != 10) {
        set_error("CUDA device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return DSDNEO_B200_ENODEV;
    }
    checked_dev = dev;
    return 0;
}

bool g_timing_on = false;

namespace {
struct TimingRec {
    const char* name;
    cudaEvent_t a, b;
};
struct TimingAcc {
    const char* name;
    unsigned long long launches;
    double ms;
};
constexpr int kMaxRec = 4096, kMaxAcc = 64;
TimingRec g_rec[kMaxRec];
int g_n_rec = 0;
TimingAcc g_acc[kMaxAcc];
int g_n_acc = 0;
cudaEvent_t g_ev_pool[2 * kMaxRec];
int g_ev_pool_n = 0;
int g_open = -1;
/* The recorder is process-wide (one table, one event pool, created on the device current at the first timed launch): a
 * bench facility for ONE device.  The mutex makes it safe against launches from several host threads: it is held from
 * timing_begin to timing_end, i.e. across the launch the pair brackets (only while timing is enabled). */
std::recursive_mutex g_timing_mu; /* recursive: a nested bracket (none today) must not deadlock */

cudaEvent_t
pool_event(int idx) {
    while (g_ev_pool_n <= idx) {
        cudaEventCreate(&g_ev_pool[g_ev_pool_n++]);
    }
    return g_ev_pool[idx];
}

void
timing_drain() {
    for (int i = 0; i < g_n_rec; i++) {
        float ms = 0.0f;
        if (cudaEventSynchronize(g_rec[i].b) != cudaSuccess || cudaEventElapsedTime(&ms, g_rec[i].a, g_rec[i].b) != cudaSuccess) {
            (void)cudaGetLastError();
            continue;
        }
        int j = 0;
        for (; j < g_n_acc; j++) {
            if (g_acc[j].name == g_rec[i].name || strcmp(g_acc[j].name, g_rec[i].name) == 0) {
                break;
            }
        }
        if (j == g_n_acc) {
            if (g_n_acc == kMaxAcc) {
                continue;
            }
            g_acc[g_n_acc++] = TimingAcc{g_rec[i].name, 0ull, 0.0};
        }
        g_acc[j].launches++;
        g_acc[j].ms += (double)ms;
    }
    g_n_rec = 0;
}
}  // namespace

void
timing_begin(const char* name, cudaStream_t s) {
    g_timing_mu.lock();
    if (g_n_rec == kMaxRec) {
        timing_drain();
    }
    g_open = g_n_rec++;
    g_rec[g_open].name = name;
    g_rec[g_open].a = pool_event(2 * g_open);
    g_rec[g_open].b = pool_event(2 * g_open + 1);
    cudaEventRecord(g_rec[g_open].a, s);
}

void
timing_end(cudaStream_t s) {
    if (g_open >= 0) {
        cudaEventRecord(g_rec[g_open].b, s);
        g_open = -1;
    }
    g_timing_mu.unlock();
}

}  // namespace dsdneo

using namespace dsdneo;

extern "C" {

int
dsdneo_b200_abi_version(void) {
    return DSDNEO_B200_ABI_VERSION;
}

const char*
dsdneo_b200_last_error(void) {
    return t_err;
}

int
dsdneo_b200_init(int device_ordinal) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); libdsdneo_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return DSDNEO_B200_ENODEV;
    }
    if (device_ordinal < 0 || device_ordinal >= n) {
        set_error("device ordinal %d out of range [0,%d)", device_ordinal, n);
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaSetDevice(device_ordinal));
    return ensure_device();
}

int
dsdneo_b200_device_sm_count(void) {
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    int dev = 0, sms = 0;
    DSDNEO_CUDA(cudaGetDevice(&dev));
    DSDNEO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return sms;
}

int
dsdneo_b200_stream_sync(void* stream) {
    DSDNEO_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

unsigned long long
dsdneo_b200_launch_count(void) {
    return g_launch_count.load(std::memory_order_relaxed);
}

int
dsdneo_b200_timing_enable(int on) {
    std::lock_guard<std::recursive_mutex> lk(g_timing_mu);
    timing_drain();
    g_n_acc = 0;
    g_timing_on = on != 0;
    return 0;
}

int
dsdneo_b200_timing_report(char* buf, size_t cap) {
    if (!buf || cap < 3) {
        set_error("timing_report: bad buffer");
        return DSDNEO_B200_EINVAL;
    }
    std::lock_guard<std::recursive_mutex> lk(g_timing_mu);
    timing_drain();
    size_t off = 0;
    off += (size_t)snprintf(buf + off, cap - off, "{");
    for (int j = 0; j < g_n_acc && off + 160 < cap; j++) {
        off += (size_t)snprintf(buf + off, cap - off, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f}", j ? ", " : "",
                                g_acc[j].name, g_acc[j].launches, g_acc[j].ms);
    }
    snprintf(buf + off, cap - off, "}");
    return g_n_acc;
}

void*
dsdneo_b200_malloc_device(size_t bytes) {
    if (ensure_device()) {
        return NULL;
    }
    void* p = NULL;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
        return NULL;
    }
    return p;
}

void
dsdneo_b200_free_device(void* d_ptr) {
    if (d_ptr) {
        cudaFree(d_ptr);
    }
}

void*
dsdneo_b200_malloc_pinned(size_t bytes) {
    if (ensure_device()) {
        return NULL;
    }
    void* p = NULL;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
        return NULL;
    }
    return p;
}

void
dsdneo_b200_free_pinned(void* h_ptr) {
    if (h_ptr) {
        cudaFreeHost(h_ptr);
    }
}

int
dsdneo_b200_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    DSDNEO_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

int
dsdneo_b200_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    DSDNEO_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

} /* extern "C" */
