// vf_tables.cpp — device-wide, reference-counted cache of the 2^24-entry function tables (baked
// colorlut tables, hsvfilter / hsvdetector / chain function tables).
//
// A table is a pure function of its key (LUT content + interpolation, or element + layout + every
// setting bit), so every context on a device that needs the same function shares one 64 MiB copy:
// two `hsvfilter` instances with equal settings, or N `colorlut` instances on one LUT file, cost one
// table of L2 footprint instead of N.  Builds are stream-ordered on the first user's stream; other
// contexts wait for the build's event on their own stream.
#include <list>
#include <mutex>

#include "vf_internal.h"

namespace vf {
namespace {

// recursive: building the chain's table runs the colorlut stage, which acquires the baked table
std::recursive_mutex g_mu;
std::list<SharedTable> g_tables;

}  // namespace

SharedTable *table_acquire(int device, const std::vector<uint8_t> &key) {
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    for (SharedTable &t : g_tables)
        if (t.device == device && t.key == key) {
            t.refs++;
            return &t;
        }
    uint32_t *data = nullptr;
    if (cudaMalloc((void **)&data, sizeof(uint32_t) << 24) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(data);
        return nullptr;
    }
    g_tables.emplace_back();
    SharedTable &t = g_tables.back();
    t.key = key;
    t.device = device;
    t.data = data;
    t.ready = ev;
    t.refs = 1;
    return &t;
}

void table_release(SharedTable *t) {
    if (!t) return;
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    if (--t->refs > 0) return;
    // cudaFree waits for all work on the device, so kernels of the releasing context that still
    // read the table finish first
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != t->device) cudaSetDevice(t->device);
    cudaFree(t->data);
    cudaEventDestroy(t->ready);
    if (cur >= 0 && cur != t->device) cudaSetDevice(cur);
    cudaGetLastError();
    for (auto it = g_tables.begin(); it != g_tables.end(); ++it)
        if (&*it == t) {
            g_tables.erase(it);
            break;
        }
}

cudaError_t table_ensure_built(SharedTable *t, cudaStream_t stream,
                               const std::function<cudaError_t(uint32_t *)> &build) {
    std::lock_guard<std::recursive_mutex> lock(g_mu);  // one builder; enqueueing is quick
    if (t->built) {
        // built (or being built) on another context's stream: order our stream after it
        return t->builder_stream == stream ? cudaSuccess : cudaStreamWaitEvent(stream, t->ready, 0);
    }
    cudaError_t e = build(t->data);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(t->ready, stream);
    if (e != cudaSuccess) return e;
    t->built = true;
    t->builder_stream = stream;
    return cudaSuccess;
}

void table_cache_stats(int device, uint64_t *tables, uint64_t *bytes) {
    std::lock_guard<std::recursive_mutex> lock(g_mu);
    uint64_t n = 0;
    for (const SharedTable &t : g_tables)
        if (t.device == device) n++;
    if (tables) *tables = n;
    if (bytes) *bytes = n * (sizeof(uint32_t) << 24);
}

}  // namespace vf
