// vf_pool.cpp — device frame pool and pointer classification: the pieces an element needs to
// negotiate `memory:CUDAMemory` buffers.  Behavioural model: the
// D3D12 buffer pool d3d12colorlut proposes / decides on (d3d12colorlut/imp.rs:385-492) and
// its device-follow check (:494-542).  No pixel work here.
#include <cuda_runtime.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200vf.h"
#include "vf_internal.h"

namespace {

struct PoolBuffer {
    void *data = nullptr;
    cudaEvent_t last_use = nullptr;  // recorded at release when a stream still touches the frame
    bool pending = false;
};

int cuda_error(cudaError_t e, const char *what) {
    cudaGetLastError();
    return vf::fail_global(B200VF_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" +
                                                cudaGetErrorString(e) + ")");
}

}  // namespace

struct b200vf_pool {
    int device = 0;
    b200vf_pool_config cfg{};
    int64_t stride = 0;
    size_t frame_bytes = 0;
    std::mutex mu;
    std::condition_variable returned;
    std::vector<PoolBuffer> idle;                     // LIFO: the most recently used frame first
    std::unordered_map<void *, PoolBuffer> outstanding;
    uint32_t allocated = 0;
};

namespace {

// Caller holds pool->mu (or owns the pool exclusively).
int allocate_one(b200vf_pool *pool, PoolBuffer &b) {
    cudaError_t e = cudaSetDevice(pool->device);
    if (e != cudaSuccess) return cuda_error(e, "cudaSetDevice");
    e = pool->cfg.host_pinned ? cudaMallocHost(&b.data, pool->frame_bytes)
                              : cudaMalloc(&b.data, pool->frame_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return vf::fail_global(B200VF_ERR_NOMEM, std::string("pool: allocation failed: ") + cudaGetErrorString(e));
    }
    e = cudaEventCreateWithFlags(&b.last_use, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (pool->cfg.host_pinned)
            cudaFreeHost(b.data);
        else
            cudaFree(b.data);
        b.data = nullptr;
        return cuda_error(e, "cudaEventCreate");
    }
    pool->allocated++;
    return B200VF_OK;
}

void free_one(PoolBuffer &b, bool host_pinned) {
    if (b.last_use) cudaEventDestroy(b.last_use);
    if (b.data) {
        if (host_pinned)
            cudaFreeHost(b.data);
        else
            cudaFree(b.data);
    }
    b = PoolBuffer{};
}

}  // namespace

extern "C" {

int b200vf_pool_create(int device, const b200vf_pool_config *config, b200vf_pool **out) {
    if (!out) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_create: out is NULL");
    *out = nullptr;
    if (!config) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_create: config is NULL");
    const uint32_t bpp = b200vf_format_bytes_per_pixel(config->format);
    if (bpp == 0) return vf::fail_global(B200VF_ERR_UNSUPPORTED_FORMAT, "pool_create: unknown format");
    if (config->width == 0 || config->height == 0)
        return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_create: empty frame geometry");
    if (config->max_buffers != 0 && config->max_buffers < config->min_buffers)
        return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_create: max_buffers < min_buffers");
    int n = 0;
    int rc = b200vf_device_count(&n);
    if (rc) return rc;
    if (device < 0 || device >= n)
        return vf::fail_global(B200VF_ERR_NO_DEVICE, "pool_create: device index out of range");
    try {
        std::unique_ptr<b200vf_pool> pool(new b200vf_pool());
        pool->device = device;
        pool->cfg = *config;
        const uint64_t row = (uint64_t)config->width * bpp;
        pool->stride = (int64_t)((row % 16 == 0) ? row : ((row + 255) / 256) * 256);
        pool->frame_bytes = (size_t)pool->stride * config->height;
        pool->idle.reserve(config->min_buffers);
        for (uint32_t i = 0; i < config->min_buffers; i++) {  // nobody else sees the pool yet
            PoolBuffer b;
            rc = allocate_one(pool.get(), b);
            if (rc) {
                for (PoolBuffer &x : pool->idle) free_one(x, config->host_pinned != 0);
                return rc;
            }
            pool->idle.push_back(b);
        }
        *out = pool.release();
        return B200VF_OK;
    } catch (...) {
        return vf::fail_global(B200VF_ERR_NOMEM, "pool_create: host allocation failed");
    }
}

void b200vf_pool_destroy(b200vf_pool *pool) {
    if (!pool) return;
    cudaSetDevice(pool->device);
    cudaDeviceSynchronize();  // frames may still be read or written by enqueued work
    {
        std::lock_guard<std::mutex> g(pool->mu);
        for (PoolBuffer &b : pool->idle) free_one(b, pool->cfg.host_pinned != 0);
        for (auto &kv : pool->outstanding) free_one(kv.second, pool->cfg.host_pinned != 0);
        pool->idle.clear();
        pool->outstanding.clear();
    }
    cudaGetLastError();
    delete pool;
}

int b200vf_pool_acquire(b200vf_pool *pool, uint32_t flags, b200vf_frame *out) {
    if (!pool || !out) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_acquire: NULL argument");
    PoolBuffer b;
    try {
        std::unique_lock<std::mutex> lk(pool->mu);
        for (;;) {
            if (!pool->idle.empty()) {
                b = pool->idle.back();
                pool->idle.pop_back();
                break;
            }
            if (pool->cfg.max_buffers == 0 || pool->allocated < pool->cfg.max_buffers) {
                int rc = allocate_one(pool, b);
                if (rc) return rc;
                break;
            }
            if (flags & B200VF_POOL_DONTWAIT)
                return vf::fail_global(B200VF_ERR_NOMEM, "pool_acquire: all buffers are in use");
            pool->returned.wait(lk);
        }
        pool->outstanding.emplace(b.data, b);
    } catch (...) {
        return vf::fail_global(B200VF_ERR_NOMEM, "pool_acquire: host allocation failed");
    }
    if (b.pending) {  // the previous user's enqueued work must be done before a new owner writes
        cudaError_t e = cudaEventSynchronize(b.last_use);
        if (e != cudaSuccess) {
            // the frame is not handed out: put it back so that it can be released / reused and a
            // bounded pool does not lose a buffer for good
            cudaGetLastError();
            {
                std::lock_guard<std::mutex> g(pool->mu);
                pool->outstanding.erase(b.data);
                try {
                    pool->idle.push_back(b);
                } catch (...) {
                    free_one(b, pool->cfg.host_pinned != 0);
                    pool->allocated--;
                }
            }
            pool->returned.notify_one();
            return cuda_error(e, "cudaEventSynchronize");
        }
        std::lock_guard<std::mutex> g(pool->mu);
        auto it = pool->outstanding.find(b.data);
        if (it != pool->outstanding.end()) it->second.pending = false;
    }
    out->data = b.data;
    out->stride = pool->stride;
    out->width = pool->cfg.width;
    out->height = pool->cfg.height;
    out->format = pool->cfg.format;
    out->memory = pool->cfg.host_pinned ? B200VF_MEM_HOST : B200VF_MEM_DEVICE;
    return B200VF_OK;
}

int b200vf_pool_release(b200vf_pool *pool, const b200vf_frame *frame, void *last_use_stream) {
    if (!pool || !frame) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_release: NULL argument");
    std::unique_lock<std::mutex> lk(pool->mu);
    auto it = pool->outstanding.find(frame->data);
    if (it == pool->outstanding.end())
        return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_release: frame does not belong to this pool");
    PoolBuffer b = it->second;
    pool->outstanding.erase(it);
    if (last_use_stream) {
        cudaError_t e = cudaSetDevice(pool->device);
        if (e == cudaSuccess) e = cudaEventRecord(b.last_use, (cudaStream_t)last_use_stream);
        if (e != cudaSuccess) {  // keep the buffer, but make the next owner wait for the device
            cudaGetLastError();
            cudaDeviceSynchronize();
            b.pending = false;
        } else {
            b.pending = true;
        }
    }
    try {
        pool->idle.push_back(b);
    } catch (...) {
        pool->allocated--;
        lk.unlock();
        free_one(b, pool->cfg.host_pinned != 0);
        return vf::fail_global(B200VF_ERR_NOMEM, "pool_release: host allocation failed");
    }
    lk.unlock();
    pool->returned.notify_one();
    return B200VF_OK;
}

int b200vf_pool_release_after(b200vf_pool *pool, const b200vf_frame *frame, const b200vf_ctx *last_user) {
    if (!last_user) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_release_after: context is NULL");
    // the context's stream whatever it is — the legacy default stream (a NULL handle) included,
    // which b200vf_pool_release cannot tell from "no pending work"
    void *s = b200vf_ctx_get_stream(last_user);
    if (s) return b200vf_pool_release(pool, frame, s);
    return b200vf_pool_release(pool, frame, (void *)cudaStreamLegacy);
}

int b200vf_pool_get_stats(b200vf_pool *pool, b200vf_pool_stats *out) {
    if (!pool || !out) return vf::fail_global(B200VF_ERR_INVALID_ARG, "pool_get_stats: NULL argument");
    std::lock_guard<std::mutex> g(pool->mu);
    out->allocated = pool->allocated;
    out->outstanding = (uint32_t)pool->outstanding.size();
    out->frame_bytes = pool->frame_bytes;
    out->stride = pool->stride;
    return B200VF_OK;
}

int b200vf_pool_device(const b200vf_pool *pool) { return pool ? pool->device : -1; }

int b200vf_pointer_info(const void *p, uint32_t *memory, int *device) {
    if (!p || !memory || !device)
        return vf::fail_global(B200VF_ERR_INVALID_ARG, "pointer_info: NULL argument");
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) return cuda_error(e, "cudaPointerGetAttributes");
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
        *memory = B200VF_MEM_DEVICE;
        *device = attr.device;
    } else {
        *memory = B200VF_MEM_HOST;
        *device = -1;
    }
    return B200VF_OK;
}

}  // extern "C"
