// cuda_memory_pipeline.cpp — the device-resident pipeline of SURVEY.md §8f rank 3, driven in
// the order GStreamer drives a BaseTransform chain:
//
//   filesrc ! cudaupload ! cudacolorlut location=<.cube> ! cudahsvfilter hue-shift=… ! cudadownload ! filesink
//
// caps negotiation (memory:CUDAMemory) → allocation queries (the uploader takes the pool
// cudacolorlut proposes; cudacolorlut decides on the pool cudahsvfilter proposes) →
// per buffer: before_transform (device follow) → transform (enqueue only) → buffers go back to
// their pools with the stream that last touched them.  Frames stay in HBM between elements.
//
// Usage: cuda_memory_pipeline <lut.cube> <in.raw> <out.raw> <width> <height> <frames>
//                             [hue_shift] [src_device]
// in.raw holds `frames` tightly packed RGBA frames; out.raw receives the results.  With
// src_device given, the "uploader" ignores the proposed pool and allocates on that device, so
// the elements (created on device 0) must follow it.  Prints one JSON line.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../gst-plugins-rs_b200/elements/vf_elements.hpp"

using namespace b200vf;

static int die(const char *what, const std::string &why) {
    std::fprintf(stderr, "%s: %s\n", what, why.c_str());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 7) {
        std::fprintf(stderr,
                     "usage: %s <lut.cube> <in.raw> <out.raw> <width> <height> <frames> "
                     "[hue_shift] [src_device]\n",
                     argv[0]);
        return 2;
    }
    const uint32_t w = (uint32_t)std::atoi(argv[4]), h = (uint32_t)std::atoi(argv[5]);
    const unsigned n = (unsigned)std::atoi(argv[6]);
    const float hue_shift = argc > 7 ? (float)std::atof(argv[7]) : 0.0f;
    const int src_device = argc > 8 ? std::atoi(argv[8]) : -1;
    const size_t frame_bytes = (size_t)w * h * 4;

    std::vector<uint8_t> input(frame_bytes * n), output(frame_bytes * n);
    FILE *f = std::fopen(argv[2], "rb");
    if (!f || std::fread(input.data(), 1, input.size(), f) != input.size())
        return die("read", argv[2]);
    std::fclose(f);

    // NULL → READY → PAUSED
    auto lut = element_factory_make("cudacolorlut"), hsv = element_factory_make("cudahsvfilter");
    lut->set_property("location", Value{std::string(argv[1])});
    hsv->set_property("hue-shift", Value{hue_shift});
    ErrorMessage e = lut->start();
    if (!e.ok()) return die("cudacolorlut start", e.message);
    e = hsv->start();
    if (!e.ok()) return die("cudahsvfilter start", e.message);

    // caps: video/x-raw(memory:CUDAMemory), format=RGBA, fixed geometry
    Caps caps;
    caps.formats = {"RGBA"};
    caps.features = {kCapsFeatureCudaMemory};
    caps.width = w, caps.height = h;
    for (VideoFilter *el : {lut.get(), hsv.get()}) {
        Caps other = el->transform_caps(PadDirection::Sink, caps, nullptr);
        if (other.formats != caps.formats || !other.has_feature(kCapsFeatureCudaMemory))
            return die("transform_caps", "CUDA-memory RGBA caps were not accepted");
        std::string err = el->set_caps(caps, other);
        if (!err.empty()) return die("set_caps", err);
    }

    // allocation, upstream side of cudacolorlut: the uploader asks it for a pool
    AllocationQuery upstream_q;
    upstream_q.caps = caps;
    std::string err = lut->propose_allocation(upstream_q);
    if (!err.empty() || upstream_q.pools.empty()) return die("propose_allocation", err);
    std::shared_ptr<DeviceBufferPool> in_pool = upstream_q.pools.front().pool;
    if (src_device >= 0) {  // an upstream that brings its own memory from another device
        in_pool = DeviceBufferPool::create(src_device, caps, 0, 0, &err);
        if (!in_pool) return die("upstream pool", err);
    }
    // the "uploader" element's own context, on the device its buffers live on
    b200vf_ctx *up_ctx = nullptr;
    if (b200vf_ctx_create(in_pool->device(), &up_ctx) != B200VF_OK)
        return die("uploader context", b200vf_last_error(nullptr));

    // allocation, downstream side: cudacolorlut's decide_allocation over what cudahsvfilter
    // proposes; redone whenever an element reports reconfigure (device change)
    std::shared_ptr<DeviceBufferPool> out_pool;
    unsigned negotiations = 0;
    auto negotiate_downstream = [&]() -> std::string {
        AllocationQuery q;
        q.caps = caps;
        std::string r = hsv->propose_allocation(q);
        if (!r.empty()) return r;
        r = lut->decide_allocation(q);
        if (!r.empty()) return r;
        out_pool = q.pools.front().pool;
        negotiations++;
        return {};
    };
    err = negotiate_downstream();
    if (!err.empty()) return die("decide_allocation", err);

    for (unsigned i = 0; i < n; i++) {
        // cudaupload
        VideoFrameRef in, out;
        if (!in_pool->acquire(in)) return die("acquire", b200vf_last_error(nullptr));
        if (b200vf_memcpy(up_ctx, in.data, input.data() + i * frame_bytes, frame_bytes, 0) != B200VF_OK)
            return die("upload", b200vf_last_error(up_ctx));

        // cudacolorlut
        lut->before_transform(in);
        if (lut->take_reconfigure()) {
            hsv->before_transform(in);  // same memory type / device question for the next element
            hsv->take_reconfigure();
            err = negotiate_downstream();
            if (!err.empty()) return die("renegotiation", err);
        }
        if (!out_pool->acquire(out)) return die("acquire", b200vf_last_error(nullptr));
        if (lut->transform_frame(in, out) != FlowReturn::Ok) return die("cudacolorlut", lut->last_error());
        in_pool->release(in, b200vf_ctx_get_stream(lut->context()));  // still being read

        // cudahsvfilter (in place on the buffer cudacolorlut produced)
        hsv->before_transform(out);
        hsv->take_reconfigure();
        // every element keeps its own stream; the hand-over of `out` is ordered on the device
        // (no host sync, no shared handle that could dangle when an element restarts)
        if (b200vf_ctx_wait_for(hsv->context(), lut->context()) != B200VF_OK)
            return die("wait_for", b200vf_last_error(hsv->context()));
        if (hsv->transform_frame_ip(out) != FlowReturn::Ok) return die("cudahsvfilter", hsv->last_error());

        // cudadownload: stream-ordered after both kernels, synchronous for the host
        if (b200vf_memcpy(hsv->context(), output.data() + i * frame_bytes, out.data, frame_bytes, 1) !=
            B200VF_OK)
            return die("download", b200vf_last_error(hsv->context()));
        out_pool->release(out, nullptr);
    }

    f = std::fopen(argv[3], "wb");
    if (!f || std::fwrite(output.data(), 1, output.size(), f) != output.size())
        return die("write", argv[3]);
    std::fclose(f);

    std::printf(
        "{\"frames\": %u, \"device\": %d, \"in_pool_allocated\": %u, \"out_pool_allocated\": %u, "
        "\"in_pool_outstanding\": %u, \"out_pool_outstanding\": %u, \"downstream_negotiations\": %u}\n",
        n, lut->device(), in_pool->allocated(), out_pool->allocated(), in_pool->outstanding(),
        out_pool->outstanding(), negotiations);
    b200vf_ctx_destroy(up_ctx);
    hsv->stop();
    lut->stop();
    return 0;
}
