// negotiation_check.cpp — checks of the CUDA-memory element variants' negotiation logic against
// the behaviour of d3d12colorlut (d3d12colorlut/imp.rs:236-266 caps, :349-383 set_caps,
// :385-492 allocation, :494-542 before_transform).  Mode "cpu": everything that must hold
// without a device.  Mode "gpu <lut.cube>": pools and negotiation on cuda:0.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../gst-plugins-rs_b200/elements/vf_elements.hpp"

using namespace b200vf;

static int g_failed = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            g_failed++;                                                 \
        }                                                               \
    } while (0)

static Caps cuda_caps(const char *format, uint32_t w, uint32_t h) {
    Caps c;
    c.formats = {format};
    c.features = {kCapsFeatureCudaMemory};
    c.width = w, c.height = h;
    return c;
}

static void check_cpu() {
    const char *names[3] = {"cudacolorlut", "cudahsvfilter", "cudahsvdetector"};
    const char *bases[3] = {"colorlut", "hsvfilter", "hsvdetector"};
    for (int i = 0; i < 3; i++) {
        auto e = element_factory_make(names[i]);
        auto b = element_factory_make(bases[i]);
        CHECK(e && b);
        CHECK(std::string(e->factory_name()) == names[i]);
        CHECK(e->mode() == b->mode());
        // same formats and properties as the system-memory element, plus the memory feature
        for (size_t t = 0; t < 2; t++) {
            CHECK(e->pad_templates()[t].formats == b->pad_templates()[t].formats);
            CHECK(e->pad_templates()[t].features ==
                  std::vector<std::string>{kCapsFeatureCudaMemory});
            CHECK(b->pad_templates()[t].features.empty());
        }
        CHECK(e->properties().size() == b->properties().size());
        for (size_t p = 0; p < e->properties().size(); p++)
            CHECK(e->properties()[p].name == b->properties()[p].name);
        // not started: no device, so no allocation can be negotiated
        AllocationQuery q;
        q.caps = cuda_caps("RGBA", 64, 32);
        CHECK(e->propose_allocation(q) == "Device not configured");
        CHECK(e->decide_allocation(q) == "Device not configured");
        CHECK(q.pools.empty());
        VideoFrameRef none;
        e->before_transform(none);
        CHECK(e->last_error() == "No device configured");
        // the system-memory elements neither offer nor need a pool
        CHECK(b->propose_allocation(q).empty() && q.pools.empty());
    }
    // caps features survive transform_caps, with and without a filter
    auto lut = element_factory_make("cudacolorlut");
    Caps in = cuda_caps("RGBA", 64, 32), filter;
    filter.formats = {"RGBA64_LE", "RGBA"};
    Caps out = lut->transform_caps(PadDirection::Sink, in, &filter);
    CHECK(out.formats == std::vector<std::string>{"RGBA"});
    CHECK(out.has_feature(kCapsFeatureCudaMemory) && out.width == 64 && out.height == 32);
    auto det = element_factory_make("cudahsvdetector");
    Caps din = cuda_caps("BGRx", 64, 32);
    Caps dout = det->transform_caps(PadDirection::Sink, din, nullptr);
    CHECK(dout.formats == (std::vector<std::string>{"RGBA", "ARGB", "BGRA", "ABGR"}));
    CHECK(dout.has_feature(kCapsFeatureCudaMemory) && dout.width == 64);
    // set_caps on a cudacolorlut that has no LUT (d3d12colorlut/imp.rs:360-362)
    CHECK(lut->set_caps(in, in) == "No LUT configured");
    // host frames are refused by the CUDA-memory variants
    uint8_t px[16] = {0};
    VideoFrameRef host{px, 16, 4, 1, "RGBA", B200VF_MEM_HOST};
    auto hsv = element_factory_make("cudahsvfilter");
    CHECK(hsv->transform_frame_ip(host) == FlowReturn::Error);
    CHECK(hsv->last_error() == "Wrong memory type");
    // pool without fixed caps / without a device
    std::string err;
    Caps loose;
    loose.formats = {"RGBA", "RGBx"};
    CHECK(!DeviceBufferPool::create(0, loose, 0, 0, &err) && err == "caps are not fixed");
    // registration: the variants live in their own plugin descriptor
    bool found = false;
    for (const PluginDescriptor &p : plugins())
        if (p.name == "b200vf") found = p.elements.size() == 3;
    CHECK(found);
}

static void check_gpu(const char *cube) {
    auto lut = element_factory_make("cudacolorlut");
    lut->set_property("location", Value{std::string(cube)});
    CHECK(lut->start().ok());
    Caps caps = cuda_caps("RGBA", 320, 200);
    Caps loose = caps;
    loose.width = 0;
    CHECK(lut->set_caps(loose, loose) == "Failed to parse output caps");
    CHECK(lut->set_caps(caps, caps).empty());

    // propose: one pool on our device, sized stride*height, video meta advertised
    AllocationQuery up;
    up.caps = caps;
    CHECK(lut->propose_allocation(up).empty());
    CHECK(up.pools.size() == 1 && up.video_meta);
    CHECK(up.pools[0].pool->device() == lut->device());
    CHECK(up.pools[0].size == 320u * 4 * 200);
    AllocationQuery no_pool;
    no_pool.caps = caps;
    no_pool.need_pool = false;
    CHECK(lut->propose_allocation(no_pool).empty() && no_pool.pools.empty() && no_pool.video_meta);

    // decide: a suitable downstream pool is kept (update), a foreign geometry is replaced,
    // an empty query gets a new pool (add)
    auto hsv = element_factory_make("cudahsvfilter");
    CHECK(hsv->start().ok());
    AllocationQuery down;
    down.caps = caps;
    CHECK(hsv->propose_allocation(down).empty());
    auto offered = down.pools[0].pool;
    CHECK(lut->decide_allocation(down).empty());
    CHECK(down.pools.size() == 1 && down.pools[0].pool == offered);
    AllocationQuery wrong;
    wrong.caps = caps;
    std::string err;
    wrong.pools.push_back({DeviceBufferPool::create(0, cuda_caps("RGBA", 64, 64), 0, 0, &err), 64 * 64 * 4, 2, 4});
    CHECK(wrong.pools[0].pool != nullptr);
    auto foreign = wrong.pools[0].pool;
    CHECK(lut->decide_allocation(wrong).empty());
    CHECK(wrong.pools.size() == 1 && wrong.pools[0].pool != foreign);
    CHECK(wrong.pools[0].min_buffers == 2 && wrong.pools[0].max_buffers == 4);
    CHECK(wrong.pools[0].size == 320u * 4 * 200);
    CHECK(wrong.pools[0].pool->allocated() == 2);  // min_buffers preallocated
    AllocationQuery empty;
    empty.caps = caps;
    CHECK(lut->decide_allocation(empty).empty() && empty.pools.size() == 1);

    // before_transform: device memory of our own device is accepted silently, host memory
    // is the reference's "Wrong memory type"
    VideoFrameRef frame;
    CHECK(offered->acquire(frame));
    lut->before_transform(frame);
    CHECK(lut->last_error().empty() && !lut->take_reconfigure());
    uint8_t px[16] = {0};
    VideoFrameRef host{px, 16, 4, 1, "RGBA", B200VF_MEM_HOST};
    lut->before_transform(host);
    CHECK(lut->last_error() == "Wrong memory type");
    CHECK(lut->transform_frame(host, frame) == FlowReturn::Error);
    CHECK(offered->outstanding() == 1);
    CHECK(offered->release(frame, nullptr) && offered->outstanding() == 0);
    CHECK(!offered->release(frame, nullptr));  // not outstanding any more

    // the system-memory elements offer / use page-locked pools, and only for system-memory caps
    {
        auto sys_lut = element_factory_make("colorlut");
        sys_lut->set_property("location", Value{std::string(cube)});
        CHECK(sys_lut->start().ok());
        Caps sys;
        sys.formats = {"RGBA"};
        sys.width = 320, sys.height = 200;
        AllocationQuery q;
        q.caps = sys;
        CHECK(sys_lut->propose_allocation(q).empty() && q.video_meta);
        CHECK(q.pools.size() == 1 && q.pools[0].pool->host_pinned());
        CHECK(q.pools[0].size == 320u * 4 * 200);
        VideoFrameRef f;
        CHECK(q.pools[0].pool->acquire(f) && f.memory == B200VF_MEM_HOST && f.stride == 320 * 4);
        static_cast<uint8_t *>(f.data)[0] = 1;  // ordinary, writable system memory
        CHECK(q.pools[0].pool->release(f, nullptr));
        AllocationQuery d;
        d.caps = sys;
        CHECK(sys_lut->decide_allocation(d).empty() && d.pools.size() == 1 &&
              d.pools[0].pool->host_pinned());
        AllocationQuery kept;  // a downstream proposal is left alone
        kept.caps = sys;
        kept.pools.push_back({offered, offered->size(), 0, 0});
        CHECK(sys_lut->decide_allocation(kept).empty() && kept.pools.size() == 1 &&
              kept.pools[0].pool == offered);
        AllocationQuery no_pool_wanted;
        no_pool_wanted.caps = sys;
        no_pool_wanted.need_pool = false;
        CHECK(sys_lut->propose_allocation(no_pool_wanted).empty() && no_pool_wanted.pools.empty());
        AllocationQuery cuda_q;  // device-memory caps are not this element's business
        cuda_q.caps = caps;
        CHECK(sys_lut->propose_allocation(cuda_q).empty() && cuda_q.pools.empty());
        auto sys_hsv = element_factory_make("hsvfilter");  // in place: proposes, decides nothing
        AllocationQuery hq, hd;
        hq.caps = hd.caps = sys;
        CHECK(sys_hsv->propose_allocation(hq).empty() && hq.pools.size() == 1);
        CHECK(sys_hsv->decide_allocation(hd).empty() && hd.pools.empty());
        // a CUDA-memory element does not accept a page-locked host pool as its device pool
        AllocationQuery mixed;
        mixed.caps = caps;
        mixed.pools.push_back({DeviceBufferPool::create(0, caps, 0, 0, &err, true), 0, 0, 0});
        auto pinned = mixed.pools[0].pool;
        CHECK(pinned && lut->decide_allocation(mixed).empty() && mixed.pools[0].pool != pinned &&
              !mixed.pools[0].pool->host_pinned());
    }

    // stop drops the LUT and the context: negotiation fails as before start
    CHECK(lut->stop().ok());
    CHECK(lut->set_caps(caps, caps) == "No LUT configured");
    AllocationQuery after;
    after.caps = caps;
    CHECK(lut->propose_allocation(after) == "Device not configured");
}

int main(int argc, char **argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "cpu"))
        check_cpu();
    else if (argc >= 3 && !std::strcmp(argv[1], "gpu"))
        check_gpu(argv[2]);
    else
        return 2;
    if (g_failed) return 1;
    std::printf("negotiation_check %s: ok\n", argv[1]);
    return 0;
}
