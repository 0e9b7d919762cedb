// cfg1_pipeline.cpp — BASELINE.json configs[0] driven through the C++ element layer exactly as a
// GStreamer streaming thread would drive the reference element:
//
//   videotestsrc num-buffers=300 ! video/x-raw,format=RGBA,width=1920,height=1080
//       ! colorlut location=<33^3 .cube> ! fakesink
//
// "videotestsrc" = a generated SMPTE-like bars frame in system memory (one fresh buffer per
// push, as a source would hand over), "fakesink" = the output buffer is dropped.  Prints one
// JSON line with frames/s.  Usage: cfg1_pipeline <lut.cube> [num_buffers] [width] [height] [copy_threads] [chunk_bytes] [pool|register|queued]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../gst-plugins-rs_b200/elements/vf_elements.hpp"

static void fill_bars(std::vector<uint8_t> &f, uint32_t w, uint32_t h, unsigned frame_no) {
    static const uint8_t bars[7][3] = {{191, 191, 191}, {191, 191, 0}, {0, 191, 191}, {0, 191, 0},
                                       {191, 0, 191},   {191, 0, 0},   {0, 0, 191}};
    const uint32_t split = h * 2 / 3;
    for (uint32_t y = 0; y < h; y++) {
        uint8_t *row = f.data() + (size_t)y * w * 4;
        for (uint32_t x = 0; x < w; x++) {
            uint8_t *p = row + 4 * x;
            if (y < split) {
                const uint8_t *c = bars[x * 7 / w];
                p[0] = c[0], p[1] = c[1], p[2] = c[2];
            } else {
                p[0] = p[1] = p[2] = (uint8_t)(x * 255 / (w - 1));
            }
            p[3] = 255;
        }
    }
    f[0] = (uint8_t)frame_no;  // buffers differ, like a live source with a moving element
}

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <lut.cube> [num_buffers] [width] [height]\n", argv[0]);
        return 2;
    }
    const unsigned n = argc > 2 ? (unsigned)std::atoi(argv[2]) : 300;
    const uint32_t w = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 1920;
    const uint32_t h = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 1080;

    auto lut = b200vf::element_factory_make("colorlut");
    if (!lut->set_property("location", b200vf::Value{std::string(argv[1])})) return 3;
    b200vf::ErrorMessage err = lut->start();  // READY → PAUSED: parse + upload the LUT
    if (!err.ok()) {
        std::fprintf(stderr, "start failed: %s\n", err.message.c_str());
        return 4;
    }

    if (argc > 5)  // dev aid: helper threads for the pageable-frame row copies
        b200vf_ctx_set_option(lut->context(), "host.copy_threads", std::atoi(argv[5]));
    if (argc > 6) b200vf_ctx_set_option(lut->context(), "host.chunk_bytes", std::atoll(argv[6]));

    // "pool": the source takes its buffers from the pool the element proposes (allocation query
    // upstream) and the element allocates its output from the pool it decides on — both are
    // page-locked system memory here, so no frame goes through the pageable bounce copy.
    // Without it: plain malloc'ed buffers, what a source that ignores the proposal hands over.
    // "queued": "pool", and the element holds one buffer back (submit_input_buffer / generate_output
    // instead of transform_frame): buffer k+1 uploads while buffer k is still on its way out.
    const bool queued = argc > 7 && !std::strcmp(argv[7], "queued");
    const bool use_pool = queued || (argc > 7 && !std::strcmp(argv[7], "pool"));
    if (queued && !lut->set_frames_in_flight(1).ok()) return 7;
    // "register": plain malloc'ed buffers again, but the context page-locks recurring ones in place
    // ("host.register": an upstream pool hands the same buffers round and round) instead of
    // bouncing every frame through pinned staging memory.  The buffers outlive the element here;
    // a shim that cannot guarantee that calls b200vf_ctx_host_memory_released from a destroy notify.
    const bool use_register = argc > 7 && !std::strcmp(argv[7], "register");
    if (use_register) b200vf_ctx_set_option(lut->context(), "host.register", 1);
    std::vector<uint8_t> src((size_t)w * h * 4), dst((size_t)w * h * 4);
    fill_bars(src, w, h, 0);
    b200vf::VideoFrameRef in{src.data(), (int64_t)w * 4, w, h, "RGBA", B200VF_MEM_HOST};
    b200vf::VideoFrameRef out{dst.data(), (int64_t)w * 4, w, h, "RGBA", B200VF_MEM_HOST};
    b200vf::AllocationQuery upstream, downstream;
    if (use_pool) {
        b200vf::Caps caps;
        caps.formats = {"RGBA"};
        caps.width = w, caps.height = h;
        upstream.caps = downstream.caps = caps;
        lut->set_caps(caps, caps);
        lut->propose_allocation(upstream);
        lut->decide_allocation(downstream);
        if (upstream.pools.empty() || downstream.pools.empty() ||
            !upstream.pools[0].pool->acquire(in) || !downstream.pools[0].pool->acquire(out)) {
            std::fprintf(stderr, "no pool was offered\n");
            return 6;
        }
        std::memcpy(in.data, src.data(), src.size());
    }
    uint8_t *const src_px = static_cast<uint8_t *>(in.data), *const dst_px = static_cast<uint8_t *>(out.data);
    // queued: a second pair of buffers, so that one pair can be in flight while the other is filled / read
    b200vf::VideoFrameRef in2 = in, out2 = out;
    if (queued) {
        if (!upstream.pools[0].pool->acquire(in2) || !downstream.pools[0].pool->acquire(out2)) return 6;
        std::memcpy(in2.data, src.data(), src.size());
    }
    auto sink = [&](const b200vf::VideoFrameRef &f) {  // fakesink: look at it, drop it
        const uint8_t *px = static_cast<const uint8_t *>(f.data);
        return (uint64_t)px[0] + px[(size_t)w * h * 2 + 1];
    };
    b200vf::VideoFrameRef done;
    for (int i = 0; i < 3; i++)  // preroll
        if (lut->transform_frame(in, out) != b200vf::FlowReturn::Ok) return 5;
    if (queued) b200vf_ctx_synchronize(lut->context());

    uint64_t checksum = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned i = 0; i < n; i++) {
        if (queued) {
            b200vf::VideoFrameRef &fi = i & 1 ? in2 : in, &fo = i & 1 ? out2 : out;
            static_cast<uint8_t *>(fi.data)[0] = (uint8_t)i;  // the source produced a new buffer
            if (lut->submit_input_frame(fi, fo) != b200vf::FlowReturn::Ok) {
                std::fprintf(stderr, "flow error: %s\n", lut->last_error().c_str());
                return 5;
            }
            const auto g = lut->generate_output(done);  // buffer i-1, complete
            if (g == b200vf::VideoFilter::GenerateOutput::Error) return 5;
            if (g == b200vf::VideoFilter::GenerateOutput::Buffer) checksum += sink(done);
            continue;
        }
        src_px[0] = (uint8_t)i;  // the source produced a new buffer
        if (lut->transform_frame(in, out) != b200vf::FlowReturn::Ok) {
            std::fprintf(stderr, "flow error: %s\n", lut->last_error().c_str());
            return 5;
        }
        (void)dst_px;
        checksum += sink(out);
    }
    while (lut->drain(done) == b200vf::VideoFilter::GenerateOutput::Buffer) checksum += sink(done);  // EOS
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    lut->stop();
    std::printf("{\"pipeline\": \"videotestsrc num-buffers=%u ! colorlut(33^3) %ux%u RGBA ! fakesink\", "
                "\"memory\": \"%s\", \"frames_per_s\": %.1f, \"seconds\": %.3f, "
                "\"checksum\": %llu}\n",
                n, w, h,
                queued ? "system (page-locked pool proposed by the element), one buffer held back (queued)"
                : use_pool ? "system (page-locked pool proposed by the element)"
                         : use_register ? "system (malloc'ed, recycled; page-locked in place on second sight)"
                                        : "system (pageable)",
                n / s, s, (unsigned long long)checksum);
    return 0;
}
