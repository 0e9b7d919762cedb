// vf_elements.cpp — see vf_elements.hpp.  Element shells only: every pixel goes through
// the C ABI of libb200vf.so; there is no CPU pixel path here.
#include "vf_elements.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "vf_elements_c.h"

namespace b200vf {

namespace {
const float kFmax = std::numeric_limits<float>::max();

ParamSpec float_spec(const char *name, const char *nick, const char *blurb, float def,
                     float lo = -kFmax, float hi = kFmax) {
    ParamSpec p;
    p.name = name, p.nick = nick, p.blurb = blurb, p.type = "gfloat";
    p.default_value = def, p.minimum = lo, p.maximum = hi;
    p.mutability = ParamMutability::Playing;  // .mutable_playing()
    return p;
}

std::vector<PadTemplate> templates(std::vector<std::string> sink, std::vector<std::string> src,
                                   std::vector<std::string> features = {}) {
    return {PadTemplate{"sink", PadDirection::Sink, "always", std::move(sink), features},
            PadTemplate{"src", PadDirection::Src, "always", std::move(src), features}};
}

bool contains(const std::vector<std::string> &v, const std::string &s) {
    return std::find(v.begin(), v.end(), s) != v.end();
}
}  // namespace

bool Caps::has_feature(const std::string &f) const { return contains(features, f); }

// =====================================================================================
// DeviceBufferPool
// =====================================================================================
std::shared_ptr<DeviceBufferPool> DeviceBufferPool::create(int device, const Caps &caps,
                                                           uint32_t min_buffers, uint32_t max_buffers,
                                                           std::string *error, bool host_pinned) {
    auto err = [&](const std::string &m) {
        if (error) *error = m;
        return std::shared_ptr<DeviceBufferPool>();
    };
    // VideoInfo::from_caps: fixed caps only
    if (caps.any_format || caps.formats.size() != 1 || caps.width == 0 || caps.height == 0)
        return err("caps are not fixed");
    int fmt = b200vf_format_from_name(caps.formats[0].c_str());
    if (fmt < 0) return err("unknown video format " + caps.formats[0]);
    b200vf_pool_config cfg{caps.width, caps.height, (uint32_t)fmt, min_buffers, max_buffers,
                           host_pinned ? 1u : 0u};
    b200vf_pool *raw = nullptr;
    if (b200vf_pool_create(device, &cfg, &raw) != B200VF_OK) return err(b200vf_last_error(nullptr));
    std::shared_ptr<DeviceBufferPool> p(new DeviceBufferPool());
    p->pool_ = raw;
    p->device_ = device;
    p->host_pinned_ = host_pinned;
    p->caps_ = caps;
    b200vf_pool_stats st{};
    b200vf_pool_get_stats(raw, &st);
    p->size_ = st.frame_bytes;
    return p;
}

DeviceBufferPool::~DeviceBufferPool() { b200vf_pool_destroy(pool_); }

bool DeviceBufferPool::acquire(VideoFrameRef &out, bool dont_wait) {
    b200vf_frame f{};
    if (b200vf_pool_acquire(pool_, dont_wait ? (uint32_t)B200VF_POOL_DONTWAIT : 0u, &f) != B200VF_OK) return false;
    out.data = f.data, out.stride = f.stride, out.width = f.width, out.height = f.height;
    out.format = caps_.formats[0];
    out.memory = (b200vf_memory)f.memory;
    return true;
}

bool DeviceBufferPool::release(const VideoFrameRef &frame, void *last_use_stream) {
    b200vf_frame f{};
    f.data = frame.data;
    return b200vf_pool_release(pool_, &f, last_use_stream) == B200VF_OK;
}

uint32_t DeviceBufferPool::outstanding() const {
    b200vf_pool_stats st{};
    b200vf_pool_get_stats(pool_, &st);
    return st.outstanding;
}

uint32_t DeviceBufferPool::allocated() const {
    b200vf_pool_stats st{};
    b200vf_pool_get_stats(pool_, &st);
    return st.allocated;
}

// =====================================================================================
// VideoFilter base
// =====================================================================================
VideoFilter::~VideoFilter() {
    if (ctx_) b200vf_ctx_destroy(ctx_);
}

bool VideoFilter::set_property(const std::string &name, const Value &v) {
    for (const ParamSpec &p : properties()) {
        if (p.name != name) continue;
        if (p.type == "gfloat") {
            const float *f = std::get_if<float>(&v);
            if (!f) return false;  // "type checked upstream"
            // g_param_value_validate: out-of-range values are refused; NaN compares false
            // against both bounds and passes, exactly as CLAMP() lets it through.
            if (*f < p.minimum || *f > p.maximum) return false;
        } else if (!std::holds_alternative<std::string>(v) &&
                   !std::holds_alternative<std::monostate>(v)) {
            return false;
        }
        return store(name, v);
    }
    return false;
}

std::optional<Value> VideoFilter::property(const std::string &name) const { return load(name); }

ErrorMessage VideoFilter::start() {
    if (!ctx_) {
        int rc = b200vf_ctx_create(device_, &ctx_);
        if (rc != B200VF_OK) {
            ctx_ = nullptr;
            return {ResourceError::Failed,
                    std::string("CUDA context: ") + b200vf_last_error(nullptr)};
        }
        if (frames_in_flight_) b200vf_ctx_set_option(ctx_, "host.async", 1);
    }
    return {};
}

ErrorMessage VideoFilter::stop() {
    if (ctx_) {
        b200vf_ctx_destroy(ctx_);  // completes whatever is still in flight
        ctx_ = nullptr;
    }
    // frames still held back (a device-following restart in mid-stream) are complete now; their
    // tickets belonged to the context that is gone: 0 = nothing to wait for when they are handed out
    for (auto &q : queued_) q.first = 0;
    return {};
}

// ---- queued operation (submit_input_buffer / generate_output) ------------------------------
ErrorMessage VideoFilter::set_frames_in_flight(unsigned frames) {
    if (frames > 14)  // the library keeps at most 15 host-frame calls in flight
        return {ResourceError::Settings, "at most 14 frames can be held back"};
    if (!queued_.empty()) return {ResourceError::Failed, "drain the element before changing its latency"};
    frames_in_flight_ = frames;
    if (ctx_ && b200vf_ctx_set_option(ctx_, "host.async", frames ? 1 : 0) != B200VF_OK)
        return {ResourceError::Failed, b200vf_last_error(ctx_)};
    return {};
}

FlowReturn VideoFilter::queued(FlowReturn rc, const VideoFrameRef &out) {
    if (rc == FlowReturn::Ok) queued_.emplace_back(b200vf_ctx_host_ticket(ctx_), out);
    return rc;
}

FlowReturn VideoFilter::submit_input_frame(const VideoFrameRef &in, VideoFrameRef &out) {
    return queued(transform_frame(in, out), out);
}

FlowReturn VideoFilter::submit_input_frame_ip(VideoFrameRef &frame) {
    return queued(transform_frame_ip(frame), frame);
}

VideoFilter::GenerateOutput VideoFilter::pop_output(VideoFrameRef &done) {
    const uint64_t ticket = queued_.front().first;
    // complete on return for synchronous calls (pageable frames) and device memory (stream-ordered)
    if (ctx_ && b200vf_ctx_host_wait(ctx_, ticket) != B200VF_OK) {
        flow_error(b200vf_last_error(ctx_));
        return GenerateOutput::Error;
    }
    done = queued_.front().second;
    queued_.pop_front();
    return GenerateOutput::Buffer;
}

VideoFilter::GenerateOutput VideoFilter::generate_output(VideoFrameRef &done) {
    return queued_.size() > frames_in_flight_ ? pop_output(done) : GenerateOutput::NoOutput;
}

VideoFilter::GenerateOutput VideoFilter::drain(VideoFrameRef &done) {
    return queued_.empty() ? GenerateOutput::NoOutput : pop_output(done);
}

Caps VideoFilter::transform_caps(PadDirection, const Caps &caps, const Caps *filter) const {
    if (!filter || filter->any_format) return caps;
    Caps out = caps;  // features and geometry pass through unchanged
    out.any_format = false;
    out.formats.clear();
    for (const std::string &f : filter->formats)
        if (caps.any_format || contains(caps.formats, f)) out.formats.push_back(f);
    return out;
}

std::string VideoFilter::set_caps(const Caps &incaps, const Caps &outcaps) {
    incaps_ = incaps;
    outcaps_ = outcaps;
    return {};
}

namespace {
bool fixed_system_caps(const Caps &c) {
    return !c.any_format && c.formats.size() == 1 && c.width && c.height && c.features.empty();
}
}  // namespace

// System memory: offer upstream a pool of page-locked frames (never an error: without a device
// or fixed caps the query simply goes on as in the reference, with no pool from this element).
std::string VideoFilter::propose_allocation(AllocationQuery &query) {
    query.video_meta = true;  // GstVideoFilter's default: GstVideoMeta is supported
    if (!query.need_pool || !fixed_system_caps(query.caps)) return {};
    if (!ctx_ && !VideoFilter::start().ok()) return {};
    std::string err;
    auto pool = DeviceBufferPool::create(device_, query.caps, 0, 0, &err, /*host_pinned=*/true);
    if (pool) query.pools.push_back({pool, pool->size(), 0, 0});
    return {};
}

// System memory, own output: keep what downstream proposed; with no proposal, allocate the output
// frames page-locked instead of from the default allocator.
std::string VideoFilter::decide_allocation(AllocationQuery &query) {
    if (!query.pools.empty() || !fixed_system_caps(query.caps)) return {};
    if (mode() == BaseTransformMode::AlwaysInPlace) return {};  // output buffer = input buffer
    if (!ctx_ && !VideoFilter::start().ok()) return {};
    std::string err;
    auto pool = DeviceBufferPool::create(device_, query.caps, 0, 0, &err, /*host_pinned=*/true);
    if (pool) query.pools.push_back({pool, pool->size(), 0, 0});
    return {};
}
void VideoFilter::before_transform(const VideoFrameRef &) {}

// d3d12colorlut/imp.rs:385-431
std::string VideoFilter::cuda_propose_allocation(AllocationQuery &query) {
    if (!ctx_) return "Device not configured";
    if (query.caps.formats.empty()) return "No caps specified";
    if (query.need_pool) {
        std::string err;
        auto pool = DeviceBufferPool::create(device_, query.caps, 0, 0, &err);
        if (!pool) return "Failed to configure pool: " + err;
        query.pools.push_back({pool, pool->size(), 0, 0});  // "gets updated size"
    }
    query.video_meta = true;
    return {};
}

// d3d12colorlut/imp.rs:433-492
std::string VideoFilter::cuda_decide_allocation(AllocationQuery &query) {
    if (!ctx_) return "Device not configured";
    if (query.caps.formats.empty()) return "No caps specified";
    const bool update_pool = !query.pools.empty();
    AllocationPool entry;
    if (update_pool) entry = query.pools.front();
    // keep the downstream pool only if it lives on our device and has our geometry
    if (entry.pool && (entry.pool->device() != device_ || entry.pool->host_pinned() ||
                       entry.pool->caps().formats != query.caps.formats ||
                       entry.pool->caps().width != query.caps.width ||
                       entry.pool->caps().height != query.caps.height))
        entry.pool.reset();
    if (!entry.pool) {
        std::string err;
        entry.pool = DeviceBufferPool::create(device_, query.caps, entry.min_buffers,
                                              entry.max_buffers, &err);
        if (!entry.pool) return "Failed to configure pool: " + err;
    }
    entry.size = entry.pool->size();
    if (update_pool)
        query.pools.front() = entry;
    else
        query.pools.push_back(entry);
    return {};
}

// d3d12colorlut/imp.rs:494-542: follow the device of the incoming memory
void VideoFilter::cuda_before_transform(const VideoFrameRef &inbuf) {
    if (!ctx_) {
        last_error_ = "No device configured";
        return;
    }
    if (!incaps_ || !outcaps_) {
        last_error_ = "No caps configured";
        return;
    }
    if (!inbuf.data) {
        last_error_ = "Empty buffer";
        return;
    }
    uint32_t memory = B200VF_MEM_HOST;
    int mem_device = -1;
    if (b200vf_pointer_info(inbuf.data, &memory, &mem_device) != B200VF_OK ||
        memory != B200VF_MEM_DEVICE) {
        last_error_ = "Wrong memory type";
        return;
    }
    if (mem_device == device_) return;
    // "Device updated from … to …": drop the context, recreate it where the memory lives
    const Caps incaps = *incaps_, outcaps = *outcaps_;
    stop();
    device_ = mem_device;
    ErrorMessage e = start();
    std::string err = e.ok() ? set_caps(incaps, outcaps) : e.message;
    if (!err.empty()) {
        last_error_ = "Failed to recreate CUDA context: " + err;
        return;
    }
    reconfigure_ = true;  // reconfigure_src: downstream allocation is renegotiated
}

bool VideoFilter::require_device_memory(const VideoFrameRef &f) {
    return f.memory == B200VF_MEM_DEVICE;
}

FlowReturn VideoFilter::transform_frame(const VideoFrameRef &, VideoFrameRef &) {
    return flow_error("transform_frame not implemented for this element");
}

FlowReturn VideoFilter::transform_frame_ip(VideoFrameRef &) {
    return flow_error("transform_frame_ip not implemented for this element");
}

bool VideoFilter::make_frame(const VideoFrameRef &f, b200vf_frame &out) {
    int fmt = b200vf_format_from_name(f.format.c_str());
    if (fmt < 0) return false;
    out.data = f.data;
    out.stride = f.stride;
    out.width = f.width;
    out.height = f.height;
    out.format = (uint32_t)fmt;
    out.memory = (uint32_t)f.memory;
    return true;
}

FlowReturn VideoFilter::flow_error(const std::string &why) {
    last_error_ = why;  // gst::error!(CAT, …)
    return FlowReturn::Error;
}

// =====================================================================================
// colorlut — video/colorlut/src/colorlut/imp.rs
// =====================================================================================
const ElementMetadata &ColorLut::metadata() const {
    static const ElementMetadata m{"Color LUT", "Filter/Effect/Video", "Apply color lookup table",
                                   "Seungha Yang <seungha@centricular.com>"};
    return m;
}

const std::vector<PadTemplate> &ColorLut::pad_templates() const {
    // little-endian host order, imp.rs:128-134
    static const auto t = templates({"RGBA64_LE", "RGBA64_BE", "RGBA"},
                                    {"RGBA64_LE", "RGBA64_BE", "RGBA"});
    return t;
}

const std::vector<ParamSpec> &ColorLut::properties() const {
    static const std::vector<ParamSpec> p = [] {
        ParamSpec s;
        s.name = "location", s.nick = "Location";
        s.blurb = "Location of the LUT file to read from";
        s.type = "gchararray";
        s.default_value = std::monostate{};  // NULL
        s.mutability = ParamMutability::Ready;  // .mutable_ready()
        return std::vector<ParamSpec>{s};
    }();
    return p;
}

bool ColorLut::store(const std::string &name, const Value &v) {
    if (name != "location") return false;
    std::lock_guard<std::mutex> g(settings_mu_);
    if (const std::string *s = std::get_if<std::string>(&v))
        location_ = *s;
    else
        location_.reset();
    return true;
}

std::optional<Value> ColorLut::load(const std::string &name) const {
    if (name != "location") return std::nullopt;
    std::lock_guard<std::mutex> g(settings_mu_);
    if (location_) return Value{*location_};
    return Value{std::monostate{}};
}

ErrorMessage ColorLut::start() {
    std::optional<std::string> location;
    {
        std::lock_guard<std::mutex> g(settings_mu_);
        location = location_;
    }
    if (!location)  // imp.rs:175-180
        return {ResourceError::Settings, "LUT file location is not configured"};
    ErrorMessage base = VideoFilter::start();
    if (!base.ok()) return base;
    int rc = b200vf_colorlut_set_lut_file(ctx_, location->c_str());
    if (rc == B200VF_ERR_PARSE || rc == B200VF_ERR_IO)  // imp.rs:182-187
        return {ResourceError::Read, b200vf_last_error(ctx_)};
    if (rc != B200VF_OK) return {ResourceError::Failed, b200vf_last_error(ctx_)};
    std::lock_guard<std::mutex> g(state_mu_);
    lut_loaded_ = true;  // imp.rs:191
    return {};
}

bool ColorLut::lut_loaded() {
    std::lock_guard<std::mutex> g(state_mu_);
    return lut_loaded_;
}

ErrorMessage ColorLut::stop() {
    {
        std::lock_guard<std::mutex> g(state_mu_);
        lut_loaded_ = false;  // imp.rs:197
    }
    return VideoFilter::stop();
}

FlowReturn ColorLut::transform_frame(const VideoFrameRef &in, VideoFrameRef &out) {
    std::lock_guard<std::mutex> g(state_mu_);  // imp.rs:208
    if (!lut_loaded_ || !ctx_) return flow_error("No LUT configured");  // imp.rs:210-213
    b200vf_frame fi, fo;
    if (!make_frame(in, fi) || !make_frame(out, fo)) return flow_error("unknown video format");
    int rc = b200vf_colorlut_process(ctx_, &fi, &fo);
    if (rc != B200VF_OK) return flow_error(b200vf_last_error(ctx_));
    return FlowReturn::Ok;
}

// =====================================================================================
// hsvfilter — video/hsv/src/hsvfilter/imp.rs
// =====================================================================================
const ElementMetadata &HsvFilter::metadata() const {
    static const ElementMetadata m{
        "HSV filter", "Filter/Effect/Converter/Video",
        "Works within the HSV colorspace to apply transformations to incoming frames",
        "Julien Bardagi <julien.bardagi@gmail.com>"};
    return m;
}

const std::vector<PadTemplate> &HsvFilter::pad_templates() const {
    static const std::vector<std::string> f{"RGBx", "xRGB", "BGRx", "xBGR", "RGBA",
                                            "ARGB", "BGRA", "ABGR", "RGB",  "BGR"};  // :278-289
    static const auto t = templates(f, f);
    return t;
}

const std::vector<ParamSpec> &HsvFilter::properties() const {
    static const std::vector<ParamSpec> p{
        float_spec("hue-shift", "Hue shift", "Hue shifting in degrees", 0.0f),
        float_spec("saturation-mul", "Saturation multiplier",
                   "Saturation multiplier to apply to the saturation value (before offset)", 1.0f),
        float_spec("saturation-off", "Saturation offset",
                   "Saturation offset to add to the saturation value (after multiplier)", 0.0f),
        float_spec("value-mul", "Value multiplier",
                   "Value multiplier to apply to the value (before offset)", 1.0f),
        float_spec("value-off", "Value offset",
                   "Value offset to add to the value (after multiplier)", 0.0f)};
    return p;
}

bool HsvFilter::store(const std::string &name, const Value &v) {
    const float f = std::get<float>(v);
    std::lock_guard<std::mutex> g(settings_mu_);
    if (name == "hue-shift") settings_.hue_shift = f;
    else if (name == "saturation-mul") settings_.saturation_mul = f;
    else if (name == "saturation-off") settings_.saturation_off = f;
    else if (name == "value-mul") settings_.value_mul = f;
    else if (name == "value-off") settings_.value_off = f;
    else return false;
    return true;
}

std::optional<Value> HsvFilter::load(const std::string &name) const {
    std::lock_guard<std::mutex> g(settings_mu_);
    if (name == "hue-shift") return Value{settings_.hue_shift};
    if (name == "saturation-mul") return Value{settings_.saturation_mul};
    if (name == "saturation-off") return Value{settings_.saturation_off};
    if (name == "value-mul") return Value{settings_.value_mul};
    if (name == "value-off") return Value{settings_.value_off};
    return std::nullopt;
}

FlowReturn HsvFilter::transform_frame_ip(VideoFrameRef &frame) {
    if (!ctx_) {  // the reference element is stateless; create the context on first use
        ErrorMessage e = VideoFilter::start();
        if (!e.ok()) return flow_error(e.message);
    }
    b200vf_hsvfilter_params snapshot;
    {
        std::lock_guard<std::mutex> g(settings_mu_);
        snapshot = settings_;  // imp.rs:85 — settings copied once per frame
    }
    b200vf_frame f;
    if (!make_frame(frame, f)) return flow_error("unknown video format");
    int rc = b200vf_hsvfilter_process(ctx_, &f, &snapshot);
    if (rc != B200VF_OK) return flow_error(b200vf_last_error(ctx_));
    return FlowReturn::Ok;
}

// =====================================================================================
// hsvdetector — video/hsv/src/hsvdetector/imp.rs
// =====================================================================================
namespace {
const std::vector<std::string> kDetectorIn{"RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"};  // :78-87
const std::vector<std::string> kDetectorOut{"RGBA", "ARGB", "BGRA", "ABGR"};               // :89-96
}  // namespace

const ElementMetadata &HsvDetector::metadata() const {
    static const ElementMetadata m{"HSV detector", "Filter/Effect/Converter/Video",
                                   "Works within the HSV colorspace to mark positive pixels",
                                   "Julien Bardagi <julien.bardagi@gmail.com>"};
    return m;
}

const std::vector<PadTemplate> &HsvDetector::pad_templates() const {
    static const auto t = templates(kDetectorIn, kDetectorOut);
    return t;
}

const std::vector<ParamSpec> &HsvDetector::properties() const {
    static const std::vector<ParamSpec> p{
        float_spec("hue-ref", "Hue reference", "Hue reference in degrees", 0.0f),
        float_spec("hue-var", "Hue variation",
                   "Allowed hue variation from the reference hue angle, in degrees", 10.0f, 0.0f,
                   180.0f),
        float_spec("saturation-ref", "Saturation reference", "Reference saturation value", 0.0f,
                   0.0f, 1.0f),
        float_spec("saturation-var", "Saturation variation",
                   "Allowed saturation variation from the reference value", 0.15f, 0.0f, 1.0f),
        float_spec("value-ref", "Value reference", "Reference value value", 0.0f, 0.0f, 1.0f),
        float_spec("value-var", "Value variation",
                   "Allowed value variation from the reference value", 0.3f, 0.0f, 1.0f)};
    return p;
}

bool HsvDetector::store(const std::string &name, const Value &v) {
    const float f = std::get<float>(v);
    std::lock_guard<std::mutex> g(settings_mu_);
    if (name == "hue-ref") settings_.hue_ref = f;
    else if (name == "hue-var") settings_.hue_var = f;
    else if (name == "saturation-ref") settings_.saturation_ref = f;
    else if (name == "saturation-var") settings_.saturation_var = f;
    else if (name == "value-ref") settings_.value_ref = f;
    else if (name == "value-var") settings_.value_var = f;
    else return false;
    return true;
}

std::optional<Value> HsvDetector::load(const std::string &name) const {
    std::lock_guard<std::mutex> g(settings_mu_);
    if (name == "hue-ref") return Value{settings_.hue_ref};
    if (name == "hue-var") return Value{settings_.hue_var};
    if (name == "saturation-ref") return Value{settings_.saturation_ref};
    if (name == "saturation-var") return Value{settings_.saturation_var};
    if (name == "value-ref") return Value{settings_.value_ref};
    if (name == "value-var") return Value{settings_.value_var};
    return std::nullopt;
}

// imp.rs:386-419: every structure's `format` becomes the full list of the OTHER pad, then
// the result is intersected with `filter` in First mode (filter's order wins).
Caps HsvDetector::transform_caps(PadDirection direction, const Caps &, const Caps *filter) const {
    Caps other;
    other.formats = direction == PadDirection::Src ? kDetectorIn : kDetectorOut;
    if (!filter || filter->any_format) return other;
    Caps out;
    for (const std::string &f : filter->formats)
        if (contains(other.formats, f)) out.formats.push_back(f);
    return out;
}

FlowReturn HsvDetector::transform_frame(const VideoFrameRef &in, VideoFrameRef &out) {
    if (!ctx_) {
        ErrorMessage e = VideoFilter::start();
        if (!e.ok()) return flow_error(e.message);
    }
    b200vf_hsvdetector_params snapshot;
    {
        std::lock_guard<std::mutex> g(settings_mu_);  // imp.rs:110
        snapshot = settings_;
    }
    b200vf_frame fi, fo;
    if (!make_frame(in, fi) || !make_frame(out, fo)) return flow_error("unknown video format");
    int rc = b200vf_hsvdetector_process(ctx_, &fi, &fo, &snapshot);
    if (rc != B200VF_OK) return flow_error(b200vf_last_error(ctx_));
    return FlowReturn::Ok;
}

// =====================================================================================
// CUDA-memory variants — after video/colorlut/src/d3d12colorlut/imp.rs
// =====================================================================================
namespace {
const std::vector<std::string> kCudaFeature{kCapsFeatureCudaMemory};
}

const ElementMetadata &CudaColorLut::metadata() const {
    static const ElementMetadata m{"CUDA Color LUT", "Filter/Video", "Apply Color LUT using CUDA",
                                   "b200vf"};
    return m;
}

const std::vector<PadTemplate> &CudaColorLut::pad_templates() const {
    static const auto t = templates({"RGBA64_LE", "RGBA64_BE", "RGBA"},
                                    {"RGBA64_LE", "RGBA64_BE", "RGBA"}, kCudaFeature);
    return t;
}

std::string CudaColorLut::set_caps(const Caps &incaps, const Caps &outcaps) {
    if (!lut_loaded()) return "No LUT configured";  // :360-362
    if (!ctx_) return "No Context configured";       // :364-366
    if (incaps.formats.size() != 1 || incaps.width == 0 || incaps.height == 0)
        return "Failed to parse output caps";
    return VideoFilter::set_caps(incaps, outcaps);
}

FlowReturn CudaColorLut::transform_frame(const VideoFrameRef &in, VideoFrameRef &out) {
    if (!require_device_memory(in) || !require_device_memory(out))
        return flow_error("Wrong memory type");
    return ColorLut::transform_frame(in, out);
}

const ElementMetadata &CudaHsvFilter::metadata() const {
    static const ElementMetadata m{
        "CUDA HSV filter", "Filter/Effect/Converter/Video",
        "Works within the HSV colorspace to apply transformations to incoming frames, using CUDA",
        "b200vf"};
    return m;
}

const std::vector<PadTemplate> &CudaHsvFilter::pad_templates() const {
    static const auto t = [this] {
        const std::vector<std::string> &f = HsvFilter::pad_templates()[0].formats;
        return templates(f, f, kCudaFeature);
    }();
    return t;
}

FlowReturn CudaHsvFilter::transform_frame_ip(VideoFrameRef &frame) {
    if (!require_device_memory(frame)) return flow_error("Wrong memory type");
    return HsvFilter::transform_frame_ip(frame);
}

const ElementMetadata &CudaHsvDetector::metadata() const {
    static const ElementMetadata m{"CUDA HSV detector", "Filter/Effect/Converter/Video",
                                   "Works within the HSV colorspace to mark positive pixels, using CUDA",
                                   "b200vf"};
    return m;
}

const std::vector<PadTemplate> &CudaHsvDetector::pad_templates() const {
    static const auto t = templates(kDetectorIn, kDetectorOut, kCudaFeature);
    return t;
}

Caps CudaHsvDetector::transform_caps(PadDirection direction, const Caps &caps,
                                     const Caps *filter) const {
    Caps out = HsvDetector::transform_caps(direction, caps, filter);
    out.features = kCudaFeature;  // format changes, memory feature and geometry do not
    out.width = caps.width, out.height = caps.height;
    return out;
}

FlowReturn CudaHsvDetector::transform_frame(const VideoFrameRef &in, VideoFrameRef &out) {
    if (!require_device_memory(in) || !require_device_memory(out))
        return flow_error("Wrong memory type");
    return HsvDetector::transform_frame(in, out);
}

// =====================================================================================
// registration
// =====================================================================================
const std::vector<PluginDescriptor> &plugins() {
    // video/colorlut/src/lib.rs:33-43, video/hsv/src/lib.rs:32-42 (+ gst_plugins_cache.json)
    static const std::vector<PluginDescriptor> p{
        {"colorlut", "GStreamer Color LUT Plugin", "gstcolorlut", "MPL-2.0", "gst-plugin-colorlut",
         {"colorlut"}},
        {"hsv", "GStreamer plugin with HSV manipulation elements", "gsthsv", "MIT/X11",
         "gst-plugin-hsv", {"hsvdetector", "hsvfilter"}},
        // no reference counterpart on Linux; the precedent is d3d12colorlut inside `colorlut`
        {"b200vf", "CUDA-memory variants of colorlut, hsvfilter and hsvdetector", "gstb200vf",
         "MPL-2.0", "gst-plugins-rs_b200", {"cudacolorlut", "cudahsvdetector", "cudahsvfilter"}}};
    return p;
}

std::unique_ptr<VideoFilter> element_factory_make(const std::string &name, int device) {
    if (name == "colorlut") return std::make_unique<ColorLut>(device);
    if (name == "hsvfilter") return std::make_unique<HsvFilter>(device);
    if (name == "hsvdetector") return std::make_unique<HsvDetector>(device);
    if (name == "cudacolorlut") return std::make_unique<CudaColorLut>(device);
    if (name == "cudahsvfilter") return std::make_unique<CudaHsvFilter>(device);
    if (name == "cudahsvdetector") return std::make_unique<CudaHsvDetector>(device);
    return nullptr;
}

namespace {
std::string json_escape(const std::string &s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        o += c;
    }
    return o;
}

// %g with 6 significant digits is what gst-inspect / the docs cache prints for gfloat
std::string fmt_float(float v) {
    char b[64];
    std::snprintf(b, sizeof b, "%g", (double)v);
    return b;
}

std::string list_json(const std::vector<std::string> &v) {
    std::string o = "[";
    for (size_t i = 0; i < v.size(); i++) o += (i ? ", \"" : "\"") + json_escape(v[i]) + "\"";
    return o + "]";
}
}  // namespace

std::string describe_element_json(const std::string &name) {
    std::unique_ptr<VideoFilter> e = element_factory_make(name, 0);
    if (!e) return "{}";
    const PluginDescriptor *plugin = nullptr;
    for (const PluginDescriptor &p : plugins())
        if (contains(p.elements, name)) plugin = &p;
    std::ostringstream o;
    o << "{\"gtype\": \"" << e->type_name() << "\", \"parent\": \"GstVideoFilter\", \"klass\": \""
      << json_escape(e->metadata().klass) << "\", \"long-name\": \""
      << json_escape(e->metadata().long_name) << "\", \"description\": \""
      << json_escape(e->metadata().description) << "\", \"author\": \""
      << json_escape(e->metadata().author) << "\", \"rank\": \"none\", \"plugin\": \""
      << plugin->name << "\", \"filename\": \"" << plugin->filename << "\", \"license\": \""
      << plugin->license << "\", \"mode\": \""
      << (e->mode() == BaseTransformMode::AlwaysInPlace ? "AlwaysInPlace" : "NeverInPlace")
      << "\"";
    for (const PadTemplate &t : e->pad_templates()) {
        o << ", \"" << t.name << "_formats\": " << list_json(t.formats);
        if (!t.features.empty()) o << ", \"" << t.name << "_features\": " << list_json(t.features);
    }
    o << ", \"properties\": {";
    bool first = true;
    for (const ParamSpec &p : e->properties()) {
        o << (first ? "" : ", ") << "\"" << p.name << "\": {\"type\": \"" << p.type
          << "\", \"mutable\": \""
          << (p.mutability == ParamMutability::Ready ? "ready" : "playing")
          << "\", \"readable\": true, \"writable\": true, \"default\": \"";
        if (const float *f = std::get_if<float>(&p.default_value))
            o << fmt_float(*f) << "\", \"min\": \"" << fmt_float(p.minimum) << "\", \"max\": \""
              << fmt_float(p.maximum) << "\"";
        else
            o << "NULL\"";
        o << "}";
        first = false;
    }
    o << "}}";
    return o.str();
}

}  // namespace b200vf

// =====================================================================================
// C view of the element layer (for the Python test / bench harness)
// =====================================================================================
using b200vf::VideoFilter;

struct b200vf_element {
    std::unique_ptr<VideoFilter> impl;
    std::string scratch;
};

static b200vf::VideoFrameRef to_ref(const b200vf_frame *f) {
    b200vf::VideoFrameRef r;
    r.data = f->data, r.stride = f->stride, r.width = f->width, r.height = f->height;
    const char *n = b200vf_format_name(f->format);
    r.format = n ? n : "";
    r.memory = (b200vf_memory)f->memory;
    return r;
}

extern "C" {

b200vf_element *b200vf_element_new(const char *factory_name, int device) {
    if (!factory_name) return nullptr;
    auto impl = b200vf::element_factory_make(factory_name, device);
    if (!impl) return nullptr;
    auto *e = new b200vf_element();
    e->impl = std::move(impl);
    return e;
}

void b200vf_element_free(b200vf_element *e) { delete e; }

int b200vf_element_set_float(b200vf_element *e, const char *name, float v) {
    return e && name && e->impl->set_property(name, b200vf::Value{v}) ? 1 : 0;
}

int b200vf_element_set_string(b200vf_element *e, const char *name, const char *v) {
    if (!e || !name) return 0;
    return e->impl->set_property(name, v ? b200vf::Value{std::string(v)}
                                         : b200vf::Value{std::monostate{}})
               ? 1
               : 0;
}

int b200vf_element_get_float(b200vf_element *e, const char *name, float *out) {
    if (!e || !name || !out) return 0;
    auto v = e->impl->property(name);
    if (!v) return 0;
    if (const float *f = std::get_if<float>(&*v)) return *out = *f, 1;
    return 0;
}

const char *b200vf_element_get_string(b200vf_element *e, const char *name) {
    if (!e || !name) return nullptr;
    auto v = e->impl->property(name);
    if (!v) return nullptr;
    if (const std::string *s = std::get_if<std::string>(&*v)) return e->scratch = *s, e->scratch.c_str();
    return nullptr;
}

int b200vf_element_start(b200vf_element *e) {
    if (!e) return (int)b200vf::ResourceError::Failed;
    b200vf::ErrorMessage m = e->impl->start();
    e->scratch = m.message;
    return (int)m.domain;
}

int b200vf_element_stop(b200vf_element *e) {
    if (!e) return (int)b200vf::ResourceError::Failed;
    return (int)e->impl->stop().domain;
}

const char *b200vf_element_message(b200vf_element *e) {
    if (!e) return "";
    return e->scratch.empty() ? e->impl->last_error().c_str() : e->scratch.c_str();
}

int b200vf_element_transform_frame(b200vf_element *e, const b200vf_frame *in,
                                   const b200vf_frame *out) {
    if (!e || !in || !out) return (int)b200vf::FlowReturn::Error;
    e->scratch.clear();
    b200vf::VideoFrameRef i = to_ref(in), o = to_ref(out);
    return (int)e->impl->transform_frame(i, o);
}

int b200vf_element_transform_frame_ip(b200vf_element *e, const b200vf_frame *frame) {
    if (!e || !frame) return (int)b200vf::FlowReturn::Error;
    e->scratch.clear();
    b200vf::VideoFrameRef f = to_ref(frame);
    return (int)e->impl->transform_frame_ip(f);
}

int b200vf_element_set_frames_in_flight(b200vf_element *e, unsigned frames) {
    if (!e) return (int)b200vf::ResourceError::Failed;
    b200vf::ErrorMessage m = e->impl->set_frames_in_flight(frames);
    e->scratch = m.message;
    return (int)m.domain;
}

int b200vf_element_submit_input_frame(b200vf_element *e, const b200vf_frame *in, const b200vf_frame *out) {
    if (!e || !in) return (int)b200vf::FlowReturn::Error;
    e->scratch.clear();
    b200vf::VideoFrameRef i = to_ref(in);
    if (!out) return (int)e->impl->submit_input_frame_ip(i);
    b200vf::VideoFrameRef o = to_ref(out);
    return (int)e->impl->submit_input_frame(i, o);
}

namespace {
int hand_out(b200vf_element *e, b200vf::VideoFilter::GenerateOutput g, const b200vf::VideoFrameRef &f,
             b200vf_frame *done) {
    if (g == b200vf::VideoFilter::GenerateOutput::Error) return (int)b200vf::FlowReturn::Error;
    if (g == b200vf::VideoFilter::GenerateOutput::NoOutput) return 0;
    int fmt = b200vf_format_from_name(f.format.c_str());
    *done = b200vf_frame{f.data, f.stride, f.width, f.height, (uint32_t)(fmt < 0 ? 0 : fmt), (uint32_t)f.memory};
    (void)e;
    return 1;
}
}  // namespace

int b200vf_element_generate_output(b200vf_element *e, b200vf_frame *done) {
    if (!e || !done) return (int)b200vf::FlowReturn::Error;
    e->scratch.clear();
    b200vf::VideoFrameRef f;
    return hand_out(e, e->impl->generate_output(f), f, done);
}

int b200vf_element_drain(b200vf_element *e, b200vf_frame *done) {
    if (!e || !done) return (int)b200vf::FlowReturn::Error;
    e->scratch.clear();
    b200vf::VideoFrameRef f;
    return hand_out(e, e->impl->drain(f), f, done);
}

const char *b200vf_element_transform_caps(b200vf_element *e, int direction_is_src,
                                          const char *formats_csv, const char *filter_csv) {
    if (!e) return nullptr;
    auto split = [](const char *csv) {
        b200vf::Caps c;
        if (!csv) {
            c.any_format = true;
            return c;
        }
        std::stringstream ss(csv);
        std::string tok;
        while (std::getline(ss, tok, ','))
            if (!tok.empty()) c.formats.push_back(tok);
        return c;
    };
    b200vf::Caps caps = split(formats_csv), filter = split(filter_csv);
    b200vf::Caps out = e->impl->transform_caps(
        direction_is_src ? b200vf::PadDirection::Src : b200vf::PadDirection::Sink, caps,
        filter_csv ? &filter : nullptr);
    e->scratch.clear();
    for (size_t i = 0; i < out.formats.size(); i++) e->scratch += (i ? "," : "") + out.formats[i];
    return e->scratch.c_str();
}

const char *b200vf_element_describe(const char *factory_name) {
    static thread_local std::string s;
    s = b200vf::describe_element_json(factory_name ? factory_name : "");
    return s.c_str();
}

void *b200vf_element_context(b200vf_element *e) { return e ? (void *)e->impl->context() : nullptr; }

}  // extern "C"
