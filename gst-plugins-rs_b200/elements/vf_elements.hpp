// vf_elements.hpp — host-side mirror of the reference's element layer (L1/L2 of
// SURVEY.md §1) for `colorlut`, `hsvfilter` and `hsvdetector`, in C++ because the
// reference's Rust toolchain and GStreamer are absent from this image.
//
// Each class keeps what the reference element declares — GType name, metadata, pad
// templates, properties (name, type, default, range, mutability), BaseTransform mode,
// start/stop, transform_caps, transform_frame[_ip] and their error behaviour — and
// replaces ONLY the per-pixel loops by one call into the C ABI (include/b200vf.h).
// Reference sources mirrored:
//   video/colorlut/src/colorlut/imp.rs:45-223      ColorLut
//   video/hsv/src/hsvfilter/imp.rs:24-377           HsvFilter
//   video/hsv/src/hsvdetector/imp.rs:25-708         HsvDetector
//   video/{colorlut,hsv}/src/lib.rs, */mod.rs       plugin / element registration
#pragma once
#include <cstdint>
#include <deque>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <variant>
#include <vector>

#include "../../include/b200vf.h"

// The C++ classes are part of the library's public surface (examples/, C++ hosts).
#pragma GCC visibility push(default)
namespace b200vf {

// ---- the slice of GStreamer vocabulary the three elements touch ---------------------
enum class PadDirection { Src, Sink };
enum class BaseTransformMode { AlwaysInPlace, NeverInPlace, Both };
enum class ParamMutability { Ready, Paused, Playing };  // GST_PARAM_MUTABLE_*
enum class FlowReturn { Ok = 0, Error = -5, NotNegotiated = -4 };  // gst::FlowSuccess / FlowError
enum class ResourceError { None, Settings, Read, Failed };         // gst::ResourceError used by start()

struct ErrorMessage {  // gst::ErrorMessage
    ResourceError domain = ResourceError::None;
    std::string message;
    bool ok() const { return domain == ResourceError::None; }
};

struct ElementMetadata {  // gst::subclass::ElementMetadata
    std::string long_name, klass, description, author;
};

struct PadTemplate {  // gst::PadTemplate with VideoCapsBuilder::format_list
    std::string name;
    PadDirection direction;
    std::string presence;  // "always"
    std::vector<std::string> formats;
    std::vector<std::string> features;  // caps features; empty = memory:SystemMemory
};

inline constexpr const char *kCapsFeatureCudaMemory = "memory:CUDAMemory";

using Value = std::variant<std::monostate, float, std::string>;  // glib::Value (NULL, gfloat, gchararray)

struct ParamSpec {  // glib::ParamSpecFloat / ParamSpecString
    std::string name, nick, blurb;
    std::string type;  // "gfloat" | "gchararray"
    Value default_value;
    float minimum = -std::numeric_limits<float>::max();
    float maximum = std::numeric_limits<float>::max();
    ParamMutability mutability = ParamMutability::Playing;
};

// video/x-raw caps reduced to what transform_caps manipulates: the format list.
struct Caps {
    std::vector<std::string> formats;
    bool any_format = false;  // field absent = unconstrained
    std::vector<std::string> features;  // empty = memory:SystemMemory
    uint32_t width = 0, height = 0;     // fixed caps only (set_caps / allocation queries)
    bool has_feature(const std::string &f) const;
};

// gst_video::VideoFrameRef: plane 0 of a mapped frame.
struct VideoFrameRef {
    void *data = nullptr;
    int64_t stride = 0;
    uint32_t width = 0, height = 0;
    std::string format;  // GstVideoFormat name, e.g. "RGBA"
    b200vf_memory memory = B200VF_MEM_HOST;
};

// ---- allocation (gst::BufferPool / gst::query::Allocation, device memory only) -----------
// A pool of device frames over b200vf_pool_* — the role gst_d3d12::D3D12BufferPool plays for
// d3d12colorlut (d3d12colorlut/imp.rs:385-492).
class DeviceBufferPool {
public:
    // set_config + set_active(true): fixed caps (one format, width, height) → a pool on `device`.
    // host_pinned: page-locked system memory instead of device memory — frames of such a pool are
    // ordinary memory:SystemMemory buffers that reach the GPU without the pageable bounce copy.
    static std::shared_ptr<DeviceBufferPool> create(int device, const Caps &caps, uint32_t min_buffers,
                                                    uint32_t max_buffers, std::string *error,
                                                    bool host_pinned = false);
    ~DeviceBufferPool();
    DeviceBufferPool(const DeviceBufferPool &) = delete;
    DeviceBufferPool &operator=(const DeviceBufferPool &) = delete;

    int device() const { return device_; }
    bool host_pinned() const { return host_pinned_; }
    uint64_t size() const { return size_; }  // the "updated size" of the pool config
    const Caps &caps() const { return caps_; }
    // acquire_buffer / buffer unref.  `last_use_stream`: see b200vf_pool_release.
    bool acquire(VideoFrameRef &out, bool dont_wait = false);
    bool release(const VideoFrameRef &frame, void *last_use_stream);
    uint32_t outstanding() const;
    uint32_t allocated() const;

private:
    DeviceBufferPool() = default;
    b200vf_pool *pool_ = nullptr;
    int device_ = 0;
    bool host_pinned_ = false;
    uint64_t size_ = 0;
    Caps caps_;
};

struct AllocationPool {  // one entry of the query's pool list
    std::shared_ptr<DeviceBufferPool> pool;
    uint64_t size = 0;
    uint32_t min_buffers = 0, max_buffers = 0;
};

struct AllocationQuery {  // gst::query::Allocation
    Caps caps;
    bool need_pool = true;
    std::vector<AllocationPool> pools;
    bool video_meta = false;  // add_allocation_meta::<VideoMeta>
};

// ---- GstVideoFilter stand-in ---------------------------------------------------------
class VideoFilter {
public:
    explicit VideoFilter(int device) : device_(device) {}
    virtual ~VideoFilter();

    virtual const char *type_name() const = 0;    // ObjectSubclass::NAME
    virtual const char *factory_name() const = 0; // gst::Element::register name
    virtual const ElementMetadata &metadata() const = 0;
    virtual const std::vector<PadTemplate> &pad_templates() const = 0;
    virtual const std::vector<ParamSpec> &properties() const = 0;
    virtual BaseTransformMode mode() const = 0;

    // g_object_set / g_object_get.  Unknown name → false (the reference: unimplemented!()).
    // A float outside [minimum, maximum] is rejected and the property keeps its value, as
    // g_object_set_property does after g_param_value_validate.
    bool set_property(const std::string &name, const Value &v);
    std::optional<Value> property(const std::string &name) const;

    // BaseTransformImpl::start / stop.  The base creates / destroys the CUDA context.
    virtual ErrorMessage start();
    virtual ErrorMessage stop();

    // BaseTransformImpl::transform_caps; default = same caps both sides (GstVideoFilter).
    virtual Caps transform_caps(PadDirection direction, const Caps &caps, const Caps *filter) const;

    // BaseTransformImpl::set_caps / propose_allocation / decide_allocation / before_transform.
    // The reference elements override none of these (GstVideoFilter defaults: caps recorded,
    // no pool offered, output from the default system-memory allocator).  The system-memory
    // elements here add one thing on top, invisible in caps and properties: when a pool is
    // wanted and the element has a device, it offers (upstream) / uses (for its own output)
    // a pool of page-locked system memory, so frames skip the pageable bounce copy; a peer
    // that ignores the offer gets the reference's behaviour.  The CUDA-memory variants below
    // override all four after d3d12colorlut/imp.rs:349-542.  Non-empty string = LoggableError.
    virtual std::string set_caps(const Caps &incaps, const Caps &outcaps);
    virtual std::string propose_allocation(AllocationQuery &query);
    virtual std::string decide_allocation(AllocationQuery &query);
    virtual void before_transform(const VideoFrameRef &inbuf);
    // set by before_transform when the element moved to another device (reconfigure_src)
    bool take_reconfigure() { bool r = reconfigure_; reconfigure_ = false; return r; }
    int device() const { return device_; }

    // VideoFilterImpl
    virtual FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out);
    virtual FlowReturn transform_frame_ip(VideoFrameRef &frame);

    // Queued operation — BaseTransformImpl::submit_input_buffer / generate_output instead of the
    // transform_frame call inside the default generate_output.  The reference elements are
    // synchronous (their CPU loop is done when transform_frame returns); with page-locked
    // system-memory frames a GPU element finishes buffer k while buffer k+1 is already uploading
    // if it may hold `frames` buffers back (it then reports that latency).  frames = 0 (default)
    // is the reference's behaviour.  submit_input_frame queues the work ("host.async", include/
    // b200vf.h); generate_output hands out the oldest queued output once more than `frames`
    // are held, complete; drain (EOS, flush, caps change, stop) completes and hands out the rest.
    ErrorMessage set_frames_in_flight(unsigned frames);
    unsigned frames_in_flight() const { return frames_in_flight_; }
    FlowReturn submit_input_frame(const VideoFrameRef &in, VideoFrameRef &out);  // NeverInPlace elements
    FlowReturn submit_input_frame_ip(VideoFrameRef &frame);                      // AlwaysInPlace elements
    enum class GenerateOutput { Buffer, NoOutput, Error };                       // GenerateOutputSuccess
    GenerateOutput generate_output(VideoFrameRef &done);
    GenerateOutput drain(VideoFrameRef &done);  // call until NoOutput

    const std::string &last_error() const { return last_error_; }
    b200vf_ctx *context() const { return ctx_; }

protected:
    virtual bool store(const std::string &name, const Value &v) = 0;
    virtual std::optional<Value> load(const std::string &name) const = 0;
    bool make_frame(const VideoFrameRef &f, b200vf_frame &out);
    FlowReturn flow_error(const std::string &why);

    // shared bodies of the CUDA-memory variants
    std::string cuda_propose_allocation(AllocationQuery &query);
    std::string cuda_decide_allocation(AllocationQuery &query);
    void cuda_before_transform(const VideoFrameRef &inbuf);
    bool require_device_memory(const VideoFrameRef &f);

    int device_;
    b200vf_ctx *ctx_ = nullptr;
    std::string last_error_;
    std::optional<Caps> incaps_, outcaps_;
    bool reconfigure_ = false;
    unsigned frames_in_flight_ = 0;
    std::deque<std::pair<uint64_t, VideoFrameRef>> queued_;  // (ticket, output frame), oldest first

private:
    FlowReturn queued(FlowReturn rc, const VideoFrameRef &out);
    GenerateOutput pop_output(VideoFrameRef &done);
};

// ---- colorlut -------------------------------------------------------------------------
class ColorLut : public VideoFilter {
public:
    using VideoFilter::VideoFilter;
    const char *type_name() const override { return "GstColorLut"; }  // imp.rs:63
    const char *factory_name() const override { return "colorlut"; }  // mod.rs:18-24
    const ElementMetadata &metadata() const override;                 // imp.rs:106-117
    const std::vector<PadTemplate> &pad_templates() const override;   // imp.rs:120-159
    const std::vector<ParamSpec> &properties() const override;        // imp.rs:69-81
    BaseTransformMode mode() const override { return BaseTransformMode::NeverInPlace; }  // :163
    ErrorMessage start() override;                                    // imp.rs:168-194
    ErrorMessage stop() override;                                     // imp.rs:196-199
    FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out) override;  // :203-223

protected:
    bool store(const std::string &name, const Value &v) override;
    std::optional<Value> load(const std::string &name) const override;
    bool lut_loaded();  // State::lut.is_some(), under the state lock

private:
    mutable std::mutex settings_mu_;       // Mutex<Settings>, imp.rs:57
    std::optional<std::string> location_;  // Settings::location
    std::mutex state_mu_;                  // Mutex<State>, imp.rs:58
    bool lut_loaded_ = false;              // State::lut.is_some()
};

// ---- hsvfilter ------------------------------------------------------------------------
class HsvFilter : public VideoFilter {
public:
    using VideoFilter::VideoFilter;
    const char *type_name() const override { return "GstHsvFilter"; }  // imp.rs:69
    const char *factory_name() const override { return "hsvfilter"; }
    const ElementMetadata &metadata() const override;                  // imp.rs:260-271
    const std::vector<PadTemplate> &pad_templates() const override;    // imp.rs:274-312
    const std::vector<ParamSpec> &properties() const override;         // imp.rs:124-161
    BaseTransformMode mode() const override { return BaseTransformMode::AlwaysInPlace; }  // :316
    FlowReturn transform_frame_ip(VideoFrameRef &frame) override;      // imp.rs:323-376

protected:
    bool store(const std::string &name, const Value &v) override;
    std::optional<Value> load(const std::string &name) const override;

private:
    mutable std::mutex settings_mu_;                           // imp.rs:56
    b200vf_hsvfilter_params settings_{0.0f, 1.0f, 0.0f, 1.0f, 0.0f};  // imp.rs:25-29
};

// ---- hsvdetector ----------------------------------------------------------------------
class HsvDetector : public VideoFilter {
public:
    using VideoFilter::VideoFilter;
    const char *type_name() const override { return "GstHsvDetector"; }  // imp.rs:73
    const char *factory_name() const override { return "hsvdetector"; }
    const ElementMetadata &metadata() const override;                    // imp.rs:330-341
    const std::vector<PadTemplate> &pad_templates() const override;      // imp.rs:344-377
    const std::vector<ParamSpec> &properties() const override;           // imp.rs:164-214
    BaseTransformMode mode() const override { return BaseTransformMode::NeverInPlace; }  // :381
    Caps transform_caps(PadDirection direction, const Caps &caps, const Caps *filter) const override;
    FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out) override;  // :423-707

protected:
    bool store(const std::string &name, const Value &v) override;
    std::optional<Value> load(const std::string &name) const override;

private:
    mutable std::mutex settings_mu_;                                            // imp.rs:61
    b200vf_hsvdetector_params settings_{0.0f, 10.0f, 0.0f, 0.15f, 0.0f, 0.3f};  // imp.rs:26-31
};

// ---- CUDA-memory variants ----------------------------------------
// Siblings of the three elements that negotiate `memory:CUDAMemory` caps, in the way the
// reference ships `d3d12colorlut` next to `colorlut` (video/colorlut/src/d3d12colorlut/imp.rs):
// same properties and formats, caps carry the memory feature (:236-266), buffers come from a
// device pool (:385-492), the element follows the device of the incoming memory (:494-542) and
// `transform` only enqueues work, the analogue of the output fence (:711-714).
class CudaColorLut : public ColorLut {
public:
    using ColorLut::ColorLut;
    const char *type_name() const override { return "GstCudaColorLut"; }
    const char *factory_name() const override { return "cudacolorlut"; }
    const ElementMetadata &metadata() const override;
    const std::vector<PadTemplate> &pad_templates() const override;
    std::string set_caps(const Caps &incaps, const Caps &outcaps) override;  // :349-383
    std::string propose_allocation(AllocationQuery &q) override { return cuda_propose_allocation(q); }
    std::string decide_allocation(AllocationQuery &q) override { return cuda_decide_allocation(q); }
    void before_transform(const VideoFrameRef &inbuf) override { cuda_before_transform(inbuf); }
    FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out) override;
};

class CudaHsvFilter : public HsvFilter {
public:
    using HsvFilter::HsvFilter;
    const char *type_name() const override { return "GstCudaHsvFilter"; }
    const char *factory_name() const override { return "cudahsvfilter"; }
    const ElementMetadata &metadata() const override;
    const std::vector<PadTemplate> &pad_templates() const override;
    std::string propose_allocation(AllocationQuery &q) override { return cuda_propose_allocation(q); }
    std::string decide_allocation(AllocationQuery &q) override { return cuda_decide_allocation(q); }
    void before_transform(const VideoFrameRef &inbuf) override { cuda_before_transform(inbuf); }
    FlowReturn transform_frame_ip(VideoFrameRef &frame) override;
};

class CudaHsvDetector : public HsvDetector {
public:
    using HsvDetector::HsvDetector;
    const char *type_name() const override { return "GstCudaHsvDetector"; }
    const char *factory_name() const override { return "cudahsvdetector"; }
    const ElementMetadata &metadata() const override;
    const std::vector<PadTemplate> &pad_templates() const override;
    Caps transform_caps(PadDirection direction, const Caps &caps, const Caps *filter) const override;
    std::string propose_allocation(AllocationQuery &q) override { return cuda_propose_allocation(q); }
    std::string decide_allocation(AllocationQuery &q) override { return cuda_decide_allocation(q); }
    void before_transform(const VideoFrameRef &inbuf) override { cuda_before_transform(inbuf); }
    FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out) override;
};

// ---- plugin registration (lib.rs / mod.rs) ---------------------------------------------
struct PluginDescriptor {  // gst::plugin_define!
    std::string name, description, filename, license, package;
    std::vector<std::string> elements;
};
const std::vector<PluginDescriptor> &plugins();
// gst::ElementFactory::make(name): "colorlut" | "hsvfilter" | "hsvdetector" and their
// "cuda…" variants; rank none.
std::unique_ptr<VideoFilter> element_factory_make(const std::string &factory_name, int device = 0);
// Machine-readable element surface in the shape of docs/plugins/gst_plugins_cache.json.
std::string describe_element_json(const std::string &factory_name);

}  // namespace b200vf
#pragma GCC visibility pop
