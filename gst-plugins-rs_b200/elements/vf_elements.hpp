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
};

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
};

// gst_video::VideoFrameRef: plane 0 of a mapped frame.
struct VideoFrameRef {
    void *data = nullptr;
    int64_t stride = 0;
    uint32_t width = 0, height = 0;
    std::string format;  // GstVideoFormat name, e.g. "RGBA"
    b200vf_memory memory = B200VF_MEM_HOST;
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

    // VideoFilterImpl
    virtual FlowReturn transform_frame(const VideoFrameRef &in, VideoFrameRef &out);
    virtual FlowReturn transform_frame_ip(VideoFrameRef &frame);

    const std::string &last_error() const { return last_error_; }
    b200vf_ctx *context() const { return ctx_; }

protected:
    virtual bool store(const std::string &name, const Value &v) = 0;
    virtual std::optional<Value> load(const std::string &name) const = 0;
    bool make_frame(const VideoFrameRef &f, b200vf_frame &out);
    FlowReturn flow_error(const std::string &why);

    int device_;
    b200vf_ctx *ctx_ = nullptr;
    std::string last_error_;
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

// ---- plugin registration (lib.rs / mod.rs) ---------------------------------------------
struct PluginDescriptor {  // gst::plugin_define!
    std::string name, description, filename, license, package;
    std::vector<std::string> elements;
};
const std::vector<PluginDescriptor> &plugins();
// gst::ElementFactory::make(name): "colorlut" | "hsvfilter" | "hsvdetector"; rank none.
std::unique_ptr<VideoFilter> element_factory_make(const std::string &factory_name, int device = 0);
// Machine-readable element surface in the shape of docs/plugins/gst_plugins_cache.json.
std::string describe_element_json(const std::string &factory_name);

}  // namespace b200vf
#pragma GCC visibility pop
