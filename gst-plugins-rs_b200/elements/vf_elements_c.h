/* vf_elements_c.h — C view of the C++ element layer (vf_elements.hpp) so that the Python
 * test / bench harness can drive the elements the way GStreamer would:
 * g_object_set → start → transform_frame[_ip] → stop.  Not part of the drop-in boundary
 * (that is include/b200vf.h); this is the host-side mirror of the reference's elements. */
#ifndef VF_ELEMENTS_C_H
#define VF_ELEMENTS_C_H
#include "../../include/b200vf.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct b200vf_element b200vf_element;

B200VF_API b200vf_element *b200vf_element_new(const char *factory_name, int device); /* NULL if unknown */
B200VF_API void b200vf_element_free(b200vf_element *e);
/* g_object_set / g_object_get; return 1 on success, 0 if unknown / wrong type / out of range */
B200VF_API int b200vf_element_set_float(b200vf_element *e, const char *name, float v);
B200VF_API int b200vf_element_set_string(b200vf_element *e, const char *name, const char *v);
B200VF_API int b200vf_element_get_float(b200vf_element *e, const char *name, float *out);
B200VF_API const char *b200vf_element_get_string(b200vf_element *e, const char *name); /* NULL = unset */
/* start / stop: 0 = ok, else ResourceError (1 Settings, 2 Read, 3 Failed); text via _message */
B200VF_API int b200vf_element_start(b200vf_element *e);
B200VF_API int b200vf_element_stop(b200vf_element *e);
B200VF_API const char *b200vf_element_message(b200vf_element *e);
/* 0 = FlowSuccess::Ok, -5 = FlowError::Error */
B200VF_API int b200vf_element_transform_frame(b200vf_element *e, const b200vf_frame *in,
                                              const b200vf_frame *out);
B200VF_API int b200vf_element_transform_frame_ip(b200vf_element *e, const b200vf_frame *frame);
/* Queued operation (VideoFilter::set_frames_in_flight / submit_input_frame / generate_output /
 * drain).  submit: `out` NULL = in place.  generate_output / drain: 1 = *done holds a completed
 * output frame, 0 = no output (yet / left), -5 = error. */
B200VF_API int b200vf_element_set_frames_in_flight(b200vf_element *e, unsigned frames);
B200VF_API int b200vf_element_submit_input_frame(b200vf_element *e, const b200vf_frame *in,
                                                 const b200vf_frame *out);
B200VF_API int b200vf_element_generate_output(b200vf_element *e, b200vf_frame *done);
B200VF_API int b200vf_element_drain(b200vf_element *e, b200vf_frame *done);
/* formats as comma-separated GstVideoFormat names; filter_csv NULL = no filter caps */
B200VF_API const char *b200vf_element_transform_caps(b200vf_element *e, int direction_is_src,
                                                     const char *formats_csv,
                                                     const char *filter_csv);
/* JSON description of the element surface, comparable with gst_plugins_cache.json */
B200VF_API const char *b200vf_element_describe(const char *factory_name);
B200VF_API void *b200vf_element_context(b200vf_element *e); /* b200vf_ctx* after start */
#ifdef __cplusplus
}
#endif
#endif
