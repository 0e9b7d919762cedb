// vf_internal.h — internal interface between the C ABI (vf_abi.cpp, vf_host.cpp), the
// kernels (vf_ops.cuh, vf_launch_*.cu) and the .cube parser (vf_cube_parser.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

struct b200vf_ctx;

namespace vf {

constexpr int kMaxBatch = 64;  // frames per launch (pointer table lives in kernel params)

struct FrameSet {
    const uint8_t *in[kMaxBatch];
    uint8_t *out[kMaxBatch];
};

struct Geom {
    long long in_stride, out_stride;  // bytes
    uint32_t width, height;           // pixels, rows
};

// Byte offsets of the channels inside one pixel; a = -1 when there is no 4th byte.
struct PixLayout {
    int bpp, r, g, b, a;
};

struct HsvFilterArgs {
    float hue_shift, sat_mul, sat_off, val_mul, val_off;
};

struct HsvDetectArgs {
    float hue_ref, hue_var, sat_ref, sat_var, val_ref, val_var;
};

// LUT resident in device memory, laid out for the kernels (see vf_ops.cuh).
struct DeviceLut {
    int kind = 0;       // 0 none, 1 = 1D, 3 = 3D
    uint32_t size = 0;  // N
    float scale[3] = {1, 1, 1}, offset[3] = {0, 0, 0};
    bool identity_domain = true;
    // 3D: (N+1)^3 float4, index x + y*(N+1) + z*(N+1)^2, far edges duplicated so that
    // corner x0+1 is always addressable (equals the reference's min(x0+1, N-1) clamp).
    // Entry = {R(x), R(x+1), G(x), B(x)} (the file's constant fourth lane is dropped).
    float4 *lut3d = nullptr;
    // 3D, R-axis resampled: [z][y][r] for r = 0..255 (8-bit) — x-lerp pre-applied with
    // the reference's own arithmetic.  (N+1)^2 * 256 float4.  Optional.
    float4 *lut3d_rx = nullptr;
    // 3D, R- and G-axis resampled: [z][g][r] for 8-bit codes r, g — x- and y-lerps
    // pre-applied; entry = {R(z), R(z+1), G(z), B(z)}.  (N+1) * 65536 float4 = (N+1) MiB.
    // Optional (built for N <= 71).
    float4 *lut3d_rg = nullptr;
    // 3D, baked to the native 8-bit resolution: [b][g][r] → R'|G'<<8|B'<<16, 2^24 * 4 B = 64 MiB.
    // Every entry is the reference's full trilinear result for that input triple, computed on
    // the device with the direct path.  Opt-in ("lut.path" = 4), built on first use.
    uint32_t *lut3d_baked = nullptr;  // borrowed from the device-wide table cache (vf_tables.cpp)
    int baked_interp = -1;  // LutInterp the baked table was built with
    // every entry finite and within [0,1] ⇒ the output clamp is the identity
    bool unit_range = false;
    // 1D: three planes of N+1 floats (last duplicated).
    float *lut1d = nullptr;
    // 3D, 16-bit frames with an identity domain (N <= 128): 32-byte entries
    // {R, G, dR, dG, B(y), B(y+1), dB(y), dB(y+1)} with dC = RN(C(x+1) - C(x)), at x + y*S + z*S^2,
    // S = 65 or 129 (ColorLut64Op, vf_ops.cuh).
    // Built on the first RGBA64 frame; `coords16_ok` = the op's coordinate arithmetic was checked
    // against the reference formula for all 65536 codes of this size (on the host).
    float *lut3d_d = nullptr;
    int lut3d_d_stride = 0;
    bool coords16_ok = false, sm1_pow2 = false, lut64_failed = false;
    float k16_hi = 0.0f, k16_lo = 0.0f;
};

enum MathMode { kMathFast = 0, kMathPlain = 1 };
enum LutPath { kLutAuto = 0, kLutDirect = 1, kLutResampledR = 2, kLutResampledRG = 3, kLutBaked = 4 };
// 3D interpolation.  Trilinear is the reference (imp.rs:493-526); the other two are extensions
// without a reference counterpart, defined in DESIGN.md §11.
enum LutInterp { kInterpTrilinear = 0, kInterpTetrahedral = 1, kInterpNearest = 2 };

// All launchers enqueue on `stream`, add the number of kernels launched to
// *launches, and return the CUDA status of the launch.
cudaError_t launch_hsvfilter(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                             const PixLayout &lay, const HsvFilterArgs &a, int math_mode,
                             uint64_t *launches);
cudaError_t launch_hsvdetector(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                               const PixLayout &in_lay, const PixLayout &out_lay,
                               const HsvDetectArgs &a, int math_mode, uint64_t *launches);
// bits = 8 (RGBA) or 16 (RGBA64); big_endian only meaningful for 16.
cudaError_t launch_colorlut(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                            int bits, bool big_endian, const DeviceLut &lut, int math_mode,
                            int lut_path, int interp, uint64_t *launches);
cudaError_t launch_chain_lut_hsv(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                                 const DeviceLut &lut, const HsvFilterArgs &a, int lut_path,
                                 int interp, uint64_t *launches);
// colorlut on any 8-bit packed layout in / out, through the baked table (vf_launch_colorlut.cu).
cudaError_t launch_colorlut_convert(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                                    const PixLayout &in_lay, const PixLayout &out_lay,
                                    const uint32_t *baked, uint64_t *launches);
// Which kernel launch_colorlut will pick: 0 direct, 1 R-resampled, 2 1D, 3 RG-resampled, 4 baked,
// 5 tetrahedral, 6 nearest.
int resolved_lut_path(const DeviceLut &lut, int bits, int math_mode, int lut_path, int interp);
// Fills `dst` (2^24 entries, blk_index order) from lut.lut3d with the given interpolation.
cudaError_t launch_build_baked(cudaStream_t stream, const DeviceLut &lut, uint32_t *dst, int interp,
                               uint64_t *launches);
// Fills lut.lut3d_d (allocated, strides lut3d_d_stride and its square) from lut.lut3d.
cudaError_t launch_build_lut64(cudaStream_t stream, const DeviceLut &lut, uint64_t *launches);
// Builds lut.lut3d_rx (when `rx`) / lut.lut3d_rg from lut.lut3d (8-bit input codes); the tables must be allocated.
cudaError_t launch_build_resampled(cudaStream_t stream, const DeviceLut &lut, bool rx, bool rg, uint64_t *launches);

// Tabulated element functions (vf_launch_table.cu).  table = 2^24 uint32.
//   fill: table[i] = i << (colour_at_1 ? 8 : 0) — every colour triple as a 4-byte pixel whose
//         pass-through byte is 0; running an element's kernel in place over it (as a 4096x4096
//         frame) turns it into that element's function table.
//   map : out = table[colour bytes of in], merged with the pixel's own byte where `keep_other`
//         (hsvfilter: alpha / x passes through; hsvdetector: the entry is the whole output pixel).
cudaError_t launch_table_fill(cudaStream_t stream, uint32_t *table, bool colour_at_1, uint64_t *launches);
cudaError_t launch_table_map(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g, int in_bpp,
                             int out_bpp, const uint32_t *table, bool colour_at_1, bool keep_other,
                             uint64_t *launches);

// ---- device-wide cache of 2^24-entry function tables (vf_tables.cpp) ---------------------------
// One entry per (device, key); shared by all contexts that ask for the same key, freed with the
// last reference.  `key` describes the function completely (LUT content hash + interpolation, or
// element + colour-byte placement + every setting bit).
struct SharedTable {
    std::vector<uint8_t> key;
    int device = 0;
    uint32_t *data = nullptr;  // 2^24 entries in blk_index order
    int refs = 0;
    bool built = false;              // the build has been enqueued
    cudaStream_t builder_stream = nullptr;
    cudaEvent_t ready = nullptr;     // recorded after the build on the builder's stream
};
SharedTable *table_acquire(int device, const std::vector<uint8_t> &key);  // nullptr: out of memory
void table_release(SharedTable *t);
// Enqueues build(table) on `stream` if nobody has yet; otherwise orders `stream` after the build.
cudaError_t table_ensure_built(SharedTable *t, cudaStream_t stream,
                               const std::function<cudaError_t(uint32_t *)> &build);
void table_cache_stats(int device, uint64_t *tables, uint64_t *bytes);

// Host evaluation of the tables' index function (vf_ops.cuh blk_index) for the layout tests.
void table_indices(const uint32_t *colours, size_t n, uint32_t *out);

// RGBA pixels (device) → 3 floats (h,s,v) per pixel (device); diagnostics for the tests.
cudaError_t launch_debug_from_rgb(cudaStream_t stream, const uint32_t *px, float *hsv, size_t n,
                                  int plain, uint64_t *launches);

// ---- .cube parser (host) -----------------------------------------------------
struct CubeData {
    int kind = 0;  // 1 or 3
    uint32_t size = 0;
    float domain_scale[3] = {1, 1, 1};
    float domain_offset[3] = {0, 0, 0};
    std::vector<float> data;  // 3D: size^3 * 4 ([r,g,b,1]); 1D: 3 planes of size
};

// Returns 0 ok, 1 = InvalidLut, 2 = Io; message in `err`.
int parse_cube_text(const char *text, size_t len, CubeData &out, std::string &err);
int parse_cube_file(const char *path, CubeData &out, std::string &err);

// Records `msg` as the calling thread's context-free error (b200vf_last_error(NULL)) and
// returns `code`; for entry points that have no context (vf_abi.cpp).
int fail_global(int code, const std::string &msg);
// Same for a context (b200vf_last_error(ctx)).
int ctx_fail(b200vf_ctx *ctx, int code, const std::string &msg);

}  // namespace vf
