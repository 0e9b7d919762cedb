// vf_launch_table.cu — tabulated element functions: fill and apply (op: vf_ops.cuh TableMapOp).
#include "vf_ops.cuh"

namespace vf {

__global__ void vf_table_fill_kernel(uint32_t *table, uint32_t shift) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // 2^24 threads
    table[blk_index(i)] = i << shift;  // the entry of colour triple i holds pixel i (blocked order)
}

void table_indices(const uint32_t *colours, size_t n, uint32_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = blk_index(colours[i]);
}

cudaError_t launch_table_fill(cudaStream_t stream, uint32_t *table, bool colour_at_1, uint64_t *launches) {
    vf_table_fill_kernel<<<(1u << 24) / 256, 256, 0, stream>>>(table, colour_at_1 ? 8u : 0u);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_table_map(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g, int in_bpp,
                             int out_bpp, const uint32_t *table, bool colour_at_1, bool keep_other,
                             uint64_t *launches) {
    TableMapOp op;
    op.table = table;
    op.idx_sel = colour_at_1 ? 0x4321u : 0x4210u;
    // bytes 0-3 of the result: entry bytes are selectors 0-3, the pixel's own bytes 4-7
    op.out_sel = !keep_other ? 0x3210u : colour_at_1 ? 0x3214u : 0x7210u;
    return launch_map(stream, fs, n, g, in_bpp, out_bpp, op, launches);
}

}  // namespace vf
