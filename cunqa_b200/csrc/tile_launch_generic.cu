// generic (non-TMA) tile-pass kernel launcher; see tile_launch.hpp
#include "tile_kernel.cuh"
#include "tile_launch.hpp"

namespace cb200 {

template <typename T>
cudaError_t launch_tile_pass_generic(const PassParams<T> &p, void *state, const void *diag_tabs, const uint8_t *flags,
                                     int num_sms, int ctas_per_sm, int device, cudaStream_t stream,
                                     const void *reg_mats, uint32_t reg_mat_stride, uint32_t reg_diag_stride)
{
    using V = typename Vec2<T>::type;
    auto kern = tile_pass_kernel<T>;
    const size_t smem = sizeof(V) * ((size_t)1 << p.t) + sizeof(uint64_t) * ((size_t)1 << (p.t - p.L));
    static size_t seen_all[16] = {};
    size_t &seen = seen_all[device & 15];
    if (smem > seen) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        seen = smem;
    }
    int occ = ctas_per_sm;
    if (occ <= 0) {
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTileThreads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    const uint64_t total_tiles = p.tiles_per_state * (uint64_t)p.batch;
    const unsigned grid = (unsigned)(total_tiles < (uint64_t)num_sms * occ ? total_tiles : (uint64_t)num_sms * occ);
    kern<<<grid, kTileThreads, smem, stream>>>(p, (V *)state, (const V *)diag_tabs, flags, (const T *)reg_mats, reg_mat_stride, reg_diag_stride);
    return cudaGetLastError();
}
template cudaError_t launch_tile_pass_generic<double>(const PassParams<double> &, void *, const void *, const uint8_t *, int, int, int, cudaStream_t, const void *, uint32_t, uint32_t);
template cudaError_t launch_tile_pass_generic<float>(const PassParams<float> &, void *, const void *, const uint8_t *, int, int, int, cudaStream_t, const void *, uint32_t, uint32_t);

}  // namespace cb200
