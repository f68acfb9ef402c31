// tile_launch_inst.cuh -- body of one tile_launch_*.cu translation unit: define CB200_T and CB200_NT before including.
#include "tile_kernel.cuh"
#include "tile_launch.hpp"

namespace cb200 {

namespace {
constexpr int kMaxDevices = 16;
template <typename K>
cudaError_t set_smem_once(K kern, size_t smem, size_t &seen)
{
    // (the attribute is sticky per function: raise it only when a launch needs more than any launch before it)
    if (smem <= seen) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) seen = smem;
    return e;
}
}  // namespace

template <>
cudaError_t launch_tile_pass_tma<CB200_T, CB200_NT>(const PassParams<CB200_T> &p, const CUtensorMap &tmap, const TmaLaunch &l)
{
    using T = CB200_T;
    using V = typename Vec2<T>::type;
    constexpr int N = CB200_NT;
    static size_t seen_all[kMaxDevices][12] = {};
    size_t *seen = seen_all[l.device & (kMaxDevices - 1)];
    const bool cnd = p.cond_mask != 0;
    // kernel variant: 0 lean, 1 lean + diagonal runs, 2 full.  The diagonal-run path needs the LUT route (no per-tile
    // conditions, 16 amplitudes per thread); anything else falls back to the full variant.
    int var = p.needs_full;
    if (cnd || (var == 1 && 16 * N != (1 << p.t))) var = 2;
    auto go = [&](auto kern, size_t &s) -> cudaError_t {
        cudaError_t e = set_smem_once(kern, l.smem, s);
        if (e != cudaSuccess) return e;
        kern<<<l.grid, N, l.smem, l.stream>>>(p, tmap, (const V *)l.diag_tabs, l.flags, l.box_bits, l.buf_stride, l.refill_after,
                                                (const T *)l.reg_mats, l.reg_mat_stride, l.reg_diag_stride);
        return cudaGetLastError();
    };
    if (p.mat_form != 0) {   // lean pass with widened register blocks (encode_pass): FP64 -> form 1, FP32 -> form 2
        constexpr int kForm = sizeof(T) == 8 ? 1 : 2;
        if (l.reg_mats != nullptr || p.needs_full != 0 || p.mat_form != kForm) return cudaErrorInvalidValue;   // (always two stages)
        if (cnd) return go(tile_pass_tma_kernel<T, N, true, 0, 2, kForm>, seen[7]);
        if (l.stages == 1) return go(tile_pass_tma_kernel<T, N, false, 0, 1, kForm>, seen[8]);   // (experiment: one buffer, more CTAs per SM)
        return go(tile_pass_tma_kernel<T, N, false, 0, 2, kForm>, seen[6]);
    }
    if (l.stages == 3) {
        if (cnd) return go(tile_pass_tma_kernel<T, N, true, 2, 3>, seen[0]);
        return go(tile_pass_tma_kernel<T, N, false, 2, 3>, seen[1]);   // (three stages: experiments only, full variant)
    }
    if (cnd) return go(tile_pass_tma_kernel<T, N, true, 2, 2>, seen[2]);
    if (var == 0) return go(tile_pass_tma_kernel<T, N, false, 0, 2>, seen[3]);
    if (var == 1 && l.stages == 1) return go(tile_pass_tma_kernel<T, N, false, 1, 1>, seen[9]);   // (experiment: one buffer, more CTAs per SM)
    if (var == 1) return go(tile_pass_tma_kernel<T, N, false, 1, 2>, seen[4]);
    return go(tile_pass_tma_kernel<T, N, false, 2, 2>, seen[5]);
}

}  // namespace cb200
