// tile_launch.hpp -- host-callable launchers of the tile-pass kernels (tile_kernel.cuh).
//
// The TMA kernel is instantiated per (amplitude type, threads per CTA) in its own translation unit
// (tile_launch_inst.cuh included by tile_launch_{f64,f32}_{32,64,128,256}.cu) so that the variants compile side by side.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "pass_types.hpp"

namespace cb200 {

struct TmaLaunch {
    unsigned grid;
    size_t smem;
    cudaStream_t stream;
    const void *diag_tabs;     // Vec2<T>::type elements
    const uint8_t *flags;
    int box_bits;
    uint32_t buf_stride;
    int refill_after;
    int stages;                // 2 or 3
    int device;                // CUDA device of the launch (function attributes are per device)
    // independent registers batched in one launch: state b of the batch uses row b of these (nullptr / 0 = one shared set)
    const void *reg_mats = nullptr;   // T scalars, reg_mat_stride per register
    uint32_t reg_mat_stride = 0;
    uint32_t reg_diag_stride = 0;     // elements of the diagonal-table buffer per register
};

// NT threads per CTA; the (COND, FULL) kernel variant is picked from p.cond_mask / p.needs_full
template <typename T, int NT>
cudaError_t launch_tile_pass_tma(const PassParams<T> &p, const CUtensorMap &tmap, const TmaLaunch &l);

// generic (non-TMA) fallback
template <typename T>
cudaError_t launch_tile_pass_generic(const PassParams<T> &p, void *state, const void *diag_tabs, const uint8_t *flags,
                                     int num_sms, int ctas_per_sm, int device, cudaStream_t stream,
                                     const void *reg_mats = nullptr, uint32_t reg_mat_stride = 0, uint32_t reg_diag_stride = 0);

}  // namespace cb200
