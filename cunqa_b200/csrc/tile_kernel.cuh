// tile_kernel.cuh -- the hot kernel: one in-place tile pass over the state (K1/K2/K3 of SURVEY 8a).
//
// Replaces the per-gate 2^n sweeps that the reference triggers through
// AerState::apply_* (aer_simulator_adapter.cpp:203-533): instead of one HBM round trip per gate,
// a pass stages each 2^t-amplitude tile in shared memory once and applies the pass's whole op
// list (dense k-qubit blocks, diagonal tables, X/SWAP permutations, all optionally controlled and
// optionally masked per state) before writing it back.  Algorithmic bytes per pass = 2*S.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "pass_types.hpp"

namespace cb200 {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// shared-memory slot swizzle: XOR-fold the higher index bits into the bits that select the
// 16-byte bank group, so that sweeps whose lanes differ only in higher tile bits stay conflict free.
template <typename T> __device__ __forceinline__ uint32_t swz(uint32_t i);
template <> __device__ __forceinline__ uint32_t swz<double>(uint32_t i)
{
    return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9) ^ (i >> 12)) & 7u);
}
template <> __device__ __forceinline__ uint32_t swz<float>(uint32_t i)
{
    return i ^ (((i >> 4) ^ (i >> 8) ^ (i >> 12)) & 15u);
}

__device__ __forceinline__ uint32_t insert0(uint32_t g, uint32_t pos)
{
    const uint32_t lo = g & ((1u << pos) - 1u);
    return ((g >> pos) << (pos + 1)) | lo;
}
__device__ __forceinline__ uint64_t insert0_64(uint64_t g, uint32_t pos)
{
    const uint64_t lo = g & ((1ull << pos) - 1ull);
    return ((g >> pos) << (pos + 1)) | lo;
}

template <typename V, typename T>
__device__ __forceinline__ V cmul(const T mr, const T mi, const V a)
{
    V r;
    r.x = mr * a.x - mi * a.y;
    r.y = mr * a.y + mi * a.x;
    return r;
}
template <typename V, typename T>
__device__ __forceinline__ void cfma(V &acc, const T mr, const T mi, const V a)
{
    acc.x = fma(mr, a.x, acc.x);
    acc.x = fma(-mi, a.y, acc.x);
    acc.y = fma(mr, a.y, acc.y);
    acc.y = fma(mi, a.x, acc.y);
}

// ---- dense k-qubit block on the tile -----------------------------------------------------------
// K <= 2: the matrix is hoisted into registers (constant-bank loads, no shared-memory traffic);
// K >= 3: matrix elements stream from the constant bank inside the row loop.
template <typename T, int K>
__device__ __forceinline__ void op_dense(typename Vec2<T>::type *tile, const DevOp &op, const T *mat,
                                         const int t, const int tid)
{
    using V = typename Vec2<T>::type;
    constexpr int D = 1 << K;
    uint32_t off[D];
#pragma unroll
    for (int j = 0; j < D; j++) {
        uint32_t o = 0;
#pragma unroll
        for (int m = 0; m < K; m++)
            if ((j >> m) & 1) o |= 1u << op.tbit[m];
        off[j] = o;
    }
    const int nfix = op.nfix;
    const uint32_t ngroups = 1u << (t - nfix);
    if constexpr (K <= 2) {
        T mr[D * D], mi[D * D];
#pragma unroll
        for (int e = 0; e < D * D; e++) { mr[e] = mat[2 * e]; mi[e] = mat[2 * e + 1]; }
        for (uint32_t g = tid; g < ngroups; g += kTileThreads) {
            uint32_t idx = g;
            for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
            idx |= op.ctrl_in;
            V in[D];
#pragma unroll
            for (int j = 0; j < D; j++) in[j] = tile[swz<T>(idx | off[j])];
#pragma unroll
            for (int r = 0; r < D; r++) {
                V acc = cmul<V, T>(mr[r * D], mi[r * D], in[0]);
#pragma unroll
                for (int c = 1; c < D; c++) cfma<V, T>(acc, mr[r * D + c], mi[r * D + c], in[c]);
                tile[swz<T>(idx | off[r])] = acc;
            }
        }
    } else {
        for (uint32_t g = tid; g < ngroups; g += kTileThreads) {
            uint32_t idx = g;
            for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
            idx |= op.ctrl_in;
            V in[D];
#pragma unroll
            for (int j = 0; j < D; j++) in[j] = tile[swz<T>(idx | off[j])];
#pragma unroll 4
            for (int r = 0; r < D; r++) {
                const T *row = mat + 2 * r * D;
                V acc = cmul<V, T>(row[0], row[1], in[0]);
#pragma unroll
                for (int c = 1; c < D; c++) cfma<V, T>(acc, row[2 * c], row[2 * c + 1], in[c]);
                uint32_t o = 0;
#pragma unroll
                for (int m = 0; m < K; m++)
                    if ((r >> m) & 1) o |= 1u << op.tbit[m];
                tile[swz<T>(idx | o)] = acc;
            }
        }
    }
}

// ---- X / SWAP permutations (optionally controlled): pure data movement ------------------------
template <typename T>
__device__ __forceinline__ void op_perm(typename Vec2<T>::type *tile, const DevOp &op, const int t, const int tid)
{
    using V = typename Vec2<T>::type;
    const int nfix = op.nfix;
    const uint32_t ngroups = 1u << (t - nfix);
    uint32_t oa, ob;
    if (op.kind == OP_PERMX) { oa = 0; ob = 1u << op.tbit[0]; }
    else { oa = 1u << op.tbit[0]; ob = 1u << op.tbit[1]; }
    for (uint32_t g = tid; g < ngroups; g += kTileThreads) {
        uint32_t idx = g;
        for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
        idx |= op.ctrl_in;
        const uint32_t sa = swz<T>(idx | oa), sb = swz<T>(idx | ob);
        const V a = tile[sa], b = tile[sb];
        tile[sa] = b;
        tile[sb] = a;
    }
}

// ---- the pass kernel ---------------------------------------------------------------------------
// grid: persistent CTAs striding over all tiles of all states of the batch.
// dynamic smem: 2^t amplitudes + 2^(t-L) run offsets.
template <typename T>
__global__ void __launch_bounds__(kTileThreads, 2)
tile_pass_kernel(const __grid_constant__ PassParams<T> p, typename Vec2<T>::type *__restrict__ state,
                 const typename Vec2<T>::type *__restrict__ diag_tabs, const uint8_t *__restrict__ flags)
{
    using V = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V *tile = reinterpret_cast<V *>(smem_raw);
    const int t = p.t, L = p.L, n = p.n;
    const uint32_t tile_amps = 1u << t;
    uint64_t *run_off = reinterpret_cast<uint64_t *>(smem_raw + sizeof(V) * tile_amps);
    const int tid = threadIdx.x;
    const uint32_t nruns = 1u << (t - L);
    const uint32_t run_mask = (1u << L) - 1u;

    // offset (in amplitudes) of run r inside a tile: deposit the bits of r at the high tile qubits
    for (uint32_t r = tid; r < nruns; r += kTileThreads) {
        uint64_t o = 0;
        for (int j = L; j < t; j++)
            if ((r >> (j - L)) & 1u) o |= 1ull << p.tile_q[j];
        run_off[r] = o;
    }
    __syncthreads();

    const uint64_t total_tiles = p.tiles_per_state * (uint64_t)p.batch;
    for (uint64_t tileno = blockIdx.x; tileno < total_tiles; tileno += gridDim.x) {
        const uint64_t b = tileno / p.tiles_per_state;
        const uint64_t tin = tileno - b * p.tiles_per_state;
        // base amplitude index inside the state: zeros inserted at the tile-qubit positions
        uint64_t base = tin;
        for (int j = L; j < t; j++) base = insert0_64(base, p.tile_q[j] - L);
        base <<= L;  // tile qubits 0..L-1 are the L lowest physical qubits

        // which ops act on this tile (controls outside the tile, per-state masks)
        uint64_t active = 0;
        for (int o = 0; o < p.n_ops; o++) {
            const DevOp &op = p.ops[o];
            bool on = (base & op.ctrl_out) == op.ctrl_out;
            if (on && op.has_mask) on = flags[(size_t)op.mask_row * p.batch + b] != 0;
            if (on) active |= 1ull << o;
        }
        if (active == 0) continue;  // untouched tile: no traffic at all

        V *gbase = state + (b << n) + base;
        // ---- stage in: 128-bit coalesced loads, swizzled shared-memory stores
#pragma unroll 4
        for (uint32_t i = tid; i < tile_amps; i += kTileThreads)
            tile[swz<T>(i)] = gbase[run_off[i >> L] + (i & run_mask)];
        __syncthreads();

        for (int o = 0; o < p.n_ops; o++) {
            if (!((active >> o) & 1ull)) continue;
            const DevOp &op = p.ops[o];
            if (op.kind == OP_DENSE) {
                const T *mat = p.mats + op.mat_off;
                switch (op.k) {
                case 1: op_dense<T, 1>(tile, op, mat, t, tid); break;
                case 2: op_dense<T, 2>(tile, op, mat, t, tid); break;
                case 3: op_dense<T, 3>(tile, op, mat, t, tid); break;
                case 4: op_dense<T, 4>(tile, op, mat, t, tid); break;
                default: op_dense<T, 5>(tile, op, mat, t, tid); break;
                }
                __syncthreads();
            } else if (op.kind == OP_DIAG) {
                // a run of consecutive diagonal ops is applied in ONE sweep of the tile
                int oe = o;
                while (oe + 1 < p.n_ops && p.ops[oe + 1].kind == OP_DIAG) oe++;
                for (uint32_t i = tid; i < tile_amps; i += kTileThreads) {
                    const uint32_t s = swz<T>(i);
                    V v = tile[s];
                    for (int q = o; q <= oe; q++) {
                        if (!((active >> q) & 1ull)) continue;
                        const DevOp &d = p.ops[q];
                        uint32_t ti = 0;
                        for (int m = 0; m < d.k; m++) {
                            const uint32_t src = d.dsrc[m];
                            const uint32_t bit = src < 64 ? ((i >> src) & 1u) : (uint32_t)((base >> (src - 64)) & 1ull);
                            ti |= bit << m;
                        }
                        const V ph = diag_tabs[d.mat_off + ti];
                        v = cmul<V, T>(ph.x, ph.y, v);
                    }
                    tile[s] = v;
                }
                o = oe;
                __syncthreads();
            } else {
                op_perm<T>(tile, op, t, tid);
                __syncthreads();
            }
        }

        // ---- stage out
#pragma unroll 4
        for (uint32_t i = tid; i < tile_amps; i += kTileThreads)
            gbase[run_off[i >> L] + (i & run_mask)] = tile[swz<T>(i)];
        __syncthreads();
    }
}

}  // namespace cb200
