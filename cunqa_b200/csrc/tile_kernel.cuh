// tile_kernel.cuh -- the hot kernel: one in-place tile pass over the state (K1/K2/K3 of SURVEY 8a).
//
// Replaces the per-gate 2^n sweeps that the reference triggers through
// AerState::apply_* (aer_simulator_adapter.cpp:203-533): instead of one HBM round trip per gate,
// a pass stages each 2^t-amplitude tile in shared memory once and applies the pass's whole op
// list (dense k-qubit blocks, diagonal tables, X/SWAP permutations, all optionally controlled and
// optionally masked per state) before writing it back.  Algorithmic bytes per pass = 2*S.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "pass_types.hpp"

namespace cb200 {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// shared-memory slot swizzle = the TMA hardware's 128-byte swizzle (CU_TENSOR_MAP_SWIZZLE_128B):
// the 16-byte chunk index inside a 128-byte row (byte-address bits 4..6) is XOR-ed with the row
// index mod 8 (byte-address bits 7..9).  Tile buffers are 1024-byte aligned, so the pattern is a
// pure function of the tile index; sweeps whose lanes differ in tile bits 3..5 (fp64) stay
// bank-conflict free.  Both kernels (TMA-staged and generic) use the same layout.
template <typename T> __device__ __forceinline__ uint32_t swz(uint32_t i);
template <> __device__ __forceinline__ uint32_t swz<double>(uint32_t i)   // 16-byte amplitudes
{
    return i ^ ((i >> 3) & 7u);
}
template <> __device__ __forceinline__ uint32_t swz<float>(uint32_t i)    // 8-byte amplitudes
{
    return i ^ (((i >> 4) & 7u) << 1);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// ---- dense gates of <= 2 qubits: complex products with three multiplications --------------------
// Matrix element e = c + i d of such a gate is stored as  mat[2e] = c,  mat[2e+1] = -(c+d),  mat[2E+e] = d-c  (E = 4^K
// elements; planner.cpp: pack_dense).  With s_j = re_j + im_j shared by all rows of the gate,
//     re_i = sum_j c_ij s_j + sum_j [-(c+d)_ij] im_j        im_i = sum_j c_ij s_j + sum_j (d-c)_ij re_j
// i.e. 13 FP64 instructions per amplitude of a 2-qubit gate instead of 16 (the kernel is FP64-pipe bound on these).  Two
// rows of the matrix are live at a time (24 values): read from the kernel parameters they stay in uniform registers
// (LDCU -> UR operands of DFMA: no vector registers, no shared-memory wavefronts), read from shared memory (per-register
// matrices, conditioned passes) they cost 12 128-bit loads per row pair.
template <typename T, int K>
__device__ __forceinline__ void rows2_mul3(const T *mat, const int r0, const typename Vec2<T>::type (&in)[1 << K], const T (&s)[1 << K],
                                           typename Vec2<T>::type (&out)[2])
{
    using V = typename Vec2<T>::type;
    constexpr int D = 1 << K, E = D * D;
    const V *mv = reinterpret_cast<const V *>(mat);   // (c, -(c+d)) pairs, then the (d-c) values two by two
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int r = r0 + rr;
        V cn[D], dm[D / 2];
#pragma unroll
        for (int c = 0; c < D; c++) cn[c] = mv[r * D + c];
#pragma unroll
        for (int c = 0; c < D / 2; c++) dm[c] = mv[E + (r * D) / 2 + c];
        T tt = cn[0].x * s[0];
#pragma unroll
        for (int c = 1; c < D; c++) tt = fma(cn[c].x, s[c], tt);
        T re = tt, im = tt;
#pragma unroll
        for (int c = 0; c < D; c++) {
            re = fma(cn[c].y, in[c].y, re);
            im = fma((c & 1) ? dm[c / 2].y : dm[c / 2].x, in[c].x, im);
        }
        out[rr].x = re;
        out[rr].y = im;
    }
}

// ---- dense k-qubit block on the tile -----------------------------------------------------------
// K <= 2: three-multiplication form when the pass stores its matrices that way (M3: PassParams::mat_form, FP64 passes
// served by the lean kernel, matrices read from the kernel parameters) and the gate does not belong to a register sweep
// (DevOp::k2); otherwise plain (re, im) elements hoisted into registers, one 128-bit load each.  K >= 3: plain matrix
// elements streamed in inside the row loop.
template <typename T, int K, int NT, bool M3>
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
    if constexpr (K <= 2 && M3) {
        if (!op.k2) {
            for (uint32_t g = tid; g < ngroups; g += NT) {
                uint32_t idx = g;
                for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
                idx |= op.ctrl_in;
                V in[D];
                T s[D];
#pragma unroll
                for (int j = 0; j < D; j++) {
                    in[j] = tile[swz<T>(idx | off[j])];
                    s[j] = in[j].x + in[j].y;
                }
#pragma unroll
                for (int r0 = 0; r0 < D; r0 += 2) {
                    V out[2];
                    rows2_mul3<T, K>(mat, r0, in, s, out);
                    tile[swz<T>(idx | off[r0])] = out[0];
                    tile[swz<T>(idx | off[r0 + 1])] = out[1];
                }
            }
            return;
        }
    }
    if constexpr (K <= 2) {
        T mr[D * D], mi[D * D];
#pragma unroll
        for (int e = 0; e < D * D; e++) { const V m = reinterpret_cast<const V *>(mat)[e]; mr[e] = m.x; mi[e] = m.y; }  // one 16-byte load per element
        for (uint32_t g = tid; g < ngroups; g += NT) {
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
        for (uint32_t g = tid; g < ngroups; g += NT) {
            uint32_t idx = g;
            for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
            idx |= op.ctrl_in;
            V in[D];
#pragma unroll
            for (int j = 0; j < D; j++) in[j] = tile[swz<T>(idx | off[j])];
#pragma unroll 4
            for (int r = 0; r < D; r++) {
                const V *row = reinterpret_cast<const V *>(mat) + r * D;  // one 16-byte load per element
                const V m0 = row[0];
                V acc = cmul<V, T>(m0.x, m0.y, in[0]);
#pragma unroll
                for (int c = 1; c < D; c++) { const V mc = row[c]; cfma<V, T>(acc, mc.x, mc.y, in[c]); }
                uint32_t o = 0;
#pragma unroll
                for (int m = 0; m < K; m++)
                    if ((r >> m) & 1) o |= 1u << op.tbit[m];
                tile[swz<T>(idx | o)] = acc;
            }
        }
    }
}

// explicit shared-space accesses with 32-bit addresses (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ double2 lds_amp(uint32_t addr, double2)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_amp(uint32_t addr, float2)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_amp(uint32_t addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sts_amp(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// ---- register block: two dense gates per shared-memory round trip ---------------------------------
// The thread holds the 2^R amplitudes spanned by the block's R local bits in registers; gate A acts on
// local bits [0,K1), gate B on local bits [S,S+K2).  All register indices are compile-time.
// the same on plain (re, im) matrix elements hoisted into registers: the dense gates of a register sweep (op_sweep), whose
// matrices come from shared memory and whose register budget is spent on the tables of the sweep
template <typename T, int R, int K, int S>
__device__ __forceinline__ void block_gate_plain(typename Vec2<T>::type (&a)[1 << R], const T *mat)
{
    using V = typename Vec2<T>::type;
    constexpr int D = 1 << K;
    T mr[D * D], mi[D * D];
#pragma unroll
    for (int e = 0; e < D * D; e++) { const V m = reinterpret_cast<const V *>(mat)[e]; mr[e] = m.x; mi[e] = m.y; }  // one 16-byte load per element
#pragma unroll
    for (int rest = 0; rest < (1 << (R - K)); rest++) {
        const int lo = rest & ((1 << S) - 1), hi = rest >> S;
        const int base = (hi << (S + K)) | lo;
        V in[D];
#pragma unroll
        for (int j = 0; j < D; j++) in[j] = a[base | (j << S)];
#pragma unroll
        for (int r = 0; r < D; r++) {
            V acc = cmul<V, T>(mr[r * D], mi[r * D], in[0]);
#pragma unroll
            for (int c = 1; c < D; c++) cfma<V, T>(acc, mr[r * D + c], mi[r * D + c], in[c]);
            a[base | (r << S)] = acc;
        }
    }
}

template <typename T, int R, int K, int S>
__device__ __forceinline__ void block_gate(typename Vec2<T>::type (&a)[1 << R], const T *mat)
{
    using V = typename Vec2<T>::type;
    constexpr int D = 1 << K, G = 1 << (R - K), NCH = D / 2;
    static_assert(K == 1 || K == 2, "register blocks hold 1- and 2-qubit gates");
    // two matrix rows at a time over all groups: rows 0,1 of a 2-qubit gate wait in `keep` until rows 2,3 are known
    V keep[NCH > 1 ? G : 1][2];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
        for (int rest = 0; rest < G; rest++) {
            const int lo = rest & ((1 << S) - 1), hi = rest >> S;
            const int base = (hi << (S + K)) | lo;
            V in[D];
            T s[D];
#pragma unroll
            for (int j = 0; j < D; j++) { in[j] = a[base | (j << S)]; s[j] = in[j].x + in[j].y; }
            V out[2];
            rows2_mul3<T, K>(mat, 2 * ch, in, s, out);
            if (ch + 1 < NCH) { keep[rest][0] = out[0]; keep[rest][1] = out[1]; }
            else {
                a[base | ((2 * ch) << S)] = out[0];
                a[base | ((2 * ch + 1) << S)] = out[1];
                if (NCH > 1) { a[base] = keep[rest][0]; a[base | (1 << S)] = keep[rest][1]; }
            }
        }
    }
}

// RW = 4: the block is widened to four local bits with spectator bits (local bits the gates do not touch; the host picks
// free tile bits): every thread then holds 16 amplitudes and makes one trip whatever the gates' widths, and K2 = 0 is a
// single gate.
template <typename T, int K1, int K2, int S, int NT, bool M3, int RW = 0>
__device__ __forceinline__ void op_pair(typename Vec2<T>::type *tile, const DevOp &op, const T *mat,
                                        const int t, const int tid)
{
    using V = typename Vec2<T>::type;
    constexpr int R = RW ? RW : ((K1 > S + K2) ? K1 : (S + K2));
    constexpr int D = 1 << R;
    // the swizzle is XOR-linear and the block bits are disjoint from the group bits, so
    // swz(idx | off_j) = swz(idx) ^ swz(off_j): one XOR per amplitude instead of re-swizzling
    constexpr int kShift = sizeof(V) == 16 ? 4 : 3;
    const uint32_t tile_s = smem_u32(tile);
    uint32_t sb[R], seg[R + 1];
#pragma unroll
    for (int m = 0; m < R; m++) sb[m] = (uint32_t)op.sboff[m] << kShift;   // sboff[m] = swz<T>(1u << op.tbit[m]) (encode_pass)
#pragma unroll
    for (int m = 0; m <= R; m++) seg[m] = op.seg[m];
    uint32_t boff[D];  // byte offset of block element j relative to the group's element 0 (XOR-combined)
#pragma unroll
    for (int j = 0; j < D; j++) {
        uint32_t o = 0;
#pragma unroll
        for (int m = 0; m < R; m++) if ((j >> m) & 1) o ^= sb[m];
        boff[j] = o;
    }
    const uint32_t ngroups = 1u << (t - R);
    for (uint32_t g = tid; g < ngroups; g += NT) {
        uint32_t idx = 0;
#pragma unroll
        for (int m = 0; m <= R; m++) idx |= (g & seg[m]) << m;
        const uint32_t bidx = swz<T>(idx) << kShift;
        uint32_t addr[D];
        V a[D];
#pragma unroll
        for (int j = 0; j < D; j++) { addr[j] = tile_s + (bidx ^ boff[j]); a[j] = lds_amp(addr[j], V()); }
        // (gate B's matrix sits 3 * 4^K1 scalars after gate A's in either form: planner.cpp, dense_scalars)
        if constexpr (M3) {
            block_gate<T, R, K1, 0>(a, mat);
            if constexpr (K2 > 0) block_gate<T, R, K2, S>(a, mat + 3 * (1 << (2 * K1)));
        } else {
            block_gate_plain<T, R, K1, 0>(a, mat);
            if constexpr (K2 > 0) block_gate_plain<T, R, K2, S>(a, mat + 3 * (1 << (2 * K1)));
        }
#pragma unroll
        for (int j = 0; j < D; j++) sts_amp(addr[j], a[j]);
    }
}

// FORM (PassParams::mat_form): 0 = blocks at their natural width, plain matrices (every kernel variant; what per-register
// matrices, conditioned passes and the full variant use); 1 / 2 = lean passes without per-tile conditions: blocks
// widened to four local bits, matrices in the three-multiplication form read from the kernel parameters (1, FP64) or plain
// from shared memory (2, FP32).
template <typename T, int NT, int VAR, int FORM>
__device__ __forceinline__ void op_pair_dispatch(typename Vec2<T>::type *tile, const DevOp &op, const T *mat,
                                                 const int t, const int tid)
{
    const int code = (op.k << 4) | (op.k2 << 2) | op.s;
    if constexpr (FORM != 0) {
        constexpr bool M3 = FORM == 1;
        if (code == ((2 << 4) | (2 << 2) | 2)) { op_pair<T, 2, 2, 2, NT, M3, 4>(tile, op, mat, t, tid); return; }
        if (code == ((2 << 4) | (2 << 2) | 1)) { op_pair<T, 2, 2, 1, NT, M3, 4>(tile, op, mat, t, tid); return; }
        if (code == ((2 << 4) | (0 << 2) | 0)) { op_pair<T, 2, 0, 0, NT, M3, 4>(tile, op, mat, t, tid); return; }
        op_pair<T, 1, 1, 1, NT, M3, 4>(tile, op, mat, t, tid);
    } else {
        if (code == ((2 << 4) | (2 << 2) | 2)) { op_pair<T, 2, 2, 2, NT, false>(tile, op, mat, t, tid); return; }
        if (code == ((2 << 4) | (2 << 2) | 1)) { op_pair<T, 2, 2, 1, NT, false>(tile, op, mat, t, tid); return; }
        if (code == ((1 << 4) | (1 << 2) | 1)) { op_pair<T, 1, 1, 1, NT, false>(tile, op, mat, t, tid); return; }
        if constexpr (VAR == 2) {
            switch (code) {
            case (1 << 4) | (1 << 2) | 0: op_pair<T, 1, 1, 0, NT, false>(tile, op, mat, t, tid); break;
            case (1 << 4) | (2 << 2) | 1: op_pair<T, 1, 2, 1, NT, false>(tile, op, mat, t, tid); break;
            case (1 << 4) | (2 << 2) | 0: op_pair<T, 1, 2, 0, NT, false>(tile, op, mat, t, tid); break;
            case (2 << 4) | (1 << 2) | 2: op_pair<T, 2, 1, 2, NT, false>(tile, op, mat, t, tid); break;
            case (2 << 4) | (1 << 2) | 1: op_pair<T, 2, 1, 1, NT, false>(tile, op, mat, t, tid); break;
            default: op_pair<T, 2, 2, 0, NT, false>(tile, op, mat, t, tid); break;
            }
        }
    }
}

// ---- X / SWAP permutations (optionally controlled): pure data movement ------------------------
template <typename T, int NT>
__device__ __forceinline__ void op_perm(typename Vec2<T>::type *tile, const DevOp &op, const int t, const int tid)
{
    using V = typename Vec2<T>::type;
    const int nfix = op.nfix;
    const uint32_t ngroups = 1u << (t - nfix);
    uint32_t oa, ob;
    if (op.kind == OP_PERMX) { oa = 0; ob = 1u << op.tbit[0]; }
    else { oa = 1u << op.tbit[0]; ob = 1u << op.tbit[1]; }
    for (uint32_t g = tid; g < ngroups; g += NT) {
        uint32_t idx = g;
        for (int f = 0; f < nfix; f++) idx = insert0(idx, op.fixpos[f]);
        idx |= op.ctrl_in;
        const uint32_t sa = swz<T>(idx | oa), sb = swz<T>(idx | ob);
        const V a = tile[sa], b = tile[sb];
        tile[sa] = b;
        tile[sb] = a;
    }
}

// ---- a run of X / CX / SWAP gates as one permutation of the tile (OP_LINPERM, pass_types.hpp) ----------------------
// Thread g owns the destinations d = g | j << (t-4), j < 16 (lanes on the low tile bits: the stores of a warp hit distinct
// banks); the slot of the source of d is lin[15] ^ XOR of lin[i] over the bits i set in d.  All loads, a barrier, all
// stores.  A thread holds 16 amplitudes per trip; FP32 tiles of 2^13 amplitudes take two trips' worth of registers.
template <typename T, int NT>
__device__ __forceinline__ void op_linperm(typename Vec2<T>::type *tile, const DevOp &op, const int t, const int tid)
{
    using V = typename Vec2<T>::type;
    constexpr int TRIPS = sizeof(T) == 4 ? 2 : 1;
    const int hb = t - 4;                       // tile bits [hb, t) count the thread's 16 amplitudes
    const uint32_t ngroups = 1u << hb;
    uint32_t jc[16];
#pragma unroll
    for (int j = 0; j < 16; j++)
        jc[j] = ((j & 1) ? (uint32_t)op.lin[hb] : 0u) ^ ((j & 2) ? (uint32_t)op.lin[hb + 1] : 0u) ^ ((j & 4) ? (uint32_t)op.lin[hb + 2] : 0u) ^
                ((j & 8) ? (uint32_t)op.lin[hb + 3] : 0u);
    // (threads past the last group repeat it: they load and store the same values as its owner, so the amplitudes stay in
    // registers without a guard around the barrier)
    V v[TRIPS][16];
#pragma unroll
    for (int tr = 0; tr < TRIPS; tr++) {
        const uint32_t g = min((uint32_t)tid + (uint32_t)tr * NT, ngroups - 1u);
        uint32_t sb = op.lin[15];
        for (int i = 0; i < hb; i++) sb ^= ((g >> i) & 1u) ? (uint32_t)op.lin[i] : 0u;
#pragma unroll
        for (int j = 0; j < 16; j++) v[tr][j] = tile[sb ^ jc[j]];
    }
    __syncthreads();
#pragma unroll
    for (int tr = 0; tr < TRIPS; tr++) {
        const uint32_t g = min((uint32_t)tid + (uint32_t)tr * NT, ngroups - 1u);
#pragma unroll
        for (int j = 0; j < 16; j++) tile[swz<T>(g | ((uint32_t)j << hb))] = v[tr][j];
    }
}

// ---- register sweeps: dense gates, diagonal runs and the phase fan on the thread's resident amplitudes ------------
// Each thread owns 16 amplitudes: the 2^4 settings of four "register bits" rb[0..3] of the tile index, chosen per sweep
// by the host (encode_pass).  Between ONE load and ONE store of the tile the sweep applies a short program:
//   * dense gates of <= 2 qubits whose targets are register bits (the host places a 2-qubit gate on register bits
//     (0,1) or (2,3), a 1-qubit gate on any of the four): block_gate on compile-time register indices;
//   * runs of diagonal tables: the in-tile part of a table index is a bit-extract of the tile index, additive over
//     disjoint bit sets, so a table none of whose qubits is a register bit costs ONE lookup and one complex multiply per
//     thread (folded into `acc`, applied once per run); tables that do depend on register bits are looked up per
//     amplitude, table by table so that the loads of one table are in flight together;
//   * the phase fan (below).
// A QFT pass -- H / ladder blocks on six qubits -- is two such sweeps: [dense, tables, dense, tables] and [dense, fan].
// (Not inlined: the sweep's register allocation must not disturb the dense-gate paths of the kernel.)
#ifndef CB200_SWEEP_INLINE
#define CB200_SWEEP_INLINE __noinline__
#endif
template <typename T> struct FanCtx {
    const typename Vec2<T>::type *fanT, *fanc;   // T[c][2^m] and this tile's center factors (shared memory)
};

template <typename T, int NT>
__device__ CB200_SWEEP_INLINE void op_sweep(typename Vec2<T>::type *tile, const PassParams<T> &p, const int sw, const T *mats,
                                      const typename Vec2<T>::type *__restrict__ diag_tabs,
                                      const typename Vec2<T>::type *const *tptr, const uint16_t *lut, const FanCtx<T> fan, const int tid)
{
    using V = typename Vec2<T>::type;
    constexpr int A = 16;
    const typename PassParams<T>::Sweep &S = p.sweeps[sw];
    uint32_t bi = tid;
#pragma unroll
    for (int f = 0; f < 4; f++) bi = insert0(bi, S.asc[f]);
    uint32_t rbit[4];
#pragma unroll
    for (int b = 0; b < 4; b++) rbit[b] = 1u << S.rb[b];
    V v[A];
    // tile index of amplitude k (k is a compile-time constant wherever this is called: at most three ORs)
    auto ti = [&](int k) -> uint32_t {
        return bi | ((k & 1) ? rbit[0] : 0u) | ((k & 2) ? rbit[1] : 0u) | ((k & 4) ? rbit[2] : 0u) | ((k & 8) ? rbit[3] : 0u);
    };
#pragma unroll
    for (int k = 0; k < A; k++) v[k] = tile[swz<T>(ti(k))];
    const int o_end = S.first + S.len;
    for (int o = S.first; o < o_end; o++) {
        const DevOp &op = p.ops[o];
        if (op.kind == OP_DENSE) {
            const T *mat = mats + op.mat_off;
            if (op.k == 2) {
                if (op.s == 0) block_gate_plain<T, 4, 2, 0>(v, mat);
                else block_gate_plain<T, 4, 2, 2>(v, mat);
            } else {
                switch (op.s) {
                case 0: block_gate_plain<T, 4, 1, 0>(v, mat); break;
                case 1: block_gate_plain<T, 4, 1, 1>(v, mat); break;
                case 2: block_gate_plain<T, 4, 1, 2>(v, mat); break;
                default: block_gate_plain<T, 4, 1, 3>(v, mat); break;
                }
            }
        } else if (op.kind == OP_DIAG) {
            // the run of consecutive tables that starts here: slots [s0, s1], the first fixpos[1] of them ignore the register bits
            int oe = o;
            while (oe + 1 < o_end && p.ops[oe + 1].kind == OP_DIAG) oe++;
            const int s0 = op.k2, s1 = p.ops[oe].k2, s_indep = s0 + op.fixpos[1];
            auto index_of = [&](int s) -> uint32_t {
                const uint16_t *L = lut + s * kDiagLut;
                return (uint32_t)(L[bi & 63u] | L[64 + (bi >> 6)]);
            };
            V acc;
            acc.x = (T)1;
            acc.y = (T)0;
            int sl = s0;
            for (; sl + 4 <= s_indep; sl += 4) {
                uint32_t ix[4];
                V q[4];
#pragma unroll
                for (int j = 0; j < 4; j++) ix[j] = index_of(sl + j);
#pragma unroll
                for (int j = 0; j < 4; j++) q[j] = tptr[sl + j][ix[j]];
                const V q01 = cmul<V, T>(q[0].x, q[0].y, q[1]), q23 = cmul<V, T>(q[2].x, q[2].y, q[3]);
                acc = cmul<V, T>(acc.x, acc.y, cmul<V, T>(q01.x, q01.y, q23));
            }
            for (; sl < s_indep; sl++) {
                const V q = tptr[sl][index_of(sl)];
                acc = cmul<V, T>(q.x, q.y, acc);
            }
            for (; sl <= s1; sl++) {
                const uint16_t *L = lut + sl * kDiagLut;
                const V *tb = tptr[sl] + index_of(sl);
                uint32_t e[4];
#pragma unroll
                for (int b = 0; b < 4; b++) e[b] = L[rbit[b] & 63u] | L[64 + (rbit[b] >> 6)];  // one of the two is entry 0 == 0
#pragma unroll
                for (int h = 0; h < A; h += 8) {  // two batches of 8 independent loads (register budget)
                    V q[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int k = h + j;
                        q[j] = tb[((k & 1) ? e[0] : 0u) | ((k & 2) ? e[1] : 0u) | ((k & 4) ? e[2] : 0u) | ((k & 8) ? e[3] : 0u)];
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) v[h + j] = cmul<V, T>(q[j].x, q[j].y, v[h + j]);
                }
            }
            if (s_indep > s0) {
#pragma unroll
                for (int k = 0; k < A; k++) v[k] = cmul<V, T>(acc.x, acc.y, v[k]);
            }
            o = oe;
        } else if (op.kind == OP_FAN) {
            // amplitude *= g * prod over the centers set in its index of [c_b * T_b[leaf bits]]: the leaf bits and the
            // centers that are not register bits are thread constants; the four register-bit factors give 16 subset products
            const int m = op.nfix, nc = op.k;
            uint32_t lv = 0;
            for (int i = 0; i < m; i++) lv |= ((bi >> op.fixpos[i]) & 1u) << i;
            auto factor = [&](int c) -> V { const V a = fan.fanc[c], b = fan.fanT[(c << m) + lv]; return cmul<V, T>(a.x, a.y, b); };
            // which center (if any) each register bit is (host: dsrc[4..7])
            int rc[4];
#pragma unroll
            for (int b = 0; b < 4; b++) rc[b] = op.dsrc[4 + b] == 0xFF ? -1 : (int)op.dsrc[4 + b];
            V s = diag_tabs[op.mat_off];
            for (int c = 0; c < nc; c++)
                if (c != rc[0] && c != rc[1] && c != rc[2] && c != rc[3] && ((bi >> op.tbit[c]) & 1u)) { const V f = factor(c); s = cmul<V, T>(s.x, s.y, f); }
            V f[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                f[b].x = (T)1; f[b].y = (T)0;
                if (rc[b] >= 0) f[b] = factor(rc[b]);
            }
            V g01[4], g23[4];
            g01[0] = s;
            g01[1] = cmul<V, T>(s.x, s.y, f[0]);
            g01[2] = cmul<V, T>(s.x, s.y, f[1]);
            g01[3] = cmul<V, T>(g01[1].x, g01[1].y, f[1]);
            g23[1] = f[2];
            g23[2] = f[3];
            g23[3] = cmul<V, T>(f[2].x, f[2].y, f[3]);
#pragma unroll
            for (int k = 0; k < A; k++) {
                V ph = g01[k & 3];
                if ((k >> 2) != 0) ph = cmul<V, T>(ph.x, ph.y, g23[k >> 2]);
                v[k] = cmul<V, T>(ph.x, ph.y, v[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < A; k++) tile[swz<T>(ti(k))] = v[k];
}

// ---- phase fan (OP_FAN, pass_types.hpp), slow form.  The fast form is part of op_sweep.
// the same fan, one amplitude at a time, everything read from the table block in global memory (generic kernel and the
// kernel variants without the 16-amplitudes-per-thread layout): correctness path, not a fast one
template <typename T>
__device__ void op_fan_slow(typename Vec2<T>::type *tile, const DevOp &fo, const uint8_t (*fan_q)[8], const int nch,
                            const typename Vec2<T>::type *__restrict__ blk, const uint64_t base, const int t, const int tid, const int nt)
{
    using V = typename Vec2<T>::type;
    const int m = fo.nfix, nc = fo.k;
    uint32_t u[kFanMaxChunks];
    for (int ch = 0; ch < nch; ch++) {
        u[ch] = 0;
        for (int bit = 0; bit < 8; bit++)
            if (fan_q[ch][bit] != 0xFF) u[ch] |= (uint32_t)((base >> fan_q[ch][bit]) & 1ull) << bit;
    }
    const V *Tt = blk + 1, *X = blk + 1 + ((size_t)nc << m);
    for (uint32_t i = tid; i < (1u << t); i += nt) {
        uint32_t v = 0;
        for (int j = 0; j < m; j++) v |= ((i >> fo.fixpos[j]) & 1u) << j;
        V ph = blk[0];
        for (int c = 0; c < nc; c++) {
            if (!((i >> fo.tbit[c]) & 1u)) continue;
            const V tv = Tt[(c << m) + v];
            ph = cmul<V, T>(ph.x, ph.y, tv);
            for (int ch = 0; ch < nch; ch++) { const V x = X[((size_t)ch * nc + c) * 256 + u[ch]]; ph = cmul<V, T>(ph.x, ph.y, x); }
        }
        const uint32_t sidx = swz<T>(i);
        const V a = tile[sidx];
        tile[sidx] = cmul<V, T>(ph.x, ph.y, a);
    }
}

// which ops of the pass act on the tile at `base` of state `b`: unconditional ops always do; ops with
// controls outside the tile or a per-state mask (rare; listed in cond_mask) are tested individually
template <typename T>
__device__ __forceinline__ uint64_t active_ops(const PassParams<T> &p, const uint8_t *__restrict__ flags,
                                               const uint64_t b, const uint64_t base)
{
    const uint64_t all = p.n_ops >= 64 ? ~0ull : ((1ull << p.n_ops) - 1ull);
    uint64_t active = all & ~p.cond_mask;
    for (uint64_t m = p.cond_mask; m != 0; m &= m - 1) {
        const int o = __ffsll((long long)m) - 1;
        const DevOp &op = p.ops[o];
        bool on = (base & op.ctrl_out) == op.ctrl_out;
        if (on && op.has_mask) on = flags[(size_t)op.mask_row * p.batch + b] != 0;
        if (on) active |= 1ull << o;
    }
    return active;
}

// ---- apply the pass's op list to one staged tile (block-wide; ends with a barrier) -------------
// COND = false: no op of the pass depends on the tile (no outside controls, no per-state masks), so the
// whole op loop is driven by kernel parameters only and stays on the uniform datapath.
// lut/dhi: per-diagonal-op bit-extract tables (build_diag_luts) and per-tile high index parts; lut == nullptr
// selects the slow per-bit gather (generic kernel).
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// hook() is invoked exactly once, right after the barrier that ends the first executed op whose index is
// >= hook_at (or at the very end): the TMA kernel uses it to recycle the previous tile's buffer while the
// remaining ops still run.
// VAR: 0 = lean (1-2 qubit dense gates, the common register blocks, permutations), 1 = lean + diagonal runs through the
// LUT / register-block path, 2 = everything (dense k >= 3, all register blocks, the generic per-bit diagonal gather)
template <typename T, int NT, bool COND, int VAR, typename Hook = NoHook, int FORM = 0>
__device__ __forceinline__ void apply_ops(const PassParams<T> &p, const T *mats, typename Vec2<T>::type *tile,
                                          const typename Vec2<T>::type *__restrict__ diag_tabs,
                                          const uint64_t active, const uint64_t base, const int t, const int tid,
                                          const uint16_t *lut = nullptr, const typename Vec2<T>::type *const *tptr = nullptr,
                                          const int hook_at = 1 << 30, Hook hook = Hook(),
                                          const typename Vec2<T>::type *fanT = nullptr, const typename Vec2<T>::type *fanc = nullptr,
                                          const T *mats_sweep = nullptr)
{
    using V = typename Vec2<T>::type;
    const uint32_t tile_amps = 1u << t;
    bool hooked = false;
    for (int o = 0; o < p.n_ops; o++) {
        if (!hooked && o > hook_at) { hook(); hooked = true; }
            if (COND && !((active >> o) & 1ull)) continue;
            const DevOp &op = p.ops[o];
            if constexpr (VAR >= 1 && !COND) {
                // a register sweep starts here (16-amplitudes-per-thread layout, LUT route): several ops, one barrier
                if (lut != nullptr && p.sweep_at[o] != 0xFF) {
                    const int sw = p.sweep_at[o];
                    FanCtx<T> fc;
                    fc.fanT = fanT;
                    fc.fanc = fanc;
                    op_sweep<T, NT>(tile, p, sw, mats_sweep, diag_tabs, tptr, lut, fc, tid);   // (not inlined: matrices from shared memory)
                    o += p.sweeps[sw].len - 1;
                    __syncthreads();
                    continue;
                }
            }
            if (op.kind == OP_DENSE) {
                const T *mat = mats + op.mat_off;
                if (op.k == 1) op_dense<T, 1, NT, FORM == 1>(tile, op, mat, t, tid);
                else if (op.k == 2) op_dense<T, 2, NT, FORM == 1>(tile, op, mat, t, tid);
                else if constexpr (VAR == 2) {
                    if (op.k == 3) op_dense<T, 3, NT, false>(tile, op, mat, t, tid);
                    else if (op.k == 4) op_dense<T, 4, NT, false>(tile, op, mat, t, tid);
                    else op_dense<T, 5, NT, false>(tile, op, mat, t, tid);
                }
                __syncthreads();
            } else if (op.kind == OP_PAIR) {
                op_pair_dispatch<T, NT, VAR, FORM>(tile, op, mats + op.mat_off, t, tid);
                __syncthreads();
            } else if (op.kind == OP_DIAG) {
                // (outside a register sweep: kernels without the 16-amplitudes-per-thread layout, or per-tile conditions)
                if constexpr (VAR == 2) {
                    int oe = o;
                    while (oe + 1 < p.n_ops && p.ops[oe + 1].kind == OP_DIAG) oe++;
                    for (uint32_t i = tid; i < tile_amps; i += NT) {
                        const uint32_t s = swz<T>(i);
                        V v = tile[s];
                        for (int q = o; q <= oe; q++) {
                            if (COND && !((active >> q) & 1ull)) continue;
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
                }
            } else if (op.kind == OP_FAN) {
                if constexpr (VAR == 2) {   // (outside a register sweep)
                    op_fan_slow<T>(tile, op, p.fan_q, p.fan_chunks, diag_tabs + op.mat_off, base, t, tid, NT);
                    __syncthreads();
                }
            } else if (op.kind == OP_LINPERM) {
                op_linperm<T, NT>(tile, op, t, tid);
                __syncthreads();
            } else {
                op_perm<T, NT>(tile, op, t, tid);
                __syncthreads();
            }
        }
    if (!hooked) hook();
}

// ---- the pass kernel ---------------------------------------------------------------------------
// grid: persistent CTAs striding over all tiles of all states of the batch.
// dynamic smem: 2^t amplitudes + 2^(t-L) run offsets.
template <typename T>
__global__ void __launch_bounds__(kTileThreads, 2)
tile_pass_kernel(const __grid_constant__ PassParams<T> p, typename Vec2<T>::type *__restrict__ state,
                 const typename Vec2<T>::type *__restrict__ diag_tabs, const uint8_t *__restrict__ flags,
                 const T *__restrict__ reg_mats, const uint32_t reg_mat_stride, const uint32_t reg_diag_stride)
{
    using V = typename Vec2<T>::type;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn;
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

        if ((base & p.chunk_mask) != p.chunk_val) continue;  // not this launch's chunk of the state
        // which ops act on this tile (controls outside the tile, per-state masks)
        const uint64_t active = active_ops(p, flags, b, base);
        if (active == 0) continue;  // untouched tile: no traffic at all

        V *gbase = state + (b << n) + base;
        // ---- stage in: 128-bit coalesced loads, swizzled shared-memory stores
#pragma unroll 4
        for (uint32_t i = tid; i < tile_amps; i += kTileThreads)
            tile[swz<T>(i)] = gbase[run_off[i >> L] + (i & run_mask)];
        __syncthreads();

        // reg_mats != nullptr: every state of the batch is its own register with its own matrices and diagonal tables
        // (same pass structure): state b reads row b of the per-register buffers
        apply_ops<T, kTileThreads, true, 2>(p, reg_mats ? reg_mats + (size_t)b * reg_mat_stride : p.mats, tile,
                                               diag_tabs + (size_t)b * reg_diag_stride, active, base, t, tid);

        // ---- stage out
#pragma unroll 4
        for (uint32_t i = tid; i < tile_amps; i += kTileThreads)
            gbase[run_off[i >> L] + (i & run_mask)] = tile[swz<T>(i)];
        __syncthreads();
    }
}

// =================================================================================================
// TMA-staged, software-pipelined pass kernel (the production path for every pass whose tiles are
// made of >= 256-byte runs).
//
//   * one persistent CTA per SM, NT threads, kStages tile buffers in shared memory;
//   * tiles move HBM -> smem and smem -> HBM with cp.async.bulk.tensor (TMA) through a 3-D tensor
//     map over the state ([128-byte row][2^20 rows][rest]) with the hardware 128-byte swizzle, one
//     TMA op per contiguous run (box = 2^Lb amplitudes); no register staging, no LSU issue slots;
//   * loads complete on an mbarrier (complete_tx), stores are tracked with bulk async-groups;
//   * while tile k is being computed, tile k+1 is landing and tile k-1 is draining.
// =================================================================================================
constexpr int kAmpsPerThreadShift = 4;  // a thread owns 2^4 amplitudes of the tile (one 4-bit register block)

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 5-D tensor map of a pass: {128-byte row, row index in the whole buffer, tile qubit a, b, c}; the box is
// {row, 2^(L-rowshift) rows, 2, 2, 2}: one TMA op moves the 8 runs spanned by three high tile qubits.
__device__ __forceinline__ void tma_load_row(void *dst, const CUtensorMap *map, uint64_t *bar, int row)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %3, %3, %3}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(row)
        : "memory");
}
__device__ __forceinline__ void tma_store_row(const CUtensorMap *map, const void *src, int row)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %2, %2, %2}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(0), "r"(row)
                 : "memory");
}

// NT threads per CTA = 2^(t-4); 256/NT CTAs are resident per SM, each running its own pipeline on its own
// tiles, so the load / math / store / barrier phases of different CTAs interleave on an SM.
// kStages tile buffers per CTA: 3 = the CTA overlaps its own load / math / store; 2 = one more CTA fits per
// SM and the overlap comes from the other resident CTAs.
// FORM = PassParams::mat_form (op_pair_dispatch).  FORM 1: the dense matrices of the register blocks and dense gates are
// read straight from the kernel parameters (uniform datapath) instead of the shared-memory copy.
template <typename T, int NT, bool COND, int VAR, int kStages, int FORM = 0>
__global__ void __launch_bounds__(NT, (kStages == 3 ? 256 : kStages == 2 ? 384 : 512) / NT)
tile_pass_tma_kernel(const __grid_constant__ PassParams<T> p, const __grid_constant__ CUtensorMap tmap,
                     const typename Vec2<T>::type *__restrict__ diag_tabs, const uint8_t *__restrict__ flags,
                     const int box_bits, const uint32_t buf_stride, const int refill_after,
                     const T *__restrict__ reg_mats, const uint32_t reg_mat_stride, const uint32_t reg_diag_stride)
{
    using V = typename Vec2<T>::type;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // tile buffers must sit on 1024-byte boundaries for the 128-byte swizzle pattern
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int t = p.t, L = p.L, n = p.n;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    // one TMA box covers tile bits [0, box_bits): the contiguous run (or a 256-row piece of it when the run
    // is longer) plus up to three high tile qubits folded into the tensor map
    const uint32_t nbox = 1u << (t - box_bits);    // TMA ops per tile per direction
    const uint32_t tile_bytes = (uint32_t)sizeof(V) << t;
    constexpr int kAmpsPerRowShift = sizeof(V) == 16 ? 3 : 4;  // amplitudes per 128-byte row
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kStages * buf_stride);
    // row offset (128-byte rows) of TMA box r relative to the tile's first row
    uint32_t *box_row = reinterpret_cast<uint32_t *>(full + kStages + 1);
    // the pass's dense matrices, staged once: broadcast shared-memory loads instead of constant-bank loads
    // (16-byte aligned: complex elements are read with one 128-bit load; plain pointer arithmetic keeps the
    // shared-memory address space visible to the compiler)
    unsigned char *mats_raw = reinterpret_cast<unsigned char *>(box_row + ((nbox + 3) & ~3u));
    mats_raw += (16u - (smem_u32(mats_raw) & 15u)) & 15u;
    T *mats_s = reinterpret_cast<T *>(mats_raw);
    // reg_mats != nullptr: many independent registers batched in one launch (same pass structure, each state of the
    // batch with its own matrices and diagonal tables): the matrices are (re)staged whenever the CTA moves on to a tile
    // of another state -- tiles are numbered state-major, so that happens once per 2^(n-t)/gridDim tiles
    if (FORM != 1 && reg_mats == nullptr)
        for (int e = tid; e < p.n_mat; e += NT) mats_s[e] = p.mats[e];
    uint64_t staged_b = ~0ull;
    // per diagonal op: this tile's sub-table pointer, the per-tile index part that comes from qubits outside the tile,
    // bit-extract tables of the in-tile index part (tile bits 0..5 and 6..12), and the store of staged sub-tables
    const V **tptr = reinterpret_cast<const V **>(mats_s + ((p.n_mat + 1) & ~1));
    V *stab = reinterpret_cast<V *>(tptr + ((p.n_diag + 1) & ~1));
    uint32_t *dhi = reinterpret_cast<uint32_t *>(stab + p.n_stab);
    uint16_t *lut = reinterpret_cast<uint16_t *>(dhi + p.n_diag);
    // phase fan: T[c][2^m] (staged with the matrices), this tile's center factors and their per-chunk parts
    const int fan_nc = p.fan_op >= 0 ? p.ops[p.fan_op].k : 0, fan_m = p.fan_op >= 0 ? p.ops[p.fan_op].nfix : 0;
    unsigned char *fan_raw = reinterpret_cast<unsigned char *>(lut + (size_t)p.n_diag * kDiagLut);
    fan_raw += (16u - (smem_u32(fan_raw) & 15u)) & 15u;
    V *fanT = reinterpret_cast<V *>(fan_raw);
    V *fanc = fanT + ((size_t)fan_nc << fan_m);
    V *fanp = fanc + fan_nc;
    if (VAR >= 1 && p.fan_op >= 0 && reg_mats == nullptr) {
        const V *src = diag_tabs + p.ops[p.fan_op].mat_off + 1;
        for (int e = tid; e < (fan_nc << fan_m); e += NT) fanT[e] = src[e];
    }
    static_assert(NT >= 32, "one warp at least");
    for (int e = tid; VAR >= 1 && e < p.n_diag * kDiagLut; e += NT) {
        const int sl = e / kDiagLut, r = e - sl * kDiagLut;
        const uint32_t mask = p.ops[p.diag_op[sl]].ctrl_in;
        const uint32_t val = r < 64 ? (uint32_t)r : (uint32_t)(r - 64) << 6;
        uint32_t out = 0, kbit = 0;
        for (uint32_t m = r < 64 ? (mask & 63u) : (mask & ~63u); m != 0; m &= m - 1, kbit++)
            out |= ((val >> (__ffs((int)m) - 1)) & 1u) << kbit;
        if (r >= 64) out <<= __popc(mask & 63u);
        lut[e] = (uint16_t)out;
    }

    // amplitude offset of box r inside a tile: tile-index bit j >= box_bits of the box number is deposited at
    // physical qubit tile_q[j] (which is j itself while j < L: the box is then a piece of the run)
    for (uint32_t r = tid; r < nbox; r += NT) {
        uint64_t o = 0;
        for (int j = box_bits; j < t; j++)
            if ((r >> (j - box_bits)) & 1u) o |= 1ull << p.tile_q[j];
        box_row[r] = (uint32_t)(o >> kAmpsPerRowShift);
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const uint64_t total_tiles = p.tiles_per_state * (uint64_t)p.batch;
    const int tps_shift = n - t;  // tiles_per_state = 2^(n-t)
    // tile number -> amplitude offset of the tile inside the whole buffer (state b, base within the state): the bits of the
    // tile number move up past the high tile qubits segment by segment (PassParams::tile_seg, 32-bit arithmetic: the
    // deposited number is base >> L < 2^(n-L))
    const int nseg = t - L + 1;
    auto tile_dep = [&](uint64_t tileno, uint64_t &b) -> uint32_t {
        b = tileno >> tps_shift;
        const uint32_t x = (uint32_t)(tileno & (p.tiles_per_state - 1ull));
        uint32_t dep = 0;   // (masks past the last segment are zero: static indices, the masks are constant-bank operands)
#pragma unroll
        for (int k = 0; k < 8; k++) dep |= (x & p.tile_seg[k]) << k;
        if (nseg > 8) {
#pragma unroll
            for (int k = 8; k <= kMaxTileBits; k++) dep |= (x & p.tile_seg[k]) << k;
        }
        return dep;
    };
    auto tile_base = [&](uint64_t tileno, uint64_t &b) -> uint64_t { return (uint64_t)tile_dep(tileno, b) << L; };
    // next tile of this CTA (stride gridDim.x) that some op acts on; ~0 = none.  Evaluated identically by
    // every thread, so the tile sequence lives in registers and needs no shared bookkeeping.
    auto next_active = [&](uint64_t from) -> uint64_t {
        if constexpr (!COND) {
            if (p.chunk_mask == 0) return from < total_tiles ? from : ~0ull;
            for (uint64_t x = from; x < total_tiles; x += gridDim.x) {   // chunked pass: skip the tiles of other chunks
                uint64_t b;
                if ((tile_base(x, b) & p.chunk_mask) == p.chunk_val) return x;
            }
            return ~0ull;
        } else {
            for (uint64_t x = from; x < total_tiles; x += gridDim.x) {
                uint64_t b;
                const uint64_t base = tile_base(x, b);
                if ((base & p.chunk_mask) != p.chunk_val) continue;
                if (active_ops(p, flags, b, base) != 0) return x;
            }
            return ~0ull;
        }
    };
    const uint32_t smem_s = smem_u32(smem_raw);
    const uint32_t box_bytes = (uint32_t)sizeof(V) << box_bits;
    // first 128-byte row of a tile in the whole buffer
    auto tile_row = [&](uint64_t tileno) -> uint32_t {
        uint64_t b;
        const uint32_t dep = tile_dep(tileno, b);
        return (uint32_t)(b << (n - kAmpsPerRowShift)) + (dep << (L - kAmpsPerRowShift));   // (the TMA path has L >= the row shift)
    };
    auto issue_load = [&](int s, uint32_t row0) {
        if (tid == 0) mbar_arrive_expect_tx(&full[s], tile_bytes);
        // TMA instructions take warp-uniform operands: lane 0 of warp w issues boxes w, w+nwarps, ...
        if (lane == 0) {
            const uint32_t dst = smem_s + (uint32_t)s * buf_stride, bar = smem_u32(&full[s]);
            for (uint32_t r = warp; r < nbox; r += NT / 32)
                asm volatile(
                    "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %3, %3, %3}], [%2];"
                    ::"r"(dst + r * box_bytes), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar), "r"(0), "r"((int)(row0 + box_row[r]))
                    : "memory");
        }
    };

    // tq[i] = the tile this CTA processes i iterations from now (register-resident, statically indexed).
    // Tile j lives in slot j % kStages.  At the START of iteration k the slot of tile k-1 is recycled for tile
    // k+kStages-1: the issuing lane only waits until its own TMA stores of tile k-1 have finished READING shared
    // memory, so every load gets kStages-1 full compute phases of lead time.
    uint64_t tq[kStages];
    uint32_t trow[kStages];  // first row of each queued tile (computed once per tile)
    tq[0] = next_active(blockIdx.x);
#pragma unroll
    for (int i = 1; i < kStages; i++) tq[i] = tq[i - 1] == ~0ull ? ~0ull : next_active(tq[i - 1] + gridDim.x);
#pragma unroll
    for (int i = 0; i < kStages; i++) trow[i] = tq[i] == ~0ull ? 0u : tile_row(tq[i]);
#pragma unroll
    for (int i = 0; i < kStages - 1; i++)
        if (tq[i] != ~0ull) issue_load(i, trow[i]);

    int s = 0;            // slot of tq[0]
    uint32_t parity = 0;  // flips every time s wraps
    while (tq[0] != ~0ull) {
        const int s2 = s == 0 ? kStages - 1 : s - 1;  // slot of the previous tile
        // Recycling that slot needs the previous tile's TMA stores to have finished READING shared memory, which
        // takes as long as HBM needs to absorb them.  Doing it after the first op(s) of this tile instead of here
        // lets the drain overlap the math, while the new load still has the rest of the tile's math to land.
        auto refill = [&]() {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (tq[kStages - 1] != ~0ull) issue_load(s2, trow[kStages - 1]);
        };

        if constexpr (kStages == 1) refill();   // single buffer: drain the previous tile's stores, then load this tile
        uint64_t b;
        const uint64_t base = tile_base(tq[0], b);
        V *tile = reinterpret_cast<V *>(smem_raw + (size_t)s * buf_stride);
        // per-tile base of every diagonal table (the index bits that come from qubits outside the tile)
        const V *tabs_b = diag_tabs + (size_t)b * reg_diag_stride;
        if constexpr (VAR >= 1) {   // (the lean variant runs passes without diagonal tables and fans)
        for (int sl = tid; sl < p.n_diag; sl += NT) {
            const DevOp &d = p.ops[p.diag_op[sl]];
            uint32_t hi = 0;
            for (int m = d.nfix; m < d.k; m++) hi |= (uint32_t)((base >> (d.dsrc[m] - 64)) & 1ull) << m;
            dhi[sl] = hi + d.mat_off;
            tptr[sl] = d.stage_off != kNotStaged ? stab + d.stage_off : tabs_b + hi + d.mat_off;
        }
        if (p.n_stab) {
            // small sub-tables (few in-tile qubits) are copied into shared memory: 2^nfix entries per op per tile
            __syncthreads();
            for (int x = tid; x < p.n_stab; x += NT) {
                const int sl = p.stab_slot[x >> 3];
                const DevOp &d = p.ops[p.diag_op[sl]];
                const uint32_t e = (uint32_t)x - d.stage_off;
                if (e < (1u << d.nfix)) stab[x] = tabs_b[dhi[sl] + e];
            }
        }
        }
        bool restaged = false;
        if (FORM == 0 && reg_mats != nullptr && b != staged_b) {
            // (mats_s is only read inside apply_ops, and every thread left the previous tile's apply_ops through a barrier)
            const T *src = reg_mats + (size_t)b * reg_mat_stride;
            for (int e = tid; e < p.n_mat; e += NT) mats_s[e] = src[e];
            if (p.fan_op >= 0) {
                const V *fsrc = tabs_b + p.ops[p.fan_op].mat_off + 1;
                for (int e = tid; e < (fan_nc << fan_m); e += NT) fanT[e] = fsrc[e];
            }
            staged_b = b;
            restaged = true;
        }
        if (VAR >= 1 && p.fan_op >= 0) {
            // out-of-tile leaves of the fan: one 256-entry table per (chunk of 8 qubits, center), indexed by the tile base
            const DevOp &fo = p.ops[p.fan_op];
            const V *X = tabs_b + fo.mat_off + 1 + ((size_t)fan_nc << fan_m);
            for (int e = tid; e < fan_nc * p.fan_chunks; e += NT) {
                const int ch = e / fan_nc;
                uint32_t u = 0;
                for (int bit = 0; bit < 8; bit++)
                    if (p.fan_q[ch][bit] != 0xFF) u |= (uint32_t)((base >> p.fan_q[ch][bit]) & 1ull) << bit;
                fanp[e] = X[(size_t)e * 256 + u];
            }
        }
        mbar_wait(&full[s], parity);
        if ((VAR >= 1 && (p.n_diag || p.fan_op >= 0)) || restaged) __syncthreads();
        if (VAR >= 1 && p.fan_op >= 0) {
            if (tid < fan_nc) {
                V c;
                c.x = (T)1; c.y = (T)0;
                for (int ch = 0; ch < p.fan_chunks; ch++) { const V x = fanp[ch * fan_nc + tid]; c = cmul<V, T>(c.x, c.y, x); }
                fanc[tid] = c;
            }
            __syncthreads();
        }

        if constexpr (kStages == 1) {
            // (one buffer: the tile's own load was issued at the top of the iteration, nothing to recycle during the ops)
            apply_ops<T, NT, COND, VAR, NoHook, FORM>(p, FORM == 1 ? p.mats + (uint32_t)b * (uint32_t)p.reg_stride : mats_s, tile, tabs_b, COND ? active_ops(p, flags, b, base) : ~0ull, base, t, tid,
                                        (COND || (16 * NT) != (1 << t)) ? nullptr : lut, tptr, 1 << 30, NoHook(),
                                        (COND || (16 * NT) != (1 << t)) ? nullptr : fanT, fanc, mats_s);
        } else {
        apply_ops<T, NT, COND, VAR, decltype(refill), FORM>(p, FORM == 1 ? p.mats + (uint32_t)b * (uint32_t)p.reg_stride : mats_s, tile, tabs_b, COND ? active_ops(p, flags, b, base) : ~0ull, base, t, tid,
                                    (COND || (16 * NT) != (1 << t)) ? nullptr : lut, tptr, refill_after, refill,
                                    (COND || (16 * NT) != (1 << t)) ? nullptr : fanT, fanc, mats_s);
        }

        // every thread's generic-proxy writes -> visible to the async proxy; then drain the tile with TMA
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (lane == 0) {
            const uint32_t src = smem_s + (uint32_t)s * buf_stride;
            for (uint32_t r = warp; r < nbox; r += NT / 32)
                asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %2, %2, %2}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(src + r * box_bytes), "r"(0), "r"((int)(trow[0] + box_row[r]))
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#pragma unroll
        for (int i = 0; i < kStages - 1; i++) { tq[i] = tq[i + 1]; trow[i] = trow[i + 1]; }
        tq[kStages - 1] = tq[kStages - 1] == ~0ull ? ~0ull : next_active(tq[kStages - 1] + gridDim.x);
        trow[kStages - 1] = tq[kStages - 1] == ~0ull ? 0u : tile_row(tq[kStages - 1]);
        if (++s == kStages) { s = 0; parity ^= 1u; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace cb200
