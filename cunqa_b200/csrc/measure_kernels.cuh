// measure_kernels.cuh -- K4 (norm / probability reductions, warp-shuffle), K5 (shot sampling from a
// hierarchical cumulative-probability scan) and K6 (measure collapse + renormalise) of SURVEY 8a.
//
// Replaces what the reference gets from AerState::apply_measure / apply_reset
// (aer_simulator_adapter.cpp:166-167,187,218) and the static path's final sample_measure (:829).
// All reductions use a fixed summation order so that seeded runs are reproducible.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cb200 {

constexpr int kRedThreads = 256;
constexpr int kChunkBits = 10;  // sampling hierarchy: 2^10 entries per level

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in thread 0
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double ws[kRedThreads / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kRedThreads / 32 ? ws[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

template <typename V> __device__ __forceinline__ double abs2(const V a)
{
    return (double)a.x * (double)a.x + (double)a.y * (double)a.y;
}

__device__ __forceinline__ double splitmix_uniform(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// |0...0> for every state of the batch
template <typename V>
__global__ void init_zero_kernel(V *state, uint64_t amps_per_state, uint64_t total, int has_origin, uint64_t origin)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        V v; v.x = (has_origin && (i & (amps_per_state - 1)) == origin) ? 1 : 0; v.y = 0;
        state[i] = v;
    }
}

// copy state 0 of a batch onto states 1..batch-1 (shared deterministic prefix of batched dynamic shots)
template <typename V>
__global__ void broadcast_state0_kernel(V *state, uint64_t amps_per_state, uint64_t total)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = amps_per_state + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        state[i] = state[i & (amps_per_state - 1)];
}

// partial[b*gridDim.x + blockIdx.x] = sum over this block's slice of state b of |a|^2 restricted to
// amplitudes with (index & sel_mask) == sel_mask (sel_mask = 0: full norm; 1<<q: P(q=1)).
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
prob_partial_kernel(const V *__restrict__ state, uint64_t amps_per_state, uint64_t sel_mask, double *__restrict__ partial)
{
    const uint64_t b = blockIdx.y;
    const V *s = state + b * amps_per_state;
    const uint64_t per_block = (amps_per_state + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    uint64_t hi = lo + per_block;
    if (hi > amps_per_state) hi = amps_per_state;
    double acc = 0.0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kRedThreads)
        if ((i & sel_mask) == sel_mask) acc += abs2(s[i]);
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = acc;
}

// both sums in ONE read of the state: partial_all = sum |a|^2, partial_sel = sum over (index & sel_mask) == sel_mask
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
prob2_partial_kernel(const V *__restrict__ state, uint64_t amps_per_state, uint64_t sel_mask,
                     double *__restrict__ partial_all, double *__restrict__ partial_sel)
{
    const uint64_t b = blockIdx.y;
    const V *s = state + b * amps_per_state;
    const uint64_t per_block = (amps_per_state + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    uint64_t hi = lo + per_block;
    if (hi > amps_per_state) hi = amps_per_state;
    double acc = 0.0, sel = 0.0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kRedThreads) {
        const double p = abs2(s[i]);
        acc += p;
        if ((i & sel_mask) == sel_mask) sel += p;
    }
    acc = block_sum(acc);
    sel = block_sum(sel);
    if (threadIdx.x == 0) { partial_all[b * gridDim.x + blockIdx.x] = acc; partial_sel[b * gridDim.x + blockIdx.x] = sel; }
}

// reduced density matrix of one qubit (physical bit `pq`), per state: partial[0] = rho00, [1] = rho11,
// [2] + i [3] = rho01 = sum_rest a(rest,0) conj(a(rest,1)).  Each of the four arrays has batch * gridDim.x entries.
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
rho_partial_kernel(const V *__restrict__ state, uint64_t amps_per_state, int pq, double *__restrict__ partial, size_t arr)
{
    const uint64_t b = blockIdx.y;
    const V *s = state + b * amps_per_state;
    const uint64_t pairs = amps_per_state >> 1;
    const uint64_t per_block = (pairs + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    uint64_t hi = lo + per_block;
    if (hi > pairs) hi = pairs;
    const uint64_t bit = 1ull << pq;
    double r00 = 0.0, r11 = 0.0, re = 0.0, im = 0.0;
    for (uint64_t g = lo + threadIdx.x; g < hi; g += kRedThreads) {
        const uint64_t i0 = ((g >> pq) << (pq + 1)) | (g & (bit - 1ull));
        const V a0 = s[i0], a1 = s[i0 | bit];
        const double x0 = (double)a0.x, y0 = (double)a0.y, x1 = (double)a1.x, y1 = (double)a1.y;
        r00 += x0 * x0 + y0 * y0;
        r11 += x1 * x1 + y1 * y1;
        re += x0 * x1 + y0 * y1;
        im += y0 * x1 - x0 * y1;
    }
    r00 = block_sum(r00);
    r11 = block_sum(r11);
    re = block_sum(re);
    im = block_sum(im);
    if (threadIdx.x == 0) {
        const size_t o = b * gridDim.x + blockIdx.x;
        partial[o] = r00; partial[arr + o] = r11; partial[2 * arr + o] = re; partial[3 * arr + o] = im;
    }
}

// out[b] = sum_j partial[b*nblk + j]   (one block per state, fixed order)
__global__ void __launch_bounds__(kRedThreads) sum_partials_kernel(const double *__restrict__ partial, int nblk, double *__restrict__ out)
{
    const int b = blockIdx.x;
    double acc = 0.0;
    for (int j = threadIdx.x; j < nblk; j += kRedThreads) acc += partial[(size_t)b * nblk + j];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[b] = acc;
}

// measurement decision for every state: outcome, collapse scale.  norm[b], p1[b] given.
__global__ void measure_decide_kernel(const double *__restrict__ norm, const double *__restrict__ p1, int batch,
                                      uint64_t seed, const uint64_t *__restrict__ counters,
                                      const int8_t *__restrict__ forced, int *__restrict__ outcome, double *__restrict__ scale)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double total = norm[b], q1 = p1[b], q0 = total - q1;
    int out;
    if (forced != nullptr && forced[b] == -2) {  // this state does not take part (conditioned measure)
        outcome[b] = -1;
        scale[b] = 1.0;
        return;
    }
    if (forced != nullptr && forced[b] >= 0) out = forced[b] ? 1 : 0;
    else out = (splitmix_uniform(seed, counters[b]) * total >= q0) ? 1 : 0;
    const double pk = out ? q1 : q0;
    outcome[b] = out;
    scale[b] = pk > 0.0 ? rsqrt(pk) : 0.0;
}

// collapse qubit q of every state onto outcome[b], scaling the kept half by scale[b].
// flip_to_zero: additionally move the kept |1> half onto |0> (reset = measure + conditional X).
template <typename V>
__global__ void collapse_kernel(V *state, int n, int q, int batch, const int *__restrict__ outcome,
                                const double *__restrict__ scale, int flip_to_zero)
{
    const uint64_t half = 1ull << (n - 1);
    const uint64_t total = half * (uint64_t)batch;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const uint64_t b = w >> (n - 1);
        const uint64_t g = w & (half - 1);
        const uint64_t lo = g & ((1ull << q) - 1ull);
        const uint64_t i0 = ((g >> q) << (q + 1)) | lo;
        const uint64_t i1 = i0 | (1ull << q);
        V *s = state + (b << n);
        const int out = outcome[b];
        if (out < 0) continue;
        const double sc = scale[b];
        V keep = out ? s[i1] : s[i0];
        keep.x = (decltype(keep.x))(keep.x * sc);
        keep.y = (decltype(keep.y))(keep.y * sc);
        V zero; zero.x = 0; zero.y = 0;
        if (out && !flip_to_zero) { s[i0] = zero; s[i1] = keep; }
        else { s[i0] = keep; s[i1] = zero; }
    }
}

// level sums for sampling: out[j] = sum of 2^bits consecutive inputs (|a|^2 of amplitudes at level 0)
template <typename V>
__global__ void __launch_bounds__(kRedThreads) chunk_sums_amp_kernel(const V *__restrict__ state, int bits, uint64_t nchunks, double *__restrict__ out)
{
    const uint64_t len = 1ull << bits;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        double acc = 0.0;
        for (uint64_t i = threadIdx.x; i < len; i += kRedThreads) acc += abs2(state[c * len + i]);
        acc = block_sum(acc);
        if (threadIdx.x == 0) out[c] = acc;
    }
}
__global__ void __launch_bounds__(kRedThreads) chunk_sums_f64_kernel(const double *__restrict__ in, int bits, uint64_t nchunks, double *__restrict__ out)
{
    const uint64_t len = 1ull << bits;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        double acc = 0.0;
        for (uint64_t i = threadIdx.x; i < len; i += kRedThreads) acc += in[c * len + i];
        acc = block_sum(acc);
        if (threadIdx.x == 0) out[c] = acc;
    }
}
// inclusive prefix of the (small) top level, one thread: m <= 2^16
__global__ void top_prefix_kernel(double *a, uint64_t m)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (uint64_t i = 0; i < m; i++) { acc += a[i]; a[i] = acc; }
    }
}

// warp-cooperative search: first j in [0,len) with sum(vals[0..j]) > u ; u is reduced by the
// mass before j.  `load(j)` returns the j-th non-negative value.
template <typename F>
__device__ __forceinline__ uint64_t warp_search(F load, uint64_t len, double &u)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (uint64_t blk = 0; blk < len; blk += 32) {
        const uint64_t j = blk + lane;
        const double v = j < len ? load(j) : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, (j < len) && (acc + inc > u));
        if (hit) {
            const int first = __ffs(hit) - 1;
            const double before = __shfl_sync(0xffffffffu, inc - v, first);
            u -= acc + before;
            return blk + first;
        }
        acc += __shfl_sync(0xffffffffu, inc, 31);
    }
    u -= acc;  // rounding: fell off the end
    return len - 1;
}

// one warp per shot.  lvl2 = inclusive prefix over top-level groups (n2 entries), lvl1 = chunk sums.
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
sample_kernel(const V *__restrict__ state, int n, int bits0, int bits1, const double *__restrict__ lvl1,
              const double *__restrict__ lvl2_prefix, uint64_t n2, uint64_t shots, uint64_t seed, uint64_t *__restrict__ out,
              double total_all, double off_lo, double off_hi, uint64_t index_prefix, uint64_t ctr0)
{
    // sharded state: total_all = norm of the whole state, this shard owns the CDF interval [off_lo, off_hi) and
    // writes 0 for shots that fall elsewhere (an all-reduce sums the ranks' answers); single GPU: total_all < 0
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t s = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); s < shots; s += warps) {
        const double total = total_all < 0.0 ? lvl2_prefix[n2 - 1] : total_all;
        double u = splitmix_uniform(seed, ctr0 + s) * total;
        if (total_all >= 0.0) {
            if (!(u >= off_lo && u < off_hi)) { if (lane == 0) out[s] = 0; continue; }
            u -= off_lo;
        }
        // top level: binary search on the inclusive prefix (first g with prefix[g] > u)
        uint64_t lo = 0, hi = n2;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (lvl2_prefix[mid] > u) hi = mid; else lo = mid + 1;
        }
        if (lo >= n2) lo = n2 - 1;
        if (lo > 0) u -= lvl2_prefix[lo - 1];
        const uint64_t g = lo;
        const uint64_t len1 = 1ull << bits1, len0 = 1ull << bits0;
        const double *l1 = lvl1 + g * len1;
        const uint64_t c = warp_search([&](uint64_t j) { return l1[j]; }, len1, u);
        const V *a = state + ((g * len1 + c) << bits0);
        const uint64_t i = warp_search([&](uint64_t j) { return abs2(a[j]); }, len0, u);
        if (lane == 0) out[s] = index_prefix | (((g * len1 + c) << bits0) + i);
    }
}

// one sample per state of the batch (terminal measurement of a dynamic shot): one warp per state.
// lvl1: chunk sums of 2^bits0 amplitudes over the whole batch; lvl2: sums of 2^bits1 lvl1 entries.
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
sample_each_kernel(const V *__restrict__ state, int n, int bits0, int bits1, const double *__restrict__ lvl1,
                   const double *__restrict__ lvl2, int batch, uint64_t seed, const uint64_t *__restrict__ counters,
                   uint64_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t n2 = 1ull << (n - bits0 - bits1), len1 = 1ull << bits1, len0 = 1ull << bits0;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < batch; b += warps) {
        const double *l2 = lvl2 + (uint64_t)b * n2;
        double total = 0.0;
        for (uint64_t j = 0; j < n2; j++) total += l2[j];  // n2 <= 2^14, same order on every lane
        double u = splitmix_uniform(seed, counters[b]) * total;
        const uint64_t g = warp_search([&](uint64_t j) { return l2[j]; }, n2, u);
        const double *l1 = lvl1 + ((uint64_t)b * n2 + g) * len1;
        const uint64_t c = warp_search([&](uint64_t j) { return l1[j]; }, len1, u);
        const V *a = state + ((uint64_t)b << n) + ((g * len1 + c) << bits0);
        const uint64_t i = warp_search([&](uint64_t j) { return abs2(a[j]); }, len0, u);
        if (lane == 0) out[b] = ((g * len1 + c) << bits0) + i;
    }
}

// `shots` samples of EVERY state of the batch (independent registers, each with its own seed): one warp per (state, shot),
// u = U(seeds[b], s) * norm_b -- the same draw cb200_sample makes for state b alone.
template <typename V>
__global__ void __launch_bounds__(kRedThreads)
sample_regs_kernel(const V *__restrict__ state, int n, int bits0, int bits1, const double *__restrict__ lvl1,
                   const double *__restrict__ lvl2, int batch, uint64_t shots, const uint64_t *__restrict__ seeds,
                   uint64_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n2 = 1ull << (n - bits0 - bits1), len1 = 1ull << bits1, len0 = 1ull << bits0;
    const uint64_t work = (uint64_t)batch * shots;
    for (uint64_t w = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < work; w += warps) {
        const uint64_t b = w / shots, s = w - b * shots;
        const double *l2 = lvl2 + b * n2;
        double total = 0.0;
        for (uint64_t j = 0; j < n2; j++) total += l2[j];  // same order on every lane and for every shot of the state
        double u = splitmix_uniform(seeds[b], s) * total;
        const uint64_t g = warp_search([&](uint64_t j) { return l2[j]; }, n2, u);
        const double *l1 = lvl1 + (b * n2 + g) * len1;
        const uint64_t c = warp_search([&](uint64_t j) { return l1[j]; }, len1, u);
        const V *a = state + (b << n) + ((g * len1 + c) << bits0);
        const uint64_t i = warp_search([&](uint64_t j) { return abs2(a[j]); }, len0, u);
        if (lane == 0) out[w] = ((g * len1 + c) << bits0) + i;
    }
}

template <typename V>
__global__ void gather_amps_kernel(const V *__restrict__ state, const uint64_t *__restrict__ idx, uint64_t n, double2 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const V a = state[idx[i]]; out[i] = make_double2((double)a.x, (double)a.y); }
}

template <typename V>
__global__ void scatter_amps_kernel(V *__restrict__ state, const double2 *__restrict__ in, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        V v; v.x = (decltype(v.x))in[i].x; v.y = (decltype(v.y))in[i].y;
        state[i] = v;
    }
}

// marginal distribution over k qubits: out[2^k] += |a|^2 (global atomics; diagnostics path)
template <typename V>
__global__ void marginal_kernel(const V *__restrict__ state, uint64_t amps, const int *__restrict__ qubits, int k, double *__restrict__ out,
                                const int *__restrict__ outbit = nullptr, uint32_t fixed = 0)
{
    // outbit[m]: position of qubits[m] in the output index (default m); fixed: output bits that are constant for this
    // shard (marginal qubits that are global: their value is the rank's bit)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < amps; i += stride) {
        uint32_t j = fixed;
        for (int m = 0; m < k; m++) j |= (uint32_t)((i >> qubits[m]) & 1ull) << (outbit ? outbit[m] : m);
        atomicAdd(&out[j], abs2(state[i]));
    }
}

}  // namespace cb200
