// pair_variants.cu -- where do the register blocks (OP_PAIR, two 2-qubit gates on the 16 amplitudes a thread holds) lose
// their time?  Same set-up as dmma_pair.cu (a 2^11-amplitude complex128 tile resident in shared memory, 3 CTAs x 128
// threads per SM, a barrier after every pair, no HBM traffic), three forms of the block:
//   smem  : production round-1/2 form: matrix elements read from shared memory (LDS.128 broadcast) into vector registers
//   const : matrix elements read from the kernel parameters (constant bank) through the uniform datapath (LDCU -> UR
//           operands of DFMA): no vector registers and no shared-memory wavefronts for the matrices
//   mul3  : const + 3-multiplication complex products: with s_j = re_j + im_j shared by the four rows of a gate,
//           re_i = sum c_ij s_j + sum [-(c+d)_ij] im_j,  im_i = sum c_ij s_j + sum (d-c)_ij re_j : 52 FP64 ops per group of 4
//           amplitudes instead of 64
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cunqa_b200/csrc tools/pair_variants.cu -o tools/bin/pair_variants
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "tile_kernel.cuh"

using namespace cb200;

struct CP {
    DevOp op[2];
    int off[2];
    alignas(16) double mats[2 * 96];     // per op: gate A (re, im) x 16 in a slot of 48 scalars, gate B likewise (production layout)
    alignas(16) double mats3[2 * 96];    // per op: gate A (c, -(c+d), d-c) x 16, gate B likewise
};

template <int R, int K, int S>
__device__ __forceinline__ void gate_const(double2 (&a)[1 << R], const double *mat)
{
    constexpr int D = 1 << K;
#pragma unroll
    for (int rest = 0; rest < (1 << (R - K)); rest++) {
        const int lo = rest & ((1 << S) - 1), hi = rest >> S;
        const int base = (hi << (S + K)) | lo;
        double2 in[D];
#pragma unroll
        for (int j = 0; j < D; j++) in[j] = a[base | (j << S)];
#pragma unroll
        for (int r = 0; r < D; r++) {
            double2 acc = cmul<double2, double>(mat[2 * (r * D)], mat[2 * (r * D) + 1], in[0]);
#pragma unroll
            for (int c = 1; c < D; c++) cfma<double2, double>(acc, mat[2 * (r * D + c)], mat[2 * (r * D + c) + 1], in[c]);
            a[base | (r << S)] = acc;
        }
    }
}

template <int R, int K, int S>
__device__ __forceinline__ void gate_mul3(double2 (&a)[1 << R], const double *mat)
{
    // two rows of the matrix at a time (24 uniform values live): rows 0,1 of every group go to temporaries, rows 2,3 are
    // written in place once both are known, then rows 0,1 follow
    constexpr int D = 1 << K, G = 1 << (R - K);
    static_assert(K == 2, "2-qubit gates");
    double2 lo01[G][2];
#pragma unroll
    for (int half = 0; half < 2; half++) {
#pragma unroll
        for (int rest = 0; rest < G; rest++) {
            const int lo = rest & ((1 << S) - 1), hi = rest >> S;
            const int base = (hi << (S + K)) | lo;
            double2 in[D];
            double s[D];
#pragma unroll
            for (int j = 0; j < D; j++) { in[j] = a[base | (j << S)]; s[j] = in[j].x + in[j].y; }
            double2 out[2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int r = 2 * half + rr;
                double t = mat[3 * (r * D)] * s[0];
#pragma unroll
                for (int c = 1; c < D; c++) t = fma(mat[3 * (r * D + c)], s[c], t);
                double re = t, im = t;
#pragma unroll
                for (int c = 0; c < D; c++) {
                    re = fma(mat[3 * (r * D + c) + 1], in[c].y, re);
                    im = fma(mat[3 * (r * D + c) + 2], in[c].x, im);
                }
                out[rr] = make_double2(re, im);
            }
            if (half == 0) { lo01[rest][0] = out[0]; lo01[rest][1] = out[1]; }
            else {
                a[base | (2 << S)] = out[0]; a[base | (3 << S)] = out[1];
                a[base] = lo01[rest][0]; a[base | (1 << S)] = lo01[rest][1];
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ void pair_block(double2 *tile, const DevOp &op, const double *mat, const int tid)
{
    constexpr int R = 4, D = 16;
    const uint32_t tile_s = smem_u32(tile);
    uint32_t sb[R], seg[R + 1];
#pragma unroll
    for (int m = 0; m < R; m++) sb[m] = swz<double>(1u << op.tbit[m]) << 4;
#pragma unroll
    for (int m = 0; m <= R; m++) seg[m] = op.seg[m];
    uint32_t idx = 0;
#pragma unroll
    for (int m = 0; m <= R; m++) idx |= ((uint32_t)tid & seg[m]) << m;
    const uint32_t bidx = swz<double>(idx) << 4;
    uint32_t addr[D];
    double2 a[D];
#pragma unroll
    for (int j = 0; j < D; j++) {
        uint32_t o = 0;
#pragma unroll
        for (int m = 0; m < R; m++) if ((j >> m) & 1) o ^= sb[m];
        addr[j] = tile_s + (bidx ^ o);
        a[j] = lds_amp(addr[j], double2());
    }
    if (MODE == 1) { gate_const<4, 2, 0>(a, mat); gate_const<4, 2, 2>(a, mat + 48); }
    else { gate_mul3<4, 2, 0>(a, mat); gate_mul3<4, 2, 2>(a, mat + 48); }
#pragma unroll
    for (int j = 0; j < D; j++) sts_amp(addr[j], a[j]);
}

// the per-op set-up (op fields, group index deposit, swizzled offsets) of the NEXT op done before the barrier that ends
// the current one: after the barrier only the 16 loads remain
struct Setup { uint32_t bidx, sb[4]; };
__device__ __forceinline__ Setup pair_setup(const DevOp &op, const int tid)
{
    Setup s;
#pragma unroll
    for (int m = 0; m < 4; m++) s.sb[m] = (uint32_t)op.sboff[m] << 4;
    uint32_t idx = 0;
#pragma unroll
    for (int m = 0; m <= 4; m++) idx |= ((uint32_t)tid & op.seg[m]) << m;
    s.bidx = swz<double>(idx) << 4;
    return s;
}
__device__ __forceinline__ void pair_run(double2 *tile, const Setup &st, const double *mat)
{
    const uint32_t tile_s = smem_u32(tile);
    uint32_t addr[16];
    double2 a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t o = 0;
#pragma unroll
        for (int m = 0; m < 4; m++) if ((j >> m) & 1) o ^= st.sb[m];
        addr[j] = tile_s + (st.bidx ^ o);
        a[j] = lds_amp(addr[j], double2());
    }
    gate_mul3<4, 2, 0>(a, mat);
    gate_mul3<4, 2, 2>(a, mat + 48);
#pragma unroll
    for (int j = 0; j < 16; j++) sts_amp(addr[j], a[j]);
}

template <int MODE, int CTAS>
__global__ void __launch_bounds__(128, CTAS) k_pair(const __grid_constant__ CP p, const double *mats_g, int iters, double2 *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    double2 *tile = reinterpret_cast<double2 *>(smem);
    double *mats = reinterpret_cast<double *>(smem + 2048 * 16);   // (192 doubles)
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += 128) tile[swz<double>(i)] = make_double2(1.0 / (1 + i), 0.5 / (1 + i));
    for (int e = tid; e < 192; e += 128) mats[e] = mats_g[e];
    __syncthreads();
    if (MODE == 3) {
        Setup st = pair_setup(p.op[0], tid);
        for (int it = 0; it < iters; it++) {
            const int o = it & 1;
            pair_run(tile, st, p.mats3 + p.off[o]);
            st = pair_setup(p.op[o ^ 1], tid);
            __syncthreads();
        }
    }
    for (int it = 0; it < iters && MODE != 3; it++) {
        const int o = it & 1;   // two ops alternate: the op index is a run-time (uniform) value, as in the pass kernel
        if (MODE == 0) op_pair<double, 2, 2, 2, 128, false>(tile, p.op[o], mats + p.off[o], 11, tid);
        else if (MODE == 1) pair_block<1>(tile, p.op[o], p.mats + p.off[o], tid);
        else pair_block<2>(tile, p.op[o], p.mats3 + p.off[o], tid);
        __syncthreads();
    }
    if (out) for (int i = tid; i < 2048; i += 128) out[(size_t)blockIdx.x * 2048 + i] = tile[swz<double>(i)];
}

static void haar4(std::mt19937_64 &rng, double *u)
{
    std::normal_distribution<double> nd;
    double re[4][4], im[4][4];
    for (auto &row : re) for (double &v : row) v = nd(rng);
    for (auto &row : im) for (double &v : row) v = nd(rng);
    for (int i = 0; i < 4; i++) {
        for (int k = 0; k < i; k++) {
            double pr = 0, pi = 0;
            for (int c = 0; c < 4; c++) { pr += re[k][c] * re[i][c] + im[k][c] * im[i][c]; pi += re[k][c] * im[i][c] - im[k][c] * re[i][c]; }
            for (int c = 0; c < 4; c++) { re[i][c] -= pr * re[k][c] - pi * im[k][c]; im[i][c] -= pr * im[k][c] + pi * re[k][c]; }
        }
        double nrm = 0;
        for (int c = 0; c < 4; c++) nrm += re[i][c] * re[i][c] + im[i][c] * im[i][c];
        nrm = std::sqrt(nrm);
        for (int c = 0; c < 4; c++) { re[i][c] /= nrm; im[i][c] /= nrm; }
    }
    for (int i = 0; i < 4; i++) for (int c = 0; c < 4; c++) { u[2 * (4 * i + c)] = re[i][c]; u[2 * (4 * i + c) + 1] = im[i][c]; }
}

static void fill_op(DevOp &d, const int bits[4])
{
    d = DevOp{};
    d.kind = OP_PAIR; d.k = 2; d.k2 = 2; d.s = 2; d.nfix = 4;
    for (int m = 0; m < 4; m++) d.sboff[m] = (uint16_t)(((1u << bits[m]) ^ (((1u << bits[m]) >> 3) & 7u)));
    std::vector<int> fix(bits, bits + 4);
    for (int m = 0; m < 4; m++) d.tbit[m] = (uint8_t)bits[m];
    std::sort(fix.begin(), fix.end());
    for (int k = 0; k <= 4; k++) {
        const int lo = k == 0 ? 0 : fix[k - 1] - (k - 1), hi = k == 4 ? 16 : fix[k] - k;
        d.seg[k] = (uint16_t)(hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u);
    }
}

template <int MODE, int CTAS>
static double run(const CP &p, const double *d_m, int grid_per_sm, int sms, int iters, size_t smem, double2 *d_out, std::vector<double2> *res)
{
    cudaFuncSetAttribute(k_pair<MODE, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (res) {
        k_pair<MODE, CTAS><<<1, 128, smem>>>(p, d_m, 2, d_out);
        res->resize(2048);
        cudaMemcpy(res->data(), d_out, 2048 * sizeof(double2), cudaMemcpyDeviceToHost);
    }
    const int grid = sms * grid_per_sm;
    float ms = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); k_pair<MODE, CTAS><<<grid, 128, smem>>>(p, d_m, iters, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (double)grid * 2048 * 2 * 32 * iters / (ms * 1e-3) / 1e12;   // "useful" TFLOP/s at 32 flop per amplitude-gate
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const int sms = prop.multiProcessorCount;
    std::mt19937_64 rng(7);
    struct Case { std::string name; int b0[4], b1[4]; };
    std::vector<Case> cases = {{"low (0,1)(3,4) / (2,5)(6,7)", {0, 1, 3, 4}, {2, 5, 6, 7}}, {"mixed (1,7)(4,9) / (0,3)(5,10)", {1, 7, 4, 9}, {0, 3, 5, 10}},
                               {"high (7,9)(8,10) / (5,6)(3,4)", {7, 9, 8, 10}, {5, 6, 3, 4}}};
    for (int k = 0; k < 5; k++) {
        std::vector<int> all = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
        std::shuffle(all.begin(), all.end(), rng);
        Case c;
        char nm[64];
        snprintf(nm, sizeof nm, "random (%d,%d)(%d,%d) / (%d,%d)(%d,%d)", all[0], all[1], all[2], all[3], all[4], all[5], all[6], all[7]);
        c.name = nm;
        for (int i = 0; i < 4; i++) { c.b0[i] = all[i]; c.b1[i] = all[4 + i]; }
        cases.push_back(c);
    }
    double2 *d_out = nullptr;
    cudaMalloc(&d_out, (size_t)2048 * sizeof(double2));
    double *d_m = nullptr;
    cudaMalloc(&d_m, 192 * sizeof(double));
    const size_t smem3 = 66 * 1024, smem4 = 2048 * 16 + 2048;   // 3 CTAs per SM as in the pass kernel (two 32 KB stages); 4-5 if the tile were single-buffered
    printf("%-40s %10s %10s %10s %10s %10s %10s\n", "bit placement (op 0 / op 1)", "smem x3", "const x3", "mul3 x3", "const x4", "mul3 x4", "mul3 pipe");
    for (const Case &cs : cases) {
        CP p{};
        fill_op(p.op[0], cs.b0);
        fill_op(p.op[1], cs.b1);
        p.off[0] = 0; p.off[1] = 96;
        for (int g = 0; g < 4; g++) haar4(rng, p.mats + 48 * g);
        for (int e = 0; e < 64; e++) {
            const double c = p.mats[48 * (e / 16) + 2 * (e % 16)], d = p.mats[48 * (e / 16) + 2 * (e % 16) + 1];
            p.mats3[3 * e] = c; p.mats3[3 * e + 1] = -(c + d); p.mats3[3 * e + 2] = d - c;
        }
        cudaMemcpy(d_m, p.mats, 192 * sizeof(double), cudaMemcpyHostToDevice);
        std::vector<double2> r0, r1, r2;
        const double t0 = run<0, 3>(p, d_m, 3, sms, iters, smem3, d_out, &r0);
        const double t1 = run<1, 3>(p, d_m, 3, sms, iters, smem3, d_out, &r1);
        const double t2 = run<2, 3>(p, d_m, 3, sms, iters, smem3, d_out, &r2);
        const double t3 = run<1, 4>(p, d_m, 4, sms, iters, smem4, d_out, nullptr);
        const double t4 = run<2, 4>(p, d_m, 4, sms, iters, smem4, d_out, nullptr);
        const double t5 = run<3, 3>(p, d_m, 3, sms, iters, smem3, d_out, &r1);
        double e1 = 0, e2 = 0, mx = 0;
        for (int i = 0; i < 2048; i++) {
            mx = std::max(mx, std::max(std::fabs(r0[i].x), std::fabs(r0[i].y)));
            e1 = std::max(e1, std::max(std::fabs(r0[i].x - r1[i].x), std::fabs(r0[i].y - r1[i].y)));
            e2 = std::max(e2, std::max(std::fabs(r0[i].x - r2[i].x), std::fabs(r0[i].y - r2[i].y)));
        }
        printf("%-40s %10.2f %10.2f %10.2f %10.2f %10.2f %10.2f   err const %.1e mul3 %.1e (max |amp| %.2f)\n", cs.name.c_str(), t0, t1, t2, t3, t4, t5, e1, e2, mx);
    }
    return 0;
}
