// dmma_pair.cu -- experiment for VERDICT r1 #6: is an FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) register block faster
// than the production DFMA register block for two 2-qubit gates on a shared-memory tile?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cunqa_b200/csrc tools/dmma_pair.cu -o tools/bin/dmma_pair
//
// Both variants apply the same "pair" (gate A on tile bits a0,a1; gate B on tile bits b0,b1) to a 2^11-amplitude
// complex128 tile that stays in shared memory (128-byte XOR swizzle, as the TMA kernel leaves it), `iters` times, with the
// barrier the production kernel has after every pair.  No HBM traffic: this isolates the compute phase.
//   dfma : op_pair<double, 2, 2, 2, 128> from tile_kernel.cuh (a thread owns the 16 amplitudes of a group)
//   dmma : one amplitude per lane, lanes <-> (a0 a1 | b0 b1 | x); gate = 2 DMMAs on the real 8x8 form of the complex 4x4
//          (state rows as the A operand, matrix as the B operand, C fragment = the lane's own amplitude again); between
//          the two gates the lanes are transposed with shuffles; stored in place
// Bit placements: tile bits 0..5 reach distinct shared-memory banks (bits 0..2 = 16-byte chunk, 3..5 = row mod 8 through
// the swizzle), bits >= 6 do not -- lanes that differ only in such bits collide.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "tile_kernel.cuh"

using namespace cb200;

struct MmaOp {
    int a0, a1, b0, b1, x;       // tile-bit positions of the lane bits
    int wb[2];                   // tile bits that select the warp (4 warps)
    int ib[4];                   // tile bits that select the iteration (16 per warp)
    double ua[32], ub[32];       // the two 4x4 complex matrices, row-major (re, im)
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t swzb(int bit) { return swz<double>(1u << bit) << 4; }  // byte offset of one tile bit

__global__ void __launch_bounds__(128, 3) k_dmma(const __grid_constant__ MmaOp op, int iters, double2 *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    double2 *tile = reinterpret_cast<double2 *>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += 128) tile[swz<double>(i)] = make_double2(1.0 / (1 + i), 0.5 / (1 + i));
    __syncthreads();
    // lane constants: B fragment of gate A / gate B.  k = lane & 3 (input amplitude j), n = lane >> 2 = 2 r + c
    const int j = lane & 3, r = lane >> 3, c = (lane >> 2) & 1;
    const double are = op.ua[2 * (4 * r + j)], aim = op.ua[2 * (4 * r + j) + 1];
    const double bre = op.ub[2 * (4 * r + j)], bim = op.ub[2 * (4 * r + j) + 1];
    const double a1 = c ? aim : are, a2 = c ? are : -aim;   // real parts of the input feed a1, imaginary parts a2
    const double b1 = c ? bim : bre, b2 = c ? bre : -bim;
    // addresses: layout A (load) and layout B (store) differ in which gate's bits sit in lane bits 0,1
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    uint32_t rest = ((lane >> 4) & 1 ? swzb(op.x) : 0) ^ ((warp & 1) ? swzb(op.wb[0]) : 0) ^ ((warp & 2) ? swzb(op.wb[1]) : 0);
    const uint32_t la = rest ^ ((lane & 1) ? swzb(op.a0) : 0) ^ ((lane & 2) ? swzb(op.a1) : 0) ^ ((lane & 4) ? swzb(op.b0) : 0) ^ ((lane & 8) ? swzb(op.b1) : 0);
    const uint32_t lb = rest ^ ((lane & 1) ? swzb(op.b0) : 0) ^ ((lane & 2) ? swzb(op.b1) : 0) ^ ((lane & 4) ? swzb(op.a0) : 0) ^ ((lane & 8) ? swzb(op.a1) : 0);
    const uint32_t i0 = swzb(op.ib[0]), i1 = swzb(op.ib[1]), i2 = swzb(op.ib[2]), i3 = swzb(op.ib[3]);
    const int src = (lane & 16) | ((lane & 3) << 2) | ((lane >> 2) & 3);   // lane transposition between the two gates
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 16; g++) {
            const uint32_t off = ((g & 1) ? i0 : 0) ^ ((g & 2) ? i1 : 0) ^ ((g & 4) ? i2 : 0) ^ ((g & 8) ? i3 : 0);
            double2 v = lds_amp(base + (la ^ off), double2());
            double c0 = 0.0, c1 = 0.0;
            dmma(c0, c1, v.x, a1);
            dmma(c0, c1, v.y, a2);
            const double t0 = __shfl_sync(0xffffffffu, c0, src), t1 = __shfl_sync(0xffffffffu, c1, src);
            double d0 = 0.0, d1 = 0.0;
            dmma(d0, d1, t0, b1);
            dmma(d0, d1, t1, b2);
            sts_amp(base + (lb ^ off), make_double2(d0, d1));
        }
        __syncthreads();
    }
    if (out) for (int i = tid; i < 2048; i += 128) out[(size_t)blockIdx.x * 2048 + i] = tile[swz<double>(i)];
}

__global__ void __launch_bounds__(128, 3) k_dfma(const __grid_constant__ DevOp op, const double *mats_g, int iters, double2 *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    double2 *tile = reinterpret_cast<double2 *>(smem);
    double *mats = reinterpret_cast<double *>(smem + 2048 * 16);   // gate A in a slot of 48 scalars, gate B after it (production layout)
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += 128) tile[swz<double>(i)] = make_double2(1.0 / (1 + i), 0.5 / (1 + i));
    if (tid < 64) mats[tid < 32 ? tid : tid + 16] = mats_g[tid];
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        op_pair<double, 2, 2, 2, 128, false>(tile, op, mats, 11, tid);
        __syncthreads();
    }
    if (out) for (int i = tid; i < 2048; i += 128) out[(size_t)blockIdx.x * 2048 + i] = tile[swz<double>(i)];
}

static void haar4(std::mt19937_64 &rng, double *u)  // random unitary 4x4 (Gram-Schmidt of a Ginibre matrix), row-major (re, im)
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

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int dev_sms = 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    dev_sms = prop.multiProcessorCount;
    const int grid = dev_sms * 3;
    const size_t smem = 2048 * 16 + 1024;
    cudaFuncSetAttribute(k_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_dfma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::mt19937_64 rng(7);
    struct Case { const char *name; int bits[4]; };
    std::vector<Case> cases = {{"bank-friendly a=(0,1) b=(3,4)", {0, 1, 3, 4}}, {"mixed a=(1,7) b=(4,9)", {1, 7, 4, 9}},
                               {"one gate high a=(7,9) b=(2,4)", {7, 9, 2, 4}}, {"all high a=(7,9) b=(8,10)", {7, 9, 8, 10}}};
    for (int k = 0; k < 6; k++) {  // placements as a QV pass produces them: 4 distinct bits of 11
        std::vector<int> all = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
        std::shuffle(all.begin(), all.end(), rng);
        static char names[6][32];
        snprintf(names[k], sizeof names[k], "random a=(%d,%d) b=(%d,%d)", all[0], all[1], all[2], all[3]);
        cases.push_back({names[k], {all[0], all[1], all[2], all[3]}});
    }
    double2 *d_out = nullptr;
    cudaMalloc(&d_out, (size_t)grid * 2048 * sizeof(double2));
    double *d_m = nullptr;
    cudaMalloc(&d_m, 64 * sizeof(double));
    printf("%-34s %12s %12s %8s   (TFLOP/s at 32 flop per amplitude-gate; %d CTAs x 128 threads, %d pair sweeps)\n", "bit placement", "dfma TF/s", "dmma TF/s", "ratio", grid, iters);
    double sum_f = 0, sum_m = 0;
    int n_rand = 0;
    for (const Case &cs : cases) {
        MmaOp mo{};
        mo.a0 = cs.bits[0]; mo.a1 = cs.bits[1]; mo.b0 = cs.bits[2]; mo.b1 = cs.bits[3];
        std::vector<int> freeb;
        for (int b = 0; b < 11; b++) if (std::find(cs.bits, cs.bits + 4, b) == cs.bits + 4) freeb.push_back(b);
        mo.x = freeb[0];                                   // lowest free bit: lanes 0-15 / 16-31
        mo.wb[0] = freeb[6]; mo.wb[1] = freeb[5];          // highest free bits pick the warp
        for (int i = 0; i < 4; i++) mo.ib[i] = freeb[1 + i];
        haar4(rng, mo.ua);
        haar4(rng, mo.ub);
        // production op: local bits [0,2) = gate A, [2,4) = gate B; seg masks as encode_pass builds them
        DevOp d{};
        d.kind = OP_PAIR; d.k = 2; d.k2 = 2; d.s = 2; d.nfix = 4;
        for (int m = 0; m < 4; m++) d.sboff[m] = (uint16_t)(((1u << cs.bits[m]) ^ (((1u << cs.bits[m]) >> 3) & 7u)));
        std::vector<int> fix(cs.bits, cs.bits + 4);
        for (int m = 0; m < 4; m++) d.tbit[m] = (uint8_t)cs.bits[m];
        std::sort(fix.begin(), fix.end());
        for (int k = 0; k <= 4; k++) {   // segment masks of the group index, as encode_pass builds them
            const int lo = k == 0 ? 0 : fix[k - 1] - (k - 1), hi = k == 4 ? 16 : fix[k] - k;
            d.seg[k] = (uint16_t)(hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u);
        }
        std::vector<double> hm(64);
        // op_pair's matrices index the local bits (m = 0,1): bit m of the row/column index <-> tbit[m]; the DMMA kernel uses
        // j = lane & 3 with lane bit 0 <-> a0, bit 1 <-> a1: the same convention
        std::copy(mo.ua, mo.ua + 32, hm.begin());
        std::copy(mo.ub, mo.ub + 32, hm.begin() + 32);
        cudaMemcpy(d_m, hm.data(), 64 * sizeof(double), cudaMemcpyHostToDevice);
        // correctness: one sweep of each must agree
        std::vector<double2> r1((size_t)2048), r2((size_t)2048);
        k_dfma<<<1, 128, smem>>>(d, d_m, 1, d_out);
        cudaMemcpy(r1.data(), d_out, 2048 * sizeof(double2), cudaMemcpyDeviceToHost);
        k_dmma<<<1, 128, smem>>>(mo, 1, d_out);
        cudaMemcpy(r2.data(), d_out, 2048 * sizeof(double2), cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < 2048; i++) err = std::max(err, std::max(std::fabs(r1[i].x - r2[i].x), std::fabs(r1[i].y - r2[i].y)));
        if (cudaGetLastError() != cudaSuccess || err > 1e-14) { printf("%-34s MISMATCH %.3e\n", cs.name, err); continue; }
        float ms_f = 0, ms_m = 0;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0); k_dfma<<<grid, 128, smem>>>(d, d_m, iters, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_f, e0, e1);
            cudaEventRecord(e0); k_dmma<<<grid, 128, smem>>>(mo, iters, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_m, e0, e1);
        }
        const double flop = (double)grid * 2048 * 2 * 32 * iters;
        const double tf = flop / (ms_f * 1e-3) / 1e12, tm = flop / (ms_m * 1e-3) / 1e12;
        printf("%-34s %12.2f %12.2f %8.2f\n", cs.name, tf, tm, tm / tf);
        if (!std::string(cs.name).compare(0, 6, "random")) { sum_f += 1 / tf; sum_m += 1 / tm; n_rand++; }
    }
    if (n_rand) printf("%-34s %12.2f %12.2f %8.2f   (harmonic mean over the random placements)\n", "random placements", n_rand / sum_f, n_rand / sum_m, (n_rand / sum_m) / (n_rand / sum_f));
    return 0;
}
