// dist.cu -- one state sharded over the GPUs of a box (BASELINE config 4).
//
// One process per GPU.  Rank r owns the amplitudes whose top log2(world) physical index bits equal r.
//   * diagonal gates and controls on global qubits: resolved per rank on the host (dist_plan.cpp), no traffic;
//   * a non-diagonal gate on a global qubit: global<->local QUBIT SWAP.  Every rank exchanges half of its shard
//     with rank ^ (1 << gbit).  The exchange is ONE kernel per rank that reads and writes the partner's HBM
//     directly over NVLink/NVSwitch (peer pointers from CUDA IPC), swapping in place -- no staging buffer,
//     which a 128 GiB shard on a 180 GB GPU could not afford.  Each rank moves half of the pair's elements,
//     so both directions of the link are busy.  Stream-ordered NCCL all-reduces fence the kernel on both sides;
//   * norms / probabilities / sampled shots / amplitude reads: local kernels + ncclAllReduce.
// NCCL is resolved with dlopen so that a process which already loaded one (PyTorch) shares it.
#include "../../include/cunqa_b200.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "dist_plan.hpp"
#include "engine.hpp"
#include "tile_kernel.cuh"

namespace cb200 {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl()
{
    static NcclApi api;
    if (!api.lib) {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) throw CudaError(std::string("cannot load NCCL: ") + dlerror());
        auto sym = [&](const char *n) {
            void *p = dlsym(api.lib, n);
            if (!p) throw CudaError(std::string("NCCL symbol missing: ") + n);
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    }
    return api;
}

#define CB200_NCCL(call)                                                                                  \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) throw CudaError(std::string(#call) + " failed: " + nccl().GetErrorString(r_)); \
    } while (0)

struct DistContext {
    int rank = 0, world = 1, gbits = 0;
    ncclComm_t comm = nullptr;
    std::vector<void *> peers;  // shard base pointer of every rank (own entry = local pointer)
    int *d_fence = nullptr;
    uint64_t swaps = 0, exchanges = 0, swap_bytes = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> swap_events;  // fence-to-fence span of every swap
    double swap_ms = 0.0;
};

// In-place exchange of m (global bit, local bit) pairs over peer memory (NVLink P2P), all m at once.
// Write x for this rank's values of the m global bits and y for an amplitude's values of the m local bits: the
// amplitude (rank x, local y) trades places with (rank y, local x), everything else fixed -- a transposition, so
// slabs with y == x stay put and each rank talks to its 2^m - 1 partners directly (one hop through the NVSwitch,
// (2^m - 1)/2^m of the shard sent instead of m/2 for m single exchanges).  Each unordered pair {x, y} is split
// between its two ranks: the smaller x handles the lower half of the remaining index space, the larger the upper.
struct SwapPlan {
    void *peer[16];   // by y (entry x unused)
    int m, n_local;
    int lbit[4];      // ascending
    int perm[4];      // lbit[j] belongs to pair perm[j] (bit perm[j] of x / y)
    uint32_t x;
};

template <typename V>
__global__ void __launch_bounds__(256) qubit_swap_kernel(V *__restrict__ mine, const __grid_constant__ SwapPlan sp)
{
    const int rest_bits = sp.n_local - sp.m - 1;  // index bits of one half of one slab
    const uint64_t half = 1ull << rest_bits;
    const uint64_t total = (uint64_t)((1u << sp.m) - 1) << rest_bits;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        // partner of step j is x ^ (j + 1): at every step the ranks pair up, so no GPU's NVLink port is oversubscribed
        const uint32_t y = sp.x ^ ((uint32_t)(w >> rest_bits) + 1u);
        uint64_t r = (w & (half - 1)) | (sp.x < y ? 0 : half);
        uint64_t mi = r, pi;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < sp.m) mi = insert0_64(mi, (uint32_t)sp.lbit[j]);
        pi = mi;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < sp.m) {
                mi |= (uint64_t)((y >> sp.perm[j]) & 1u) << sp.lbit[j];
                pi |= (uint64_t)((sp.x >> sp.perm[j]) & 1u) << sp.lbit[j];
            }
        V *__restrict__ theirs = (V *)sp.peer[y];
        const V a = mine[mi];
        const V b = theirs[pi];
        mine[mi] = b;
        theirs[pi] = a;
    }
}

// measurement outcome of a GLOBAL qubit (device-side outcome[0], scale[0]): the ranks whose bit equals the outcome keep
// their shard, rescaled; the others hold amplitude 0 from here on
template <typename V>
__global__ void __launch_bounds__(256) collapse_global_kernel(V *__restrict__ shard, uint64_t amps, int my_bit,
                                                               const int *__restrict__ outcome, const double *__restrict__ scale)
{
    const int out = outcome[0];
    if (out < 0) return;
    const bool keep = out == my_bit;
    const double sc = scale[0];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < amps; i += stride) {
        V v = shard[i];
        v.x = keep ? (decltype(v.x))(v.x * sc) : 0;
        v.y = keep ? (decltype(v.y))(v.y * sc) : 0;
        shard[i] = v;
    }
}

void State::dist_collapse_global(int my_bit)
{
    const uint64_t amps = 1ull << n_local_;
    const int blocks = (int)std::min<uint64_t>((amps + 255) / 256, (uint64_t)num_sms_ * 16);
    if (precision == 64) collapse_global_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, amps, my_bit, d_outcome_, d_small_ + 2);
    else collapse_global_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, amps, my_bit, d_outcome_, d_small_ + 2);
    CB200_CUDA(cudaGetLastError());
}

void State::dist_fence()
{
    CB200_NCCL(nccl().AllReduce(dist->d_fence, dist->d_fence, 1, ncclInt, ncclSum, dist->comm, stream_));
    stats.launches++;
}

void State::dist_swap(const std::vector<std::pair<int, int>> &pairs)
{
    if ((int)dist->peers.size() != dist->world) throw InvalidArg("sharded state: peer handles were not imported");
    const int m = (int)pairs.size();
    if (m < 1 || m > 4 || m >= n_local_) throw InvalidArg("sharded state: 1..4 qubit pairs per exchange");
    for (int a = 0; a < m; a++)
        for (int b = a + 1; b < m; b++)
            if (pairs[a].first == pairs[b].first || pairs[a].second == pairs[b].second)
                throw InvalidArg("sharded state: repeated global or local bit in one exchange");
    SwapPlan sp{};
    sp.m = m;
    sp.n_local = n_local_;
    std::vector<int> order(m);
    for (int j = 0; j < m; j++) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return pairs[a].second < pairs[b].second; });
    for (int j = 0; j < m; j++) { sp.lbit[j] = pairs[order[j]].second; sp.perm[j] = order[j]; }
    uint32_t x = 0;
    for (int j = 0; j < m; j++) x |= (uint32_t)((dist->rank >> pairs[j].first) & 1) << j;
    sp.x = x;
    for (uint32_t y = 0; y < (1u << m); y++) {
        int r = dist->rank;
        for (int j = 0; j < m; j++) r = (r & ~(1 << pairs[j].first)) | (int)((y >> j) & 1u) << pairs[j].first;
        sp.peer[y] = dist->peers[r];
    }
    const uint64_t total = (uint64_t)((1u << m) - 1) << (n_local_ - m - 1);
    cudaEvent_t e0, e1;
    CB200_CUDA(cudaEventCreate(&e0));
    CB200_CUDA(cudaEventCreate(&e1));
    dist_fence();  // every shard quiescent
    CB200_CUDA(cudaEventRecord(e0, stream_));
    const int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)num_sms_ * 16);
    if (precision == 64) qubit_swap_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, sp);
    else qubit_swap_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, sp);
    CB200_CUDA(cudaGetLastError());
    dist_fence();  // the partners' writes into this shard have landed
    CB200_CUDA(cudaEventRecord(e1, stream_));
    dist->swap_events.emplace_back(e0, e1);
    if (dist->swap_events.size() >= 64) dist_swap_ms();  // fold the spans into the running total, free the events
    stats.launches++;
    dist->swaps += m;
    dist->exchanges++;
    dist->swap_bytes += (state_bytes_ >> m) * ((1u << m) - 1);  // sent (= received) per rank
}

template <typename T>
void State::flush_dist_t()
{
    std::vector<HostOp> &ops = queue.ops();
    if (ops.empty()) return;
    CB200_CUDA(cudaSetDevice(device));
    const DistPlan plan = plan_distributed(ops, n, n_local_, dist->rank);
    static const int max_group = getenv("CB200_SWAP_GROUP") ? std::max(1, std::min(4, atoi(getenv("CB200_SWAP_GROUP")))) : 4;
    // swaps planned together (same batch id) become one multi-qubit exchange.  Grouping by batch id -- not by adjacency
    // in this rank's segment list -- keeps the number and shape of exchanges identical on every rank: a LOCAL segment
    // that separates two batches on one rank can be empty (all its ops dropped by a global control) on another.
    std::vector<std::pair<int, int>> group;
    int group_batch = -1;
    auto exchange = [&]() {
        if (!group.empty()) dist_swap(group);
        group.clear();
    };
    for (const DistSegment &seg : plan.segments) {
        if (seg.is_swap) {
            if (seg.batch != group_batch) exchange();
            group_batch = seg.batch;
            group.emplace_back(seg.gbit, seg.lbit);
            if ((int)group.size() == max_group) exchange();
        } else {
            exchange();
            std::vector<HostOp> local = seg.ops;
            launch_ops_t<T>(local);
        }
    }
    exchange();
    for (int &p : queue.l2p) p = plan.sigma[p];
    queue.clear_pending();
}
template void State::flush_dist_t<double>();
template void State::flush_dist_t<float>();

void State::dist_allreduce_f64(double *d_buf, size_t count)
{
    CB200_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, dist->comm, stream_));
    stats.launches++;
}
void State::dist_allreduce_u64(uint64_t *d_buf, size_t count)
{
    CB200_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclUint64, ncclSum, dist->comm, stream_));
    stats.launches++;
}

void State::dist_init(int rank, int world, const uint8_t *id)
{
    dist = new DistContext();
    dist->rank = rank;
    dist->world = world;
    while ((1 << dist->gbits) < world) dist->gbits++;
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == CB200_DIST_ID_BYTES, "ncclUniqueId size");
    std::memcpy(&uid, id, sizeof(uid));
    CB200_NCCL(nccl().CommInitRank(&dist->comm, world, uid, rank));
    CB200_CUDA(cudaMalloc(&dist->d_fence, sizeof(int)));
    CB200_CUDA(cudaMemset(dist->d_fence, 0, sizeof(int)));
}

void State::dist_destroy()
{
    if (!dist) return;
    for (int r = 0; r < (int)dist->peers.size(); r++)
        if (r != dist->rank && dist->peers[r]) cudaIpcCloseMemHandle(dist->peers[r]);
    if (dist->comm) nccl().CommDestroy(dist->comm);
    for (auto &ev : dist->swap_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    cudaFree(dist->d_fence);
    delete dist;
    dist = nullptr;
}

void State::dist_export(uint8_t *out)
{
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == CB200_IPC_HANDLE_BYTES, "IPC handle size");
    CB200_CUDA(cudaIpcGetMemHandle(&h, d_state_));
    std::memcpy(out, &h, sizeof(h));
}

void State::dist_import(const uint8_t *all)
{
    dist->peers.assign(dist->world, nullptr);
    for (int r = 0; r < dist->world; r++) {
        if (r == dist->rank) { dist->peers[r] = d_state_; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all + (size_t)r * CB200_IPC_HANDLE_BYTES, sizeof(h));
        CB200_CUDA(cudaIpcOpenMemHandle(&dist->peers[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
}

int State::dist_rank() const { return dist ? dist->rank : 0; }
int State::dist_world() const { return dist ? dist->world : 1; }
uint64_t State::dist_swaps() const { return dist ? dist->swaps : 0; }
uint64_t State::dist_swap_bytes() const { return dist ? dist->swap_bytes : 0; }
// total device time spent in swaps so far (drains the stream)
double State::dist_swap_ms()
{
    if (!dist) return 0.0;
    cudaStreamSynchronize(stream_);
    for (auto &ev : dist->swap_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) dist->swap_ms += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    dist->swap_events.clear();
    return dist->swap_ms;
}

}  // namespace cb200

using namespace cb200;
struct cb200_state { State *impl; };
extern thread_local std::string g_cb200_err;

extern "C" {

int cb200_dist_unique_id(uint8_t *id_out)
{
    try {
        ncclUniqueId uid;
        CB200_NCCL(nccl().GetUniqueId(&uid));
        std::memcpy(id_out, &uid, sizeof(uid));
        return CB200_OK;
    } catch (const std::exception &e) { g_cb200_err = e.what(); return CB200_ERR_NCCL; }
}

int cb200_state_create_dist(cb200_state **out, int n_qubits_total, int precision, int device, int rank, int world,
                            const uint8_t *id)
{
    if (!out || !id) { g_cb200_err = "null argument"; return CB200_ERR_INVALID; }
    *out = nullptr;
    try {
        if (world < 2 || (world & (world - 1)) || rank < 0 || rank >= world) throw InvalidArg("world must be a power of two >= 2 and 0 <= rank < world");
        State *s = new State(n_qubits_total, precision, device, 1, rank, world, id);
        *out = new cb200_state{s};
        return CB200_OK;
    } catch (const InvalidArg &e) { g_cb200_err = e.what(); return CB200_ERR_INVALID; }
    catch (const std::exception &e) { g_cb200_err = e.what(); return CB200_ERR_CUDA; }
}

int cb200_dist_export_handle(cb200_state *s, uint8_t *handle_out)
{
    if (!s || !s->impl || !handle_out) { g_cb200_err = "null argument"; return CB200_ERR_INVALID; }
    try { s->impl->dist_export(handle_out); return CB200_OK; }
    catch (const std::exception &e) { g_cb200_err = e.what(); return CB200_ERR_CUDA; }
}

int cb200_dist_import_handles(cb200_state *s, const uint8_t *handles_all_ranks)
{
    if (!s || !s->impl || !handles_all_ranks || !s->impl->dist) { g_cb200_err = "not a sharded state"; return CB200_ERR_INVALID; }
    try { s->impl->dist_import(handles_all_ranks); return CB200_OK; }
    catch (const std::exception &e) { g_cb200_err = e.what(); return CB200_ERR_CUDA; }
}

int cb200_dist_stats(cb200_state *s, uint64_t out[5])
{
    if (!s || !s->impl) { g_cb200_err = "null handle"; return CB200_ERR_INVALID; }
    out[0] = (uint64_t)s->impl->dist_rank();
    out[1] = (uint64_t)s->impl->dist_world();
    out[2] = s->impl->dist_swaps();
    out[3] = s->impl->dist_swap_bytes();
    out[4] = (uint64_t)(s->impl->dist_swap_ms() * 1000.0);
    return CB200_OK;
}

}  // extern "C"
