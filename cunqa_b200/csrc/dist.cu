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
#include <array>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
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

// ---- the shards of one state driven by ONE process -----------------------------------------------------
// config.device.target_devices = [d0, d1, ...] reaches a backend as a list (/root/reference/src/utils/helpers/
// net_functions.hpp:198-222, forwarded at AER/aer_helpers.hpp:111-116) and there is one process per vQPU
// (src/cli/setup_qpus.cpp:69-82): a sharded state behind the plugin must therefore live in that one process.  Every shard
// is an ordinary sharded State (same planner, same kernels as the one-process-per-GPU mode) run SPMD on its own worker
// thread bound to its GPU; what differs is the plumbing: peers are reached through cudaDeviceEnablePeerAccess instead of
// CUDA IPC, exchanges are fenced with cross-device stream events instead of NCCL all-reduces, and the few scalars of a
// reduction are summed on the host.
struct GroupAborted : std::runtime_error { GroupAborted() : std::runtime_error("shard group aborted") {} };

class ShardGroup {
public:
    ShardGroup(int n, int precision, const std::vector<int> &devices);
    ~ShardGroup();
    int world() const { return (int)shards_.size(); }
    State &shard(int r) { return *shards_[r]; }
    void each(const std::function<void(int, State &)> &f);
    // collectives: called by every rank, in the same order, from inside each()
    void barrier();
    void fence(int rank, cudaStream_t stream);
    template <typename W> void allreduce(int rank, W *d_buf, size_t count, cudaStream_t stream);
    static constexpr int kMaxChunks = 8;
    cudaEvent_t chunk_event(int rank, int c) { return chunk_events_[rank][c]; }
    uint64_t all_or(int rank, uint64_t v);   // bitwise OR of one host word over the ranks (collective)

private:
    std::vector<std::array<cudaEvent_t, kMaxChunks>> chunk_events_;   // [rank][chunk of the exchange in flight]
    void worker(int rank);
    std::vector<std::unique_ptr<State>> shards_;
    std::vector<int> devices_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_job_, cv_done_, cv_bar_;
    const std::function<void(int, State &)> *job_ = nullptr;
    uint64_t generation_ = 0;
    int remaining_ = 0;
    bool stop_ = false, aborted_ = false;
    std::exception_ptr error_;
    int bar_count_ = 0;
    uint64_t bar_gen_ = 0;
    std::vector<std::array<cudaEvent_t, 2>> events_;   // [rank][fence parity]
    std::vector<uint64_t> fence_no_;
    std::vector<std::vector<uint64_t>> host_;          // [rank]: staging of one all-reduce (8-byte words)
};

struct DistContext {
    int rank = 0, world = 1, gbits = 0;
    ShardGroup *group = nullptr;   // single-process mode (no NCCL, no IPC)
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
    // chunked exchange (pipelined with the tile passes around it): this launch only moves the elements whose top
    // `chunk_bits` rest bits (below the bit that splits a pair's work between its two ranks) equal `chunk`
    int chunk_bits;
    uint32_t chunk;
    int cpos[3];      // ascending positions of the chunk bits in the pair-half index r (below its top bit)
};

template <typename V>
__global__ void __launch_bounds__(256) qubit_swap_kernel(V *__restrict__ mine, const __grid_constant__ SwapPlan sp)
{
    const int rest_bits = sp.n_local - sp.m - 1;  // index bits of one half of one slab
    const uint64_t half = 1ull << rest_bits;
    const int cr_bits = rest_bits - sp.chunk_bits;  // index bits of this launch's chunk of a half
    const uint64_t total = (uint64_t)((1u << sp.m) - 1) << cr_bits;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t wc = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wc < total; wc += stride) {
        // partner of step j is x ^ (j + 1): at every step the ranks pair up, so no GPU's NVLink port is oversubscribed
        const uint32_t y = sp.x ^ ((uint32_t)(wc >> cr_bits) + 1u);
        uint64_t rl = wc & ((1ull << cr_bits) - 1ull);
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (c < sp.chunk_bits) rl = insert0_64(rl, (uint32_t)sp.cpos[c]) | ((uint64_t)((sp.chunk >> c) & 1u) << sp.cpos[c]);
        uint64_t r = rl | (sp.x < y ? 0 : half);
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
    if (dist->group) { dist->group->fence(dist->rank, stream_); return; }
    CB200_NCCL(nccl().AllReduce(dist->d_fence, dist->d_fence, 1, ncclInt, ncclSum, dist->comm, stream_));
    stats.launches++;
}

static SwapPlan make_swap_plan(const DistContext &dc, int n_local, const std::vector<std::pair<int, int>> &pairs)
{
    if ((int)dc.peers.size() != dc.world) throw InvalidArg("sharded state: peer handles were not imported");
    const int m = (int)pairs.size();
    if (m < 1 || m > 4 || m >= n_local) throw InvalidArg("sharded state: 1..4 qubit pairs per exchange");
    for (int a = 0; a < m; a++)
        for (int b = a + 1; b < m; b++)
            if (pairs[a].first == pairs[b].first || pairs[a].second == pairs[b].second)
                throw InvalidArg("sharded state: repeated global or local bit in one exchange");
    SwapPlan sp{};
    sp.m = m;
    sp.n_local = n_local;
    std::vector<int> order(m);
    for (int j = 0; j < m; j++) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return pairs[a].second < pairs[b].second; });
    for (int j = 0; j < m; j++) { sp.lbit[j] = pairs[order[j]].second; sp.perm[j] = order[j]; }
    uint32_t x = 0;
    for (int j = 0; j < m; j++) x |= (uint32_t)((dc.rank >> pairs[j].first) & 1) << j;
    sp.x = x;
    for (uint32_t y = 0; y < (1u << m); y++) {
        int r = dc.rank;
        for (int j = 0; j < m; j++) r = (r & ~(1 << pairs[j].first)) | (int)((y >> j) & 1u) << pairs[j].first;
        sp.peer[y] = dc.peers[r];
    }
    return sp;
}

void State::dist_swap(const std::vector<std::pair<int, int>> &pairs)
{
    const int m = (int)pairs.size();
    SwapPlan sp = make_swap_plan(*dist, n_local_, pairs);
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

// The same exchange in 2^chunk_bits chunks on a second stream, so that the tile passes AFTER it can start on the chunks
// that have arrived while the rest is still crossing NVLink (single-process groups: chunk completion is a CUDA event per
// rank and chunk, waited on across devices).  The chunk is selected by the highest local positions that are neither
// exchanged nor the position that splits a pair's work between its two ranks; `pipe` tells the passes which they are.
void State::dist_swap_chunked(const std::vector<std::pair<int, int>> &pairs, int chunk_bits, ChunkPipe &pipe, uint64_t avoid)
{
    ShardGroup *grp = dist->group;
    const int m = (int)pairs.size();
    SwapPlan sp = make_swap_plan(*dist, n_local_, pairs);
    const int rest_bits = n_local_ - m - 1;
    chunk_bits = std::max(0, std::min(chunk_bits, std::min(3, rest_bits - 8)));
    // every rank must cut the same chunks: the positions to keep clear are the union over the ranks (their next passes
    // differ where ops controlled by global qubits were dropped)
    avoid = grp->all_or(dist->rank, avoid);
    // local positions that are not exchanged, ascending: the top one splits the pair's work, the next chunk_bits select the chunk
    std::vector<int> rest;
    for (int p = 0; p < n_local_; p++) {
        bool ex = false;
        for (auto &pr : pairs) ex = ex || pr.second == p;
        if (!ex) rest.push_back(p);
    }
    // chunk bits: the highest of those positions (below the top one) that are not tile qubits of the pass that follows --
    // a tile must lie inside one chunk for the pass to run chunk by chunk
    std::vector<int> cidx;   // indices into `rest` (= bit positions in the pair-half index r), descending
    for (int i = (int)rest.size() - 2; i >= 6 && (int)cidx.size() < chunk_bits; i--)
        if (!((avoid >> rest[i]) & 1ull)) cidx.push_back(i);
    if ((int)cidx.size() < chunk_bits) { chunk_bits = 0; cidx.clear(); }
    std::sort(cidx.begin(), cidx.end());
    const int K2 = 1 << chunk_bits;
    pipe.mask = 0;
    pipe.vals.assign(K2, 0);
    for (int b = 0; b < chunk_bits; b++) {
        const int pos = rest[cidx[b]];
        sp.cpos[b] = cidx[b];
        pipe.mask |= 1ull << pos;
        for (int c = 0; c < K2; c++) if ((c >> b) & 1) pipe.vals[c] |= 1ull << pos;
    }
    if (!xstream_) CB200_CUDA(cudaStreamCreateWithFlags(&xstream_, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CB200_CUDA(cudaEventCreate(&e0));
    CB200_CUDA(cudaEventCreate(&e1));
    dist_fence();  // every shard quiescent (and every rank has enqueued its waits on the previous exchange's chunk events)
    CB200_CUDA(cudaEventRecord(e0, stream_));
    CB200_CUDA(cudaStreamWaitEvent(xstream_, e0, 0));
    const uint64_t total = (uint64_t)((1u << m) - 1) << (rest_bits - chunk_bits);
    // few CTAs per SM: the exchange has to share the SMs with the tile pass that runs beside it (CB200_SWAP_CTAS_PER_SM)
    static const int swap_ctas = getenv("CB200_SWAP_CTAS_PER_SM") ? std::max(1, atoi(getenv("CB200_SWAP_CTAS_PER_SM"))) : 2;
    const int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)num_sms_ * swap_ctas);
    const int K = 1 << chunk_bits;
    {   // an SM runs kernels side by side only under ONE shared-memory / L1 split: the exchange kernel (no shared memory)
        // asks for the split the tile-pass kernel forces, otherwise every SM drains before switching between the two
        static bool once[16] = {};
        if (!once[device & 15]) {
            cudaFuncSetAttribute(qubit_swap_kernel<double2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(qubit_swap_kernel<float2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaGetLastError();
            once[device & 15] = true;
        }
    }
    for (int c = 0; c < K; c++) {
        sp.chunk_bits = chunk_bits;
        sp.chunk = (uint32_t)c;
        if (precision == 64) qubit_swap_kernel<double2><<<blocks, 256, 0, xstream_>>>((double2 *)d_state_, sp);
        else qubit_swap_kernel<float2><<<blocks, 256, 0, xstream_>>>((float2 *)d_state_, sp);
        CB200_CUDA(cudaGetLastError());
        CB200_CUDA(cudaEventRecord(grp->chunk_event(dist->rank, c), xstream_));
        stats.launches++;
    }
    CB200_CUDA(cudaEventRecord(e1, xstream_));
    dist->swap_events.emplace_back(e0, e1);
    if (dist->swap_events.size() >= 64) dist_swap_ms();
    grp->barrier();   // every rank has recorded its chunk events: they can be waited on now
    // the ranks that write into this shard: the 2^m - 1 partners of the exchange (and this rank's own kernel)
    std::vector<int> writers;
    for (uint32_t y = 0; y < (1u << m); y++) {
        int r = dist->rank;
        for (int j = 0; j < m; j++) r = (r & ~(1 << pairs[j].first)) | (int)((y >> j) & 1u) << pairs[j].first;
        writers.push_back(r);
    }
    cudaStream_t s = stream_;
    pipe.wait_chunk = [grp, writers, s](int c) {
        for (int r : writers) CB200_CUDA(cudaStreamWaitEvent(s, grp->chunk_event(r, c), 0));
    };
    pipe.wait_all = [grp, writers, s, K]() {
        for (int c = 0; c < K; c++)
            for (int r : writers) CB200_CUDA(cudaStreamWaitEvent(s, grp->chunk_event(r, c), 0));
    };
    dist->swaps += m;
    dist->exchanges++;
    dist->swap_bytes += (state_bytes_ >> m) * ((1u << m) - 1);
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
    // single-process groups pipeline the exchange with the pass that follows it (CB200_SWAP_CHUNK_BITS; 0 = one
    // exchange kernel on the main stream as in the one-process-per-GPU mode): the exchange runs in 2^bits chunks on a
    // second stream, with few CTAs per SM and the shared-memory carve-out of the pass kernel so that both can be
    // resident on an SM, and the pass starts on the chunks that have arrived.  Measured on 2 x B200
    // (profiles/exchange_overlap_r02.txt): QFT-32 215.7 -> 206.6 ms (-4 %).  The gain is far below the 17 % a perfect
    // pipeline would give: side by side on the same SMs the two kernels slow each other down (the pass kernel alone holds
    // 3 CTAs x 128 threads x 168 registers = the whole register file); hiding the exchange completely needs the copy
    // engines (peer DMA through a staging buffer) instead of a kernel.
    // Chunks per exchange, re-measured with the round-2 kernels (wall clock of QFT-(33+g), best of 3): an exchange of ONE
    // qubit pair (2 GPUs) is best in 4 chunks (671 against 697 ms in 2 and 724 unchunked), exchanges of two and three pairs
    // (every rank talks to 3 / 7 partners per chunk) in 2 (4 GPUs: 819 against 854 ms; 8 GPUs: 887 against 935 and 938
    // unchunked).  CB200_SWAP_CHUNK_BITS overrides both.
    static const int chunk_bits_cfg = getenv("CB200_SWAP_CHUNK_BITS") ? std::max(0, std::min(3, atoi(getenv("CB200_SWAP_CHUNK_BITS")))) : -1;
    const bool may_chunk = dist->group && n_local_ >= 24;
    std::vector<std::pair<int, int>> group;
    int group_batch = -1;
    ChunkPipe pipe;   // the exchange in flight, if any
    auto exchange = [&](const DistSegment *next_local) {
        if (!group.empty()) {
            if (pipe.wait_all) { pipe.wait_all(); pipe = ChunkPipe(); }   // two exchanges back to back: the first one completes first
            const int chunk_bits = !may_chunk ? 0 : chunk_bits_cfg >= 0 ? chunk_bits_cfg : (group.size() >= 2 ? 1 : 2);
            if (chunk_bits > 0) {
                // (chunked or not is decided by rank-independent facts only: every rank takes the same path)
                // the tile qubits of the pass that will follow the exchange on this rank (same schedule launch_ops_t derives)
                uint64_t avoid = 0;
                if (next_local && !next_local->ops.empty()) {
                    const std::vector<Pass> nxt = schedule_passes(next_local->ops, n_local_, opt);
                    if (!nxt.empty()) for (int q : nxt[0].tile_q) avoid |= 1ull << q;
                }
                dist_swap_chunked(group, chunk_bits, pipe, avoid);
            } else {
                dist_swap(group);
            }
        }
        group.clear();
    };
    for (size_t si = 0; si < plan.segments.size(); si++) {
        const DistSegment &seg = plan.segments[si];
        // the LOCAL segment that follows the exchange being assembled, if the next segment is one
        const DistSegment *after = (si + 1 < plan.segments.size() && !plan.segments[si + 1].is_swap) ? &plan.segments[si + 1] : nullptr;
        if (seg.is_swap) {
            if (seg.batch != group_batch) exchange(nullptr);
            group_batch = seg.batch;
            group.emplace_back(seg.gbit, seg.lbit);
            if ((int)group.size() == max_group) exchange(after);
        } else {
            exchange(&seg);
            std::vector<HostOp> local = seg.ops;
            launch_ops_t<T>(local, 0, -1, pipe.wait_all ? &pipe : nullptr);
            pipe = ChunkPipe();
        }
    }
    exchange(nullptr);
    if (pipe.wait_all) pipe.wait_all();   // an exchange at the very end: the stream's next user sees it complete
    for (int &p : queue.l2p) p = plan.sigma[p];
    queue.clear_pending();
}
template void State::flush_dist_t<double>();
template void State::flush_dist_t<float>();

void State::dist_allreduce_f64(double *d_buf, size_t count)
{
    if (dist->group) { dist->group->allreduce<double>(dist->rank, d_buf, count, stream_); return; }
    CB200_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, dist->comm, stream_));
    stats.launches++;
}
void State::dist_allreduce_u64(uint64_t *d_buf, size_t count)
{
    if (dist->group) { dist->group->allreduce<uint64_t>(dist->rank, d_buf, count, stream_); return; }
    CB200_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclUint64, ncclSum, dist->comm, stream_));
    stats.launches++;
}

void State::dist_init(int rank, int world, const uint8_t *id, ShardGroup *group)
{
    dist = new DistContext();
    dist->rank = rank;
    dist->world = world;
    while ((1 << dist->gbits) < world) dist->gbits++;
    if (group) { dist->group = group; return; }   // peers are filled in once every shard exists (ShardGroup ctor)
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == CB200_DIST_ID_BYTES, "ncclUniqueId size");
    std::memcpy(&uid, id, sizeof(uid));
    CB200_NCCL(nccl().CommInitRank(&dist->comm, world, uid, rank));
    CB200_CUDA(cudaMalloc(&dist->d_fence, sizeof(int)));
    CB200_CUDA(cudaMemset(dist->d_fence, 0, sizeof(int)));
}

void State::dist_destroy()
{
    if (group_) { delete group_; group_ = nullptr; }
    if (!dist) return;
    for (int r = 0; r < (int)dist->peers.size(); r++)
        if (!dist->group && r != dist->rank && dist->peers[r]) cudaIpcCloseMemHandle(dist->peers[r]);
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
int State::dist_world() const { return group_ ? group_->world() : (dist ? dist->world : 1); }
uint64_t State::dist_swaps() const { return group_ ? group_->shard(0).dist_swaps() : (dist ? dist->swaps : 0); }
uint64_t State::dist_swap_bytes() const { return group_ ? group_->shard(0).dist_swap_bytes() : (dist ? dist->swap_bytes : 0); }
// total device time spent in swaps so far (drains the stream)
double State::dist_swap_ms()
{
    if (group_) {
        double ms = 0.0;
        group_->each([&](int r, State &sh) { const double v = sh.dist_swap_ms(); if (r == 0) ms = v; });
        return ms;
    }
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


// ---- ShardGroup ------------------------------------------------------------------------------------
ShardGroup::ShardGroup(int n, int precision, const std::vector<int> &devices) : devices_(devices)
{
    const int w = (int)devices.size();
    for (int a = 0; a < w; a++)
        for (int b = a + 1; b < w; b++)
            if (devices[a] == devices[b]) throw InvalidArg("target_devices lists a GPU twice");
    // peer access between every pair of devices (NVLink / NVSwitch)
    for (int a = 0; a < w; a++) {
        CB200_CUDA(cudaSetDevice(devices[a]));
        for (int b = 0; b < w; b++) {
            if (a == b) continue;
            int can = 0;
            CB200_CUDA(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) throw Unsupported("GPUs " + std::to_string(devices[a]) + " and " + std::to_string(devices[b]) + " have no peer access");
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CB200_CUDA(e);
            cudaGetLastError();
        }
    }
    for (int r = 0; r < w; r++) shards_.emplace_back(new State(n, precision, devices[r], r, w, this));
    for (int r = 0; r < w; r++) {
        State &sh = *shards_[r];
        sh.dist->peers.assign(w, nullptr);
        for (int p = 0; p < w; p++) sh.dist->peers[p] = shards_[p]->device_ptr();
    }
    events_.resize(w);
    fence_no_.assign(w, 0);
    host_.resize(w);
    chunk_events_.resize(w);
    for (int r = 0; r < w; r++) {
        CB200_CUDA(cudaSetDevice(devices[r]));
        for (int k = 0; k < 2; k++) CB200_CUDA(cudaEventCreateWithFlags(&events_[r][k], cudaEventDisableTiming));
        for (int c = 0; c < kMaxChunks; c++) CB200_CUDA(cudaEventCreateWithFlags(&chunk_events_[r][c], cudaEventDisableTiming));
    }
    for (int r = 0; r < w; r++) threads_.emplace_back([this, r] { worker(r); });
}

ShardGroup::~ShardGroup()
{
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
    }
    cv_job_.notify_all();
    for (std::thread &t : threads_) t.join();
    shards_.clear();
    for (size_t r = 0; r < events_.size(); r++) {
        cudaSetDevice(devices_[r]);
        for (int k = 0; k < 2; k++) if (events_[r][k]) cudaEventDestroy(events_[r][k]);
        for (int c = 0; c < kMaxChunks; c++) if (chunk_events_[r][c]) cudaEventDestroy(chunk_events_[r][c]);
    }
}

void ShardGroup::worker(int rank)
{
    cudaSetDevice(devices_[rank]);
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int, State &)> *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_job_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            job = job_;
        }
        try {
            (*job)(rank, *shards_[rank]);
        } catch (const GroupAborted &) {
        } catch (...) {
            std::lock_guard<std::mutex> lk(m_);
            if (!error_) error_ = std::current_exception();
            aborted_ = true;      // ranks waiting in a collective give up instead of waiting for this one
            cv_bar_.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            if (--remaining_ == 0) cv_done_.notify_all();
        }
    }
}

void ShardGroup::each(const std::function<void(int, State &)> &f)
{
    std::unique_lock<std::mutex> lk(m_);
    job_ = &f;
    remaining_ = world();
    generation_++;
    cv_job_.notify_all();
    cv_done_.wait(lk, [&] { return remaining_ == 0; });
    job_ = nullptr;
    aborted_ = false;
    bar_count_ = 0;
    if (error_) {
        std::exception_ptr e = error_;
        error_ = nullptr;
        std::rethrow_exception(e);
    }
}

void ShardGroup::barrier()
{
    std::unique_lock<std::mutex> lk(m_);
    if (aborted_) throw GroupAborted();
    const uint64_t gen = bar_gen_;
    if (++bar_count_ == world()) {
        bar_count_ = 0;
        bar_gen_++;
        cv_bar_.notify_all();
        return;
    }
    cv_bar_.wait(lk, [&] { return bar_gen_ != gen || aborted_; });
    if (bar_gen_ == gen) throw GroupAborted();
}

// every shard's stream waits until every other shard's stream has reached this point.  Two event sets alternate: a rank
// can re-record set k only after passing the barrier of fence k+1, which every rank reaches after enqueueing its waits
// on set k.
void ShardGroup::fence(int rank, cudaStream_t stream)
{
    const int k = (int)(fence_no_[rank]++ & 1);
    CB200_CUDA(cudaEventRecord(events_[rank][k], stream));
    barrier();
    for (int p = 0; p < world(); p++)
        if (p != rank) CB200_CUDA(cudaStreamWaitEvent(stream, events_[p][k], 0));
}

// sum of `count` scalars over the shards, the same bits on every rank (fixed rank order); host-mediated: these are the
// handful of doubles of a norm / probability / marginal, or the shot indices of a sample
uint64_t ShardGroup::all_or(int rank, uint64_t v)
{
    host_[rank].assign(1, v);
    barrier();
    uint64_t out = 0;
    for (int p = 0; p < world(); p++) out |= host_[p][0];
    barrier();
    return out;
}

template <typename W>
void ShardGroup::allreduce(int rank, W *d_buf, size_t count, cudaStream_t stream)
{
    static_assert(sizeof(W) == 8, "8-byte scalars");
    host_[rank].resize(count);
    CB200_CUDA(cudaMemcpyAsync(host_[rank].data(), d_buf, count * sizeof(W), cudaMemcpyDeviceToHost, stream));
    CB200_CUDA(cudaStreamSynchronize(stream));
    barrier();
    std::vector<W> sum(count, W(0));
    for (int p = 0; p < world(); p++) {
        const W *src = reinterpret_cast<const W *>(host_[p].data());
        for (size_t i = 0; i < count; i++) sum[i] += src[i];
    }
    barrier();  // nobody overwrites its staging while another rank still reads it
    CB200_CUDA(cudaMemcpyAsync(d_buf, sum.data(), count * sizeof(W), cudaMemcpyHostToDevice, stream));
    CB200_CUDA(cudaStreamSynchronize(stream));
}
template void ShardGroup::allreduce<double>(int, double *, size_t, cudaStream_t);
template void ShardGroup::allreduce<uint64_t>(int, uint64_t *, size_t, cudaStream_t);

// ---- the facade State of a single-process group ------------------------------------------------------------
State::State(int n_qubits, int prec, const std::vector<int> &devices)
    : n(n_qubits), precision(prec), device(devices.empty() ? 0 : devices[0]), batch(1), queue(std::max(1, n_qubits), 1, PlanOptions())
{
    const int w = (int)devices.size();
    if (w < 2 || (w & (w - 1))) throw InvalidArg("a sharded state needs a power of two (>= 2) of target devices");
    if (precision != 64 && precision != 32) throw InvalidArg("precision must be 64 or 32");
    int gbits = 0;
    while ((1 << gbits) < w) gbits++;
    if (n - gbits < 12 || n > 62) throw InvalidArg("sharded states need >= 12 local qubits");
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) + "); cunqa_b200 has no CPU fallback");
    for (int d : devices) if (d < 0 || d >= count) throw InvalidArg("device index out of range");
    init_options(gbits);
    try { group_ = new ShardGroup(n, precision, devices); }
    catch (...) { release(); throw; }
    reset();
}

void State::group_each_rank(const std::function<void(int, State &)> &f) { group_->each(f); }
void State::group_each(const std::function<void(State &)> &f) { group_->each([&](int, State &sh) { f(sh); }); }
int State::group_world() const { return group_ ? group_->world() : 1; }
std::vector<int> State::group_l2p() { return group_->shard(0).queue.l2p; }
cudaStream_t State::group_stream0() const { return group_->shard(0).stream(); }

void State::facade_sync_stats()
{
    const Stats &s0 = group_->shard(0).stats;
    stats = s0;
}

// gates were queued and fused ONCE on the facade; every shard plans its own segments of the fused list (dist_plan.cpp)
// and runs them, exchanges in lockstep
void State::facade_flush()
{
    if (queue.pending() == 0 && !pending_init_) {
        // nothing to run, but uncontrolled SWAPs may have relabelled qubits since the last flush: the shards map logical
        // qubits themselves (measure, amplitudes, marginals), so their tables follow (the workers are idle here)
        for (int r = 0; r < group_->world(); r++) group_->shard(r).queue.l2p = queue.l2p;
        return;
    }
    const std::vector<HostOp> &ops = queue.ops();
    group_->each([&](int, State &sh) {
        sh.queue.l2p = queue.l2p;
        sh.queue.basis = queue.basis;
        sh.queue.ops() = ops;
        sh.queue.mask_rows.clear();
        sh.flush();
    });
    queue.l2p = group_->shard(0).queue.l2p;
    queue.clear_pending();
    queue.fresh = false;
    pending_init_ = false;
    facade_sync_stats();
}

// a global qubit becomes local: one pairwise exchange with the highest local position (SPMD: every rank calls it)
void State::localize(int qubit)
{
    if (qubit < 0 || qubit >= n) throw InvalidArg("qubit out of range");
    if (group_) { flush(); group_each([&](State &sh) { sh.localize(qubit); }); queue.l2p = group_l2p(); facade_sync_stats(); return; }
    if (!dist) return;
    flush();
    const int pq = queue.l2p[qubit];
    if (pq < n_local_) return;
    const int victim = n_local_ - 1;
    dist_swap({{pq - n_local_, victim}});
    for (int &p : queue.l2p) {
        if (p == victim) p = pq;
        else if (p == pq) p = victim;
    }
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
