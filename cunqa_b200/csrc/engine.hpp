// engine.hpp -- device state + op queue of the B200 state-vector engine.
//
// A State is what the reference gets from AER::AerState (configure/allocate_qubits/initialize/
// apply_*/apply_measure/clear, aer_simulator_adapter.cpp:185-533,887-892,937-952): a register of
// n qubits, here x `batch` independent copies so that many shots / many small vQPUs advance in one
// kernel launch.  Gates are queued, fused and scheduled into tile passes at flush time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <utility>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "opqueue.hpp"

namespace cb200 {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CB200_CUDA(call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            cudaGetLastError(); /* clear the (non-sticky) error so that later launch checks are clean */ \
            throw cb200::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e_));        \
        }                                                                                             \
    } while (0)

struct Stats {
    uint64_t passes = 0, gates = 0, fused_ops = 0, bytes = 0, launches = 0;
    uint64_t relabelled_swaps = 0, h2d_bytes = 0, d2h_bytes = 0;
};

// one chunked qubit exchange as seen by the tile passes that follow it (dist.cu): the exchange moves the shard in 2^c
// chunks on a second stream; the first pass after it runs chunk by chunk as the chunks arrive
struct ChunkPipe {
    uint64_t mask = 0;                        // physical local bits that select the chunk (0: nothing in flight)
    std::vector<uint64_t> vals;               // vals[c]: those bits for chunk c, in the order the chunks are exchanged
    std::function<void(int)> wait_chunk;      // this shard's stream waits until chunk c has arrived from every partner
    std::function<void()> wait_all;
};

struct DistContext;  // dist.cu
class ShardGroup;    // dist.cu: the shards of one state driven by ONE process (one worker thread per GPU)

class State {
public:
    State(int n_qubits, int precision, int device, int batch);
    // sharded over `world` ranks (one process per GPU): n_qubits is the TOTAL width
    State(int n_qubits, int precision, int device, int batch, int rank, int world, const uint8_t *nccl_id);
    // ONE process, several GPUs (config.device.target_devices = [d0, d1, ...], a power of two of them): this object is a
    // facade that owns one shard State per device; gates are queued and fused here once, every other call is forwarded to
    // the shards (SPMD on worker threads; peer access instead of IPC, stream events instead of NCCL fences)
    State(int n_qubits, int precision, const std::vector<int> &devices);
    // shard of a single-process group (constructed by ShardGroup)
    State(int n_qubits, int precision, int device, int rank, int world, ShardGroup *group);
    ~State();
    State(const State &) = delete;
    State &operator=(const State &) = delete;

    void reset();
    // ---- queueing (logical qubits; `flags` = per-state condition for batched feed-forward)
    void apply_matrix(const int *qubits, int k, const int *ctrls, int nc, const cplx *mat, const uint8_t *flags = nullptr);
    void apply_diagonal(const int *qubits, int k, const cplx *diag, const uint8_t *flags = nullptr);
    void apply_named(const std::string &name, const int *qubits, int nq, const double *params, int np,
                     const uint8_t *flags = nullptr);
    void global_phase(double theta, const uint8_t *flags = nullptr);
    void flush();
    void sync();

    // ---- measurement
    void norm(double *out);
    void prob1(int qubit, double *out);
    // reduced density matrix of one qubit of state 0: out = {rho00, rho11, Re rho01, Im rho01}
    void reduced_qubit(int qubit, double out[4], int b = 0);
    void probabilities(int b, const int *qubits, int k, double *out);
    void measure(int qubit, uint64_t seed, const uint64_t *counters, const int8_t *forced, int *bits_out, bool reset_after);
    // u_s = U(seed, ctr0 + s) * norm
    void sample(int b, uint64_t shots, uint64_t seed, uint64_t *out, uint64_t ctr0 = 0);
    // one sample per state of the batch, draw U(seed, counters[b]) (terminal measurement of dynamic shots)
    void sample_each(uint64_t seed, const uint64_t *counters, uint64_t *out);
    void broadcast_state0();  // states 1..batch-1 := state 0
    void get_amplitudes(int b, const uint64_t *idx, uint64_t cnt, double *out);
    void set_amplitudes(int b, const double *in);

    // ---- independent registers (BASELINE config 5: many small vQPU groups batched as separate state vectors per GPU).
    // Every state of the batch becomes its own register with its own op queue and qubit map.  flush() runs all of them
    // in the SAME tile passes when their fused op lists have the same structure (one launch per pass over the whole
    // batch, per-register matrices and diagonal tables), and register by register otherwise.
    void use_registers(bool on);
    bool registers() const { return !regq_.empty(); }
    void select_register(int r);
    // the queue ops currently go to; a host thread may bind itself to one register (bind_register_for_this_thread) so that
    // several threads queue into different registers at the same time
    OpQueue &rq();
    static void bind_register_for_this_thread(int r);  // -1: unbind (follow select_register)
    const OpQueue &qof(int b) const { return regq_.empty() ? queue : regq_[b]; }  // qubit map of state b
    // `shots` samples of every register, u = U(seeds[r], s): out[r * shots + s] (logical indices)
    void sample_registers(uint64_t shots, const uint64_t *seeds, uint64_t *out);
    uint64_t uniform_flushes = 0, split_flushes = 0;  // register-mode flushes that ran batched / register by register

    int n, precision, device, batch;
    PlanOptions opt;
    Stats stats;
    OpQueue queue;  // pending ops + logical->physical qubit map
    DistContext *dist = nullptr;
    int n_local() const { return n_local_; }
    void *device_ptr() { return d_state_; }
    size_t state_bytes() const { return state_bytes_; }
    cudaStream_t stream() const;

    // ---- sharded mode (dist.cu)
    void dist_export(uint8_t *handle_out);
    void dist_import(const uint8_t *handles_all_ranks);
    int dist_rank() const;
    int dist_world() const;
    uint64_t dist_swaps() const;
    uint64_t dist_swap_bytes() const;
    double dist_swap_ms();
    bool is_facade() const { return group_ != nullptr; }
    bool sharded() const { return dist != nullptr || group_ != nullptr; }
    // bring a global qubit into the shard (one pairwise exchange with the highest free local position)
    void localize(int qubit);

private:
    friend class ShardGroup;
    ShardGroup *group_ = nullptr;         // facade only (owned; destroyed in dist_destroy)
    ShardGroup *member_of_ = nullptr;     // shard of a single-process group
    void init_options(int gbits);
    template <typename T, typename F> void group_out(size_t count, T *out, F f);
    void construct(int rank, int world, const uint8_t *nccl_id, ShardGroup *group = nullptr);
    void release();  // frees every device resource (also used when construction fails half-way)
    void dist_init(int rank, int world, const uint8_t *id, ShardGroup *group);
    void dist_destroy();
    void dist_fence();
    void dist_swap(const std::vector<std::pair<int, int>> &pairs);
    void dist_collapse_global(int my_bit);  // measurement of a global qubit: keep-and-rescale or clear this shard  // (global bit, local bit) pairs, one exchange
    void dist_allreduce_f64(double *d_buf, size_t count);
    void dist_allreduce_u64(uint64_t *d_buf, size_t count);
    template <typename T> void flush_dist_t();
    // plan `ops` into tile passes and launch them on states [b0, b0 + nb) of the batch (nb < 0: the whole batch)
    template <typename T> void launch_ops_t(std::vector<HostOp> &ops, int b0 = 0, int nb = -1, ChunkPipe *pipe = nullptr);
    // chunked exchange on a second stream (single-process groups): fills `pipe` for the passes that follow
    void dist_swap_chunked(const std::vector<std::pair<int, int>> &pairs, int chunk_bits, ChunkPipe &pipe, uint64_t avoid);
    cudaStream_t xstream_ = nullptr;
    template <typename T> void flush_registers_t();
    template <typename T> void launch_uniform_t(int b0, int nb);
    // pinned staging for host->device uploads of tables / flags / matrices (two slots, reused once their copy is done)
    const std::vector<Pass> &cached_schedule(const std::vector<HostOp> &ops);
    std::vector<std::pair<std::vector<uint64_t>, std::vector<Pass>>> sched_cache_;
    void *stage(size_t bytes);
    void stage_commit();
    struct StageSlot { void *ptr = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool busy = false; };
    StageSlot stage_[2];
    int stage_cur_ = 0;
    std::vector<OpQueue> regq_;
    int cur_reg_ = 0;
    void *d_regmats_ = nullptr; size_t regmats_cap_ = 0;
    uint64_t *d_seeds_ = nullptr; uint64_t *d_samples_ = nullptr; size_t samples_cap_ = 0;
    template <typename T> void flush_t();
    template <typename T> void reduce2_t(uint64_t sel_mask, double *d_all, double *d_sel);
    void ensure_scratch(size_t partial_doubles);
    void ensure_init();        // lazily writes the initial basis state (reset() only marks it)
    bool pending_init_ = true;
    void facade_flush();
    void facade_sync_stats();
    void group_each(const std::function<void(State &)> &f);
    void group_each_rank(const std::function<void(int, State &)> &f);
    int group_world() const;
    std::vector<int> group_l2p();
    cudaStream_t group_stream0() const;
    void maybe_flush() { if (regq_.empty() && queue.pending() >= 4096) flush(); }

    int n_local_;
    size_t amp_bytes_, state_bytes_;
    void *d_state_ = nullptr;
    cudaStream_t stream_ = nullptr;
    int num_sms_ = 148;
    size_t max_smem_ = 232448, max_smem_sm_ = 233472;
    int tma_threads_ = 0;
    int ctas_per_sm_ = 0;  // 0 = occupancy query (generic kernel only)
    bool use_tma_ = true;
    int diag_stages_ = 1;      // tile buffers per CTA for passes made of register sweeps (engine.cu: launch_ops_t)
    bool const_mats_ = true;   // dense matrices read from the kernel parameters (uniform datapath) where one set serves the launch
    void make_tensor_map(CUtensorMap *out, int run_bits, const int hi_qubits[3], int b0 = 0, int nb = -1);
    bool tma_fold_dims_ = true;
    int tma_stages_ = 2;
    int tma_max_ctas_ = 0;
    int tma_refill_after_ = 0;  // recycle the previous tile's buffer after this op index (-1: before the first op)
    // device scratch
    void *d_diag_ = nullptr; size_t diag_cap_ = 0;       // diagonal tables (elements of V)
    uint8_t *d_flags_ = nullptr; size_t flags_cap_ = 0;  // mask rows
    double *d_partial_ = nullptr; size_t partial_cap_ = 0;
    double *d_small_ = nullptr;   // [4*batch]: norm, p1, scale, (spare)
    int *d_outcome_ = nullptr;    // [batch]
    uint64_t *d_counters_ = nullptr; int8_t *d_forced_ = nullptr;
    double *d_lvl1_ = nullptr, *d_lvl2_ = nullptr; size_t lvl1_cap_ = 0, lvl2_cap_ = 0;
};

// facade: run f(shard, out_r) on every shard; rank 0 writes the caller's buffer (every rank computes the same values)
template <typename T, typename F>
void State::group_out(size_t count, T *out, F f)
{
    const int w = group_world();
    std::vector<std::vector<T>> tmp((size_t)w);
    for (int r = 1; r < w; r++) tmp[r].resize(count > 0 ? count : 1);
    group_each_rank([&](int r, State &sh) { f(sh, r == 0 ? out : tmp[r].data()); });
    facade_sync_stats();
}

}  // namespace cb200
