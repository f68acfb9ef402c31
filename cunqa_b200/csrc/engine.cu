// engine.cu -- State implementation: device memory, op queue, tile-pass launches, measurement.
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "measure_kernels.cuh"
#include "tile_kernel.cuh"
#include "tile_launch.hpp"

namespace cb200 {

static int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

State::State(int n_qubits, int prec, int dev, int nbatch)
    : n(n_qubits), precision(prec), device(dev), batch(nbatch), queue(std::max(1, n_qubits), std::max(1, nbatch), PlanOptions())
{
    try { construct(0, 1, nullptr); }
    catch (...) { release(); throw; }
}

State::State(int n_qubits, int prec, int dev, int nbatch, int rank, int world, const uint8_t *nccl_id)
    : n(n_qubits), precision(prec), device(dev), batch(nbatch), queue(std::max(1, n_qubits), std::max(1, nbatch), PlanOptions())
{
    try { construct(rank, world, nccl_id); }
    catch (...) { release(); throw; }
}

State::State(int n_qubits, int prec, int dev, int rank, int world, ShardGroup *group)
    : n(n_qubits), precision(prec), device(dev), batch(1), queue(std::max(1, n_qubits), 1, PlanOptions())
{
    try { construct(rank, world, nullptr, group); }
    catch (...) { release(); throw; }
}

void State::init_options(int gbits)
{
    n_local_ = n - gbits;
    amp_bytes_ = precision == 64 ? 16 : 8;
    state_bytes_ = amp_bytes_ * ((size_t)batch << n_local_);
    opt.tile_bits = env_int("CB200_TILE_BITS", precision == 64 ? 11 : 13);
    opt.min_run_bits = env_int("CB200_MIN_RUN_BITS", precision == 64 ? 4 : 5);
    opt.max_pass_cost = env_int("CB200_MAX_PASS_COST", 0);
    opt.lookahead = env_int("CB200_LOOKAHEAD", 256);
    opt.lookahead_min_qubits = env_int("CB200_LOOKAHEAD_MIN_QUBITS", 24);
    opt.max_diag_qubits = env_int("CB200_MAX_DIAG_QUBITS", 10);
    opt.fuse = env_int("CB200_FUSE", 1) != 0;
    opt.pair_ops = env_int("CB200_PAIR_OPS", 1) != 0;
    opt.diag_fuse_always = env_int("CB200_DIAG_FUSE_ALWAYS", 1) != 0;
    opt.stage_diag = env_int("CB200_STAGE_DIAG", 0) != 0;
    opt.fan = env_int("CB200_FAN", 1) != 0;
    opt.lin_perm = env_int("CB200_LIN_PERM", 1) != 0;
    if (gbits) opt.global_mask = ((1ull << gbits) - 1ull) << n_local_;
    queue = OpQueue(n, batch, opt);
}

void State::construct(int rank, int world, const uint8_t *nccl_id, ShardGroup *group)
{
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    if (n < 1 || n - gbits > 40 || n > 62) throw InvalidArg("n_qubits out of range");
    if (world > 1 && (n - gbits < 12 || batch != 1)) throw InvalidArg("sharded states need >= 12 local qubits and batch 1");
    if (precision != 64 && precision != 32) throw InvalidArg("precision must be 64 or 32");
    if (batch < 1) throw InvalidArg("batch must be >= 1");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                        "); cunqa_b200 has no CPU fallback");
    if (device < 0 || device >= count) throw InvalidArg("device index out of range");
    CB200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB200_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw CudaError("cunqa_b200 kernels are built for sm_100a only (found sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    num_sms_ = prop.multiProcessorCount;
    max_smem_ = prop.sharedMemPerBlockOptin;
    max_smem_sm_ = prop.sharedMemPerMultiprocessor;
    tma_threads_ = env_int("CB200_TMA_THREADS", 0);
    tma_fold_dims_ = env_int("CB200_TMA_FOLD", 1) != 0;
    tma_refill_after_ = env_int("CB200_TMA_REFILL_AFTER", 0);
    tma_max_ctas_ = env_int("CB200_TMA_MAX_CTAS", 0);
    tma_stages_ = std::max(1, std::min(3, env_int("CB200_TMA_STAGES", 2)));
    diag_stages_ = std::max(1, std::min(2, env_int("CB200_DIAG_STAGES", 1)));
    init_options(gbits);
    ctas_per_sm_ = env_int("CB200_CTAS_PER_SM", 0);
    use_tma_ = env_int("CB200_USE_TMA", 1) != 0;
    const_mats_ = env_int("CB200_CONST_MATS", 1) != 0;
    CB200_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CB200_CUDA(cudaMalloc(&d_state_, state_bytes_));
    CB200_CUDA(cudaMalloc(&d_small_, sizeof(double) * (4 * batch + 64)));
    CB200_CUDA(cudaMalloc(&d_outcome_, sizeof(int) * batch));
    CB200_CUDA(cudaMalloc(&d_counters_, sizeof(uint64_t) * batch));
    CB200_CUDA(cudaMalloc(&d_forced_, sizeof(int8_t) * batch));
    if (world > 1) dist_init(rank, world, nccl_id, group);
    member_of_ = group;
    reset();
}

State::~State() { release(); }

cudaStream_t State::stream() const { return group_ ? group_stream0() : stream_; }

void State::release()
{
    if (device >= 0) cudaSetDevice(device);
    if (stream_) cudaStreamSynchronize(stream_);
    dist_destroy();
    cudaFree(d_state_); cudaFree(d_diag_); cudaFree(d_flags_); cudaFree(d_partial_); cudaFree(d_small_);
    cudaFree(d_outcome_); cudaFree(d_counters_); cudaFree(d_forced_); cudaFree(d_lvl1_); cudaFree(d_lvl2_);
    cudaFree(d_regmats_); cudaFree(d_seeds_); cudaFree(d_samples_);
    d_regmats_ = nullptr; d_seeds_ = nullptr; d_samples_ = nullptr;
    regmats_cap_ = samples_cap_ = 0;
    for (StageSlot &sl : stage_) {
        if (sl.ptr) cudaFreeHost(sl.ptr);
        if (sl.ev) cudaEventDestroy(sl.ev);
        sl = StageSlot();
    }
    d_state_ = d_diag_ = nullptr;
    d_flags_ = nullptr; d_partial_ = d_small_ = d_lvl1_ = d_lvl2_ = nullptr;
    d_outcome_ = nullptr; d_counters_ = nullptr; d_forced_ = nullptr;
    if (stream_) cudaStreamDestroy(stream_);
    stream_ = nullptr;
    if (xstream_) cudaStreamDestroy(xstream_);
    xstream_ = nullptr;
    cudaGetLastError();
}

void State::reset()
{
    if (group_) { queue.reset(); pending_init_ = true; group_each([](State &sh) { sh.reset(); }); return; }
    queue.reset();
    for (OpQueue &q : regq_) { q.reset(); q.fresh = false; }
    pending_init_ = true;  // the write happens at the next flush, once the initial basis state is known
}

// ---- pinned staging ---------------------------------------------------------------------------------
// Uploads of diagonal tables, condition flags and per-register matrices go through page-locked memory so that the
// copy is a real asynchronous DMA and the flush never drains the stream.  Two slots alternate; a slot is reused only
// after the copy that read it has completed (the event is normally long done by then).
void *State::stage(size_t bytes)
{
    stage_cur_ ^= 1;
    StageSlot &sl = stage_[stage_cur_];
    if (sl.busy) { CB200_CUDA(cudaEventSynchronize(sl.ev)); sl.busy = false; }
    if (bytes > sl.cap) {
        if (sl.ptr) cudaFreeHost(sl.ptr);
        sl.ptr = nullptr;
        sl.cap = std::max<size_t>(bytes * 2, 1 << 16);
        CB200_CUDA(cudaMallocHost(&sl.ptr, sl.cap));
    }
    if (!sl.ev) CB200_CUDA(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    return sl.ptr;
}
void State::stage_commit()
{
    StageSlot &sl = stage_[stage_cur_];
    CB200_CUDA(cudaEventRecord(sl.ev, stream_));
    sl.busy = true;
}

void State::use_registers(bool on)
{
    if ((dist || group_) && on) throw Unsupported("independent registers on a sharded state");
    if (on == registers()) return;
    regq_.clear();
    if (on) {
        for (int b = 0; b < batch; b++) { regq_.emplace_back(n, 1, opt); regq_.back().fresh = false; }
        cur_reg_ = 0;
    }
}
static thread_local int tl_bound_register = -1;
void State::bind_register_for_this_thread(int r) { tl_bound_register = r; }
OpQueue &State::rq()
{
    if (regq_.empty()) return queue;
    const int r = tl_bound_register >= 0 ? tl_bound_register : cur_reg_;
    if (r >= (int)regq_.size()) throw InvalidArg("register index out of range");
    return regq_[r];
}

void State::select_register(int r)
{
    if (regq_.empty()) throw InvalidArg("state is not in register mode");
    if (r < 0 || r >= batch) throw InvalidArg("register index out of range");
    cur_reg_ = r;
}

void State::ensure_init()
{
    if (!pending_init_) return;
    pending_init_ = false;
    queue.fresh = false;
    CB200_CUDA(cudaSetDevice(device));
    const uint64_t amps = 1ull << n_local_;
    const uint64_t total = (uint64_t)batch << n_local_;
    const int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)num_sms_ * 16);
    // the basis state |queue.basis> (X gates folded into the initialisation); sharded: it lives on one rank
    const int owner = (int)(queue.basis >> n_local_);
    const uint64_t local = queue.basis & (amps - 1ull);
    const int origin = dist_rank() == owner;
    if (precision == 64) init_zero_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, amps, total, origin, local);
    else init_zero_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, amps, total, origin, local);
    CB200_CUDA(cudaGetLastError());
    stats.launches++;
}

// ---- queueing (forwarded to the host-only OpQueue) ------------------------------------------------
void State::apply_matrix(const int *qubits, int k, const int *ctrls, int nc, const cplx *mat, const uint8_t *flags)
{
    rq().apply_matrix(qubits, k, ctrls, nc, mat, flags);
    maybe_flush();
}
void State::apply_diagonal(const int *qubits, int k, const cplx *diag, const uint8_t *flags)
{
    rq().apply_diagonal(qubits, k, diag, flags);
    maybe_flush();
}
void State::global_phase(double theta, const uint8_t *flags)
{
    rq().global_phase(theta, flags);
    maybe_flush();
}
void State::apply_named(const std::string &name, const int *qubits, int nq, const double *params, int np,
                        const uint8_t *flags)
{
    rq().apply_named(name, qubits, nq, params, np, flags);
    maybe_flush();
}

// ---- TMA tensor maps ------------------------------------------------------------------------------
// The state is described to the TMA unit as a 3-D tensor of scalars: [128-byte row][2^20 rows][rest],
// box = {one row wide, 2^(Lb - amps-per-row-shift) rows, 1} = one contiguous run of 2^Lb amplitudes,
// hardware 128-byte swizzle on the shared-memory side (matches swz<T>() in tile_kernel.cuh).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CB200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !ptr) throw CudaError("cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// {128-byte row, row index in the whole buffer, qa, qb, qc}: box = one run of 2^run_bits amplitudes (<= 256
// rows) times the three high tile qubits qa/qb/qc (size-2 dimensions with stride 2^q amplitudes; unused
// ones are passed as -1).  Hardware 128-byte swizzle on the shared-memory side (= swz<T>() in the kernel).
void State::make_tensor_map(CUtensorMap *out, int run_bits, const int hi_qubits[3], int b0, int nb)
{
    const int aprs = precision == 64 ? 3 : 4;  // log2(amplitudes per 128-byte row)
    const size_t one = amp_bytes_ << n_local_;  // bytes of one state of the batch
    const size_t view_bytes = nb < 0 ? state_bytes_ : one * (size_t)nb;
    void *view_ptr = (char *)d_state_ + (nb < 0 ? 0 : one * (size_t)b0);
    const uint64_t nrows = view_bytes >> 7;
    if (nrows >= (1ull << 31)) throw Unsupported("state too large for one tensor map");
    const cuuint32_t inner = precision == 64 ? 16 : 32;
    cuuint64_t gdim[5] = {inner, nrows, 1, 1, 1};
    cuuint64_t gstride[4] = {128, 128, 128, 128};
    cuuint32_t box[5] = {inner, 1u << (run_bits - aprs), 1, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int d = 0; d < 3; d++) {
        if (hi_qubits[d] < 0) continue;
        gdim[2 + d] = 2;
        box[2 + d] = 2;
        gstride[1 + d] = (cuuint64_t)amp_bytes_ << hi_qubits[d];
    }
    const CUresult r = encode_tiled_fn()(out, precision == 64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                         5, view_ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

// ---- flush: plan + launch -----------------------------------------------------------------------
template <typename T>
void State::flush_t()
{
    std::vector<HostOp> &ops = queue.ops();
    if (ops.empty()) return;
    CB200_CUDA(cudaSetDevice(device));
    launch_ops_t<T>(ops);
    queue.clear_pending();
}

// The pass schedule depends on the STRUCTURE of the fused op list only (kinds, targets, controls), not on the numbers in
// the matrices: a VQE loop that re-sends angles (`upgrade_parameters`, /root/reference/src/quantum_task.cpp:39-108) gets
// the same fused structure every time, so the schedule of the last few structures is kept.
const std::vector<Pass> &State::cached_schedule(const std::vector<HostOp> &ops)
{
    std::vector<uint64_t> sig;
    sig.reserve(ops.size() * 3 + 4);
    sig.push_back(((uint64_t)n_local_ << 48) | ((uint64_t)opt.tile_bits << 40) | ((uint64_t)opt.min_run_bits << 32) |
                  ((uint64_t)opt.max_pass_cost << 16) | (uint64_t)opt.lookahead);
    for (const HostOp &op : ops) {
        uint64_t tm = 0, cm = 0;
        for (int q : op.q) tm |= 1ull << q;
        for (int q : op.ctrl) cm |= 1ull << q;
        sig.push_back(((uint64_t)op.kind << 60) | ((uint64_t)op.q.size() << 52) | (uint64_t)(op.mask_row >= 0));
        sig.push_back(tm);
        sig.push_back(cm);
    }
    for (auto &e : sched_cache_)
        if (e.first == sig) return e.second;
    if (sched_cache_.size() >= 8) sched_cache_.erase(sched_cache_.begin());
    sched_cache_.emplace_back(std::move(sig), schedule_passes(ops, n_local_, opt));
    return sched_cache_.back().second;
}

// plan `ops` (physical qubits < n_local) into tile passes and launch them on this shard
template <typename T>
void State::launch_ops_t(std::vector<HostOp> &ops, int b0, int nb, ChunkPipe *pipe)
{
    using V = typename Vec2<T>::type;
    if (ops.empty()) { if (pipe && pipe->wait_all) pipe->wait_all(); return; }
    const int vbatch = nb < 0 ? batch : nb;
    const size_t view_bytes = nb < 0 ? state_bytes_ : (amp_bytes_ << n_local_) * (size_t)nb;
    void *view_ptr = (char *)d_state_ + (nb < 0 ? 0 : (amp_bytes_ << n_local_) * (size_t)b0);
    const std::vector<Pass> &passes = cached_schedule(ops);
    std::vector<cplx> diag_host;
    std::vector<PassParams<T>> params(passes.size());
    std::string err;
    for (size_t i = 0; i < passes.size(); i++)
        if (!encode_pass<T>(passes[i], ops, n_local_, vbatch, params[i], diag_host, err, opt.enc_flags() | (const_mats_ ? 0 : kEncPlainMats))) throw Unsupported(err);
    // diagonal tables and condition flags: staged in pinned memory, copied asynchronously (no stream drain)
    const std::vector<std::vector<uint8_t>> &mask_rows = rq().mask_rows;
    const size_t diag_bytes = diag_host.size() * sizeof(V);
    const size_t flag_bytes = nb < 0 ? mask_rows.size() * (size_t)batch : 0;
    if (diag_bytes + flag_bytes > 0) {
        if (diag_host.size() > diag_cap_) {
            CB200_CUDA(cudaStreamSynchronize(stream_));
            cudaFree(d_diag_);
            d_diag_ = nullptr;
            diag_cap_ = std::max<size_t>(diag_host.size() * 2, 4096);
            CB200_CUDA(cudaMalloc(&d_diag_, diag_cap_ * sizeof(V)));
        }
        if (flag_bytes > flags_cap_) {
            CB200_CUDA(cudaStreamSynchronize(stream_));
            cudaFree(d_flags_);
            d_flags_ = nullptr;
            flags_cap_ = flag_bytes * 2;
            CB200_CUDA(cudaMalloc(&d_flags_, flags_cap_));
        }
        char *host = (char *)stage(diag_bytes + flag_bytes);
        V *tmp = reinterpret_cast<V *>(host);
        for (size_t i = 0; i < diag_host.size(); i++) { tmp[i].x = (T)diag_host[i].real(); tmp[i].y = (T)diag_host[i].imag(); }
        for (size_t r = 0; r < mask_rows.size() && flag_bytes; r++) std::memcpy(host + diag_bytes + r * batch, mask_rows[r].data(), batch);
        if (diag_bytes) CB200_CUDA(cudaMemcpyAsync(d_diag_, host, diag_bytes, cudaMemcpyHostToDevice, stream_));
        if (flag_bytes) CB200_CUDA(cudaMemcpyAsync(d_flags_, host + diag_bytes, flag_bytes, cudaMemcpyHostToDevice, stream_));
        stage_commit();
        stats.h2d_bytes += diag_bytes + flag_bytes;
    }
    const int aprs = sizeof(V) == 16 ? 3 : 4;
    for (size_t i = 0; i < passes.size(); i++) {
        const PassParams<T> &p = params[i];
        const uint64_t total_tiles = p.tiles_per_state * (uint64_t)vbatch;
        // a chunked exchange is still in flight: the first pass runs chunk by chunk as the chunks arrive -- possible when
        // none of its tile qubits selects the chunk (every tile then lies inside one chunk)
        int n_chunks = 1;
        if (i == 0 && pipe && pipe->wait_all) {
            bool clash = pipe->mask == 0 || pipe->vals.size() < 2;
            for (int j = 0; j < p.t; j++) clash = clash || ((pipe->mask >> p.tile_q[j]) & 1ull);
            if (clash) pipe->wait_all();
            else n_chunks = (int)pipe->vals.size();
        }
        // TMA box: the contiguous run (at most 256 rows of it) plus up to three high tile qubits
        const int run_bits = std::min(p.L, aprs + 8);
        int hi_q[3] = {-1, -1, -1};
        int box_bits = run_bits;
        if (run_bits == p.L && tma_fold_dims_)
            for (int d = 0; d < 3 && p.L + d < p.t; d++) { hi_q[d] = p.tile_q[p.L + d]; box_bits++; }
        const uint32_t buf_stride = (uint32_t)std::max<size_t>(1024, sizeof(V) << p.t);
        int nt = 1 << std::max(5, std::min(8, p.t - kAmpsPerThreadShift));  // 32 / 64 / 128 / 256 threads
        if (tma_threads_ > 0) nt = tma_threads_;
        // Tile buffers per CTA.  Two: the CTA overlaps its own load / math / store (three CTAs of 128 threads per SM).
        // Passes made of register sweeps (diagonal runs, fans: QFT) are latency bound, not FP64 bound: ONE buffer and four
        // CTAs per SM serve them better on big states (QFT-33 500 against 533 ms; QFT-30 unchanged); the FP64-bound lean
        // passes lose with one buffer (QV-30 615 against 578 ms) and take it only on request (CB200_TMA_STAGES=1).
        int stages = std::max(2, tma_stages_);
        if (p.cond_mask == 0 && 16 * nt == (1 << p.t)) {
            if (p.needs_full == 1) stages = diag_stages_;
            else if (p.mat_form != 0 && tma_stages_ == 1) stages = 1;
        }
        size_t smem_tma = (size_t)stages * buf_stride + sizeof(uint64_t) * (stages + 1) + sizeof(uint32_t) * ((((size_t)1 << (p.t - box_bits)) + 3) & ~(size_t)3) +
                          sizeof(T) * (size_t)((p.n_mat + 1) & ~1) + (size_t)((p.n_diag + 1) & ~1) * 8 + (size_t)p.n_stab * sizeof(V) +
                          (size_t)p.n_diag * (4 + 2 * kDiagLut) +
                          (p.fan_op >= 0 ? (((size_t)p.ops[p.fan_op].k << p.ops[p.fan_op].nfix) + (size_t)p.ops[p.fan_op].k * (1 + p.fan_chunks)) * sizeof(V) + 16 : 0) + 16 + 1024;
        int ctas = std::max(1, (stages == 3 ? 256 : stages == 2 ? 384 : 512) / nt);
        {   // staged diagonal sub-tables must not cost a resident CTA: drop them when they would
            const size_t without = smem_tma - (size_t)p.n_stab * sizeof(V);
            const int c_with = std::min<int>(ctas, (int)(max_smem_sm_ / (smem_tma + 1024)));
            const int c_without = std::min<int>(ctas, (int)(max_smem_sm_ / (without + 1024)));
            if (p.n_stab && c_with < c_without) {
                PassParams<T> &pm = params[i];
                for (int o = 0; o < pm.n_ops; o++) pm.ops[o].stage_off = kNotStaged;
                pm.n_stab = 0;
                smem_tma = without;
            }
        }
        ctas = std::max(1, std::min<int>(ctas, (int)(max_smem_sm_ / (smem_tma + 1024))));  // what shared memory allows
        if (tma_max_ctas_ > 0) ctas = std::min(ctas, tma_max_ctas_);
        // (a run of one 128-byte row is enough once the three folded qubits make the box 1 KiB = one swizzle atom)
        const bool tma = use_tma_ && p.L >= aprs + (box_bits >= aprs + 3 ? 0 : 1) && view_bytes >= 128 && smem_tma <= max_smem_;
        if (!tma && p.mat_form != 0) {
            // the generic kernel reads plain matrix elements: encode this pass again without the three-multiplication form
            // (such a pass has no diagonal tables: nothing is appended to the table buffer)
            std::vector<cplx> none;
            if (!encode_pass<T>(passes[i], ops, n_local_, vbatch, params[i], none, err, opt.enc_flags() | kEncPlainMats)) throw Unsupported(err);
        }
        for (int c = 0; c < n_chunks; c++) {
        if (n_chunks > 1) {
            pipe->wait_chunk(c);
            params[i].chunk_mask = pipe->mask;
            params[i].chunk_val = pipe->vals[c];
        }
        if (tma) {
            CUtensorMap tmap;
            make_tensor_map(&tmap, run_bits, hi_q, b0, nb);
            TmaLaunch l;
            // (a pass that runs beside a chunked exchange leaves room for the exchange kernel's CTAs on every SM)
            static const int overlap_ctas = env_int("CB200_OVERLAP_PASS_CTAS", 3);
            const int use_ctas = n_chunks > 1 ? std::max(1, std::min(ctas, overlap_ctas)) : ctas;
            l.grid = (unsigned)std::min<uint64_t>(total_tiles / n_chunks, (uint64_t)num_sms_ * use_ctas);
            l.smem = smem_tma;
            l.stream = stream_;
            l.diag_tabs = d_diag_;
            l.flags = d_flags_;
            l.box_bits = box_bits;
            l.buf_stride = buf_stride;
            l.refill_after = tma_refill_after_;
            l.stages = stages;
            l.device = device;
            cudaError_t e;
            if (nt == 256) e = launch_tile_pass_tma<T, 256>(p, tmap, l);
            else if (nt == 128) e = launch_tile_pass_tma<T, 128>(p, tmap, l);
            else if (nt == 64) e = launch_tile_pass_tma<T, 64>(p, tmap, l);
            else e = launch_tile_pass_tma<T, 32>(p, tmap, l);
            CB200_CUDA(e);
        } else {
            CB200_CUDA(launch_tile_pass_generic<T>(p, view_ptr, d_diag_, d_flags_, num_sms_, ctas_per_sm_, device, stream_));
        }
        CB200_CUDA(cudaGetLastError());
        stats.launches++;
        }
        stats.passes++;
        stats.h2d_bytes += sizeof(PassParams<T>) + (tma ? sizeof(CUtensorMap) : 0);
        stats.bytes += 2ull * view_bytes;
        stats.fused_ops += (uint64_t)p.n_ops;
    }
}
template void State::launch_ops_t<double>(std::vector<HostOp> &, int, int, ChunkPipe *);
template void State::launch_ops_t<float>(std::vector<HostOp> &, int, int, ChunkPipe *);

// ---- independent registers ------------------------------------------------------------------------
static bool same_structure(const std::vector<HostOp> &a, const std::vector<HostOp> &b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) {
        const HostOp &x = a[i], &y = b[i];
        if (x.kind != y.kind || x.q != y.q || x.ctrl != y.ctrl || x.m.size() != y.m.size() || x.mask_row >= 0 || y.mask_row >= 0)
            return false;
    }
    return true;
}

template <typename T>
void State::flush_registers_t()
{
    bool any = false;
    for (OpQueue &q : regq_) any = any || q.pending() > 0;
    if (!any) return;
    CB200_CUDA(cudaSetDevice(device));
    // maximal runs of consecutive registers whose fused op lists have the same structure share their launches
    // (Backend::execute_many orders the registers so that equal circuits are neighbours)
    const int G = (int)regq_.size();
    for (int b0 = 0; b0 < G;) {
        int b1 = b0 + 1;
        if (regq_[b0].pending() == 0) { b0 = b1; continue; }
        const bool plain = regq_[b0].mask_rows.empty();
        while (plain && b1 < G && regq_[b1].mask_rows.empty() && same_structure(regq_[b0].ops(), regq_[b1].ops()) &&
               regq_[b1].l2p == regq_[b0].l2p)
            b1++;
        cur_reg_ = b0;
        bool shared = false;
        static const bool uniform_ok = env_int("CB200_REG_UNIFORM", 1) != 0;   // (0: every register by itself -- A/B switch)
        if (b1 - b0 >= 2 && uniform_ok) {
            // (a layout that depends on the numbers after all -- e.g. a product that happens to be diagonal for one
            // register only -- is detected before anything is launched: those registers then run one by one)
            try { launch_uniform_t<T>(b0, b1 - b0); shared = true; uniform_flushes++; }
            catch (const Unsupported &) { shared = false; }
        }
        if (shared) {
        } else if (b1 - b0 >= 2) {
            for (int g = b0; g < b1; g++) { cur_reg_ = g; launch_ops_t<T>(regq_[g].ops(), g, 1); }
            split_flushes++;
        } else {
            launch_ops_t<T>(regq_[b0].ops(), b0, 1);
            split_flushes++;
        }
        b0 = b1;
    }
    for (OpQueue &q : regq_) q.clear_pending();
}

// every register has the same fused op list up to the numbers in its matrices: ONE schedule, ONE launch per pass over
// the whole batch; register b reads row b of the per-register matrix / diagonal-table buffers
template <typename T>
void State::launch_uniform_t(int b0, int nb)
{
    using V = typename Vec2<T>::type;
    const int G = nb;
    const size_t view_bytes = (amp_bytes_ << n_local_) * (size_t)nb;
    void *view_ptr = (char *)d_state_ + (amp_bytes_ << n_local_) * (size_t)b0;
    std::vector<HostOp> &ops0 = regq_[b0].ops();
    const std::vector<Pass> passes = schedule_passes(ops0, n_local_, opt);
    std::vector<PassParams<T>> params(passes.size());
    std::vector<cplx> diag0;
    std::string err;
    // A pass that the lean FP64 kernel variant serves keeps the fast forms of its matrices (encode_pass: mat_form 1), which
    // the kernel reads from its parameters: such a pass is launched over as many registers at a time as their matrices fit
    // into one parameter block (state b reads mats + b * reg_stride); every other pass reads row b of a matrix buffer in
    // global memory (reg_mats).
    std::vector<int> pflags(passes.size(), opt.enc_flags() | kEncPlainMats);
    for (size_t i = 0; i < passes.size(); i++) {
        if (sizeof(T) == 8 && const_mats_ && use_tma_) {
            std::vector<cplx> none;
            if (!encode_pass<T>(passes[i], ops0, n_local_, G, params[i], none, err, opt.enc_flags())) throw Unsupported(err);
            if (params[i].mat_form == 1 && params[i].n_mat > 0) pflags[i] = opt.enc_flags();
        }
        if (!encode_pass<T>(passes[i], ops0, n_local_, G, params[i], diag0, err, pflags[i])) throw Unsupported(err);
    }
    auto fast = [&](size_t i) { return !(pflags[i] & kEncPlainMats); };
    std::vector<std::vector<T>> fast_mats(passes.size());   // [pass][register * stride]: matrices of the fast passes
    const size_t D = diag0.size();
    std::vector<size_t> mat_off(passes.size() + 1, 0);   // scalar offset of pass i's [G][stride_i] block
    std::vector<uint32_t> mat_stride(passes.size());
    for (size_t i = 0; i < passes.size(); i++) {
        mat_stride[i] = (uint32_t)((params[i].n_mat + 1) & ~1);
        mat_off[i + 1] = mat_off[i] + (fast(i) ? 0 : (size_t)mat_stride[i] * G);
        if (fast(i)) fast_mats[i].assign((size_t)mat_stride[i] * G, (T)0);
    }
    const size_t mats_bytes = mat_off.back() * sizeof(T), diag_bytes = (size_t)G * D * sizeof(V);
    if (mats_bytes > regmats_cap_) {
        CB200_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(d_regmats_);
        d_regmats_ = nullptr;
        regmats_cap_ = mats_bytes * 2;
        CB200_CUDA(cudaMalloc(&d_regmats_, regmats_cap_));
    }
    if ((size_t)G * D > diag_cap_) {
        CB200_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(d_diag_);
        d_diag_ = nullptr;
        diag_cap_ = std::max<size_t>((size_t)G * D * 2, 4096);
        CB200_CUDA(cudaMalloc(&d_diag_, diag_cap_ * sizeof(V)));
    }
    char *host = (char *)stage(mats_bytes + diag_bytes + 16);
    T *hm = reinterpret_cast<T *>(host);
    V *hd = reinterpret_cast<V *>(host + ((mats_bytes + 15) & ~(size_t)15));
    std::vector<cplx> dg;
    PassParams<T> tmp;
    for (int g = 0; g < G; g++) {
        dg.clear();
        std::vector<HostOp> &ops = regq_[b0 + g].ops();
        for (size_t i = 0; i < passes.size(); i++) {
            const PassParams<T> *src = &params[i];
            if (g > 0) {
                if (!encode_pass<T>(passes[i], ops, n_local_, G, tmp, dg, err, pflags[i])) throw Unsupported(err);
                if (tmp.n_mat != params[i].n_mat || tmp.n_ops != params[i].n_ops || tmp.mat_form != params[i].mat_form)
                    throw Unsupported("internal: register passes differ in layout");
                src = &tmp;
            }
            T *dst = fast(i) ? fast_mats[i].data() + (size_t)g * mat_stride[i] : hm + mat_off[i] + (size_t)g * mat_stride[i];
            std::memcpy(dst, src->mats, sizeof(T) * (size_t)src->n_mat);
        }
        const std::vector<cplx> &d = g == 0 ? diag0 : dg;
        if (d.size() != D) throw Unsupported("internal: register diagonal tables differ in layout");
        for (size_t e = 0; e < D; e++) { hd[(size_t)g * D + e].x = (T)d[e].real(); hd[(size_t)g * D + e].y = (T)d[e].imag(); }
    }
    if (mats_bytes) CB200_CUDA(cudaMemcpyAsync(d_regmats_, hm, mats_bytes, cudaMemcpyHostToDevice, stream_));
    if (diag_bytes) CB200_CUDA(cudaMemcpyAsync(d_diag_, hd, diag_bytes, cudaMemcpyHostToDevice, stream_));
    stage_commit();
    stats.h2d_bytes += mats_bytes + diag_bytes;
    const int aprs = sizeof(V) == 16 ? 3 : 4;
    for (size_t i = 0; i < passes.size(); i++) {
        const PassParams<T> &p = params[i];
        const uint64_t total_tiles = p.tiles_per_state * (uint64_t)G;
        const int run_bits = std::min(p.L, aprs + 8);
        int hi_q[3] = {-1, -1, -1};
        int box_bits = run_bits;
        if (run_bits == p.L && tma_fold_dims_)
            for (int d = 0; d < 3 && p.L + d < p.t; d++) { hi_q[d] = p.tile_q[p.L + d]; box_bits++; }
        const uint32_t buf_stride = (uint32_t)std::max<size_t>(1024, sizeof(V) << p.t);
        const int stages = std::max(2, tma_stages_);
        const size_t smem_tma = (size_t)stages * buf_stride + sizeof(uint64_t) * (stages + 1) + sizeof(uint32_t) * ((((size_t)1 << (p.t - box_bits)) + 3) & ~(size_t)3) +
                                sizeof(T) * (size_t)((p.n_mat + 1) & ~1) + (size_t)((p.n_diag + 1) & ~1) * 8 + (size_t)p.n_stab * sizeof(V) +
                                (size_t)p.n_diag * (4 + 2 * kDiagLut) +
                          (p.fan_op >= 0 ? (((size_t)p.ops[p.fan_op].k << p.ops[p.fan_op].nfix) + (size_t)p.ops[p.fan_op].k * (1 + p.fan_chunks)) * sizeof(V) + 16 : 0) + 16 + 1024;
        int nt = 1 << std::max(5, std::min(8, p.t - kAmpsPerThreadShift));
        if (tma_threads_ > 0) nt = tma_threads_;
        int ctas = std::max(1, (stages == 3 ? 256 : 384) / nt);
        ctas = std::max(1, std::min<int>(ctas, (int)(max_smem_sm_ / (smem_tma + 1024))));
        if (tma_max_ctas_ > 0) ctas = std::min(ctas, tma_max_ctas_);
        const bool tma = use_tma_ && p.L >= aprs + (box_bits >= aprs + 3 ? 0 : 1) && view_bytes >= 128 && smem_tma <= max_smem_;
        const T *rm = reinterpret_cast<const T *>(d_regmats_) + mat_off[i];
        if (fast(i) && tma) {
            // registers per launch: as many as fit into the parameter block's matrix storage
            const int K = std::max(1, std::min<int>(G, kMatScalars / (int)mat_stride[i]));
            for (int c0 = 0; c0 < G; c0 += K) {
                const int kc = std::min(K, G - c0);
                PassParams<T> pc = p;
                pc.batch = kc;
                pc.reg_stride = kc > 1 ? (int32_t)mat_stride[i] : 0;
                std::memcpy(pc.mats, fast_mats[i].data() + (size_t)c0 * mat_stride[i], sizeof(T) * (size_t)mat_stride[i] * kc);
                CUtensorMap tmap;
                make_tensor_map(&tmap, run_bits, hi_q, b0 + c0, kc);
                TmaLaunch l;
                l.grid = (unsigned)std::min<uint64_t>(p.tiles_per_state * (uint64_t)kc, (uint64_t)num_sms_ * ctas);
                l.smem = smem_tma;
                l.stream = stream_;
                l.diag_tabs = d_diag_;
                l.flags = d_flags_;
                l.box_bits = box_bits;
                l.buf_stride = buf_stride;
                l.refill_after = tma_refill_after_;
                l.stages = stages;
                l.device = device;
                cudaError_t e;
                if (nt == 256) e = launch_tile_pass_tma<T, 256>(pc, tmap, l);
                else if (nt == 128) e = launch_tile_pass_tma<T, 128>(pc, tmap, l);
                else if (nt == 64) e = launch_tile_pass_tma<T, 64>(pc, tmap, l);
                else e = launch_tile_pass_tma<T, 32>(pc, tmap, l);
                CB200_CUDA(e);
                stats.launches++;
                stats.h2d_bytes += sizeof(PassParams<T>) + sizeof(CUtensorMap);
            }
            stats.passes++;
            stats.bytes += 2ull * view_bytes;
            stats.fused_ops += (uint64_t)p.n_ops * (uint64_t)G;
            continue;
        }
        if (fast(i)) throw Unsupported("internal: fast register pass without the TMA path");
        if (tma) {
            CUtensorMap tmap;
            make_tensor_map(&tmap, run_bits, hi_q, b0, nb);
            TmaLaunch l;
            l.grid = (unsigned)std::min<uint64_t>(total_tiles, (uint64_t)num_sms_ * ctas);
            l.smem = smem_tma;
            l.stream = stream_;
            l.diag_tabs = d_diag_;
            l.flags = d_flags_;
            l.box_bits = box_bits;
            l.buf_stride = buf_stride;
            l.refill_after = tma_refill_after_;
            l.stages = stages;
            l.device = device;
            l.reg_mats = rm;
            l.reg_mat_stride = mat_stride[i];
            l.reg_diag_stride = (uint32_t)D;
            cudaError_t e;
            if (nt == 256) e = launch_tile_pass_tma<T, 256>(p, tmap, l);
            else if (nt == 128) e = launch_tile_pass_tma<T, 128>(p, tmap, l);
            else if (nt == 64) e = launch_tile_pass_tma<T, 64>(p, tmap, l);
            else e = launch_tile_pass_tma<T, 32>(p, tmap, l);
            CB200_CUDA(e);
        } else {
            CB200_CUDA(launch_tile_pass_generic<T>(p, view_ptr, d_diag_, d_flags_, num_sms_, ctas_per_sm_, device, stream_, rm, mat_stride[i], (uint32_t)D));
        }
        stats.passes++;
        stats.launches++;
        stats.h2d_bytes += sizeof(PassParams<T>) + (tma ? sizeof(CUtensorMap) : 0);
        stats.bytes += 2ull * view_bytes;
        stats.fused_ops += (uint64_t)p.n_ops * (uint64_t)G;
    }
}
template void State::launch_uniform_t<double>(int, int);
template void State::launch_uniform_t<float>(int, int);

void State::flush()
{
    if (group_) { facade_flush(); return; }
    ensure_init();
    if (dist) {
        if (precision == 64) flush_dist_t<double>();
        else flush_dist_t<float>();
        return;
    }
    if (!regq_.empty()) {
        if (precision == 64) flush_registers_t<double>();
        else flush_registers_t<float>();
        return;
    }
    if (precision == 64) flush_t<double>();
    else flush_t<float>();
}

void State::sync()
{
    flush();
    if (group_) { group_each([](State &sh) { sh.sync(); }); facade_sync_stats(); return; }
    CB200_CUDA(cudaSetDevice(device));
    CB200_CUDA(cudaStreamSynchronize(stream_));
}

// ---- measurement --------------------------------------------------------------------------------
void State::ensure_scratch(size_t partial_doubles)
{
    if (partial_doubles > partial_cap_) {
        CB200_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(d_partial_);
        partial_cap_ = partial_doubles;
        CB200_CUDA(cudaMalloc(&d_partial_, partial_cap_ * sizeof(double)));
    }
}

// d_all[b] = sum |a|^2 ; d_sel[b] = sum over (index & sel_mask)==sel_mask   (either may be null)
template <typename T>
void State::reduce2_t(uint64_t sel_mask, double *d_all, double *d_sel)
{
    using V = typename Vec2<T>::type;
    const uint64_t amps = 1ull << n_local_;
    int nblk = (int)std::min<uint64_t>((amps + kRedThreads * 8 - 1) / (kRedThreads * 8),
                                       std::max<uint64_t>(1, (uint64_t)num_sms_ * 8 / batch));
    if (nblk < 1) nblk = 1;
    ensure_scratch((size_t)nblk * batch);
    dim3 grid(nblk, batch);
    if (d_all) {
        prob_partial_kernel<V><<<grid, kRedThreads, 0, stream_>>>((const V *)d_state_, amps, 0ull, d_partial_);
        sum_partials_kernel<<<batch, kRedThreads, 0, stream_>>>(d_partial_, nblk, d_all);
        stats.launches += 2;
    }
    if (d_sel) {
        prob_partial_kernel<V><<<grid, kRedThreads, 0, stream_>>>((const V *)d_state_, amps, sel_mask, d_partial_);
        sum_partials_kernel<<<batch, kRedThreads, 0, stream_>>>(d_partial_, nblk, d_sel);
        stats.launches += 2;
    }
    CB200_CUDA(cudaGetLastError());
}

void State::norm(double *out)
{
    if (group_) { flush(); group_out<double>(1, out, [&](State &sh, double *o) { sh.norm(o); }); return; }
    flush();
    if (precision == 64) reduce2_t<double>(0, d_small_, nullptr); else reduce2_t<float>(0, d_small_, nullptr);
    if (dist) dist_allreduce_f64(d_small_, 1);
    CB200_CUDA(cudaMemcpyAsync(out, d_small_, sizeof(double) * batch, cudaMemcpyDeviceToHost, stream_));
    CB200_CUDA(cudaStreamSynchronize(stream_));
    stats.d2h_bytes += sizeof(double) * batch;
}

void State::prob1(int qubit, double *out)
{
    if (qubit < 0 || qubit >= n) throw InvalidArg("qubit out of range");
    if (registers()) throw Unsupported("prob1 over the whole batch is not defined for independent registers");
    if (group_) { flush(); group_out<double>(1, out, [&](State &sh, double *o) { sh.prob1(qubit, o); }); return; }
    flush();
    const int pq1 = queue.l2p[qubit];
    uint64_t sel = 1ull << pq1;
    if (dist && pq1 >= n_local_) {
        // global qubit: the whole shard counts, or none of it
        if ((dist_rank() >> (pq1 - n_local_)) & 1) { if (precision == 64) reduce2_t<double>(0, d_small_ + batch, nullptr); else reduce2_t<float>(0, d_small_ + batch, nullptr); }
        else CB200_CUDA(cudaMemsetAsync(d_small_ + batch, 0, sizeof(double), stream_));
        sel = 0;
    } else
    if (precision == 64) reduce2_t<double>(sel, nullptr, d_small_ + batch); else reduce2_t<float>(sel, nullptr, d_small_ + batch);
    if (dist) dist_allreduce_f64(d_small_ + batch, 1);
    CB200_CUDA(cudaMemcpyAsync(out, d_small_ + batch, sizeof(double) * batch, cudaMemcpyDeviceToHost, stream_));
    CB200_CUDA(cudaStreamSynchronize(stream_));
}

void State::reduced_qubit(int qubit, double out[4], int b)
{
    if (qubit < 0 || qubit >= n) throw InvalidArg("qubit out of range");
    if (b < 0 || b >= batch) throw InvalidArg("state index out of range");
    if (group_) { flush(); group_out<double>(4, out, [&](State &sh, double *o) { sh.reduced_qubit(qubit, o, b); }); queue.l2p = group_l2p(); return; }
    if (dist) localize(qubit);  // rho01 pairs amplitudes that differ in this qubit: bring it into the shard first
    flush();
    const int pq = (regq_.empty() ? queue : regq_[b]).l2p[qubit];
    const uint64_t amps = 1ull << n_local_;
    int nblk = (int)std::min<uint64_t>((amps / 2 + kRedThreads * 8 - 1) / (kRedThreads * 8), (uint64_t)num_sms_ * 8);
    if (nblk < 1) nblk = 1;
    ensure_scratch((size_t)4 * nblk);
    dim3 grid(nblk, 1);  // one state only
    if (precision == 64) rho_partial_kernel<double2><<<grid, kRedThreads, 0, stream_>>>((const double2 *)d_state_ + ((size_t)b << n_local_), amps, pq, d_partial_, (size_t)nblk);
    else rho_partial_kernel<float2><<<grid, kRedThreads, 0, stream_>>>((const float2 *)d_state_ + ((size_t)b << n_local_), amps, pq, d_partial_, (size_t)nblk);
    for (int k = 0; k < 4; k++) sum_partials_kernel<<<1, kRedThreads, 0, stream_>>>(d_partial_ + (size_t)k * nblk, nblk, d_small_ + k);
    CB200_CUDA(cudaGetLastError());
    if (dist) dist_allreduce_f64(d_small_, 4);
    stats.launches += 5;
    stats.bytes += state_bytes_ / (uint64_t)batch;
    CB200_CUDA(cudaMemcpyAsync(out, d_small_, sizeof(double) * 4, cudaMemcpyDeviceToHost, stream_));
    CB200_CUDA(cudaStreamSynchronize(stream_));
    stats.d2h_bytes += sizeof(double) * 4;
}

void State::measure(int qubit, uint64_t seed, const uint64_t *counters, const int8_t *forced, int *bits_out, bool reset_after)
{
    if (qubit < 0 || qubit >= n) throw InvalidArg("qubit out of range");
    if (registers()) throw Unsupported("batch-wide measurement is not defined for independent registers");
    if (group_) {
        flush();
        const bool global = queue.l2p[qubit] >= n_local_;
        int bit = 0;
        group_out<int>(1, &bit, [&](State &sh, int *o) { sh.measure(qubit, seed, counters, forced, o, reset_after); });
        if (bits_out) bits_out[0] = bit;
        // reset of a global qubit that read 1: the kept shards sit on the ranks with that bit set; an X moves them
        if (global && reset_after && bit == 1) apply_named("x", &qubit, 1, nullptr, 0, nullptr);
        return;
    }
    if (dist) {
        // Sharded state: norm and P(q=1) are local reductions + one all-reduce each (every rank then holds the same
        // two doubles and draws the same outcome from the shared counter stream); a local qubit collapses inside the
        // shard, a global one keeps or clears the whole shard -- no amplitude crosses NVLink.
        flush();
        const int pq = queue.l2p[qubit];
        const bool global = pq >= n_local_;
        if (precision == 64) reduce2_t<double>(0, d_small_, nullptr); else reduce2_t<float>(0, d_small_, nullptr);
        if (global) {
            if ((dist_rank() >> (pq - n_local_)) & 1) CB200_CUDA(cudaMemcpyAsync(d_small_ + 1, d_small_, sizeof(double), cudaMemcpyDeviceToDevice, stream_));
            else CB200_CUDA(cudaMemsetAsync(d_small_ + 1, 0, sizeof(double), stream_));
        } else {
            if (precision == 64) reduce2_t<double>(1ull << pq, nullptr, d_small_ + 1); else reduce2_t<float>(1ull << pq, nullptr, d_small_ + 1);
        }
        dist_allreduce_f64(d_small_, 2);
        const uint64_t ctr = counters ? counters[0] : 0;
        CB200_CUDA(cudaMemcpyAsync(d_counters_, &ctr, sizeof(uint64_t), cudaMemcpyHostToDevice, stream_));
        if (forced) CB200_CUDA(cudaMemcpyAsync(d_forced_, forced, sizeof(int8_t), cudaMemcpyHostToDevice, stream_));
        measure_decide_kernel<<<1, 128, 0, stream_>>>(d_small_, d_small_ + 1, 1, seed, d_counters_, forced ? d_forced_ : nullptr,
                                                      d_outcome_, d_small_ + 2);
        const uint64_t work = (1ull << n_local_) >> 1;
        const int blocks = (int)std::min<uint64_t>((work + 255) / 256, (uint64_t)num_sms_ * 16);
        if (global) dist_collapse_global((dist_rank() >> (pq - n_local_)) & 1);
        else if (precision == 64) collapse_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, n_local_, pq, 1, d_outcome_, d_small_ + 2, reset_after ? 1 : 0);
        else collapse_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, n_local_, pq, 1, d_outcome_, d_small_ + 2, reset_after ? 1 : 0);
        CB200_CUDA(cudaGetLastError());
        stats.launches += 4;
        stats.bytes += 2ull * state_bytes_ + 2ull * state_bytes_;
        int out = 0;
        CB200_CUDA(cudaMemcpyAsync(&out, d_outcome_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CB200_CUDA(cudaStreamSynchronize(stream_));
        stats.d2h_bytes += sizeof(int);
        if (bits_out) bits_out[0] = out;
        // reset of a global qubit that read 1: the kept shards sit on the ranks with that bit set; an X moves them
        // (queued like any other gate -- it brings the qubit in with one exchange when it is next flushed)
        if (global && reset_after && out == 1 && !member_of_) apply_named("x", &qubit, 1, nullptr, 0, nullptr);  // (a group's facade queues it)
        return;
    }
    flush();
    const int pq = queue.l2p[qubit];
    const uint64_t sel = 1ull << pq;
    {   // norm and P(q=1) of every state in one read of the batch
        const uint64_t amps = 1ull << n_local_;
        int nblk = (int)std::min<uint64_t>((amps + kRedThreads * 8 - 1) / (kRedThreads * 8), std::max<uint64_t>(1, (uint64_t)num_sms_ * 8 / batch));
        if (nblk < 1) nblk = 1;
        ensure_scratch((size_t)2 * nblk * batch);
        dim3 grid(nblk, batch);
        double *pa = d_partial_, *ps = d_partial_ + (size_t)nblk * batch;
        if (precision == 64) prob2_partial_kernel<double2><<<grid, kRedThreads, 0, stream_>>>((const double2 *)d_state_, amps, sel, pa, ps);
        else prob2_partial_kernel<float2><<<grid, kRedThreads, 0, stream_>>>((const float2 *)d_state_, amps, sel, pa, ps);
        sum_partials_kernel<<<batch, kRedThreads, 0, stream_>>>(pa, nblk, d_small_);
        sum_partials_kernel<<<batch, kRedThreads, 0, stream_>>>(ps, nblk, d_small_ + batch);
        CB200_CUDA(cudaGetLastError());
        stats.launches += 3;
    }
    std::vector<uint64_t> zero_ctr;
    if (!counters) { zero_ctr.assign(batch, 0); counters = zero_ctr.data(); }
    CB200_CUDA(cudaMemcpyAsync(d_counters_, counters, sizeof(uint64_t) * batch, cudaMemcpyHostToDevice, stream_));
    if (forced) CB200_CUDA(cudaMemcpyAsync(d_forced_, forced, sizeof(int8_t) * batch, cudaMemcpyHostToDevice, stream_));
    measure_decide_kernel<<<(batch + 127) / 128, 128, 0, stream_>>>(d_small_, d_small_ + batch, batch, seed, d_counters_,
                                                                     forced ? d_forced_ : nullptr, d_outcome_, d_small_ + 2 * batch);
    const uint64_t work = ((uint64_t)batch << n_local_) >> 1;
    const int blocks = (int)std::min<uint64_t>((work + 255) / 256, (uint64_t)num_sms_ * 16);
    if (precision == 64)
        collapse_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, n_local_, pq, batch, d_outcome_, d_small_ + 2 * batch, reset_after ? 1 : 0);
    else
        collapse_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, n_local_, pq, batch, d_outcome_, d_small_ + 2 * batch, reset_after ? 1 : 0);
    CB200_CUDA(cudaGetLastError());
    stats.launches += 2;
    stats.bytes += state_bytes_ + 2ull * state_bytes_;  // one fused reduction (1*S) + collapse (2*S)
    std::vector<int> tmp;
    int *dst = bits_out;
    if (!dst) { tmp.resize(batch); dst = tmp.data(); }
    CB200_CUDA(cudaMemcpyAsync(dst, d_outcome_, sizeof(int) * batch, cudaMemcpyDeviceToHost, stream_));
    CB200_CUDA(cudaStreamSynchronize(stream_));
    stats.h2d_bytes += (sizeof(uint64_t) + (forced ? 1 : 0)) * (size_t)batch;
    stats.d2h_bytes += sizeof(int) * batch;
}

void State::sample(int b, uint64_t shots, uint64_t seed, uint64_t *out, uint64_t ctr0)
{
    if (b < 0 || b >= batch) throw InvalidArg("state index out of range");
    if (group_) { flush(); group_out<uint64_t>(shots, out, [&](State &sh, uint64_t *o) { sh.sample(b, shots, seed, o, ctr0); }); return; }
    flush();
    const int nl = n_local_;
    const int bits0 = std::min(nl, kChunkBits);
    const int bits1 = std::min(nl - bits0, kChunkBits);
    const uint64_t n1 = 1ull << (nl - bits0);           // level-1 entries (chunk sums)
    const uint64_t n2 = 1ull << (nl - bits0 - bits1);   // level-2 entries
    if (n1 > lvl1_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl1_); lvl1_cap_ = n1; CB200_CUDA(cudaMalloc(&d_lvl1_, n1 * sizeof(double))); }
    if (n2 > lvl2_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl2_); lvl2_cap_ = n2; CB200_CUDA(cudaMalloc(&d_lvl2_, n2 * sizeof(double))); }
    const int g1 = (int)std::min<uint64_t>(n1, (uint64_t)num_sms_ * 16);
    const int g2 = (int)std::min<uint64_t>(n2, (uint64_t)num_sms_ * 16);
    if (std::max<uint64_t>(shots, 1) > samples_cap_) {
        CB200_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(d_samples_);
        d_samples_ = nullptr;
        samples_cap_ = std::max<uint64_t>(shots, 4096);
        CB200_CUDA(cudaMalloc(&d_samples_, samples_cap_ * sizeof(uint64_t)));
    }
    uint64_t *d_out = d_samples_;
    double total_all = -1.0, off_lo = 0.0, off_hi = 0.0;
    uint64_t prefix = 0;
    if (dist) {
        // every rank learns all shard norms, then owns one interval of the global CDF
        const int w = dist_world(), r = dist_rank();
        if (precision == 64) reduce2_t<double>(0, d_small_, nullptr); else reduce2_t<float>(0, d_small_, nullptr);
        double mine = 0.0;
        CB200_CUDA(cudaMemcpyAsync(&mine, d_small_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB200_CUDA(cudaStreamSynchronize(stream_));
        std::vector<double> all(w, 0.0);
        all[r] = mine;
        CB200_CUDA(cudaMemcpyAsync(d_small_ + 8, all.data(), sizeof(double) * w, cudaMemcpyHostToDevice, stream_));
        dist_allreduce_f64(d_small_ + 8, w);
        CB200_CUDA(cudaMemcpyAsync(all.data(), d_small_ + 8, sizeof(double) * w, cudaMemcpyDeviceToHost, stream_));
        CB200_CUDA(cudaStreamSynchronize(stream_));
        total_all = 0.0;
        for (int i = 0; i < w; i++) { if (i == r) off_lo = total_all; total_all += all[i]; }
        off_hi = r == w - 1 ? 1e300 : off_lo + all[r];
        prefix = (uint64_t)r << nl;
    }
    const int sblocks = (int)std::min<uint64_t>((shots * 32 + kRedThreads - 1) / kRedThreads, (uint64_t)num_sms_ * 8);
    if (precision == 64) {
        const double2 *st = (const double2 *)d_state_ + ((size_t)b << nl);
        chunk_sums_amp_kernel<double2><<<g1, kRedThreads, 0, stream_>>>(st, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        top_prefix_kernel<<<1, 32, 0, stream_>>>(d_lvl2_, n2);
        if (shots) sample_kernel<double2><<<std::max(sblocks, 1), kRedThreads, 0, stream_>>>(st, nl, bits0, bits1, d_lvl1_, d_lvl2_, n2, shots, seed, d_out, total_all, off_lo, off_hi, prefix, ctr0);
    } else {
        const float2 *st = (const float2 *)d_state_ + ((size_t)b << nl);
        chunk_sums_amp_kernel<float2><<<g1, kRedThreads, 0, stream_>>>(st, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        top_prefix_kernel<<<1, 32, 0, stream_>>>(d_lvl2_, n2);
        if (shots) sample_kernel<float2><<<std::max(sblocks, 1), kRedThreads, 0, stream_>>>(st, nl, bits0, bits1, d_lvl1_, d_lvl2_, n2, shots, seed, d_out, total_all, off_lo, off_hi, prefix, ctr0);
    }
    cudaError_t e = cudaGetLastError();
    stats.launches += 4;
    stats.bytes += state_bytes_ / batch;
    if (e == cudaSuccess && dist && shots) dist_allreduce_u64(d_out, shots);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(uint64_t) * shots, cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) throw CudaError(std::string("sampling failed: ") + cudaGetErrorString(e));
    stats.d2h_bytes += sizeof(uint64_t) * shots;
    if (!qof(b).identity_perm())
        for (uint64_t s = 0; s < shots; s++) out[s] = qof(b).logical_index(out[s]);
}

void State::broadcast_state0()
{
    if (group_) { flush(); return; }
    flush();
    if (batch == 1) return;
    const uint64_t amps = 1ull << n_local_, total = (uint64_t)batch << n_local_;
    const int blocks = (int)std::min<uint64_t>((total - amps + 255) / 256, (uint64_t)num_sms_ * 16);
    if (precision == 64) broadcast_state0_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_, amps, total);
    else broadcast_state0_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_, amps, total);
    CB200_CUDA(cudaGetLastError());
    stats.launches++;
    stats.bytes += state_bytes_;
}

void State::sample_each(uint64_t seed, const uint64_t *counters, uint64_t *out)
{
    if (group_ || dist) {   // one state: a single draw with the caller's counter
        sample(0, 1, seed, out, counters[0]);
        return;
    }
    if (registers()) throw Unsupported("sample_each is not defined for independent registers (use sample_registers)");
    flush();
    const int nl = n_local_;
    const int bits0 = std::min(nl, kChunkBits);
    const int bits1 = std::min(nl - bits0, kChunkBits);
    const uint64_t n1 = ((uint64_t)batch << nl) >> bits0;   // level-1 entries over the whole batch
    const uint64_t n2 = n1 >> bits1;
    if (n1 > lvl1_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl1_); lvl1_cap_ = n1; CB200_CUDA(cudaMalloc(&d_lvl1_, n1 * sizeof(double))); }
    if (n2 > lvl2_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl2_); lvl2_cap_ = n2; CB200_CUDA(cudaMalloc(&d_lvl2_, n2 * sizeof(double))); }
    const int g1 = (int)std::min<uint64_t>(n1, (uint64_t)num_sms_ * 16);
    const int g2 = (int)std::min<uint64_t>(n2, (uint64_t)num_sms_ * 16);
    uint64_t *d_out = nullptr;
    CB200_CUDA(cudaMalloc(&d_out, sizeof(uint64_t) * batch));
    cudaError_t e = cudaMemcpyAsync(d_counters_, counters, sizeof(uint64_t) * batch, cudaMemcpyHostToDevice, stream_);
    const int sblocks = std::max(1, std::min((batch * 32 + kRedThreads - 1) / kRedThreads, num_sms_ * 8));
    if (precision == 64) {
        chunk_sums_amp_kernel<double2><<<g1, kRedThreads, 0, stream_>>>((const double2 *)d_state_, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        sample_each_kernel<double2><<<sblocks, kRedThreads, 0, stream_>>>((const double2 *)d_state_, nl, bits0, bits1, d_lvl1_, d_lvl2_, batch, seed, d_counters_, d_out);
    } else {
        chunk_sums_amp_kernel<float2><<<g1, kRedThreads, 0, stream_>>>((const float2 *)d_state_, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        sample_each_kernel<float2><<<sblocks, kRedThreads, 0, stream_>>>((const float2 *)d_state_, nl, bits0, bits1, d_lvl1_, d_lvl2_, batch, seed, d_counters_, d_out);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    stats.launches += 3;
    stats.bytes += state_bytes_;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(uint64_t) * batch, cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d_out);
    if (e != cudaSuccess) throw CudaError(std::string("sample_each failed: ") + cudaGetErrorString(e));
    stats.h2d_bytes += sizeof(uint64_t) * batch;
    stats.d2h_bytes += sizeof(uint64_t) * batch;
    if (!queue.identity_perm())
        for (int b = 0; b < batch; b++) out[b] = queue.logical_index(out[b]);
}

void State::sample_registers(uint64_t shots, const uint64_t *seeds, uint64_t *out)
{
    if (dist) throw Unsupported("sample_registers on a sharded state");
    flush();
    if (shots == 0) return;
    const int nl = n_local_;
    const int bits0 = std::min(nl, kChunkBits);
    const int bits1 = std::min(nl - bits0, kChunkBits);
    const uint64_t n1 = ((uint64_t)batch << nl) >> bits0;
    const uint64_t n2 = n1 >> bits1;
    if (n1 > lvl1_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl1_); d_lvl1_ = nullptr; lvl1_cap_ = n1; CB200_CUDA(cudaMalloc(&d_lvl1_, n1 * sizeof(double))); }
    if (n2 > lvl2_cap_) { CB200_CUDA(cudaStreamSynchronize(stream_)); cudaFree(d_lvl2_); d_lvl2_ = nullptr; lvl2_cap_ = n2; CB200_CUDA(cudaMalloc(&d_lvl2_, n2 * sizeof(double))); }
    const size_t total = (size_t)batch * shots;
    if (total > samples_cap_) {
        CB200_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(d_samples_);
        d_samples_ = nullptr;
        samples_cap_ = total;
        CB200_CUDA(cudaMalloc(&d_samples_, total * sizeof(uint64_t)));
    }
    if (!d_seeds_) CB200_CUDA(cudaMalloc(&d_seeds_, sizeof(uint64_t) * batch));
    uint64_t *hs = (uint64_t *)stage(sizeof(uint64_t) * batch);
    std::memcpy(hs, seeds, sizeof(uint64_t) * batch);
    CB200_CUDA(cudaMemcpyAsync(d_seeds_, hs, sizeof(uint64_t) * batch, cudaMemcpyHostToDevice, stream_));
    stage_commit();
    const int g1 = (int)std::min<uint64_t>(n1, (uint64_t)num_sms_ * 16);
    const int g2 = (int)std::min<uint64_t>(n2, (uint64_t)num_sms_ * 16);
    const int sblocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((total * 32 + kRedThreads - 1) / kRedThreads, (uint64_t)num_sms_ * 8));
    if (precision == 64) {
        chunk_sums_amp_kernel<double2><<<g1, kRedThreads, 0, stream_>>>((const double2 *)d_state_, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        sample_regs_kernel<double2><<<sblocks, kRedThreads, 0, stream_>>>((const double2 *)d_state_, nl, bits0, bits1, d_lvl1_, d_lvl2_, batch, shots, d_seeds_, d_samples_);
    } else {
        chunk_sums_amp_kernel<float2><<<g1, kRedThreads, 0, stream_>>>((const float2 *)d_state_, bits0, n1, d_lvl1_);
        chunk_sums_f64_kernel<<<g2, kRedThreads, 0, stream_>>>(d_lvl1_, bits1, n2, d_lvl2_);
        sample_regs_kernel<float2><<<sblocks, kRedThreads, 0, stream_>>>((const float2 *)d_state_, nl, bits0, bits1, d_lvl1_, d_lvl2_, batch, shots, d_seeds_, d_samples_);
    }
    CB200_CUDA(cudaGetLastError());
    stats.launches += 3;
    stats.bytes += state_bytes_;
    CB200_CUDA(cudaMemcpyAsync(out, d_samples_, total * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream_));
    CB200_CUDA(cudaStreamSynchronize(stream_));
    stats.h2d_bytes += sizeof(uint64_t) * batch;
    stats.d2h_bytes += total * sizeof(uint64_t);
    for (int b = 0; b < batch; b++) {
        const OpQueue &q = regq_.empty() ? queue : regq_[b];
        if (q.identity_perm()) continue;
        for (uint64_t s = 0; s < shots; s++) out[(size_t)b * shots + s] = q.logical_index(out[(size_t)b * shots + s]);
    }
}

void State::get_amplitudes(int b, const uint64_t *idx, uint64_t cnt, double *out)
{
    if (b < 0 || b >= batch) throw InvalidArg("state index out of range");
    if (group_) { flush(); group_out<double>(2 * cnt, out, [&](State &sh, double *o) { sh.get_amplitudes(b, idx, cnt, o); }); return; }
    flush();
    if (cnt == 0) return;
    std::vector<uint64_t> phys(cnt);
    const uint64_t dim = 1ull << n;
    const uint64_t local_mask = (1ull << n_local_) - 1ull;
    std::vector<uint8_t> owned(cnt, 1);
    for (uint64_t i = 0; i < cnt; i++) {
        if (idx[i] >= dim) throw InvalidArg("amplitude index out of range");
        phys[i] = qof(b).phys_index(idx[i]);
        if (dist) {
            owned[i] = (int)(phys[i] >> n_local_) == dist_rank();
            phys[i] = owned[i] ? (phys[i] & local_mask) : 0;  // not owned: read any valid slot, zeroed below
        }
    }
    uint64_t *d_idx = nullptr; double2 *d_out = nullptr;
    CB200_CUDA(cudaMalloc(&d_idx, cnt * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&d_out, cnt * sizeof(double2));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, phys.data(), cnt * sizeof(uint64_t), cudaMemcpyHostToDevice, stream_);
    if (e == cudaSuccess) {
        const int blocks = (int)((cnt + 255) / 256);
        if (precision == 64) gather_amps_kernel<double2><<<blocks, 256, 0, stream_>>>((const double2 *)d_state_ + ((size_t)b << n_local_), d_idx, cnt, d_out);
        else gather_amps_kernel<float2><<<blocks, 256, 0, stream_>>>((const float2 *)d_state_ + ((size_t)b << n_local_), d_idx, cnt, d_out);
        e = cudaGetLastError();
        stats.launches++;
    }
    if (e == cudaSuccess && dist) {
        // zero what this rank does not own, then sum over ranks
        std::vector<double2> tmp(cnt);
        e = cudaMemcpyAsync(tmp.data(), d_out, cnt * sizeof(double2), cudaMemcpyDeviceToHost, stream_);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
        for (uint64_t i = 0; i < cnt; i++) if (!owned[i]) tmp[i] = make_double2(0.0, 0.0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_out, tmp.data(), cnt * sizeof(double2), cudaMemcpyHostToDevice, stream_);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
        if (e == cudaSuccess) dist_allreduce_f64(reinterpret_cast<double *>(d_out), 2 * cnt);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, cnt * sizeof(double2), cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d_idx); cudaFree(d_out);
    if (e != cudaSuccess) throw CudaError(std::string("get_amplitudes failed: ") + cudaGetErrorString(e));
    stats.h2d_bytes += cnt * sizeof(uint64_t);
    stats.d2h_bytes += cnt * sizeof(double2);
}

void State::set_amplitudes(int b, const double *in)
{
    if (b < 0 || b >= batch) throw InvalidArg("state index out of range");
    if (group_) { flush(); group_each([&](State &sh) { sh.set_amplitudes(b, in); }); return; }
    flush();
    CB200_CUDA(cudaStreamSynchronize(stream_));
    const uint64_t dim = 1ull << n_local_;
    // logical -> physical layout on the host (a sharded state keeps the slice whose global bits are this rank's), one upload
    std::vector<double2> host(dim);
    const uint64_t full = 1ull << n, mine = (uint64_t)dist_rank();
    for (uint64_t i = 0; i < full; i++) {
        const uint64_t p = qof(b).phys_index(i);
        if ((p >> n_local_) == mine) host[p & (dim - 1)] = make_double2(in[2 * i], in[2 * i + 1]);
    }
    double2 *d_in = nullptr;
    CB200_CUDA(cudaMalloc(&d_in, dim * sizeof(double2)));
    cudaError_t e = cudaMemcpyAsync(d_in, host.data(), dim * sizeof(double2), cudaMemcpyHostToDevice, stream_);
    if (e == cudaSuccess) {
        const int blocks = (int)std::min<uint64_t>((dim + 255) / 256, (uint64_t)num_sms_ * 16);
        if (precision == 64) scatter_amps_kernel<double2><<<blocks, 256, 0, stream_>>>((double2 *)d_state_ + ((size_t)b << n_local_), d_in, dim);
        else scatter_amps_kernel<float2><<<blocks, 256, 0, stream_>>>((float2 *)d_state_ + ((size_t)b << n_local_), d_in, dim);
        e = cudaGetLastError();
        stats.launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d_in);
    if (e != cudaSuccess) throw CudaError(std::string("set_amplitudes failed: ") + cudaGetErrorString(e));
    stats.h2d_bytes += dim * sizeof(double2);
}

void State::probabilities(int b, const int *qubits, int k, double *out)
{
    if (b < 0 || b >= batch) throw InvalidArg("state index out of range");
    if (k < 1 || k > 20) throw InvalidArg("marginal over 1..20 qubits");
    if (group_) { flush(); group_out<double>((size_t)1 << k, out, [&](State &sh, double *o) { sh.probabilities(b, qubits, k, o); }); return; }
    flush();
    for (int i = 0; i < k; i++) if (qubits[i] < 0 || qubits[i] >= n) throw InvalidArg("qubit out of range");
    // local qubits of the list are reduced over the shard; global ones are constant here (the rank's bits)
    std::vector<int> pq, ob;
    uint32_t fixed = 0;
    for (int i = 0; i < k; i++) {
        const int p = qof(b).l2p[qubits[i]];
        if (p < n_local_) { pq.push_back(p); ob.push_back(i); }
        else if ((dist_rank() >> (p - n_local_)) & 1) fixed |= 1u << i;
    }
    const int kl = (int)pq.size();
    pq.insert(pq.end(), ob.begin(), ob.end());  // one upload: positions, then output bits
    pq.resize(2 * (size_t)k + 2, 0);
    int *d_q = nullptr; double *d_o = nullptr;
    CB200_CUDA(cudaMalloc(&d_q, pq.size() * sizeof(int)));
    cudaError_t e = cudaMalloc(&d_o, sizeof(double) << k);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, pq.data(), pq.size() * sizeof(int), cudaMemcpyHostToDevice, stream_);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_o, 0, sizeof(double) << k, stream_);
    if (e == cudaSuccess) {
        const uint64_t amps = 1ull << n_local_;
        const int blocks = (int)std::min<uint64_t>((amps + 255) / 256, (uint64_t)num_sms_ * 16);
        if (precision == 64) marginal_kernel<double2><<<blocks, 256, 0, stream_>>>((const double2 *)d_state_ + ((size_t)b << n_local_), amps, d_q, kl, d_o, d_q + kl, fixed);
        else marginal_kernel<float2><<<blocks, 256, 0, stream_>>>((const float2 *)d_state_ + ((size_t)b << n_local_), amps, d_q, kl, d_o, d_q + kl, fixed);
        e = cudaGetLastError();
        stats.launches++;
    }
    if (e == cudaSuccess && dist) dist_allreduce_f64(d_o, (size_t)1 << k);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, sizeof(double) << k, cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d_q); cudaFree(d_o);
    if (e != cudaSuccess) throw CudaError(std::string("probabilities failed: ") + cudaGetErrorString(e));
}

}  // namespace cb200
