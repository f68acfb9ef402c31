// api.cu -- extern "C" surface of libcunqa_b200.so (declared in include/cunqa_b200.h).
#include "../../include/cunqa_b200.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine.hpp"
#include "executor.hpp"

using namespace cb200;

struct cb200_state { State *impl; };
struct cb200_backend { Backend *impl; };

thread_local std::string g_cb200_err;
#define g_err g_cb200_err
namespace {

template <typename F>
int guard(F f)
{
    try {
        f();
        return CB200_OK;
    } catch (const CudaError &e) { g_err = e.what(); return CB200_ERR_CUDA; }
    catch (const Unsupported &e) { g_err = e.what(); return CB200_ERR_UNSUPPORTED; }
    catch (const InvalidArg &e) { g_err = e.what(); return CB200_ERR_INVALID; }
    catch (const std::exception &e) { g_err = e.what(); return CB200_ERR_INVALID; }
}

char *dup_string(const std::string &s)
{
    char *p = (char *)std::malloc(s.size() + 1);
    if (p) std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

void fill_stats(const Stats &st, uint64_t gates, uint64_t relabelled, uint64_t out[8])
{
    out[0] = st.passes; out[1] = gates; out[2] = st.fused_ops; out[3] = st.bytes; out[4] = st.launches;
    out[5] = relabelled; out[6] = st.h2d_bytes; out[7] = st.d2h_bytes;
}
}  // namespace

extern "C" {

const char *cb200_version(void) { return "cunqa_b200 0.2.0 (sm_100a)"; }
const char *cb200_last_error(void) { return g_err.c_str(); }
int cb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
void cb200_free(void *p) { std::free(p); }

int cb200_state_create(cb200_state **out, int n_qubits, int precision, int device, int batch)
{
    if (!out) { g_err = "null output pointer"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] {
        State *s = new State(n_qubits, precision, device, batch);
        *out = new cb200_state{s};
    });
}
int cb200_state_create_sharded(cb200_state **out, int n_qubits, int precision, const int *devices, int n_devices)
{
    if (!out || !devices) { g_err = "null argument"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] {
        State *s = new State(n_qubits, precision, std::vector<int>(devices, devices + n_devices));
        *out = new cb200_state{s};
    });
}
int cb200_state_destroy(cb200_state *s)
{
    if (!s) return CB200_OK;
    delete s->impl;
    delete s;
    return CB200_OK;
}
#define NEED(s) if (!(s) || !(s)->impl) { g_err = "null handle"; return CB200_ERR_INVALID; }

int cb200_state_reset(cb200_state *s) { NEED(s); return guard([&] { s->impl->reset(); }); }
int cb200_state_num_qubits(const cb200_state *s) { return (s && s->impl) ? s->impl->n : -1; }
int cb200_state_batch(const cb200_state *s) { return (s && s->impl) ? s->impl->batch : -1; }

int cb200_apply_matrix(cb200_state *s, const int *qubits, int k, const int *ctrls, int nc, const double *mat)
{
    NEED(s);
    return guard([&] { s->impl->apply_matrix(qubits, k, ctrls, nc, reinterpret_cast<const cplx *>(mat)); });
}
int cb200_apply_diagonal(cb200_state *s, const int *qubits, int k, const double *diag)
{
    NEED(s);
    return guard([&] { s->impl->apply_diagonal(qubits, k, reinterpret_cast<const cplx *>(diag)); });
}
int cb200_apply_gate(cb200_state *s, const char *name, const int *qubits, int nq, const double *params, int np)
{
    NEED(s);
    return guard([&] { s->impl->apply_named(name ? name : "", qubits, nq, params, np); });
}
int cb200_apply_gate_masked(cb200_state *s, const char *name, const int *qubits, int nq, const double *params, int np,
                            const uint8_t *flags_host)
{
    NEED(s);
    return guard([&] { s->impl->apply_named(name ? name : "", qubits, nq, params, np, flags_host); });
}
int cb200_flush(cb200_state *s) { NEED(s); return guard([&] { s->impl->flush(); }); }
int cb200_sync(cb200_state *s) { NEED(s); return guard([&] { s->impl->sync(); }); }

int cb200_norm(cb200_state *s, double *out) { NEED(s); return guard([&] { s->impl->norm(out); }); }
int cb200_prob1(cb200_state *s, int qubit, double *out) { NEED(s); return guard([&] { s->impl->prob1(qubit, out); }); }
int cb200_probabilities(cb200_state *s, int b, const int *qubits, int k, double *out)
{
    NEED(s);
    return guard([&] { s->impl->probabilities(b, qubits, k, out); });
}
int cb200_measure(cb200_state *s, int qubit, uint64_t seed, const uint64_t *counters, const int8_t *forced, int *bits_out)
{
    NEED(s);
    return guard([&] { s->impl->measure(qubit, seed, counters, forced, bits_out, false); });
}
int cb200_reset_qubit(cb200_state *s, int qubit, uint64_t seed, const uint64_t *counters, const int8_t *forced)
{
    NEED(s);
    return guard([&] { s->impl->measure(qubit, seed, counters, forced, nullptr, true); });
}
int cb200_sample(cb200_state *s, int b, uint64_t shots, uint64_t seed, uint64_t *out)
{
    NEED(s);
    return guard([&] { s->impl->sample(b, shots, seed, out); });
}
int cb200_get_amplitudes(cb200_state *s, int b, const uint64_t *idx, uint64_t n, double *out)
{
    NEED(s);
    return guard([&] { s->impl->get_amplitudes(b, idx, n, out); });
}
int cb200_set_amplitudes(cb200_state *s, int b, const double *in)
{
    NEED(s);
    return guard([&] { s->impl->set_amplitudes(b, in); });
}

void *cb200_state_stream(cb200_state *s) { return (s && s->impl) ? (void *)s->impl->stream() : nullptr; }
void *cb200_backend_stream(cb200_backend *b) { return (b && b->impl) ? b->impl->stream_ptr() : nullptr; }

int cb200_stats(const cb200_state *s, uint64_t stats[8])
{
    NEED(s);
    fill_stats(s->impl->stats, s->impl->queue.gates, s->impl->queue.relabelled_swaps, stats);
    return CB200_OK;
}
int cb200_plan_json(const char *task_json, const char *options_json, char **plan_json_out)
{
    if (!task_json || !plan_json_out) { g_err = "null argument"; return CB200_ERR_INVALID; }
    *plan_json_out = nullptr;
    return guard([&] { *plan_json_out = dup_string(plan_task_json(task_json, options_json ? options_json : "")); });
}
int cb200_update_task_json(const char *task_json, const char *update_json, char **out)
{
    if (!task_json || !update_json || !out) { g_err = "null argument"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] { *out = dup_string(update_task_json(task_json, update_json)); });
}
int cb200_qasm2_to_json(const char *qasm, char **out)
{
    if (!qasm || !out) { g_err = "null argument"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] { *out = dup_string(qasm2_to_json(qasm)); });
}
int cb200_gate_matrix(const char *name, int nq, const double *params, int np, double *out_reim)
{
    if (!name || !out_reim || nq < 1 || nq > 6) { g_err = "bad argument"; return CB200_ERR_INVALID; }
    return guard([&] {
        GateSpec g;
        std::string err;
        if (!lookup_gate(name, nq, params, np, g, err)) throw Unsupported(err);
        if (g.is_global_phase) throw Unsupported("gp has no matrix");
        HostOp op;
        op.kind = HostOp::DENSE;
        const int nc = nq - g.n_targets;
        std::vector<int> all;
        for (int i = 0; i < nq; i++) all.push_back(i);
        for (int i = 0; i < g.n_targets; i++) op.q.push_back(nc + i);
        for (int i = 0; i < nc; i++) op.ctrl.push_back(i);
        op.m = g.mat;
        const std::vector<cplx> M = embed_dense(op, all);
        std::memcpy(out_reim, M.data(), M.size() * sizeof(cplx));
    });
}

int cb200_backend_create(cb200_backend **out, int device)
{
    if (!out) { g_err = "null output pointer"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] { *out = new cb200_backend{new Backend(device)}; });
}
int cb200_backend_create_multi(cb200_backend **out, const int *devices, int n_devices)
{
    if (!out || !devices || n_devices < 1) { g_err = "bad argument"; return CB200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] {
        if (n_devices & (n_devices - 1)) throw InvalidArg("the number of target devices must be a power of two");
        *out = new cb200_backend{new Backend(std::vector<int>(devices, devices + n_devices))};
    });
}
int cb200_backend_destroy(cb200_backend *b)
{
    if (!b) return CB200_OK;
    delete b->impl;
    delete b;
    return CB200_OK;
}
int cb200_backend_set_channel(cb200_backend *b, cb200_send_fn send, cb200_recv_fn recv, void *user)
{
    NEED(b);
    b->impl->set_channel(send, recv, user);
    return CB200_OK;
}
int cb200_backend_execute(cb200_backend *b, const char *tasks_json, char **result_json_out)
{
    NEED(b);
    if (!tasks_json || !result_json_out) { g_err = "null argument"; return CB200_ERR_INVALID; }
    *result_json_out = nullptr;
    return guard([&] { *result_json_out = dup_string(b->impl->execute(tasks_json)); });
}
int cb200_backend_execute_many(cb200_backend *b, const char *const *messages, int n, char **results_json_out)
{
    NEED(b);
    if (!messages || !results_json_out || n < 0) { g_err = "bad argument"; return CB200_ERR_INVALID; }
    for (int i = 0; i < n; i++) results_json_out[i] = nullptr;
    return guard([&] {
        std::vector<std::string> in;
        for (int i = 0; i < n; i++) in.emplace_back(messages[i] ? messages[i] : "");
        const std::vector<std::string> out = b->impl->execute_many(in);
        for (int i = 0; i < n; i++) results_json_out[i] = dup_string(out[i]);
    });
}
int cb200_backend_batch_info(const cb200_backend *b, uint64_t info[3])
{
    NEED(b);
    info[0] = b->impl->last_uniform_flushes;
    info[1] = b->impl->last_split_flushes;
    info[2] = b->impl->last_solo;
    return CB200_OK;
}
int cb200_backend_stats(const cb200_backend *b, uint64_t stats[8])
{
    NEED(b);
    fill_stats(b->impl->last_stats, b->impl->last_stats.gates, b->impl->last_stats.relabelled_swaps, stats);
    return CB200_OK;
}

int cb200_backend_last_mode(const cb200_backend *b, int *mode_out)
{
    NEED(b);
    if (!mode_out) { g_err = "null output pointer"; return CB200_ERR_INVALID; }
    *mode_out = b->impl->last_mode;
    return CB200_OK;
}

}  // extern "C"
