/*
 * cunqa_b200.h -- C ABI of libcunqa_b200.so, the B200-native state-vector engine for CUNQA.
 *
 * The shape (opaque handle, plain pointers and sizes, int status, caller owns host buffers,
 * library owns device memory) follows the in-tree precedent of a C-ABI simulator behind a CUNQA
 * plugin: Maestro's handle API used at
 *   /root/reference/src/backends/simulators/Maestro/maestro_adapters/maestro_simulator_adapter.cpp:774,792,880,885
 * Each entry point names the reference interface it replaces.  All paths are relative to
 * /root/reference/src/backends/simulators/AER/aer_adapters/ unless stated otherwise.
 *
 * Conventions: qubit 0 = least-significant bit of the amplitude index (Qiskit/Aer little
 * endian); dense matrices are row-major, (re,im) interleaved doubles, matrix-index bit m <->
 * qubits[m]; counts bitstrings have clbit 0 as the rightmost character (aer_helpers.hpp:175-179).
 * Status: 0 = ok, <0 = error; the message is available from cb200_last_error().
 * There is NO CPU fallback: every compute entry point fails with CB200_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef CUNQA_B200_H
#define CUNQA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_OK 0
#define CB200_ERR_INVALID (-1) /* bad argument / malformed task */
#define CB200_ERR_CUDA (-2)    /* CUDA runtime failure or no device */
#define CB200_ERR_UNSUPPORTED (-3)
#define CB200_ERR_NCCL (-4)

#define CB200_PRECISION_DOUBLE 64 /* complex128 (reference default; dynamic path forces it, aer_simulator_adapter.cpp:946) */
#define CB200_PRECISION_SINGLE 32 /* complex64 (Aer "precision":"single", aer_helpers.hpp:20) */

typedef struct cb200_state cb200_state;     /* one register of n qubits x batch independent states */
typedef struct cb200_backend cb200_backend; /* per-vQPU engine: state cache + plan cache          */

/* ---- library -------------------------------------------------------------------------------- */
const char *cb200_version(void);
/* last error of this thread (entry points that fail before a handle exists) or of the handle */
const char *cb200_last_error(void);
int cb200_device_count(void);
void cb200_free(void *p); /* frees strings returned by the *_json entry points (Maestro FreeResult analogue) */

/* ---- state lifecycle: replaces AerState::{configure,allocate_qubits,initialize,clear}
 *      (aer_simulator_adapter.cpp:887-892, 905-911, 937-952) ---------------------------------- */
int cb200_state_create(cb200_state **out, int n_qubits, int precision, int device, int batch);
int cb200_state_destroy(cb200_state *s);
int cb200_state_reset(cb200_state *s); /* every state of the batch back to |0...0> */
int cb200_state_num_qubits(const cb200_state *s);
int cb200_state_batch(const cb200_state *s);

/* ---- gate application: ops are queued, fused and scheduled into tile passes at the next flush
 *      (the analogue of AerState's op buffer + flush_ops, aer_simulator_adapter.cpp:564,574) --- */
/* replaces apply_unitary / apply_mcu / apply_mcx ... (aer_simulator_adapter.cpp:203-470):
 * dense 2^k x 2^k matrix on `qubits`, applied on the subspace where all `ctrls` are 1. */
int cb200_apply_matrix(cb200_state *s, const int *qubits, int k, const int *ctrls, int nc,
                       const double *mat_rowmajor_reim);
/* replaces apply_diagonal_matrix / apply_mcphase / apply_mcrz (:337,:394,:531) */
int cb200_apply_diagonal(cb200_state *s, const int *qubits, int k, const double *diag_reim);
/* named gate with the reference's vocabulary and argument order (constants.hpp.in:271-465;
 * controls first, targets last): "h","cx","mcrz","cswap","ecr",...  */
int cb200_apply_gate(cb200_state *s, const char *name, const int *qubits, int nq,
                     const double *params, int np);
/* classically conditioned variant for batched shots: state b receives the gate iff
 * flags[b] != 0 (device feed-forward for CIF, aer_simulator_adapter.cpp:582-600). */
int cb200_apply_gate_masked(cb200_state *s, const char *name, const int *qubits, int nq,
                            const double *params, int np, const uint8_t *flags_host);
int cb200_flush(cb200_state *s); /* plan + launch everything queued (asynchronous) */
int cb200_sync(cb200_state *s);  /* flush + wait for the device                    */

/* ---- measurement: replaces apply_measure / apply_reset (:166-167,:187,:218) and the static
 *      path's sample_measure (:829) ------------------------------------------------------------ */
int cb200_norm(cb200_state *s, double *out_per_state);
/* P(qubit == 1) of every state of the batch */
int cb200_prob1(cb200_state *s, int qubit, double *out_per_state);
/* marginal distribution over `qubits` (k <= 20) of state `b`: out[2^k], index bit m <-> qubits[m] */
int cb200_probabilities(cb200_state *s, int b, const int *qubits, int k, double *out);
/* measure one qubit of every state: draw r_b = U(seed, counters[b]) (splitmix64 stream shared
 * with the oracle), outcome = r_b*norm >= P0, collapse + renormalise.  forced[b] in {0,1}
 * overrides the draw (deterministic test mode); pass NULL for random. */
int cb200_measure(cb200_state *s, int qubit, uint64_t seed, const uint64_t *counters,
                  const int8_t *forced, int *bits_out);
int cb200_reset_qubit(cb200_state *s, int qubit, uint64_t seed, const uint64_t *counters,
                      const int8_t *forced);
/* sample `shots` basis states of state `b` by inverse CDF, u_s = U(seed, s)*norm */
int cb200_sample(cb200_state *s, int b, uint64_t shots, uint64_t seed, uint64_t *out_indices);
/* read amplitudes at the given basis indices of state `b` -> out[2*n] (re,im); fp32 states are widened */
int cb200_get_amplitudes(cb200_state *s, int b, const uint64_t *idx, uint64_t n, double *out_reim);
/* upload a full state vector (2^n (re,im) doubles) into state `b` (tests / save_state round trips) */
int cb200_set_amplitudes(cb200_state *s, int b, const double *in_reim);

/* the CUDA stream (cudaStream_t) every kernel of this state is launched on, for timing with CUDA events */
void *cb200_state_stream(cb200_state *s);

/* ---- planner introspection (host only, no device needed) ------------------------------------- */
/* counters of everything flushed so far on this state:
 * stats[0]=tile passes launched, [1]=gates queued (original), [2]=fused ops executed,
 * [3]=algorithmic bytes moved by passes (2*S each), [4]=kernel launches (all kinds),
 * [5]=swaps done by relabelling, [6]=host->device bytes (kernel parameters, tables, flags, uploads),
 * [7]=device->host bytes (outcomes, samples, amplitudes) */
int cb200_stats(const cb200_state *s, uint64_t stats[8]);
/* plan a wire-format task without executing it; returns a JSON description of the fused ops and
 * tile passes (used by the CPU tests to prove the plan is equivalent to the circuit). */
int cb200_plan_json(const char *task_json, const char *options_json, char **plan_json_out);
/* apply a `{"params":[...],"shots":N}` update message to a wire-format task exactly as the backend does
 * between two executes (replaces QuantumTask::update_params_, /root/reference/src/quantum_task.cpp:39-108);
 * returns the updated task as JSON text.  Host only. */
int cb200_update_task_json(const char *task_json, const char *update_json, char **task_json_out);
/* OpenQASM 2 text -> CUNQA circuit JSON {"instructions","num_qubits","num_clbits","quantum_registers",
 * "classical_registers"} (replaces qasm2_to_json, /root/reference/src/utils/helpers/qasm2_to_json.hpp:520-549).
 * Host only. */
int cb200_qasm2_to_json(const char *qasm, char **circuit_json_out);
/* matrix of a named gate exactly as the engine builds it: out = 2^nq x 2^nq row-major (re,im),
 * controls expanded.  nq <= 6. */
int cb200_gate_matrix(const char *name, int nq, const double *params, int np, double *out_reim);

/* ---- whole-task execution: what the plugin's execute() calls.  Replaces
 *      AerSimulatorAdapter::simulate(const Backend*) (:806-842) for static tasks and
 *      AerSimulatorAdapter::simulate(ClassicalChannel*, bool) + execute_shot_ (:120-792,:845-935)
 *      for dynamic ones. ----------------------------------------------------------------------- */
int cb200_backend_create(cb200_backend **out, int device);
int cb200_backend_destroy(cb200_backend *b);
/* classical-channel hooks for SEND/RECV between vQPU processes; the plugin wires them to
 * ClassicalChannel::send_measure / recv_measure (src/classical_channel/classical_channel.hpp:25-26).
 * The engine drains its CUDA stream before every blocking recv (cf. flush_ops at :564,:574). */
typedef void (*cb200_send_fn)(void *user, int bit, const char *target_qpu);
typedef int (*cb200_recv_fn)(void *user, const char *origin_qpu);
int cb200_backend_set_channel(cb200_backend *b, cb200_send_fn send, cb200_recv_fn recv, void *user);
/* tasks_json: one wire-format quantum task (src/quantum_task.cpp:18-36) or a JSON array of them
 * (quantum-communication executor: one joint register, aer_simulator_adapter.cpp:131-159).
 * A `{"params":[...],"shots":N}` message re-runs the cached task with new angles
 * (src/quantum_task.cpp:33-35,39-108).  Result: `{"counts":{...},"time_taken":s}` for one task,
 * `{"id_counts":{id:{...}},"time_taken":s}` for several; `{"ERROR":"..."}` on failure
 * (status stays CB200_OK so that the plugin can forward the message, cf. :836-840). */
int cb200_backend_execute(cb200_backend *b, const char *tasks_json, char **result_json_out);
/* MANY independent messages in one call, run as SEPARATE registers batched on the device (BASELINE config 5: "many
 * small independent virtual QPUs are batched as separate state vectors per GPU").  messages[i] is anything
 * cb200_backend_execute accepts -- one task, a JSON array of tasks (= one quantum-communication group with its own joint
 * register, aer_simulator_adapter.cpp:131-159), or a {"params":[...],"shots":N} update of the message last sent in slot i
 * (src/quantum_task.cpp:33-35).  In the reference each vQPU / each executor group is its own process calling simulate()
 * (src/cli/setup_qpus.cpp:69-82, AER/aer_executor.cpp:41-85); on one GPU they share this engine instead: registers of
 * equal width live in ONE device buffer and advance in the same tile passes -- one launch per pass over all of them with
 * per-register matrices when their circuits have the same structure (a VQE ansatz evaluated at many parameter sets),
 * register by register otherwise; every measurement is deferred and all shots of all registers are drawn in one
 * sampling pass.  A message that needs a real mid-circuit collapse or an external classical channel runs on its own
 * through the ordinary path.  results_json_out[i] (free with cb200_free) has the shape cb200_backend_execute returns
 * for messages[i].  The same call is reachable as the message {"b200_batch":[...]} -> {"b200_batch_results":[...]}. */
int cb200_backend_execute_many(cb200_backend *b, const char *const *messages, int n, char **results_json_out);
/* how the last execute_many ran: info[0] = flushes that ran all registers in the same launches, [1] = flushes that ran
 * register by register (structures differed), [2] = messages that ran alone */
int cb200_backend_batch_info(const cb200_backend *b, uint64_t info[3]);
/* counters of the backend's last execute: same layout as cb200_stats */
int cb200_backend_stats(const cb200_backend *b, uint64_t stats[8]);
/* how the last execute ran its measurements: 0 = static task (one state, shots sampled), 1 = dynamic task with every
 * measurement deferred (one state, feed-forward as quantum control, shots sampled at the end), 2 = dynamic task on
 * the batched per-shot interpreter (execute_shot_, aer_simulator_adapter.cpp:120-792) */
int cb200_backend_last_mode(const cb200_backend *b, int *mode_out);
/* stream of the backend's current state (NULL before the first execute) */
void *cb200_backend_stream(cb200_backend *b);

/* ---- sharded state across the GPUs of one box, ONE process (what the plugin uses) -------------------
 * The reference hands a backend its devices as config.device = {"device_name":"GPU","target_devices":[...]}
 * (/root/reference/src/utils/helpers/net_functions.hpp:198-222, forwarded at AER/aer_helpers.hpp:111-116) and runs one
 * process per vQPU (src/cli/setup_qpus.cpp:69-82).  cb200_backend_create_multi is the backend for such a list: registers
 * of at least `b200_shard_min_qubits` (config key, default 30) qubits are sharded over all listed GPUs -- top
 * log2(n_devices) qubits global, exchanges over NVLink peer memory (cudaDeviceEnablePeerAccess), one worker thread per
 * GPU inside the library, no NCCL / IPC / rendezvous -- and everything cb200_backend_execute accepts works on them
 * (static and dynamic tasks, {"params"} updates, amplitude / marginal return); smaller registers run on devices[0].
 * cb200_state_create_sharded gives the same sharded register to the state-level calls (every entry point of the
 * "state lifecycle", "gate application" and "measurement" sections, plus cb200_dist_stats). */
int cb200_backend_create_multi(cb200_backend **out, const int *devices, int n_devices);
int cb200_state_create_sharded(cb200_state **out, int n_qubits, int precision, const int *devices, int n_devices);

/* ---- sharded state, one PROCESS per GPU (no reference analogue) -------------
 * One process per GPU.  The top log2(world) qubits are global (index bits select the rank).
 * Rendezvous is done by the caller (torch.distributed / MPI): rank 0 obtains an id blob with
 * cb200_dist_unique_id, broadcasts it, every rank calls cb200_state_create_dist.
 * Every rank then issues the SAME sequence of calls on its handle (gates, flush, norm, prob1, measure,
 * reset_qubit, sample, get_amplitudes): reductions are all-reduced, so each call returns the same values on every
 * rank (cb200_sample: every rank resolves the shots that fall into its interval of the global cumulative
 * distribution, the shot indices are then all-reduced). */
#define CB200_DIST_ID_BYTES 128
int cb200_dist_unique_id(uint8_t id_out[CB200_DIST_ID_BYTES]);
int cb200_state_create_dist(cb200_state **out, int n_qubits_total, int precision, int device,
                            int rank, int world, const uint8_t id[CB200_DIST_ID_BYTES]);
/* peer-memory exchange: every rank exports an IPC handle of its shard, the caller all-gathers
 * them, and the engine maps the peers so that the global<->local qubit swap runs as ONE kernel
 * reading and writing peer HBM over NVLink (no staging buffer). */
#define CB200_IPC_HANDLE_BYTES 64
int cb200_dist_export_handle(cb200_state *s, uint8_t handle_out[CB200_IPC_HANDLE_BYTES]);
int cb200_dist_import_handles(cb200_state *s, const uint8_t *handles_all_ranks);
/* out[0]=rank, [1]=world, [2]=global<->local qubit swaps performed, [3]=bytes this rank sent over NVLink,
 * [4]=device time spent in swaps, microseconds (drains the stream) */
int cb200_dist_stats(cb200_state *s, uint64_t out[5]);

#ifdef __cplusplus
}
#endif
#endif /* CUNQA_B200_H */
