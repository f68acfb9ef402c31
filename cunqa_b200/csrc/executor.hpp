// executor.hpp -- whole-task execution behind cb200_backend_execute.
//
// Static tasks:  AerSimulatorAdapter::simulate(const Backend*)           (aer_simulator_adapter.cpp:806-842)
// Dynamic tasks: AerSimulatorAdapter::simulate(ClassicalChannel*, bool)  (:845-935) + execute_shot_ (:120-792)
// with shots batched on the device (B independent states advance per launch) instead of the
// reference's one-state-per-shot loop / OpenMP team (:877-925).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "engine.hpp"
#include "json_min.hpp"

namespace cb200 {

typedef void (*send_fn)(void *user, int bit, const char *target);
typedef int (*recv_fn)(void *user, const char *origin);

class Backend {
public:
    explicit Backend(int device) : device_(device), devices_{device} {}
    // several GPUs of one box (config.device.target_devices, net_functions.hpp:198-222): registers of at least
    // `b200_shard_min_qubits` qubits (default 30) are sharded over all of them, smaller ones run on the first
    explicit Backend(const std::vector<int> &devices) : device_(devices.empty() ? 0 : devices[0]), devices_(devices) {}
    // returns the result JSON text (never throws: failures become {"ERROR": "..."} )
    std::string execute(const std::string &tasks_json);
    // n independent messages (each what execute() accepts: a task, a JSON array of tasks = one quantum-communication
    // group, or a {"params":..} update of the task last sent in the same slot) run as SEPARATE registers batched on the
    // device: one State with batch = n, one launch per tile pass over all registers when their circuits share their
    // structure (BASELINE config 5: many small vQPU groups per GPU).  Result i has the shape execute() gives message i.
    std::vector<std::string> execute_many(const std::vector<std::string> &messages);
    uint64_t last_uniform_flushes = 0, last_split_flushes = 0, last_solo = 0;
    void set_channel(send_fn s, recv_fn r, void *user) { send_ = s; recv_ = r; user_ = user; }
    Stats last_stats;
    int last_mode = 0;  // 0 static, 1 dynamic with deferred measurements, 2 dynamic on the batched interpreter
    void *stream_ptr() { return state_ ? (void *)state_->stream() : nullptr; }

private:
    Json run(const Json &msg);
    Json run_static(const Json &task);
    struct RegJob;  // one register of a batched run (executor.cpp)
    Json run_dynamic(const std::vector<Json> &tasks, RegJob *job = nullptr);
    State &get_state(int n, int precision, int batch, bool shard = false);
    bool shards(int n, int shard_min_qubits) const;
    std::vector<int> devices_;
    std::unique_ptr<State> batch_state_;      // execute_many: the registers
    std::vector<Json> cached_many_;           // execute_many: last full message per slot

    int device_;
    std::unique_ptr<State> state_;
    Json cached_;          // last full task message (for {"params":...} updates)
    bool has_cached_ = false;
    std::shared_ptr<void> parsed_;   // the cached task, parsed (static tasks only): {"params"} updates rewrite its angles in place
    Json run_static_parsed(const void *parsed_task);
    send_fn send_ = nullptr;
    recv_fn recv_ = nullptr;
    void *user_ = nullptr;
};

// apply a {"params":..,"shots":..} message to a task (host only; cb200_update_task_json)
std::string update_task_json(const std::string &task_json, const std::string &update_json);

// plan a task without a device: fused ops + tile passes as JSON (cb200_plan_json)
std::string plan_task_json(const std::string &task_json, const std::string &options_json);
// OpenQASM 2 text -> CUNQA circuit JSON (qasm2.cpp; replaces src/utils/helpers/qasm2_to_json.hpp:520-549)
std::string qasm2_to_json(const std::string &qasm);

}  // namespace cb200
