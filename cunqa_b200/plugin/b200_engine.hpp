// b200_engine.hpp -- RAII owner of the C-ABI engine handle, shared by the B200 plugin classes.
//
// One instance per vQPU process, owned by the strategy object, which the backend keeps alive for the
// process lifetime (/root/reference/src/backends/simple_backend.hpp:72-74,96-97): this is where the CUDA
// context, the device state buffer and the stream live between jobs.  Exactly one thread ever calls
// execute() (/root/reference/src/qpu.cpp:21-31,33-63), so no locking is needed.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "cunqa_b200.h"
#include "quantum_task.hpp"
#include "utils/json.hpp"

namespace cunqa {
namespace sim {

class B200Engine {
public:
    B200Engine() = default;
    B200Engine(const B200Engine&) = delete;
    B200Engine& operator=(const B200Engine&) = delete;
    ~B200Engine() { if (handle_) cb200_backend_destroy(handle_); }

    // Lazily binds to the GPU named by the job itself: config.device = {"device_name":"GPU",
    // "target_devices":[k]} is computed at server start from CUDA_VISIBLE_DEVICES / SLURM_PROCID
    // (/root/reference/src/utils/helpers/net_functions.hpp:198-222) and echoed into every job.
    cb200_backend* get(const JSON& task_config)
    {
        if (!handle_) {
            // the whole list is handed over: with 2^g GPUs, registers of >= `b200_shard_min_qubits` qubits are sharded over
            // all of them inside the library (cb200_backend_create_multi), smaller ones run on the first
            std::vector<int> devices;
            if (task_config.contains("device") && task_config.at("device").contains("target_devices")) {
                const auto& devs = task_config.at("device").at("target_devices");
                if (devs.is_array())
                    for (const auto& d : devs) devices.push_back(d.get<int>());
            }
            if (devices.empty()) devices.push_back(0);
            const bool pow2 = (devices.size() & (devices.size() - 1)) == 0;
            const int rc = devices.size() > 1 && pow2 ? cb200_backend_create_multi(&handle_, devices.data(), (int)devices.size())
                                                      : cb200_backend_create(&handle_, devices.front());
            if (rc != CB200_OK) throw std::runtime_error(std::string("B200 backend: ") + cb200_last_error());
        }
        return handle_;
    }

    // Runs one wire-format message (a task, or a JSON array of tasks for the QC executor) and returns
    // the engine's result JSON; engine-side failures arrive as {"ERROR": "..."} like every other
    // CUNQA adapter (aer_simulator_adapter.cpp:836-840).
    JSON run(const std::string& message, const JSON& task_config)
    {
        char* out = nullptr;
        if (cb200_backend_execute(get(task_config), message.c_str(), &out) != CB200_OK || out == nullptr)
            return {{"ERROR", std::string(cb200_last_error())}};
        JSON result = JSON::parse(out);
        cb200_free(out);
        return result;
    }

private:
    cb200_backend* handle_ = nullptr;
};

} // End of sim namespace
} // End of cunqa namespace
