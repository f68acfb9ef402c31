#include "b200_qc_simulator.hpp"

#include <cstdlib>

using namespace std::string_literals;

namespace cunqa {
namespace sim {

// Same rendezvous as AER/aer_qc_simulator.cpp:13-32: publish, wait for the executor's "ready", connect.
B200QCSimulator::B200QCSimulator() :
    executor_id{std::getenv("SLURM_JOB_ID") + "_executor"s},
    classical_channel{std::getenv("SLURM_JOB_ID") + "_"s + std::getenv("SLURM_TASK_PID")}
{
    classical_channel.publish();
    auto ready = classical_channel.recv_info(executor_id);
    classical_channel.connect(executor_id);
}

JSON B200QCSimulator::execute([[maybe_unused]] const QCBackend& backend, const QuantumTask& quantum_task)
{
    auto circuit = to_string(quantum_task);
    classical_channel.send_info(circuit, executor_id);  // "" = idle this round
    if (circuit != "") {
        auto results = classical_channel.recv_info(executor_id);
        return JSON::parse(results);
    }
    return JSON();
}

} // End namespace sim
} // End namespace cunqa
