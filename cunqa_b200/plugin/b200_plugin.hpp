// b200_plugin.hpp -- everything CUNQA needs to raise vQPUs on the B200 backend.
//
//   B200SimpleSimulator / B200CCSimulator / B200QCSimulator : the three SimulatorStrategy<> flavours that
//       turn_ON_QPU<Simulator, Config, Backend> instantiates (src/cli/setup_qpus.cpp:69-82,127-231);
//   B200Executor : the per-family executor process of the quantum-communication scheme
//       (src/cli/setup_executor.cpp:35-87).
// All of them are thin: the work happens behind the C ABI (cunqa_b200.h) through one B200Engine per process.
#pragma once

#include <cstddef>
#include <string>
#include <vector>

#include "backends/cc_backend.hpp"
#include "backends/qc_backend.hpp"
#include "backends/simple_backend.hpp"
#include "backends/simulators/simulator_strategy.hpp"
#include "classical_channel/classical_channel.hpp"
#include "quantum_task.hpp"
#include "utils/json.hpp"

#include "b200_engine.hpp"

namespace cunqa {
namespace sim {

namespace b200 {
// endpoint names of this SLURM job's processes on the classical channel: "<job>_<task pid>" for a vQPU,
// "<job>_executor" for the family's executor (the convention of every CUNQA simulator plugin)
std::string vqpu_endpoint();
std::string executor_endpoint();
}  // namespace b200

// what the three flavours share: the registered name (a key of SIMULATORS_MAP, because
// get_basis_gates(simulator->get_name()) is called on it, setup_qpus.cpp:77)
template <class BackendT>
class B200Strategy : public SimulatorStrategy<BackendT> {
public:
    std::string get_name() const override { return "B200"; }
};

// no communication: one engine call per job
class B200SimpleSimulator final : public B200Strategy<SimpleBackend> {
public:
    JSON execute(const SimpleBackend& backend, const QuantumTask& task) override;

private:
    B200Engine engine_;
};

// classical communication: measured bits travel between vQPU processes over this process's channel
class B200CCSimulator final : public B200Strategy<CCBackend> {
public:
    B200CCSimulator();
    JSON execute(const CCBackend& backend, const QuantumTask& task) override;

private:
    comm::ClassicalChannel channel_;
    B200Engine engine_;
};

// quantum communication: the vQPU forwards its task to the executor, which simulates the whole family
class B200QCSimulator final : public B200Strategy<QCBackend> {
public:
    B200QCSimulator();
    JSON execute(const QCBackend& backend, const QuantumTask& task) override;

private:
    const std::string executor_;
    comm::ClassicalChannel channel_;
};

class B200Executor {
public:
    explicit B200Executor(std::size_t n_qpus);
    [[noreturn]] void run();

private:
    comm::ClassicalChannel channel_;
    std::vector<std::string> vqpus_;
    B200Engine engine_;
};

}  // namespace sim
}  // namespace cunqa
