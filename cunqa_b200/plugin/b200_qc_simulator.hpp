// B200 plugin with quantum communications: the sibling of AER/aer_qc_simulator.{hpp,cpp}.  As in every
// CUNQA QC simulator the vQPU only forwards its task to the family's executor process, which simulates
// all co-scheduled tasks in one joint register (b200_executor.hpp).
#pragma once

#include "quantum_task.hpp"
#include "backends/qc_backend.hpp"
#include "backends/simulators/simulator_strategy.hpp"
#include "classical_channel/classical_channel.hpp"
#include "utils/json.hpp"

namespace cunqa {
namespace sim {

class B200QCSimulator final : public SimulatorStrategy<QCBackend> {
public:
    B200QCSimulator();
    ~B200QCSimulator() = default;

    inline std::string get_name() const override { return "B200"; }
    JSON execute([[maybe_unused]] const QCBackend& backend, const QuantumTask& circuit) override;

private:
    std::string executor_id;
    comm::ClassicalChannel classical_channel;
};

} // End namespace sim
} // End namespace cunqa
