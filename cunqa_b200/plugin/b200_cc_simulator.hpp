// B200 plugin with classical communications: the sibling of AER/aer_cc_simulator.{hpp,cpp}.
#pragma once

#include "quantum_task.hpp"
#include "backends/cc_backend.hpp"
#include "backends/simulators/simulator_strategy.hpp"
#include "classical_channel/classical_channel.hpp"
#include "utils/json.hpp"

#include "b200_engine.hpp"

namespace cunqa {
namespace sim {

class B200CCSimulator final : public SimulatorStrategy<CCBackend> {
public:
    B200CCSimulator();
    ~B200CCSimulator() = default;

    inline std::string get_name() const override { return "B200"; }
    JSON execute(const CCBackend& backend, const QuantumTask& circuit) override;

private:
    comm::ClassicalChannel classical_channel;
    B200Engine engine_;
};

} // End namespace sim
} // End namespace cunqa
