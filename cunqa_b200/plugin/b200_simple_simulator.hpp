// B200 plugin, no-communication flavour: the sibling of AER/aer_simple_simulator.{hpp,cpp}.
#pragma once

#include "quantum_task.hpp"
#include "backends/simple_backend.hpp"
#include "backends/simulators/simulator_strategy.hpp"
#include "utils/json.hpp"

#include "b200_engine.hpp"

namespace cunqa {
namespace sim {

class B200SimpleSimulator final : public SimulatorStrategy<SimpleBackend> {
public:
    B200SimpleSimulator() = default;
    ~B200SimpleSimulator() = default;

    inline std::string get_name() const override { return "B200"; }
    JSON execute(const SimpleBackend& backend, const QuantumTask& circuit) override;

private:
    B200Engine engine_;
};

} // End of sim namespace
} // End of cunqa namespace
