#include "b200_simple_simulator.hpp"

namespace cunqa {
namespace sim {

// Static and dynamic tasks both go through one engine call: the engine itself decides between
// "evolve once + sample" and the per-shot interpreter (cf. AER/aer_simple_simulator.cpp:8-22), and
// already returns the flat {"counts","time_taken"} shape that cunqa/result.py:113 accepts.
JSON B200SimpleSimulator::execute([[maybe_unused]] const SimpleBackend& backend, const QuantumTask& quantum_task)
{
    if (backend.config.contains("noise_model") && !backend.config.at("noise_model").empty())
        return {{"ERROR", "The B200 backend is an ideal state-vector simulator: noise models are not supported."}};
    return engine_.run(to_string(quantum_task), quantum_task.config);
}

} // End namespace sim
} // End namespace cunqa
