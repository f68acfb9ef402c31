// Executor process of a quantum-communication family on the B200 backend: the sibling of
// AER/aer_executor.{hpp,cpp}.  Collects one task per vQPU per round and hands them to the engine as ONE
// message; the engine lays them out in a joint register (+ communication pairs) and runs the teledata /
// telegate interpreter with shots batched on the GPU.
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "classical_channel/classical_channel.hpp"

#include "b200_engine.hpp"

namespace cunqa {
namespace sim {

class B200Executor {
public:
    B200Executor(const std::size_t& n_qpus);
    ~B200Executor() = default;

    void run();

private:
    comm::ClassicalChannel classical_channel;
    std::vector<std::string> qpu_ids;
    B200Engine engine_;
};

} // End of sim namespace
} // End of cunqa namespace
