#include "b200_cc_simulator.hpp"

#include <cstdlib>

using namespace std::string_literals;

namespace {
// C-ABI trampolines: SEND/RECV of one measured bit between vQPU processes
// (ClassicalChannel::send_measure / recv_measure, src/classical_channel/classical_channel.hpp:25-26).
void send_bit(void* user, int bit, const char* target)
{
    static_cast<cunqa::comm::ClassicalChannel*>(user)->send_measure(bit, std::string(target));
}
int recv_bit(void* user, const char* origin)
{
    return static_cast<cunqa::comm::ClassicalChannel*>(user)->recv_measure(std::string(origin));
}
} // End of anonymous namespace

namespace cunqa {
namespace sim {

// The channel id convention is the one every CUNQA CC simulator uses (AER/aer_cc_simulator.cpp:10-14).
B200CCSimulator::B200CCSimulator() :
    classical_channel{std::getenv("SLURM_JOB_ID") + "_"s + std::getenv("SLURM_TASK_PID")}
{
    classical_channel.publish();
}

JSON B200CCSimulator::execute([[maybe_unused]] const CCBackend& backend, const QuantumTask& quantum_task)
{
    for (const auto& qpu_id : quantum_task.sending_to)
        classical_channel.connect(qpu_id);

    cb200_backend_set_channel(engine_.get(quantum_task.config), send_bit, recv_bit, &classical_channel);
    return engine_.run(to_string(quantum_task), quantum_task.config);
}

} // End namespace sim
} // End namespace cunqa
