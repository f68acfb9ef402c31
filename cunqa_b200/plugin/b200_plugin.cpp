#include "b200_plugin.hpp"

#include <cstdlib>
#include <map>
#include <stdexcept>

#include "utils/constants.hpp"

namespace cunqa {
namespace sim {

namespace b200 {

static std::string env_or_throw(const char* name)
{
    const char* v = std::getenv(name);
    if (!v) throw std::runtime_error(std::string("B200 plugin: environment variable ") + name + " is not set");
    return v;
}
std::string vqpu_endpoint() { return env_or_throw("SLURM_JOB_ID") + "_" + env_or_throw("SLURM_TASK_PID"); }
std::string executor_endpoint() { return env_or_throw("SLURM_JOB_ID") + "_executor"; }

// C-ABI trampolines for SEND / RECV of one measured bit (ClassicalChannel::send_measure / recv_measure)
static void send_bit(void* channel, int bit, const char* to)
{
    static_cast<comm::ClassicalChannel*>(channel)->send_measure(bit, std::string(to));
}
static int recv_bit(void* channel, const char* from)
{
    return static_cast<comm::ClassicalChannel*>(channel)->recv_measure(std::string(from));
}

}  // namespace b200

// ---- no communication ---------------------------------------------------------------------------
// Static and dynamic tasks take the same call: the engine decides between "evolve once and sample", deferred
// measurement and the per-shot interpreter, and answers in the flat {"counts","time_taken"} shape.
JSON B200SimpleSimulator::execute(const SimpleBackend& backend, const QuantumTask& task)
{
    const bool noisy = backend.config.contains("noise_model") && !backend.config.at("noise_model").empty();
    if (noisy) return {{"ERROR", "The B200 backend is an ideal state-vector simulator: noise models are not supported."}};
    return engine_.run(to_string(task), task.config);
}

// ---- classical communication --------------------------------------------------------------------
B200CCSimulator::B200CCSimulator() : channel_{b200::vqpu_endpoint()} { channel_.publish(); }

JSON B200CCSimulator::execute(const CCBackend&, const QuantumTask& task)
{
    for (const std::string& peer : task.sending_to) channel_.connect(peer);
    cb200_backend_set_channel(engine_.get(task.config), b200::send_bit, b200::recv_bit, &channel_);
    return engine_.run(to_string(task), task.config);
}

// ---- quantum communication ----------------------------------------------------------------------
// Handshake with the executor: publish our endpoint, wait until the executor says it is up, connect back.
B200QCSimulator::B200QCSimulator() : executor_{b200::executor_endpoint()}, channel_{b200::vqpu_endpoint()}
{
    channel_.publish();
    (void)channel_.recv_info(executor_);
    channel_.connect(executor_);
}

JSON B200QCSimulator::execute(const QCBackend&, const QuantumTask& task)
{
    const std::string message = to_string(task);  // empty = this vQPU sits the round out
    channel_.send_info(message, executor_);
    return message.empty() ? JSON() : JSON::parse(channel_.recv_info(executor_));
}

// ---- executor -----------------------------------------------------------------------------------
// The family's vQPUs are the entries of communications.json that carry this job's id; wait until all are there.
B200Executor::B200Executor(std::size_t n_qpus) : channel_{b200::executor_endpoint()}
{
    const std::string prefix = b200::env_or_throw("SLURM_JOB_ID") + "_";
    while (vqpus_.size() != n_qpus) {
        vqpus_.clear();
        const JSON published = read_file(constants::COMM_FILEPATH);
        for (const auto& entry : published.items())
            if (entry.key().compare(0, prefix.size(), prefix) == 0 && entry.key() != b200::executor_endpoint()) vqpus_.push_back(entry.key());
    }
    channel_.publish();
    for (const std::string& vqpu : vqpus_) {
        channel_.connect(vqpu);
        channel_.send_info("ready", vqpu);
    }
}

// One round = one message from every vQPU (empty when idle).  The non-empty ones become ONE engine message: the
// engine lays the tasks out in a joint register with the communication pairs and runs teledata / telegate on it.
void B200Executor::run()
{
    for (;;) {
        JSON batch = JSON::array();
        std::map<std::string, std::string> task_id_of;  // vQPU endpoint -> (unique) task id in this round
        for (std::size_t slot = 0; slot < vqpus_.size(); slot++) {
            const std::string text = channel_.recv_info(vqpus_[slot]);
            if (text.empty()) continue;
            JSON task = JSON::parse(text);
            std::string id = task.at("id").get<std::string>();
            for (const auto& taken : task_id_of)
                if (taken.second == id) { id += "_" + std::to_string(slot); break; }
            task["id"] = id;
            task_id_of[vqpus_[slot]] = id;
            batch.push_back(std::move(task));
        }
        if (batch.empty()) continue;

        const JSON result = engine_.run(batch.dump(), batch.front().at("config"));
        for (const auto& [vqpu, id] : task_id_of) {
            const JSON reply = result.contains("ERROR")
                                   ? result
                                   : JSON{{"counts", result.at("id_counts").at(id)}, {"time_taken", result.at("time_taken")}};
            channel_.send_info(reply.dump(), vqpu);
        }
    }
}

}  // namespace sim
}  // namespace cunqa
