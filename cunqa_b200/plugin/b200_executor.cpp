#include "b200_executor.hpp"

#include <cstdlib>

#include "quantum_task.hpp"
#include "utils/constants.hpp"
#include "utils/json.hpp"

using namespace std::string_literals;

namespace cunqa {
namespace sim {

// Discovery of the family's vQPUs through communications.json, as in AER/aer_executor.cpp:20-39.
B200Executor::B200Executor(const std::size_t& n_qpus) :
    classical_channel{std::getenv("SLURM_JOB_ID") + "_executor"s}
{
    const std::string job = std::getenv("SLURM_JOB_ID");
    JSON ids;
    do {
        ids = JSON::object();
        JSON all = read_file(constants::COMM_FILEPATH);
        for (const auto& [key, value] : all.items())
            if (key.substr(0, key.find('_')) == job) ids[key] = value;
    } while (ids.size() != n_qpus);

    classical_channel.publish();
    for (const auto& [key, _] : ids.items()) {
        qpu_ids.push_back(key);
        classical_channel.connect(key);
        classical_channel.send_info("ready", key);
    }
}

void B200Executor::run()
{
    while (true) {
        // one polling round: an empty string means "this vQPU is idle" (AER/aer_executor.cpp:49-66)
        std::vector<std::string> working;
        std::unordered_map<std::string, std::string> task_of_qpu;
        JSON batch = JSON::array();
        JSON first_config;
        int position = 0;
        for (const auto& qpu_id : qpu_ids) {
            const std::string message = classical_channel.recv_info(qpu_id);
            if (!message.empty()) {
                JSON task = JSON::parse(message);
                std::string id = task.at("id").get<std::string>();
                for (const auto& [other, used] : task_of_qpu)
                    if (used == id) { id += "_" + std::to_string(position); break; }  // de-duplicate ids
                task["id"] = id;
                if (batch.empty()) first_config = task.at("config");
                task_of_qpu[qpu_id] = id;
                working.push_back(qpu_id);
                batch.push_back(std::move(task));
            }
            position++;
        }
        if (batch.empty()) continue;

        JSON result = engine_.run(batch.dump(), first_config);
        for (const auto& qpu : working) {
            JSON reply;
            if (result.contains("ERROR")) reply = result;
            else reply = {{"counts", result.at("id_counts").at(task_of_qpu[qpu])}, {"time_taken", result.at("time_taken")}};
            classical_channel.send_info(reply.dump(), qpu);
        }
    }
}

} // End of sim namespace
} // End of cunqa namespace
