// plugin_run.cpp -- links and RUNS the B200 plugin classes against the reference's own boundary code.  TEST INFRASTRUCTURE.
//
// Linked from: cunqa_b200/plugin/b200_plugin.cpp (the product's plugin), /root/reference/src/quantum_task.cpp and
// /root/reference/src/utils/json.cpp where they lie (the reference's task parser, parameter updater and file registry),
// fake_channel.cpp (in-process ClassicalChannel) and libcunqa_b200.so.  It drives the plugin exactly as the reference's
// processes do:
//   simple : turn_ON_QPU<B200SimpleSimulator, SimpleConfig, SimpleBackend> (src/cli/setup_qpus.cpp:69-82) followed by the
//            compute loop of QPU::compute_result_ (src/qpu.cpp:33-63): quantum_task.update_circuit(message);
//            backend->execute(quantum_task) -- for a full task and then for {"params":..} updates;
//   cc     : two vQPU "processes" (threads) with B200CCSimulator exchanging measured bits (aer_cc_simulator.cpp:17-33);
//   qc     : two vQPUs with B200QCSimulator + one B200Executor (setup_executor.cpp:35-87, aer_executor.cpp:41-85).
// stdin: one JSON document (see each mode); stdout: one JSON line per result.
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <spdlog/sinks/null_sink.h>
#include <spdlog/spdlog.h>

#include "b200_plugin.hpp"
#include "quantum_task.hpp"
#include "utils/json.hpp"

std::shared_ptr<spdlog::logger> logger = std::make_shared<spdlog::logger>("ref", std::make_shared<spdlog::sinks::null_sink_mt>());

using cunqa::JSON;
using cunqa::QuantumTask;
using namespace cunqa::sim;

static std::mutex out_mutex;
static void emit(const JSON& j)
{
    std::lock_guard<std::mutex> lk(out_mutex);
    std::cout << j.dump() << std::endl;
}

// the server loop body of a vQPU (src/qpu.cpp:33-63): parse/patch the cached task, execute, report errors as {"ERROR"}
template <class BackendT>
static JSON serve(const BackendT& backend, QuantumTask& task, const std::string& message)
{
    try {
        task.update_circuit(message);
        return backend.execute(task);
    } catch (const std::exception& e) {
        return {{"ERROR", std::string(e.what())}};
    }
}

static int run_simple(const JSON& doc)
{
    auto simulator = std::make_unique<B200SimpleSimulator>();
    if (simulator->get_name() != "B200") return 2;
    SimpleConfig config;
    SimpleBackend backend(config, std::move(simulator));
    QuantumTask task;
    for (const JSON& m : doc.at("messages")) emit(serve(backend, task, m.dump()));
    // a noise model is refused, not ignored
    SimpleConfig noisy;
    noisy.noise_model = {{"errors", JSON::array({1})}};
    SimpleBackend nb(noisy, std::make_unique<B200SimpleSimulator>());
    QuantumTask t2;
    emit(serve(nb, t2, doc.at("messages").at(0).dump()));
    return 0;
}

static int run_cc(const JSON& doc)
{
    const std::vector<std::string> pids = {"101", "102"};
    std::vector<std::unique_ptr<CCBackend>> backends;
    for (const std::string& pid : pids) {
        setenv("SLURM_TASK_PID", pid.c_str(), 1);   // the endpoint name is read in the simulator's constructor
        CCConfig config;
        backends.push_back(std::make_unique<CCBackend>(config, std::make_unique<B200CCSimulator>()));
    }
    std::vector<JSON> results(2);
    std::vector<std::thread> th;
    for (int i = 0; i < 2; i++)
        th.emplace_back([&, i] {
            QuantumTask task;
            results[i] = serve(*backends[i], task, doc.at("tasks").at(i).dump());
        });
    for (auto& t : th) t.join();
    for (const JSON& r : results) emit(r);
    return 0;
}

static int run_qc(const JSON& doc)
{
    const std::vector<std::string> pids = {"201", "202"};
    const int rounds = doc.value("rounds", 1);
    std::vector<std::vector<JSON>> results(2);
    // executor "process": constructed concurrently with the vQPUs (it waits for their endpoints, they wait for "ready")
    std::thread exec([] {
        B200Executor executor(2);
        executor.run();  // never returns; the process exits once the vQPUs have their answers
    });
    exec.detach();
    std::vector<std::thread> th;
    const std::string job = std::getenv("SLURM_JOB_ID");
    for (int i = 0; i < 2; i++)
        th.emplace_back([&, i] {
            // the endpoint name comes from the environment (one process per vQPU in the reference): vQPU 1 waits until
            // vQPU 0 has published its endpoint, i.e. has read the variable, before setting its own
            if (i == 1)
                for (;;) {
                    const JSON published = cunqa::read_file(cunqa::constants::COMM_FILEPATH);
                    if (published.contains(job + "_" + pids[0])) break;
                    std::this_thread::yield();
                }
            setenv("SLURM_TASK_PID", pids[i].c_str(), 1);
            QCConfig config;
            // (the constructor publishes the endpoint and blocks until the executor says "ready")
            auto backend = std::make_unique<QCBackend>(config, std::make_unique<B200QCSimulator>());
            QuantumTask task;
            for (int r = 0; r < rounds; r++) results[i].push_back(serve(*backend, task, doc.at("tasks").at(i).dump()));
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < 2; i++)
        for (const JSON& r : results[i]) emit(JSON{{"vqpu", i}, {"result", r}});
    std::cout.flush();
    std::_Exit(0);
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::cerr << "usage: plugin_run simple|cc|qc < scenario.json" << std::endl; return 64; }
    const std::string mode = argv[1];
    const std::string text((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
    try {
        const JSON doc = JSON::parse(text);
        if (mode == "simple") return run_simple(doc);
        if (mode == "cc") return run_cc(doc);
        if (mode == "qc") return run_qc(doc);
        std::cerr << "unknown mode " << mode << std::endl;
        return 64;
    } catch (const std::exception& e) {
        emit({{"HARNESS_ERROR", std::string(e.what())}});
        return 1;
    }
}
