// ref_parse.cpp -- driver around the REFERENCE'S OWN wire-format code.  TEST INFRASTRUCTURE ONLY.
//
// Compiled (by oracle/ref_boundary/build.sh, into oracle/_ref/) together with the reference sources
// where they lie: /root/reference/src/quantum_task.cpp and /root/reference/src/utils/json.cpp, against
// /root/reference/src/utils/constants.hpp.in (instantiated).  It lets the tests check the engine's
// JSON front end against what the reference itself parses:
//   stdin : one JSON message per line (a full task, then optional {"params":[...],"shots":N} updates)
//   stdout: per message one JSON line {"id","is_dynamic","sending_to","config","structured":[...]} where
//           `structured` is from_quantum_task_to_structuredqtask() flattened (name/qubits/clbits/params/
//           matrix/diagonal/qpus/condition/operation/instructions), i.e. after update_params_().
#include <iostream>
#include <memory>
#include <string>

#include <spdlog/sinks/null_sink.h>
#include <spdlog/spdlog.h>

#include "quantum_task.hpp"
#include "utils/json.hpp"

std::shared_ptr<spdlog::logger> logger = std::make_shared<spdlog::logger>("ref", std::make_shared<spdlog::sinks::null_sink_mt>());

using cunqa::JSON;

static JSON dump_instruction(const cunqa::constants::CUNQAInstruction& in)
{
    JSON j = {{"name", in.name}, {"qubits", in.qubits}, {"clbits", in.clbits}, {"params", in.params},
              {"qpus", in.qpus}, {"condition", in.condition}, {"operation", in.operation},
              {"l_clbits", in.l_clbits}, {"r_clbits", in.r_clbits}};
    if (!in.matrix.empty()) j["matrix"] = in.matrix;
    if (!in.diagonal.empty()) j["diagonal"] = in.diagonal;
    JSON sub = JSON::array();
    for (const auto& s : in.instructions) sub.push_back(dump_instruction(s));
    j["instructions"] = sub;
    return j;
}

int main()
{
    cunqa::QuantumTask task;
    std::string line;
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        try {
            task.update_circuit(line);
            auto st = cunqa::from_quantum_task_to_structuredqtask(task);
            JSON out = {{"id", task.id}, {"is_dynamic", task.is_dynamic}, {"sending_to", task.sending_to},
                        {"config", task.config}, {"n_qubits", st.n_qubits}, {"n_clbits", st.n_clbits}};
            JSON arr = JSON::array();
            for (const auto& in : st.instructions) arr.push_back(dump_instruction(in));
            out["structured"] = arr;
            out["to_string"] = cunqa::to_string(task);
            std::cout << out.dump() << std::endl;
        } catch (const std::exception& e) {
            std::cout << JSON({{"ERROR", e.what()}}).dump() << std::endl;
        }
    }
    return 0;
}
