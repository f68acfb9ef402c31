// gates.hpp -- named-gate table of the engine.
//
// Vocabulary and argument order are the reference's: names from INSTRUCTIONS_MAP
// (/root/reference/src/utils/constants.hpp.in:271-465), restricted to the union of AER_BASIS_GATES
// (:468-485), CUNQASIM_BASIS_GATES (:561-567) and the cases of the Aer dynamic interpreter
// (aer_simulator_adapter.cpp:183-752); controls first, targets last (apply_mcx/mcu/mcswap call
// sites :245-434).  Definitions are the standard Qiskit matrices for those names.
#pragma once
#include <complex>
#include <string>
#include <vector>

namespace cb200 {

using cplx = std::complex<double>;

struct GateSpec {
    int n_targets = 0;       // the LAST n_targets qubits of the instruction are targets
    std::vector<cplx> mat;   // 2^n_targets x 2^n_targets, row-major, index bit m <-> target m
    bool is_global_phase = false;
    double phase = 0.0;
};

// returns false (and sets err) for an unknown name / wrong arity
bool lookup_gate(const std::string &name, int n_qubits, const double *params, int n_params,
                 GateSpec &out, std::string &err);

// number of angle parameters update_params_ rewrites for this gate name
// (/root/reference/src/quantum_task.cpp:51-97); 0 = not parametric there.
int update_param_arity(const std::string &name);

bool is_diagonal(const std::vector<cplx> &mat, int dim);

}  // namespace cb200
