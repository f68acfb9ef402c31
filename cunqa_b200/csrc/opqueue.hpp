// opqueue.hpp -- host-only op queue: gate lookup, logical->physical qubit map, fusion.
// (the analogue of AerState's buffered ops between flush_ops calls, aer_simulator_adapter.cpp:564,574)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "planner.hpp"

namespace cb200 {

struct InvalidArg : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

class OpQueue {
public:
    OpQueue(int n_qubits, int batch, const PlanOptions &opt);
    void reset();  // drop pending ops and the qubit relabelling
    void clear_pending();
    // logical qubits; flags = per-state condition (batched feed-forward), nullptr = unconditional
    void apply_matrix(const int *qubits, int k, const int *ctrls, int nc, const cplx *mat, const uint8_t *flags = nullptr);
    void apply_diagonal(const int *qubits, int k, const cplx *diag, const uint8_t *flags = nullptr);
    void apply_named(const std::string &name, const int *qubits, int nq, const double *params, int np,
                     const uint8_t *flags = nullptr);
    void global_phase(double theta, const uint8_t *flags = nullptr);

    std::vector<HostOp> &ops() { return fuser_.ops; }
    size_t pending() const { return fuser_.ops.size(); }
    uint64_t phys_index(uint64_t logical) const;
    uint64_t logical_index(uint64_t phys) const;
    bool identity_perm() const;
    // per-task fusion switches (Aer's `fusion_enable` / `fusion_max_qubit` config keys, aer_helpers.hpp:17-86)
    void set_fusion(bool enable, int max_dense_qubits);

    int n, batch;
    PlanOptions opt;
    std::vector<int> l2p;  // logical -> physical qubit; uncontrolled SWAPs only edit this table
    std::vector<std::vector<uint8_t>> mask_rows;
    uint64_t gates = 0, relabelled_swaps = 0;
    // state-preparation peephole: while the register is still |0...0> (nothing flushed since reset), X gates on
    // qubits nothing else has touched only change WHICH basis state is initialised -- no pass at all
    bool fresh = true;
    uint64_t basis = 0;     // physical index of the initial basis state
    uint64_t touched = 0;   // physical qubits some queued op already acts on

private:
    void push_op(HostOp op, const uint8_t *flags);
    void check_qubits(const int *q, int k, const int *c, int nc) const;
    Fuser fuser_;
};

}  // namespace cb200
