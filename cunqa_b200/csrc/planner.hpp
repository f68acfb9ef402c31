// planner.hpp -- host side of the hot path (H1 of SURVEY 8a): gate fusion and tile-pass scheduling.
//
// Plays the role of Aer's fusion pass + op buffer on the reference's CPU path (the work behind
// controller_execute / flush_ops at aer_simulator_adapter.cpp:829,564), but plans for the B200
// memory system: ops are grouped into *tile passes* (one HBM round trip each, see pass_types.hpp)
// instead of one 2^n sweep per fused gate.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "gates.hpp"
#include "pass_types.hpp"

namespace cb200 {

// one factor of a diagonal op: multiply the amplitude by z where qubit a is 1 (and qubit b is 1 when b >= 0); a = b = -1: a scalar
struct PhaseTerm { int a = -1, b = -1; cplx z = cplx(1, 0); };

struct HostOp {
    enum Kind { DENSE, DIAG, PERMX, SWAP } kind = DENSE;
    std::vector<int> q;     // DENSE/PERMX/SWAP: target qubits; DIAG: table qubits (bit m of the table index <-> q[m])
    std::vector<int> ctrl;  // control qubits (must all be 1); always empty for DIAG (folded into the table)
    std::vector<cplx> m;    // DENSE: 2^k x 2^k row-major; DIAG: 2^k entries
    int mask_row = -1;      // per-state flag row (batched feed-forward), -1 = unconditional
    int n_orig = 1;         // original circuit gates folded into this op
    // DIAG ops built from 1- and 2-qubit phase gates only (z s t rz p cp cz crz rzz ...): the table is the product of these
    // terms.  encode_pass uses them to pull the part of a phase ladder that commutes with the rest of its pass out of the
    // per-amplitude tables (the "fan" at the end of the pass, pass_types.hpp: OP_FAN).
    std::vector<PhaseTerm> terms;
    bool has_terms = false;

    uint64_t qmask() const;   // all qubits touched
    uint64_t xmask() const;   // qubits acted on non-diagonally
};

// encode_pass flags
constexpr int kEncPairOps = 1;         // register-block adjacent dense gates
constexpr int kEncDiagFuseAlways = 2;  // always fold a diagonal run into the dense gate that follows it (experiments)
constexpr int kEncNoStage = 4;
constexpr int kEncLinPerm = 64;        // runs of unconditioned X / CX / SWAP gates as one linear permutation of the tile (OP_LINPERM)
constexpr int kEncPlainMats = 32;      // never the three-multiplication matrix form (per-register matrices, A/B experiments)
constexpr int kEncFan = 16;            // pull commuting phase-ladder terms into a fan at the end of the pass (TMA kernel, no per-tile conditions)         // no per-tile shared-memory staging of diagonal sub-tables (experiments)

struct PlanOptions {
    int tile_bits = 12;        // t: 2^t amplitudes staged per tile
    int min_run_bits = 4;      // L_min: low physical qubits forced into every tile (coalescing)
    int max_diag_qubits = 10;  // width of a merged diagonal table
    int max_dense_qubits = 5;  // widest dense block the kernel accepts
    int max_pass_cost = 0;     // cap on sum over dense ops of 2^k per pass (0 = unlimited)
    int lookahead = 256;       // pending ops examined when choosing a pass's tile qubits (0 = first-come greedy);
                               // states below 2^28 amplitudes use at most 64 (a pass is too short to pay for more)
    int lookahead_min_qubits = 24;  // smaller states: a pass costs less than the look-ahead that would save it
    bool fuse = true;
    bool pair_ops = true;      // register-block adjacent dense gates (two gates per shared-memory round trip)
    // measured on QFT-30 (profiles/qft30_diag_ab_r02.txt): folding a diagonal run into the dense gate after it always pays
    // (81.6 ms against 84.9 when it is skipped where it makes more tables depend on register bits), and staging the
    // per-amplitude sub-tables in shared memory does not (82.6 against 81.6): the pass is bound by its instruction count
    // (258 per amplitude), not by the table loads.  Both stay switchable: CB200_DIAG_FUSE_ALWAYS, CB200_STAGE_DIAG.
    bool diag_fuse_always = true, stage_diag = false;
    bool fan = true;           // CB200_FAN: phase fans (OP_FAN) for controlled-phase ladders
    bool lin_perm = true;      // CB200_LIN_PERM: runs of X / CX / SWAP gates as one permutation of the tile (OP_LINPERM)
    int enc_flags() const { return (pair_ops ? kEncPairOps : 0) | (diag_fuse_always ? kEncDiagFuseAlways : 0) | (stage_diag ? 0 : kEncNoStage) | (fan ? kEncFan : 0) | (lin_perm ? kEncLinPerm : 0); }
    uint64_t global_mask = 0;  // sharded state: physical qubits that select the rank (never fused into dense blocks)
};

struct Pass {
    std::vector<int> tile_q;       // ascending physical qubits spanning the tile
    int L = 0;                     // leading run bits
    std::vector<int> op_idx;       // indices into the fused op list, program order
};

// Fuser: feeds ops one at a time and merges them into the pending list.
class Fuser {
public:
    explicit Fuser(const PlanOptions &o) : opt(o) {}
    void push(HostOp op);
    std::vector<HostOp> ops;
    PlanOptions opt;

private:
    bool try_merge(HostOp &prev, const HostOp &nw);
};

// expand controls into a dense matrix over (targets..., ctrls...) ; index bit order = [q..., ctrl...]
HostOp expand_controls(const HostOp &op);
// dense product on the union qubit set `uq` (bit m <-> uq[m]):  result = B_embedded * A_embedded (A first)
std::vector<cplx> embed_dense(const HostOp &op, const std::vector<int> &uq);
std::vector<cplx> matmul(const std::vector<cplx> &b, const std::vector<cplx> &a, int dim);

// group the fused ops into tile passes for an n-qubit register
std::vector<Pass> schedule_passes(const std::vector<HostOp> &ops, int n, const PlanOptions &opt);

// build the kernel parameter block of one pass.  diag tables are appended to `diag_host`
// (element units) and referenced by offset.  Returns false + err if the pass cannot be encoded.
template <typename T>
bool encode_pass(const Pass &pass, const std::vector<HostOp> &ops, int n, int batch, PassParams<T> &out,
                 std::vector<cplx> &diag_host, std::string &err, int flags = kEncPairOps);

}  // namespace cb200
