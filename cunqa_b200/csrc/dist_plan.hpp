// dist_plan.hpp -- host-only planning for a state sharded over 2^g ranks (no reference analogue).
//
// Physical qubits [0, n_local) are local (index bits inside the shard), [n_local, n_total) are global (bit
// p - n_local of the rank).  For one rank, a list of ops on physical qubits is turned into segments:
//   * LOCAL: ops that only need this rank's shard.  Diagonal tables are sliced by the rank's global bits,
//     controls on global qubits are resolved (rank bit 0: op dropped, 1: control removed);
//   * SWAP(gbit, lbit): exchange global qubit n_local+gbit with local qubit lbit -- every rank swaps half of
//     its shard with rank ^ (1 << gbit).  Emitted only when a non-diagonal target is global.
// All ranks derive the same swap sequence AND the same batch ids (both depend on the op list only, never on which
// ops a rank dropped), so the exchange steps line up even when the LOCAL segments between them differ per rank.
#pragma once
#include <vector>

#include "planner.hpp"

namespace cb200 {

struct DistSegment {
    bool is_swap = false;
    int gbit = 0, lbit = 0;       // SWAP
    int batch = -1;               // SWAP: exchanges planned together carry the same id (identical on every rank); only
                                  // swaps of one batch may be executed as ONE multi-qubit exchange
    std::vector<HostOp> ops;      // LOCAL (qubits < n_local)
};

struct DistPlan {
    std::vector<DistSegment> segments;
    std::vector<int> sigma;       // where each original physical qubit ended up
};

DistPlan plan_distributed(const std::vector<HostOp> &ops, int n_total, int n_local, int rank);

}  // namespace cb200
