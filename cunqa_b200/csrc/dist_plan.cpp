// dist_plan.cpp -- see dist_plan.hpp
#include "dist_plan.hpp"

#include <algorithm>
#include <stdexcept>

namespace cb200 {

namespace {

// next position in `ops` (from index `from`) at which physical qubit q (in the ORIGINAL numbering) is used
// non-diagonally; ops.size() if never
size_t next_x_use(const std::vector<HostOp> &ops, size_t from, int q_orig)
{
    for (size_t i = from; i < ops.size(); i++)
        if ((ops[i].xmask() >> q_orig) & 1ull) return i;
    return ops.size();
}

}  // namespace

DistPlan plan_distributed(const std::vector<HostOp> &ops, int n_total, int n_local, int rank)
{
    DistPlan plan;
    const int g = n_total - n_local;
    std::vector<int> sigma(n_total), inv(n_total);  // sigma[orig] = current position; inv[pos] = orig
    for (int i = 0; i < n_total; i++) sigma[i] = inv[i] = i;
    DistSegment cur;
    auto flush_cur = [&]() {
        if (!cur.ops.empty()) plan.segments.push_back(std::move(cur));
        cur = DistSegment();
    };
    auto rank_bit = [&](int pos) { return (rank >> (pos - n_local)) & 1; };
    int swap_batch = 0;

    for (size_t i = 0; i < ops.size(); i++) {
        HostOp op = ops[i];
        for (int &x : op.q) x = sigma[x];
        for (int &x : op.ctrl) x = sigma[x];
        for (PhaseTerm &t : op.terms) { if (t.a >= 0) t.a = sigma[t.a]; if (t.b >= 0) t.b = sigma[t.b]; }
        // non-diagonal targets on global positions: bring them in.  When one swap is needed, every global qubit
        // that some op within the look-ahead window also needs non-diagonally is brought in at the same time, so
        // that exchanges cluster (back-to-back swaps) and the local segments between them stay long.
        if (op.kind != HostOp::DIAG) {
            bool need_swap = false;
            for (int x : op.q) if (x >= n_local) need_swap = true;
            if (need_swap) {
                flush_cur();
                const size_t window = 4096;
                std::vector<std::pair<size_t, int>> wanted;  // (first non-diagonal use, global position)
                for (int gp = n_local; gp < n_total; gp++) {
                    const size_t use = next_x_use(ops, i, inv[gp]);
                    if (use < std::min(ops.size(), i + window)) wanted.emplace_back(use, gp);
                }
                std::sort(wanted.begin(), wanted.end());
                uint64_t taken = 0;  // local positions already used as victims in this batch
                for (auto &w : wanted) {
                    const int gpos = w.second;
                    // victim: a high local position (large contiguous halves) whose next non-diagonal use lies
                    // farthest in the future, not touched by the op that triggered the exchange
                    int best = -1;
                    size_t best_use = 0;
                    const uint64_t mine = op.qmask();
                    for (int lp = n_local - 1; lp >= std::max(0, n_local - 8); lp--) {
                        if (((mine | taken) >> lp) & 1ull) continue;
                        const size_t use = next_x_use(ops, i, inv[lp]);
                        if (best < 0 || use > best_use) { best = lp; best_use = use; }
                    }
                    if (best < 0) throw std::runtime_error("distributed planner: no local qubit available for a global swap");
                    if (best_use <= w.first && w.first != i) continue;  // the victim is needed even sooner: not worth it
                    taken |= 1ull << best;
                    DistSegment sw;
                    sw.is_swap = true;
                    sw.gbit = gpos - n_local;
                    sw.lbit = best;
                    sw.batch = swap_batch;
                    plan.segments.push_back(sw);
                    const int og = inv[gpos], ol = inv[best];
                    sigma[og] = best; sigma[ol] = gpos;
                    inv[best] = og; inv[gpos] = ol;
                    for (int &x : op.q) if (x == gpos) x = best; else if (x == best) x = gpos;
                    for (int &x : op.ctrl) if (x == gpos) x = best; else if (x == best) x = gpos;
                    // (only non-diagonal ops trigger an exchange: no phase terms to remap here)
                }
                swap_batch++;
                for (int x : op.q)
                    if (x >= n_local) throw std::runtime_error("distributed planner: internal error (target still global)");
            }
        }
        // controls on global positions
        bool dropped = false;
        for (size_t c = 0; c < op.ctrl.size();) {
            if (op.ctrl[c] >= n_local) {
                if (!rank_bit(op.ctrl[c])) { dropped = true; break; }
                op.ctrl.erase(op.ctrl.begin() + c);
            } else {
                c++;
            }
        }
        if (dropped) continue;
        // diagonal table sliced by this rank's global bits
        if (op.kind == HostOp::DIAG) {
            std::vector<int> lq;
            std::vector<int> lbits;
            int fixed = 0;
            for (size_t m = 0; m < op.q.size(); m++) {
                if (op.q[m] >= n_local) { if (rank_bit(op.q[m])) fixed |= 1 << m; }
                else { lq.push_back(op.q[m]); lbits.push_back((int)m); }
            }
            if (lq.size() != op.q.size()) {
                HostOp d;
                d.kind = HostOp::DIAG;
                d.mask_row = op.mask_row;
                d.n_orig = op.n_orig;
                // the phase terms sliced the same way: a global qubit is this rank's bit
                d.has_terms = op.has_terms;
                for (PhaseTerm t : op.terms) {
                    bool gone = false;
                    for (int *e : {&t.a, &t.b}) {
                        if (*e >= n_local) {
                            if (!rank_bit(*e)) gone = true;   // factor 1 on this rank
                            *e = -1;
                        }
                    }
                    if (gone) continue;
                    if (t.a < 0 && t.b >= 0) std::swap(t.a, t.b);
                    d.terms.push_back(t);
                }
                if (lq.empty()) {  // a rank-dependent scalar phase
                    d.q = {0};
                    d.m = {op.m[fixed], op.m[fixed]};
                } else {
                    d.q = lq;
                    d.m.resize((size_t)1 << lq.size());
                    for (int idx = 0; idx < (1 << lq.size()); idx++) {
                        int full = fixed;
                        for (size_t b = 0; b < lbits.size(); b++) full |= ((idx >> b) & 1) << lbits[b];
                        d.m[idx] = op.m[full];
                    }
                }
                op = std::move(d);
            }
        }
        cur.ops.push_back(std::move(op));
    }
    flush_cur();
    plan.sigma = sigma;
    (void)g;
    return plan;
}

}  // namespace cb200
