// opqueue.cpp -- see opqueue.hpp
#include "opqueue.hpp"

#include <algorithm>
#include <cmath>

namespace cb200 {

OpQueue::OpQueue(int n_qubits, int nbatch, const PlanOptions &o) : n(n_qubits), batch(nbatch), opt(o), fuser_(o)
{
    l2p.resize(n);
    reset();
}

void OpQueue::reset()
{
    clear_pending();
    for (int i = 0; i < n; i++) l2p[i] = i;
    fresh = true;
    basis = 0;
    touched = 0;
}

void OpQueue::clear_pending()
{
    fuser_.ops.clear();
    mask_rows.clear();
}

void OpQueue::set_fusion(bool enable, int max_dense_qubits)
{
    opt.fuse = fuser_.opt.fuse = enable;
    opt.max_dense_qubits = fuser_.opt.max_dense_qubits = std::max(1, std::min(max_dense_qubits, kMaxDenseK));
}

bool OpQueue::identity_perm() const
{
    for (int i = 0; i < n; i++) if (l2p[i] != i) return false;
    return true;
}
uint64_t OpQueue::phys_index(uint64_t logical) const
{
    uint64_t p = 0;
    for (int i = 0; i < n; i++) p |= ((logical >> i) & 1ull) << l2p[i];
    return p;
}
uint64_t OpQueue::logical_index(uint64_t phys) const
{
    uint64_t l = 0;
    for (int i = 0; i < n; i++) l |= ((phys >> l2p[i]) & 1ull) << i;
    return l;
}

// a diagonal over one or two qubits as a product of phase terms (exact for any such table)
static void terms_of_small_diag(HostOp &d)
{
    const size_t k = d.q.size();
    if (k == 0 || k > 2) return;
    for (const cplx &z : d.m) if (std::abs(z) == 0.0) return;
    d.terms.clear();
    // (factors that are exactly 1 -- the |0>, |01>, |10> entries of a controlled phase -- are not terms)
    auto add = [&](int a, int b, cplx z) {
        if (z == cplx(1, 0)) return;
        PhaseTerm t;
        t.a = a; t.b = b; t.z = z;
        d.terms.push_back(t);
    };
    add(-1, -1, d.m[0]);                                    // scalar: the |0..0> entry
    add(d.q[0], -1, d.m[1] / d.m[0]);
    if (k == 2) {
        add(d.q[1], -1, d.m[2] / d.m[0]);
        add(d.q[0], d.q[1], d.m[3] * d.m[0] / (d.m[1] * d.m[2]));
    }
    d.has_terms = true;
}

void OpQueue::push_op(HostOp op, const uint8_t *flags)
{
    if (op.kind == HostOp::DIAG && !flags) terms_of_small_diag(op);
    if (flags) {
        op.mask_row = (int)mask_rows.size();
        mask_rows.emplace_back(flags, flags + batch);
    }
    gates += 1;
    touched |= op.qmask();
    fuser_.push(std::move(op));
}

void OpQueue::check_qubits(const int *q, int k, const int *c, int nc) const
{
    uint64_t seen = 0;
    auto chk = [&](int x) {
        if (x < 0 || x >= n) throw InvalidArg("qubit index " + std::to_string(x) + " out of range");
        if ((seen >> x) & 1ull) throw InvalidArg("duplicate qubit " + std::to_string(x) + " in one instruction");
        seen |= 1ull << x;
    };
    for (int i = 0; i < k; i++) chk(q[i]);
    for (int i = 0; i < nc; i++) chk(c[i]);
}

void OpQueue::apply_matrix(const int *qubits, int k, const int *ctrls, int nc, const cplx *mat, const uint8_t *flags)
{
    if (k < 1 || k > kMaxDenseK)
        throw Unsupported("dense blocks of 1..5 qubits are supported (got " + std::to_string(k) + ")");
    check_qubits(qubits, k, ctrls, nc);
    const int D = 1 << k;
    std::vector<cplx> m(mat, mat + (size_t)D * D);
    HostOp op;
    for (int i = 0; i < k; i++) op.q.push_back(l2p[qubits[i]]);
    for (int i = 0; i < nc; i++) op.ctrl.push_back(l2p[ctrls[i]]);
    if (is_diagonal(m, D) && k + nc <= opt.max_diag_qubits) {
        // diagonal (controls folded into the table): needs no tile slot at all
        HostOp d;
        d.kind = HostOp::DIAG;
        d.q = op.q;
        d.q.insert(d.q.end(), op.ctrl.begin(), op.ctrl.end());
        const int kk = k + nc;
        d.m.assign((size_t)1 << kk, cplx(1, 0));
        for (int idx = 0; idx < (1 << kk); idx++) {
            bool on = true;
            for (int c = 0; c < nc; c++) if (!((idx >> (k + c)) & 1)) on = false;
            if (on) d.m[idx] = m[(size_t)(idx & (D - 1)) * D + (idx & (D - 1))];
        }
        push_op(std::move(d), flags);
        return;
    }
    if (k == 1 && m[0] == cplx(0, 0) && m[3] == cplx(0, 0) && m[1] == cplx(1, 0) && m[2] == cplx(1, 0)) {
        if (fresh && nc == 0 && flags == nullptr && !((touched >> op.q[0]) & 1ull)) {
            basis ^= 1ull << op.q[0];  // X on an untouched qubit of |0...0>: a different initial basis state
            gates += 1;
            return;
        }
        op.kind = HostOp::PERMX;
        push_op(std::move(op), flags);
        return;
    }
    op.kind = HostOp::DENSE;
    op.m = std::move(m);
    push_op(std::move(op), flags);
}

void OpQueue::apply_diagonal(const int *qubits, int k, const cplx *diag, const uint8_t *flags)
{
    if (k < 1 || k > kMaxDiagK) throw Unsupported("diagonal of 1..12 qubits supported");
    check_qubits(qubits, k, nullptr, 0);
    HostOp d;
    d.kind = HostOp::DIAG;
    for (int i = 0; i < k; i++) d.q.push_back(l2p[qubits[i]]);
    d.m.assign(diag, diag + ((size_t)1 << k));
    push_op(std::move(d), flags);
}

void OpQueue::global_phase(double theta, const uint8_t *flags)
{
    const int q0 = 0;
    const cplx ph = std::exp(cplx(0, theta));
    const cplx d[2] = {ph, ph};
    apply_diagonal(&q0, 1, d, flags);
}

void OpQueue::apply_named(const std::string &name, const int *qubits, int nq, const double *params, int np,
                          const uint8_t *flags)
{
    if (name == "barrier" || name == "id" || name == "save_state") return;
    GateSpec g;
    std::string err;
    if (!lookup_gate(name, nq, params, np, g, err)) throw Unsupported(err);
    if (g.is_global_phase) { global_phase(g.phase, flags); return; }
    const int nt = g.n_targets, nc = nq - nt;
    const int *ctrls = qubits, *targets = qubits + nc;
    const bool is_swap = nt == 2 && g.mat[0] == cplx(1, 0) && g.mat[15] == cplx(1, 0) && g.mat[6] == cplx(1, 0) &&
                         g.mat[9] == cplx(1, 0) && g.mat[5] == cplx(0, 0) && g.mat[10] == cplx(0, 0);
    if (is_swap) {
        check_qubits(targets, 2, ctrls, nc);
        if (nc == 0 && flags == nullptr) {
            // uncontrolled, unconditional SWAP = pure relabelling of the qubit map: zero passes
            std::swap(l2p[targets[0]], l2p[targets[1]]);
            gates += 1;
            relabelled_swaps += 1;
            return;
        }
        HostOp op;
        op.kind = HostOp::SWAP;
        op.q = {l2p[targets[0]], l2p[targets[1]]};
        for (int i = 0; i < nc; i++) op.ctrl.push_back(l2p[ctrls[i]]);
        push_op(std::move(op), flags);
        return;
    }
    apply_matrix(targets, nt, ctrls, nc, g.mat.data(), flags);
}

}  // namespace cb200
