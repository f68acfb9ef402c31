// planner.cpp -- gate fusion + tile-pass scheduling (host only; see planner.hpp).
#include "planner.hpp"

#include <algorithm>
#include <cstring>

namespace cb200 {

uint64_t HostOp::qmask() const
{
    uint64_t m = 0;
    for (int x : q) m |= 1ull << x;
    for (int x : ctrl) m |= 1ull << x;
    return m;
}
uint64_t HostOp::xmask() const
{
    if (kind == DIAG) return 0;
    uint64_t m = 0;
    for (int x : q) m |= 1ull << x;
    return m;
}

static int popcount64(uint64_t v) { return __builtin_popcountll(v); }

static std::vector<int> mask_to_list(uint64_t m)
{
    std::vector<int> out;
    for (int i = 0; i < 64; i++) if ((m >> i) & 1ull) out.push_back(i);
    return out;
}

// full dense matrix of `op` over the qubit list uq (index bit m <-> uq[m]); uq must cover op's qubits
std::vector<cplx> embed_dense(const HostOp &op, const std::vector<int> &uq)
{
    const int K = (int)uq.size();
    const int D = 1 << K;
    auto pos = [&](int qubit) {
        for (int i = 0; i < K; i++) if (uq[i] == qubit) return i;
        return -1;
    };
    std::vector<int> tp, cp;
    for (int x : op.q) tp.push_back(pos(x));
    for (int x : op.ctrl) cp.push_back(pos(x));
    const int k = (int)op.q.size();
    std::vector<cplx> M((size_t)D * D, cplx(0, 0));
    for (int c = 0; c < D; c++) {
        bool on = true;
        for (int p : cp) if (!((c >> p) & 1)) on = false;
        if (!on) { M[(size_t)c * D + c] = 1; continue; }
        int j = 0;  // sub-index over op.q
        for (int m = 0; m < k; m++) j |= ((c >> tp[m]) & 1) << m;
        int rest = c;
        for (int m = 0; m < k; m++) rest &= ~(1 << tp[m]);
        auto row_of = [&](int jj) {
            int r = rest;
            for (int m = 0; m < k; m++) if ((jj >> m) & 1) r |= 1 << tp[m];
            return r;
        };
        switch (op.kind) {
        case HostOp::DENSE:
            for (int jj = 0; jj < (1 << k); jj++) M[(size_t)row_of(jj) * D + c] = op.m[(size_t)jj * (1 << k) + j];
            break;
        case HostOp::DIAG: M[(size_t)c * D + c] = op.m[j]; break;
        case HostOp::PERMX: M[(size_t)row_of(j ^ 1) * D + c] = 1; break;
        case HostOp::SWAP: {
            const int jj = ((j & 1) << 1) | ((j >> 1) & 1);
            M[(size_t)row_of(jj) * D + c] = 1;
            break;
        }
        }
    }
    return M;
}

std::vector<cplx> matmul(const std::vector<cplx> &b, const std::vector<cplx> &a, int dim)
{
    std::vector<cplx> out((size_t)dim * dim, cplx(0, 0));
    for (int r = 0; r < dim; r++)
        for (int k = 0; k < dim; k++) {
            const cplx v = b[(size_t)r * dim + k];
            if (v == cplx(0, 0)) continue;
            for (int c = 0; c < dim; c++) out[(size_t)r * dim + c] += v * a[(size_t)k * dim + c];
        }
    return out;
}

HostOp expand_controls(const HostOp &op)
{
    if (op.ctrl.empty() && op.kind == HostOp::DENSE) return op;
    HostOp out;
    out.kind = HostOp::DENSE;
    out.q = op.q;
    out.q.insert(out.q.end(), op.ctrl.begin(), op.ctrl.end());
    out.m = embed_dense(op, out.q);
    out.mask_row = op.mask_row;
    out.n_orig = op.n_orig;
    return out;
}

bool Fuser::try_merge(HostOp &prev, const HostOp &nw)
{
    const uint64_t P = prev.qmask(), N = nw.qmask(), U = P | N;
    const int nu = popcount64(U);
    if (prev.kind == HostOp::DIAG && nw.kind == HostOp::DIAG) {
        if (nu > opt.max_diag_qubits) return false;
        const std::vector<int> uq = mask_to_list(U);
        std::vector<cplx> tab((size_t)1 << nu);
        auto sub = [&](const HostOp &o, int idx) {
            int j = 0;
            for (size_t m = 0; m < o.q.size(); m++) {
                const int p = (int)(std::find(uq.begin(), uq.end(), o.q[m]) - uq.begin());
                j |= ((idx >> p) & 1) << m;
            }
            return o.m[j];
        };
        for (int idx = 0; idx < (1 << nu); idx++) tab[idx] = sub(prev, idx) * sub(nw, idx);
        prev.q = uq;
        prev.m = std::move(tab);
        prev.n_orig += nw.n_orig;
        prev.has_terms = prev.has_terms && nw.has_terms;
        if (prev.has_terms) prev.terms.insert(prev.terms.end(), nw.terms.begin(), nw.terms.end());
        else prev.terms.clear();
        return true;
    }
    // dense absorb: one op's qubit set contains the other's
    if (U != P && U != N) return false;
    if (U & opt.global_mask) return false;  // keep controls / diagonals on global qubits communication-free
    if (nu > opt.max_dense_qubits) return false;
    const HostOp &big = (U == P) ? prev : nw;
    const bool big_plain_dense = big.kind == HostOp::DENSE && big.ctrl.empty();
    if (nu > 2 && !big_plain_dense) return false;
    // keep the qubit order of the containing op so that a plain dense block keeps its matrix layout
    std::vector<int> uq = big.q;
    for (int c : big.ctrl) uq.push_back(c);
    const std::vector<cplx> A = embed_dense(prev, uq), B = embed_dense(nw, uq);
    HostOp merged;
    merged.kind = HostOp::DENSE;
    merged.q = uq;
    merged.m = matmul(B, A, 1 << nu);
    merged.n_orig = prev.n_orig + nw.n_orig;
    // a product that is diagonal goes back to the (tile-free) diagonal form
    if (is_diagonal(merged.m, 1 << nu)) {
        std::vector<cplx> d((size_t)1 << nu);
        for (int i = 0; i < (1 << nu); i++) d[i] = merged.m[(size_t)i * (1 << nu) + i];
        merged.kind = HostOp::DIAG;
        merged.m = std::move(d);
    }
    prev = std::move(merged);
    return true;
}

void Fuser::push(HostOp op)
{
    if (!opt.fuse || op.mask_row >= 0) { ops.push_back(std::move(op)); return; }
    // `self`: where the op lives once it has merged with an earlier one.  A merge that widened the earlier op (a 1-qubit
    // gate swallowed by the 2-qubit gate that follows it) goes on looking further back from there: the block may now
    // contain other earlier gates too (ry(a), ry(b), cx(a,b): both rotations end up in the block, not one of them)
    int self = -1;
    int steps = 0;
    for (int i = (int)ops.size() - 1; i >= 0 && steps < 64; --i, ++steps) {
        const HostOp &cur = self >= 0 ? ops[self] : op;
        const uint64_t qm = cur.qmask(), xm = cur.xmask();
        HostOp &prev = ops[i];
        const uint64_t pm = prev.qmask();
        if (!(pm & qm)) continue;      // disjoint: commutes
        if (prev.mask_row >= 0) break; // conditioned op: ordering barrier
        if (try_merge(prev, cur)) {
            if (self >= 0) ops.erase(ops.begin() + self);
            if (ops[i].qmask() == pm || ops[i].kind != HostOp::DENSE) return;   // the earlier op kept its width: nothing new in reach
            self = i;
            continue;
        }
        const bool commute = ((prev.xmask() & qm) == 0) && ((xm & pm) == 0);
        if (commute) continue;         // both act diagonally on the shared qubits
        break;
    }
    if (self < 0) ops.push_back(std::move(op));
}

// scalars of kernel-parameter storage a dense k-qubit matrix takes (pass_types.hpp: 3 per element up to two qubits)
static int dense_scalars(int k) { return (k <= 2 ? 3 : 2) * (1 << (2 * k)); }

// Dense matrices are written as plain (re, im) pairs while a pass is encoded; a pass that qualifies (encode_pass, end)
// then has its <= 2-qubit matrices rewritten in place into the three-multiplication form of tile_kernel.cuh: element
// e = c + i d becomes m[2e] = c, m[2e+1] = -(c+d), m[2E+e] = d-c (E = 4^k; the slot holds 3E scalars in either form).
template <typename T>
static void pack_dense(T *dst, const cplx *m, int k)
{
    const int E = 1 << (2 * k);
    for (int e = 0; e < E; e++) { dst[2 * e] = (T)m[e].real(); dst[2 * e + 1] = (T)m[e].imag(); }
}
template <typename T>
static void to_mul3(T *m, int k)
{
    const int E = 1 << (2 * k);
    for (int e = 0; e < E; e++) {
        const T c = m[2 * e], d = m[2 * e + 1];
        m[2 * e + 1] = -(c + d);
        m[2 * E + e] = d - c;
    }
}

static int dense_cost(const HostOp &op)
{
    switch (op.kind) {
    case HostOp::DENSE: return 1 << op.q.size();
    case HostOp::DIAG: return 1;
    default: return 1;
    }
}

namespace {

// One pass's admission loop: walk the pending ops in program order starting at `first`; an op joins the pass when
// none of its qubits is blocked by an earlier op that stayed behind (diagonal uses only block later non-diagonal
// uses) and its non-diagonal targets fit into the tile.  grow = true: targets may claim free tile slots (Q grows);
// grow = false: only ops whose targets already lie in Q.  `window` bounds how many pending ops are examined.
// Returns the admitted op indices (and the number of them through the return value's size).
std::vector<int> admit(const std::vector<HostOp> &ops, const std::vector<char> &done, size_t first, uint64_t &Q, bool grow,
                       int t, const PlanOptions &opt, uint64_t all, int window)
{
    std::vector<int> sel;
    uint64_t blockedX = 0, blockedZ = 0;
    int cost = 0, mat_scalars = 0, seen = 0;
    for (size_t i = first; i < ops.size(); i++) {
        if (done[i]) continue;
        if (window > 0 && seen++ >= window) break;
        const HostOp &op = ops[i];
        const uint64_t qx = op.xmask();
        const uint64_t qz = op.qmask() & ~qx;
        const bool can = !(qx & (blockedX | blockedZ)) && !(qz & blockedX);
        if (can) {
            const uint64_t newQ = Q | qx;
            const int c = dense_cost(op);
            const int ms = op.kind == HostOp::DENSE ? dense_scalars((int)op.q.size()) : 0;
            bool fits = (grow ? popcount64(newQ) <= t : newQ == Q) && (int)sel.size() < kMaxOps && mat_scalars + ms <= kMatScalars;
            if (fits && opt.max_pass_cost > 0 && !sel.empty() && cost + c > opt.max_pass_cost) fits = false;
            if (fits) {
                Q = newQ;
                cost += c;
                mat_scalars += ms;
                sel.push_back((int)i);
                continue;
            }
        }
        blockedX |= qx;
        blockedZ |= qz;
        if ((blockedX & all) == all) break;
    }
    return sel;
}

}  // namespace

std::vector<Pass> schedule_passes(const std::vector<HostOp> &ops, int n, const PlanOptions &opt)
{
    std::vector<Pass> passes;
    const int t = std::min(n, std::min(opt.tile_bits, kMaxTileBits));
    const int lmin = std::min(t, std::max(0, opt.min_run_bits));
    std::vector<char> done(ops.size(), 0);
    size_t first = 0, remaining = ops.size();
    const uint64_t all = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    // tile choice by look-ahead pays when a pass is expensive (a big state) and there is a choice to make: its host cost
    // is ~7 * window^2 op checks per pass (0.03 ms at 64, 0.5 ms at 256) against 0.1 ms (n = 24) ... 100 ms (n = 33) per pass
    const bool look = opt.lookahead > 0 && n >= opt.lookahead_min_qubits && n > t;
    const int window = n >= 28 ? opt.lookahead : std::min(opt.lookahead, 64);
    while (remaining > 0) {
        while (first < ops.size() && done[first]) first++;
        // the first pending op always goes in; if it does not fit beside the forced low qubits,
        // shrink the forced set (its targets still span >= 1 run of the lowest qubits)
        const uint64_t need = ops[first].xmask();
        int forced = lmin;
        while (forced > 0 && popcount64((((1ull << forced) - 1ull)) | need) > t) forced--;
        uint64_t Q = (forced > 0 ? ((1ull << forced) - 1ull) : 0ull) | need;
        if (look) {
            // Grow the tile one op's worth of qubits at a time, each time by the candidate that lets the most
            // additional pending ops run in this pass per tile slot it costs (first-come greedy fills the tile with
            // whatever comes next in program order; on QV-33 this look-ahead needs 66 passes instead of 92).
            while (popcount64(Q) < t) {
                uint64_t q0 = Q;
                const int base = (int)admit(ops, done, first, q0, false, t, opt, all, window).size();
                std::vector<uint64_t> cands;
                int seen = 0;
                for (size_t i = first; i < ops.size() && seen < window; i++) {
                    if (done[i]) continue;
                    seen++;
                    const uint64_t add = ops[i].xmask() & ~Q;
                    if (add == 0 || popcount64(Q | add) > t) continue;
                    if (std::find(cands.begin(), cands.end(), add) == cands.end()) cands.push_back(add);
                }
                uint64_t best = 0;
                double best_gain = 0.0;
                for (uint64_t add : cands) {
                    uint64_t q1 = Q | add;
                    const int cnt = (int)admit(ops, done, first, q1, false, t, opt, all, window).size();
                    const double gain = (double)(cnt - base) / popcount64(add);
                    if (gain > best_gain) { best_gain = gain; best = add; }
                }
                if (best == 0) break;
                Q |= best;
            }
        }
        Pass pass;
        pass.op_idx = admit(ops, done, first, Q, true, t, opt, all, 0);
        // pad the tile with the lowest unused physical qubits (longer contiguous runs)
        for (int qb = 0; qb < n && popcount64(Q) < t; qb++) Q |= 1ull << qb;
        pass.tile_q = mask_to_list(Q);
        pass.L = 0;
        while (pass.L < (int)pass.tile_q.size() && pass.tile_q[pass.L] == pass.L) pass.L++;
        for (int i : pass.op_idx) { done[i] = 1; remaining--; }
        passes.push_back(std::move(pass));
    }
    return passes;
}

template <typename T>
bool encode_pass(const Pass &pass, const std::vector<HostOp> &ops, int n, int batch, PassParams<T> &out,
                 std::vector<cplx> &diag_host, std::string &err, int flags)
{
    const bool pair_ops = (flags & kEncPairOps) != 0;
    std::memset(&out, 0, sizeof(out));
    const int t = (int)pass.tile_q.size();
    out.n = n;
    out.t = t;
    out.L = pass.L;
    out.batch = batch;
    out.tiles_per_state = 1ull << (n - t);
    for (int j = 0; j < t; j++) out.tile_q[j] = (uint8_t)pass.tile_q[j];
    for (int k = 0, lo = 0; k <= t - pass.L; k++) {   // segment masks of the tile number (tile_kernel.cuh: tile_dep)
        const int hi = k == t - pass.L ? 32 : std::max(lo, pass.tile_q[pass.L + k] - pass.L - k);
        out.tile_seg[k] = (uint32_t)((hi >= 32 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull));
        lo = hi;
    }
    auto tile_pos = [&](int qubit) {
        for (int j = 0; j < t; j++) if (pass.tile_q[j] == qubit) return j;
        return -1;
    };
    int nops = 0, mat_used = 0;
    // local bits of a register block (OP_PAIR): the gates' tile bits; widen: filled up to four with spectator bits -- free
    // tile bits, highest first, so that the lanes of a warp keep the low tile bits (distinct shared-memory banks): a thread
    // then holds 16 amplitudes whatever the gates' widths (tile_kernel.cuh: op_pair, RW = 4)
    auto set_block_bits = [&](DevOp &d, std::vector<int> lpos, bool widen) {
        for (int pos = t - 1; pos >= 0 && lpos.size() < 4 && widen; pos--)
            if (std::find(lpos.begin(), lpos.end(), pos) == lpos.end()) lpos.push_back(pos);
        const int R = (int)lpos.size();
        if (R < 1 || R > 4) return false;
        for (int m = 0; m < R; m++) d.tbit[m] = (uint8_t)lpos[m];
        std::vector<int> fix = lpos;
        std::sort(fix.begin(), fix.end());
        d.nfix = (uint8_t)R;
        for (int m = 0; m < 4; m++) {   // swizzled tile index of every local bit (tile_kernel.cuh: swz)
            const uint32_t i = m < R ? 1u << lpos[m] : 0u;
            const uint32_t sw = sizeof(T) == 8 ? (i ^ ((i >> 3) & 7u)) : (i ^ (((i >> 4) & 7u) << 1));
            d.sboff[m] = (uint16_t)sw;
        }
        // segment masks of the group index: bits between consecutive fixed positions move up together
        for (int k = 0; k <= R; k++) {
            const int lo = k == 0 ? 0 : fix[k - 1] - (k - 1);
            const int hi = k == R ? 16 : fix[k] - k;
            d.seg[k] = (uint16_t)(hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u);
        }
        for (int k = R + 1; k < 6; k++) d.seg[k] = 0;
        return true;
    };
    auto pairable = [&](const HostOp &o) {
        return o.kind == HostOp::DENSE && o.q.size() <= 2 && o.ctrl.empty() && o.mask_row < 0;
    };
    out.fan_op = -1;
    // ---- the ops of this pass; phase-ladder terms that commute with everything after them are pulled into a fan -----
    // A term z^(x_a x_b) of a diagonal op can move to the END of the pass when no later op of the pass acts non-
    // diagonally on a or b.  The terms that join a "center" (an in-tile qubit some op of this pass targets) to a "leaf"
    // (a qubit no op of this pass targets, in the tile or not) are collected: together they form a fan (OP_FAN), which
    // the kernel applies in one sweep at ~14 instructions per amplitude whatever the number of terms.  For a QFT pass
    // (H + controlled-phase ladders on six qubits) that is every ladder term whose control lies outside those six --
    // all but a handful -- and the per-amplitude tables shrink to the six qubits themselves.
    std::vector<HostOp> rebuilt;                 // diagonal ops that lost terms to the fan
    rebuilt.reserve(pass.op_idx.size());
    std::vector<const HostOp *> pops;            // the ops the pass executes, in order
    struct FanTerm { int center, leaf; cplx z; };   // leaf = -1: the center's own phase
    std::vector<FanTerm> fan_terms;
    cplx fan_scalar(1, 0);
    for (int oi : pass.op_idx) pops.push_back(&ops[oi]);
    if ((flags & kEncFan) && t >= 9 && t <= 12) {
        uint64_t targets_all = 0, tile_mask = 0;
        bool conditional = false;
        for (int q : pass.tile_q) tile_mask |= 1ull << q;
        for (const HostOp *o : pops) {
            targets_all |= o->xmask();
            conditional = conditional || o->mask_row >= 0;
            for (int c : o->ctrl) conditional = conditional || !((tile_mask >> c) & 1ull);
        }
        std::vector<uint64_t> later(pops.size() + 1, 0);   // non-diagonal targets of the ops after position pi
        for (size_t pi = pops.size(); pi-- > 0;) later[pi] = later[pi + 1] | pops[pi]->xmask();
        std::vector<std::vector<PhaseTerm>> keep(pops.size());
        std::vector<FanTerm> cand;
        cplx cand_scalar(1, 0);
        std::vector<char> touched(pops.size(), 0);
        for (size_t pi = 0; pi < pops.size() && !conditional; pi++) {
            const HostOp &o = *pops[pi];
            if (o.kind != HostOp::DIAG || !o.has_terms) continue;
            const uint64_t lat = later[pi + 1];
            for (const PhaseTerm &tm : o.terms) {
                auto is_center = [&](int q) { return q >= 0 && ((tile_mask >> q) & 1ull) && ((targets_all >> q) & 1ull); };
                auto is_leaf = [&](int q) { return q >= 0 && !((targets_all >> q) & 1ull); };
                bool sunk = false;
                if (tm.a < 0) { cand_scalar *= tm.z; sunk = true; }
                else if (!((lat >> tm.a) & 1ull) && (tm.b < 0 || !((lat >> tm.b) & 1ull))) {
                    if (tm.b < 0 && is_center(tm.a)) { cand.push_back({tm.a, -1, tm.z}); sunk = true; }
                    else if (tm.b >= 0 && is_center(tm.a) && is_leaf(tm.b)) { cand.push_back({tm.a, tm.b, tm.z}); sunk = true; }
                    else if (tm.b >= 0 && is_center(tm.b) && is_leaf(tm.a)) { cand.push_back({tm.b, tm.a, tm.z}); sunk = true; }
                }
                if (sunk) touched[pi] = 1;
                else keep[pi].push_back(tm);
            }
        }
        // is the fan worth a sweep, and does it fit the kernel's tables?
        uint64_t cmask = 0, lin = 0, lout = 0;
        size_t pair_terms = 0;
        for (const FanTerm &f : cand) {
            cmask |= 1ull << f.center;
            if (f.leaf >= 0) { pair_terms++; (((tile_mask >> f.leaf) & 1ull) ? lin : lout) |= 1ull << f.leaf; }
        }
        const int neutral = t - popcount64(cmask) - popcount64(lin);   // tile bits that are neither center nor leaf
        const bool fits = popcount64(cmask) <= kFanMaxCenters && popcount64(lin) <= kFanMaxLeafBits &&
                          popcount64(lout) <= 8 * kFanMaxChunks && std::min(popcount64(cmask), 4) + neutral >= 4;
        if (pair_terms >= 8 && fits && (int)pops.size() < kMaxOps) {
            fan_terms = std::move(cand);
            fan_scalar = cand_scalar;
            std::vector<const HostOp *> np;
            for (size_t pi = 0; pi < pops.size(); pi++) {
                if (!touched[pi]) { np.push_back(pops[pi]); continue; }
                if (keep[pi].empty()) continue;            // the whole op went into the fan
                HostOp r;
                r.kind = HostOp::DIAG;
                r.n_orig = pops[pi]->n_orig;
                uint64_t qm = 0;
                for (const PhaseTerm &tm : keep[pi]) { qm |= 1ull << tm.a; if (tm.b >= 0) qm |= 1ull << tm.b; }
                r.q = mask_to_list(qm);
                r.m.assign((size_t)1 << r.q.size(), cplx(1, 0));
                for (const PhaseTerm &tm : keep[pi]) {
                    const int pa = (int)(std::find(r.q.begin(), r.q.end(), tm.a) - r.q.begin());
                    const int pb = tm.b < 0 ? -1 : (int)(std::find(r.q.begin(), r.q.end(), tm.b) - r.q.begin());
                    for (size_t idx = 0; idx < r.m.size(); idx++)
                        if (((idx >> pa) & 1) && (pb < 0 || ((idx >> pb) & 1))) r.m[idx] *= tm.z;
                }
                // two shrunken tables in a row (what was between them went into the fan) become one while it stays small
                if (!rebuilt.empty() && !np.empty() && np.back() == &rebuilt.back()) {
                    HostOp &prev = rebuilt.back();
                    const uint64_t um = prev.qmask() | r.qmask();
                    if (popcount64(um) <= 6) {
                        const std::vector<int> uq = mask_to_list(um);
                        std::vector<cplx> tab((size_t)1 << uq.size());
                        auto sub = [&](const HostOp &o, size_t idx) {
                            size_t j = 0;
                            for (size_t b = 0; b < o.q.size(); b++) {
                                const size_t pos = (size_t)(std::find(uq.begin(), uq.end(), o.q[b]) - uq.begin());
                                j |= ((idx >> pos) & 1) << b;
                            }
                            return o.m[j];
                        };
                        for (size_t idx = 0; idx < tab.size(); idx++) tab[idx] = sub(prev, idx) * sub(r, idx);
                        prev.q = uq;
                        prev.m = std::move(tab);
                        prev.n_orig += r.n_orig;
                        continue;
                    }
                }
                rebuilt.push_back(std::move(r));
                np.push_back(&rebuilt.back());
            }
            pops = std::move(np);
        }
    }
    for (size_t pi = 0; pi < pops.size(); pi++) {
        const HostOp &op = *pops[pi];
        if (nops >= kMaxOps) { err = "too many ops in one pass"; return false; }
        DevOp &d = out.ops[nops++];
        d.has_mask = op.mask_row >= 0;
        d.mask_row = op.mask_row;
        // register block: this dense gate and the next one share one shared-memory round trip
        if (pair_ops && pairable(op) && pi + 1 < pops.size() && pairable(*pops[pi + 1])) {
            const HostOp &A = op, &B = *pops[pi + 1];
            std::vector<int> a_only, shared, b_only;
            for (int x : A.q) (std::find(B.q.begin(), B.q.end(), x) == B.q.end() ? a_only : shared).push_back(x);
            for (int x : B.q) if (std::find(A.q.begin(), A.q.end(), x) == A.q.end()) b_only.push_back(x);
            std::vector<int> la = a_only, lb = shared, local = a_only;
            la.insert(la.end(), shared.begin(), shared.end());
            lb.insert(lb.end(), b_only.begin(), b_only.end());
            local.insert(local.end(), shared.begin(), shared.end());
            local.insert(local.end(), b_only.begin(), b_only.end());
            const int K1 = (int)A.q.size(), K2 = (int)B.q.size(), S = K1 - (int)shared.size(), R = (int)local.size();
            const int sc = dense_scalars(K1) + dense_scalars(K2);
            if (R <= 4 && mat_used + sc <= kMatScalars) {
                d.kind = OP_PAIR;
                d.k = (uint8_t)K1;
                d.k2 = (uint8_t)K2;
                d.s = (uint8_t)S;
                std::vector<int> lpos;
                for (int m = 0; m < R; m++) {
                    const int p = tile_pos(local[m]);
                    if (p < 0) { err = "internal: target outside tile"; return false; }
                    lpos.push_back(p);
                }
                if (!set_block_bits(d, lpos, false)) { err = "internal: register block without local bits"; return false; }
                d.mat_off = (uint32_t)mat_used;
                const std::vector<cplx> MA = embed_dense(A, la), MB = embed_dense(B, lb);
                pack_dense(out.mats + mat_used, MA.data(), K1);
                pack_dense(out.mats + mat_used + dense_scalars(K1), MB.data(), K2);
                mat_used += sc;
                pi++;  // B consumed
                continue;
            }
        }
        if (op.kind == HostOp::DIAG) {
            // table index layout chosen for the kernel: in-tile qubits first, in ascending tile-bit order (so that
            // the in-tile part of the index is a bit-extract of the tile index), then the out-of-tile qubits
            d.kind = OP_DIAG;
            const int k = (int)op.q.size();
            d.k = (uint8_t)k;
            if (k > kMaxDiagK) { err = "diagonal table too wide"; return false; }
            std::vector<std::pair<int, int>> in;  // (tile pos, op bit)
            std::vector<int> outb;
            for (int m = 0; m < k; m++) {
                const int p = tile_pos(op.q[m]);
                if (p >= 0) in.emplace_back(p, m);
                else outb.push_back(m);
            }
            std::sort(in.begin(), in.end());
            std::vector<int> order;  // new index bit -> old index bit
            uint32_t mask_in = 0;
            for (auto &pm : in) { order.push_back(pm.second); mask_in |= 1u << pm.first; }
            for (int m : outb) order.push_back(m);
            for (int m = 0; m < k; m++)
                d.dsrc[m] = (uint8_t)(m < (int)in.size() ? in[m].first : 64 + op.q[order[m]]);
            d.nfix = (uint8_t)in.size();
            d.ctrl_in = mask_in;  // DIAG: tile-bit mask of the in-tile table qubits
            d.k2 = (uint8_t)out.n_diag;
            out.diag_op[out.n_diag++] = (uint8_t)(nops - 1);
            d.mat_off = (uint32_t)diag_host.size();
            for (int idx = 0; idx < (1 << k); idx++) {
                int old = 0;
                for (int m = 0; m < k; m++) old |= ((idx >> m) & 1) << order[m];
                diag_host.push_back(op.m[old]);
            }
            continue;
        }
        d.kind = op.kind == HostOp::DENSE ? OP_DENSE : (op.kind == HostOp::PERMX ? OP_PERMX : OP_SWAP);
        d.k = (uint8_t)op.q.size();
        if (op.kind == HostOp::DENSE && d.k > kMaxDenseK) { err = "dense block wider than 5 qubits"; return false; }
        std::vector<int> fix;
        for (int m = 0; m < d.k; m++) {
            const int p = tile_pos(op.q[m]);
            if (p < 0) { err = "internal: target outside tile"; return false; }
            d.tbit[m] = (uint8_t)p;
            fix.push_back(p);
        }
        for (int c : op.ctrl) {
            const int p = tile_pos(c);
            if (p >= 0) { d.ctrl_in |= 1u << p; fix.push_back(p); }
            else d.ctrl_out |= 1ull << c;
        }
        std::sort(fix.begin(), fix.end());
        if ((int)fix.size() > kMaxFix) { err = "too many in-tile controls"; return false; }
        d.nfix = (uint8_t)fix.size();
        for (int f = 0; f < d.nfix && f < kMaxFix; f++) d.fixpos[f] = (uint8_t)fix[f];
        if (op.kind == HostOp::DENSE) {
            const int sc = dense_scalars(d.k);
            if (mat_used + sc > kMatScalars) { err = "pass matrix storage exhausted"; return false; }
            d.mat_off = (uint32_t)mat_used;
            pack_dense(out.mats + mat_used, op.m.data(), d.k);
            mat_used += sc;
        }
    }
    if (!fan_terms.empty()) {
        // ---- the fan op (last op of the pass) and its tables
        DevOp &d = out.ops[nops++];
        d.kind = OP_FAN;
        d.mask_row = -1;
        std::vector<int> centers, leaf_in, leaf_out;   // tile positions / tile positions / physical qubits
        for (const auto &f : fan_terms) {
            const int cp = tile_pos(f.center);
            if (std::find(centers.begin(), centers.end(), cp) == centers.end()) centers.push_back(cp);
            if (f.leaf < 0) continue;
            const int lp = tile_pos(f.leaf);
            std::vector<int> &dst = lp >= 0 ? leaf_in : leaf_out;
            const int v = lp >= 0 ? lp : f.leaf;
            if (std::find(dst.begin(), dst.end(), v) == dst.end()) dst.push_back(v);
        }
        std::sort(centers.begin(), centers.end());
        std::sort(leaf_in.begin(), leaf_in.end());
        std::sort(leaf_out.begin(), leaf_out.end());
        const int nc = std::min((int)centers.size(), kFanMaxCenters), m = std::min((int)leaf_in.size(), kFanMaxLeafBits);
        const int nch = std::min(((int)leaf_out.size() + 7) / 8, kFanMaxChunks);   // (the limits were checked when the fan was accepted)
        d.k = (uint8_t)nc;
        d.nfix = (uint8_t)m;
        for (int c = 0; c < nc; c++) d.tbit[c] = (uint8_t)centers[c];
        for (int i = 0; i < m; i++) { d.fixpos[i] = (uint8_t)leaf_in[i]; d.ctrl_in |= 1u << leaf_in[i]; }
        // register bits of the sweep: four tile bits that are not leaves -- centers first, then neutral bits
        std::vector<int> rb;
        for (int c = nc - 1; c >= 0 && rb.size() < 4; c--) rb.push_back(centers[c]);
        for (int pos = t - 1; pos >= 0 && rb.size() < 4; pos--)
            if (std::find(centers.begin(), centers.end(), pos) == centers.end() && std::find(leaf_in.begin(), leaf_in.end(), pos) == leaf_in.end())
                rb.push_back(pos);
        if (rb.size() < 4) { err = "internal: phase fan without four register bits"; return false; }
        std::sort(rb.begin(), rb.end());
        for (int r = 0; r < 4; r++) {
            d.dsrc[r] = (uint8_t)rb[r];
            const auto it = std::find(centers.begin(), centers.end(), rb[r]);
            d.dsrc[4 + r] = it == centers.end() ? (uint8_t)0xFF : (uint8_t)(it - centers.begin());
        }
        d.k2 = (uint8_t)nch;
        d.mat_off = (uint32_t)diag_host.size();
        out.fan_op = nops - 1;
        out.fan_chunks = nch;
        std::memset(out.fan_q, 0xFF, sizeof(out.fan_q));
        for (size_t i = 0; i < leaf_out.size(); i++) out.fan_q[i / 8][i % 8] = (uint8_t)leaf_out[i];
        // data block: [0] scalar, T[c][2^m], X[chunk][c][256]
        const size_t base = diag_host.size();
        diag_host.resize(base + 1 + ((size_t)nc << m) + (size_t)nch * nc * 256, cplx(1, 0));
        diag_host[base] = fan_scalar;
        for (const auto &f : fan_terms) {
            const int c = (int)(std::find(centers.begin(), centers.end(), tile_pos(f.center)) - centers.begin());
            const int lp = f.leaf < 0 ? -1 : tile_pos(f.leaf);
            if (f.leaf < 0 || lp >= 0) {
                const int bit = f.leaf < 0 ? -1 : (int)(std::find(leaf_in.begin(), leaf_in.end(), lp) - leaf_in.begin());
                cplx *T0 = &diag_host[base + 1 + ((size_t)c << m)];
                for (int v = 0; v < (1 << m); v++)
                    if (bit < 0 || ((v >> bit) & 1)) T0[v] *= f.z;
            } else {
                const int li = (int)(std::find(leaf_out.begin(), leaf_out.end(), f.leaf) - leaf_out.begin());
                cplx *X0 = &diag_host[base + 1 + ((size_t)nc << m) + ((size_t)(li / 8) * nc + c) * 256];
                for (int u = 0; u < 256; u++)
                    if ((u >> (li % 8)) & 1) X0[u] *= f.z;
            }
        }
    }
    // ---- runs of unconditioned X / CX / SWAP gates collapse into one linear permutation of the tile (OP_LINPERM) -----
    if ((flags & kEncLinPerm) && t >= 4 && t <= (sizeof(T) == 8 ? 12 : 13)) {
        auto lin_ok = [&](const DevOp &d) {
            if (d.has_mask || d.ctrl_out != 0) return false;
            if (d.kind == OP_PERMX) return popcount64(d.ctrl_in) <= 1;
            return d.kind == OP_SWAP && d.ctrl_in == 0;
        };
        int remap[kMaxOps + 1];
        int w = 0;
        for (int o = 0; o < nops;) {
            int e = o;
            while (e < nops && lin_ok(out.ops[e])) e++;
            if (e - o < 2) { remap[o] = w; if (w != o) out.ops[w] = out.ops[o]; w++; o++; continue; }
            // state_after[d] = state_before[s(d)] with s = f_o . f_(o+1) . ... . f_(e-1) (every gate is its own inverse):
            // evaluate s on 0 and on the unit vectors -- it is affine over GF(2)
            auto src = [&](uint32_t idx) {
                for (int q = e - 1; q >= o; q--) {
                    const DevOp &g = out.ops[q];
                    if (g.kind == OP_PERMX) {
                        if ((idx & g.ctrl_in) == g.ctrl_in) idx ^= 1u << g.tbit[0];
                    } else {
                        const uint32_t a = (idx >> g.tbit[0]) & 1u, b = (idx >> g.tbit[1]) & 1u;
                        if (a != b) idx ^= (1u << g.tbit[0]) | (1u << g.tbit[1]);
                    }
                }
                return idx;
            };
            auto swz_t = [&](uint32_t i) { return sizeof(T) == 8 ? (i ^ ((i >> 3) & 7u)) : (i ^ (((i >> 4) & 7u) << 1)); };
            DevOp L;
            std::memset(&L, 0, sizeof(L));
            L.kind = OP_LINPERM;
            L.k = (uint8_t)std::min(255, e - o);
            L.mask_row = -1;
            const uint32_t c0 = src(0);
            for (int i = 0; i < t; i++) L.lin[i] = (uint16_t)swz_t(src(1u << i) ^ c0);
            L.lin[15] = (uint16_t)swz_t(c0);
            for (int q = o; q < e; q++) remap[q] = w;
            out.ops[w++] = L;
            o = e;
        }
        remap[nops] = w;
        if (w != nops) {
            for (int sl = 0; sl < out.n_diag; sl++) out.diag_op[sl] = (uint8_t)remap[out.diag_op[sl]];
            if (out.fan_op >= 0) out.fan_op = remap[out.fan_op];
            for (int o = w; o < nops; o++) std::memset(&out.ops[o], 0, sizeof(DevOp));
            nops = w;
        }
    }
    out.n_ops = nops;
    out.n_mat = mat_used;
    out.needs_full = (out.n_diag > 0 || out.fan_op >= 0) ? 1 : 0;  // the lean kernel variant carries no diagonal-run / fan code
    for (int o = 0; o < nops; o++) {
        const DevOp &d = out.ops[o];
        if (d.kind == OP_DENSE && d.k >= 3) out.needs_full = 2;
        if (d.kind == OP_PAIR) {
            const int code = (d.k << 4) | (d.k2 << 2) | d.s;
            if (code != ((2 << 4) | (2 << 2) | 2) && code != ((2 << 4) | (2 << 2) | 1) && code != ((1 << 4) | (1 << 2) | 1))
                out.needs_full = 2;
        }
    }
    // ---- register sweeps (tile_kernel.cuh: op_sweep).  Consecutive device ops -- uncontrolled dense gates of <= 2 qubits,
    // runs of diagonal tables, the phase fan -- are applied to the 16 amplitudes a thread holds in registers (the 2^4
    // settings of four "register bits" of the tile index) between one load and one store of the tile.  A dense gate needs
    // its targets among the register bits (a 2-qubit gate on bits (0,1) or (2,3), a 1-qubit gate on any); a table none of
    // whose qubits is a register bit costs one lookup per thread instead of one per amplitude, so the free register bits
    // are the ones the sweep's tables depend on least; the fan needs register bits that are not leaves.
    std::memset(out.sweep_at, 0xFF, sizeof(out.sweep_at));
    out.n_sweeps = 0;
    if (t >= 4 && (out.n_diag > 0 || out.fan_op >= 0)) {
        auto sweepable_dense = [&](const DevOp &d) {
            return d.kind == OP_DENSE && d.k <= 2 && d.nfix == d.k && d.ctrl_out == 0 && !d.has_mask;
        };
        for (int o = 0; o < nops && out.n_sweeps < kMaxSweeps;) {
            const DevOp &first = out.ops[o];
            if (!(sweepable_dense(first) || (first.kind == OP_DIAG && !first.has_mask) || first.kind == OP_FAN)) { o++; continue; }
            std::vector<int> rb;           // register bits in register order (tile-bit positions)
            bool has_fan = false, has_diag = false;
            int oe = o;                    // one past the last op of the sweep
            for (; oe < nops; oe++) {
                DevOp &d = out.ops[oe];
                if (sweepable_dense(d)) {
                    if (d.k == 2) {
                        const auto i0 = std::find(rb.begin(), rb.end(), (int)d.tbit[0]) - rb.begin();
                        const auto i1 = std::find(rb.begin(), rb.end(), (int)d.tbit[1]) - rb.begin();
                        if (i0 < (long)rb.size() || i1 < (long)rb.size()) {
                            // (targets already register bits: only the aligned pairs (0,1) / (2,3) in this order are instantiated)
                            if (!(i1 == i0 + 1 && (i0 == 0 || i0 == 2) && i1 < (long)rb.size())) break;
                            d.s = (uint8_t)i0;
                        } else {
                            if (rb.size() != 0 && rb.size() != 2) break;
                            d.s = (uint8_t)rb.size();
                            rb.push_back(d.tbit[0]);
                            rb.push_back(d.tbit[1]);
                        }
                    } else {
                        const auto i0 = std::find(rb.begin(), rb.end(), (int)d.tbit[0]) - rb.begin();
                        if (i0 == (long)rb.size()) {
                            if (rb.size() >= 4) break;
                            rb.push_back(d.tbit[0]);
                        }
                        d.s = (uint8_t)i0;
                    }
                } else if (d.kind == OP_DIAG && !d.has_mask) {
                    // a long run of tables after a dense gate starts its own sweep: with the gate's targets as register
                    // bits most of its tables would be looked up per amplitude (measured on a QFT without fans: 92 against
                    // 84 ms); short runs of small tables (what a fan leaves behind) stay with the gate
                    if (!rb.empty() && (oe == o || out.ops[oe - 1].kind != OP_DIAG)) {
                        int run = 0;
                        bool small = true;
                        while (oe + run < nops && out.ops[oe + run].kind == OP_DIAG) { small = small && out.ops[oe + run].k <= 6; run++; }
                        if (run > 4 || !small) break;
                    }
                    has_diag = true;
                } else if (d.kind == OP_FAN) {
                    has_fan = true;
                    oe++;
                    break;                  // (the fan is the last op of its pass)
                } else {
                    break;
                }
            }
            if (oe - o == 1 && out.ops[o].kind == OP_DENSE) { o++; continue; }   // a lone dense gate: the ordinary path
            // fill up the register bits: fewest dependent tables first; then bits >= 3 (a warp then reads whole 128-byte
            // rows); then high bits.  Leaves of the fan are excluded (its factors must be thread constants).
            std::vector<int> users(t, 0);
            uint32_t fan_leaves = 0;
            for (int q = o; q < oe; q++) {
                if (out.ops[q].kind == OP_DIAG)
                    for (int pos = 0; pos < t; pos++) users[pos] += (out.ops[q].ctrl_in >> pos) & 1u;
                if (out.ops[q].kind == OP_FAN) fan_leaves = out.ops[q].ctrl_in;
            }
            std::vector<int> cand;
            for (int pos = 0; pos < t; pos++)
                if (std::find(rb.begin(), rb.end(), pos) == rb.end() && !((fan_leaves >> pos) & 1u)) cand.push_back(pos);
            std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
                if (users[a] != users[b]) return users[a] < users[b];
                if ((a < 3) != (b < 3)) return b < 3;
                return a > b;
            });
            bool ok = true;
            for (size_t ci = 0; rb.size() < 4; ci++) {
                if (ci >= cand.size()) { ok = false; break; }
                rb.push_back(cand[ci]);
            }
            for (int b : rb) if ((fan_leaves >> b) & 1u) ok = false;   // (a dense target is never a leaf; belt and braces)
            if (!ok) { err = "internal: no register bits for a sweep"; return false; }
            uint32_t rmask = 0;
            for (int b : rb) rmask |= 1u << b;
            // every run of tables inside the sweep: the tables that ignore the register bits go first (fixpos[1] of them)
            for (int q = o; q < oe;) {
                if (out.ops[q].kind != OP_DIAG) { q++; continue; }
                int qe = q;
                while (qe + 1 < oe && out.ops[qe + 1].kind == OP_DIAG) qe++;
                std::stable_partition(out.ops + q, out.ops + qe + 1, [&](const DevOp &x) { return (x.ctrl_in & rmask) == 0; });
                const int slot0 = std::min_element(out.ops + q, out.ops + qe + 1, [](const DevOp &a, const DevOp &b) { return a.k2 < b.k2; })->k2;
                int n_indep = 0;
                for (int r = q; r <= qe; r++) {
                    out.ops[r].k2 = (uint8_t)(slot0 + (r - q));
                    out.diag_op[slot0 + (r - q)] = (uint8_t)r;
                    n_indep += (out.ops[r].ctrl_in & rmask) == 0;
                    out.ops[r].s = (out.ops[r].ctrl_in & rmask) != 0;  // DIAG: 1 = looked up per amplitude
                }
                out.ops[q].fixpos[0] = 0;
                out.ops[q].fixpos[1] = (uint8_t)n_indep;
                q = qe + 1;
            }
            typename PassParams<T>::Sweep &S = out.sweeps[out.n_sweeps];
            S.first = (uint8_t)o;
            S.len = (uint8_t)(oe - o);
            std::vector<int> asc = rb;
            std::sort(asc.begin(), asc.end());
            for (int b = 0; b < 4; b++) { S.rb[b] = (uint8_t)rb[b]; S.asc[b] = (uint8_t)asc[b]; }
            if (has_fan) {   // which center (if any) each register bit of the sweep is
                DevOp &f = out.ops[oe - 1];
                for (int b = 0; b < 4; b++) {
                    f.dsrc[b] = (uint8_t)rb[b];
                    f.dsrc[4 + b] = 0xFF;
                    for (int c = 0; c < f.k; c++) if (f.tbit[c] == rb[b]) f.dsrc[4 + b] = (uint8_t)c;
                }
            }
            out.sweep_at[o] = (uint8_t)out.n_sweeps++;
            (void)has_fan; (void)has_diag;
            o = oe;
        }
        // every diagonal run and the fan must sit in a sweep (the lean + diagonal kernel variant has no other path for them)
        for (int o = 0, covered = -1; o < nops; o++) {
            if (out.sweep_at[o] != 0xFF) covered = o + out.sweeps[out.sweep_at[o]].len - 1;
            if ((out.ops[o].kind == OP_DIAG || out.ops[o].kind == OP_FAN) && o > covered && !out.ops[o].has_mask) out.needs_full = 2;
        }
    }
    out.cond_mask = 0;
    for (int o = 0; o < nops; o++)
        if (out.ops[o].ctrl_out != 0 || out.ops[o].has_mask) out.cond_mask |= 1ull << o;
    // Passes that the lean kernel variant serves (register blocks, <= 2-qubit gates and permutations only:
    // the FP64-pipe-bound passes of dense circuits) get the fast forms of tile_kernel.cuh: every register block is widened
    // to four local bits, a lone uncontrolled 2-qubit gate becomes a block of one gate, and in FP64 the matrices are kept
    // in the three-multiplication form, which the kernel reads from its parameters through the uniform datapath
    // (mat_form 1; FP32: plain elements, mat_form 2).  Per-register matrices (kEncPlainMats: they live in global memory) and
    // every other pass keep natural blocks and plain elements (mat_form 0).
    out.mat_form = 0;
    if (pair_ops && !(flags & kEncPlainMats) && out.needs_full == 0 && t >= 4) {
        out.mat_form = sizeof(T) == 8 ? 1 : 2;
        for (int o = 0; o < nops; o++) {
            DevOp &d = out.ops[o];
            if (d.kind == OP_DENSE && d.k == 2 && d.nfix == 2 && d.ctrl_out == 0 && !d.has_mask) {
                d.kind = OP_PAIR;
                d.k2 = 0;
                d.s = 0;
            }
            if (d.kind == OP_PAIR) {
                const int R = std::max<int>(d.k, d.s + d.k2);
                std::vector<int> lpos(d.tbit, d.tbit + R);
                if (!set_block_bits(d, lpos, true)) { err = "internal: register block without local bits"; return false; }
            }
            if (out.mat_form != 1) continue;
            if (d.kind == OP_DENSE && d.k <= 2) to_mul3(out.mats + d.mat_off, d.k);
            if (d.kind == OP_PAIR) {
                to_mul3(out.mats + d.mat_off, d.k);
                if (d.k2) to_mul3(out.mats + d.mat_off + dense_scalars(d.k), d.k2);
            }
        }
    }
    // Per-tile staging of small diagonal sub-tables.  For a given tile the table-index bits that come from qubits outside
    // the tile are constants, so the op only ever reads the 2^nfix consecutive entries that start at its per-tile base
    // (in-tile qubits are the low index bits, see above).  Ops with few in-tile qubits -- the CP ladders of a QFT are
    // tables over ten qubits of which one to five lie in the tile -- are copied into shared memory at the start of the
    // tile and looked up there instead of in L2 (one global load per ENTRY per tile instead of one per amplitude per op).
    out.n_stab = 0;
    for (int o = 0; o < nops; o++) out.ops[o].stage_off = kNotStaged;
    for (int sl = 0; sl < out.n_diag; sl++) {
        DevOp &d = out.ops[out.diag_op[sl]];
        // (only tables that are looked up per amplitude are worth a slot: the others cost one lookup per thread anyway)
        if ((flags & kEncNoStage) || !d.s || d.nfix > kStabMaxBits || d.has_mask || d.ctrl_out) continue;
        const int need = std::max(8, 1 << d.nfix);  // (chunks of 8 entries)
        if (out.n_stab + need > kStabEntries) continue;
        d.stage_off = (uint16_t)out.n_stab;
        for (int c = 0; c < need / 8; c++) out.stab_slot[out.n_stab / 8 + c] = (uint8_t)sl;
        out.n_stab += need;
    }
    return true;
}

template bool encode_pass<double>(const Pass &, const std::vector<HostOp> &, int, int, PassParams<double> &,
                                  std::vector<cplx> &, std::string &, int);
template bool encode_pass<float>(const Pass &, const std::vector<HostOp> &, int, int, PassParams<float> &,
                                 std::vector<cplx> &, std::string &, int);

}  // namespace cb200
