// gates.cpp -- see gates.hpp.  Standard Qiskit gate definitions, little-endian
// (matrix index bit m <-> m-th target qubit).
#include "gates.hpp"

#include <cmath>
#include <map>

namespace cb200 {

namespace {
const cplx I(0.0, 1.0);
const double PI = 3.14159265358979323846;

std::vector<cplx> u3(double th, double ph, double la)
{
    const double c = std::cos(th / 2), s = std::sin(th / 2);
    return {cplx(c), -std::exp(I * la) * s, std::exp(I * ph) * s, std::exp(I * (ph + la)) * c};
}

struct Base {
    int nt;  // targets
    int np;  // minimum params
    std::vector<cplx> (*make)(const double *p, int np);
};

std::vector<cplx> diag2(cplx a, cplx b) { return {a, 0, 0, b}; }

const std::map<std::string, Base> &bases()
{
    static const std::map<std::string, Base> B = {
        {"id", {1, 0, [](const double *, int) { return std::vector<cplx>{1, 0, 0, 1}; }}},
        {"x", {1, 0, [](const double *, int) { return std::vector<cplx>{0, 1, 1, 0}; }}},
        {"y", {1, 0, [](const double *, int) { return std::vector<cplx>{0, -I, I, 0}; }}},
        {"z", {1, 0, [](const double *, int) { return diag2(1, -1); }}},
        {"h", {1, 0, [](const double *, int) {
                   const double r = 1.0 / std::sqrt(2.0);
                   return std::vector<cplx>{r, r, r, -r};
               }}},
        {"s", {1, 0, [](const double *, int) { return diag2(1, I); }}},
        {"sdg", {1, 0, [](const double *, int) { return diag2(1, -I); }}},
        {"t", {1, 0, [](const double *, int) { return diag2(1, std::exp(I * (PI / 4))); }}},
        {"tdg", {1, 0, [](const double *, int) { return diag2(1, std::exp(-I * (PI / 4))); }}},
        {"sx", {1, 0, [](const double *, int) {
                    return std::vector<cplx>{cplx(0.5, 0.5), cplx(0.5, -0.5), cplx(0.5, -0.5), cplx(0.5, 0.5)};
                }}},
        {"sxdg", {1, 0, [](const double *, int) {
                      return std::vector<cplx>{cplx(0.5, -0.5), cplx(0.5, 0.5), cplx(0.5, 0.5), cplx(0.5, -0.5)};
                  }}},
        {"sy", {1, 0, [](const double *, int) {
                    return std::vector<cplx>{cplx(0.5, 0.5), cplx(-0.5, -0.5), cplx(0.5, 0.5), cplx(0.5, 0.5)};
                }}},
        {"sydg", {1, 0, [](const double *, int) {
                      return std::vector<cplx>{cplx(0.5, -0.5), cplx(0.5, -0.5), cplx(-0.5, 0.5), cplx(0.5, -0.5)};
                  }}},
        {"rx", {1, 1, [](const double *p, int) {
                    const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                    return std::vector<cplx>{c, -I * s, -I * s, c};
                }}},
        {"ry", {1, 1, [](const double *p, int) {
                    const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                    return std::vector<cplx>{c, -s, s, c};
                }}},
        {"rz", {1, 1, [](const double *p, int) { return diag2(std::exp(-I * (p[0] / 2)), std::exp(I * (p[0] / 2))); }}},
        {"p", {1, 1, [](const double *p, int) { return diag2(1, std::exp(I * p[0])); }}},
        {"u2", {1, 2, [](const double *p, int) { return u3(PI / 2, p[0], p[1]); }}},
        {"u3", {1, 3, [](const double *p, int np) {
                    auto m = u3(p[0], p[1], p[2]);
                    if (np > 3) for (auto &v : m) v *= std::exp(I * p[3]);
                    return m;
                }}},
        {"r", {1, 2, [](const double *p, int) {
                   const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                   return std::vector<cplx>{c, -I * std::exp(-I * p[1]) * s, -I * std::exp(I * p[1]) * s, c};
               }}},
        {"swap", {2, 0, [](const double *, int) {
                      return std::vector<cplx>{1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
                  }}},
        {"iswap", {2, 0, [](const double *, int) {
                       return std::vector<cplx>{1, 0, 0, 0, 0, 0, I, 0, 0, I, 0, 0, 0, 0, 0, 1};
                   }}},
        {"dcx", {2, 0, [](const double *, int) {
                     return std::vector<cplx>{1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0};
                 }}},
        {"ecr", {2, 0, [](const double *, int) {
                     const double r = 1.0 / std::sqrt(2.0);
                     return std::vector<cplx>{0, r, 0, I * r, r, 0, -I * r, 0, 0, I * r, 0, r, -I * r, 0, r, 0};
                 }}},
        {"rxx", {2, 1, [](const double *p, int) {
                     const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                     return std::vector<cplx>{c, 0, 0, -I * s, 0, c, -I * s, 0, 0, -I * s, c, 0, -I * s, 0, 0, c};
                 }}},
        {"ryy", {2, 1, [](const double *p, int) {
                     const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                     return std::vector<cplx>{c, 0, 0, I * s, 0, c, -I * s, 0, 0, -I * s, c, 0, I * s, 0, 0, c};
                 }}},
        {"rzz", {2, 1, [](const double *p, int) {
                     const cplx a = std::exp(-I * (p[0] / 2)), b = std::exp(I * (p[0] / 2));
                     return std::vector<cplx>{a, 0, 0, 0, 0, b, 0, 0, 0, 0, b, 0, 0, 0, 0, a};
                 }}},
        {"rzx", {2, 1, [](const double *p, int) {
                     const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                     return std::vector<cplx>{c, 0, -I * s, 0, 0, c, 0, I * s, -I * s, 0, c, 0, 0, I * s, 0, c};
                 }}},
        {"xxmyy", {2, 2, [](const double *p, int) {
                       const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                       const cplx e = std::exp(I * p[1]);
                       return std::vector<cplx>{c, 0, 0, -I * s / e, 0, 1, 0, 0, 0, 0, 1, 0, -I * s * e, 0, 0, c};
                   }}},
        {"xxpyy", {2, 2, [](const double *p, int) {
                       const double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                       const cplx e = std::exp(I * p[1]);
                       return std::vector<cplx>{1, 0, 0, 0, 0, c, -I * s / e, 0, 0, -I * s * e, c, 0, 0, 0, 0, 1};
                   }}},
    };
    return B;
}

// instruction name -> base gate name; every qubit before the base's targets is a control
const std::map<std::string, std::string> &aliases()
{
    static const std::map<std::string, std::string> A = {
        {"u1", "p"}, {"u", "u3"}, {"sz", "s"}, {"szdg", "sdg"}, {"v", "sx"}, {"vdg", "sxdg"}, {"id2", "id"},
        {"cx", "x"}, {"ccx", "x"}, {"mcx", "x"}, {"cy", "y"}, {"ccy", "y"}, {"mcy", "y"},
        {"cz", "z"}, {"ccz", "z"}, {"mcz", "z"}, {"ch", "h"}, {"mch", "h"},
        {"cs", "s"}, {"mcs", "s"}, {"csdg", "sdg"}, {"ct", "t"}, {"mct", "t"},
        {"csx", "sx"}, {"mcsx", "sx"}, {"csxdg", "sxdg"},
        {"crx", "rx"}, {"mcrx", "rx"}, {"cry", "ry"}, {"mcry", "ry"}, {"crz", "rz"}, {"mcrz", "rz"},
        {"cp", "p"}, {"mcp", "p"}, {"cu1", "p"}, {"mcu1", "p"},
        {"cu2", "u2"}, {"mcu2", "u2"}, {"cu3", "u3"}, {"mcu3", "u3"}, {"cu", "u3"}, {"mcu", "u3"},
        {"cr", "r"}, {"mcr", "r"}, {"cswap", "swap"}, {"mcswap", "swap"}, {"cecr", "ecr"},
    };
    return A;
}
}  // namespace

bool lookup_gate(const std::string &name, int n_qubits, const double *params, int n_params,
                 GateSpec &out, std::string &err)
{
    out = GateSpec();
    if (name == "gp") {
        if (n_params < 1) { err = "gp needs 1 parameter"; return false; }
        out.is_global_phase = true;
        out.phase = params[0];
        return true;
    }
    std::string base = name;
    bool aliased = false;
    auto a = aliases().find(name);
    if (a != aliases().end()) { base = a->second; aliased = true; }
    auto b = bases().find(base);
    if (b == bases().end()) { err = "unsupported instruction '" + name + "'"; return false; }
    const Base &B = b->second;
    if (n_params < B.np) { err = "instruction '" + name + "' needs " + std::to_string(B.np) + " parameter(s)"; return false; }
    if (n_qubits < B.nt) { err = "instruction '" + name + "' needs at least " + std::to_string(B.nt) + " qubit(s)"; return false; }
    // un-aliased / uncontrolled names must have exactly the base arity
    const bool controlled = aliased && name != "u1" && name != "u" && name != "sz" && name != "szdg" && name != "v" &&
                            name != "vdg" && name != "id2";
    if (!controlled && n_qubits != B.nt && !(name == "id2" && n_qubits == 2)) {
        err = "instruction '" + name + "' takes " + std::to_string(B.nt) + " qubit(s)";
        return false;
    }
    if (controlled && n_qubits < B.nt + 1) { err = "instruction '" + name + "' needs control qubit(s)"; return false; }
    out.n_targets = (name == "id2") ? 1 : B.nt;
    out.mat = B.make(params, n_params);
    return true;
}

int update_param_arity(const std::string &n)
{
    static const std::map<std::string, int> A = {
        {"rx", 1}, {"ry", 1}, {"rz", 1}, {"p", 1}, {"u1", 1}, {"crx", 1}, {"cry", 1}, {"crz", 1}, {"cp", 1},
        {"cu1", 1}, {"rxx", 1}, {"ryy", 1}, {"rzz", 1}, {"rzx", 1},
        {"u2", 2}, {"r", 2}, {"cu2", 2}, {"cr", 2}, {"mcu2", 2}, {"mcr", 2},
        {"u3", 3}, {"cu3", 3}, {"mcu3", 3}, {"u", 4}, {"cu", 4},
    };
    auto it = A.find(n);
    return it == A.end() ? 0 : it->second;
}

bool is_diagonal(const std::vector<cplx> &mat, int dim)
{
    for (int r = 0; r < dim; r++)
        for (int c = 0; c < dim; c++)
            if (r != c && mat[(size_t)r * dim + c] != cplx(0.0, 0.0)) return false;
    return true;
}

}  // namespace cb200
