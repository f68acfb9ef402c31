// qasm2.cpp -- OpenQASM 2 text -> CUNQA circuit JSON (host only).
// Replaces /root/reference/src/utils/helpers/qasm2_to_json.hpp:520-549 (qasm2_to_json): same output object
// {"instructions", "num_qubits", "num_clbits", "quantum_registers", "classical_registers"}, registers numbered in
// declaration order, gate statements mapped by name onto the wire instruction of the same name.  Differences, all
// supersets: statements are split at ';' (not at line ends), angle expressions are a full + - * / ( ) pi grammar
// (the reference evaluates only chains of * and /), a gate applied to a whole register is broadcast, `reset`,
// `barrier` and `id` are accepted, the builtins U / CX map to u3 / cx, and arities are checked against the gate table.
#include "executor.hpp"

#include <cctype>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "gates.hpp"
#include "json_min.hpp"

namespace cb200 {
namespace {

struct Reg { int offset, size; };

std::string strip(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

// recursive-descent evaluator: expr = term {(+|-) term}; term = unary {(*|/) unary}; unary = [-|+] atom;
// atom = number | pi | ( expr )
struct Expr {
    const std::string &s;
    size_t p = 0;
    explicit Expr(const std::string &text) : s(text) {}
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
    double atom()
    {
        ws();
        if (p < s.size() && s[p] == '(') {
            p++;
            const double v = expr();
            ws();
            if (p >= s.size() || s[p] != ')') throw std::runtime_error("qasm: missing ')' in '" + s + "'");
            p++;
            return v;
        }
        if (s.compare(p, 2, "pi") == 0) { p += 2; return M_PI; }
        size_t used = 0;
        double v;
        try { v = std::stod(s.substr(p), &used); } catch (...) { throw std::runtime_error("qasm: cannot evaluate '" + s + "'"); }
        p += used;
        return v;
    }
    double unary()
    {
        ws();
        if (p < s.size() && (s[p] == '-' || s[p] == '+')) { const bool neg = s[p] == '-'; p++; const double v = unary(); return neg ? -v : v; }
        return atom();
    }
    double term()
    {
        double v = unary();
        for (;;) {
            ws();
            if (p < s.size() && s[p] == '*') { p++; v *= unary(); }
            else if (p < s.size() && s[p] == '/') { p++; v /= unary(); }
            else return v;
        }
    }
    double expr()
    {
        double v = term();
        for (;;) {
            ws();
            if (p < s.size() && s[p] == '+') { p++; v += term(); }
            else if (p < s.size() && s[p] == '-') { p++; v -= term(); }
            else return v;
        }
    }
    double all()
    {
        const double v = expr();
        ws();
        if (p != s.size()) throw std::runtime_error("qasm: cannot evaluate '" + s + "'");
        return v;
    }
};

std::vector<std::string> split_top(const std::string &s, char sep)
{
    std::vector<std::string> out;
    int depth = 0;
    std::string cur;
    for (char ch : s) {
        if (ch == '(') depth++;
        if (ch == ')') depth--;
        if (ch == sep && depth == 0) { out.push_back(strip(cur)); cur.clear(); }
        else cur.push_back(ch);
    }
    if (!strip(cur).empty() || !out.empty()) out.push_back(strip(cur));
    return out;
}

// "q[3]" -> {offset+3}; "q" -> every index of the register
std::vector<int> resolve(const std::string &arg, const std::map<std::string, Reg> &regs, const char *what)
{
    const size_t lb = arg.find('[');
    const std::string name = strip(arg.substr(0, lb));
    auto it = regs.find(name);
    if (it == regs.end()) throw std::runtime_error(std::string("qasm: unknown ") + what + " register '" + name + "'");
    std::vector<int> out;
    if (lb == std::string::npos) {
        for (int i = 0; i < it->second.size; i++) out.push_back(it->second.offset + i);
        return out;
    }
    const size_t rb = arg.find(']', lb);
    if (rb == std::string::npos) throw std::runtime_error("qasm: missing ']' in '" + arg + "'");
    const int idx = std::stoi(arg.substr(lb + 1, rb - lb - 1));
    if (idx < 0 || idx >= it->second.size) throw std::runtime_error("qasm: index out of range in '" + arg + "'");
    out.push_back(it->second.offset + idx);
    return out;
}

Json int_array(const std::vector<int> &v)
{
    Json a = Json::array();
    for (int x : v) a.arr.push_back(Json::integer(x));
    return a;
}

}  // namespace

std::string qasm2_to_json(const std::string &qasm)
{
    // comments out, then statements at ';'
    std::string text;
    for (size_t i = 0; i < qasm.size(); i++) {
        if (qasm[i] == '/' && i + 1 < qasm.size() && qasm[i + 1] == '/') { while (i < qasm.size() && qasm[i] != '\n') i++; text.push_back('\n'); }
        else text.push_back(qasm[i]);
    }
    std::map<std::string, Reg> qregs, cregs;
    Json qreg_json = Json::object(), creg_json = Json::object(), insts = Json::array();
    int nq = 0, nc = 0, ignored = 0;
    auto declare = [&](const std::string &body, std::map<std::string, Reg> &regs, Json &out, int &total) {
        const size_t lb = body.find('['), rb = body.find(']');
        if (lb == std::string::npos || rb == std::string::npos) throw std::runtime_error("qasm: malformed register declaration '" + body + "'");
        const std::string name = strip(body.substr(0, lb));
        const int size = std::stoi(body.substr(lb + 1, rb - lb - 1));
        if (size < 1 || regs.count(name)) throw std::runtime_error("qasm: bad register declaration '" + body + "'");
        regs[name] = Reg{total, size};
        std::vector<int> idx;
        for (int i = 0; i < size; i++) idx.push_back(total + i);
        out.set(name, int_array(idx));
        total += size;
    };
    for (const std::string &raw : split_top(text, ';')) {
        const std::string st = strip(raw);
        if (st.empty()) continue;
        size_t e = 0;
        while (e < st.size() && (std::isalnum((unsigned char)st[e]) || st[e] == '_')) e++;
        const std::string head = st.substr(0, e);
        const std::string rest = strip(st.substr(e));
        if (head == "OPENQASM" || head == "include" || head == "barrier") continue;
        if (head == "qreg") { declare(rest, qregs, qreg_json, nq); continue; }
        if (head == "creg") { declare(rest, cregs, creg_json, nc); continue; }
        if (head == "gate" || head == "opaque" || head == "if") throw std::runtime_error("qasm: '" + head + "' statements are not supported");
        if (head == "measure") {
            const size_t arrow = rest.find("->");
            if (arrow == std::string::npos) throw std::runtime_error("qasm: measure without '->'");
            const std::vector<int> qs = resolve(strip(rest.substr(0, arrow)), qregs, "quantum");
            const std::vector<int> cs = resolve(strip(rest.substr(arrow + 2)), cregs, "classical");
            if (qs.size() != cs.size()) throw std::runtime_error("qasm: measure register sizes differ");
            for (size_t i = 0; i < qs.size(); i++) {  // one qubit per instruction (aer_simulator_adapter.cpp:185-190)
                Json m = Json::object();
                m.set("name", Json::string("measure"));
                m.set("qubits", int_array({qs[i]}));
                m.set("clbits", int_array({cs[i]}));
                insts.arr.push_back(std::move(m));
            }
            continue;
        }
        // gate application: name [ (params) ] arg {, arg}
        std::string name = head, args = rest;
        std::vector<double> params;
        if (!args.empty() && args[0] == '(') {
            int depth = 0;
            size_t close = std::string::npos;
            for (size_t i = 0; i < args.size(); i++) {
                if (args[i] == '(') depth++;
                if (args[i] == ')' && --depth == 0) { close = i; break; }
            }
            if (close == std::string::npos) throw std::runtime_error("qasm: missing ')' in '" + st + "'");
            for (const std::string &p : split_top(args.substr(1, close - 1), ',')) params.push_back(Expr(p).all());
            args = strip(args.substr(close + 1));
        }
        if (name == "U") name = "u3";
        if (name == "CX") name = "cx";
        if (qregs.empty()) { ignored++; continue; }
        std::vector<std::vector<int>> operands;
        for (const std::string &a : split_top(args, ',')) operands.push_back(resolve(a, qregs, "quantum"));
        if (operands.empty()) throw std::runtime_error("qasm: '" + st + "' has no operands");
        // broadcast: whole-register operands must agree in size; single qubits are repeated
        size_t reps = 1;
        for (auto &o : operands) if (o.size() > 1) { if (reps != 1 && reps != o.size()) throw std::runtime_error("qasm: register sizes differ in '" + st + "'"); reps = o.size(); }
        for (size_t r = 0; r < reps; r++) {
            std::vector<int> qs;
            for (auto &o : operands) qs.push_back(o.size() > 1 ? o[r] : o[0]);
            if (name != "reset" && name != "id") {
                GateSpec g;
                std::string err;
                if (!lookup_gate(name, (int)qs.size(), params.data(), (int)params.size(), g, err)) throw std::runtime_error("qasm: " + err);
            } else if (qs.size() != 1 || !params.empty()) {
                throw std::runtime_error("qasm: '" + name + "' takes one qubit");
            }
            Json in = Json::object();
            in.set("name", Json::string(name));
            in.set("qubits", int_array(qs));
            if (!params.empty()) {
                Json pa = Json::array();
                for (double v : params) pa.arr.push_back(Json::number(v));
                in.set("params", std::move(pa));
            }
            insts.arr.push_back(std::move(in));
        }
    }
    Json out = Json::object();
    out.set("instructions", std::move(insts));
    out.set("num_qubits", Json::integer(nq));
    out.set("num_clbits", Json::integer(nc));
    out.set("quantum_registers", std::move(qreg_json));
    out.set("classical_registers", std::move(creg_json));
    (void)ignored;
    return out.dump();
}

}  // namespace cb200
