// executor.cpp -- whole-task execution (see executor.hpp).
#include "executor.hpp"
#include "dist_plan.hpp"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <random>
#include <set>
#include <thread>
#include <atomic>

namespace cb200 {

namespace {

struct Inst {
    std::string name;
    std::vector<int> qubits, clbits, l_clbits, r_clbits;
    std::vector<double> params;
    std::vector<cplx> matrix;  // unitary / cunitary (row-major) or diagonal entries
    int matrix_dim = 0;
    std::vector<std::string> qpus;
    std::vector<Inst> sub;
    std::string operation = "and";
    int condition = 1;
};

std::vector<int> int_list(const Json &j)
{
    std::vector<int> out;
    for (const Json &v : j.arr) out.push_back((int)v.as_int());
    return out;
}

Inst parse_inst(const Json &j);

std::vector<Inst> parse_insts(const Json &arr)
{
    if (!arr.is_array()) throw std::runtime_error("'instructions' must be an array");
    std::vector<Inst> out;
    for (const Json &j : arr.arr) out.push_back(parse_inst(j));
    return out;
}

// field layout per instruction family: /root/reference/src/utils/constants.hpp.in:570-916
Inst parse_inst(const Json &j)
{
    Inst in;
    in.name = j.at("name").as_string();
    if (const Json *q = j.find("qubits")) in.qubits = int_list(*q);
    if (const Json *c = j.find("clbits")) in.clbits = int_list(*c);
    if (const Json *c = j.find("l_clbits")) in.l_clbits = int_list(*c);
    if (const Json *c = j.find("r_clbits")) in.r_clbits = int_list(*c);
    if (const Json *p = j.find("params"))
        for (const Json &v : p->arr) in.params.push_back(v.as_double());
    if (const Json *q = j.find("qpus"))
        for (const Json &v : q->arr) in.qpus.push_back(v.as_string());
    if (const Json *s = j.find("instructions")) in.sub = parse_insts(*s);
    if (const Json *o = j.find("operation")) in.operation = o->as_string();
    if (const Json *c = j.find("condition")) in.condition = (int)c->as_int();
    if (const Json *m = j.find("matrix")) {
        // unitary: [[ [ [re,im], ... ], ... ]]   diagonal: [[ [re,im], ... ]]   (one extra list level)
        const Json &inner = m->at(0);
        if (in.name == "diagonal") {
            for (const Json &z : inner.arr) in.matrix.emplace_back(z.at(0).as_double(), z.at(1).as_double());
            in.matrix_dim = (int)in.matrix.size();
        } else {
            in.matrix_dim = (int)inner.size();
            for (const Json &row : inner.arr) {
                if ((int)row.size() != in.matrix_dim) throw std::runtime_error("matrix of '" + in.name + "' is not square");
                for (const Json &z : row.arr) in.matrix.emplace_back(z.at(0).as_double(), z.at(1).as_double());
            }
        }
    }
    return in;
}

int log2_exact(int v)
{
    int k = 0;
    while ((1 << k) < v) k++;
    if ((1 << k) != v) throw std::runtime_error("matrix dimension is not a power of two");
    return k;
}

bool is_classical_or_remote(const std::string &n)
{
    static const std::set<std::string> S = {"measure", "reset", "cif", "copy", "send", "recv", "qsend", "qrecv", "expose", "rcontrol"};
    return S.count(n) != 0;
}

// queue one unitary instruction.  qmap maps instruction qubit -> register qubit.
template <typename Q, typename MapF>
void queue_gate(Q &q, const Inst &in, MapF qmap, const uint8_t *flags, bool reverse_unitary_qubits)
{
    std::vector<int> qs;
    for (int x : in.qubits) qs.push_back(qmap(x));
    if (in.name == "unitary") {
        const int k = log2_exact(in.matrix_dim);
        if ((int)qs.size() != k) throw std::runtime_error("unitary: matrix size does not match the qubit list");
        if (reverse_unitary_qubits) std::reverse(qs.begin(), qs.end());
        q.apply_matrix(qs.data(), k, nullptr, 0, in.matrix.data(), flags);
    } else if (in.name == "cunitary") {
        const int k = log2_exact(in.matrix_dim);
        const int nc = (int)qs.size() - k;
        if (nc < 1) throw std::runtime_error("cunitary: needs control qubit(s)");
        q.apply_matrix(qs.data() + nc, k, qs.data(), nc, in.matrix.data(), flags);
    } else if (in.name == "diagonal") {
        const int k = log2_exact(in.matrix_dim);
        if ((int)qs.size() != k) throw std::runtime_error("diagonal: size does not match the qubit list");
        q.apply_diagonal(qs.data(), k, in.matrix.data(), flags);
    } else {
        q.apply_named(in.name, qs.data(), (int)qs.size(), in.params.data(), (int)in.params.size(), flags);
    }
}

struct TaskCfg {
    std::string id;
    int num_qubits = 0, num_clbits = 0;
    uint64_t shots = 1024;
    bool has_seed = false;
    uint64_t seed = 0;
    int precision = 64;
    bool is_dynamic = false;
    std::vector<std::string> sending_to;
    std::vector<Inst> insts;
    int n_comm = -1;
    bool bugcompat_unitary_order = false;
    bool bugcompat_teledata = true;   // X <- sent qubit's outcome, Z <- comm qubit's: the reference's literal code (:653-658), quirk Q6
    bool terminal_sampling = true;
    bool share_prefix = true;
    bool defer_measurement = true;
    int fusion_enable = -1, fusion_max_qubit = -1;  // -1 = the engine's default
    int max_batch = 0;
    int shard_min_qubits = 30;      // with several target devices: registers at least this wide are sharded over them
    // out-of-band state return: amplitudes at chosen basis states / a marginal distribution in the result JSON
    std::vector<uint64_t> amplitude_indices;
    std::vector<int> marginal_qubits;
};

// config keys honoured (cf. quantum_task_to_AER, aer_helpers.hpp:93-140)
TaskCfg parse_task(const Json &t)
{
    TaskCfg c;
    c.id = t.contains("id") ? t.at("id").as_string() : std::string("task");
    const Json &cfg = t.at("config");
    c.num_qubits = (int)cfg.at("num_qubits").as_int();
    c.num_clbits = cfg.contains("num_clbits") ? (int)cfg.at("num_clbits").as_int() : c.num_qubits;
    if (cfg.contains("shots")) c.shots = (uint64_t)cfg.at("shots").as_int();
    if (cfg.contains("seed")) { c.has_seed = true; c.seed = (uint64_t)cfg.at("seed").as_int(); }
    if (cfg.contains("precision")) {
        const std::string p = cfg.at("precision").as_string();
        if (p == "single") c.precision = 32;
        else if (p == "double") c.precision = 64;
        else throw std::runtime_error("precision must be 'single' or 'double'");
    }
    if (cfg.contains("method")) {
        const std::string m = cfg.at("method").as_string();
        if (m != "automatic" && m != "statevector")
            throw std::runtime_error("method '" + m + "' is not supported by the B200 backend (statevector only)");
    }
    if (cfg.contains("fusion_enable")) c.fusion_enable = cfg.at("fusion_enable").as_bool() ? 1 : 0;
    if (cfg.contains("fusion_max_qubit")) c.fusion_max_qubit = (int)cfg.at("fusion_max_qubit").as_int();
    if (cfg.contains("n_communication_qubits")) c.n_comm = (int)cfg.at("n_communication_qubits").as_int();
    if (cfg.contains("b200_bugcompat_unitary_order")) c.bugcompat_unitary_order = cfg.at("b200_bugcompat_unitary_order").as_bool();
    if (cfg.contains("b200_bugcompat_teledata")) c.bugcompat_teledata = cfg.at("b200_bugcompat_teledata").as_bool();
    if (cfg.contains("b200_teledata_corrections")) {
        const std::string v = cfg.at("b200_teledata_corrections").as_string();
        if (v == "reference") c.bugcompat_teledata = true;
        else if (v == "protocol") c.bugcompat_teledata = false;
        else throw std::runtime_error("b200_teledata_corrections must be 'reference' or 'protocol'");
    }
    if (cfg.contains("b200_share_prefix")) c.share_prefix = cfg.at("b200_share_prefix").as_bool();
    if (cfg.contains("b200_terminal_sampling")) c.terminal_sampling = cfg.at("b200_terminal_sampling").as_bool();
    if (cfg.contains("b200_defer_measurement")) c.defer_measurement = cfg.at("b200_defer_measurement").as_bool();
    if (cfg.contains("b200_max_batch")) c.max_batch = (int)cfg.at("b200_max_batch").as_int();
    if (cfg.contains("b200_shard_min_qubits")) c.shard_min_qubits = (int)cfg.at("b200_shard_min_qubits").as_int();
    if (const Json *a = cfg.find("b200_amplitude_indices"))
        for (const Json &v : a->arr) {
            if (v.as_int() < 0) throw std::runtime_error("b200_amplitude_indices: negative index");
            c.amplitude_indices.push_back((uint64_t)v.as_int());
        }
    if (const Json *m = cfg.find("b200_marginal_qubits")) c.marginal_qubits = int_list(*m);
    c.is_dynamic = t.contains("is_dynamic") ? t.at("is_dynamic").as_bool() : false;
    if (const Json *s = t.find("sending_to"))
        for (const Json &v : s->arr) c.sending_to.push_back(v.as_string());
    c.insts = parse_insts(t.at("instructions"));
    if (c.num_qubits < 1) throw std::runtime_error("num_qubits must be >= 1");
    return c;
}

uint64_t fresh_seed()
{
    std::random_device rd;
    return ((uint64_t)rd() << 32) ^ rd();
}

std::string bits_to_string(const std::vector<uint8_t> &creg, int zero, int ncl)
{
    // clbit c at string index ncl-1-c (aer_simulator_adapter.cpp:780-789; aer_helpers.hpp:175-179)
    std::string s((size_t)ncl, '0');
    for (int c = 0; c < ncl; c++)
        if (creg[zero + c]) s[ncl - 1 - c] = '1';
    return s;
}

Json counts_json(const std::map<std::string, uint64_t> &counts)
{
    Json j = Json::object();
    j.obj.reserve(counts.size());  // keys of a std::map are unique: append directly (Json::set searches linearly)
    for (auto &kv : counts) j.obj.emplace_back(kv.first, Json::integer((long long)kv.second));
    return j;
}

// Aer's `fusion_enable` / `fusion_max_qubit` (aer_helpers.hpp:17-86); absent keys restore the engine's defaults so that a
// cached State never inherits the previous job's setting
void apply_fusion_cfg(State &st, const TaskCfg &c)
{
    st.rq().set_fusion(c.fusion_enable < 0 ? st.opt.fuse : c.fusion_enable != 0,
                       c.fusion_max_qubit < 0 ? st.opt.max_dense_qubits : c.fusion_max_qubit);
}

bool cif_op(const std::string &op, bool a, bool b)
{
    if (op == "and") return a & b;
    if (op == "or") return a | b;
    if (op == "xor") return a ^ b;
    if (op == "andn") return a & !b;
    if (op == "orn") return a | !b;
    if (op == "xorn") return a ^ !b;
    throw std::runtime_error("unknown cif operation '" + op + "'");
}

}  // namespace

bool Backend::shards(int n, int shard_min_qubits) const
{
    const int w = (int)devices_.size();
    if (w < 2 || (w & (w - 1))) return false;
    int g = 0;
    while ((1 << g) < w) g++;
    return n >= shard_min_qubits && n - g >= 12;
}

State &Backend::get_state(int n, int precision, int batch, bool shard)
{
    if (state_ && state_->n == n && state_->precision == precision && state_->batch == batch && state_->is_facade() == shard) {
        state_->reset();
        return *state_;
    }
    state_.reset();
    if (shard) state_.reset(new State(n, precision, devices_));   // one state over all target devices (dist.cu: ShardGroup)
    else state_.reset(new State(n, precision, device_, batch));
    return *state_;
}

std::string Backend::execute(const std::string &tasks_json)
{
    try {
        const Json msg = Json::parse(tasks_json);
        if (msg.is_object() && msg.contains("b200_batch")) {
            // {"b200_batch": [message, ...]}: independent messages as separate registers (execute_many) in one call
            std::vector<std::string> texts;
            for (const Json &m : msg.at("b200_batch").arr) texts.push_back(m.dump());
            Json res = Json::array();
            for (const std::string &r : execute_many(texts)) res.arr.push_back(Json::parse(r));
            Json out = Json::object();
            out.set("b200_batch_results", std::move(res));
            return out.dump();
        }
        return run(msg).dump();
    } catch (const std::exception &e) {
        // same contract as the reference adapters (aer_simulator_adapter.cpp:836-840)
        Json err = Json::object();
        err.set("ERROR", Json::string(e.what()));
        return err.dump();
    }
}

// {"params":[...],"shots":N}: rewrite the cached task in circuit order by gate arity
// (/root/reference/src/quantum_task.cpp:39-108)
static void update_params(Json &task, const Json &msg)
{
    const Json &params = msg.at("params");
    size_t counter = 0;
    Json *insts = task.find("instructions");
    if (!insts || insts->arr.empty()) throw std::runtime_error("Circuit not sent before updating parameters.");
    for (Json &inst : insts->arr) {
        const std::string name = inst.at("name").as_string();
        int ar = update_param_arity(name);
        if (ar == 0) continue;
        Json *p = inst.find("params");
        if (!p) throw std::runtime_error("Error updating parameters: instruction '" + name + "' has no params");
        if ((name == "u" || name == "cu") && p->arr.size() < 4) ar = (int)p->arr.size();  // builder emits 3-parameter u
        for (int k = 0; k < ar; k++) {
            if (counter >= params.size()) throw std::runtime_error("Error updating parameters: too few parameters (check correct size).");
            if ((size_t)k >= p->arr.size()) p->arr.push_back(Json::number(0));
            p->arr[k] = Json::number(params.at(counter++).as_double());
        }
    }
    Json *cfg = task.find("config");
    if (cfg && msg.contains("shots")) cfg->set("shots", Json::integer(msg.at("shots").as_int()));
}

Json Backend::run(const Json &msg)
{
    if (msg.is_array()) {
        std::vector<Json> tasks(msg.arr.begin(), msg.arr.end());
        if (tasks.empty()) throw std::runtime_error("empty task list");
        return run_dynamic(tasks);
    }
    if (!msg.is_object()) throw std::runtime_error("quantum task must be a JSON object");
    if (msg.contains("instructions") && msg.contains("config")) {
        cached_ = msg;
        has_cached_ = true;
        parsed_.reset();
    } else if (msg.contains("params")) {
        if (!has_cached_) throw std::runtime_error("Circuit not sent before updating parameters.");
        update_params(cached_, msg);
        if (parsed_) {
            // the structure is unchanged: copy the rewritten angles (and shots) into the parsed task instead of re-parsing
            TaskCfg &pc = *std::static_pointer_cast<TaskCfg>(parsed_);
            const Json &insts = cached_.at("instructions");
            bool same = insts.arr.size() == pc.insts.size();
            for (size_t i = 0; same && i < pc.insts.size(); i++) {
                const Json *pj = insts.arr[i].find("params");
                if (!pj) continue;
                pc.insts[i].params.resize(pj->arr.size());
                for (size_t k = 0; k < pj->arr.size(); k++) pc.insts[i].params[k] = pj->arr[k].as_double();
            }
            if (!same) parsed_.reset();
            else if (msg.contains("shots")) pc.shots = (uint64_t)msg.at("shots").as_int();
        }
    } else {
        throw std::runtime_error("message has neither instructions+config nor params");
    }
    const Json &task = cached_;
    // static vs dynamic (aer_simple_simulator.cpp:8-22): the flag decides; a "static" task that
    // measures mid-circuit is also sent down the per-shot path, as Aer does internally.
    bool dynamic = task.contains("is_dynamic") && task.at("is_dynamic").as_bool();
    if (!dynamic) {
        const Json &insts = task.at("instructions");
        uint64_t measured = 0;
        for (const Json &j : insts.arr) {
            const std::string name = j.at("name").as_string();
            if (name == "measure") {
                for (const Json &q : j.at("qubits").arr)
                    if (q.as_int() >= 0 && q.as_int() < 64) measured |= 1ull << q.as_int();
            } else if (is_classical_or_remote(name)) {
                dynamic = true;
            } else if (const Json *qs = j.find("qubits")) {
                for (const Json &q : qs->arr)
                    if (q.as_int() >= 0 && q.as_int() < 64 && ((measured >> q.as_int()) & 1ull)) dynamic = true;
            }
        }
    }
    if (dynamic) {
        Json r = run_dynamic({task});
        Json out = Json::object();
        const std::string id = task.contains("id") ? task.at("id").as_string() : std::string("task");
        out.set("counts", r.at("id_counts").at(id));
        out.set("time_taken", r.at("time_taken"));
        if (r.contains("b200")) out.set("b200", r.at("b200"));
        return out;
    }
    if (!parsed_) parsed_ = std::make_shared<TaskCfg>(parse_task(task));
    return run_static_parsed(parsed_.get());
}

Json Backend::run_static(const Json &task)
{
    const TaskCfg c = parse_task(task);
    return run_static_parsed(&c);
}

Json Backend::run_static_parsed(const void *parsed_task)
{
    const auto t_in = std::chrono::high_resolution_clock::now();
    const TaskCfg &c = *static_cast<const TaskCfg *>(parsed_task);
    const auto t0 = std::chrono::high_resolution_clock::now();
    State &st = get_state(c.num_qubits, c.precision, 1, shards(c.num_qubits, c.shard_min_qubits));
    apply_fusion_cfg(st, c);
    const Stats before = st.stats;
    const uint64_t gates_before = st.queue.gates;
    std::vector<std::pair<int, int>> meas;  // (qubit, clbit)
    bool save_state = false;
    for (const Inst &in : c.insts) {
        if (in.name == "measure") {
            if (in.qubits.size() != in.clbits.size()) throw std::runtime_error("measure: qubits/clbits size mismatch");
            for (size_t i = 0; i < in.qubits.size(); i++) {
                if (in.qubits[i] < 0 || in.qubits[i] >= c.num_qubits) throw std::runtime_error("measure: qubit out of range");
                if (in.clbits[i] < 0 || in.clbits[i] >= c.num_clbits) throw std::runtime_error("measure: clbit out of range");
                meas.emplace_back(in.qubits[i], in.clbits[i]);
            }
        } else if (in.name == "save_state") {
            save_state = true;
        } else {
            queue_gate(st, in, [](int q) { return q; }, nullptr, false);
        }
    }
    const uint64_t seed = c.has_seed ? c.seed : fresh_seed();
    std::vector<uint64_t> samples(c.shots);
    const auto t_q = std::chrono::high_resolution_clock::now();
    st.flush();
    const auto t_f = std::chrono::high_resolution_clock::now();
    if (getenv("CB200_TRACE")) st.sync();  // (tracing only) split the wait for the tile passes from the sampling
    const auto t_w = std::chrono::high_resolution_clock::now();
    st.sample(0, c.shots, seed, samples.data());
    const std::vector<uint64_t> raw_samples = (save_state && c.num_qubits > 20) ? samples : std::vector<uint64_t>();
    const auto t_s = std::chrono::high_resolution_clock::now();
    Json out = Json::object();
    std::string key((size_t)c.num_clbits, '0');
    if (c.num_clbits <= 64) {
        // classical register of every shot as an integer, sorted: equal outcomes are adjacent, one string each
        bool identity = (int)meas.size() == c.num_clbits;   // measure q -> c with q == c for every classical bit, each once
        for (size_t i = 0; i < meas.size() && identity; i++) identity = meas[i].first == (int)i && meas[i].second == (int)i;
        if (!identity) {
            for (uint64_t &s : samples) {
                uint64_t r = 0;
                for (auto &qc : meas) r = (r & ~(1ull << qc.second)) | (((s >> qc.first) & 1ull) << qc.second);
                s = r;
            }
        } else if (c.num_clbits < 64) {
            const uint64_t keep = (1ull << c.num_clbits) - 1ull;
            for (uint64_t &s : samples) s &= keep;
        }
        std::sort(samples.begin(), samples.end());
        static const std::array<std::array<char, 8>, 256> kByteBits = [] {   // "00000000" .. "11111111", most significant bit first
            std::array<std::array<char, 8>, 256> tb{};
            for (int v = 0; v < 256; v++)
                for (int b = 0; b < 8; b++) tb[v][7 - b] = ((v >> b) & 1) ? '1' : '0';
            return tb;
        }();
        // the counts object is rendered straight into JSON text: {"<bits>":n,...} (a 16-qubit ansatz at 1024 shots has
        // about a thousand distinct outcomes; building and dumping a node per outcome was a quarter of the evaluation)
        std::string text;
        text.reserve(samples.size() * (size_t)(c.num_clbits + 8) + 2);
        text += '{';
        char num[24];
        for (size_t i = 0; i < samples.size();) {
            size_t j = i;
            while (j < samples.size() && samples[j] == samples[i]) j++;
            {   // classical bit 0 is the rightmost character: whole bytes from the table, the leading bits one by one
                int b = 0;
                for (; b + 8 <= c.num_clbits; b += 8) std::memcpy(&key[c.num_clbits - 8 - b], kByteBits[(samples[i] >> b) & 0xFFull].data(), 8);
                for (; b < c.num_clbits; b++) key[c.num_clbits - 1 - b] = ((samples[i] >> b) & 1ull) ? '1' : '0';
            }
            if (i) text += ',';
            text += '"';
            text += key;
            text += "\":";
            size_t v = j - i;
            int len = 0;
            do { num[sizeof num - 1 - len++] = (char)('0' + v % 10); v /= 10; } while (v);
            text.append(num + sizeof num - len, (size_t)len);
            i = j;
        }
        text += '}';
        out.set("counts", Json::raw(std::move(text)));
    } else {
        std::map<std::string, uint64_t> counts;
        for (uint64_t s : samples) {
            std::fill(key.begin(), key.end(), '0');
            for (auto &qc : meas) key[c.num_clbits - 1 - qc.second] = ((s >> qc.first) & 1ull) ? '1' : '0';
            counts[key]++;
        }
        out.set("counts", counts_json(counts));
    }
    auto amplitudes_json = [&](const std::vector<uint64_t> &idx) {
        std::vector<double> amps(2 * idx.size());
        st.get_amplitudes(0, idx.data(), idx.size(), amps.data());
        Json sv = Json::array();
        sv.arr.reserve(idx.size());
        for (size_t i = 0; i < idx.size(); i++) {
            Json z = Json::array();
            z.arr.push_back(Json::number(amps[2 * i]));
            z.arr.push_back(Json::number(amps[2 * i + 1]));
            sv.arr.push_back(std::move(z));
        }
        return sv;
    };
    if (save_state) {
        // the reference returns the whole vector as [[re, im], ...] (cunqa/result.py:151-187): feasible up to 20 qubits.
        // Above that the amplitudes of the basis states the shots landed on are returned instead (at most `shots` of
        // them, the heavy part of the distribution), with their indices; b200_amplitude_indices asks for chosen ones.
        if (c.num_qubits <= 20) {
            std::vector<uint64_t> idx((size_t)1 << c.num_qubits);
            for (uint64_t i = 0; i < idx.size(); i++) idx[i] = i;
            out.set("statevector", amplitudes_json(idx));
        } else {
            std::vector<uint64_t> idx(raw_samples);
            std::sort(idx.begin(), idx.end());
            idx.erase(std::unique(idx.begin(), idx.end()), idx.end());
            Json sampled = Json::object();
            Json ji = Json::array();
            for (uint64_t i : idx) ji.arr.push_back(Json::integer((long long)i));
            sampled.set("indices", std::move(ji));
            sampled.set("amplitudes", amplitudes_json(idx));
            out.set("statevector_sampled", std::move(sampled));
        }
    }
    if (!c.amplitude_indices.empty()) {
        for (uint64_t i : c.amplitude_indices)
            if (c.num_qubits < 64 && (i >> c.num_qubits)) throw std::runtime_error("b200_amplitude_indices: index out of range");
        out.set("b200_amplitudes", amplitudes_json(c.amplitude_indices));
    }
    if (!c.marginal_qubits.empty()) {
        // P(qubits = j), index bit m <-> b200_marginal_qubits[m]: the on-device equivalent of recombine_probs
        // (/root/reference/src/utils/probabilities/process_counts.hpp:71-131), exact instead of estimated from counts
        const int k = (int)c.marginal_qubits.size();
        if (k > 20) throw std::runtime_error("b200_marginal_qubits: at most 20 qubits");
        std::vector<double> p((size_t)1 << k);
        st.probabilities(0, c.marginal_qubits.data(), k, p.data());
        Json jp = Json::array();
        jp.arr.reserve(p.size());
        for (double v : p) jp.arr.push_back(Json::number(v));
        out.set("b200_marginal", std::move(jp));
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    out.set("time_taken", Json::number(std::chrono::duration<double>(t1 - t0).count()));
    if (getenv("CB200_TRACE")) {
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[cb200] static task: parse %.0f us, queue+fuse %.0f us, plan+launch %.0f us, wait for passes %.0f us, sample %.0f us, counts %.0f us\n",
                us(t_in, t0), us(t0, t_q), us(t_q, t_f), us(t_f, t_w), us(t_w, t_s), us(t_s, t1));
    }
    last_mode = 0;
    last_stats = st.stats;
    last_stats.passes -= before.passes;
    last_stats.launches -= before.launches;
    last_stats.bytes -= before.bytes;
    last_stats.fused_ops -= before.fused_ops;
    last_stats.h2d_bytes -= before.h2d_bytes;
    last_stats.d2h_bytes -= before.d2h_bytes;
    last_stats.gates = st.queue.gates - gates_before;
    last_stats.relabelled_swaps = st.queue.relabelled_swaps;
    return out;
}

// ---- dynamic path: per-shot interpreter, shots batched on the device ---------------------------
namespace {

struct CommPair {
    int q0, q1;
    bool idle = true;
    std::string sendr, recvr, proto;
    int label = 0;
};

struct TaskRt {
    const TaskCfg *cfg;
    int zq = 0, zc = 0;
    size_t pc = 0;
    bool finished = false, b_td = false, b_tg = false, b_cc = false, cat = false;
};

using Bits = std::vector<uint8_t>;

// Deferred measurement (one state, all shots sampled at the end): a classical bit is then SYMBOLIC -- 0 / 1 are
// constants, kSym + q means "the outcome recorded in register qubit q", which is read from the final sample.
constexpr uint8_t kSym = 2;
constexpr uint8_t kNotSym = 64;   // kNotSym + q: "NOT the outcome recorded in qubit q" (a cif that tests for 0)
constexpr uint8_t kConj = 128;    // kConj + i: the conjunction of literals stored in entry i of the batch's table
inline bool sym_negated(uint8_t v) { return v >= kNotSym && v < kConj; }
inline int sym_qubit(uint8_t v) { return v >= kNotSym ? v - kNotSym : v - kSym; }
struct DeferAbort {};  // the circuit needs a real mid-circuit collapse: fall back to the batched interpreter
enum QState : uint8_t { Q_LIVE = 0, Q_FROZEN, Q_FROZEN_RESET };

// how an instruction acts on each of its qubits: 'k' = control, 'd' = target of a diagonal matrix (both commute
// with a Z-basis measurement of that qubit), 'x' = anything else
std::vector<char> action_kinds(const Inst &in)
{
    const int nq = (int)in.qubits.size();
    std::vector<char> kinds((size_t)nq, 'x');
    int nt = nq;
    bool diag = false;
    if (in.name == "diagonal") {
        diag = true;
    } else if (in.name == "unitary" || in.name == "cunitary") {
        nt = log2_exact(in.matrix_dim);
        diag = is_diagonal(in.matrix, 1 << nt);
    } else {
        GateSpec g;
        std::string err;
        if (!lookup_gate(in.name, nq, in.params.data(), (int)in.params.size(), g, err)) throw std::runtime_error(err);
        if (g.is_global_phase) return std::vector<char>((size_t)nq, 'd');
        nt = g.n_targets;
        diag = is_diagonal(g.mat, 1 << nt);
    }
    for (int i = 0; i < nq; i++) kinds[i] = i < nq - nt ? 'k' : (diag ? 'd' : 'x');
    return kinds;
}

// the instruction with one more control qubit (register index), for a condition that is a deferred outcome
void queue_gate_controlled_pos(State &st, const Inst &in, const std::vector<int> &qs, const std::vector<int> &extra, bool reverse_unitary_qubits);

// lits: kSym + q (fires on outcome 1) or kNotSym + q (fires on outcome 0: that control qubit is X-conjugated);
// the instruction fires when ALL literals hold
void queue_gate_controlled(State &st, const Inst &in, const std::vector<int> &qs, const std::vector<uint8_t> &lits, bool reverse_unitary_qubits)
{
    std::vector<int> extra;
    for (uint8_t l : lits) extra.push_back(sym_qubit(l));
    for (uint8_t l : lits) if (sym_negated(l)) { const int q = sym_qubit(l); st.apply_named("x", &q, 1, nullptr, 0, nullptr); }
    queue_gate_controlled_pos(st, in, qs, extra, reverse_unitary_qubits);
    for (uint8_t l : lits) if (sym_negated(l)) { const int q = sym_qubit(l); st.apply_named("x", &q, 1, nullptr, 0, nullptr); }
}

void queue_gate_controlled_pos(State &st, const Inst &in, const std::vector<int> &qs, const std::vector<int> &extra, bool reverse_unitary_qubits)
{
    std::vector<int> ctrls, targets;
    std::vector<cplx> mat;
    if (in.name == "unitary" || in.name == "cunitary") {
        const int k = log2_exact(in.matrix_dim);
        const int nc = (int)qs.size() - k;
        if (nc < 0 || (in.name == "unitary" && nc != 0)) throw std::runtime_error("unitary: matrix size does not match the qubit list");
        ctrls.assign(qs.begin(), qs.begin() + nc);
        targets.assign(qs.begin() + nc, qs.end());
        if (in.name == "unitary" && reverse_unitary_qubits) std::reverse(targets.begin(), targets.end());
        mat = in.matrix;
    } else if (in.name == "diagonal") {
        const int k = log2_exact(in.matrix_dim);
        if ((int)qs.size() != k) throw std::runtime_error("diagonal: size does not match the qubit list");
        std::vector<int> all = qs;
        all.insert(all.end(), extra.begin(), extra.end());
        const int ne = (int)extra.size();
        std::vector<cplx> d((size_t)1 << (k + ne), cplx(1, 0));
        const size_t on = (((size_t)1 << ne) - 1) << k;  // every extra control set
        for (int i = 0; i < (1 << k); i++) d[(size_t)i | on] = in.matrix[i];
        st.apply_diagonal(all.data(), k + ne, d.data(), nullptr);
        return;
    } else {
        GateSpec g;
        std::string err;
        if (!lookup_gate(in.name, (int)qs.size(), in.params.data(), (int)in.params.size(), g, err)) throw std::runtime_error(err);
        if (g.is_global_phase) {  // a conditioned global phase is a (multi-controlled) phase gate on the condition qubits
            const int ne = (int)extra.size();
            std::vector<cplx> d((size_t)1 << ne, cplx(1, 0));
            d.back() = std::exp(cplx(0, g.phase));
            st.apply_diagonal(extra.data(), ne, d.data(), nullptr);
            return;
        }
        const int nc = (int)qs.size() - g.n_targets;
        ctrls.assign(qs.begin(), qs.begin() + nc);
        targets.assign(qs.begin() + nc, qs.end());
        mat = g.mat;
    }
    ctrls.insert(ctrls.end(), extra.begin(), extra.end());
    st.apply_matrix(targets.data(), (int)targets.size(), ctrls.data(), (int)ctrls.size(), mat.data(), nullptr);
}

}  // namespace

// one register of a batched run: the deferred interpreter queues the group's gates on register `reg` of the shared
// State and leaves its symbolic classical register here; sampling and decoding happen once for all registers
struct Backend::RegJob {
    State *st = nullptr;
    int reg = 0;
    // outputs
    std::vector<uint8_t> symbols;                 // classical register after the run: 0 / 1 / kSym + q
    struct Layout { std::string id; int zc, ncl; };
    std::vector<Layout> tasks;
    uint64_t shots = 0, seed = 0, gates = 0;
    bool used_teledata = false, teledata_reference = true, multi = false;
};

Json Backend::run_dynamic(const std::vector<Json> &tasks_json, RegJob *job)
{
    std::vector<TaskCfg> cfgs;
    for (const Json &t : tasks_json) cfgs.push_back(parse_task(t));
    const bool multi = cfgs.size() > 1;
    const TaskCfg &c0 = cfgs[0];
    // joint register: sum of task widths + communication pairs (aer_simulator_adapter.cpp:855-874)
    int n_comm = 0;
    if (multi) {
        n_comm = c0.n_comm >= 0 ? c0.n_comm : 2;
        if (n_comm % 2) n_comm++;
    }
    int n_total = n_comm, ncl_total = 0;
    for (const TaskCfg &c : cfgs) { n_total += c.num_qubits; ncl_total += c.num_clbits; }
    const uint64_t shots = c0.shots;
    const uint64_t seed = c0.has_seed ? c0.seed : fresh_seed();
    const int precision = 64;  // the reference forces double on this path (:946)

    // does any task talk to another process?  then shots run one by one, in the reference's order
    bool external_cc = false;
    if (!multi)
        for (const Inst &in : c0.insts)
            if (in.name == "send" || in.name == "recv") external_cc = true;
    if (external_cc && (!send_ || !recv_)) throw std::runtime_error("send/recv used but no classical channel is attached");

    // batch size: as many shots as fit a memory budget
    uint64_t B = shots;
    {
        const double state_bytes = 16.0 * std::ldexp(1.0, n_total);
        const double budget = 48.0 * 1073741824.0;
        uint64_t cap = (uint64_t)std::max(1.0, std::floor(budget / state_bytes));
        cap = std::min<uint64_t>(cap, 4096);
        if (c0.max_batch > 0) cap = std::min<uint64_t>(cap, (uint64_t)c0.max_batch);
        B = std::max<uint64_t>(1, std::min(B, cap));
        if (external_cc) B = 1;
        if (!job && shards(n_total, c0.shard_min_qubits)) B = 1;   // a sharded register holds one state: shots run one by one
    }

    const auto t0 = std::chrono::high_resolution_clock::now();
    std::map<std::string, std::map<std::string, uint64_t>> counters;
    for (const TaskCfg &c : cfgs) counters[c.id];
    if (shots == 0) B = 1;
    State *stp = nullptr;
    Stats before;
    uint64_t gates_before = 0;
    bool deferred_used = false;
    bool used_teledata = false;

    // one batch of nb shots starting at shot0 (the last batch may carry surplus states; their results are dropped).
    // deferred = one state for ALL shots: measurements only freeze their qubit, classical bits are symbolic, feed-
    // forward becomes quantum control, and every shot is drawn from the final state (principle of deferred
    // measurement; same joint distribution of all classical outcomes).  Throws DeferAbort when the circuit needs a
    // real collapse (a measured qubit is driven again, a live qubit is reset, a condition is not a single outcome).
    auto run_batch = [&](State &st, const int nb, const uint64_t shot0, const uint64_t live, const bool deferred) {
        std::vector<uint8_t> qstate((size_t)n_total, Q_LIVE);
        std::vector<char> touched((size_t)n_total, 0);
        std::vector<std::vector<uint8_t>> conj;  // deferred mode: conjunctions of literals (nested / multi-bit conditions)
        auto literals = [&](uint8_t v) { return v >= kConj ? conj[v - kConj] : std::vector<uint8_t>{v}; };
        auto make_cond = [&](std::vector<uint8_t> lits) -> uint8_t {
            std::sort(lits.begin(), lits.end());
            lits.erase(std::unique(lits.begin(), lits.end()), lits.end());
            for (uint8_t a : lits)
                for (uint8_t b : lits)
                    if (sym_qubit(a) == sym_qubit(b) && sym_negated(a) != sym_negated(b)) return 0;  // q and not q
            if (lits.size() == 1) return lits[0];
            if (conj.size() >= 127) throw DeferAbort();
            conj.push_back(lits);
            return (uint8_t)(kConj + conj.size() - 1);
        };
        std::vector<TaskRt> T(cfgs.size());
        std::map<std::string, size_t> by_id;
        {
            int zq = 0, zc = 0;
            for (size_t i = 0; i < cfgs.size(); i++) {
                T[i].cfg = &cfgs[i];
                T[i].zq = zq;
                T[i].zc = zc;
                zq += cfgs[i].num_qubits;
                zc += cfgs[i].num_clbits;
                by_id[cfgs[i].id] = i;
            }
        }
        std::vector<CommPair> pairs;
        for (int i = 0; i < n_comm; i += 2) pairs.push_back(CommPair{n_total - n_comm + i, n_total - n_comm + i + 1});
        std::vector<Bits> creg((size_t)ncl_total, Bits(nb, 0));
        std::map<std::string, std::deque<Bits>> q_td, q_tg;
        std::map<std::pair<std::string, std::string>, std::deque<Bits>> q_cc;
        std::vector<uint64_t> ordinal(nb, 0), ctr(nb);
        // shared prefix: until the first measurement every shot of the batch is the same state, so gates are
        // applied to state 0 only (per-state mask: the other states' tiles are skipped, no traffic) and state 0 is
        // broadcast right before the first measurement
        bool diverged = nb == 1 || !c0.share_prefix;
        Bits only0(nb, 0);
        only0[0] = 1;
        auto diverge = [&]() {
            if (diverged) return;
            st.broadcast_state0();
            diverged = true;
        };

        auto peer = [&](const std::string &id) -> TaskRt & {
            auto it = by_id.find(id);
            if (it == by_id.end()) throw std::runtime_error("remote instruction refers to unknown task '" + id + "'");
            return T[it->second];
        };
        // measure (optionally conditioned by `flags`) every state; returns the outcome bits
        auto measure = [&](int q, bool reset_after, const uint8_t *flags) -> Bits {
            if (deferred) {
                if (flags) {
                    if (flags[0] == 0) return Bits(1, 0);
                    if (flags[0] != 1) throw DeferAbort();  // measurement under an undecided condition
                }
                if (qstate[q] == Q_FROZEN_RESET) return Bits(1, 0);           // |0> again: outcome 0, nothing to do
                if (!touched[q] && qstate[q] == Q_LIVE) return Bits(1, 0);    // still the initial |0>
                qstate[q] = reset_after ? Q_FROZEN_RESET : Q_FROZEN;          // (a frozen qubit keeps its record)
                return Bits(1, (uint8_t)(kSym + q));
            }
            diverge();
            std::vector<int8_t> forced;
            for (int b = 0; b < nb; b++) ctr[b] = ((shot0 + b + 1) << 20) + ordinal[b];
            if (flags) {
                forced.assign(nb, -1);
                for (int b = 0; b < nb; b++) if (!flags[b]) forced[b] = -2;
            }
            std::vector<int> out(nb);
            st.measure(q, seed, ctr.data(), flags ? forced.data() : nullptr, out.data(), reset_after);
            Bits bits(nb, 0);
            for (int b = 0; b < nb; b++) {
                if (out[b] >= 0) { bits[b] = (uint8_t)out[b]; ordinal[b]++; }
            }
            return bits;
        };
        auto any = [&](const Bits &f) { return std::any_of(f.begin(), f.end(), [](uint8_t v) { return v != 0; }); };
        auto all = [&](const Bits &f) { return std::all_of(f.begin(), f.end(), [](uint8_t v) { return v != 0; }); };
        // per-state condition of a gate: the caller's flags, restricted to state 0 while the batch has not diverged
        Bits pre_flags;
        auto gate_flags = [&](const Bits *flags) -> const uint8_t * {
            if (!diverged) {
                pre_flags = only0;
                if (flags) pre_flags[0] = (*flags)[0];
                return pre_flags.data();
            }
            return (flags && !all(*flags)) ? flags->data() : nullptr;
        };
        // deferred mode: bring a qubit that holds a recorded outcome (or any other state) back to |0> WITHOUT a collapse.
        // Possible when no classical bit or pending message still refers to its record and the qubit has become
        // disentangled from the rest of the register -- as the measured qubits of a completed teleportation or remote
        // gate have (the corrections leave them in |+>): then one single-qubit unitary maps its pure state to |0>.
        auto sym_referenced = [&](int q) -> bool {
            const uint8_t v = (uint8_t)(kSym + q);  // (negated symbols only live inside a cif body, never in storage)
            for (const Bits &b : creg) if (b[0] == v) return true;
            for (auto &kv : q_td) for (const Bits &b : kv.second) if (b[0] == v) return true;
            for (auto &kv : q_tg) for (const Bits &b : kv.second) if (b[0] == v) return true;
            for (auto &kv : q_cc) for (const Bits &b : kv.second) if (b[0] == v) return true;
            return false;
        };
        auto recycle = [&](int q) {
            if (job) throw DeferAbort();  // batched registers: no mid-circuit device queries (the group runs on its own instead)
            if (sym_referenced(q)) throw DeferAbort();
            double r[4];
            st.reduced_qubit(q, r);
            const double tr = r[0] + r[1];
            const double purity = (r[0] * r[0] + r[1] * r[1] + 2.0 * (r[2] * r[2] + r[3] * r[3])) / (tr * tr);
            if (!(purity > 1.0 - 1e-9)) throw DeferAbort();  // still entangled: only a real measurement resets it
            // pure state (phi0, phi1) with rho01 = phi0 conj(phi1); fix the phase on the larger component
            cplx phi0, phi1;
            if (r[0] >= r[1]) {
                phi0 = std::sqrt(r[0] / tr);
                phi1 = std::conj(cplx(r[2], r[3])) / (tr * phi0);
            } else {
                phi1 = std::sqrt(r[1] / tr);
                phi0 = cplx(r[2], r[3]) / (tr * phi1);
            }
            const cplx u[4] = {std::conj(phi0), std::conj(phi1), -phi1, phi0};  // u (phi0, phi1)^T = (1, 0)^T
            st.apply_matrix(&q, 1, nullptr, 0, u, nullptr);
            qstate[q] = Q_LIVE;
            touched[q] = 0;
        };
        // reset whose outcome nobody reads (the `reset` instruction, the protocol's preparation of a communication pair)
        auto reset_q = [&](int q, const uint8_t *flags) {
            if (!deferred) { measure(q, true, flags); return; }
            if (flags) {
                if (flags[0] == 0) return;
                if (flags[0] != 1) throw DeferAbort();
            }
            if (qstate[q] == Q_FROZEN) { qstate[q] = Q_FROZEN_RESET; return; }  // record kept, logically |0> from here on
            if (qstate[q] == Q_FROZEN_RESET || !touched[q]) return;
            recycle(q);
        };
        // deferred mode: may this instruction (register qubits qs) still run?  false = skip it (a control sits on a
        // qubit that was measured and reset: logically |0>), DeferAbort = it would disturb a recorded outcome
        auto admit = [&](const Inst &in, const std::vector<int> &qs) -> bool {
            bool frozen = false;
            for (int q : qs) frozen = frozen || qstate[q] != Q_LIVE;
            if (frozen) {
                const std::vector<char> kinds = action_kinds(in);
                bool skip = false;
                for (size_t i = 0; i < qs.size(); i++) {
                    if (qstate[qs[i]] == Q_FROZEN && kinds[i] == 'x') throw DeferAbort();
                    if (qstate[qs[i]] == Q_FROZEN_RESET) {
                        if (kinds[i] == 'k') skip = true;  // controlled on |0>: never fires
                        else recycle(qs[i]);               // driven again: needs the physical |0> (or DeferAbort)
                    }
                }
                if (skip) return false;
            }
            for (int q : qs) touched[q] = 1;
            return true;
        };
        auto named = [&](const char *name, std::vector<int> qs, const Bits *flags) {
            if (deferred) {
                Inst in;
                in.name = name;
                in.qubits = qs;
                const uint8_t f = flags ? (*flags)[0] : 1;
                if (f == 0 || !admit(in, qs)) return;
                if (f == 1) st.apply_named(name, qs.data(), (int)qs.size(), nullptr, 0, nullptr);
                else queue_gate_controlled(st, in, qs, literals(f), false);
                return;
            }
            if (flags && !any(*flags)) return;
            if (!diverged && flags && !(*flags)[0]) return;
            st.apply_named(name, qs.data(), (int)qs.size(), nullptr, 0, gate_flags(flags));
        };
        auto gen_ent = [&](size_t npairs) {
            std::vector<int> idx;
            // (deferred mode: pairs are interchangeable, so never-used ones go first -- a used pair needs a real reset)
            if (deferred)
                for (size_t i = 0; i < pairs.size() && idx.size() < npairs; i++)
                    if (pairs[i].idle && !touched[pairs[i].q0] && !touched[pairs[i].q1]) idx.push_back((int)i);
            for (size_t i = 0; i < pairs.size() && idx.size() < npairs; i++)
                if (pairs[i].idle && std::find(idx.begin(), idx.end(), (int)i) == idx.end()) idx.push_back((int)i);
            if (idx.size() < npairs) return std::vector<int>();
            for (int i : idx) {
                pairs[i].idle = false;
                reset_q(pairs[i].q1, nullptr);
                reset_q(pairs[i].q0, nullptr);
                named("h", {pairs[i].q0}, nullptr);
                named("cx", {pairs[i].q0, pairs[i].q1}, nullptr);
            }
            return idx;
        };
        auto my_pairs = [&](const std::string &s, const std::string &r, const std::string &proto, size_t npairs) {
            std::vector<int> out;
            for (size_t i = 0; i < pairs.size(); i++) {
                if (npairs && out.size() == npairs) break;
                if (!pairs[i].idle && pairs[i].sendr == s && pairs[i].recvr == r && pairs[i].proto == proto) out.push_back((int)i);
            }
            return out;
        };

        // one instruction of task t (cond = per-state condition from enclosing cif, may be null)
        std::function<void(TaskRt &, const Inst &, const std::vector<int> &, const Bits *)> step =
            [&](TaskRt &t, const Inst &in, const std::vector<int> &comm_idx, const Bits *cond) {
                const TaskCfg &c = *t.cfg;
                auto qmap = [&](int q) {
                    if (q < 0) {
                        for (int i : comm_idx)
                            if (!pairs[i].idle && pairs[i].label == q) return pairs[i].q1;
                        throw std::runtime_error("unresolved remote control qubit");
                    }
                    if (q >= c.num_qubits) throw std::runtime_error("qubit index out of range in task '" + c.id + "'");
                    return q + t.zq;
                };
                auto clb = [&](int cl) -> Bits & {
                    if (cl < 0 || cl >= c.num_clbits) throw std::runtime_error("clbit index out of range in task '" + c.id + "'");
                    return creg[(size_t)(cl + t.zc)];
                };
                auto assign = [&](Bits &dst, const Bits &src) {
                    if (deferred && cond && (*cond)[0] >= kSym) throw DeferAbort();  // classically conditioned assignment
                    for (int b = 0; b < nb; b++) if (!cond || (*cond)[b]) dst[b] = src[b];
                };
                const std::string &name = in.name;
                if (name == "measure") {
                    assign(clb(in.clbits.at(0)), measure(qmap(in.qubits.at(0)), false, cond ? cond->data() : nullptr));
                } else if (name == "copy") {
                    if (in.l_clbits.size() != in.r_clbits.size())
                        throw std::runtime_error("The number of copied clbits and the number of clbits copied on does not match.");
                    for (size_t i = 0; i < in.l_clbits.size(); i++) { const Bits src = clb(in.r_clbits[i]); assign(clb(in.l_clbits[i]), src); }
                } else if (name == "reset") {
                    reset_q(qmap(in.qubits.at(0)), cond ? cond->data() : nullptr);
                } else if (name == "send") {
                    if (multi) {
                        for (int cl : in.clbits) q_cc[{c.id, in.qpus.at(0)}].push_back(clb(cl));
                    } else {
                        for (int cl : in.clbits) send_(user_, clb(cl)[0], in.qpus.at(0).c_str());
                    }
                } else if (name == "recv") {
                    if (multi) {
                        auto &qq = q_cc[{in.qpus.at(0), c.id}];
                        if (qq.size() >= in.clbits.size()) {
                            for (int cl : in.clbits) { assign(clb(cl), qq.front()); qq.pop_front(); }
                            t.b_cc = false;
                        } else {
                            t.b_cc = true;
                        }
                    } else {
                        st.sync();  // drain the stream before blocking on the network (cf. flush_ops, :574)
                        for (int cl : in.clbits) {
                            Bits v(nb, (uint8_t)(recv_(user_, in.qpus.at(0).c_str()) == 1));
                            assign(clb(cl), v);
                        }
                    }
                } else if (name == "cif") {
                    // fold the clbits with cif_ops (:582-600)
                    const bool cnd = in.condition != 0;
                    Bits res(nb, 0);
                    bool symbolic = false;
                    if (deferred) {
                        for (int cl : in.clbits) symbolic = symbolic || clb(cl)[0] >= kSym;
                        if (symbolic) {
                            // one recorded outcome tested for 1 or for 0, or the AND of several bits; an enclosing
                            // undecided condition joins the conjunction.  Anything else (or / xor folds) needs real bits.
                            std::vector<uint8_t> lits;
                            bool never = cond && (*cond)[0] == 0;
                            if (in.clbits.size() == 1) {
                                const uint8_t v = clb(in.clbits[0])[0];
                                lits.push_back(cnd ? v : (uint8_t)(kNotSym + sym_qubit(v)));
                            } else {
                                if (!(in.operation == "and" && cnd)) throw DeferAbort();
                                for (int cl : in.clbits) {
                                    const uint8_t v = clb(cl)[0];
                                    if (v == 0) never = true;
                                    else if (v >= kSym) lits.push_back(v);
                                }
                            }
                            if (cond && (*cond)[0] >= kSym)
                                for (uint8_t l : literals((*cond)[0])) lits.push_back(l);
                            res[0] = never ? 0 : make_cond(lits);
                        }
                    }
                    for (int b = 0; b < nb && !symbolic; b++) {
                        bool accb = cnd ? clb(in.clbits.at(0))[b] != 0 : clb(in.clbits.at(0))[b] == 0;
                        for (size_t k = 1; k < in.clbits.size(); k++) accb = cif_op(in.operation, accb, clb(in.clbits[k])[b] != 0);
                        const bool r = cnd ? accb : !accb;
                        res[b] = (uint8_t)((cnd == r) && (!cond || (*cond)[b]));
                        if (deferred && res[b] && cond && (*cond)[b] >= kSym) res[b] = (*cond)[b];  // decided here, undecided outside
                    }
                    if (any(res))
                        for (const Inst &sub : in.sub) step(t, sub, {}, &res);
                } else if (name == "qsend") {
                    const std::vector<int> idx = gen_ent(1);
                    if (idx.empty()) { t.b_td = true; return; }
                    t.b_td = false;
                    CommPair &p = pairs[idx[0]];
                    p.proto = "teledata";
                    const int q = qmap(in.qubits.at(0));
                    named("cx", {q, p.q0}, nullptr);
                    named("h", {q}, nullptr);
                    q_td[c.id].push_back(measure(q, true, nullptr));  // measure + "reset if 1" (:618-624)
                    q_td[c.id].push_back(measure(p.q0, false, nullptr));
                    peer(in.qpus.at(0)).b_td = false;
                    p.sendr = c.id;
                    p.recvr = in.qpus.at(0);
                } else if (name == "qrecv") {
                    auto &qq = q_td[in.qpus.at(0)];
                    if (qq.empty()) { t.b_td = true; return; }
                    if (t.b_td) return;
                    const Bits m1 = qq.front(); qq.pop_front();
                    const Bits m2 = qq.front(); qq.pop_front();
                    const std::vector<int> mine = my_pairs(in.qpus.at(0), c.id, "teledata", 1);
                    if (mine.empty()) throw std::runtime_error("qrecv without an entangled pair");
                    CommPair &p = pairs[mine[0]];
                    // m1 = outcome of the sent qubit, m2 = outcome of the sender's comm qubit.  The reference code
                    // (:653-658) applies X <- m1, Z <- m2 and that is the default here (drop-in: same histograms);
                    // the documented protocol (docs/source/_static/teledata.png) is X <- m2, Z <- m1, selected with
                    // b200_teledata_corrections = "protocol".  The result JSON says which one ran.
                    const bool lit = cfgs[0].bugcompat_teledata;
                    used_teledata = true;
                    named("x", {p.q1}, lit ? &m1 : &m2);
                    named("z", {p.q1}, lit ? &m2 : &m1);
                    named("swap", {p.q1, qmap(in.qubits.at(0))}, nullptr);
                    p.idle = true;
                } else if (name == "expose") {
                    if (!t.cat) {
                        const std::vector<int> idx = gen_ent(in.qubits.size());
                        if (idx.empty()) { t.b_tg = true; return; }
                        int qid = 0;
                        for (int i : idx) {
                            CommPair &p = pairs[i];
                            p.proto = "telegate";
                            p.label = -(qid + 1);
                            named("cx", {qmap(in.qubits[qid]), p.q0}, nullptr);
                            q_tg[c.id].push_back(measure(p.q0, false, nullptr));
                            t.cat = true;
                            t.b_tg = true;
                            peer(in.qpus.at(0)).b_tg = false;
                            p.sendr = c.id;
                            p.recvr = in.qpus.at(0);
                            qid++;
                        }
                        return;
                    }
                    auto &qq = q_tg[in.qpus.at(0)];
                    for (size_t i = 0; i < in.qubits.size(); i++) {
                        if (qq.empty()) throw std::runtime_error("expose: missing telegate measurement");
                        const Bits m = qq.front(); qq.pop_front();
                        named("z", {qmap(in.qubits[i])}, &m);
                    }
                    t.cat = false;
                    for (int i : my_pairs(c.id, in.qpus.at(0), "telegate", in.qubits.size())) pairs[i].idle = true;
                } else if (name == "rcontrol") {
                    auto &qq = q_tg[in.qpus.at(0)];
                    if (qq.empty()) { t.b_tg = true; return; }
                    if (t.b_tg) return;
                    const std::vector<int> idx = my_pairs(in.qpus.at(0), c.id, "telegate", 0);
                    for (int i : idx) {
                        const Bits m = qq.front(); qq.pop_front();
                        named("x", {pairs[i].q1}, &m);
                    }
                    for (const Inst &sub : in.sub) step(t, sub, idx, cond);
                    for (int i : idx) {
                        named("h", {pairs[i].q1}, nullptr);
                        q_tg[c.id].push_back(measure(pairs[i].q1, false, nullptr));
                    }
                    peer(in.qpus.at(0)).b_tg = false;
                    t.b_tg = false;
                } else {
                    if (cond && !any(*cond)) return;
                    if (deferred) {
                        std::vector<int> qs;
                        for (int x : in.qubits) qs.push_back(qmap(x));
                        if (name == "barrier" || name == "id" || name == "save_state") return;
                        // active reset, "measure q -> c; if (c) x q": the qubit is |0> afterwards, its record stays
                        if (name == "x" && qs.size() == 1 && cond && (*cond)[0] == kSym + qs[0] && qstate[qs[0]] == Q_FROZEN) {
                            qstate[qs[0]] = Q_FROZEN_RESET;
                            return;
                        }
                        if (!admit(in, qs)) return;
                        if (cond && (*cond)[0] >= kSym) queue_gate_controlled(st, in, qs, literals((*cond)[0]), c.bugcompat_unitary_order);
                        else queue_gate(st, in, qmap, nullptr, c.bugcompat_unitary_order);
                        return;
                    }
                    if (!diverged && cond && !(*cond)[0]) return;
                    queue_gate(st, in, qmap, gate_flags(cond), c.bugcompat_unitary_order);
                }
            };

        // terminal sampling: once every unfinished task has nothing but `measure` left (and no condition is
        // pending), the m remaining measure+collapse rounds (3*S each) are replaced by ONE draw per state from
        // |a|^2 (1*S): same joint distribution, one RNG draw instead of m.  The oracle restates the same rule.
        auto try_terminal = [&]() -> bool {
            if (!c0.terminal_sampling || deferred) return false;
            std::vector<std::pair<int, int>> pend;  // (register qubit, global clbit) in scheduler order
            bool any_left = false;
            for (TaskRt &t : T) {
                if (t.finished) continue;
                if (t.b_td || t.b_tg || t.b_cc) return false;
                const std::vector<Inst> &ins = t.cfg->insts;
                for (size_t k = t.pc; k < ins.size(); k++) {
                    if (ins[k].name != "measure") return false;
                    const int q = ins[k].qubits.at(0), cl = ins[k].clbits.at(0);
                    if (q < 0 || q >= t.cfg->num_qubits || cl < 0 || cl >= t.cfg->num_clbits) return false;
                    pend.emplace_back(q + t.zq, cl + t.zc);
                    any_left = true;
                }
            }
            if (!any_left || pend.size() < 2) return false;
            diverge();
            for (int b = 0; b < nb; b++) ctr[b] = ((shot0 + b + 1) << 20) + ordinal[b];
            std::vector<uint64_t> idx(nb);
            st.sample_each(seed, ctr.data(), idx.data());
            std::vector<char> seen((size_t)n_total, 0);
            for (auto &qc : pend) {
                for (int b = 0; b < nb; b++) creg[(size_t)qc.second][b] = (uint8_t)((idx[b] >> qc.first) & 1ull);
                seen[qc.first] = 1;
            }
            for (int b = 0; b < nb; b++) ordinal[b]++;
            for (TaskRt &t : T) { t.pc = t.cfg->insts.size(); t.finished = true; }
            return true;
        };

        // round-robin scheduler (aer_simulator_adapter.cpp:755-778)
        bool ended = false;
        uint64_t guard = 0;
        while (!ended) {
            if (try_terminal()) break;
            ended = true;
            bool progressed = false;
            for (TaskRt &t : T) {
                if (t.finished) continue;
                const std::vector<Inst> &ins = t.cfg->insts;
                if (t.pc >= ins.size()) { t.finished = true; continue; }
                if (t.b_td || t.b_tg || t.b_cc) {
                    ended = false;
                    if (!t.b_cc) continue;
                    // a task blocked on a local RECV re-polls its queue (the reference would spin forever here)
                }
                const size_t pc0 = t.pc;
                const bool was_blocked = t.b_td || t.b_tg || t.b_cc;
                step(t, ins[t.pc], {}, nullptr);
                if (!(t.b_td || t.b_tg || t.b_cc)) t.pc++;
                if (t.pc != pc0 || !was_blocked) progressed = true;
                if (t.pc != ins.size()) ended = false;
                else t.finished = true;
            }
            if (!ended && !progressed && ++guard > 1000)
                throw std::runtime_error("dynamic circuit deadlock: every task is blocked on communication");
            if (progressed) guard = 0;
        }
        if (deferred && job) {
            // batched registers: the draw happens once for all registers (Backend::execute_many)
            job->symbols.resize((size_t)ncl_total);
            for (int cl = 0; cl < ncl_total; cl++) job->symbols[cl] = creg[cl][0];
            for (const TaskRt &t : T) job->tasks.push_back({t.cfg->id, t.zc, t.cfg->num_clbits});
            return;
        }
        if (deferred) {
            // every shot is one draw from the final state; a symbolic bit reads its qubit from the drawn index
            std::vector<uint64_t> samples(live);
            st.sample(0, live, seed, samples.data());
            std::vector<uint8_t> bits((size_t)ncl_total);
            for (uint64_t s = 0; s < live; s++) {
                for (int cl = 0; cl < ncl_total; cl++) {
                    const uint8_t v = creg[cl][0];
                    bits[cl] = v >= kSym ? (uint8_t)((samples[s] >> sym_qubit(v)) & 1ull) : v;
                }
                for (const TaskRt &t : T) counters[t.cfg->id][bits_to_string(bits, t.zc, t.cfg->num_clbits)]++;
            }
            return;
        }
        st.sync();
        for (const TaskRt &t : T)
            for (uint64_t b = 0; b < live; b++) {
                std::vector<uint8_t> bits((size_t)ncl_total);
                for (int cl = 0; cl < ncl_total; cl++) bits[cl] = creg[cl][b];
                counters[t.cfg->id][bits_to_string(bits, t.zc, t.cfg->num_clbits)]++;
            }
    };

    if (job) {
        // one register of a batched run: deferred measurement or nothing (DeferAbort & co. propagate to execute_many)
        if (external_cc || n_total + (int)kSym > (int)kNotSym) throw DeferAbort();
        State &sb = *job->st;
        State::bind_register_for_this_thread(job->reg);   // (execute_many runs the registers' interpreters on several threads)
        if (sb.n != n_total || sb.precision != precision) throw std::runtime_error("internal: register width mismatch");
        apply_fusion_cfg(sb, c0);
        const uint64_t g0 = sb.rq().gates;
        run_batch(sb, 1, 0, shots, true);
        job->shots = shots;
        job->seed = seed;
        job->gates = sb.rq().gates - g0;
        job->used_teledata = used_teledata;
        job->teledata_reference = c0.bugcompat_teledata;
        job->multi = multi;
        return Json::object();
    }
    // first choice: defer every measurement (one state, all shots from one final sample)
    bool done = false;
    if (shots > 0 && c0.defer_measurement && !external_cc && n_total + (int)kSym <= (int)kNotSym) {  // (symbol encoding: <= 62 qubits)
        State &s1 = get_state(n_total, precision, 1, shards(n_total, c0.shard_min_qubits));
        apply_fusion_cfg(s1, c0);
        const Stats b1 = s1.stats;
        const uint64_t g1 = s1.queue.gates;
        auto undo = [&]() {
            s1.reset();
            for (auto &kv : counters) kv.second.clear();
        };
        try {
            run_batch(s1, 1, 0, shots, true);
            done = true;
            stp = &s1;
            before = b1;
            gates_before = g1;
            deferred_used = true;
        } catch (const DeferAbort &) {
            undo();
        } catch (const Unsupported &) {
            // turning a classical condition into extra quantum controls can exceed what the queue accepts (a diagonal
            // wider than kMaxDiagK, ...): the per-shot interpreter runs the same task without those controls
            undo();
        } catch (const InvalidArg &) {
            undo();  // e.g. a condition literal that is also an operand of the conditioned gate
        }
    }
    if (!done) {
        stp = shots ? &get_state(n_total, precision, (int)B, shards(n_total, c0.shard_min_qubits)) : nullptr;
        if (stp) { apply_fusion_cfg(*stp, c0); before = stp->stats; gates_before = stp->queue.gates; }
        for (uint64_t shot0 = 0; shot0 < shots; shot0 += B) {
            if (shot0 > 0) stp->reset();
            run_batch(*stp, (int)B, shot0, std::min<uint64_t>(B, shots - shot0), false);
        }
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    Json idc = Json::object();
    for (auto &kv : counters) idc.set(kv.first, counts_json(kv.second));
    Json out = Json::object();
    out.set("id_counts", std::move(idc));
    out.set("time_taken", Json::number(std::chrono::duration<double>(t1 - t0).count()));
    {
        Json info = Json::object();
        info.set("mode", Json::string(deferred_used ? "deferred_measurement" : "per_shot_batched"));
        if (used_teledata) info.set("teledata_corrections", Json::string(c0.bugcompat_teledata ? "reference" : "protocol"));
        out.set("b200", std::move(info));
    }
    last_mode = deferred_used ? 1 : 2;
    if (stp) {
        last_stats = stp->stats;
        last_stats.passes -= before.passes;
        last_stats.launches -= before.launches;
        last_stats.bytes -= before.bytes;
        last_stats.fused_ops -= before.fused_ops;
        last_stats.h2d_bytes -= before.h2d_bytes;
        last_stats.d2h_bytes -= before.d2h_bytes;
        last_stats.gates = stp->queue.gates - gates_before;
    }
    return out;
}

// ---- many independent messages as separate registers of one batched State ------------------------------
namespace {
// joint-register width of one message without parsing its instructions (allocation needs it before the run)
int message_width(const Json &msg)
{
    auto width1 = [](const Json &t) { return (int)t.at("config").at("num_qubits").as_int(); };
    if (!msg.is_array()) return width1(msg);
    if (msg.arr.empty()) throw std::runtime_error("empty task list");
    int n = 0;
    for (const Json &t : msg.arr) n += width1(t);
    if (msg.arr.size() > 1) {
        const Json &cfg = msg.arr[0].at("config");
        int nc = cfg.contains("n_communication_qubits") ? (int)cfg.at("n_communication_qubits").as_int() : 2;
        if (nc % 2) nc++;
        n += nc;
    }
    return n;
}
}  // namespace

namespace {
// f(i) for i in [0, count) on up to `max_threads` host threads (exceptions must be handled inside f)
template <typename F>
void parallel_for(size_t count, F f, unsigned max_threads = 16)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>({count, (size_t)hw, (size_t)max_threads});
    if (nt <= 1) { for (size_t i = 0; i < count; i++) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&] { for (size_t i = next++; i < count; i = next++) f(i); });
    for (std::thread &t : th) t.join();
}
}  // namespace

std::vector<std::string> Backend::execute_many(const std::vector<std::string> &messages)
{
    const auto t0 = std::chrono::high_resolution_clock::now();
    const size_t M = messages.size();
    std::vector<std::string> results(M);
    std::vector<int> width(M, -1);
    auto error_json = [](const std::string &what) {
        Json err = Json::object();
        err.set("ERROR", Json::string(what));
        return err.dump();
    };
    if (cached_many_.size() < M) cached_many_.resize(M);
    // 1. parse (in parallel); {"params"} updates rewrite the slot's cached message (src/quantum_task.cpp:33-35,39-108)
    std::vector<const Json *> mp(M, nullptr);
    parallel_for(M, [&](size_t i) {
        try {
            Json m = Json::parse(messages[i]);
            if (m.is_object() && !m.contains("instructions") && m.contains("params")) {
                if (!cached_many_[i].is_object() || !cached_many_[i].contains("instructions"))
                    throw std::runtime_error("Circuit not sent before updating parameters.");
                update_params(cached_many_[i], m);
            } else {
                cached_many_[i] = std::move(m);
            }
            const Json &c = cached_many_[i];
            if (c.is_object() && !(c.contains("instructions") && c.contains("config")))
                throw std::runtime_error("message has neither instructions+config nor params");
            if (c.is_object() && c.at("config").contains("precision") && c.at("config").at("precision").as_string() == "single")
                width[i] = -2;  // batched registers are complex128 (the dynamic path's precision, :946): single runs on its own
            else
                width[i] = message_width(c);
            mp[i] = &c;
        } catch (const std::exception &e) {
            results[i] = error_json(e.what());
            width[i] = -1;
        }
    });
    // structure signature of a message: instruction names and operands, no angles
    std::vector<std::string> signature(M);
    for (size_t i = 0; i < M; i++) {
        if (width[i] < 1) continue;
        std::function<void(const Json &, std::string &)> sig = [&](const Json &insts, std::string &out) {
            for (const Json &in : insts.arr) {
                out += in.at("name").as_string();
                for (const char *key : {"qubits", "clbits", "qpus"})
                    if (const Json *v = in.find(key)) out += v->dump();
                if (const Json *sub = in.find("instructions")) { out += '{'; sig(*sub, out); out += '}'; }
                out += ';';
            }
        };
        try {
            if (mp[i]->is_array()) for (const Json &t : mp[i]->arr) { sig(t.at("instructions"), signature[i]); signature[i] += '|'; }
            else sig(mp[i]->at("instructions"), signature[i]);
        } catch (const std::exception &) { signature[i] = "?"; }
    }
    last_uniform_flushes = last_split_flushes = last_solo = 0;
    Stats total;
    const bool trace = getenv("CB200_TRACE") != nullptr;
    auto now = [] { return std::chrono::high_resolution_clock::now(); };
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    const auto t_parsed = now();
    double us_queue = 0, us_flush = 0, us_sample = 0, us_decode = 0;
    // 2. registers of equal width share one State (as many per chunk as fit the memory budget)
    std::vector<char> handled(M, 0);
    for (size_t i = 0; i < M; i++) {
        if (handled[i] || width[i] < 1) continue;
        std::vector<size_t> members;
        for (size_t j = i; j < M; j++)
            if (!handled[j] && width[j] == width[i]) members.push_back(j);
        // equal circuits become neighbouring registers (State::flush_registers_t batches runs of equal structure)
        std::stable_sort(members.begin(), members.end(), [&](size_t a, size_t b) { return signature[a] < signature[b]; });
        const int n = width[i];
        const double budget = 96.0 * 1073741824.0;
        const size_t per_chunk = (size_t)std::max(1.0, std::min(4096.0, std::floor(budget / (16.0 * std::ldexp(1.0, n)))));
        for (size_t c0 = 0; c0 < members.size(); c0 += per_chunk) {
            const size_t G = std::min(per_chunk, members.size() - c0);
            std::vector<RegJob> jobs(G);
            std::vector<char> ok(G, 0);
            try {
                if (!batch_state_ || batch_state_->n != n || batch_state_->batch != (int)G) {
                    batch_state_.reset();
                    batch_state_.reset(new State(n, 64, device_, (int)G));
                }
                State &st = *batch_state_;
                st.use_registers(true);
                st.reset();
                const Stats before = st.stats;
                const uint64_t uf0 = st.uniform_flushes, sf0 = st.split_flushes;
                const auto tq0 = now();
                std::atomic<bool> cuda_failed{false};
                parallel_for(G, [&](size_t g) {
                    const Json &m = *mp[members[c0 + g]];
                    jobs[g].st = &st;
                    jobs[g].reg = (int)g;
                    try {
                        run_dynamic(m.is_array() ? std::vector<Json>(m.arr.begin(), m.arr.end()) : std::vector<Json>{m}, &jobs[g]);
                        ok[g] = 1;
                    } catch (const DeferAbort &) {
                    } catch (const CudaError &) {
                        cuda_failed = true;
                    } catch (const std::exception &) {
                        // Unsupported / InvalidArg from the extra controls of a deferred condition, or a malformed task:
                        // the ordinary path below gives this message its own answer (or its own error)
                    }
                    if (!ok[g]) {   // needs a real collapse / an external channel: leave the register idle, run it alone below
                        State::bind_register_for_this_thread((int)g);
                        st.rq().clear_pending();
                    }
                    State::bind_register_for_this_thread(-1);
                });
                if (cuda_failed) throw CudaError("CUDA failure while queueing a register");
                uint64_t max_shots = 0;
                std::vector<uint64_t> seeds(G, 0);
                for (size_t g = 0; g < G; g++) if (ok[g]) { max_shots = std::max(max_shots, jobs[g].shots); seeds[g] = jobs[g].seed; }
                std::vector<uint64_t> samples((size_t)G * max_shots);
                const auto tq1 = now();
                st.flush();
                const auto tq2 = now();
                st.sample_registers(max_shots, seeds.data(), samples.data());
                const auto tq3 = now();
                us_queue += us(tq0, tq1); us_flush += us(tq1, tq2); us_sample += us(tq2, tq3);
                const double secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
                last_uniform_flushes += st.uniform_flushes - uf0;
                last_split_flushes += st.split_flushes - sf0;
                total.passes += st.stats.passes - before.passes;
                total.launches += st.stats.launches - before.launches;
                total.bytes += st.stats.bytes - before.bytes;
                total.fused_ops += st.stats.fused_ops - before.fused_ops;
                total.h2d_bytes += st.stats.h2d_bytes - before.h2d_bytes;
                total.d2h_bytes += st.stats.d2h_bytes - before.d2h_bytes;
                for (size_t g = 0; g < G; g++) if (ok[g]) total.gates += jobs[g].gates;
                parallel_for(G, [&](size_t g) {
                    if (!ok[g]) return;
                    const RegJob &jb = jobs[g];
                    // every shot is one draw from the register's final state; a symbolic bit reads its qubit from the index.
                    // Per task: the classical register of every shot as an integer, sorted -> one string per distinct outcome
                    std::vector<Json> per_task;
                    for (const RegJob::Layout &t : jb.tasks) {
                        Json cj = Json::object();
                        if (t.ncl <= 64) {
                            uint64_t konst = 0;
                            std::vector<std::pair<int, int>> src;  // (register qubit, clbit of this task)
                            for (int c = 0; c < t.ncl; c++) {
                                const uint8_t v = jb.symbols[(size_t)(t.zc + c)];
                                if (v >= kSym) src.emplace_back(sym_qubit(v), c);
                                else if (v) konst |= 1ull << c;
                            }
                            std::vector<uint64_t> keys(jb.shots);
                            for (uint64_t s = 0; s < jb.shots; s++) {
                                const uint64_t idx = samples[g * max_shots + s];
                                uint64_t k = konst;
                                for (auto &qc : src) k |= ((idx >> qc.first) & 1ull) << qc.second;
                                keys[s] = k;
                            }
                            std::sort(keys.begin(), keys.end());
                            std::string key((size_t)t.ncl, '0');
                            for (size_t a = 0; a < keys.size();) {
                                size_t b = a;
                                while (b < keys.size() && keys[b] == keys[a]) b++;
                                for (int c = 0; c < t.ncl; c++) key[t.ncl - 1 - c] = ((keys[a] >> c) & 1ull) ? '1' : '0';
                                cj.obj.emplace_back(key, Json::integer((long long)(b - a)));
                                a = b;
                            }
                        } else {
                            std::map<std::string, uint64_t> counts;
                            std::vector<uint8_t> bits(jb.symbols.size());
                            for (uint64_t s = 0; s < jb.shots; s++) {
                                const uint64_t idx = samples[g * max_shots + s];
                                for (size_t cl = 0; cl < jb.symbols.size(); cl++) {
                                    const uint8_t v = jb.symbols[cl];
                                    bits[cl] = v >= kSym ? (uint8_t)((idx >> sym_qubit(v)) & 1ull) : v;
                                }
                                counts[bits_to_string(bits, t.zc, t.ncl)]++;
                            }
                            cj = counts_json(counts);
                        }
                        per_task.push_back(std::move(cj));
                    }
                    Json out = Json::object();
                    const Json &m = *mp[members[c0 + g]];
                    if (m.is_array()) {
                        // (id_counts is keyed by task id; ids are unique within a message)
                        std::vector<size_t> order(jb.tasks.size());
                        for (size_t k = 0; k < order.size(); k++) order[k] = k;
                        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return jb.tasks[a].id < jb.tasks[b].id; });
                        Json idc = Json::object();
                        for (size_t k : order) idc.set(jb.tasks[k].id, std::move(per_task[k]));
                        out.set("id_counts", std::move(idc));
                    } else {
                        out.set("counts", std::move(per_task[0]));
                    }
                    out.set("time_taken", Json::number(secs));
                    Json info = Json::object();
                    info.set("mode", Json::string("deferred_measurement"));
                    info.set("batched_registers", Json::integer((long long)G));
                    if (jb.used_teledata) info.set("teledata_corrections", Json::string(jb.teledata_reference ? "reference" : "protocol"));
                    out.set("b200", std::move(info));
                    results[members[c0 + g]] = out.dump();
                    handled[members[c0 + g]] = 1;
                });
                us_decode += us(tq3, now());
            } catch (const std::exception &e) {
                for (size_t g = 0; g < G; g++)
                    if (!handled[members[c0 + g]]) { results[members[c0 + g]] = error_json(e.what()); handled[members[c0 + g]] = 1; }
            }
            for (size_t g = 0; g < G; g++)
                if (!handled[members[c0 + g]]) width[members[c0 + g]] = -2;  // -> solo
        }
    }
    // 3. whatever could not ride in a register runs through the ordinary path, one message at a time
    for (size_t i = 0; i < M; i++) {
        if (handled[i] || width[i] != -2) continue;
        const Json keep_cached = cached_;
        const bool keep_has = has_cached_;
        results[i] = execute(mp[i]->dump());
        cached_ = keep_cached;
        has_cached_ = keep_has;
        parsed_.reset();
        last_solo++;
        total.passes += last_stats.passes; total.launches += last_stats.launches; total.bytes += last_stats.bytes;
        total.fused_ops += last_stats.fused_ops; total.h2d_bytes += last_stats.h2d_bytes; total.d2h_bytes += last_stats.d2h_bytes;
        total.gates += last_stats.gates;
    }
    last_stats = total;
    last_mode = 1;
    if (trace)
        fprintf(stderr, "[cb200] execute_many(%zu): parse %.0f us, interpret+queue+fuse %.0f us, plan+encode+launch %.0f us, wait+sample %.0f us, "
                        "decode+counts %.0f us, total %.0f us\n", M, us(t0, t_parsed), us_queue, us_flush, us_sample, us_decode, us(t0, now()));
    return results;
}

std::string update_task_json(const std::string &task_json, const std::string &update_json)
{
    Json task = Json::parse(task_json);
    update_params(task, Json::parse(update_json));
    return task.dump();
}

// ---- planning without a device -------------------------------------------------------------------
std::string plan_task_json(const std::string &task_json, const std::string &options_json)
{
    const Json task = Json::parse(task_json);
    const TaskCfg c = parse_task(task);
    PlanOptions opt;
    int world = 1, rank = 0;
    opt.tile_bits = c.precision == 64 ? 11 : 13;
    opt.min_run_bits = c.precision == 64 ? 4 : 5;
    if (!options_json.empty()) {
        const Json o = Json::parse(options_json);
        if (o.contains("tile_bits")) opt.tile_bits = (int)o.at("tile_bits").as_int();
        if (o.contains("min_run_bits")) opt.min_run_bits = (int)o.at("min_run_bits").as_int();
        if (o.contains("max_diag_qubits")) opt.max_diag_qubits = (int)o.at("max_diag_qubits").as_int();
        if (o.contains("max_pass_cost")) opt.max_pass_cost = (int)o.at("max_pass_cost").as_int();
        if (o.contains("lookahead")) opt.lookahead = (int)o.at("lookahead").as_int();
        if (o.contains("lookahead_min_qubits")) opt.lookahead_min_qubits = (int)o.at("lookahead_min_qubits").as_int();
        if (o.contains("fuse")) opt.fuse = o.at("fuse").as_bool();
        if (o.contains("fan")) opt.fan = o.at("fan").as_bool();
        if (o.contains("lin_perm")) opt.lin_perm = o.at("lin_perm").as_bool();
        if (o.contains("pair_ops")) opt.pair_ops = o.at("pair_ops").as_bool();
        if (o.contains("world")) world = (int)o.at("world").as_int();
        if (o.contains("rank")) rank = (int)o.at("rank").as_int();
    }
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    if ((1 << gbits) != world || gbits >= c.num_qubits) throw std::runtime_error("world must be a power of two smaller than 2^num_qubits");
    const int n_local = c.num_qubits - gbits;
    if (gbits) opt.global_mask = ((1ull << gbits) - 1ull) << n_local;
    OpQueue q(c.num_qubits, 1, opt);
    for (const Inst &in : c.insts) {
        if (is_classical_or_remote(in.name)) continue;
        queue_gate(q, in, [](int x) { return x; }, nullptr, false);
    }
    auto cplx_arr = [](const std::vector<cplx> &v) {
        Json a = Json::array();
        for (const cplx &z : v) {
            Json e = Json::array();
            e.arr.push_back(Json::number(z.real()));
            e.arr.push_back(Json::number(z.imag()));
            a.arr.push_back(std::move(e));
        }
        return a;
    };
    auto int_arr = [](const std::vector<int> &v) {
        Json a = Json::array();
        for (int x : v) a.arr.push_back(Json::integer(x));
        return a;
    };
    static const char *kinds[] = {"dense", "diag", "permx", "swap"};
    auto ops_json = [&](const std::vector<HostOp> &ops) {
        Json jops = Json::array();
        for (const HostOp &op : ops) {
            Json o = Json::object();
            o.set("kind", Json::string(kinds[(int)op.kind]));
            o.set("q", int_arr(op.q));
            o.set("ctrl", int_arr(op.ctrl));
            o.set("m", cplx_arr(op.m));
            o.set("n_orig", Json::integer(op.n_orig));
            jops.arr.push_back(std::move(o));
        }
        return jops;
    };
    auto passes_json = [&](const std::vector<HostOp> &ops, int nq) {
        Json jp = Json::array();
        for (const Pass &p : schedule_passes(ops, nq, opt)) {
            Json o = Json::object();
            o.set("tile_q", int_arr(p.tile_q));
            o.set("L", Json::integer(p.L));
            o.set("ops", int_arr(p.op_idx));
            // encode to make sure the pass is launchable, and report the kernel-parameter footprint
            PassParams<double> pp;
            std::vector<cplx> diag;
            std::string err;
            if (!encode_pass<double>(p, ops, nq, 1, pp, diag, err, opt.enc_flags())) throw std::runtime_error(err);
            o.set("diag_entries", Json::integer((long long)diag.size()));
            o.set("device_ops", Json::integer(pp.n_ops));
            {   // one letter per device op: D dense, T diagonal table, X / S permutations, P register-block pair, F phase fan, L linear permutation (a run of X / CX / SWAP);
                // lower case = applied in the same sweep as the op before it (a diagonal run folded into the next dense gate)
                std::string kinds;
                for (int q = 0; q < pp.n_ops; q++) {
                    static const char letter[] = {'D', 'T', 'X', 'S', 'P', 'F', 'L'};
                    kinds += letter[pp.ops[q].kind];
                    if (pp.ops[q].kind == OP_DIAG || pp.ops[q].kind == OP_LINPERM) kinds += std::to_string((int)pp.ops[q].k);   // table width / gates in the run
                    if (pp.ops[q].kind == OP_PAIR && !(pp.ops[q].k == 2 && pp.ops[q].k2 == 2 && pp.ops[q].s == 2))   // register block other than the common 4-bit one
                        kinds += "(" + std::to_string((int)pp.ops[q].k) + std::to_string((int)pp.ops[q].k2) + std::to_string((int)pp.ops[q].s) + ")";
                }
                o.set("device_kinds", Json::string(kinds));
                Json pb = Json::array();   // tile bits of every register block (ascending)
                for (int q = 0; q < pp.n_ops; q++) {
                    if (pp.ops[q].kind != OP_PAIR) continue;
                    Json e = Json::array();
                    std::vector<int> bits(pp.ops[q].tbit, pp.ops[q].tbit + pp.ops[q].nfix);
                    std::sort(bits.begin(), bits.end());
                    for (int b : bits) e.arr.push_back(Json::integer(b));
                    pb.arr.push_back(std::move(e));
                }
                o.set("pair_bits", std::move(pb));
                Json pd = Json::array();   // ... and its descriptor: local bits in block order, group-index segment masks, swizzled offsets
                for (int q = 0; q < pp.n_ops; q++) {
                    if (pp.ops[q].kind != OP_PAIR) continue;
                    Json e = Json::object(), tb = Json::array(), sg = Json::array(), so = Json::array();
                    for (int m = 0; m < pp.ops[q].nfix; m++) { tb.arr.push_back(Json::integer(pp.ops[q].tbit[m])); so.arr.push_back(Json::integer(pp.ops[q].sboff[m])); }
                    for (int m = 0; m <= pp.ops[q].nfix; m++) sg.arr.push_back(Json::integer(pp.ops[q].seg[m]));
                    e.set("tbit", std::move(tb));
                    e.set("seg", std::move(sg));
                    e.set("sboff", std::move(so));
                    pd.arr.push_back(std::move(e));
                }
                o.set("pair_desc", std::move(pd));
                Json lp = Json::array();   // every linear permutation (OP_LINPERM): its swizzled index-map columns, [15] = constant
                for (int q = 0; q < pp.n_ops; q++) {
                    if (pp.ops[q].kind != OP_LINPERM) continue;
                    Json e = Json::array();
                    for (int i = 0; i < 16; i++) e.arr.push_back(Json::integer(pp.ops[q].lin[i]));
                    lp.arr.push_back(std::move(e));
                }
                o.set("lin_perms", std::move(lp));
                Json ts = Json::array();   // tile number -> (tile base >> L) segment masks (PassParams::tile_seg)
                for (int k = 0; k <= pp.t - pp.L; k++) ts.arr.push_back(Json::integer((long long)pp.tile_seg[k]));
                o.set("tile_seg", std::move(ts));
                Json sw = Json::array();
                for (int q = 0; q < pp.n_sweeps; q++) {
                    Json e = Json::array();
                    e.arr.push_back(Json::integer(pp.sweeps[q].first));
                    e.arr.push_back(Json::integer(pp.sweeps[q].len));
                    sw.arr.push_back(std::move(e));
                }
                o.set("sweeps", std::move(sw));   // [first device op, number of ops] of every register sweep
            }
            o.set("diag_tables", Json::integer(pp.n_diag));
            if (pp.fan_op >= 0) {
                Json f = Json::object();
                f.set("centers", Json::integer(pp.ops[pp.fan_op].k));
                f.set("leaf_bits_in_tile", Json::integer(pp.ops[pp.fan_op].nfix));
                f.set("leaf_chunks_outside", Json::integer(pp.fan_chunks));
                o.set("fan", std::move(f));
            }
            jp.arr.push_back(std::move(o));
        }
        return jp;
    };
    Json out = Json::object();
    out.set("num_qubits", Json::integer(c.num_qubits));
    out.set("gates", Json::integer((long long)q.gates));
    out.set("relabelled_swaps", Json::integer((long long)q.relabelled_swaps));
    out.set("basis", Json::integer((long long)q.basis));  // physical index of the initial basis state (folded X-prep)
    if (world > 1) {
        // sharded state: this rank's segments (local pass groups and global<->local qubit swaps)
        const DistPlan dp = plan_distributed(q.ops(), c.num_qubits, n_local, rank);
        std::vector<int> l2p(q.l2p);
        for (int &p : l2p) p = dp.sigma[p];
        out.set("l2p", int_arr(l2p));
        out.set("n_local", Json::integer(n_local));
        Json segs = Json::array();
        for (const DistSegment &s : dp.segments) {
            Json o = Json::object();
            if (s.is_swap) {
                Json sw = Json::array();
                sw.arr.push_back(Json::integer(s.gbit));
                sw.arr.push_back(Json::integer(s.lbit));
                o.set("swap", std::move(sw));
                o.set("batch", Json::integer(s.batch));
            } else {
                o.set("ops", ops_json(s.ops));
                o.set("passes", passes_json(s.ops, n_local));
            }
            segs.arr.push_back(std::move(o));
        }
        out.set("segments", std::move(segs));
        return out.dump();
    }
    out.set("l2p", int_arr(q.l2p));
    out.set("ops", ops_json(q.ops()));
    Json jp = passes_json(q.ops(), c.num_qubits);
    out.set("passes", std::move(jp));
    return out.dump();
}

}  // namespace cb200
