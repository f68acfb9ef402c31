// fake_channel.cpp -- in-process stand-in for the reference's ClassicalChannel.  TEST INFRASTRUCTURE ONLY.
//
// Implements the class declared in /root/reference/src/classical_channel/classical_channel.hpp (same members, same
// semantics as the ZMQ implementation, zmq_classical_channel.cpp:17-138: per-origin FIFO queues, blocking receive,
// measurements sent as decimal strings, endpoints published in $STORE/.cunqa/communications.json) with mutex-guarded
// queues instead of sockets, so that the B200 plugin classes can be linked and run end to end in one process: every
// "vQPU process" and the "executor process" of a quantum-communication family become threads.
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>

#include "classical_channel/classical_channel.hpp"
#include "utils/constants.hpp"
#include "utils/json.hpp"

namespace cunqa {
namespace comm {

namespace {
struct Mailboxes {
    std::mutex m;
    std::condition_variable cv;
    // inbox[receiver][sender] = FIFO of messages
    std::map<std::string, std::map<std::string, std::deque<std::string>>> inbox;
};
Mailboxes& boxes()
{
    static Mailboxes b;
    return b;
}
}  // namespace

struct ClassicalChannel::Impl {};

ClassicalChannel::ClassicalChannel(const std::string& id) : endpoint{id}, pimpl_{std::make_unique<Impl>()}, qpu_id{id} {}
ClassicalChannel::~ClassicalChannel() = default;

void ClassicalChannel::publish()
{
    // what zmq_classical_channel.cpp:109-113 does: one entry per endpoint in communications.json
    JSON entry = {{"communications_endpoint", "inproc://" + qpu_id}};
    write_on_file(entry, constants::COMM_FILEPATH, qpu_id);
}

void ClassicalChannel::connect(const std::string&) {}

void ClassicalChannel::send_info(const std::string& data, const std::string& target)
{
    Mailboxes& b = boxes();
    {
        std::lock_guard<std::mutex> lk(b.m);
        b.inbox[target][qpu_id].push_back(data);
    }
    b.cv.notify_all();
}

std::string ClassicalChannel::recv_info(const std::string& origin)
{
    Mailboxes& b = boxes();
    std::unique_lock<std::mutex> lk(b.m);
    b.cv.wait(lk, [&] { return !b.inbox[qpu_id][origin].empty(); });
    std::string out = std::move(b.inbox[qpu_id][origin].front());
    b.inbox[qpu_id][origin].pop_front();
    return out;
}

void ClassicalChannel::send_measure(const int& measurement, const std::string& target) { send_info(std::to_string(measurement), target); }
int ClassicalChannel::recv_measure(const std::string& origin) { return std::stoi(recv_info(origin)); }

}  // namespace comm
}  // namespace cunqa
