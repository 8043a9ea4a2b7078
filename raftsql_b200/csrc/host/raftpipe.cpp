// raftpipe.cpp — newRaftNode (reference raft.go:62-78), startRaft (raft.go:144-188) and serveChannels
// (raft.go:204-246) around a HostNode; see raftpipe.hpp for the seam's protocol.
#include "raftpipe.hpp"

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>

#include "../../../include/mrq.h"

namespace raftsql {

// ---- the product core: libmrq.so through its C-ABI ------------------------------------------------------
namespace {
class EngineCore : public Core {
 public:
  EngineCore(uint32_t npeers, uint32_t id, int device) : R_(npeers) {
    mrq_config cfg;
    mrq_config_default(&cfg);  // ElectionTick 10, HeartbeatTick 1 (raft.go:154-155)
    cfg.n_groups = 1;
    cfg.n_replicas = npeers;
    cfg.self_id = id;  // raft.Config.ID (raft.go:153)
    cfg.seed = 0x5EED + id;
    cfg.device = device;
    cfg.inbox_slots = 1;
    if (mrq_create(&cfg, &e_) != MRQ_OK) throw std::runtime_error(std::string("mrq_create: ") + mrq_last_error(nullptr));
  }
  ~EngineCore() override { mrq_destroy(e_); }

  void import_hardstate(uint64_t term, uint64_t vote, uint64_t committed, uint64_t last_index, uint64_t last_term) override {
    mrq_state st{};
    st.term = &term;
    st.vote = &vote;
    st.committed = &committed;
    st.last_index = &last_index;
    st.last_term = &last_term;
    check(mrq_import_state(e_, &st));
  }

  CoreState tick(const std::vector<CoreMsg> &msgs, uint32_t nprop) override {
    std::vector<mrq_msg> m(msgs.size());
    for (size_t k = 0; k < msgs.size(); ++k) {
      m[k] = mrq_msg{};
      m[k].group = 0;
      m[k].from = (uint8_t)msgs[k].from;
      m[k].type = (uint8_t)msgs[k].type;
      m[k].term = msgs[k].term;
      m[k].index = msgs[k].index;
      m[k].logterm = msgs[k].logterm;
      m[k].commit = msgs[k].commit;
    }
    check(mrq_post_inbox_delta(e_, 0, m.empty() ? nullptr : m.data(), m.size(), 0));
    if (nprop) {
      const uint64_t g = 0;
      check(mrq_propose(e_, 0, &g, &nprop, 1));
    }
    check(mrq_tick(e_, 0));
    CoreState s;
    s.match.resize(R_);
    uint8_t role = 0, lead = 0;
    mrq_state st{};
    st.term = &s.term;
    st.vote = &s.vote;
    st.committed = &s.committed;
    st.last_index = &s.last_index;
    st.last_term = &s.last_term;
    st.match = s.match.data();
    st.role = &role;
    st.lead = &lead;
    check(mrq_export_state(e_, &st));
    check(mrq_sync_out(e_, &s.out));
    s.role = role;
    s.lead = lead;
    return s;
  }

 private:
  void check(int rc) {
    if (rc != MRQ_OK) throw std::runtime_error(std::string("mrq: ") + mrq_last_error(e_));
  }
  mrq_engine *e_ = nullptr;
  uint32_t R_;
};
}  // namespace

std::unique_ptr<Core> make_engine_core(uint32_t npeers, uint32_t id, int device) {
  return std::unique_ptr<Core>(new EngineCore(npeers, id, device));
}

// Nodes of one process that were given the same peer list share an in-process transport — the role loopback TCP
// + rafthttp play in the reference's in-process test cluster (raftsql_test.go:16-41).
std::shared_ptr<LocalTransport> transport_for(const std::vector<std::string> &peers) {
  static std::mutex mu;
  static std::map<std::string, std::shared_ptr<LocalTransport>> reg;
  std::string key;
  for (const auto &p : peers) key += p + ",";
  std::lock_guard<std::mutex> lk(mu);
  auto &t = reg[key];
  if (!t) t = std::make_shared<LocalTransport>();
  return t;
}

std::string RaftPipe::Close() {
  ProposeC->close();
  std::string err;
  const bool ok = ErrorC->recv(err);
  if (thread_.joinable()) thread_.join();
  return ok ? err : std::string();
}

RaftPipe::~RaftPipe() {
  if (thread_.joinable()) {
    ProposeC->close();
    ErrorC->close();  // nobody will read it any more: a node that died with an error must not block on reporting it
    thread_.join();
  }
}

std::unique_ptr<RaftPipe> NewRaftPipe(int id, const std::vector<std::string> &peers, std::shared_ptr<StrChan> proposeC,
                                      const RaftPipeOptions &opt) {
  std::unique_ptr<RaftPipe> rp(new RaftPipe());
  rp->ProposeC = proposeC;
  rp->CommitC = std::make_shared<CommitChan>();  // unbuffered (raft.go:65)
  rp->ErrorC = std::make_shared<StrChan>();      // unbuffered (raft.go:66)
  auto tr = opt.transport ? opt.transport : transport_for(peers);
  const uint32_t n = (uint32_t)peers.size();
  std::unique_ptr<Core> core = opt.core_factory ? opt.core_factory(n, (uint32_t)id) : make_engine_core(n, (uint32_t)id);
  const std::string waldir = opt.waldir == "auto" ? "raftsql-" + std::to_string(id) : opt.waldir;  // raft.go:69
  rp->node_ = std::make_shared<HostNode>(std::move(core), (uint32_t)id, n, tr, waldir);

  auto node = rp->node_;
  auto commitC = rp->CommitC;
  auto errorC = rp->ErrorC;
  const double tick_seconds = opt.tick_seconds;
  rp->thread_ = std::thread([node, commitC, errorC, proposeC, tick_seconds]() {
    std::atomic<bool> stop{false};
    std::mutex plock;
    std::string err;
    bool failed = false;
    std::thread feeder;
    auto publish = [&](const std::vector<std::string> &payloads) -> bool {  // raft.go:82-96
      for (const auto &d : payloads) {
        if (stop.load()) return false;
        if (!commitC->send(std::make_shared<std::string>(d), &stop)) return false;
      }
      return true;
    };
    try {
      const std::vector<std::string> replay = node->start();  // replayWAL (raft.go:122-134)
      if (publish(replay) && commitC->send(nullptr, &stop)) {  // nil: "commit channel is current" (raft.go:131-132)
        feeder = std::thread([&]() {  // raft.go:211-218: proposals -> raft; closing proposeC shuts the node down
          std::string prop;
          while (proposeC->recv(prop)) {
            std::lock_guard<std::mutex> lk(plock);
            node->propose(prop);
          }
          stop.store(true);
        });
        auto next = std::chrono::steady_clock::now();
        while (!stop.load()) {
          std::vector<std::string> out;
          {
            std::lock_guard<std::mutex> lk(plock);
            out = node->step_tick();
          }
          if (!publish(out)) break;
          next += std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(tick_seconds));
          const auto now = std::chrono::steady_clock::now();
          if (next > now) {
            while (!stop.load() && std::chrono::steady_clock::now() < next)
              std::this_thread::sleep_for(std::min<std::chrono::steady_clock::duration>(
                  next - std::chrono::steady_clock::now(), std::chrono::milliseconds(5)));
          } else {
            next = now;
          }
        }
      }
    } catch (const ChanClosed &) {
    } catch (const std::exception &ex) {  // writeError (raft.go:136-142)
      err = ex.what();
      failed = true;
    }
    stop.store(true);
    node->stop();
    commitC->close();
    if (failed) {
      try {
        errorC->send(err);
      } catch (const ChanClosed &) {
      }
    }
    errorC->close();
    if (feeder.joinable()) {
      proposeC->close();  // unblock the feeder if the node died on its own
      feeder.join();
    }
  });
  return rp;
}

// ---- the multi-group seam -----------------------------------------------------------------------------------------
namespace {
// One engine of G groups through the C-ABI: one sparse post, one mrq_propose, ONE mrq_tick, one export per tick.
class EngineMultiCore : public MultiCore {
 public:
  EngineMultiCore(uint32_t npeers, uint32_t id, size_t n_groups, int device) : G_(n_groups), R_(npeers) {
    mrq_config cfg;
    mrq_config_default(&cfg);  // ElectionTick 10, HeartbeatTick 1 (raft.go:154-155)
    cfg.n_groups = n_groups;
    cfg.n_replicas = npeers;
    cfg.self_id = id;  // this node's id in every group
    cfg.seed = 0x5EED + id;
    cfg.device = device;
    cfg.inbox_slots = 1;
    if (mrq_create(&cfg, &e_) != MRQ_OK) throw std::runtime_error(std::string("mrq_create: ") + mrq_last_error(nullptr));
  }
  ~EngineMultiCore() override { mrq_destroy(e_); }

  void import_hardstate(const std::vector<uint64_t> &term, const std::vector<uint64_t> &vote, const std::vector<uint64_t> &committed,
                        const std::vector<uint64_t> &last_index, const std::vector<uint64_t> &last_term) override {
    mrq_state st{};
    st.term = const_cast<uint64_t *>(term.data());
    st.vote = const_cast<uint64_t *>(vote.data());
    st.committed = const_cast<uint64_t *>(committed.data());
    st.last_index = const_cast<uint64_t *>(last_index.data());
    st.last_term = const_cast<uint64_t *>(last_term.data());
    check(mrq_import_state(e_, &st));
  }

  std::vector<CoreState> tick(const std::vector<std::vector<CoreMsg>> &msgs, const std::vector<uint32_t> &nprop) override {
    std::vector<mrq_msg> m;
    std::vector<uint64_t> pg;
    std::vector<uint32_t> pn;
    for (size_t g = 0; g < G_; ++g) {
      for (const CoreMsg &c : msgs[g]) {
        mrq_msg x{};
        x.group = g;
        x.from = (uint8_t)c.from;
        x.type = (uint8_t)c.type;
        x.term = c.term;
        x.index = c.index;
        x.logterm = c.logterm;
        x.commit = c.commit;
        m.push_back(x);
      }
      if (nprop[g]) {
        pg.push_back(g);
        pn.push_back(nprop[g]);
      }
    }
    check(mrq_post_inbox_delta(e_, 0, m.empty() ? nullptr : m.data(), m.size(), 0));
    if (!pg.empty()) check(mrq_propose(e_, 0, pg.data(), pn.data(), pg.size()));
    check(mrq_tick(e_, 0));
    std::vector<uint64_t> term(G_), vote(G_), committed(G_), li(G_), lt(G_), match(G_ * R_);
    std::vector<uint8_t> role(G_), lead(G_);
    std::vector<uint32_t> out(G_);
    mrq_state st{};
    st.term = term.data();
    st.vote = vote.data();
    st.committed = committed.data();
    st.last_index = li.data();
    st.last_term = lt.data();
    st.match = match.data();  // [R][G]
    st.role = role.data();
    st.lead = lead.data();
    check(mrq_export_state(e_, &st));
    check(mrq_sync_out(e_, out.data()));
    std::vector<CoreState> s(G_);
    for (size_t g = 0; g < G_; ++g) {
      s[g].term = term[g];
      s[g].vote = vote[g];
      s[g].committed = committed[g];
      s[g].last_index = li[g];
      s[g].last_term = lt[g];
      s[g].role = role[g];
      s[g].lead = lead[g];
      s[g].out = out[g];
      s[g].match.resize(R_);
      for (uint32_t r = 0; r < R_; ++r) s[g].match[r] = match[(size_t)r * G_ + g];
    }
    return s;
  }

 private:
  void check(int rc) {
    if (rc != MRQ_OK) throw std::runtime_error(std::string("mrq: ") + mrq_last_error(e_));
  }
  mrq_engine *e_ = nullptr;
  size_t G_;
  uint32_t R_;
};
}  // namespace

std::unique_ptr<MultiCore> make_engine_multicore(uint32_t npeers, uint32_t id, size_t n_groups, int device) {
  return std::unique_ptr<MultiCore>(new EngineMultiCore(npeers, id, n_groups, device));
}

std::string MultiRaftPipe::Close() {
  for (auto &c : ProposeC) c->close();
  std::string err;
  const bool ok = ErrorC->recv(err);
  if (thread_.joinable()) thread_.join();
  return ok ? err : std::string();
}

MultiRaftPipe::~MultiRaftPipe() {
  if (thread_.joinable()) {
    for (auto &c : ProposeC) c->close();
    ErrorC->close();  // nobody will read it any more
    thread_.join();
  }
}

std::unique_ptr<MultiRaftPipe> NewMultiRaftPipe(int id, const std::vector<std::string> &peers, size_t n_groups,
                                                const MultiRaftPipeOptions &opt) {
  std::unique_ptr<MultiRaftPipe> mp(new MultiRaftPipe());
  for (size_t g = 0; g < n_groups; ++g) {
    mp->ProposeC.push_back(std::make_shared<StrChan>());
    mp->CommitC.push_back(std::make_shared<CommitChan>());  // unbuffered (raft.go:65)
  }
  mp->ErrorC = std::make_shared<StrChan>();
  auto tr = opt.transport ? opt.transport : std::make_shared<MultiLocalTransport>(n_groups);
  const uint32_t n = (uint32_t)peers.size();
  std::unique_ptr<MultiCore> core =
      opt.core_factory ? opt.core_factory(n, (uint32_t)id, n_groups) : make_engine_multicore(n, (uint32_t)id, n_groups);
  const std::string waldir = opt.waldir == "auto" ? "raftsql-" + std::to_string(id) : opt.waldir;
  mp->node_ = std::make_shared<MultiHostNode>(std::move(core), (uint32_t)id, n, n_groups, tr, waldir);

  auto node = mp->node_;
  auto proposeC = mp->ProposeC;
  auto commitC = mp->CommitC;
  auto errorC = mp->ErrorC;
  const double tick_seconds = opt.tick_seconds;
  mp->thread_ = std::thread([node, proposeC, commitC, errorC, tick_seconds]() {
    std::atomic<bool> stop{false};
    std::string err;
    bool failed = false;
    auto publish = [&](size_t g, const std::vector<std::string> &payloads) -> bool {  // raft.go:82-96
      for (const auto &d : payloads) {
        if (stop.load()) return false;
        if (!commitC[g]->send(std::make_shared<std::string>(d), &stop)) return false;
      }
      return true;
    };
    // raft.go:211-218 for every group without a thread per group: take what is offered right now; the node shuts
    // down once every ProposeC has been closed
    auto feed = [&]() -> bool {
      bool open_any = false;
      for (size_t g = 0; g < proposeC.size(); ++g) {
        for (;;) {
          std::string p;
          const int r = proposeC[g]->try_recv(p);
          if (r == 0) {
            open_any = true;
            break;
          }
          if (r < 0) break;
          node->propose(g, p);
        }
      }
      return open_any;
    };
    try {
      const auto replay = node->start();  // replayWAL per group (raft.go:122-134)
      bool ok = true;
      for (size_t g = 0; ok && g < replay.size(); ++g)
        ok = publish(g, replay[g]) && commitC[g]->send(nullptr, &stop);  // nil: "commit channel is current"
      auto next = std::chrono::steady_clock::now();
      while (ok && !stop.load()) {
        if (!feed()) break;
        const auto out = node->step_tick();
        for (size_t g = 0; ok && g < out.size(); ++g) ok = publish(g, out[g]);
        next += std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(tick_seconds));
        const auto now = std::chrono::steady_clock::now();
        if (next > now)
          std::this_thread::sleep_for(next - now);
        else
          next = now;
      }
    } catch (const ChanClosed &) {
    } catch (const std::exception &ex) {  // writeError (raft.go:136-142)
      err = ex.what();
      failed = true;
    }
    stop.store(true);
    node->stop();
    for (auto &c : commitC) c->close();
    if (failed) {
      try {
        errorC->send(err);
      } catch (const ChanClosed &) {
      }
    }
    errorC->close();
  });
  return mp;
}

}  // namespace raftsql
