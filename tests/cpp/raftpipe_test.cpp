// raftpipe_test.cpp — scenario tests for the C++ host side (raftsql_b200/csrc/host/): the seam's channel
// protocol, a 3-node in-process cluster, and stop / restart with WAL replay — the C++ counterpart of
// tests/test_plumbing.py, which restates the reference's raftsql_test.go:92-171 at the raftPipe level.
//
//   raftpipe_test oracle <tmpdir>     consensus core = the CPU oracle (TEST ONLY: this file lives under tests/)
//   raftpipe_test engine <tmpdir>     consensus core = the GPU engine through the C-ABI (needs a B200)
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include "../../include/mrq.h"
#include "../../oracle/raft_oracle.h"
#include "../../raftsql_b200/csrc/host/raftpipe.hpp"

using namespace raftsql;

namespace {

// TEST-ONLY core: the oracle wearing the Core interface (one group, dense inbox of R slots).
class OracleCore : public Core {
 public:
  OracleCore(uint32_t npeers, uint32_t id) : R_(npeers) {
    e_ = orc_create(1, npeers, 0, 10, 1, 0x5EED + id, id);
    if (!e_) throw std::runtime_error("orc_create failed");
  }
  ~OracleCore() override { orc_destroy(e_); }
  void import_hardstate(uint64_t term, uint64_t vote, uint64_t committed, uint64_t last_index, uint64_t last_term) override {
    orc_import(e_, &term, &vote, &committed, &last_index, &last_term, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
               nullptr, nullptr, nullptr);
  }
  CoreState tick(const std::vector<CoreMsg> &msgs, uint32_t nprop) override {
    std::vector<uint8_t> type(R_, 0);
    std::vector<uint64_t> term(R_, 0), index(R_, 0), logterm(R_, 0), commit(R_, 0);
    for (const CoreMsg &m : msgs) {
      const uint32_t r = m.from - 1;
      type[r] = (uint8_t)m.type;
      term[r] = m.term;
      index[r] = m.index;
      logterm[r] = m.logterm;
      commit[r] = m.commit;
    }
    orc_tick(e_, type.data(), term.data(), index.data(), logterm.data(), commit.data(), &nprop, 1);
    CoreState s;
    s.match.resize(R_);
    uint8_t role = 0, lead = 0;
    orc_export(e_, &s.term, &s.vote, &s.committed, &s.last_index, &s.last_term, nullptr, s.match.data(), &role, &lead, nullptr,
               nullptr, nullptr, nullptr, nullptr, &s.out);
    s.role = role;
    s.lead = lead;
    return s;
  }

 private:
  orc_engine *e_;
  uint32_t R_;
};

// TEST-ONLY multi-group core: the oracle with G groups behind the MultiCore interface.
class OracleMultiCore : public MultiCore {
 public:
  OracleMultiCore(uint32_t npeers, uint32_t id, size_t n_groups) : G_(n_groups), R_(npeers) {
    e_ = orc_create(n_groups, npeers, 0, 10, 1, 0x5EED + id, id);
    if (!e_) throw std::runtime_error("orc_create failed");
  }
  ~OracleMultiCore() override { orc_destroy(e_); }
  void import_hardstate(const std::vector<uint64_t> &term, const std::vector<uint64_t> &vote, const std::vector<uint64_t> &committed,
                        const std::vector<uint64_t> &last_index, const std::vector<uint64_t> &last_term) override {
    orc_import(e_, term.data(), vote.data(), committed.data(), last_index.data(), last_term.data(), nullptr, nullptr, nullptr,
               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  std::vector<CoreState> tick(const std::vector<std::vector<CoreMsg>> &msgs, const std::vector<uint32_t> &nprop) override {
    std::vector<uint8_t> type(R_ * G_, 0);
    std::vector<uint64_t> term(R_ * G_, 0), index(R_ * G_, 0), logterm(R_ * G_, 0), commit(R_ * G_, 0);
    for (size_t g = 0; g < G_; ++g)
      for (const CoreMsg &m : msgs[g]) {
        const size_t o = (size_t)(m.from - 1) * G_ + g;
        type[o] = (uint8_t)m.type;
        term[o] = m.term;
        index[o] = m.index;
        logterm[o] = m.logterm;
        commit[o] = m.commit;
      }
    orc_tick(e_, type.data(), term.data(), index.data(), logterm.data(), commit.data(), nprop.data(), 1);
    std::vector<uint64_t> t(G_), v(G_), c(G_), li(G_), lt(G_), match(R_ * G_);
    std::vector<uint8_t> role(G_), lead(G_);
    std::vector<uint32_t> out(G_);
    orc_export(e_, t.data(), v.data(), c.data(), li.data(), lt.data(), nullptr, match.data(), role.data(), lead.data(), nullptr,
               nullptr, nullptr, nullptr, nullptr, out.data());
    std::vector<CoreState> s(G_);
    for (size_t g = 0; g < G_; ++g) {
      s[g].term = t[g];
      s[g].vote = v[g];
      s[g].committed = c[g];
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
  orc_engine *e_;
  size_t G_;
  uint32_t R_;
};

int failures = 0;
#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);     \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
      ++failures;                                          \
    }                                                      \
  } while (0)

struct Collector {  // drains a CommitC on its own thread, like db.go's readCommits
  std::shared_ptr<CommitChan> ch;
  std::thread th;
  std::mutex mu;
  std::vector<std::string> got;
  std::atomic<int> nils{0};
  explicit Collector(std::shared_ptr<CommitChan> c) : ch(std::move(c)) {
    th = std::thread([this]() {
      std::shared_ptr<std::string> v;
      while (ch->recv(v)) {
        if (!v) {
          ++nils;
          continue;
        }
        std::lock_guard<std::mutex> lk(mu);
        got.push_back(*v);
      }
    });
  }
  ~Collector() {
    if (th.joinable()) th.join();
  }
  size_t size() {
    std::lock_guard<std::mutex> lk(mu);
    return got.size();
  }
  std::vector<std::string> snapshot() {
    std::lock_guard<std::mutex> lk(mu);
    return got;
  }
  bool wait_for(size_t n, int ms = 30000) {
    for (int t = 0; t < ms / 5; ++t) {
      if (size() >= n) return true;
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    return false;
  }
};

RaftPipeOptions options(const std::string &core, const std::string &dir, int id, std::shared_ptr<LocalTransport> tr, double tick) {
  RaftPipeOptions o;
  o.tick_seconds = tick;
  o.waldir = dir + "/raftsql-" + std::to_string(id);
  o.transport = std::move(tr);
  if (core == "oracle")
    o.core_factory = [](uint32_t n, uint32_t i) { return std::unique_ptr<Core>(new OracleCore(n, i)); };
  return o;
}

void test_single_node(const std::string &core, const std::string &dir) {
  std::vector<std::string> peers = {"http://127.0.0.1:9021"};  // server/main.go:25 default
  auto proposeC = std::make_shared<StrChan>();
  auto rp = NewRaftPipe(1, peers, proposeC, options(core, dir + "/single", 1, std::make_shared<LocalTransport>(), 0.005));
  std::shared_ptr<std::string> first;
  CHECK(rp->CommitC->recv(first, 20000) && !first, "the first value on CommitC must be the nil sentinel");
  Collector col(rp->CommitC);
  for (int i = 0; i < 20; ++i) proposeC->send("entry-" + std::to_string(i));
  CHECK(col.wait_for(20), "20 proposals must commit (got %zu)", col.size());
  auto got = col.snapshot();
  for (int i = 0; i < 20 && i < (int)got.size(); ++i) CHECK(got[i] == "entry-" + std::to_string(i), "order broken at %d: %s", i, got[i].c_str());
  CHECK(rp->Close().empty(), "Close() must return no error");
  std::shared_ptr<std::string> v;
  CHECK(!rp->CommitC->recv(v), "CommitC must be closed after Close()");
}

void test_cluster_and_restart(const std::string &core, const std::string &dir) {
  const std::vector<std::string> peers = {"http://127.0.0.1:10000", "http://127.0.0.1:10001", "http://127.0.0.1:10002"};
  auto tr = std::make_shared<LocalTransport>();
  std::vector<std::shared_ptr<StrChan>> prop(3);
  std::vector<std::unique_ptr<RaftPipe>> rp(3);
  std::vector<std::unique_ptr<Collector>> col(3);
  auto start = [&](int i) {
    prop[i] = std::make_shared<StrChan>();
    rp[i] = NewRaftPipe(i + 1, peers, prop[i], options(core, dir + "/clus", i + 1, tr, 0.01));
    col[i].reset(new Collector(rp[i]->CommitC));
  };
  for (int i = 0; i < 3; ++i) start(i);
  // createEntries (raftsql_test.go:54-69): one entry through node 0, then one through every node
  prop[0]->send("CREATE");
  for (int i = 0; i < 3; ++i) CHECK(col[i]->wait_for(1), "node %d never saw CREATE", i);
  for (int i = 0; i < 3; ++i) prop[i]->send("INSERT-" + std::to_string(i));
  for (int i = 0; i < 3; ++i) CHECK(col[i]->wait_for(4), "node %d committed %zu of 4", i, col[i]->size());
  auto ref = col[0]->snapshot();
  for (int i = 1; i < 3; ++i) CHECK(col[i]->snapshot() == ref, "node %d applied a different sequence", i);
  CHECK(ref.size() == 4 && ref[0] == "CREATE", "unexpected log head");
  for (int i = 0; i < 3; ++i) CHECK(col[i]->nils.load() == 1, "node %d: %d nil sentinels", i, col[i]->nils.load());

  // TestRestartDB (raftsql_test.go:117-171): stop a FOLLOWER (a proposal forwarded to a dead leader is dropped by
  // raft, upstream too), add an entry through another node, restart the victim from its WAL
  int victim = rp[1]->node()->role() != MRQ_ROLE_LEADER ? 1 : 2;
  int proposer = 3 - victim;
  CHECK(rp[victim]->Close().empty(), "clean stop of node %d", victim);
  col[victim].reset();
  prop[proposer]->send("foo");
  CHECK(col[proposer]->wait_for(5), "foo must commit with 2 of 3 nodes");
  CHECK(col[0]->wait_for(5) || victim == 0, "node 0 must apply foo");
  start(victim);
  CHECK(col[victim]->wait_for(5), "restarted node must replay 4 entries and catch up to foo (has %zu)", col[victim]->size());
  auto again = col[victim]->snapshot();
  CHECK(again.size() == 5 && std::vector<std::string>(again.begin(), again.begin() + 4) == ref, "replay must be the original 4 entries in order");
  CHECK(again.size() == 5 && again[4] == "foo", "catch-up entry must be foo");
  CHECK(col[victim]->nils.load() == 1, "exactly one nil sentinel after the replay");
  for (int i = 0; i < 3; ++i) CHECK(rp[i]->Close().empty(), "Close node %d", i);
}

// Go channel semantics the seam relies on (raft.go:89-93 `select { case commitC <- v: case <-stopc: }`).
void test_chan_semantics() {
  {  // an aborted send withdraws its value: the receiver must never see it
    Chan<int> ch;
    std::atomic<bool> stop{false};
    std::atomic<int> result{-1};
    std::thread sender([&]() { result = ch.send(7, &stop) ? 1 : 0; });
    std::this_thread::sleep_for(std::chrono::milliseconds(40));
    stop = true;
    sender.join();
    CHECK(result.load() == 0, "a stopped send must report false");
    ch.close();
    int v = 0;
    CHECK(!ch.recv(v), "the withdrawn value must not be delivered");
  }
  {  // unbuffered rendezvous: send returns only after the value was taken; FIFO across blocked senders
    Chan<int> ch;
    std::atomic<int> returned{0};
    std::thread s1([&]() { ch.send(1); ++returned; });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    std::thread s2([&]() { ch.send(2); ++returned; });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    CHECK(returned.load() == 0, "unbuffered send returned before any receive");
    int a = 0, b = 0;
    CHECK(ch.recv(a, 2000) && ch.recv(b, 2000) && a == 1 && b == 2, "FIFO order across senders (%d, %d)", a, b);
    s1.join();
    s2.join();
    CHECK(returned.load() == 2, "both senders must be released");
  }
  {  // close() fails a blocked sender (Go panics) and later sends; buffered values stay drainable
    Chan<int> ch;
    std::atomic<bool> threw{false};
    std::thread s([&]() {
      try {
        ch.send(5);
      } catch (const ChanClosed &) {
        threw = true;
      }
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    ch.close();
    s.join();
    CHECK(threw.load(), "send blocked on a channel that gets closed must throw");
    bool again = false;
    try {
      ch.send(6);
    } catch (const ChanClosed &) {
      again = true;
    }
    CHECK(again, "send on a closed channel must throw");
    Chan<int> buf(2);
    buf.send(1);
    buf.send(2);
    buf.close();
    int v = 0;
    CHECK(buf.recv(v) && v == 1 && buf.recv(v) && v == 2 && !buf.recv(v), "buffered values drain after close");
  }
}

// wal.ReadAll tolerates a torn tail (a crash mid-append); a WAL that cannot be opened is fatal (raft.go:102-114).
void test_wal_recovery(const std::string &dir) {
  const std::string wdir = dir + "/walrec";
  {
    Wal w(wdir);
    CHECK(w.open(), "open a fresh wal");
    std::vector<Entry> es(3);
    for (int i = 0; i < 3; ++i) {
      es[i].term = 2;
      es[i].data = "payload-" + std::to_string(i);
    }
    const uint64_t hs[3] = {2, 1, 2};
    w.save(hs, es, 1, false, 0);
    w.close();
  }
  {  // a record header promising 3 GB, then nothing: must be ignored without allocating it
    FILE *f = std::fopen((wdir + "/wal.bin").c_str(), "ab");
    const char kind = 'E';
    const uint64_t a = 4, b = 2, c = 0;
    const uint32_t len = 0xC0000000u;
    const uint32_t grp = 0, crc = 0;
    std::fwrite(&kind, 1, 1, f);
    std::fwrite(&grp, 4, 1, f);
    std::fwrite(&a, 8, 1, f);
    std::fwrite(&b, 8, 1, f);
    std::fwrite(&c, 8, 1, f);
    std::fwrite(&len, 4, 1, f);
    std::fwrite(&crc, 4, 1, f);
    std::fwrite("xy", 1, 2, f);
    std::fclose(f);
  }
  Wal w(wdir);
  std::vector<Entry> ents;
  bool has_hs = false;
  uint64_t hs[3] = {0, 0, 0};
  w.read_all(&ents, &has_hs, hs);
  CHECK(ents.size() == 3 && ents[2].data == "payload-2" && ents[0].term == 2, "intact prefix must survive a torn tail (%zu entries)", ents.size());
  CHECK(has_hs && hs[0] == 2 && hs[1] == 1 && hs[2] == 2, "hardstate must survive a torn tail");
  {  // restart #1 appends AFTER the tear: open() must cut the garbage off first, or restart #2 loses these records
    Wal w1(wdir);
    CHECK(w1.open(), "reopen the torn wal");
    std::vector<Entry> more(1);
    more[0].term = 3;
    more[0].data = "after-the-crash";
    const uint64_t hs3[3] = {3, 2, 2};
    w1.save(hs3, more, 4, false, 0);
    w1.close();
    Wal w2(wdir);
    w2.read_all(&ents, &has_hs, hs);
    CHECK(ents.size() == 4 && ents[3].data == "after-the-crash" && ents[3].term == 3,
          "records saved after a torn tail must be replayed by the next restart (%zu entries)", ents.size());
    CHECK(has_hs && hs[0] == 3 && hs[1] == 2, "the vote saved after a torn tail must survive (term %llu)", (unsigned long long)hs[0]);
  }
  {  // a flipped payload bit: the CRC ends the valid prefix there
    const std::string path = wdir + "/wal.bin";
    FILE *f = std::fopen(path.c_str(), "rb+");
    std::fseek(f, -3, SEEK_END);  // inside the last record (the H record has no payload: this hits its crc/len area)
    int ch = std::fgetc(f);
    std::fseek(f, -3, SEEK_END);
    std::fputc(ch ^ 1, f);
    std::fclose(f);
    Wal w3(wdir);
    w3.read_all(&ents, &has_hs, hs);
    CHECK(has_hs && hs[0] == 2 && ents.size() == 4, "a damaged last record is dropped, the prefix stands (term %llu, %zu entries)",
          (unsigned long long)hs[0], ents.size());
  }

  // an unopenable wal directory (a regular file stands where the directory should be) fails start()
  const std::string blocker = dir + "/not-a-dir";
  FILE *f = std::fopen(blocker.c_str(), "wb");
  std::fclose(f);
  auto tr = std::make_shared<LocalTransport>();
  HostNode node(std::unique_ptr<Core>(new OracleCore(1, 1)), 1, 1, tr, blocker);
  bool threw = false;
  try {
    node.start();
  } catch (const std::runtime_error &) {
    threw = true;
  }
  CHECK(threw, "start() must fail when the wal cannot be opened");
}

// The multi-group seam (raftpipe.hpp NewMultiRaftPipe): 3 nodes x G groups in one process, one core per node ticked
// once per tick for all groups.  Every group replicates independently and in its own log order, nil once per group,
// a stopped node replays every group's WAL and catches up (tests/test_multipipe_cpu.py is the Python twin).
void test_multi_group_cluster(const std::string &core, const std::string &dir) {
  const size_t G = 4;
  const std::vector<std::string> peers = {"http://127.0.0.1:10000", "http://127.0.0.1:10001", "http://127.0.0.1:10002"};
  auto tr = std::make_shared<MultiLocalTransport>(G);
  std::vector<std::unique_ptr<MultiRaftPipe>> mp(3);
  std::vector<std::vector<std::unique_ptr<Collector>>> col(3);
  auto start = [&](int i) {
    MultiRaftPipeOptions o;
    o.tick_seconds = 0.005;
    o.waldir = dir + "/multi/raftsql-" + std::to_string(i + 1);
    o.transport = tr;
    if (core == "oracle")
      o.core_factory = [](uint32_t n, uint32_t id, size_t g) { return std::unique_ptr<MultiCore>(new OracleMultiCore(n, id, g)); };
    mp[i] = NewMultiRaftPipe(i + 1, peers, G, o);
    col[i].clear();
    for (size_t g = 0; g < G; ++g) col[i].emplace_back(new Collector(mp[i]->CommitC[g]));
  };
  ::mkdir((dir + "/multi").c_str(), 0750);
  for (int i = 0; i < 3; ++i) start(i);
  // A raft client retries: an entry accepted by a leader that is deposed before replicating it is lost (upstream
  // too), so proposals are at-least-once and the applied sequence is read modulo repeats.  Retries only ever fire
  // when the machine stalls for seconds; they keep the scenario meaningful instead of flaky when it does.
  auto uniq = [](const std::vector<std::string> &v) {
    std::vector<std::string> out;
    for (const auto &x : v)
      if (std::find(out.begin(), out.end(), x) == out.end()) out.push_back(x);
    return out;
  };
  auto has = [&](int i, size_t g, const std::string &text) {
    const auto v = col[i][g]->snapshot();
    return std::find(v.begin(), v.end(), text) != v.end();
  };
  auto leader_of = [&](size_t g, const std::vector<int> &alive) {
    for (int i : alive)
      if (mp[i]->node()->group(g)->role() == MRQ_ROLE_LEADER) return i;
    return -1;
  };
  auto propose_until_committed = [&](size_t g, const std::string &text, int via, const std::vector<int> &alive) {
    for (int attempt = 0; attempt < 8; ++attempt) {
      const int ld = leader_of(g, alive);
      mp[attempt == 0 || ld < 0 ? via : ld]->ProposeC[g]->send(text);
      for (int t = 0; t < 1000; ++t) {  // 5 s
        bool all = true;
        for (int i : alive) all = all && has(i, g, text);
        if (all) return true;
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
      }
    }
    return false;
  };
  const std::vector<int> all3 = {0, 1, 2}, two = {0, 2};
  // group g gets its own sequence, first through node g % 3 (forwarded to whoever leads); groups interleave in time
  for (int k = 0; k < 4; ++k)
    for (size_t g = 0; g < G; ++g)
      CHECK(propose_until_committed(g, "g" + std::to_string(g) + "-entry-" + std::to_string(k), (int)(g % 3), all3),
            "group %zu entry %d never committed on all nodes", g, k);
  for (size_t g = 0; g < G; ++g) {
    std::vector<std::string> want;
    for (int k = 0; k < 4; ++k) want.push_back("g" + std::to_string(g) + "-entry-" + std::to_string(k));
    for (int i = 0; i < 3; ++i) {
      CHECK(uniq(col[i][g]->snapshot()) == want, "node %d group %zu: wrong sequence (groups must not leak into each other)", i, g);
      CHECK(col[i][g]->snapshot() == col[0][g]->snapshot(), "node %d group %zu: applied a different sequence than node 0", i, g);
      CHECK(col[i][g]->nils.load() == 1, "node %d group %zu: %d nil sentinels", i, g, col[i][g]->nils.load());
    }
  }
  // stop node 2; every group re-elects among the other two and keeps committing with 2 of 3
  CHECK(mp[1]->Close().empty(), "clean stop of node 2");
  col[1].clear();
  for (size_t g = 0; g < G; ++g) {
    int leader = -1;
    for (int t = 0; t < 6000 && leader < 0; ++t) {
      leader = leader_of(g, two);
      if (leader < 0) std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    CHECK(leader >= 0, "group %zu must re-elect with 2 of 3 nodes", g);
    CHECK(propose_until_committed(g, "while-down-" + std::to_string(g), leader < 0 ? 0 : leader, two),
          "group %zu must commit with 2 of 3 nodes", g);
  }
  start(1);
  for (size_t g = 0; g < G; ++g) {
    const std::string last = "while-down-" + std::to_string(g);
    bool caught_up = false;
    for (int t = 0; t < 6000 && !caught_up; ++t) {
      caught_up = has(1, g, last);
      if (!caught_up) std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    CHECK(caught_up, "restarted node, group %zu: never caught up (has %zu entries)", g, col[1][g]->size());
    const auto got = uniq(col[1][g]->snapshot());
    CHECK(got.size() == 5 && got[0] == "g" + std::to_string(g) + "-entry-0" && got[4] == last,
          "restarted node, group %zu: replay + catch-up in order", g);
    CHECK(col[1][g]->nils.load() == 1, "restarted node, group %zu: one nil after the replay", g);
  }
  for (int i = 0; i < 3; ++i) CHECK(mp[i]->Close().empty(), "Close node %d", i);
}

// The multi-group node's WAL is ONE file for all groups, made durable once per tick, and nothing leaves the node
// before that fsync (wal.Save precedes transport.Send, raft.go:228-230).  tests/test_multipipe_cpu.py is the twin.
void test_group_commit_wal(const std::string &dir) {
  const size_t G = 5;
  struct Checked : LocalTransport {
    MultiHostNode **node;
    int *violations;
    void send(const std::vector<raftsql::Message> &msgs) override {
      if (!msgs.empty() && *node && (*node)->wal()->dirty()) ++*violations;
      LocalTransport::send(msgs);
    }
  };
  MultiHostNode *nodep = nullptr;
  int violations = 0;
  auto tr = std::make_shared<MultiLocalTransport>(0);
  for (size_t g = 0; g < G; ++g) {
    auto c = std::make_shared<Checked>();
    c->node = &nodep;
    c->violations = &violations;
    c->add(2);  // peers 2 and 3 exist as mailboxes only
    c->add(3);
    tr->groups.push_back(c);
  }
  const std::string wdir = dir + "/groupcommit";
  MultiHostNode node(std::unique_ptr<MultiCore>(new OracleMultiCore(3, 1, G)), 1, 3, G, tr, wdir);
  nodep = &node;
  node.start();
  int ticks = 0;
  for (int t = 0; t < 60; ++t) {  // every group times out and campaigns at its own tick: HardState changes -> records
    const uint64_t before = node.wal()->syncs();
    node.step_tick();
    ++ticks;
    CHECK(node.wal()->syncs() - before <= 1, "at most one fsync per tick, however many groups wrote");
    for (size_t g = 0; g < G; ++g) {  // node 2 grants every vote request it sees
      for (const raftsql::Message &m : tr->groups[g]->drain(2))
        if (m.type == kMsgVote) {
          raftsql::Message r;
          r.type = kMsgVoteResp;
          r.to = 1;
          r.from = 2;
          r.term = m.term;
          tr->groups[g]->send({r});
        }
      tr->groups[g]->drain(3);
    }
  }
  for (size_t g = 0; g < G; ++g) CHECK(node.group(g)->role() == MRQ_ROLE_LEADER, "group %zu must have elected this node", g);
  CHECK(violations == 0, "%d messages left the node before their tick's WAL records were durable", violations);
  CHECK(node.wal()->syncs() > 0 && node.wal()->syncs() <= (uint64_t)ticks && node.wal()->syncs() < G * 3,
        "%llu fsyncs for %zu groups over %d ticks", (unsigned long long)node.wal()->syncs(), G, ticks);
  node.stop();
  nodep = nullptr;
  // and the one file replays every group: term, vote for self, the leader's empty entry
  MultiWal again(wdir);
  for (size_t g = 0; g < G; ++g) {
    std::vector<Entry> ents;
    bool has_hs = false;
    uint64_t hs[3] = {0, 0, 0};
    again.view((uint32_t)g)->read_all(&ents, &has_hs, hs);
    CHECK(has_hs && hs[0] >= 1 && hs[1] == 1 && !ents.empty() && ents[0].data.empty(), "group %zu: replay of the shared file", g);
  }
}

// upstream raft_test.go TestHandleMsgApp (recalled; each row re-derived from raftLog.maybeAppend / commitTo in
// tests/test_hostnode_kat_cpu.py, which runs the same table over the Python host): follower of term 2 with the log
// [1:t1, 2:t2], committed 0; the host resolves the append against its log, the core applies the outcome.
void test_handle_msgapp_table(const std::string &dir) {
  struct Row {
    uint64_t index, logterm, commit;
    std::vector<uint64_t> ents;
    uint64_t windex, wcommit;
    bool wreject;
  };
  const std::vector<Row> rows = {
      {2, 3, 3, {}, 2, 0, true},     {3, 3, 3, {}, 2, 0, true},     {1, 1, 1, {}, 2, 1, false},    {0, 0, 1, {2}, 1, 1, false},
      {2, 2, 3, {2, 2}, 4, 3, false}, {2, 2, 4, {2}, 3, 3, false},   {1, 1, 4, {2}, 2, 2, false},   {1, 1, 3, {}, 2, 1, false},
      {1, 1, 3, {2}, 2, 2, false},   {2, 2, 3, {}, 2, 2, false},    {2, 2, 4, {}, 2, 2, false},
  };
  for (size_t k = 0; k < rows.size(); ++k) {
    const Row &r = rows[k];
    const std::string wdir = dir + "/kat-" + std::to_string(k);
    {
      Wal w(wdir);
      CHECK(w.open(), "open wal");
      std::vector<Entry> es(2);
      es[0].term = 1;
      es[0].data = "a";
      es[1].term = 2;
      es[1].data = "b";
      const uint64_t hs[3] = {2, 0, 0};
      w.save(hs, es, 1, false, 0);
    }
    auto tr = std::make_shared<LocalTransport>();
    tr->add(2);
    HostNode node(std::unique_ptr<Core>(new OracleCore(3, 1)), 1, 3, tr, wdir);
    node.start();
    raftsql::Message m;  // (the oracle's C header has a Message of its own)
    m.type = kMsgApp;
    m.to = 1;
    m.from = 2;
    m.term = 2;
    m.index = r.index;
    m.logterm = r.logterm;
    m.commit = r.commit;
    for (uint64_t t : r.ents) {
      Entry e;
      e.term = t;
      e.data = "y";
      m.entries.push_back(e);
    }
    tr->send({m});
    node.step_tick();
    const std::vector<raftsql::Message> rep = tr->drain(2);
    CHECK(rep.size() == 1 && rep[0].type == kMsgAppResp && rep[0].term == 2, "row %zu: one MsgAppResp of term 2", k);
    if (rep.size() != 1) continue;
    CHECK(rep[0].reject == r.wreject, "row %zu: reject", k);
    if (r.wreject)
      CHECK(rep[0].index == r.index && rep[0].reject_hint == 2, "row %zu: reject carries Index = m.Index, hint = lastIndex", k);
    else
      CHECK(rep[0].index == r.index + r.ents.size(), "row %zu: ack carries lastnewi", k);
    CHECK(node.commit() == r.wcommit, "row %zu: commit %llu, want %llu", k, (unsigned long long)node.commit(), (unsigned long long)r.wcommit);
    node.stop();
    // the log the node persisted is the log upstream ends with: replay it
    Wal w(wdir);
    std::vector<Entry> ents;
    bool has_hs = false;
    uint64_t hs[3];
    w.read_all(&ents, &has_hs, hs);
    CHECK(ents.size() == r.windex, "row %zu: lastIndex %zu, want %llu", k, ents.size(), (unsigned long long)r.windex);
    CHECK(has_hs && hs[2] == r.wcommit, "row %zu: persisted commit", k);
  }
}

// node 3 (term 2, log [1:t1, 2:t2]) hears in ONE tick MsgVote(term 7) from node 1 and MsgApp(term 2) from
// node 2.  The engine Steps in sender order: term 7 first, then it drops the append on the term rule — the host must
// not have appended / truncated / saved anything for it, and nothing may be acknowledged to the old leader.
void test_msgapp_behind_a_higher_term_message(const std::string &dir) {
  const std::string wdir = dir + "/stale-app";
  {
    Wal w(wdir);
    CHECK(w.open(), "open wal");
    std::vector<Entry> es(2);
    es[0].term = 1;
    es[0].data = "a";
    es[1].term = 2;
    es[1].data = "b";
    const uint64_t hs[3] = {2, 0, 0};
    w.save(hs, es, 1, false, 0);
  }
  auto tr = std::make_shared<LocalTransport>();
  tr->add(1);
  tr->add(2);
  HostNode node(std::unique_ptr<Core>(new OracleCore(3, 3)), 3, 3, tr, wdir);
  node.start();
  raftsql::Message app, vote;
  app.type = kMsgApp;
  app.to = 3;
  app.from = 2;
  app.term = 2;
  app.index = 1;
  app.logterm = 1;
  Entry e;
  e.term = 2;
  e.data = "would-truncate";
  app.entries.push_back(e);
  e.data = "more";
  app.entries.push_back(e);
  e.term = 2;
  vote.type = kMsgVote;
  vote.to = 3;
  vote.from = 1;
  vote.term = 7;
  vote.index = 9;
  vote.logterm = 9;
  tr->send({app, vote});
  node.step_tick();
  uint64_t term, v, commit, li, lt;
  node.hardstate(&term, &v, &commit, &li, &lt);
  CHECK(term == 7 && v == 1, "the vote is granted at the new term (term %llu vote %llu)", (unsigned long long)term, (unsigned long long)v);
  CHECK(li == 2 && lt == 2, "the stale append changed nothing (last %llu t%llu)", (unsigned long long)li, (unsigned long long)lt);
  bool acked = false;
  for (const raftsql::Message &m : tr->drain(2)) acked = acked || m.type == kMsgAppResp;
  CHECK(!acked, "nothing is acknowledged to the deposed leader");
  node.stop();
  Wal w(wdir);
  std::vector<Entry> ents;
  bool has_hs = false;
  uint64_t hs[3];
  w.read_all(&ents, &has_hs, hs);
  CHECK(ents.size() == 2 && ents[1].data == "b", "the WAL still holds the old log (%zu entries)", ents.size());
  CHECK(has_hs && hs[0] == 7 && hs[1] == 1, "and the hardstate of the new term");
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    std::printf("usage: raftpipe_test oracle|engine <tmpdir>\n");
    return 2;
  }
  const std::string core = argv[1], dir = argv[2];
  ::mkdir(dir.c_str(), 0750);
  ::mkdir((dir + "/single").c_str(), 0750);
  ::mkdir((dir + "/clus").c_str(), 0750);
  try {
    test_chan_semantics();
    test_wal_recovery(dir);
    test_handle_msgapp_table(dir);
    test_msgapp_behind_a_higher_term_message(dir);
    test_group_commit_wal(dir);
    test_single_node(core, dir);
    test_cluster_and_restart(core, dir);
    test_multi_group_cluster(core, dir);
  } catch (const std::exception &ex) {
    std::printf("FAIL exception: %s\n", ex.what());
    ++failures;
  }
  std::printf(failures ? "raftpipe_test: %d failure(s)\n" : "raftpipe_test: ok\n", failures);
  return failures ? 1 : 0;
}
