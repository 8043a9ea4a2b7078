// hostnode.hpp — the C++ host side of one raft node around the engine core (C-ABI, include/mrq.h).
//
// Mirrors what surrounds the hot path in the reference's raftNode (reference raft.go:38-273): the
// serveChannels loop (raft.go:204-246) — tick, drain Ready, persist, send, publish — with the consensus
// arithmetic (everything etcd-raft's node.Tick/Step/Propose/Ready computed: raft.go:214,224,227,269) done by the
// engine on the GPU.  Same design as raftsql_b200/hostnode.py (the two are tested against the same scenarios):
// the host keeps what the engine deliberately does not — entry payloads and per-entry terms (the log),
// follower-side log matching (maybeAppend), Progress.Next bookkeeping for sendAppend, and the WAL.
#pragma once

#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

namespace raftsql {

// raftpb.MessageType values used on this path (v2.2/2.3 numbering; include/mrq.h MRQ_MSG_*)
enum : int { kMsgProp = 2, kMsgApp = 3, kMsgAppResp = 4, kMsgVote = 5, kMsgVoteResp = 6, kMsgHeartbeat = 8, kMsgHeartbeatResp = 9 };

struct Entry {
  uint64_t term = 0;
  std::string data;
};

struct Message {  // raftpb.Message, the fields this system uses
  int type = 0;
  uint32_t to = 0, from = 0;
  uint64_t term = 0, logterm = 0, index = 0, commit = 0, reject_hint = 0;
  bool reject = false;
  std::vector<Entry> entries;
};

// What the host needs from the consensus core for ONE group.  The product implementation is EngineCore
// (libmrq.so through the C-ABI); the test suite also plugs its CPU checker in through this interface.
struct CoreState {
  uint64_t term = 0, vote = 0, committed = 0, last_index = 0, last_term = 0;
  uint32_t role = 0, lead = 0, out = 0;
  std::vector<uint64_t> match;  // [R]
};
struct CoreMsg {
  uint32_t from = 0;
  uint32_t type = 0;  // MRQ_MSG_* | MRQ_MSG_REJECT
  uint64_t term = 0, index = 0, logterm = 0, commit = 0;
};
class Core {
 public:
  virtual ~Core() {}
  virtual void import_hardstate(uint64_t term, uint64_t vote, uint64_t committed, uint64_t last_index, uint64_t last_term) = 0;
  // one tick: Step the messages (one per sender), apply n_proposals, Tick(); then report the Ready state
  virtual CoreState tick(const std::vector<CoreMsg> &msgs, uint32_t n_proposals) = 0;
};

// raft.MemoryStorage + the entry half of raftLog (reference raft.go:70,129,229); index-1 based, no compaction.
class Log {
 public:
  std::vector<Entry> ents;
  uint64_t last_index() const { return ents.size(); }
  uint64_t term(uint64_t i) const { return (i >= 1 && i <= ents.size()) ? ents[i - 1].term : 0; }
  uint64_t last_term() const { return term(ents.size()); }
  std::vector<Entry> slice(uint64_t lo, size_t max_bytes = 1u << 20) const;  // MaxSizePerMsg (raft.go:157)
  // upstream raftLog.maybeAppend: returns ok; lastnewi, first index written (0: none), whether a suffix was truncated
  bool maybe_append(uint64_t index, uint64_t logterm, const std::vector<Entry> &es, uint64_t *lastnewi, uint64_t *first_written,
                    bool *truncated);
};

// A minimal write-ahead log standing in for etcd `wal` (reference raft.go:100-124,228): directory raftsql-<id>,
// one append-only file of binary records  [u8 kind][u64 a][u64 b][u64 c][u32 len][len bytes]:
//   'H' hardstate (a=term b=vote c=commit) | 'E' entry (a=index b=term, payload) | 'T' truncate after index a.
class Wal {
 public:
  explicit Wal(const std::string &dir) : dir_(dir), path_(dir + "/wal.bin") {}
  virtual ~Wal() { close(); }
  static bool exist(const std::string &dir);  // wal.Exist (raft.go:100,145)
  virtual bool existed() const { return exist(dir_); }
  virtual bool open();
  virtual void close();
  // wal.ReadAll (raft.go:124): entries and the last hardstate (has_hs false if none)
  virtual void read_all(std::vector<Entry> *ents, bool *has_hs, uint64_t hs[3]);
  virtual void save(const uint64_t *hs /*nullable [3]*/, const std::vector<Entry> &new_entries, uint64_t first_index, bool truncate,
                    uint64_t truncate_after);  // wal.Save (raft.go:228): fsync'ed
  const std::string &dir() const { return dir_; }

 private:
  void put(char kind, uint64_t a, uint64_t b, uint64_t c, const std::string &payload);
  std::string dir_, path_;
  FILE *f_ = nullptr;
};

// Group-commit write-ahead log of a multi-group node: ONE append-only file for all groups (<dir>/multiwal.bin,
// Wal's records with a u32 group tag after the kind byte), made durable with ONE fsync per tick however many groups
// wrote.  view(g) is the Wal interface scoped to group g: its save() only appends; MultiHostNode calls sync() once per
// tick before anything is sent.
class MultiWal {
 public:
  explicit MultiWal(const std::string &dir) : dir_(dir), path_(dir + "/multiwal.bin") {}
  ~MultiWal() { close(); }
  bool existed() const;
  bool open();
  void close();
  void sync();  // the tick's one fsync (wal.Save's durability point, raft.go:228, for every group at once)
  uint64_t syncs() const { return syncs_; }
  bool dirty() const { return dirty_; }
  std::unique_ptr<Wal> view(uint32_t g);
  const std::string &dir() const { return dir_; }

 private:
  friend class GroupWal;
  struct Replayed {
    std::vector<Entry> ents;
    bool has_hs = false;
    uint64_t hs[3] = {0, 0, 0};
  };
  void put(uint32_t g, char kind, uint64_t a, uint64_t b, uint64_t c, const std::string &payload);
  void parse();                          // replays the file's valid prefix, once
  const Replayed &replayed(uint32_t g);
  std::string dir_, path_;
  FILE *f_ = nullptr;
  bool dirty_ = false, parsed_ = false;
  uint64_t syncs_ = 0;
  std::map<uint32_t, Replayed> groups_;
};

// In-process stand-in for rafthttp.Transport (reference raft.go:170-184,230,259).
class LocalTransport {
 public:
  virtual ~LocalTransport() {}
  virtual void add(uint32_t id);
  virtual void remove(uint32_t id);
  virtual void send(const std::vector<Message> &msgs);  // unknown / stopped peers lose messages, like a dead TCP peer
  virtual std::vector<Message> drain(uint32_t id);

 private:
  std::mutex mu_;
  std::map<uint32_t, std::vector<Message>> boxes_;
};

// One raft node for ONE group (the raftsql shape: G = 1, R = len(peers)).
class HostNode {
 public:
  HostNode(std::unique_ptr<Core> core, uint32_t id, uint32_t npeers, std::shared_ptr<LocalTransport> tr, const std::string &waldir);
  // with a WAL object of the caller's (a group's view of a shared group-commit WAL); nullptr = no WAL
  HostNode(std::unique_ptr<Core> core, uint32_t id, uint32_t npeers, std::shared_ptr<LocalTransport> tr, std::unique_ptr<Wal> wal);
  ~HostNode();
  // replayWAL (raft.go:122-134): rebuild the log, restore HardState; returns the committed payloads to replay
  std::vector<std::string> start();
  void propose(const std::string &data);
  // one iteration of serveChannels (raft.go:221-245); returns the payloads newly committed, in log order
  std::vector<std::string> step_tick();
  // step_tick in two halves around the core's tick, so that a multi-group host can prepare every group, tick
  // the shared engine ONCE, and finish every group (MultiHostNode below): step_tick() is
  // finish_tick(core->tick(p.msgs, p.nprop), p) with p = prepare_tick().
  struct Prepared {
    std::vector<CoreMsg> msgs;  // this tick's inbox of the group (MsgApp already resolved against the log)
    uint32_t nprop = 0;         // proposals handed to the core this tick
    std::map<uint32_t, Message> replies;
  };
  Prepared prepare_tick();
  std::vector<std::string> finish_tick(const CoreState &s, Prepared &p);
  // what start() restored from the WAL, for hosts that import it into a shared core themselves
  void hardstate(uint64_t *term, uint64_t *vote, uint64_t *commit, uint64_t *last_index, uint64_t *last_term) const;
  void stop();
  uint32_t role() const { return role_; }
  uint64_t term() const { return term_; }
  uint64_t commit() const { return commit_; }

 private:
  bool resolve_append(const Message &m, std::map<uint32_t, Message> *replies, CoreMsg *out, uint64_t eff_term, uint32_t eff_role);
  std::vector<std::string> ready(const CoreState &s, std::map<uint32_t, Message> &replies);
  std::vector<uint32_t> peers() const;

  std::unique_ptr<Core> core_;
  uint32_t id_, n_;
  std::shared_ptr<LocalTransport> tr_;
  std::unique_ptr<Wal> wal_;
  Log log_;
  std::vector<std::string> pending_, inflight_;
  std::vector<uint64_t> next_;
  std::set<uint32_t> behind_;
  std::vector<Message> backlog_;
  uint64_t applied_ = 0, vote_ = 0;
  uint32_t lead_ = 0;
  // read by role() / term() / commit() from other threads (operators, tests) while the node's thread advances them
  std::atomic<uint64_t> term_{0}, commit_{0};
  std::atomic<uint32_t> role_{0};
  bool stopped_ = false;
};

// ---- many groups of one node over ONE core (the multi-group seam) -----------------------------------
// What a multi-group host needs from the consensus core: the same two calls as Core, for G groups at once — the
// product implementation is one engine of G groups (ONE mrq_tick per tick for all of them).
class MultiCore {
 public:
  virtual ~MultiCore() {}
  // restored columns, [G] each
  virtual void import_hardstate(const std::vector<uint64_t> &term, const std::vector<uint64_t> &vote,
                                const std::vector<uint64_t> &committed, const std::vector<uint64_t> &last_index,
                                const std::vector<uint64_t> &last_term) = 0;
  // one tick of every group: msgs[g] (one per sender), nprop[g]; returns the Ready state of every group
  virtual std::vector<CoreState> tick(const std::vector<std::vector<CoreMsg>> &msgs, const std::vector<uint32_t> &nprop) = 0;
};

// In-process stand-in for rafthttp with one mailbox set per group.
struct MultiLocalTransport {
  explicit MultiLocalTransport(size_t n_groups) {
    for (size_t g = 0; g < n_groups; ++g) groups.push_back(std::make_shared<LocalTransport>());
  }
  std::vector<std::shared_ptr<LocalTransport>> groups;
};

// G HostNodes (the single-group host logic, unchanged: log, maybeAppend, Progress.Next, Ready.Messages, WAL)
// around one MultiCore: per-group prepare, ONE core tick, per-group Ready.
class MultiHostNode {
 public:
  MultiHostNode(std::unique_ptr<MultiCore> core, uint32_t id, uint32_t npeers, size_t n_groups,
                std::shared_ptr<MultiLocalTransport> tr, const std::string &waldir /* "" = none; else one MultiWal there */);
  std::vector<std::vector<std::string>> start();  // replayWAL for every group; committed payloads per group
  void propose(size_t g, const std::string &data) { nodes_[g]->propose(data); }
  // newly committed payloads per group, in log order.  Durability is batched like the arithmetic: the groups' WAL
  // records of the tick become durable with ONE fsync, and only then do the tick's messages leave the node
  // (persist before send: wal.Save precedes transport.Send, raft.go:228-230).
  std::vector<std::vector<std::string>> step_tick();
  void stop();
  size_t n_groups() const { return nodes_.size(); }
  HostNode *group(size_t g) { return nodes_[g].get(); }
  MultiWal *wal() { return wal_.get(); }
  // messages waiting for the tick's fsync: (group's transport, messages)
  std::vector<std::pair<std::shared_ptr<LocalTransport>, std::vector<Message>>> outbox;

 private:
  void flush();
  std::unique_ptr<MultiCore> core_;
  std::shared_ptr<MultiWal> wal_;
  std::vector<std::unique_ptr<HostNode>> nodes_;
};

}  // namespace raftsql
