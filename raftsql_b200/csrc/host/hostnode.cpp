// hostnode.cpp — see hostnode.hpp.  Reference anchors are cited per function.
#include "hostnode.hpp"

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "../../../include/mrq.h"

namespace raftsql {

// ---- Log -------------------------------------------------------------------------------------------------
std::vector<Entry> Log::slice(uint64_t lo, size_t max_bytes) const {
  std::vector<Entry> out;
  size_t size = 0;
  for (uint64_t i = lo; i <= ents.size(); ++i) {
    size += ents[i - 1].data.size();
    if (!out.empty() && size > max_bytes) break;
    out.push_back(ents[i - 1]);
  }
  return out;
}

bool Log::maybe_append(uint64_t index, uint64_t logterm, const std::vector<Entry> &es, uint64_t *lastnewi, uint64_t *first_written,
                       bool *truncated) {
  *first_written = 0;
  *truncated = false;
  *lastnewi = 0;
  if (term(index) != logterm && !(index == 0 && logterm == 0)) return false;  // matchTerm failed
  *lastnewi = index + es.size();
  for (size_t k = 0; k < es.size(); ++k) {
    const uint64_t i = index + 1 + k;
    if (i <= ents.size()) {
      if (ents[i - 1].term != es[k].term) {  // findConflict: truncate the suffix, then append
        ents.resize(i - 1);
        *truncated = true;
        ents.push_back(es[k]);
        if (!*first_written) *first_written = i;
      }
    } else {
      ents.push_back(es[k]);
      if (!*first_written) *first_written = i;
    }
  }
  return true;
}

// ---- Wal -------------------------------------------------------------------------------------------------
bool Wal::exist(const std::string &dir) {
  struct stat st;
  return ::stat((dir + "/wal.bin").c_str(), &st) == 0;
}

bool Wal::open() {
  ::mkdir(dir_.c_str(), 0750);  // raft.go:101
  f_ = std::fopen(path_.c_str(), "ab");
  return f_ != nullptr;
}

void Wal::close() {
  if (f_) {
    std::fclose(f_);
    f_ = nullptr;
  }
}

void Wal::put(char kind, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) {
  const uint32_t len = (uint32_t)payload.size();
  std::fwrite(&kind, 1, 1, f_);
  std::fwrite(&a, 8, 1, f_);
  std::fwrite(&b, 8, 1, f_);
  std::fwrite(&c, 8, 1, f_);
  std::fwrite(&len, 4, 1, f_);
  if (len) std::fwrite(payload.data(), 1, len, f_);
}

void Wal::save(const uint64_t *hs, const std::vector<Entry> &new_entries, uint64_t first_index, bool truncate, uint64_t truncate_after) {
  if (!f_) return;
  if (truncate) put('T', truncate_after, 0, 0, std::string());
  for (size_t k = 0; k < new_entries.size(); ++k) put('E', first_index + k, new_entries[k].term, 0, new_entries[k].data);
  if (hs) put('H', hs[0], hs[1], hs[2], std::string());
  std::fflush(f_);
  ::fsync(fileno(f_));
}

void Wal::read_all(std::vector<Entry> *ents, bool *has_hs, uint64_t hs[3]) {
  ents->clear();
  *has_hs = false;
  FILE *f = std::fopen(path_.c_str(), "rb");
  if (!f) return;
  std::fseek(f, 0, SEEK_END);
  const long file_size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  for (;;) {
    char kind;
    uint64_t a, b, c;
    uint32_t len;
    if (std::fread(&kind, 1, 1, f) != 1 || std::fread(&a, 8, 1, f) != 1 || std::fread(&b, 8, 1, f) != 1 ||
        std::fread(&c, 8, 1, f) != 1 || std::fread(&len, 4, 1, f) != 1)
      break;
    if ((long)len > file_size - std::ftell(f)) break;  // torn tail record: its length field outruns the file
    std::string payload(len, '\0');
    if (len && std::fread(&payload[0], 1, len, f) != len) break;  // torn tail record
    if (kind == 'H') {
      *has_hs = true;
      hs[0] = a;
      hs[1] = b;
      hs[2] = c;
    } else if (kind == 'E') {
      if (a >= 1 && a <= ents->size() + 1) {
        ents->resize(a - 1);
        Entry e;
        e.term = b;
        e.data = payload;
        ents->push_back(e);
      }
    } else if (kind == 'T') {
      if (a < ents->size()) ents->resize(a);
    }
  }
  std::fclose(f);
}

// ---- MultiWal ------------------------------------------------------------------------------------------------
// records: [u8 kind][u32 group][u64 a][u64 b][u64 c][u32 len][len bytes]
bool MultiWal::existed() const {
  struct stat st;
  return ::stat(path_.c_str(), &st) == 0;
}

bool MultiWal::open() {
  if (f_) return true;
  ::mkdir(dir_.c_str(), 0750);
  f_ = std::fopen(path_.c_str(), "ab");
  return f_ != nullptr;
}

void MultiWal::close() {
  if (f_) {
    sync();
    std::fclose(f_);
    f_ = nullptr;
  }
}

void MultiWal::put(uint32_t g, char kind, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) {
  if (!f_) return;
  const uint32_t len = (uint32_t)payload.size();
  std::fwrite(&kind, 1, 1, f_);
  std::fwrite(&g, 4, 1, f_);
  std::fwrite(&a, 8, 1, f_);
  std::fwrite(&b, 8, 1, f_);
  std::fwrite(&c, 8, 1, f_);
  std::fwrite(&len, 4, 1, f_);
  if (len) std::fwrite(payload.data(), 1, len, f_);
  dirty_ = true;
}

void MultiWal::sync() {
  if (!f_ || !dirty_) return;
  std::fflush(f_);
  ::fsync(fileno(f_));
  dirty_ = false;
  ++syncs_;
}

const MultiWal::Replayed &MultiWal::replayed(uint32_t g) {
  if (!parsed_) {
    parsed_ = true;
    FILE *f = std::fopen(path_.c_str(), "rb");
    if (f) {
      std::fseek(f, 0, SEEK_END);
      const long file_size = std::ftell(f);
      std::fseek(f, 0, SEEK_SET);
      for (;;) {
        char kind;
        uint32_t grp, len;
        uint64_t a, b, c;
        if (std::fread(&kind, 1, 1, f) != 1 || std::fread(&grp, 4, 1, f) != 1 || std::fread(&a, 8, 1, f) != 1 ||
            std::fread(&b, 8, 1, f) != 1 || std::fread(&c, 8, 1, f) != 1 || std::fread(&len, 4, 1, f) != 1)
          break;
        if ((long)len > file_size - std::ftell(f)) break;  // torn tail record
        std::string payload(len, '\0');
        if (len && std::fread(&payload[0], 1, len, f) != len) break;
        Replayed &r = groups_[grp];
        if (kind == 'H') {
          r.has_hs = true;
          r.hs[0] = a;
          r.hs[1] = b;
          r.hs[2] = c;
        } else if (kind == 'E') {
          if (a >= 1 && a <= r.ents.size() + 1) {
            r.ents.resize(a - 1);
            Entry e;
            e.term = b;
            e.data = payload;
            r.ents.push_back(e);
          }
        } else if (kind == 'T') {
          if (a < r.ents.size()) r.ents.resize(a);
        }
      }
      std::fclose(f);
    }
  }
  return groups_[g];
}

// Wal's interface for one group of a MultiWal: appends only, the owner syncs once per tick.
class GroupWal : public Wal {
 public:
  GroupWal(MultiWal *w, uint32_t g) : Wal(w->dir()), w_(w), g_(g) {}
  bool existed() const override { return w_->existed(); }
  bool open() override { return w_->open(); }
  void close() override {}  // the owner closes the shared file
  void read_all(std::vector<Entry> *ents, bool *has_hs, uint64_t hs[3]) override {
    const MultiWal::Replayed &r = w_->replayed(g_);
    *ents = r.ents;
    *has_hs = r.has_hs;
    for (int k = 0; k < 3; ++k) hs[k] = r.hs[k];
  }
  void save(const uint64_t *hs, const std::vector<Entry> &new_entries, uint64_t first_index, bool truncate,
            uint64_t truncate_after) override {
    if (truncate) w_->put(g_, 'T', truncate_after, 0, 0, std::string());
    for (size_t k = 0; k < new_entries.size(); ++k) w_->put(g_, 'E', first_index + k, new_entries[k].term, 0, new_entries[k].data);
    if (hs) w_->put(g_, 'H', hs[0], hs[1], hs[2], std::string());
  }

 private:
  MultiWal *w_;
  uint32_t g_;
};

std::unique_ptr<Wal> MultiWal::view(uint32_t g) { return std::unique_ptr<Wal>(new GroupWal(this, g)); }

// ---- LocalTransport ------------------------------------------------------------------------------------------
void LocalTransport::add(uint32_t id) {
  std::lock_guard<std::mutex> lk(mu_);
  boxes_[id];
}
void LocalTransport::remove(uint32_t id) {
  std::lock_guard<std::mutex> lk(mu_);
  boxes_.erase(id);
}
void LocalTransport::send(const std::vector<Message> &msgs) {
  std::lock_guard<std::mutex> lk(mu_);
  for (const Message &m : msgs) {
    auto it = boxes_.find(m.to);
    if (it != boxes_.end()) it->second.push_back(m);
  }
}
std::vector<Message> LocalTransport::drain(uint32_t id) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = boxes_.find(id);
  if (it == boxes_.end()) return {};
  std::vector<Message> out;
  out.swap(it->second);
  return out;
}

// ---- HostNode ---------------------------------------------------------------------------------------------------
HostNode::HostNode(std::unique_ptr<Core> core, uint32_t id, uint32_t npeers, std::shared_ptr<LocalTransport> tr,
                   const std::string &waldir)
    : core_(std::move(core)), id_(id), n_(npeers), tr_(std::move(tr)), next_(npeers + 1, 1) {
  if (!waldir.empty()) wal_.reset(new Wal(waldir));
  tr_->add(id_);
}

HostNode::HostNode(std::unique_ptr<Core> core, uint32_t id, uint32_t npeers, std::shared_ptr<LocalTransport> tr,
                   std::unique_ptr<Wal> wal)
    : core_(std::move(core)), id_(id), n_(npeers), tr_(std::move(tr)), wal_(std::move(wal)), next_(npeers + 1, 1) {
  tr_->add(id_);
}

HostNode::~HostNode() { stop(); }

void HostNode::stop() {
  if (stopped_) return;
  stopped_ = true;
  tr_->remove(id_);
  if (wal_) wal_->close();
}

std::vector<uint32_t> HostNode::peers() const {
  std::vector<uint32_t> p;
  for (uint32_t i = 1; i <= n_; ++i)
    if (i != id_) p.push_back(i);
  return p;
}

std::vector<std::string> HostNode::start() {
  std::vector<std::string> replay;
  if (!wal_) return replay;
  const bool old = wal_->existed();
  bool has_hs = false;
  uint64_t hs[3] = {0, 0, 0};
  std::vector<Entry> ents;
  if (old) wal_->read_all(&ents, &has_hs, hs);
  // "raftsql: create wal error" / "error loading wal" are fatal in the reference (raft.go:102,107,114)
  if (!wal_->open()) throw std::runtime_error("raftsql: cannot open the wal in " + wal_->dir());
  log_.ents = ents;
  if (has_hs) {  // unlike the reference (raft.go:124 discards it) the HardState is restored
    term_ = hs[0];
    vote_ = hs[1];
    commit_ = hs[2];
  }
  commit_ = std::min<uint64_t>(commit_, log_.last_index());
  if (core_) core_->import_hardstate(term_, vote_, commit_, log_.last_index(), log_.last_term());
  for (uint64_t i = 1; i <= commit_; ++i)
    if (!log_.ents[i - 1].data.empty()) replay.push_back(log_.ents[i - 1].data);
  applied_ = commit_;
  return replay;
}

void HostNode::hardstate(uint64_t *term, uint64_t *vote, uint64_t *commit, uint64_t *last_index, uint64_t *last_term) const {
  *term = term_;
  *vote = vote_;
  *commit = commit_;
  *last_index = log_.last_index();
  *last_term = log_.last_term();
}

void HostNode::propose(const std::string &data) { pending_.push_back(data); }

bool HostNode::resolve_append(const Message &m, std::map<uint32_t, Message> *replies, CoreMsg *out) {
  out->from = m.from;
  out->type = MRQ_MSG_APP;
  out->term = m.term;
  out->index = out->logterm = out->commit = 0;
  if (m.term < term_ || (role_ == MRQ_ROLE_LEADER && m.term == term_)) return true;  // the engine drops it on the term rule
  Message rep;
  rep.type = kMsgAppResp;
  rep.to = m.from;
  rep.from = id_;
  if (m.index < commit_) {  // handleAppendEntries: already committed past it
    rep.index = commit_;
    (*replies)[m.from] = rep;
    out->index = log_.last_index();
    out->logterm = log_.last_term();
    out->commit = commit_;
    return true;
  }
  uint64_t lastnewi = 0, first_written = 0;
  bool truncated = false;
  if (!log_.maybe_append(m.index, m.logterm, m.entries, &lastnewi, &first_written, &truncated)) {
    rep.index = m.index;
    rep.reject = true;
    rep.reject_hint = log_.last_index();
    (*replies)[m.from] = rep;
    out->type = MRQ_MSG_APP | MRQ_MSG_REJECT;
    return true;
  }
  if (wal_ && first_written) {  // persist before acknowledging (wal.Save precedes transport.Send, raft.go:228-230)
    std::vector<Entry> tail(log_.ents.begin() + (first_written - 1), log_.ents.end());
    wal_->save(nullptr, tail, first_written, truncated, first_written - 1);
  }
  rep.index = lastnewi;
  (*replies)[m.from] = rep;
  out->index = log_.last_index();
  out->logterm = log_.last_term();
  out->commit = std::min(m.commit, lastnewi);
  return true;
}

std::vector<std::string> HostNode::step_tick() {
  Prepared p = prepare_tick();
  const CoreState s = core_->tick(p.msgs, p.nprop);
  return finish_tick(s, p);
}

std::vector<std::string> HostNode::finish_tick(const CoreState &s, Prepared &p) { return ready(s, p.replies); }

HostNode::Prepared HostNode::prepare_tick() {
  std::vector<Message> inbound = std::move(backlog_);
  backlog_.clear();
  {
    std::vector<Message> fresh = tr_->drain(id_);
    inbound.insert(inbound.end(), fresh.begin(), fresh.end());
  }
  Prepared prep;
  std::vector<CoreMsg> &eng_msgs = prep.msgs;
  std::map<uint32_t, Message> &replies = prep.replies;
  std::set<uint32_t> seen;
  for (const Message &m : inbound) {
    if (m.type == kMsgProp) {  // a follower forwarded client proposals to us
      for (const Entry &e : m.entries) pending_.push_back(e.data);
      continue;
    }
    if (seen.count(m.from)) {  // the engine inbox holds one message per sender per tick (include/mrq.h)
      backlog_.push_back(m);
      continue;
    }
    seen.insert(m.from);
    CoreMsg cm;
    if (m.type == kMsgApp) {
      resolve_append(m, &replies, &cm);
      eng_msgs.push_back(cm);
      continue;
    }
    if (m.type == kMsgAppResp && m.reject && role_ == MRQ_ROLE_LEADER && m.term == term_)
      next_[m.from] = std::max<uint64_t>(1, std::min(m.index, m.reject_hint + 1));  // Progress.maybeDecrTo
    if (m.type == kMsgHeartbeatResp && role_ == MRQ_ROLE_LEADER && m.term == term_) behind_.insert(m.from);
    cm.from = m.from;
    cm.type = (uint32_t)m.type | (m.reject ? MRQ_MSG_REJECT : 0u);
    cm.term = m.term;
    cm.index = m.index;
    cm.logterm = m.logterm;
    cm.commit = m.commit;
    eng_msgs.push_back(cm);
  }
  // node.Propose blocks until there is a leader (upstream node.run: propc is nil while lead == None)
  inflight_.clear();
  uint32_t nprop = 0;
  if (!pending_.empty() && role_ == MRQ_ROLE_LEADER) {
    const size_t n = std::min<size_t>(pending_.size(), 255);
    inflight_.assign(pending_.begin(), pending_.begin() + n);
    pending_.erase(pending_.begin(), pending_.begin() + n);
    nprop = (uint32_t)n;
  } else if (!pending_.empty() && lead_ != 0 && lead_ != id_) {
    Message fwd;
    fwd.type = kMsgProp;
    fwd.to = lead_;
    fwd.from = id_;
    for (const std::string &d : pending_) {
      Entry e;
      e.data = d;
      fwd.entries.push_back(e);
    }
    pending_.clear();
    tr_->send({fwd});
  }
  prep.nprop = nprop;
  return prep;
}

std::vector<std::string> HostNode::ready(const CoreState &s, std::map<uint32_t, Message> &replies) {
  const bool was_leader = role_ == MRQ_ROLE_LEADER;
  const bool hs_changed = s.term != term_ || s.vote != vote_ || s.committed != commit_;
  term_ = s.term;
  vote_ = s.vote;
  role_ = s.role;
  lead_ = s.lead;
  std::vector<Message> msgs;
  // entries the engine appended as leader: the empty entry of a new term, then the accepted proposals
  std::vector<Entry> new_entries;
  uint64_t first = 0;
  if (role_ == MRQ_ROLE_LEADER && s.last_index > log_.last_index()) {
    uint64_t n_new = s.last_index - log_.last_index();
    if (s.out & MRQ_OUT_BECAME_LEADER) {
      Entry e;
      e.term = term_;
      new_entries.push_back(e);
      --n_new;
    }
    const size_t acc = std::min<size_t>(n_new, inflight_.size());
    for (size_t k = 0; k < acc; ++k) {
      Entry e;
      e.term = term_;
      e.data = inflight_[k];
      new_entries.push_back(e);
    }
    pending_.insert(pending_.begin(), inflight_.begin() + acc, inflight_.end());
    first = log_.last_index() + 1;
    log_.ents.insert(log_.ents.end(), new_entries.begin(), new_entries.end());
    if (s.out & MRQ_OUT_BECAME_LEADER) std::fill(next_.begin(), next_.end(), first);  // reset(): Next = lastIndex+1
  } else if (!inflight_.empty()) {  // stepped down before the proposals were applied: nothing was appended
    pending_.insert(pending_.begin(), inflight_.begin(), inflight_.end());
  }
  inflight_.clear();
  if (wal_ && (!new_entries.empty() || hs_changed)) {
    const uint64_t hs[3] = {s.term, s.vote, s.committed};
    wal_->save(hs, new_entries, first, false, 0);
  }
  // ---- Ready.Messages, rebuilt from the out word (include/mrq.h MRQ_OUT_*) ------------------------------
  if (s.out & MRQ_OUT_CAMPAIGN) {
    for (uint32_t p : peers()) {
      Message m;
      m.type = kMsgVote;
      m.to = p;
      m.from = id_;
      m.term = term_;
      m.index = s.last_index;
      m.logterm = s.last_term;
      msgs.push_back(m);
    }
  }
  for (uint32_t p : peers()) {
    const uint32_t rep = (s.out >> (MRQ_OUT_VOTE_REPLY_SHIFT + 2 * (p - 1))) & 3u;
    if (rep) {
      Message m;
      m.type = kMsgVoteResp;
      m.to = p;
      m.from = id_;
      m.term = term_;
      m.reject = rep == 2;
      msgs.push_back(m);
    }
    if ((s.out >> (MRQ_OUT_ACK_REPLY_SHIFT + (p - 1))) & 1u) {
      Message m;
      auto it = replies.find(p);
      if (it != replies.end()) {
        m = it->second;
      } else {
        m.type = kMsgHeartbeatResp;
        m.to = p;
        m.from = id_;
      }
      m.term = term_;
      msgs.push_back(m);
    }
  }
  if (role_ == MRQ_ROLE_LEADER && (s.out & (MRQ_OUT_BCAST_APPEND | MRQ_OUT_BCAST_HEARTBEAT | MRQ_OUT_BECAME_LEADER))) {
    for (uint32_t p : peers()) {
      const uint64_t match = p - 1 < s.match.size() ? s.match[p - 1] : 0;
      if (behind_.count(p) && match < log_.last_index()) next_[p] = match + 1;  // stepLeader MsgHeartbeatResp
      const uint64_t nxt = std::max(next_[p], match + 1);
      if (nxt <= log_.last_index()) {  // sendAppend
        Message m;
        m.type = kMsgApp;
        m.to = p;
        m.from = id_;
        m.term = term_;
        m.index = nxt - 1;
        m.logterm = log_.term(nxt - 1);
        m.entries = log_.slice(nxt);
        m.commit = s.committed;
        next_[p] = nxt + m.entries.size();  // optimistic, like ProgressStateReplicate
        msgs.push_back(m);
      } else if (s.out & MRQ_OUT_BCAST_HEARTBEAT) {
        Message m;
        m.type = kMsgHeartbeat;
        m.to = p;
        m.from = id_;
        m.term = term_;
        m.commit = std::min(match, s.committed);
        msgs.push_back(m);
      }
    }
  }
  behind_.clear();
  if (was_leader && role_ != MRQ_ROLE_LEADER) std::fill(next_.begin(), next_.end(), 1);
  tr_->send(msgs);  // transport.Send(rd.Messages) (raft.go:230)
  // ---- publish (raft.go:82-96, gated on commit) ----------------------------------------------------------
  commit_ = s.committed;
  std::vector<std::string> published;
  while (applied_ < std::min<uint64_t>(commit_, log_.last_index())) {
    ++applied_;
    const std::string &d = log_.ents[applied_ - 1].data;
    if (!d.empty()) published.push_back(d);  // "ignore conf changes and empty messages" (raft.go:84-86)
  }
  return published;
}

// ---- MultiHostNode ------------------------------------------------------------------------------------------------
namespace {
// A group's transport whose sends wait in the node's outbox until the tick's WAL records are durable.
class DeferredTransport : public LocalTransport {
 public:
  DeferredTransport(std::shared_ptr<LocalTransport> tr, MultiHostNode *owner) : tr_(std::move(tr)), owner_(owner) {}
  void add(uint32_t id) override { tr_->add(id); }
  void remove(uint32_t id) override { tr_->remove(id); }
  std::vector<Message> drain(uint32_t id) override { return tr_->drain(id); }
  void send(const std::vector<Message> &msgs) override { owner_->outbox.emplace_back(tr_, msgs); }

 private:
  std::shared_ptr<LocalTransport> tr_;
  MultiHostNode *owner_;
};
}  // namespace

MultiHostNode::MultiHostNode(std::unique_ptr<MultiCore> core, uint32_t id, uint32_t npeers, size_t n_groups,
                             std::shared_ptr<MultiLocalTransport> tr, const std::string &waldir)
    : core_(std::move(core)) {
  if (!waldir.empty()) wal_ = std::make_shared<MultiWal>(waldir);  // group commit: one file, one fsync per tick
  for (size_t g = 0; g < n_groups; ++g)
    nodes_.emplace_back(new HostNode(nullptr, id, npeers, std::make_shared<DeferredTransport>(tr->groups.at(g), this),
                                     wal_ ? wal_->view((uint32_t)g) : std::unique_ptr<Wal>()));
}

std::vector<std::vector<std::string>> MultiHostNode::start() {
  const size_t G = nodes_.size();
  std::vector<std::vector<std::string>> replay(G);
  std::vector<uint64_t> term(G), vote(G), commit(G), li(G), lt(G);
  bool any = false;
  for (size_t g = 0; g < G; ++g) {
    replay[g] = nodes_[g]->start();
    nodes_[g]->hardstate(&term[g], &vote[g], &commit[g], &li[g], &lt[g]);
    any = any || term[g] || vote[g] || commit[g] || li[g];
  }
  if (any) core_->import_hardstate(term, vote, commit, li, lt);  // ONE import of the restored columns
  return replay;
}

std::vector<std::vector<std::string>> MultiHostNode::step_tick() {
  const size_t G = nodes_.size();
  std::vector<HostNode::Prepared> prep(G);
  std::vector<std::vector<CoreMsg>> msgs(G);
  std::vector<uint32_t> nprop(G, 0);
  for (size_t g = 0; g < G; ++g) {
    prep[g] = nodes_[g]->prepare_tick();
    msgs[g] = prep[g].msgs;
    nprop[g] = prep[g].nprop;
  }
  const std::vector<CoreState> st = core_->tick(msgs, nprop);  // ONE tick for all groups
  std::vector<std::vector<std::string>> out(G);
  for (size_t g = 0; g < G; ++g) out[g] = nodes_[g]->finish_tick(st.at(g), prep[g]);
  flush();
  return out;
}

// make this tick's WAL records of every group durable with one fsync, THEN let the messages out
void MultiHostNode::flush() {
  if (wal_) wal_->sync();
  auto pending = std::move(outbox);
  outbox.clear();
  for (auto &p : pending) p.first->send(p.second);
}

void MultiHostNode::stop() {
  for (auto &n : nodes_) n->stop();
  if (wal_) wal_->close();
}

}  // namespace raftsql
