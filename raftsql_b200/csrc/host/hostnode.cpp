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

// ---- WAL records ---------------------------------------------------------------------------------------------
// One format for both logs:  [u8 kind][u32 group][u64 a][u64 b][u64 c][u32 len][u32 crc][len bytes]
// crc = CRC-32 (IEEE) of everything before it plus the payload.  A scan stops at the first record that is cut
// short or fails its CRC (a crash mid-append); open() truncates the file to that valid prefix BEFORE appending,
// as etcd's wal repairs its tail — otherwise every record written after the first crash would sit behind the
// garbage and be dropped by the next replay (a node could vote twice in a term or lose acknowledged entries).
namespace {
uint32_t crc32_update(uint32_t crc, const void *data, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  const unsigned char *p = (const unsigned char *)data;
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
  return ~crc;
}
constexpr size_t kWalHeader = 1 + 4 + 8 + 8 + 8 + 4;  // bytes before the crc field

void wal_put(FILE *f, char kind, uint32_t g, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) {
  unsigned char h[kWalHeader + 4];
  const uint32_t len = (uint32_t)payload.size();
  h[0] = (unsigned char)kind;
  std::memcpy(h + 1, &g, 4);
  std::memcpy(h + 5, &a, 8);
  std::memcpy(h + 13, &b, 8);
  std::memcpy(h + 21, &c, 8);
  std::memcpy(h + 29, &len, 4);
  uint32_t crc = crc32_update(0, h, kWalHeader);
  if (len) crc = crc32_update(crc, payload.data(), len);
  std::memcpy(h + kWalHeader, &crc, 4);
  std::fwrite(h, 1, sizeof h, f);
  if (len) std::fwrite(payload.data(), 1, len, f);
}

// calls rec(kind, group, a, b, c, payload) for every valid record; returns the byte offset where the valid prefix ends
template <class F>
long wal_scan(const std::string &path, F rec) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return 0;
  std::fseek(f, 0, SEEK_END);
  const long file_size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  long good = 0;
  for (;;) {
    unsigned char h[kWalHeader + 4];
    if (std::fread(h, 1, sizeof h, f) != sizeof h) break;  // half a header
    uint32_t g, len, crc;
    uint64_t a, b, c;
    std::memcpy(&g, h + 1, 4);
    std::memcpy(&a, h + 5, 8);
    std::memcpy(&b, h + 13, 8);
    std::memcpy(&c, h + 21, 8);
    std::memcpy(&len, h + 29, 4);
    std::memcpy(&crc, h + kWalHeader, 4);
    if ((long)len > file_size - std::ftell(f)) break;  // its length field outruns the file
    std::string payload(len, '\0');
    if (len && std::fread(&payload[0], 1, len, f) != len) break;
    uint32_t want = crc32_update(0, h, kWalHeader);
    if (len) want = crc32_update(want, payload.data(), len);
    if (want != crc) break;  // bytes that never all reached the disk
    rec((char)h[0], g, a, b, c, payload);
    good = std::ftell(f);
  }
  std::fclose(f);
  return good;
}

// cut a torn tail off (and make the cut durable) so that appends land right behind the last valid record
void wal_repair_tail(const std::string &path) {
  struct stat st;
  if (::stat(path.c_str(), &st) != 0) return;
  const long good = wal_scan(path, [](char, uint32_t, uint64_t, uint64_t, uint64_t, const std::string &) {});
  if (good < (long)st.st_size) {
    if (::truncate(path.c_str(), good) == 0) {
      FILE *f = std::fopen(path.c_str(), "rb+");
      if (f) {
        ::fsync(fileno(f));
        std::fclose(f);
      }
    }
  }
}
}  // namespace

// ---- Wal -------------------------------------------------------------------------------------------------
bool Wal::exist(const std::string &dir) {
  struct stat st;
  return ::stat((dir + "/wal.bin").c_str(), &st) == 0;
}

bool Wal::open() {
  ::mkdir(dir_.c_str(), 0750);  // raft.go:101
  wal_repair_tail(path_);
  f_ = std::fopen(path_.c_str(), "ab");
  return f_ != nullptr;
}

void Wal::close() {
  if (f_) {
    std::fclose(f_);
    f_ = nullptr;
  }
}

void Wal::put(char kind, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) { wal_put(f_, kind, 0, a, b, c, payload); }

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
  wal_scan(path_, [&](char kind, uint32_t, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) {
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
  });
}

// ---- MultiWal ------------------------------------------------------------------------------------------------
bool MultiWal::existed() const {
  struct stat st;
  return ::stat(path_.c_str(), &st) == 0;
}

bool MultiWal::open() {
  if (f_) return true;
  ::mkdir(dir_.c_str(), 0750);
  parse();                   // replay first: the records of the valid prefix
  wal_repair_tail(path_);    // then cut the torn tail off before anything is appended
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
  wal_put(f_, kind, g, a, b, c, payload);
  dirty_ = true;
}

void MultiWal::sync() {
  if (!f_ || !dirty_) return;
  std::fflush(f_);
  ::fsync(fileno(f_));
  dirty_ = false;
  ++syncs_;
}

void MultiWal::parse() {
  if (parsed_) return;
  parsed_ = true;
  wal_scan(path_, [&](char kind, uint32_t grp, uint64_t a, uint64_t b, uint64_t c, const std::string &payload) {
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
  });
}

const MultiWal::Replayed &MultiWal::replayed(uint32_t g) {
  parse();
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

// Resolved against the (term, role) the engine will have when it Steps this message: prepare_tick walks the tick's
// messages in sender order, as the engine does, so an append the engine is going to drop on the term rule never
// touches the host log or the WAL.
bool HostNode::resolve_append(const Message &m, std::map<uint32_t, Message> *replies, CoreMsg *out, uint64_t eff_term,
                              uint32_t eff_role) {
  out->from = m.from;
  out->type = MRQ_MSG_APP;
  out->term = m.term;
  out->index = out->logterm = out->commit = 0;
  if (m.term < eff_term || (eff_role == MRQ_ROLE_LEADER && m.term == eff_term)) return true;  // the engine drops it on the term rule
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
  // One message per sender per tick (include/mrq.h), and at most ONE MsgApp per tick (a second one, from another
  // sender, was matched against a log the first is about to change): the rest waits in the backlog, in order.
  std::map<uint32_t, Message> chosen;
  bool have_app = false;
  for (const Message &m : inbound) {
    if (m.type == kMsgProp) {  // a follower forwarded client proposals to us
      for (const Entry &e : m.entries) pending_.push_back(e.data);
      continue;
    }
    if (chosen.count(m.from) || (m.type == kMsgApp && have_app)) {
      backlog_.push_back(m);
      continue;
    }
    chosen[m.from] = m;
    have_app = have_app || m.type == kMsgApp;
  }
  // The engine Steps the tick's messages in SENDER order: a higher-term message from a lower sender id
  // moves it to that term before it sees a later sender's MsgApp.  Track that effective (term, role) here.
  uint64_t eff_term = term_;
  uint32_t eff_role = role_;
  for (auto &kv : chosen) {  // std::map: ascending sender id
    const Message &m = kv.second;
    if (m.term > eff_term) {  // Step(): becomeFollower(m.Term, ...)
      eff_term = m.term;
      eff_role = MRQ_ROLE_FOLLOWER;
    }
    CoreMsg cm;
    if (m.type == kMsgApp) {
      bool votes_before = false;
      for (auto &kv2 : chosen) votes_before = votes_before || (kv2.first < kv.first && kv2.second.type == kMsgVoteResp);
      if (eff_role == MRQ_ROLE_CANDIDATE && m.term == eff_term && votes_before) {
        backlog_.insert(backlog_.begin(), m);  // the votes may make us leader within this tick: resolve it next tick
        continue;
      }
      resolve_append(m, &replies, &cm, eff_term, eff_role);
      eng_msgs.push_back(cm);
      if (eff_role == MRQ_ROLE_CANDIDATE && m.term == eff_term) eff_role = MRQ_ROLE_FOLLOWER;
      continue;
    }
    if (m.type == kMsgHeartbeat && eff_role == MRQ_ROLE_CANDIDATE && m.term == eff_term) eff_role = MRQ_ROLE_FOLLOWER;
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
  // the engine tracks (lastIndex, lastTerm) of the log the host keeps: after Ready they must agree (fatal otherwise,
  // like the reference's log.Fatalf sites: every later vote / commit decision would be about a log that does not exist)
  if (s.last_index != log_.last_index() || s.last_term != log_.last_term())
    throw std::runtime_error("host log and engine diverged: log (" + std::to_string(log_.last_index()) + ", t" +
                             std::to_string(log_.last_term()) + ") engine (" + std::to_string(s.last_index) + ", t" +
                             std::to_string(s.last_term) + ")");
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
