/*
 * raft_oracle.c — CPU oracle (see raft_oracle.h: TEST INFRASTRUCTURE ONLY, PARITY UNPINNED).
 *
 * Every function names the upstream etcd-raft (v2.2–v2.3 era) symbol it restates.  The upstream
 * source is not available in this environment; the reference reaches these
 * only through its call sites raft.go:152-165,214,224,227,235,269.
 */
#include "raft_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../include/mrq.h"
#include "../include/mrq_trace.h"

/* ------------------------------------------------------------------------------------------
 * raft/log.go
 * ---------------------------------------------------------------------------------------- */

/* upstream raftLog.term(i): term of entry i; 0 outside [dummy, lastIndex] */
uint64_t raftLog_term(const raftLog *l, uint64_t i) {
  if (i == 0 || i > l->lastIndex_) return 0;
  for (int k = l->nruns - 1; k >= 0; --k)
    if (l->runs[k].first <= i) return l->runs[k].term;
  return 0;
}

/* upstream raftLog.lastTerm() */
uint64_t raftLog_lastTerm(const raftLog *l) { return raftLog_term(l, l->lastIndex_); }

/* upstream raftLog.isUpToDate(lasti, term):
 *   term > l.lastTerm() || (term == l.lastTerm() && lasti >= l.lastIndex()) */
int raftLog_isUpToDate(const raftLog *l, uint64_t lasti, uint64_t term) {
  uint64_t lt = raftLog_lastTerm(l);
  return term > lt || (term == lt && lasti >= l->lastIndex_);
}

/* upstream raftLog.commitTo(tocommit): never decrease commit; panics if tocommit > lastIndex */
void raftLog_commitTo(raftLog *l, uint64_t tocommit, uint32_t *errors) {
  if (l->committed < tocommit) {
    if (l->lastIndex_ < tocommit) { /* upstream: l.logger.Panicf("tocommit(%d) is out of range") */
      if (errors) ++*errors;
      return;
    }
    l->committed = tocommit;
  }
}

/* upstream raftLog.maybeCommit(maxIndex, term):
 *   if maxIndex > l.committed && l.zeroTermOnErrCompacted(l.term(maxIndex)) == term { commitTo } */
int raftLog_maybeCommit(raftLog *l, uint64_t maxIndex, uint64_t term) {
  if (maxIndex > l->committed && raftLog_term(l, maxIndex) == term) {
    raftLog_commitTo(l, maxIndex, NULL);
    return 1;
  }
  return 0;
}

static void log_push_run(raftLog *l, uint64_t first, uint64_t term) {
  if (l->nruns == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 4;
    l->runs = (TermRun *)realloc(l->runs, sizeof(TermRun) * (size_t)l->cap);
  }
  l->runs[l->nruns].first = first;
  l->runs[l->nruns].term = term;
  ++l->nruns;
}

/* upstream raftLog.append(ents...) for n entries that all carry `term` (appendEntry stamps r.Term) */
void raftLog_append(raftLog *l, uint32_t n, uint64_t term) {
  if (n == 0) return;
  if (l->nruns == 0 || l->runs[l->nruns - 1].term != term) log_push_run(l, l->lastIndex_ + 1, term);
  l->lastIndex_ += n;
}

/* Host-resolved follower append (include/mrq.h MSG_APP): after the host's maybeAppend the log's
 * last entry is (index, logterm).  Keeps the run list consistent for term() queries. */
static void log_set_last(raftLog *l, uint64_t index, uint64_t logterm) {
  if (index < l->lastIndex_) { /* conflict: truncate */
    while (l->nruns > 0 && l->runs[l->nruns - 1].first > index) --l->nruns;
    l->lastIndex_ = index;
  }
  if (index > l->lastIndex_) {
    if (l->nruns == 0 || l->runs[l->nruns - 1].term != logterm) log_push_run(l, l->lastIndex_ + 1, logterm);
    l->lastIndex_ = index;
  } else if (index > 0 && raftLog_term(l, index) != logterm) { /* same length, last entry replaced */
    while (l->nruns > 0 && l->runs[l->nruns - 1].first >= index) --l->nruns;
    log_push_run(l, index, logterm);
  }
}

/* ------------------------------------------------------------------------------------------
 * raft/progress.go
 * ---------------------------------------------------------------------------------------- */

/* upstream Progress.maybeUpdate(n) */
int Progress_maybeUpdate(Progress *pr, uint64_t n) {
  int updated = 0;
  if (pr->Match < n) {
    pr->Match = n;
    updated = 1;
  }
  if (pr->Next < n + 1) pr->Next = n + 1;
  return updated;
}

/* ------------------------------------------------------------------------------------------
 * raft/raft.go
 * ---------------------------------------------------------------------------------------- */

/* upstream raft.q(): len(r.prs)/2 + 1 */
int raft_q(const raft *r) { return r->nprs / 2 + 1; }

/* upstream raft.poll(id, v): first vote from id wins; returns the number granted */
int raft_poll(raft *r, uint64_t id, int v) {
  if (r->votes[id - 1] < 0) r->votes[id - 1] = (int8_t)(v ? 1 : 0);
  int granted = 0;
  for (int i = 0; i < r->nprs; ++i)
    if (r->votes[i] == 1) ++granted;
  return granted;
}

static int votes_len(const raft *r) { /* len(r.votes) */
  int n = 0;
  for (int i = 0; i < r->nprs; ++i)
    if (r->votes[i] >= 0) ++n;
  return n;
}

/* upstream raft.reset(term) (+ v3's resetRandomizedElectionTimeout, drawn from the counter RNG) */
void raft_reset(raft *r, uint64_t term) {
  if (r->Term != term) {
    r->Term = term;
    r->Vote = ORC_None;
  }
  r->lead = ORC_None;
  r->electionElapsed = 0;
  r->heartbeatElapsed = 0;
  r->randomizedElectionTimeout =
      (int)mrq_randomized_timeout(r->rand_seed, r->rand_group, *r->tick_no, (uint32_t)r->electionTimeout);
  for (int i = 0; i < r->nprs; ++i) r->votes[i] = -1;
  for (int i = 0; i < r->nprs; ++i) {
    r->prs[i].Match = 0;
    r->prs[i].Next = r->raftLog.lastIndex_ + 1;
    if ((uint64_t)(i + 1) == r->id) r->prs[i].Match = r->raftLog.lastIndex_;
  }
}

/* upstream raft.becomeFollower(term, lead) */
void raft_becomeFollower(raft *r, uint64_t term, uint64_t lead) {
  if (r->state != StateFollower) r->out |= MRQ_OUT_STEPPED_DOWN;
  raft_reset(r, term);
  r->lead = lead;
  r->state = StateFollower;
}

/* upstream raft.becomeCandidate() */
void raft_becomeCandidate(raft *r) {
  raft_reset(r, r->Term + 1);
  r->Vote = r->id;
  r->state = StateCandidate;
}

/* upstream raft.maybeCommit(): sort matches descending, mci = mis[q()-1] */
static int cmp_desc(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? 1 : (x > y ? -1 : 0);
}
int raft_maybeCommit(raft *r) {
  uint64_t mis[ORC_MAXR];
  for (int i = 0; i < r->nprs; ++i) mis[i] = r->prs[i].Match;
  qsort(mis, (size_t)r->nprs, sizeof(uint64_t), cmp_desc);
  uint64_t mci = mis[raft_q(r) - 1];
  int ok = raftLog_maybeCommit(&r->raftLog, mci, r->Term);
  if (ok) r->out |= MRQ_OUT_COMMIT_ADVANCED;
  return ok;
}

/* upstream raft.appendEntry(es...): stamp Term/Index, append, prs[id].maybeUpdate(lastIndex), maybeCommit */
void raft_appendEntry(raft *r, uint32_t n) {
  raftLog_append(&r->raftLog, n, r->Term);
  Progress_maybeUpdate(&r->prs[r->id - 1], r->raftLog.lastIndex_);
  raft_maybeCommit(r);
}

/* upstream raft.becomeLeader(): reset(Term), lead = id, appendEntry(empty) */
void raft_becomeLeader(raft *r) {
  raft_reset(r, r->Term);
  r->lead = r->id;
  r->state = StateLeader;
  r->out |= MRQ_OUT_BECAME_LEADER;
  raft_appendEntry(r, 1);
}

/* upstream raft.campaign() */
void raft_campaign(raft *r) {
  raft_becomeCandidate(r);
  if (raft_q(r) == raft_poll(r, r->id, 1)) {
    raft_becomeLeader(r);
    return;
  }
  r->out |= MRQ_OUT_CAMPAIGN; /* send MsgVote{Index: lastIndex, LogTerm: lastTerm} to every other peer */
}

static void reply_vote(raft *r, uint64_t to, int reject) {
  r->out |= (uint32_t)(reject ? 2u : 1u) << (MRQ_OUT_VOTE_REPLY_SHIFT + 2u * (uint32_t)(to - 1));
}
static void reply_ack(raft *r, uint64_t to) { r->out |= 1u << (MRQ_OUT_ACK_REPLY_SHIFT + (uint32_t)(to - 1)); }

/* upstream raft.handleAppendEntries(m), with the log-matching half resolved by the host (mrq.h) */
static void raft_handleAppendEntries(raft *r, const Message *m) {
  if (!m->Reject) {
    log_set_last(&r->raftLog, m->Index, m->LogTerm);
    uint64_t before = r->raftLog.committed;
    raftLog_commitTo(&r->raftLog, m->Commit, &r->errors);
    if (r->raftLog.committed != before) r->out |= MRQ_OUT_COMMIT_ADVANCED;
  }
  reply_ack(r, m->From); /* MsgAppResp either way */
}

/* upstream raft.handleHeartbeat(m): commitTo(m.Commit); reply MsgHeartbeatResp */
static void raft_handleHeartbeat(raft *r, const Message *m) {
  uint64_t before = r->raftLog.committed;
  raftLog_commitTo(&r->raftLog, m->Commit, &r->errors);
  if (r->raftLog.committed != before) r->out |= MRQ_OUT_COMMIT_ADVANCED;
  reply_ack(r, m->From);
}

/* upstream stepLeader(r, m) */
static void stepLeader(raft *r, const Message *m) {
  switch (m->Type) {
    case MsgBeat:
      r->out |= MRQ_OUT_BCAST_HEARTBEAT; /* bcastHeartbeat() */
      return;
    case MsgProp:
      raft_appendEntry(r, m->nEntries);
      r->out |= MRQ_OUT_BCAST_APPEND; /* bcastAppend() */
      return;
    case MsgVote:
      reply_vote(r, m->From, 1);
      return;
  }
  Progress *pr = &r->prs[m->From - 1];
  switch (m->Type) {
    case MsgAppResp:
      if (m->Reject) {
        /* pr.maybeDecrTo only moves Next (message construction; derived in the engine) */
      } else if (Progress_maybeUpdate(pr, m->Index)) {
        if (raft_maybeCommit(r)) r->out |= MRQ_OUT_BCAST_APPEND;
      }
      break;
    case MsgHeartbeatResp: /* RecentActive / sendAppend only */
      break;
  }
}

/* upstream stepCandidate(r, m) */
static void stepCandidate(raft *r, const Message *m) {
  switch (m->Type) {
    case MsgProp:
      r->out |= MRQ_OUT_PROP_DROPPED;
      return;
    case MsgApp:
      raft_becomeFollower(r, r->Term, m->From);
      raft_handleAppendEntries(r, m);
      break;
    case MsgHeartbeat:
      raft_becomeFollower(r, r->Term, m->From);
      raft_handleHeartbeat(r, m);
      break;
    case MsgVote:
      reply_vote(r, m->From, 1);
      break;
    case MsgVoteResp: {
      int gr = raft_poll(r, m->From, !m->Reject);
      if (raft_q(r) == gr) {
        raft_becomeLeader(r);
        r->out |= MRQ_OUT_BCAST_APPEND;
      } else if (raft_q(r) == votes_len(r) - gr) {
        raft_becomeFollower(r, r->Term, ORC_None);
      }
      break;
    }
  }
}

/* upstream stepFollower(r, m) */
static void stepFollower(raft *r, const Message *m) {
  switch (m->Type) {
    case MsgProp:
      if (r->lead == ORC_None)
        r->out |= MRQ_OUT_PROP_DROPPED;
      else
        r->out |= MRQ_OUT_PROP_FORWARD;
      break;
    case MsgApp:
      r->electionElapsed = 0;
      r->lead = m->From;
      raft_handleAppendEntries(r, m);
      break;
    case MsgHeartbeat:
      r->electionElapsed = 0;
      r->lead = m->From;
      raft_handleHeartbeat(r, m);
      break;
    case MsgVote:
      if ((r->Vote == ORC_None || r->Vote == m->From) && raftLog_isUpToDate(&r->raftLog, m->Index, m->LogTerm)) {
        r->electionElapsed = 0;
        r->Vote = m->From;
        reply_vote(r, m->From, 0);
      } else {
        reply_vote(r, m->From, 1);
      }
      break;
  }
}

/* upstream raft.Step(m) */
void raft_Step(raft *r, const Message *m) {
  if (m->Type == MsgHup) {
    if (r->state != StateLeader) raft_campaign(r);
    return;
  }
  if (m->Term == 0) {
    /* local message */
  } else if (m->Term > r->Term) {
    uint64_t lead = m->From;
    if (m->Type == MsgVote) lead = ORC_None;
    raft_becomeFollower(r, m->Term, lead);
  } else if (m->Term < r->Term) {
    return; /* ignore */
  }
  switch (r->state) {
    case StateLeader: stepLeader(r, m); break;
    case StateCandidate: stepCandidate(r, m); break;
    default: stepFollower(r, m); break;
  }
}

/* upstream raft.tickElection() with v3's pastElectionTimeout(): elapsed >= randomizedElectionTimeout */
static void raft_tickElection(raft *r) {
  r->electionElapsed++;
  if (r->electionElapsed >= r->randomizedElectionTimeout) {
    r->electionElapsed = 0;
    Message hup;
    memset(&hup, 0, sizeof hup);
    hup.Type = MsgHup;
    hup.From = r->id;
    raft_Step(r, &hup);
  }
}

/* upstream raft.tickHeartbeat() (checkQuorum is off: reference raft.go:152-159 never sets it) */
static void raft_tickHeartbeat(raft *r) {
  r->heartbeatElapsed++;
  r->electionElapsed++;
  if (r->electionElapsed >= r->electionTimeout) r->electionElapsed = 0;
  if (r->state != StateLeader) return;
  if (r->heartbeatElapsed >= r->heartbeatTimeout) {
    r->heartbeatElapsed = 0;
    Message beat;
    memset(&beat, 0, sizeof beat);
    beat.Type = MsgBeat;
    beat.From = r->id;
    raft_Step(r, &beat);
  }
}

/* upstream r.tick (tickElection for followers/candidates, tickHeartbeat for leaders) */
void raft_tick(raft *r) {
  if (r->state == StateLeader)
    raft_tickHeartbeat(r);
  else
    raft_tickElection(r);
}

/* ------------------------------------------------------------------------------------------
 * test helpers
 * ---------------------------------------------------------------------------------------- */
static uint64_t g_zero_tick = 0;

static void raft_init(raft *r, uint64_t id, int npeers, int et, int ht) {
  memset(r, 0, sizeof *r);
  r->id = id;
  r->nprs = npeers;
  r->electionTimeout = et;
  r->heartbeatTimeout = ht;
  r->tick_no = &g_zero_tick;
  /* upstream newRaft(): becomeFollower(r.Term, None) */
  raft_becomeFollower(r, 0, ORC_None);
  r->out = 0;
}

raft *orc_raft_new(uint64_t id, int npeers, int election_tick, int heartbeat_tick) {
  raft *r = (raft *)malloc(sizeof(raft));
  raft_init(r, id, npeers, election_tick, heartbeat_tick);
  return r;
}

void orc_raft_free(raft *r) {
  if (!r) return;
  free(r->raftLog.runs);
  free(r);
}

void orc_raft_set_log(raft *r, const uint64_t *entry_terms, size_t n) {
  r->raftLog.nruns = 0;
  r->raftLog.lastIndex_ = 0;
  for (size_t i = 0; i < n; ++i) raftLog_append(&r->raftLog, 1, entry_terms[i]);
}

/* upstream raft_test.go TestCommit: a raft with prs[i].Match = matches[i], a log with the given
 * entry terms, HardState{Term: smTerm}; run maybeCommit(); report raftLog.committed. */
uint64_t orc_kat_commit(const uint64_t *matches, int n, const uint64_t *entry_terms, size_t nlog, uint64_t smTerm) {
  raft *r = orc_raft_new(1, n, 5, 1);
  orc_raft_set_log(r, entry_terms, nlog);
  r->Term = smTerm;
  for (int i = 0; i < n; ++i) {
    r->prs[i].Match = matches[i];
    r->prs[i].Next = matches[i] + 1;
  }
  raft_maybeCommit(r);
  uint64_t c = r->raftLog.committed;
  orc_raft_free(r);
  return c;
}

/* Independent definition (no sort): the largest value x among m[] that at least q entries reach. */
uint64_t orc_quorum_index_bruteforce(const uint64_t *m, int n) {
  int q = n / 2 + 1;
  uint64_t best = 0;
  for (int i = 0; i < n; ++i) {
    int cnt = 0;
    for (int j = 0; j < n; ++j)
      if (m[j] >= m[i]) ++cnt;
    if (cnt >= q && m[i] > best) best = m[i];
  }
  return best;
}

/* ------------------------------------------------------------------------------------------
 * multi-group driver
 * ---------------------------------------------------------------------------------------- */
struct orc_engine {
  uint64_t G, base;
  uint32_t R;
  raft *groups;
  uint64_t tick_no;
};

orc_engine *orc_create(uint64_t n_groups, uint32_t n_replicas, uint64_t group_base, uint32_t election_tick,
                       uint32_t heartbeat_tick, uint64_t seed, uint32_t self_id) {
  if (n_replicas < 1 || n_replicas > ORC_MAXR || self_id > n_replicas) return NULL;
  orc_engine *e = (orc_engine *)calloc(1, sizeof *e);
  e->G = n_groups;
  e->R = n_replicas;
  e->base = group_base;
  e->groups = (raft *)calloc(n_groups ? n_groups : 1, sizeof(raft));
  for (uint64_t g = 0; g < n_groups; ++g) {
    raft *r = &e->groups[g];
    uint64_t gg = group_base + g;
    uint64_t id = self_id ? self_id : (gg % n_replicas) + 1;
    memset(r, 0, sizeof *r);
    r->id = id;
    r->nprs = (int)n_replicas;
    r->electionTimeout = (int)election_tick;
    r->heartbeatTimeout = (int)heartbeat_tick;
    r->rand_seed = seed;
    r->rand_group = gg;
    r->tick_no = &e->tick_no;
    raft_becomeFollower(r, 0, ORC_None);
    r->out = 0;
  }
  return e;
}

void orc_destroy(orc_engine *e) {
  if (!e) return;
  for (uint64_t g = 0; g < e->G; ++g) free(e->groups[g].raftLog.runs);
  free(e->groups);
  free(e);
}

raft *orc_group(orc_engine *e, uint64_t g) { return &e->groups[g]; }

/* Step ONE message on ONE group (KAT harnesses drive the per-message functions directly). */
void orc_step(orc_engine *e, uint64_t g, int type, uint64_t from, uint64_t term, uint64_t index, uint64_t logterm,
              uint64_t commit, int reject, uint32_t n_entries) {
  Message m;
  memset(&m, 0, sizeof m);
  m.Type = type;
  m.From = from;
  m.To = e->groups[g].id;
  m.Term = term;
  m.Index = index;
  m.LogTerm = logterm;
  m.Commit = commit;
  m.Reject = reject;
  m.nEntries = n_entries;
  raft_Step(&e->groups[g], &m);
}
void orc_group_set_log(orc_engine *e, uint64_t g, const uint64_t *entry_terms, size_t n) {
  orc_raft_set_log(&e->groups[g], entry_terms, n);
}
void orc_group_clear_out(orc_engine *e, uint64_t g) { e->groups[g].out = 0; }
uint64_t orc_tick_count(const orc_engine *e) { return e->tick_no; }
void orc_set_tick_count(orc_engine *e, uint64_t t) { e->tick_no = t; }
uint64_t orc_errors(const orc_engine *e) {
  uint64_t n = 0;
  for (uint64_t g = 0; g < e->G; ++g) n += e->groups[g].errors;
  return n;
}
int orc_hw_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef struct job {
  orc_engine *e;
  uint64_t g0, g1;
  int kind; /* 0 tick, 1 quorum, 2 gen */
  const uint8_t *type;
  const uint64_t *term, *index, *logterm, *commit;
  const uint32_t *prop;
  /* gen outputs */
  uint8_t *otype;
  uint64_t *oterm, *oindex, *ologterm, *ocommit;
  uint32_t *oprop;
  const mrq_trace_params *tp;
  uint64_t tick;
} job;

/* The canonical per-tick serialisation: Step every inbox message in sender order,
 * then the tick's proposals as one MsgProp, then tick(). */
static void tick_range(const job *j) {
  orc_engine *e = j->e;
  const uint64_t G = e->G;
  for (uint64_t g = j->g0; g < j->g1; ++g) {
    raft *r = &e->groups[g];
    r->out = 0;
    for (uint32_t s = 0; s < e->R; ++s) {
      uint8_t t = j->type ? j->type[s * G + g] : 0;
      if ((t & MRQ_MSG_TYPE_MASK) == 0) continue;
      if ((uint64_t)(s + 1) == r->id) continue; /* a node does not message itself */
      Message m;
      memset(&m, 0, sizeof m);
      m.Type = t & MRQ_MSG_TYPE_MASK;
      m.Reject = (t & MRQ_MSG_REJECT) != 0;
      m.From = s + 1;
      m.To = r->id;
      m.Term = j->term ? j->term[s * G + g] : 0;
      m.Index = j->index ? j->index[s * G + g] : 0;
      m.LogTerm = j->logterm ? j->logterm[s * G + g] : 0;
      m.Commit = j->commit ? j->commit[s * G + g] : 0;
      raft_Step(r, &m);
    }
    uint32_t np = j->prop ? j->prop[g] : 0;
    if (np) {
      Message m;
      memset(&m, 0, sizeof m);
      m.Type = MsgProp;
      m.From = r->id;
      m.nEntries = np;
      raft_Step(r, &m);
    }
    raft_tick(r);
  }
}

static void quorum_range(const job *j) {
  for (uint64_t g = j->g0; g < j->g1; ++g) {
    raft *r = &j->e->groups[g];
    if (r->state == StateLeader) raft_maybeCommit(r);
  }
}

static uint32_t votes_word(const raft *r) {
  uint32_t w = 0;
  for (int i = 0; i < r->nprs; ++i) w |= (uint32_t)(r->votes[i] < 0 ? 0 : (r->votes[i] ? 1 : 2)) << (2 * i);
  return w;
}

static void gen_range(const job *j) {
  orc_engine *e = j->e;
  const uint64_t G = e->G;
  for (uint64_t g = j->g0; g < j->g1; ++g) {
    const raft *r = &e->groups[g];
    mrq_trace_view v;
    v.term = r->Term;
    v.last_index = r->raftLog.lastIndex_;
    v.last_term = raftLog_lastTerm(&r->raftLog);
    v.committed = r->raftLog.committed;
    v.role = (uint32_t)r->state;
    v.lead = (uint32_t)r->lead;
    v.self_id = (uint32_t)r->id;
    v.votes = votes_word(r);
    for (uint32_t s = 0; s < e->R; ++s) {
      mrq_trace_msg m = mrq_trace_cell(j->tp, j->tick, e->base + g, s, &v);
      j->otype[s * G + g] = m.type;
      j->oterm[s * G + g] = m.term;
      j->oindex[s * G + g] = m.index;
      j->ologterm[s * G + g] = m.logterm;
      j->ocommit[s * G + g] = m.commit;
    }
    if (j->oprop) j->oprop[g] = mrq_trace_props(j->tp, j->tick, e->base + g, &v);
  }
}

static void *job_main(void *p) {
  const job *j = (const job *)p;
  if (j->kind == 0)
    tick_range(j);
  else if (j->kind == 1)
    quorum_range(j);
  else
    gen_range(j);
  return NULL;
}

static void run_jobs(job *proto, int nthreads) {
  orc_engine *e = proto->e;
  if (nthreads <= 1 || e->G < (uint64_t)nthreads * 64) {
    proto->g0 = 0;
    proto->g1 = e->G;
    job_main(proto);
    return;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  job *jobs = (job *)malloc(sizeof(job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = *proto;
    jobs[t].g0 = e->G * (uint64_t)t / (uint64_t)nthreads;
    jobs[t].g1 = e->G * (uint64_t)(t + 1) / (uint64_t)nthreads;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(jobs);
  free(th);
}

void orc_tick(orc_engine *e, const uint8_t *type, const uint64_t *term, const uint64_t *index, const uint64_t *logterm,
              const uint64_t *commit, const uint32_t *prop_count, int nthreads) {
  job j;
  memset(&j, 0, sizeof j);
  j.e = e;
  j.kind = 0;
  j.type = type;
  j.term = term;
  j.index = index;
  j.logterm = logterm;
  j.commit = commit;
  j.prop = prop_count;
  run_jobs(&j, nthreads);
  e->tick_no++;
}

void orc_quorum_commit(orc_engine *e, int nthreads) {
  job j;
  memset(&j, 0, sizeof j);
  j.e = e;
  j.kind = 1;
  run_jobs(&j, nthreads);
}

void orc_gen_trace(orc_engine *e, const mrq_trace_params *p, uint64_t tick, uint8_t *type, uint64_t *term,
                   uint64_t *index, uint64_t *logterm, uint64_t *commit, uint32_t *prop_count, int nthreads) {
  job j;
  memset(&j, 0, sizeof j);
  j.e = e;
  j.kind = 2;
  j.tp = p;
  j.tick = tick;
  j.otype = type;
  j.oterm = term;
  j.oindex = index;
  j.ologterm = logterm;
  j.ocommit = commit;
  j.oprop = prop_count;
  run_jobs(&j, nthreads);
}

void orc_export(orc_engine *e, uint64_t *term, uint64_t *vote, uint64_t *committed, uint64_t *last_index,
                uint64_t *last_term, uint64_t *term_start, uint64_t *match, uint8_t *role, uint8_t *lead,
                uint8_t *self_id, uint8_t *votes, uint16_t *election_elapsed, uint16_t *heartbeat_elapsed,
                uint16_t *randomized_timeout, uint32_t *out) {
  const uint64_t G = e->G;
  for (uint64_t g = 0; g < G; ++g) {
    const raft *r = &e->groups[g];
    if (term) term[g] = r->Term;
    if (vote) vote[g] = r->Vote;
    if (committed) committed[g] = r->raftLog.committed;
    if (last_index) last_index[g] = r->raftLog.lastIndex_;
    if (last_term) last_term[g] = raftLog_lastTerm(&r->raftLog);
    if (term_start) {
      /* first index whose entry carries the leader's own term: the index becomeLeader appended at */
      uint64_t ts = UINT64_MAX;
      if (r->state == StateLeader) {
        for (int k = r->raftLog.nruns - 1; k >= 0; --k)
          if (r->raftLog.runs[k].term == r->Term) ts = r->raftLog.runs[k].first;
      }
      term_start[g] = ts;
    }
    for (uint32_t s = 0; s < e->R; ++s) {
      if (match) match[s * G + g] = r->prs[s].Match;
      if (votes) votes[s * G + g] = (uint8_t)(r->votes[s] < 0 ? 0 : (r->votes[s] ? 1 : 2));
    }
    if (role) role[g] = (uint8_t)r->state;
    if (lead) lead[g] = (uint8_t)r->lead;
    if (self_id) self_id[g] = (uint8_t)r->id;
    if (election_elapsed) election_elapsed[g] = (uint16_t)r->electionElapsed;
    if (heartbeat_elapsed) heartbeat_elapsed[g] = (uint16_t)r->heartbeatElapsed;
    if (randomized_timeout) randomized_timeout[g] = (uint16_t)r->randomizedElectionTimeout;
    if (out) out[g] = r->out;
  }
}

void orc_import(orc_engine *e, const uint64_t *term, const uint64_t *vote, const uint64_t *committed,
                const uint64_t *last_index, const uint64_t *last_term, const uint64_t *term_start,
                const uint64_t *match, const uint8_t *role, const uint8_t *lead, const uint8_t *self_id,
                const uint8_t *votes, const uint16_t *election_elapsed, const uint16_t *heartbeat_elapsed,
                const uint16_t *randomized_timeout) {
  const uint64_t G = e->G;
  for (uint64_t g = 0; g < G; ++g) {
    raft *r = &e->groups[g];
    if (term) r->Term = term[g];
    if (vote) r->Vote = vote[g];
    if (role) r->state = role[g];
    if (lead) r->lead = lead[g];
    if (self_id) r->id = self_id[g];
    if (last_index) {
      /* rebuild a run-length log consistent with (last_index, last_term[, term_start]) */
      uint64_t li = last_index[g];
      uint64_t lt = last_term ? last_term[g] : r->Term;
      r->raftLog.nruns = 0;
      r->raftLog.lastIndex_ = 0;
      uint64_t ts = term_start ? term_start[g] : UINT64_MAX;
      if (r->state == StateLeader && ts != UINT64_MAX && ts >= 1 && ts <= li) {
        if (ts > 1) { /* entries before the leader's own term carry some smaller term */
          log_push_run(&r->raftLog, 1, r->Term > 0 ? r->Term - 1 : 0);
          r->raftLog.lastIndex_ = ts - 1;
        }
        log_push_run(&r->raftLog, ts, r->Term);
        r->raftLog.lastIndex_ = li;
      } else if (li > 0) {
        log_push_run(&r->raftLog, 1, lt);
        r->raftLog.lastIndex_ = li;
      }
    }
    if (committed) r->raftLog.committed = committed[g];
    for (uint32_t s = 0; s < e->R; ++s) {
      if (match) {
        r->prs[s].Match = match[s * G + g];
        r->prs[s].Next = match[s * G + g] + 1;
      }
      if (votes) {
        uint8_t v = votes[s * G + g];
        r->votes[s] = (int8_t)(v == 0 ? -1 : (v == 1 ? 1 : 0));
      }
    }
    if (election_elapsed) r->electionElapsed = election_elapsed[g];
    if (heartbeat_elapsed) r->heartbeatElapsed = heartbeat_elapsed[g];
    if (randomized_timeout) r->randomizedElectionTimeout = randomized_timeout[g];
  }
}
