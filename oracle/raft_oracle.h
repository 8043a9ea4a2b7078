/*
 * raft_oracle.h — CPU oracle for the multi-raft quorum hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the arithmetic of this path is not in /root/reference; it is the un-vendored,
 * un-pinned third-party module github.com/coreos/etcd/raft (reference raft.go:30, era v2.2–v2.3,
 * SURVEY §8c) and the reference's own tests assert no term / vote / commit value
 * (raftsql_test.go:109-111,144,155,167).  This file restates upstream's published per-message
 * algorithm from knowledge of that project, one function per upstream symbol, and is pinned by
 * (a) the upstream known-answer tables recalled in SURVEY §8c (TestCommit, TestVoter,
 * TestLeaderElection) and (b) independent-definition property tests (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (raftsql_b200/) never does.
 *
 * Style: array-of-structs, one group at a time, one message at a time, maybeCommit = copy the R
 * match values, sort descending, take element q-1 (upstream raft.go maybeCommit) — deliberately
 * the reference's shape, not a tuned CPU kernel.
 */
#ifndef RAFT_ORACLE_H
#define RAFT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXR 8
#define ORC_None 0ull

enum { StateFollower = 0, StateCandidate = 1, StateLeader = 2 };

/* raftpb.MessageType (v2.2/2.3 numbering) */
enum {
  MsgHup = 0, MsgBeat = 1, MsgProp = 2, MsgApp = 3, MsgAppResp = 4, MsgVote = 5, MsgVoteResp = 6,
  MsgSnap = 7, MsgHeartbeat = 8, MsgHeartbeatResp = 9
};

typedef struct Message {   /* raftpb.Message, the fields this path reads */
  int Type;
  uint64_t To, From, Term, LogTerm, Index, Commit;
  int Reject;
  uint32_t nEntries;       /* len(m.Entries) for MsgProp */
} Message;

typedef struct Progress {  /* upstream raft/progress.go */
  uint64_t Match, Next;
} Progress;

typedef struct TermRun {   /* entries [first, next run's first) carry `term` */
  uint64_t first, term;
} TermRun;

typedef struct raftLog {   /* upstream raft/log.go, payload-free: a run-length list of entry terms */
  uint64_t lastIndex_, committed;
  TermRun *runs;
  int nruns, cap;
} raftLog;

typedef struct raft {      /* upstream raft/raft.go `type raft struct` */
  uint64_t id, Term, Vote, lead;
  int state;
  int nprs;                          /* len(r.prs) = R; peer ids are 1..R (reference raft.go:150) */
  Progress prs[ORC_MAXR];            /* prs[id-1] */
  int8_t votes[ORC_MAXR];            /* r.votes map: -1 absent, 0 false, 1 true */
  raftLog raftLog;
  int electionElapsed, heartbeatElapsed;
  int electionTimeout, heartbeatTimeout, randomizedElectionTimeout;
  /* stand-in for r.rand (see include/mrq_trace.h mrq_randomized_timeout) */
  uint64_t rand_seed, rand_group;
  const uint64_t *tick_no;
  /* stand-in for r.msgs: the MRQ_OUT_* word of include/mrq.h */
  uint32_t out;
  uint32_t errors;
} raft;

/* ---- upstream functions, restated (each cites its upstream symbol in the .c file) ---- */
int raft_q(const raft *r);
int raft_poll(raft *r, uint64_t id, int v);
void raft_reset(raft *r, uint64_t term);
void raft_becomeFollower(raft *r, uint64_t term, uint64_t lead);
void raft_becomeCandidate(raft *r);
void raft_becomeLeader(raft *r);
void raft_campaign(raft *r);
void raft_appendEntry(raft *r, uint32_t n);
int raft_maybeCommit(raft *r);
void raft_Step(raft *r, const Message *m);
void raft_tick(raft *r);
int Progress_maybeUpdate(Progress *pr, uint64_t n);
uint64_t raftLog_term(const raftLog *l, uint64_t i);
uint64_t raftLog_lastTerm(const raftLog *l);
int raftLog_isUpToDate(const raftLog *l, uint64_t lasti, uint64_t term);
int raftLog_maybeCommit(raftLog *l, uint64_t maxIndex, uint64_t term);
void raftLog_commitTo(raftLog *l, uint64_t tocommit, uint32_t *errors);
void raftLog_append(raftLog *l, uint32_t n, uint64_t term);

/* ---- test helpers ------------------------------------------------------------------ */
/* Build a lone raft with an explicit log (terms of entries 1..n) — the KAT form. */
raft *orc_raft_new(uint64_t id, int npeers, int election_tick, int heartbeat_tick);
void orc_raft_free(raft *r);
void orc_raft_set_log(raft *r, const uint64_t *entry_terms, size_t n);
/* upstream TestCommit form: matches[], log terms, leader term -> commit index */
uint64_t orc_kat_commit(const uint64_t *matches, int n, const uint64_t *entry_terms, size_t nlog, uint64_t smTerm);
/* independent definition used by property tests: max{x in m : |{r: m[r] >= x}| >= q}, 0 if none */
uint64_t orc_quorum_index_bruteforce(const uint64_t *m, int n);

/* ---- the multi-group driver (same shape as the engine's C-ABI) ------------------------- */
typedef struct orc_engine orc_engine;

orc_engine *orc_create(uint64_t n_groups, uint32_t n_replicas, uint64_t group_base, uint32_t election_tick,
                       uint32_t heartbeat_tick, uint64_t seed, uint32_t self_id);
void orc_destroy(orc_engine *e);
raft *orc_group(orc_engine *e, uint64_t g);
void orc_step(orc_engine *e, uint64_t g, int type, uint64_t from, uint64_t term, uint64_t index, uint64_t logterm,
              uint64_t commit, int reject, uint32_t n_entries);
void orc_group_set_log(orc_engine *e, uint64_t g, const uint64_t *entry_terms, size_t n);
void orc_group_clear_out(orc_engine *e, uint64_t g);

/* One tick over all groups with a dense inbox laid out as include/mrq.h's mrq_inbox (replica-major).
 * nthreads <= 1: sequential; otherwise a static contiguous partition over pthreads.           */
void orc_tick(orc_engine *e, const uint8_t *type, const uint64_t *term, const uint64_t *index, const uint64_t *logterm,
              const uint64_t *commit, const uint32_t *prop_count, int nthreads);
/* maybeCommit() on every leader group (the standalone quorum pass). */
void orc_quorum_commit(orc_engine *e, int nthreads);
/* Generate the synthetic trace of include/mrq_trace.h for the current state into caller arrays. */
struct mrq_trace_params;
void orc_gen_trace(orc_engine *e, const struct mrq_trace_params *p, uint64_t tick, uint8_t *type, uint64_t *term,
                   uint64_t *index, uint64_t *logterm, uint64_t *commit, uint32_t *prop_count, int nthreads);

/* SoA export/import matching mrq_state (NULL columns skipped). */
void orc_export(orc_engine *e, uint64_t *term, uint64_t *vote, uint64_t *committed, uint64_t *last_index,
                uint64_t *last_term, uint64_t *term_start, uint64_t *match, uint8_t *role, uint8_t *lead,
                uint8_t *self_id, uint8_t *votes, uint16_t *election_elapsed, uint16_t *heartbeat_elapsed,
                uint16_t *randomized_timeout, uint32_t *out);
void orc_import(orc_engine *e, const uint64_t *term, const uint64_t *vote, const uint64_t *committed,
                const uint64_t *last_index, const uint64_t *last_term, const uint64_t *term_start,
                const uint64_t *match, const uint8_t *role, const uint8_t *lead, const uint8_t *self_id,
                const uint8_t *votes, const uint16_t *election_elapsed, const uint16_t *heartbeat_elapsed,
                const uint16_t *randomized_timeout);
uint64_t orc_tick_count(const orc_engine *e);
void orc_set_tick_count(orc_engine *e, uint64_t t);
uint64_t orc_errors(const orc_engine *e);
int orc_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
