/*
 * mrq_trace.h — the synthetic vote/append trace model (workload definition).
 *
 * BASELINE.json's configs ask for "synthetic vote/append traces" of a named (groups x replicas)
 * shape.  The reference has no trace generator (no benchmark of any kind), so the
 * trace is defined here, once, as plain C99 that also compiles as CUDA: the device generator
 * (mrq_gen_trace), the CPU baseline driver and the tests all include this file, so every side
 * sees byte-identical inputs.  It is an INPUT model — it contains none of the raft arithmetic
 * under test (that lives in raftsql_b200/csrc/ and, independently, in oracle/).
 *
 * The generator plays "the other R-1 replicas of group g" as seen from this node: given this
 * node's state of group g at the start of tick t it emits at most one message per peer,
 * drawn from a counter-based RNG keyed by (seed, tick, global group id, peer slot).  No hidden
 * PRNG stream: any (tick, g, r) cell can be generated independently, on any device.
 */
#ifndef MRQ_TRACE_H
#define MRQ_TRACE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MRQ_HD __host__ __device__ __forceinline__
#else
#define MRQ_HD static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Trace parameters (all probabilities are integer thresholds so host and device agree exactly). */
typedef struct mrq_trace_params {
  uint64_t seed;
  uint32_t p_ack_256;      /* leader: a healthy follower acks this tick w.p. p/256                */
  uint32_t p_grant_256;    /* candidate: an undecided voter grants w.p. p/256                     */
  uint32_t p_reject_256;   /* candidate: ... rejects w.p. p/256 (else stays silent this tick)     */
  uint32_t p_heartbeat_256;/* follower with a leader: that leader heartbeats w.p. p/256           */
  uint32_t churn_65536;    /* any role: a peer shows up with a higher term w.p. c/65536 per tick  */
  uint32_t lagging_pct;    /* static share (percent) of (group, peer) pairs that are lagging      */
  uint32_t max_prop;       /* leader: proposals this tick uniform in 0..max_prop                  */
  uint32_t lag_kind;       /* 0: lag in {0,1,2} w.p. {.7,.2,.1}; 1: geometric, mean ~4            */
} mrq_trace_params;

/* One generated inbox cell. */
typedef struct mrq_trace_msg {
  uint64_t term, index, logterm, commit;
  uint8_t type; /* MRQ_MSG_* | MRQ_MSG_REJECT; 0 = empty */
} mrq_trace_msg;

/* What the generator may look at: this node's state of the group at tick start. */
typedef struct mrq_trace_view {
  uint64_t term, last_index, last_term, committed;
  uint32_t role;    /* MRQ_ROLE_* */
  uint32_t lead;    /* 0 = None */
  uint32_t self_id; /* 1..R */
  uint32_t votes;   /* 2 bits per slot r: 0 absent, 1 granted, 2 rejected */
} mrq_trace_view;

/* splitmix64 finaliser */
MRQ_HD uint64_t mrq_mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

/* counter-based draw keyed by (seed, a, b, c) */
MRQ_HD uint64_t mrq_rand4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t x = mrq_mix64(seed + 0x9E3779B97F4A7C15ull * (a + 1));
  x = mrq_mix64(x ^ (0xD1B54A32D192ED03ull * (b + 1)));
  x = mrq_mix64(x ^ (0x8CB92BA72F3D8DD7ull * (c + 1)));
  return x;
}

/* The engine's stand-in for etcd-raft's r.rand (a Go math/rand stream that cannot be restated):
 * the randomized election timeout drawn by reset() during tick `tick_no` for global group `gg`.
 * Range [election_tick, 2*election_tick - 1] as upstream.                                      */
MRQ_HD uint32_t mrq_randomized_timeout(uint64_t seed, uint64_t gg, uint64_t tick_no, uint32_t election_tick) {
  return election_tick + (uint32_t)(mrq_rand4(seed, gg, tick_no, 0x7133u) % election_tick);
}

/* Number of proposals arriving at this node for global group gg during tick `tick` (only the
 * leader's count matters; others are dropped/forwarded and flagged).                          */
MRQ_HD uint32_t mrq_trace_props(const mrq_trace_params *p, uint64_t tick, uint64_t gg, const mrq_trace_view *v) {
  if (p->max_prop == 0 || v->role != 2u) return 0;
  return (uint32_t)(mrq_rand4(p->seed, tick, gg, 0xF00Du) % (uint64_t)(p->max_prop + 1u));
}

MRQ_HD uint64_t mrq_trace_lag(const mrq_trace_params *p, uint64_t u) {
  if (p->lag_kind == 0) {
    uint32_t x = (uint32_t)(u % 10u);
    return x < 7 ? 0 : (x < 9 ? 1 : 2);
  }
  /* geometric with success prob 1/5 => mean 4: count leading "failures" of base-5 digits */
  uint64_t lag = 0;
  while (lag < 24 && (u % 5u) != 0) {
    u /= 5u;
    ++lag;
  }
  return lag;
}

/* The message peer slot r (id r+1) sends to this node for global group gg in tick `tick`. */
MRQ_HD mrq_trace_msg mrq_trace_cell(const mrq_trace_params *p, uint64_t tick, uint64_t gg, uint32_t r,
                                    const mrq_trace_view *v) {
  mrq_trace_msg m;
  m.term = m.index = m.logterm = m.commit = 0;
  m.type = 0;
  const uint32_t id = r + 1u;
  if (id == v->self_id) return m;
  const uint64_t u = mrq_rand4(p->seed, tick, gg, 0x100u + r);

  /* leader churn: someone has moved on to a higher term (configs[4]) */
  if ((uint32_t)(u & 0xFFFFu) < p->churn_65536) {
    m.term = v->term + 1u + ((u >> 16) & 1u);
    if ((u >> 17) & 1u) {
      m.type = 8u; /* MsgHeartbeat from a new leader */
      m.commit = v->committed;
    } else {
      m.type = 5u; /* MsgVote from a new candidate; its log is ahead, level, or behind ours */
      uint64_t d = (u >> 20) % 3u;
      m.index = (v->last_index + d > 0) ? v->last_index + d - 1u : 0;
      m.logterm = v->last_term;
    }
    return m;
  }
  const uint32_t c = (uint32_t)((u >> 24) & 0xFFu);   /* main coin */
  const uint64_t w = u >> 32;                          /* payload entropy */
  if (v->role == 2u) { /* leader: followers acknowledge appends */
    const uint64_t h = mrq_rand4(p->seed, 0xA11CEull, gg, r); /* static per (group, peer) */
    uint64_t lag;
    if ((uint32_t)(h % 100u) < p->lagging_pct) {
      if (((tick + (h >> 8)) & 15u) != 0) return m; /* acks only every 16th tick */
      lag = 64u + (w % 4033u);                      /* stale match, 64..4096     */
    } else {
      if (c >= p->p_ack_256) return m;
      lag = mrq_trace_lag(p, w);
    }
    m.type = 4u; /* MsgAppResp */
    m.term = v->term;
    m.index = v->last_index > lag ? v->last_index - lag : 0;
    return m;
  }
  if (v->role == 1u) { /* candidate: voters answer once */
    if (((v->votes >> (2u * r)) & 3u) != 0) return m;
    if (c < p->p_grant_256) {
      m.type = 6u; /* MsgVoteResp granted */
      m.term = v->term;
    } else if (c < p->p_grant_256 + p->p_reject_256) {
      m.type = 6u | 0x80u; /* MsgVoteResp rejected */
      m.term = v->term;
    }
    return m;
  }
  /* follower: its leader (if any) heartbeats, carrying a commit index at or a little past ours */
  if (v->lead == id && c < p->p_heartbeat_256) {
    m.type = 8u;
    m.term = v->term;
    uint64_t cm = v->committed + (w % 3u);
    m.commit = cm < v->last_index ? cm : v->last_index;
  }
  return m;
}

/* Presets for BASELINE.json's configs. */
MRQ_HD mrq_trace_params mrq_trace_preset(uint32_t config_no) {
  mrq_trace_params p;
  p.seed = 0x5EED0000ull + config_no;
  p.p_ack_256 = 256;
  p.p_grant_256 = 230;  /* ~0.9 */
  p.p_reject_256 = 0;
  p.p_heartbeat_256 = 0;
  p.churn_65536 = 0;
  p.lagging_pct = 0;
  p.max_prop = 3;
  p.lag_kind = 0;
  if (config_no == 3u || config_no == 4u) { /* 1,048,576 x 5 steady state */
    p.lag_kind = 1;
  } else if (config_no == 5u) { /* 262,144 x 7, 20% lagging followers + leader churn */
    p.p_grant_256 = 205;  /* 0.8 */
    p.p_reject_256 = 26;  /* 0.1 */
    p.churn_65536 = 43;   /* per peer slot; ~1/256 per group per tick at R=7 */
    p.lagging_pct = 20;
    p.lag_kind = 1;
  }
  return p;
}

#ifdef __cplusplus
}
#endif
#endif /* MRQ_TRACE_H */
