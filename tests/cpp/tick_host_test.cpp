// tick_host_test.cpp — the DEVICE source of the tick, compiled for the host and run against the oracle.
//
// raftsql_b200/csrc/mrq_kernels.cuh is included as it is (tests/cpp/device_on_host.hpp supplies the CUDA vocabulary;
// MRQ_HOST_EMULATION turns the cache-hinted PTX accessors into plain loads/stores and the borrow-chain asm of
// delta32 into the same arithmetic in C).  What runs here is, line for line, what one GPU thread runs for one group:
//
//     fast_group_tick<R>(args, g, slow, ev);   if (slow) general_group_tick<R>(args, g);
//
// over host arrays in the engine's layout ([R][gs] replica-major columns, packed meta word), for every group, every
// tick, on the synthetic traces — and after every tick all state columns and the out word must equal the oracle's.
// The byte-form inbox goes the same way: unpack8_group — the body of unpack8_inbox_kernel — decodes frames built
// with the shared inline codec, the ticks run on what it wrote, and the device's sliding window must stay in step
// with the frame builder's (run_case8).
// This does not replace the GPU parity tests (it cannot see launch geometry, warp collectives or memory ordering);
// it lets the arithmetic of BOTH tick paths be checked on a machine without a GPU, which is where device code is
// written between GPU sessions.
//
//   tick_host_test                 all cases;   exit code 0 and "tick_host_test: ok" on success
#include "device_on_host.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../raftsql_b200/csrc/mrq_kernels.cuh"
#include "../../oracle/raft_oracle.h"

using namespace mrq;

namespace {

int failures = 0;

struct HostEngine {  // the engine's state and one inbox slot, in host memory, engine layout
  uint64_t G, gs;
  uint32_t R;
  std::vector<uint64_t> term, meta, last_index, last_term, committed, term_start, match;
  std::vector<uint32_t> out;
  std::vector<uint8_t> itype;
  std::vector<uint64_t> iterm, iindex, ilogterm, icommit;
  std::vector<uint32_t> iprop;
  uint64_t tick_no[2] = {0, 0};
  uint64_t fast_groups = 0, slow_groups = 0;

  HostEngine(uint64_t G_, uint32_t R_) : G(G_), gs(((G_ + 127) / 128) * 128), R(R_) {
    for (auto *v : {&term, &meta, &last_index, &last_term, &committed, &term_start}) v->assign(gs, 0);
    match.assign(gs * R, 0);
    out.assign(gs, 0);
    itype.assign(gs * R, 0);
    for (auto *v : {&iterm, &iindex, &ilogterm, &icommit}) v->assign(gs * R, 0);
    iprop.assign(gs, 0);
  }

  TickArgs args(uint64_t group_base, uint64_t seed, uint32_t et, uint32_t ht, bool with_inbox) {
    TickArgs a{};
    a.s = StateView{term.data(), meta.data(), last_index.data(), last_term.data(), committed.data(), term_start.data(),
                    match.data(), out.data()};
    if (with_inbox) a.in = InboxView{itype.data(), iterm.data(), iindex.data(), ilogterm.data(), icommit.data(), iprop.data()};
    a.G = G;
    a.gs = gs;
    a.group_base = group_base;
    a.seed = seed;
    a.tick_cur = &tick_no[0];
    a.tick_next = &tick_no[1];
    a.election_tick = et;
    a.heartbeat_tick = ht;
    a.world = 1;
    a.l2_policy = 1;
    return a;
  }
};

// what tick_fast_kernel + tick_slow_kernel do, one "thread" at a time (mode 0), or tick_general_kernel (mode 1)
template <int R>
void host_tick(HostEngine &e, const TickArgs &a, int mode) {
  for (uint64_t i = 0; i < e.G; ++i) {
    if (mode == 1) {
      general_group_tick<R>(a, i);
      ++e.slow_groups;
      continue;
    }
    bool slow = false;
    uint32_t ev = 0;
    fast_group_tick<R>(a, i, slow, ev);
    if (slow) {
      general_group_tick<R>(a, i);
      ++e.slow_groups;
    } else {
      ++e.fast_groups;
    }
  }
  e.tick_no[0] += 1;
}

void dispatch_tick(HostEngine &e, const TickArgs &a, int mode) {
  switch (e.R) {
    case 1: host_tick<1>(e, a, mode); break;
    case 2: host_tick<2>(e, a, mode); break;
    case 3: host_tick<3>(e, a, mode); break;
    case 4: host_tick<4>(e, a, mode); break;
    case 5: host_tick<5>(e, a, mode); break;
    case 6: host_tick<6>(e, a, mode); break;
    case 7: host_tick<7>(e, a, mode); break;
    case 8: host_tick<8>(e, a, mode); break;
  }
}

struct OracleCols {
  std::vector<uint64_t> term, vote, committed, last_index, last_term, term_start, match;
  std::vector<uint8_t> role, lead, self_id, votes;
  std::vector<uint16_t> el, hb, rto;
  std::vector<uint32_t> out;
  OracleCols(uint64_t G, uint32_t R)
      : term(G), vote(G), committed(G), last_index(G), last_term(G), term_start(G), match(G * R), role(G), lead(G), self_id(G),
        votes(G * R), el(G), hb(G), rto(G), out(G) {}
  void load(orc_engine *o) {
    orc_export(o, term.data(), vote.data(), committed.data(), last_index.data(), last_term.data(), term_start.data(), match.data(),
               role.data(), lead.data(), self_id.data(), votes.data(), el.data(), hb.data(), rto.data(), out.data());
  }
};

// oracle state -> engine layout (what mrq_import_state + pack_meta_kernel + fix_strict_kernel do)
void import_from_oracle(HostEngine &e, const OracleCols &c) {
  for (uint64_t g = 0; g < e.G; ++g) {
    e.term[g] = c.term[g];
    e.last_index[g] = c.last_index[g];
    e.last_term[g] = c.last_term[g];
    e.committed[g] = c.committed[g];
    e.term_start[g] = c.term_start[g];
    Meta m{};
    m.role = c.role[g];
    m.lead = c.lead[g];
    m.vote = (uint32_t)c.vote[g];
    m.self = c.self_id[g];
    m.elapsed = c.el[g];
    m.rto = c.rto[g];
    m.hb = c.hb[g];
    m.votes = 0;
    m.strict = 0;
    for (uint32_t r = 0; r < e.R; ++r) {
      e.match[(uint64_t)r * e.gs + g] = c.match[(uint64_t)r * e.G + g];
      m.votes |= (uint32_t)(c.votes[(uint64_t)r * e.G + g] & 3u) << (2 * r);
      if (c.match[(uint64_t)r * e.G + g] > c.last_index[g]) m.strict = 1;
    }
    m.ltok = c.last_term[g] == c.term[g] ? 1u : 0u;
    e.meta[g] = meta_pack(m);
  }
}

bool compare(const HostEngine &e, const OracleCols &c, const char *where, uint64_t tick) {
  for (uint64_t g = 0; g < e.G; ++g) {
    const Meta m = meta_unpack(e.meta[g]);
    const uint64_t lt = m.ltok ? e.term[g] : e.last_term[g];
    bool ok = e.term[g] == c.term[g] && m.vote == c.vote[g] && e.committed[g] == c.committed[g] &&
              e.last_index[g] == c.last_index[g] && lt == c.last_term[g] && m.role == c.role[g] && m.lead == c.lead[g] &&
              m.elapsed == c.el[g] && m.hb == c.hb[g] && m.rto == c.rto[g] && e.out[g] == c.out[g];
    for (uint32_t r = 0; ok && r < e.R; ++r) {
      ok = ((m.votes >> (2 * r)) & 3u) == c.votes[(uint64_t)r * e.G + g];
      if (ok && c.role[g] == MRQ_ROLE_LEADER) ok = e.match[(uint64_t)r * e.gs + g] == c.match[(uint64_t)r * e.G + g];
    }
    if (ok && c.role[g] == MRQ_ROLE_LEADER) ok = e.term_start[g] == c.term_start[g];
    if (!ok) {
      std::printf("FAIL %s tick %llu group %llu: term %llu/%llu commit %llu/%llu li %llu/%llu role %u/%u lead %u/%u el %u/%u rto %u/%u "
                  "out %08x/%08x vote %u/%llu lt %llu/%llu hb %u/%u ltok %u\n",
                  where, (unsigned long long)tick, (unsigned long long)g, (unsigned long long)e.term[g], (unsigned long long)c.term[g],
                  (unsigned long long)e.committed[g], (unsigned long long)c.committed[g], (unsigned long long)e.last_index[g],
                  (unsigned long long)c.last_index[g], m.role, c.role[g], m.lead, c.lead[g], m.elapsed, c.el[g], m.rto, c.rto[g],
                  e.out[g], c.out[g], m.vote, (unsigned long long)c.vote[g], (unsigned long long)lt, (unsigned long long)c.last_term[g], m.hb,
                  c.hb[g], m.ltok);
      std::printf("     term_start %llu/%llu strict %u self %u; match", (unsigned long long)e.term_start[g],
                  (unsigned long long)c.term_start[g], m.strict, m.self);
      for (uint32_t r = 0; r < e.R; ++r)
        std::printf(" %llu/%llu", (unsigned long long)e.match[(uint64_t)r * e.gs + g], (unsigned long long)c.match[(uint64_t)r * e.G + g]);
      std::printf("; votes");
      for (uint32_t r = 0; r < e.R; ++r) std::printf(" %u/%u", (m.votes >> (2 * r)) & 3u, c.votes[(uint64_t)r * e.G + g]);
      std::printf("\n");
      ++failures;
      return false;
    }
  }
  return true;
}

mrq_trace_params preset(uint32_t cfg) {
  mrq_trace_params p = mrq_trace_preset(cfg);
  if (cfg == 6) {  // follower heavy: deposed by heartbeats, then heart-beaten (tests/test_oracle_vs_pymodel.py)
    p = mrq_trace_preset(5);
    p.seed = 0x5EED0006ull;
    p.p_grant_256 = 150;
    p.p_reject_256 = 60;
    p.churn_65536 = 900;
    p.p_heartbeat_256 = 235;
    p.lagging_pct = 0;
  }
  return p;
}

void run_case(uint64_t G, uint32_t R, uint32_t cfg, int T, int mode, uint32_t self_id, uint64_t group_base, uint32_t et = 10,
              uint32_t ht = 1) {
  char where[96];
  std::snprintf(where, sizeof where, "G=%llu R=%u cfg=%u mode=%d%s", (unsigned long long)G, R, cfg, mode,
                (et != 10 || ht != 1) ? " timers" : "");
  const uint64_t seed = 0x5EED0000ull + cfg * 131 + R;
  orc_engine *o = orc_create(G, R, group_base, et, ht, seed, self_id);
  HostEngine e(G, R);
  OracleCols c(G, R);
  c.load(o);
  import_from_oracle(e, c);
  const mrq_trace_params p = preset(cfg);
  std::vector<uint8_t> type(G * R);
  std::vector<uint64_t> term(G * R), index(G * R), logterm(G * R), commit(G * R);
  std::vector<uint32_t> prop(G);
  for (int t = 0; t < T; ++t) {
    orc_gen_trace(o, &p, (uint64_t)t, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    const bool idle = (t % 17) == 16;  // every so often an idle tick (no inbox at all: timers only)
    if (!idle) {
      for (uint32_t r = 0; r < R; ++r) {  // dense [R][G] -> engine [R][gs]
        std::memcpy(&e.itype[(uint64_t)r * e.gs], &type[(uint64_t)r * G], G);
        std::memcpy(&e.iterm[(uint64_t)r * e.gs], &term[(uint64_t)r * G], G * 8);
        std::memcpy(&e.iindex[(uint64_t)r * e.gs], &index[(uint64_t)r * G], G * 8);
        std::memcpy(&e.ilogterm[(uint64_t)r * e.gs], &logterm[(uint64_t)r * G], G * 8);
        std::memcpy(&e.icommit[(uint64_t)r * e.gs], &commit[(uint64_t)r * G], G * 8);
      }
      std::memcpy(e.iprop.data(), prop.data(), G * 4);
      orc_tick(o, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    } else {
      orc_tick(o, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
    }
    const TickArgs a = e.args(group_base, seed, et, ht, !idle);
    dispatch_tick(e, a, mode);
    c.load(o);
    if (!compare(e, c, where, (uint64_t)t)) break;
  }
  if (orc_errors(o) != 0) {
    std::printf("FAIL %s: the oracle counted %llu errors\n", where, (unsigned long long)orc_errors(o));
    ++failures;
  }
  std::printf("  %-34s %4d ticks  fast %9llu  general %9llu\n", where, T, (unsigned long long)e.fast_groups,
              (unsigned long long)e.slow_groups);
  orc_destroy(o);
}

// The byte form of the packed inbox (include/mrq_packed8.h) through the DEVICE decode: every tick's wide inbox is
// encoded to R-1 bytes per group with the shared inline codec (escapes to a wide list), decoded into the engine's
// inbox columns by unpack8_group — the body of unpack8_inbox_kernel — the escapes scattered over it as
// scatter_msgs_kernel does, and the tick functions run on the result; state must still equal the oracle's, which
// was fed the wide inbox, and the device's sliding window must stay in step with the encoder's.
//
// direct = true: no unpack pass at all — fast_group_tick8 reads the bytes itself and general_group_tick8 materialises
// only the groups the fast function declines (the design for the tick on the byte form; not launched by the library yet).
template <int R>
void host_tick8(HostEngine &e, const TickArgs &a, const Inbox8 &b) {
  for (uint64_t i = 0; i < e.G; ++i) {
    bool slow = false;
    uint32_t ev = 0;
    fast_group_tick8<R>(a, b, i, slow, ev);
    if (slow) {
      general_group_tick8<R>(a, b, i);
      ++e.slow_groups;
    } else {
      ++e.fast_groups;
    }
  }
  e.tick_no[0] += 1;
}

void dispatch_tick8(HostEngine &e, const TickArgs &a, const Inbox8 &b) {
  switch (e.R) {
    case 1: host_tick8<1>(e, a, b); break;
    case 2: host_tick8<2>(e, a, b); break;
    case 3: host_tick8<3>(e, a, b); break;
    case 4: host_tick8<4>(e, a, b); break;
    case 5: host_tick8<5>(e, a, b); break;
    case 6: host_tick8<6>(e, a, b); break;
    case 7: host_tick8<7>(e, a, b); break;
    case 8: host_tick8<8>(e, a, b); break;
  }
}

void run_case8(uint64_t G, uint32_t R, uint32_t cfg, int T, int rebase_every, int also_rebase_at = -1, bool direct = false) {
  char where[96];
  std::snprintf(where, sizeof where, "byte form%s G=%llu R=%u cfg=%u", direct ? " direct" : "", (unsigned long long)G, R, cfg);
  const uint64_t seed = 0x5EED8000ull + cfg * 131 + R;
  orc_engine *o = orc_create(G, R, 0, 10, 1, seed, 0);
  HostEngine e(G, R);
  OracleCols c(G, R);
  c.load(o);
  import_from_oracle(e, c);
  const mrq_trace_params p = preset(cfg);
  std::vector<uint8_t> type(G * R), word(e.gs * (R > 1 ? R - 1 : 1)), prop8(e.gs);
  std::vector<uint64_t> term(G * R), index(G * R), logterm(G * R), commit(G * R);
  std::vector<uint32_t> prop(G);
  std::vector<uint64_t> dev_base(e.gs, 0), enc_base(G, 0), base_term(e.gs, 0);
  uint64_t n_bytes = 0, n_escapes = 0;
  for (int t = 0; t < T; ++t) {
    if (t % rebase_every == 0 || t == also_rebase_at) {  // the host re-bases now and then (terms move), as a real one would
      c.load(o);
      for (uint64_t g = 0; g < G; ++g) {
        enc_base[g] = dev_base[g] = c.last_index[g] > 30 ? c.last_index[g] - 30 : 0;
        base_term[g] = c.term[g];
      }
    }
    orc_gen_trace(o, &p, (uint64_t)t, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    struct Esc {
      uint64_t g;
      uint32_t r;
    };
    std::vector<Esc> esc;
    for (uint64_t g = 0; g < G; ++g) {  // the frame builder (what mrq_pack8 does), on the shared inline codec
      const uint32_t self = c.self_id[g];
      uint32_t min_ack = MRQ_P8_NO_ACK;
      for (uint32_t r = 0; r < R; ++r) {
        const uint32_t row = mrq_p8_row(r, self, R);
        if (row >= R - 1u) continue;
        const uint64_t w = (uint64_t)r * G + g;
        const uint8_t b = mrq_p8_encode(type[w], term[w], index[w], commit[w], enc_base[g], base_term[g]);
        word[(uint64_t)row * e.gs + g] = b;
        if (b == MRQ_P8_ESCAPE) {
          esc.push_back({g, r});
          ++n_escapes;
        } else if (b != 0) {
          ++n_bytes;
          if ((b & 3u) == 1u && (uint32_t)(b >> 2) < min_ack) min_ack = b >> 2;
        }
      }
      enc_base[g] = mrq_p8_next_base(enc_base[g], min_ack);
      prop8[g] = (uint8_t)prop[g];
    }
    const TickArgs a = e.args(0, seed, 10, 1, true);
    if (!direct)
      for (uint64_t g = 0; g < G; ++g)  // the DEVICE decode, one "thread" per group
        unpack8_group(a.in, e.meta.data(), dev_base.data(), base_term.data(), e.gs, R, word.data(), prop8.data(), g);
    for (const Esc &x : esc) {  // scatter_msgs_kernel: wide messages override their slot
      const uint64_t w = (uint64_t)x.r * G + x.g, d = (uint64_t)x.r * e.gs + x.g;
      e.itype[d] = type[w];
      e.iterm[d] = term[w];
      e.iindex[d] = index[w];
      e.ilogterm[d] = logterm[w];
      e.icommit[d] = commit[w];
    }
    orc_tick(o, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    if (direct) {
      const Inbox8 b8{word.data(), prop8.data(), dev_base.data(), base_term.data()};
      dispatch_tick8(e, a, b8);
    } else {
      dispatch_tick(e, a, 0);
    }
    for (uint64_t g = 0; g < G; ++g)  // whoever decoded the group slid its window: it must match the frame builder's
      if (dev_base[g] != enc_base[g]) {
        std::printf("FAIL %s tick %d group %llu: device window %llu, frame builder's %llu\n", where, t, (unsigned long long)g,
                    (unsigned long long)dev_base[g], (unsigned long long)enc_base[g]);
        ++failures;
        orc_destroy(o);
        return;
      }
    c.load(o);
    if (!compare(e, c, where, (uint64_t)t)) break;
  }
  std::printf("  %-34s %4d ticks  bytes %9llu  escapes %9llu  fast %8llu  general %8llu\n", where, T, (unsigned long long)n_bytes,
              (unsigned long long)n_escapes, (unsigned long long)e.fast_groups, (unsigned long long)e.slow_groups);
  if (R > 1 && n_bytes <= n_escapes) {  // (a sanity floor: the traces must exercise the bytes, not just the escapes)
    std::printf("FAIL %s: the byte form carried too little (%llu bytes, %llu escapes)\n", where, (unsigned long long)n_bytes,
                (unsigned long long)n_escapes);
    ++failures;
  }
  orc_destroy(o);
}

// Nothing protocol-shaped about it (tests/test_oracle_vs_pymodel.py's message soup, here against the DEVICE source): any
// type from any sender with terms around the receiver's, rejects, indices below / at / beyond the log, commits a
// little beyond it (upstream would panic: both sides count an error and skip), bursts of proposals.
// mode 0 / 1: wide inbox through fast + general / general only.  mode 8: the same soup as BYTE frames consumed by
// fast_group_tick8 / general_group_tick8 (re-based every tick; whatever does not fit a byte rides the wide list).
void run_soup(uint64_t G, uint32_t R, int T, uint64_t seed, int mode) {
  char where[96];
  std::snprintf(where, sizeof where, "message soup G=%llu R=%u mode=%d", (unsigned long long)G, R, mode);
  std::vector<uint8_t> word(((G + 127) / 128) * 128 * (R > 1 ? R - 1 : 1)), prop8(((G + 127) / 128) * 128);
  std::vector<uint64_t> base_index(((G + 127) / 128) * 128, 0), base_term(((G + 127) / 128) * 128, 0);
  orc_engine *o = orc_create(G, R, 0, 5, 1, seed, 0);
  HostEngine e(G, R);
  OracleCols c(G, R);
  c.load(o);
  import_from_oracle(e, c);
  std::vector<uint8_t> type(G * R);
  std::vector<uint64_t> term(G * R), index(G * R), logterm(G * R), commit(G * R);
  std::vector<uint32_t> prop(G);
  static const uint8_t kTypes[] = {0, 0, 0, 3, 4, 4, 4, 5, 6, 6, 6, 6, 6, 8, 9, 4 | 0x80, 6 | 0x80, 3 | 0x80};
  uint64_t x = seed * 77 + 5;
  auto rnd = [&]() { return x = mrq_mix64(x + 0x9E3779B97F4A7C15ull); };
  auto around = [](uint64_t v, int64_t d) -> uint64_t { return (d < 0 && v < (uint64_t)(-d)) ? 0 : v + (uint64_t)d; };
  uint32_t seen_roles = 0;
  for (int t = 0; t < T; ++t) {
    for (uint64_t g = 0; g < G; ++g) {
      for (uint32_t r = 0; r < R; ++r) {
        const uint64_t w = (uint64_t)r * G + g;
        const uint8_t ty = kTypes[rnd() % sizeof kTypes];
        type[w] = ty;
        if ((ty & 0x0F) == 0) {
          term[w] = index[w] = logterm[w] = commit[w] = 0;
          continue;
        }
        const uint64_t u = rnd() % 64;  // mostly the receiver's own term; now and then stale or ahead
        term[w] = around(c.term[g], u == 0 ? -2 : u == 1 ? -1 : u == 2 ? 1 : u == 3 ? 2 : 0);
        index[w] = around(c.last_index[g], (int64_t)(rnd() % 7) - 3);
        logterm[w] = around(c.last_term[g], (int64_t)(rnd() % 3) - 1);
        commit[w] = c.committed[g] + rnd() % 4;
        if ((ty & 0x0F) == 3) {  // a host-resolved MsgApp reports the log AFTER the append (include/mrq.h):
          if (commit[w] > index[w]) commit[w] = index[w];  // commit <= its new last index,
          if (index[w] == 0) logterm[w] = 0;               // an empty log has no last term,
          if (logterm[w] > term[w]) logterm[w] = term[w];  // and no entry is newer than the leader that sent it: with
          // that raft invariant a new leader's own entries are the only ones of its term, which is what lets the
          // engine keep `term_start` instead of per-entry terms (tests/test_golden.py, TestCommit)
        }
      }
      static const uint32_t kProps[] = {0, 0, 1, 4};
      prop[g] = kProps[rnd() % 4];
    }
    for (uint32_t r = 0; r < R; ++r) {
      std::memcpy(&e.itype[(uint64_t)r * e.gs], &type[(uint64_t)r * G], G);
      std::memcpy(&e.iterm[(uint64_t)r * e.gs], &term[(uint64_t)r * G], G * 8);
      std::memcpy(&e.iindex[(uint64_t)r * e.gs], &index[(uint64_t)r * G], G * 8);
      std::memcpy(&e.ilogterm[(uint64_t)r * e.gs], &logterm[(uint64_t)r * G], G * 8);
      std::memcpy(&e.icommit[(uint64_t)r * e.gs], &commit[(uint64_t)r * G], G * 8);
    }
    std::memcpy(e.iprop.data(), prop.data(), G * 4);
    const TickArgs a = e.args(0, seed, 5, 1, true);
    if (mode == 8) {  // re-encode as a byte frame: the wide inbox keeps only what escapes (as scatter_msgs_kernel leaves it)
      for (uint64_t g = 0; g < G; ++g) {
        base_index[g] = c.last_index[g] > 2 ? c.last_index[g] - 2 : 0;
        base_term[g] = c.term[g];
        for (uint32_t r = 0; r < R; ++r) {
          const uint32_t row = mrq_p8_row(r, c.self_id[g], R);
          const uint64_t w = (uint64_t)r * G + g, d = (uint64_t)r * e.gs + g;
          if (row >= R - 1u) continue;
          const uint8_t b = mrq_p8_encode(type[w], term[w], index[w], commit[w], base_index[g], base_term[g]);
          word[(uint64_t)row * e.gs + g] = b;
          if (b != MRQ_P8_ESCAPE) e.itype[d] = 0xEE, e.iterm[d] = e.iindex[d] = e.icommit[d] = ~0ull;  // poison: must come from the byte
        }
        prop8[g] = (uint8_t)prop[g];
        e.iprop[g] = 0xEEEEEEEEu;
      }
    }
    orc_tick(o, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    if (mode == 8) {
      const Inbox8 b8{word.data(), prop8.data(), base_index.data(), base_term.data()};
      dispatch_tick8(e, a, b8);
    } else {
      dispatch_tick(e, a, mode);
    }
    c.load(o);
    if (!compare(e, c, where, (uint64_t)t)) break;
    for (uint64_t g = 0; g < G; ++g) seen_roles |= 1u << c.role[g];
  }
  std::printf("  %-34s %4d ticks  fast %9llu  general %9llu  oracle errors %llu\n", where, T, (unsigned long long)e.fast_groups,
              (unsigned long long)e.slow_groups, (unsigned long long)orc_errors(o));
  if (R >= 2 && ((seen_roles & 3u) != 3u || (R <= 5 && seen_roles != 7u))) {  // (with 6+ noisy peers nobody holds a term long enough to win)
    std::printf("FAIL %s: the soup did not drive the groups through the roles (mask %u)\n", where, seen_roles);
    ++failures;
  }
  orc_destroy(o);
}

// ---- tick mode 4 on the host: compact state (32-bit offsets from a per-group base) + the byte inbox ------------------
// What tick_fast4_kernel does per group — compact_step on registers loaded from the compact columns, the group handed
// to slow_group_ticks_c (materialise -> general tick on the wide columns -> re-compact) from the first tick that needs
// the general path — over K-tick batches (K = 1: the per-tick launch; K > 1: mrq_tick_many's single launch).  After
// every batch the wide view (materialise_group over a scratch copy is not needed: it is idempotent) must equal the
// oracle, every tick's out word and commit advance must equal the oracle's, and the window must match the frame builder's.
struct CompactHost {
  uint64_t gs;
  uint32_t R;
  std::vector<uint8_t> flag;
  std::vector<uint32_t> commit, win, gate, iblo, match;
  std::vector<uint64_t> ibase;
  uint64_t fast_ticks = 0, slow_entries = 0, hot_ticks = 0;
  CompactHost(uint64_t gs_, uint32_t R_) : gs(gs_), R(R_), flag(gs_, 0), commit(gs_, 0), win(gs_, 0), gate(gs_, 0), iblo(gs_, 0),
                                            match(gs_ * R_, 0), ibase(gs_, 0) {}
  CompactView view() { return CompactView{flag.data(), commit.data(), win.data(), gate.data(), iblo.data(), match.data(), ibase.data()}; }
  void invalidate() { std::fill(flag.begin(), flag.end(), 0); }
};

struct HostSlot {  // one tick's frame, wide inbox slot (escapes + the general path's decode target) and outputs
  std::vector<uint8_t> word, prop8, itype, delta;
  std::vector<uint64_t> iterm, iindex, ilogterm, icommit;
  std::vector<uint32_t> iprop, out;
  HostSlot(uint64_t gs, uint32_t R)
      : word(gs * (R > 1 ? R - 1 : 1), 0), prop8(gs, 0), itype(gs * R, 0xEE), delta(gs, 0xEE), iterm(gs * R, ~0ull), iindex(gs * R, ~0ull),
        ilogterm(gs * R, ~0ull), icommit(gs * R, ~0ull), iprop(gs, 0xEEEEEEEEu), out(gs, 0xEEEEEEEEu) {}
  TickDesc desc() {
    return TickDesc{word.data(), prop8.data(), InboxView{itype.data(), iterm.data(), iindex.data(), ilogterm.data(), icommit.data(), iprop.data()},
                    out.data(), delta.data()};
  }
};

bool hot_path = true;  // (off: every compact tick through compact_step alone — the two must agree with the oracle either way)
template <int R>
void host_ticks_c(HostEngine &e, CompactHost &ch, const TickArgs &a, std::vector<TickDesc> &descs, uint64_t *base_index,
                  const uint64_t *base_term) {
  Tick4Args A{};
  A.t = a;
  A.c = ch.view();
  A.base_index = base_index;
  A.base_term = base_term;
  A.nticks = (uint32_t)descs.size();
  A.d0 = descs[0];
  A.descs = descs.size() > 1 ? descs.data() : nullptr;
  A.write_through = 1;
  for (uint64_t i = 0; i < e.G; ++i) {
    CGroup<R> g;
    g.meta = e.meta[i];
    g.commit = ch.commit[i];
    g.win = ch.win[i];
    for (int r = 0; r < R; ++r) g.m[r] = ch.match[(uint64_t)r * e.gs + i];
    const uint32_t flag = ch.flag[i];
    uint32_t t = 0;
    bool slow = false;
    for (; t < A.nticks; ++t) {
      const TickDesc &d = descs[t];
      uint32_t wb[R > 1 ? R - 1 : 1] = {0};
      for (int j = 0; j < R - 1; ++j) wb[j] = d.word8[(uint64_t)j * e.gs + i];
      uint32_t o, adv, dirty;
      const uint32_t np = d.prop8 ? d.prop8[i] : 0u;
      slow = false;
      if (!hot_path || !compact_hot_step<R>(g, flag, wb, np, ch.gate[i], a.election_tick, a.heartbeat_tick, o, adv, dirty))
        slow = compact_step<R>(g, flag, wb, np, ch.gate[i], a.election_tick, a.heartbeat_tick, o, adv, dirty);
      else
        ++ch.hot_ticks;
      if (slow) break;
      d.out[i] = o;
      d.delta[i] = (uint8_t)(adv > 255u ? 255u : adv);
      ++ch.fast_ticks;
    }
    e.meta[i] = g.meta;  // the state as of the last tick the fast path handled
    ch.commit[i] = g.commit;
    ch.win[i] = g.win;
    for (int r = 0; r < R; ++r) ch.match[(uint64_t)r * e.gs + i] = g.m[r];
    if (slow) {
      slow_group_ticks_c<R>(A, i, t);
      ++ch.slow_entries;
    }
  }
  e.tick_no[0] += A.nticks;
}

void dispatch_ticks_c(HostEngine &e, CompactHost &ch, const TickArgs &a, std::vector<TickDesc> &descs, uint64_t *bi, const uint64_t *bt) {
  switch (e.R) {
    case 1: host_ticks_c<1>(e, ch, a, descs, bi, bt); break;
    case 2: host_ticks_c<2>(e, ch, a, descs, bi, bt); break;
    case 3: host_ticks_c<3>(e, ch, a, descs, bi, bt); break;
    case 4: host_ticks_c<4>(e, ch, a, descs, bi, bt); break;
    case 5: host_ticks_c<5>(e, ch, a, descs, bi, bt); break;
    case 6: host_ticks_c<6>(e, ch, a, descs, bi, bt); break;
    case 7: host_ticks_c<7>(e, ch, a, descs, bi, bt); break;
    case 8: host_ticks_c<8>(e, ch, a, descs, bi, bt); break;
  }
}

// soup = false: the synthetic trace `cfg` (as run_case8);  soup = true: the adversarial message soup (as run_soup),
// re-based every `rebase_every` ticks with the window sliding in between.
void run_case_c(uint64_t G, uint32_t R, uint32_t cfg, int T, int rebase_every, uint32_t K, bool soup = false, uint64_t big = 0,
                int also_rebase_at = -1, bool edges = false) {
  char where[112];
  std::snprintf(where, sizeof where, "compact%s K=%u G=%llu R=%u %s=%u%s", soup ? " soup" : "", K, (unsigned long long)G, R, soup ? "seed" : "cfg", cfg,
                edges ? " window edges" : big ? " near 2^31" : "");
  const uint64_t seed = 0x5EEDC000ull + cfg * 131 + R;
  const uint32_t et = soup ? 5 : 10;
  orc_engine *o = orc_create(G, R, 0, et, 1, seed, 0);
  HostEngine e(G, R);
  CompactHost ch(e.gs, R);
  OracleCols c(G, R);
  c.load(o);
  if (big) {  // steady-state leaders whose indices sit `big` above a multiple of 2^32: offsets run towards the re-base limit
    uint64_t x = 7;
    auto rnd = [&]() { return x = mrq_mix64(x + 0x9E3779B97F4A7C15ull); };
    for (uint64_t g = 0; g < G; ++g) {
      const uint32_t self = (uint32_t)(g % R) + 1;
      c.self_id[g] = (uint8_t)self;
      c.role[g] = MRQ_ROLE_LEADER;
      c.lead[g] = (uint8_t)self;
      c.term[g] = 1 + rnd() % 8;
      c.vote[g] = self;
      c.last_index[g] = big + rnd() % 40;
      c.last_term[g] = c.term[g];
      for (uint32_t r = 0; r < R; ++r) c.match[(uint64_t)r * G + g] = (g % 7 == 3 && r == (self % R)) ? 5 : c.last_index[g] - 1 - rnd() % 6;  // some far-behind followers
      c.match[(uint64_t)(self - 1) * G + g] = c.last_index[g];
      c.committed[g] = c.last_index[g] - 8;
      c.term_start[g] = (g % 5 == 0) ? c.last_index[g] - 2 : c.committed[g] - 5;  // some gates still closed
      c.rto[g] = 10;
    }
    orc_import(o, c.term.data(), c.vote.data(), c.committed.data(), c.last_index.data(), c.last_term.data(), c.term_start.data(),
               c.match.data(), c.role.data(), c.lead.data(), c.self_id.data(), nullptr, nullptr, nullptr, c.rto.data());
    c.load(o);
  }
  import_from_oracle(e, c);
  const mrq_trace_params p = preset(soup ? 3 : cfg);
  std::vector<uint8_t> type(G * R);
  std::vector<uint64_t> term(G * R), index(G * R), logterm(G * R), commit(G * R);
  std::vector<uint32_t> prop(G);
  std::vector<uint64_t> dev_base(e.gs, 0), enc_base(G, 0), base_term(e.gs, 0);
  std::vector<HostSlot> slots;
  for (uint32_t k = 0; k < K; ++k) slots.emplace_back(e.gs, R);
  std::vector<std::vector<uint32_t>> want_out(K, std::vector<uint32_t>(G));
  std::vector<std::vector<uint64_t>> want_commit(K + 1, std::vector<uint64_t>(G));
  static const uint8_t kTypes[] = {0, 0, 0, 3, 4, 4, 4, 4, 4, 5, 6, 6, 6, 8, 8, 9, 4 | 0x80, 6 | 0x80, 3 | 0x80};
  uint64_t x = seed * 77 + 5;
  auto rnd = [&]() { return x = mrq_mix64(x + 0x9E3779B97F4A7C15ull); };
  auto around = [](uint64_t v, int64_t d) -> uint64_t { return (d < 0 && v < (uint64_t)(-d)) ? 0 : v + (uint64_t)d; };
  uint64_t n_bytes = 0, n_escapes = 0;
  bool bad = false;
  for (int t0 = 0; t0 < T && !bad; t0 += (int)K) {
    if (t0 % rebase_every == 0 || t0 == also_rebase_at) {  // mrq_set_packed_base: the wide columns are made exact, the compact copies die
      const TickArgs a0 = e.args(0, seed, et, 1, true);
      for (uint64_t g = 0; g < G; ++g) materialise_group(a0.s, ch.view(), dev_base.data(), e.gs, R, g);
      ch.invalidate();
      c.load(o);
      for (uint64_t g = 0; g < G; ++g) {
        const uint64_t back = soup ? 6 : 30;
        enc_base[g] = dev_base[g] = c.last_index[g] > back ? c.last_index[g] - back : 0;
        base_term[g] = c.term[g];
      }
    }
    c.load(o);
    want_commit[0] = c.committed;
    for (uint32_t k = 0; k < K; ++k) {  // K ticks of the trace: frames built as the oracle advances
      HostSlot &s = slots[k];
      if (!soup) {
        orc_gen_trace(o, &p, (uint64_t)(t0 + k), type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
      } else {
        for (uint64_t g = 0; g < G; ++g) {
          for (uint32_t r = 0; r < R; ++r) {
            const uint64_t w = (uint64_t)r * G + g;
            const uint8_t ty = kTypes[rnd() % sizeof kTypes];
            type[w] = ty;
            if ((ty & 0x0F) == 0) {
              term[w] = index[w] = logterm[w] = commit[w] = 0;
              continue;
            }
            const uint64_t u = rnd() % (edges ? 4096 : 64);  // (the window-edge soup: terms mostly stay put)
            term[w] = around(c.term[g], u == 0 ? -2 : u == 1 ? -1 : u == 2 ? 1 : u == 3 ? 2 : 0);
            index[w] = around(c.last_index[g], (int64_t)(rnd() % 7) - 4);
            logterm[w] = around(c.last_term[g], (int64_t)(rnd() % 3) - 1);
            commit[w] = c.committed[g] + rnd() % 4;
            if (edges) {  // mostly compact-representable traffic aimed at the EDGES of the 64-entry window and of the log
              static const uint8_t kEdge[] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 0, 0, 8, 9, 6, 4 | 0x80};
              const uint8_t t2 = (rnd() % 16 == 0) ? kEdge[14 + rnd() % 6] : kEdge[rnd() % 14];  // 6 % non-ack cells
              type[w] = t2;
              if ((t2 & 0x0F) == 0) {
                term[w] = index[w] = logterm[w] = commit[w] = 0;
                continue;
              }
              const uint64_t wbase = enc_base[g], li = c.last_index[g];
              const uint64_t top = wbase + 63 < li ? wbase + 63 : li;
              switch (rnd() % 64) {
                case 0: index[w] = wbase + 64; break;              // one past the window: escapes
                case 1: index[w] = wbase ? wbase - 1 : 0; break;   // one below it: escapes
                case 2: index[w] = (rnd() % 64 == 0) ? li + 1 : li; break;  // (rarely) beyond the log: upstream's strict path, sticky
                case 3: case 4: case 5: case 6: index[w] = wbase; break;   // offset 0
                case 7: case 8: case 9: case 10: index[w] = top; break;    // the last offset the window (or the log) allows
                case 11: case 12: index[w] = li; break;
                default: index[w] = top > wbase ? wbase + rnd() % (top - wbase + 1) : wbase; break;
              }
              commit[w] = (rnd() % 4 == 0) ? wbase + rnd() % 64 : c.committed[g] + rnd() % 3;  // heartbeat commits around the window too
              if (commit[w] > li) commit[w] = li;
            }
            const uint8_t ty2 = type[w];
            if ((ty2 & 0x0F) == 3) {
              if (commit[w] > index[w]) commit[w] = index[w];
              if (index[w] == 0) logterm[w] = 0;
              if (logterm[w] > term[w]) logterm[w] = term[w];
            }
          }
          static const uint32_t kProps[] = {0, 0, 1, 4};
          prop[g] = kProps[rnd() % 4];
        }
      }
      std::fill(s.itype.begin(), s.itype.end(), 0xEE);  // poison: whatever the tick reads must come from the bytes / the escapes
      for (uint64_t g = 0; g < G; ++g) {
        const uint32_t self = c.self_id[g];
        uint32_t min_ack = MRQ_P8_NO_ACK;
        for (uint32_t r = 0; r < R; ++r) {
          const uint32_t row = mrq_p8_row(r, self, R);
          if (row >= R - 1u) continue;
          const uint64_t w = (uint64_t)r * G + g, d = (uint64_t)r * e.gs + g;
          const uint8_t b = mrq_p8_encode(type[w], term[w], index[w], commit[w], enc_base[g], base_term[g]);
          s.word[(uint64_t)row * e.gs + g] = b;
          if (b == MRQ_P8_ESCAPE) {  // scatter_msgs_kernel: the wide message rides in the slot
            s.itype[d] = type[w];
            s.iterm[d] = term[w];
            s.iindex[d] = index[w];
            s.ilogterm[d] = logterm[w];
            s.icommit[d] = commit[w];
            ++n_escapes;
          } else if (b != 0) {
            ++n_bytes;
            if ((b & 3u) == 1u && (uint32_t)(b >> 2) < min_ack) min_ack = b >> 2;
          }
        }
        enc_base[g] = mrq_p8_next_base(enc_base[g], min_ack);
        s.prop8[g] = (uint8_t)prop[g];
      }
      orc_tick(o, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
      c.load(o);
      want_out[k] = c.out;
      want_commit[k + 1] = c.committed;
    }
    const TickArgs a = e.args(0, seed, et, 1, true);
    std::vector<TickDesc> descs;
    for (uint32_t k = 0; k < K; ++k) descs.push_back(slots[k].desc());
    dispatch_ticks_c(e, ch, a, descs, dev_base.data(), base_term.data());
    // the wide view of the result (what every reading entry point of the library does first)
    for (uint64_t g = 0; g < G; ++g) materialise_group(a.s, ch.view(), dev_base.data(), e.gs, R, g);
    for (uint32_t k = 0; k < K && !bad; ++k)
      for (uint64_t g = 0; g < G; ++g) {
        const uint64_t adv = want_commit[k + 1][g] - want_commit[k][g];
        if (slots[k].out[g] != want_out[k][g] || slots[k].delta[g] != (adv > 255 ? 255 : adv)) {
          std::printf("FAIL %s tick %d group %llu: out %08x want %08x, commit advance %u want %llu\n", where, t0 + (int)k, (unsigned long long)g,
                      slots[k].out[g], want_out[k][g], slots[k].delta[g], (unsigned long long)adv);
          ++failures;
          bad = true;
          break;
        }
      }
    for (uint64_t g = 0; g < G && !bad; ++g)
      if (dev_base[g] != enc_base[g]) {
        std::printf("FAIL %s tick %d group %llu: device window %llu, frame builder's %llu\n", where, t0, (unsigned long long)g,
                    (unsigned long long)dev_base[g], (unsigned long long)enc_base[g]);
        ++failures;
        bad = true;
      }
    std::memcpy(e.out.data(), slots[K - 1].out.data(), e.gs * 4);  // compare() reads the last tick's out words from e.out
    if (!bad && !compare(e, c, where, (uint64_t)(t0 + K - 1))) bad = true;
  }
  uint64_t n_compact = 0;
  for (uint64_t g = 0; g < G; ++g) n_compact += ch.flag[g] & CF_COMPACT;
  std::printf("  %-44s %4d ticks  bytes %8llu  escapes %7llu  fast group-ticks %8llu (hot %8llu)  general entries %7llu  compact now %llu/%llu\n",
              where, T, (unsigned long long)n_bytes, (unsigned long long)n_escapes, (unsigned long long)ch.fast_ticks,
              (unsigned long long)ch.hot_ticks, (unsigned long long)ch.slow_entries, (unsigned long long)n_compact, (unsigned long long)G);
  if (!bad && !soup && R > 1 && ch.fast_ticks < (uint64_t)T * G / 4) {
    std::printf("FAIL %s: the compact fast path handled too little (%llu of %llu group-ticks)\n", where, (unsigned long long)ch.fast_ticks,
                (unsigned long long)((uint64_t)T * G));
    ++failures;
  }
  orc_destroy(o);
}

// The fused all-gather of the tick (multi-GPU mode 1): every group's commit index is stored into every rank's gather
// buffer as a low word every tick and a high word only when it changes (or when priming).  Steady-state leaders
// whose commit indices sit just below a multiple of 2^32, so that they CROSS it during the run: after every tick the
// stitched (hi << 32 | lo) view in both "ranks'" buffers must equal committed[] — a stale high word would show here.
void run_gather_case(uint64_t G, uint32_t R, int T) {
  const char *where = "fused gather: low bytes + full on change, across 2^32";
  const uint64_t seed = 0x5EED6A7Eull;
  orc_engine *o = orc_create(G, R, 0, 10, 1, seed, 0);
  OracleCols c(G, R);
  c.load(o);
  uint64_t x = 99;
  auto rnd = [&]() { return x = mrq_mix64(x + 0x9E3779B97F4A7C15ull); };
  for (uint64_t g = 0; g < G; ++g) {  // steady-state leaders (bench.py steady_state), placed under the boundary
    const uint32_t self = (uint32_t)(g % R) + 1;
    c.self_id[g] = (uint8_t)self;
    c.role[g] = MRQ_ROLE_LEADER;
    c.lead[g] = (uint8_t)self;
    c.term[g] = 1 + rnd() % 8;
    c.vote[g] = self;
    c.last_index[g] = ((3 + g % 3) << 32) - 1 - rnd() % 40;
    c.last_term[g] = c.term[g];
    for (uint32_t r = 0; r < R; ++r) c.match[(uint64_t)r * G + g] = c.last_index[g] - 1 - rnd() % 6;
    c.match[(uint64_t)(self - 1) * G + g] = c.last_index[g];
    c.committed[g] = c.last_index[g] - 8;
    c.term_start[g] = c.committed[g] - 5;
    c.rto[g] = 10;
  }
  orc_import(o, c.term.data(), c.vote.data(), c.committed.data(), c.last_index.data(), c.last_term.data(), c.term_start.data(),
             c.match.data(), c.role.data(), c.lead.data(), c.self_id.data(), nullptr, nullptr, nullptr, c.rto.data());
  c.load(o);
  HostEngine e(G, R);
  import_from_oracle(e, c);
  const mrq_trace_params p = preset(3);
  std::vector<uint8_t> type(G * R);
  std::vector<uint64_t> term(G * R), index(G * R), logterm(G * R), commit(G * R);
  std::vector<uint32_t> prop(G);
  const uint32_t world = 2, rank = 1;  // this shard is rank 1 of 2: its words land at [rank * G + g]
  std::vector<uint8_t> lo0(world * G, 0xAB), lo1(world * G, 0xAB);  // low bytes, every tick
  std::vector<uint64_t> hi0(world * G, 0xDEADBEEFDEADBEEFull), hi1(world * G, 0xDEADBEEFDEADBEEFull);  // full indices, on change
  uint64_t crossed = 0;
  for (int t = 0; t < T; ++t) {
    orc_gen_trace(o, &p, (uint64_t)t, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    for (uint32_t r = 0; r < R; ++r) {
      std::memcpy(&e.itype[(uint64_t)r * e.gs], &type[(uint64_t)r * G], G);
      std::memcpy(&e.iterm[(uint64_t)r * e.gs], &term[(uint64_t)r * G], G * 8);
      std::memcpy(&e.iindex[(uint64_t)r * e.gs], &index[(uint64_t)r * G], G * 8);
      std::memcpy(&e.ilogterm[(uint64_t)r * e.gs], &logterm[(uint64_t)r * G], G * 8);
      std::memcpy(&e.icommit[(uint64_t)r * e.gs], &commit[(uint64_t)r * G], G * 8);
    }
    std::memcpy(e.iprop.data(), prop.data(), G * 4);
    orc_tick(o, type.data(), term.data(), index.data(), logterm.data(), commit.data(), prop.data(), 1);
    TickArgs a = e.args(0, seed, 10, 1, true);
    a.world = world;
    a.rank = rank;
    a.gather_prime = t == 0 ? 1u : 0u;  // the first tick publishes every high word, later ticks only the changed ones
    a.peer_lo[0] = lo0.data();
    a.peer_full[0] = hi0.data();
    a.peer_lo[1] = lo1.data();
    a.peer_full[1] = hi1.data();
    const std::vector<uint64_t> before(e.committed.begin(), e.committed.begin() + G);
    dispatch_tick(e, a, t % 2);  // alternate: fast + general, general only
    c.load(o);
    if (!compare(e, c, where, (uint64_t)t)) break;
    for (uint64_t g = 0; g < G; ++g) {
      crossed += (before[g] >> 8) != (e.committed[g] >> 8);
      for (const auto &bufs : {std::make_pair(&lo0, &hi0), std::make_pair(&lo1, &hi1)}) {
        const uint64_t got = ((*bufs.second)[rank * G + g] & ~0xFFull) | (*bufs.first)[rank * G + g];
        if (got != e.committed[g]) {
          std::printf("FAIL %s tick %d group %llu: gathered %llx, committed %llx\n", where, t, (unsigned long long)g,
                      (unsigned long long)got, (unsigned long long)e.committed[g]);
          ++failures;
          orc_destroy(o);
          return;
        }
      }
    }
  }
  for (uint64_t k = 0; k < G; ++k)  // and nothing was written into the other rank's half
    if (lo0[k] != 0xAB || hi1[k] != 0xDEADBEEFDEADBEEFull) {
      std::printf("FAIL %s: a word outside this rank's range was written\n", where);
      ++failures;
      break;
    }
  std::printf("  %-34s %4d ticks  group-ticks whose index changed above its low byte: %llu\n", where, T, (unsigned long long)crossed);
  if (crossed < G / 2) {
    std::printf("FAIL %s: only %llu crossings — the case does not exercise the high-word path\n", where, (unsigned long long)crossed);
    ++failures;
  }
  orc_destroy(o);
}

// The standalone quorum kernel's arithmetic (quorum_commit_one<R>: a15 + a16 on the 32-bit delta network with its
// 64-bit fallback) against the oracle's independent definition max{x : |{r : m[r] >= x}| >= q}: ties, zeros,
// 2^64-1, values exactly 2^32 around the commit index (the fallback boundary), closed and open gates.
template <int R>
void quorum_cases(uint64_t seed, int n) {
  uint64_t x = seed;
  auto rnd = [&]() { return x = mrq_mix64(x + 0x9E3779B97F4A7C15ull); };
  for (int it = 0; it < n; ++it) {
    const uint64_t around = (rnd() % 3 == 0) ? (rnd() >> (rnd() % 64)) : (1ull << 40) + (rnd() & 0xFFFF);
    auto pick = [&]() -> uint64_t {
      switch (rnd() % 8) {
        case 0: return rnd() % 4;
        case 1: return ~0ull - (rnd() % 3);
        case 2: return around + (rnd() % 11) - 5;
        case 3: return around + (1ull << 32) + (rnd() % 5) - 2;
        case 4: return around - (1ull << 32) + (rnd() % 5) - 2;
        case 5: return rnd();
        default: return around + (rnd() & 0xFFFFF);
      }
    };
    uint64_t m[R], mm[R];
    for (int r = 0; r < R; ++r) mm[r] = m[r] = pick();
    const uint64_t committed = pick();
    const uint64_t gate = (rnd() % 4 == 0) ? ~0ull : (rnd() % 2 ? committed : pick());
    const uint64_t mci = orc_quorum_index_bruteforce(mm, R);
    const bool want_moved = mci > committed && mci >= gate;
    bool moved = false;
    const uint64_t got = quorum_commit_one<R>(m, committed, gate, moved);
    if (moved != want_moved || got != (want_moved ? mci : committed)) {
      std::printf("FAIL quorum_commit_one<%d>: committed %llu gate %llu mci %llu -> got %llu moved %d\n", R,
                  (unsigned long long)committed, (unsigned long long)gate, (unsigned long long)mci, (unsigned long long)got, (int)moved);
      ++failures;
      return;
    }
  }
}

}  // namespace

int main(int argc, char **) {
  if (const char *soak = std::getenv("MRQ_SOAK")) {  // MRQ_SOAK=<n>: n more seeds of every soup, every R, every mode
    const int n = std::atoi(soak);
    for (int s = 0; s < n && failures == 0; ++s)
      for (uint32_t R = 1; R <= 8; ++R) {
        for (int mode : {0, 1, 8}) run_soup(96, R, 300, 1000 + 37 * (uint64_t)s + R, mode);
        run_case_c(96, R, 1000 + 37 * (uint32_t)s + R, 300, 8, 1 + (uint32_t)s % 4, true);
        run_case_c(96, R, 2000 + 37 * (uint32_t)s + R, 300, 16, 1 + (uint32_t)(s + 1) % 4, true, 1ull << 22, -1, true);  // window-edge soup
      }
    std::printf(failures ? "tick_host_test soak: %d failure(s)\n" : "tick_host_test soak: ok\n", failures);
    return failures ? 1 : 0;
  }
  const bool quick = argc > 1;  // any argument: the sanitizer builds run a smaller matrix (they are ~20x slower)
  const uint64_t k = quick ? 4 : 1;
  const int nq = quick ? 20000 : 200000;
  quorum_cases<1>(11, nq), quorum_cases<2>(12, nq), quorum_cases<3>(13, nq), quorum_cases<4>(14, nq);
  quorum_cases<5>(15, nq), quorum_cases<6>(16, nq), quorum_cases<7>(17, nq), quorum_cases<8>(18, nq);
  std::printf("  quorum_commit_one<1..8>             %d adversarial cases each vs the brute-force definition\n", nq);
  // election + replication (cfg 2), lag + churn (5), follower/heartbeat heavy (6), steady state (3); every R;
  // mode 0 = fast path with the general path for the rest (the product's default), mode 1 = general path only
  for (uint32_t R = 1; R <= 8; ++R) {
    run_case(257 / k + 1, R, 2, 220, 0, 0, 0);
    run_case(200 / k, R, 5, 260, 0, 0, 1000);
  }
  run_case(300 / k, 3, 6, 300, 0, 0, 0);
  run_case(300 / k, 5, 6, 300, 0, 0, 7);
  run_case(300 / k, 7, 6, 300, 1, 0, 0);
  run_case(513 / k, 5, 5, 300, 1, 0, 0);
  run_case(400 / k, 5, 2, 200, 0, 3, 0);  // a fixed self id (the G = 1 per node shape, many at once)
  run_case(400 / k, 5, 3, 120, 0, 0, 0);  // steady-state preset of the bench (from a cold start)
  run_case(300 / k, 5, 5, 400, 0, 0, 0, 7, 3);     // other timer settings: ElectionTick 7, HeartbeatTick 3
  run_case(300 / k, 3, 6, 400, 0, 0, 0, 23, 5);    // ... 23 / 5 (heartbeats every 5th tick: followers count between them)
  run_case(200 / k, 4, 2, 300, 1, 0, 0, 3, 1);     // ... a 3-tick election timeout (campaigns all the time)
  run_case(48, 3, 2, 9000 / (int)k, 0, 0, 0, 2047, 255);  // ... the largest the packed meta word holds (12-bit timers, 8-bit heartbeat)
  if (!quick) run_case(1000, 3, 2, 1024, 0, 0, 0);  // BASELINE configs[1] shape, all 1,024 ticks
  run_gather_case(600 / k, 5, 80);
  for (uint32_t R : {2u, 3u, 4u, 5u, 7u, 8u}) run_soup(120 / k + 8, R, 220, 100 + R, (int)(R % 2));
  for (uint32_t R : {2u, 3u, 5u, 8u}) run_soup(120 / k + 8, R, 220, 200 + R, 8);
  // the byte-form inbox through the device decode (unpack8_group), then the ticks
  for (uint32_t R : {1u, 2u, 3u, 5u, 7u, 8u}) run_case8(300 / k, R, 5, 200, 25);
  run_case8(400 / k, 5, 2, 300, 40);   // elections: votes and vote responses ride the bytes / the escapes
  run_case8(400 / k, 5, 3, 400, 1000, 60);  // steady state: one base once the leaders stand, then the window slides by itself
  // the tick consuming the bytes itself (fast_group_tick8 / general_group_tick8): no unpack pass
  for (uint32_t R : {1u, 2u, 3u, 4u, 5u, 6u, 7u, 8u}) run_case8(300 / k, R, 5, 220, 25, -1, true);
  run_case8(400 / k, 5, 2, 300, 40, -1, true);
  run_case8(400 / k, 3, 6, 300, 30, -1, true);    // follower / heartbeat heavy: the follower fast path on bytes
  run_case8(400 / k, 5, 3, 400, 1000, 60, true);  // steady state: nearly everything stays on the byte fast path
  // tick mode 4: compact state + the byte inbox; K = 1 (per-tick launch) and K > 1 (mrq_tick_many's single launch)
  for (uint32_t R : {1u, 2u, 3u, 4u, 5u, 6u, 7u, 8u}) run_case_c(300 / k, R, 5, 224, 32, R % 2 ? 1 : 4);
  run_case_c(400 / k, 5, 2, 300, 60, 1);           // elections from a cold start
  run_case_c(400 / k, 5, 2, 300, 60, 5);
  run_case_c(400 / k, 3, 6, 300, 30, 3);           // follower / heartbeat heavy: the follower fast path
  run_case_c(400 / k, 5, 3, 400, 1000, 1, false, 0, 60);  // steady state: one base once the leaders stand, then the window slides by itself
  run_case_c(400 / k, 5, 3, 400, 1000, 8, false, 0, 64);
  run_case_c(300 / k, 5, 3, 300, 1000, 4, false, (1ull << 33) + 100);          // 64-bit bases, far-behind followers, closed gates
  kCompactSpan = 120;  // a tiny offset span: every group outgrows it every few dozen ticks and is re-based by the general path
  run_case_c(200 / k, 5, 3, 240, 1000, 4, false, (1ull << 31) - 300);
  run_case_c(200 / k, 3, 5, 240, 40, 1);
  kCompactSpan = 0x7FFFFFFFu;
  for (uint32_t R : {2u, 3u, 5u, 8u}) run_case_c(120 / k + 8, R, 300 + R, 240, 8, R == 5 ? 4 : 1, true);  // message soups
  // the window-edge soup on steady-state leaders: acks / heartbeats at offset 0, 63, 64, -1, at and beyond lastIndex, rare term changes
  for (uint32_t R : {2u, 3u, 4u, 5u, 7u, 8u}) run_case_c(160 / k + 8, R, 500 + R, 320, 16, R % 2 ? 1 : 4, true, 1ull << 22, -1, true);
  hot_path = false;  // the complete compact step alone (the hot step is an early exit of it, never a different answer)
  run_case_c(300 / k, 5, 5, 224, 32, 1);
  run_case_c(300 / k, 3, 3, 300, 1000, 4, false, 0, 60);
  hot_path = true;
  std::printf(failures ? "tick_host_test: %d failure(s)\n" : "tick_host_test: ok\n", failures);
  return failures ? 1 : 0;
}
