/*
 * selftest.c — runs the oracle (TEST INFRASTRUCTURE, see raft_oracle.h) under AddressSanitizer +
 * UndefinedBehaviorSanitizer: the checker itself must be free of out-of-bounds accesses, leaks and UB.
 *   gcc -O1 -g -std=c11 -pthread -fsanitize=address,undefined -fno-sanitize-recover=all selftest.c raft_oracle.c -o selftest_asan
 * Exercises: the upstream TestCommit table, Step() of every message type on every role, the synthetic trace for
 * R = 1..8 (sequential and threaded drivers), export / import round trips, and the log's truncation paths.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mrq.h"
#include "../include/mrq_trace.h"
#include "raft_oracle.h"

static int fails = 0;
#define CHECK(c, msg)                                         \
  do {                                                        \
    if (!(c)) {                                               \
      printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg);    \
      ++fails;                                                \
    }                                                         \
  } while (0)

static void kat_commit(void) {
  struct { uint64_t m[4]; int n; uint64_t t[2]; size_t nt; uint64_t sm, want; } rows[] = {
      {{1}, 1, {1}, 1, 1, 1},          {{1}, 1, {1}, 1, 2, 0},          {{2}, 1, {1, 2}, 2, 2, 2},
      {{1}, 1, {2}, 1, 2, 1},          {{2, 1, 1}, 3, {1, 2}, 2, 1, 1}, {{2, 1, 1}, 3, {1, 1}, 2, 2, 0},
      {{2, 1, 2}, 3, {1, 2}, 2, 2, 2}, {{2, 1, 2}, 3, {1, 1}, 2, 2, 0}, {{2, 1, 1, 1}, 4, {1, 2}, 2, 1, 1},
      {{2, 1, 1, 1}, 4, {1, 1}, 2, 2, 0}, {{2, 1, 1, 2}, 4, {1, 2}, 2, 1, 1}, {{2, 1, 1, 2}, 4, {1, 1}, 2, 2, 0},
      {{2, 1, 2, 2}, 4, {1, 2}, 2, 2, 2}, {{2, 1, 2, 2}, 4, {1, 1}, 2, 2, 0}};
  for (size_t i = 0; i < sizeof rows / sizeof rows[0]; ++i)
    CHECK(orc_kat_commit(rows[i].m, rows[i].n, rows[i].t, rows[i].nt, rows[i].sm) == rows[i].want, "TestCommit row");
}

static void every_message_on_every_role(void) {
  const int types[] = {MsgHup, MsgBeat, MsgProp, MsgApp, MsgAppResp, MsgVote, MsgVoteResp, MsgHeartbeat, MsgHeartbeatResp};
  for (int R = 1; R <= ORC_MAXR; ++R) {
    orc_engine *e = orc_create(3, (uint32_t)R, 0, 10, 1, 42, 0);
    /* group 0 stays follower, group 1 becomes candidate (or leader when R == 1), group 2 leader where possible */
    orc_step(e, 1, MsgHup, 1, 0, 0, 0, 0, 0, 0);
    orc_step(e, 2, MsgHup, 1, 0, 0, 0, 0, 0, 0);
    for (int v = 1; v <= R; ++v) orc_step(e, 2, MsgVoteResp, (uint64_t)v, 1, 0, 0, 0, 0, 0);
    for (uint64_t g = 0; g < 3; ++g)
      for (size_t t = 0; t < sizeof types / sizeof types[0]; ++t)
        for (int from = 1; from <= R; ++from)
          for (uint64_t term = 0; term <= 3; ++term)
            for (int rej = 0; rej <= 1; ++rej)
              orc_step(e, g, types[t], (uint64_t)from, term, 5 + term, term, 3, rej, types[t] == MsgProp ? 2 : 0);
    /* log truncation paths through the host-resolved MsgApp */
    orc_step(e, 0, MsgApp, 1, 9, 40, 9, 10, 0, 0);
    orc_step(e, 0, MsgApp, 1, 9, 20, 8, 10, 0, 0);
    orc_step(e, 0, MsgApp, 1, 9, 20, 9, 10, 0, 0);
    orc_step(e, 0, MsgApp, 1, 9, 0, 0, 0, 0, 0);
    orc_destroy(e);
  }
}

static void traces(void) {
  for (uint32_t R = 1; R <= ORC_MAXR; ++R) {
    const uint64_t G = 257;
    orc_engine *a = orc_create(G, R, 1000, 10, 1, 7, 0), *b = orc_create(G, R, 1000, 10, 1, 7, 0);
    mrq_trace_params p = mrq_trace_preset(5);
    uint8_t *type = malloc(R * G);
    uint64_t *term = malloc(R * G * 8), *index = malloc(R * G * 8), *logterm = malloc(R * G * 8), *commit = malloc(R * G * 8);
    uint32_t *prop = malloc(G * 4);
    uint64_t *ca = malloc(G * 8), *cb = malloc(G * 8), *ta = malloc(G * 8), *li = malloc(G * 8), *lt = malloc(G * 8), *ts = malloc(G * 8),
             *vo = malloc(G * 8), *ma = malloc(R * G * 8);
    uint8_t *ro = malloc(G), *le = malloc(G), *se = malloc(G), *vs = malloc(R * G);
    uint16_t *el = malloc(G * 2), *hb = malloc(G * 2), *rt = malloc(G * 2);
    for (uint64_t t = 0; t < 200; ++t) {
      orc_gen_trace(a, &p, t, type, term, index, logterm, commit, prop, 1);
      orc_tick(a, type, term, index, logterm, commit, prop, 1);
      orc_tick(b, type, term, index, logterm, commit, prop, 3);
      if (t == 120) { /* export / import round trip on b */
        orc_export(b, ta, vo, cb, li, lt, ts, ma, ro, le, se, vs, el, hb, rt, NULL);
        orc_import(b, ta, vo, cb, li, lt, ts, ma, ro, le, se, vs, el, hb, rt);
      }
    }
    orc_export(a, NULL, NULL, ca, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    orc_export(b, NULL, NULL, cb, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    CHECK(memcmp(ca, cb, G * 8) == 0, "threaded driver / round trip diverged");
    CHECK(orc_errors(a) == 0, "trace raised commit-out-of-range errors");
    orc_quorum_commit(a, 2);
    free(type); free(term); free(index); free(logterm); free(commit); free(prop);
    free(ca); free(cb); free(ta); free(li); free(lt); free(ts); free(vo); free(ma);
    free(ro); free(le); free(se); free(vs); free(el); free(hb); free(rt);
    orc_destroy(a);
    orc_destroy(b);
  }
}

int main(void) {
  kat_commit();
  every_message_on_every_role();
  traces();
  printf(fails ? "oracle selftest: %d failure(s)\n" : "oracle selftest: ok\n", fails);
  return fails ? 1 : 0;
}
