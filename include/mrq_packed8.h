/*
 * mrq_packed8.h — the byte form of the packed inbox (include/mrq.h, mrq_inbox_packed.word_bits = 8).
 *
 * The host->device link bounds the end-to-end tick rate, so this form ships ONE BYTE per remote sender per
 * group: R-1 rows (a node never sends to itself: the row of the group's own replica slot is left out and the
 * remaining senders close up), decoded on the device against the same two per-group base columns as the wider
 * forms.  The codec is defined here once, as plain C99 that also compiles as CUDA: the device decode kernel,
 * the host-side frame builder (mrq_pack8) and the CPU tests all include this file.
 *
 *   bits 0..1  kind     bits 2..7  payload p (0..63)
 *     kind 0 (p = 0)    no message
 *     kind 1            MSG_APP_RESP   term = base_term[g], index  = base_index[g] + p
 *     kind 2            p = 0  MSG_HEARTBEAT_RESP            term = base_term[g]
 *                       p = 1  MSG_VOTE_RESP                 term = base_term[g]
 *                       p = 2  MSG_VOTE_RESP | REJECT        term = base_term[g]
 *                       p = 63 escaped: the message rides in the wide list and overrides this slot
 *     kind 3            MSG_HEARTBEAT  term = base_term[g], commit = base_index[g] + p
 *   Everything else (MSG_VOTE, MSG_APP, rejected acks, another term, an index outside the 64-entry window)
 *   escapes.  Exact: nothing is approximated.
 *
 * The window slides by itself.  After a group's bytes of one frame are decoded, with m = the smallest payload
 * among that frame's kind-1 bytes of the group (if any):  base_index[g] += m - MRQ_P8_SLACK  when m > MRQ_P8_SLACK.
 * The rule reads nothing but the bytes, so the frame builder applies it to its own copy of the base and the two
 * stay in step however far ahead the host posts; the window's low end trails the slowest acknowledging follower
 * by MRQ_P8_SLACK entries (a follower that falls further behind, or returns after a pause, escapes).
 */
#ifndef MRQ_PACKED8_H
#define MRQ_PACKED8_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MRQ_P8_HD __host__ __device__ __forceinline__
#else
#define MRQ_P8_HD static inline
#endif

#define MRQ_P8_ESCAPE 0xFEu  /* kind 2, p = 63 */
#define MRQ_P8_SLACK 16u
#define MRQ_P8_NO_ACK 64u    /* "no kind-1 byte seen" for mrq_p8_next_base */

/* What one byte decodes to.  `type` is MRQ_MSG_* | MRQ_MSG_REJECT, 0 for none / escaped. */
typedef struct mrq_p8_cell {
  uint8_t type;
  uint8_t is_ack;   /* kind 1: `value` is the acknowledged index, `pay` its window offset */
  uint8_t is_hb;    /* kind 3: `value` is the heartbeat's commit index                    */
  uint8_t pay;
  uint64_t value;
} mrq_p8_cell;

MRQ_P8_HD mrq_p8_cell mrq_p8_decode(uint32_t w, uint64_t base_index) {
  mrq_p8_cell c;
  const uint32_t kind = w & 3u, p = (w >> 2) & 63u;
  c.type = 0;
  c.is_ack = 0;
  c.is_hb = 0;
  c.pay = (uint8_t)p;
  c.value = 0;
  if (kind == 1u) {
    c.type = 4u; /* MRQ_MSG_APP_RESP */
    c.is_ack = 1;
    c.value = base_index + p;
  } else if (kind == 3u) {
    c.type = 8u; /* MRQ_MSG_HEARTBEAT */
    c.is_hb = 1;
    c.value = base_index + p;
  } else if (kind == 2u) {
    if (p == 0u) c.type = 9u;              /* MRQ_MSG_HEARTBEAT_RESP */
    else if (p == 1u) c.type = 6u;         /* MRQ_MSG_VOTE_RESP */
    else if (p == 2u) c.type = 6u | 0x80u; /* MRQ_MSG_VOTE_RESP | MRQ_MSG_REJECT */
    /* p == 63: escaped; other payloads are never produced and decode as "no message" */
  }
  return c;
}

/* Encode one message against (base_index, base_term); returns MRQ_P8_ESCAPE when it does not fit.
 * `type` is MRQ_MSG_* | MRQ_MSG_REJECT (0 = no message -> byte 0).                                 */
MRQ_P8_HD uint8_t mrq_p8_encode(uint32_t type, uint64_t term, uint64_t index, uint64_t commit, uint64_t base_index,
                                uint64_t base_term) {
  const uint32_t kind = type & 0x0Fu;
  const uint32_t reject = type & 0x80u;
  if (kind == 0u) return 0;
  if (term != base_term) return (uint8_t)MRQ_P8_ESCAPE;
  if (kind == 4u && !reject) {
    const uint64_t d = index - base_index; /* wraps for index < base: then it is > 63 */
    return d <= 63u ? (uint8_t)(1u | ((uint32_t)d << 2)) : (uint8_t)MRQ_P8_ESCAPE;
  }
  if (kind == 8u && !reject) {
    const uint64_t d = commit - base_index;
    return d <= 63u ? (uint8_t)(3u | ((uint32_t)d << 2)) : (uint8_t)MRQ_P8_ESCAPE;
  }
  if (kind == 9u && !reject) return (uint8_t)2u;
  if (kind == 6u) return (uint8_t)(2u | ((reject ? 2u : 1u) << 2));
  return (uint8_t)MRQ_P8_ESCAPE;
}

/* The sliding-window rule: min_ack_pay = the smallest kind-1 payload of the group's bytes in this frame,
 * MRQ_P8_NO_ACK if there was none.                                                                    */
MRQ_P8_HD uint64_t mrq_p8_next_base(uint64_t base_index, uint32_t min_ack_pay) {
  return (min_ack_pay < MRQ_P8_NO_ACK && min_ack_pay > MRQ_P8_SLACK) ? base_index + (min_ack_pay - MRQ_P8_SLACK) : base_index;
}

/* Row of sender slot r (0-based) in the R-1 row layout of a group whose own id is self_id (1..R);
 * returns R-1 (out of range) for the group's own slot.                                             */
MRQ_P8_HD uint32_t mrq_p8_row(uint32_t r, uint32_t self_id, uint32_t R) {
  if (r + 1u == self_id) return R - 1u;
  return r + 1u > self_id ? r - 1u : r;
}

#endif /* MRQ_PACKED8_H */
