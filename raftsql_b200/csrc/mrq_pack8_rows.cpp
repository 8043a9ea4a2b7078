// mrq_pack8_rows.cpp — the inner loop of mrq_pack8 (host CPU code; include/mrq_packed8.h is the codec).
//
// One sender row over a block of groups: the bytes of the common case — an accepted MsgAppResp of the window's term
// whose index lies in the 64-entry window — computed in a branch-free loop the host compiler vectorises (an AVX2 clone
// is picked at run time where the CPU has it); an empty cell is 0; every other cell is marked MRQ_P8_ESCAPE and
// re-encoded exactly by the caller (mrq_p8_encode: heartbeats, responses, genuine escapes).  A plain C++ translation
// unit, so that the host compiler, not nvcc's front end, sees the target attributes.
#include <stdint.h>

#include "../../include/mrq_packed8.h"

#define MRQI_ROW_BODY                                                                     \
  for (uint64_t i = 0; i < n; ++i) {                                                      \
    const uint32_t t = ty[i];                                                             \
    const uint64_t d = ix[i] - bi[i];                                                     \
    const bool fits = (t == 4u) & (tm[i] == bt[i]) & (d <= 63u);                          \
    out[i] = fits ? (uint8_t)(1u | ((uint32_t)d << 2)) : (uint8_t)((t & 0x0Fu) ? MRQ_P8_ESCAPE : 0u); \
  }

static void row_generic(const uint8_t *__restrict ty, const uint64_t *__restrict tm, const uint64_t *__restrict ix,
                        const uint64_t *__restrict bi, const uint64_t *__restrict bt, uint64_t n, uint8_t *__restrict out) {
  MRQI_ROW_BODY
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void row_avx2(const uint8_t *__restrict ty, const uint64_t *__restrict tm,
                                                     const uint64_t *__restrict ix, const uint64_t *__restrict bi,
                                                     const uint64_t *__restrict bt, uint64_t n, uint8_t *__restrict out) {
  MRQI_ROW_BODY
}
#endif

// place a sender row's bytes into the frame: sender slot r lands in frame row r for groups whose own id is above it
// (self > r + 1), in row r - 1 for groups whose own id is below it (self < r + 1), nowhere for the group's own slot;
// and fold the acks' window offsets into the per-group minimum.  Byte-wide blends: 32 groups per AVX2 instruction.
#define MRQI_PLACE_BODY                                                             \
  const uint8_t own = (uint8_t)(r + 1u);                                            \
  for (uint64_t i = 0; i < n; ++i) {                                                \
    const uint8_t v = b[i], sf = self[i];                                           \
    if (hi) hi[i] = sf > own ? v : hi[i];                                           \
    if (lo) lo[i] = (sf < own && sf >= 1u) ? v : lo[i];                             \
    const uint8_t p = ((v & 3u) == 1u && sf != own) ? (uint8_t)(v >> 2) : (uint8_t)MRQ_P8_NO_ACK; \
    mn[i] = p < mn[i] ? p : mn[i];                                                  \
  }

static void place_generic(const uint8_t *__restrict b, const uint8_t *__restrict self, uint32_t r, uint64_t n, uint8_t *__restrict lo,
                          uint8_t *__restrict hi, uint8_t *__restrict mn) {
  MRQI_PLACE_BODY
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void place_avx2(const uint8_t *__restrict b, const uint8_t *__restrict self, uint32_t r, uint64_t n,
                                                       uint8_t *__restrict lo, uint8_t *__restrict hi, uint8_t *__restrict mn) {
  MRQI_PLACE_BODY
}
#endif

extern "C" __attribute__((visibility("hidden"))) void mrqi_pack8_place(const uint8_t *b, const uint8_t *self, uint32_t r, uint64_t n,
                                                                       uint8_t *lo, uint8_t *hi, uint8_t *mn) {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) {
    place_avx2(b, self, r, n, lo, hi, mn);
    return;
  }
#endif
  place_generic(b, self, r, n, lo, hi, mn);
}

// does the row hold any marked cell (MRQ_P8_ESCAPE)?
extern "C" __attribute__((visibility("hidden"))) int mrqi_pack8_marked(const uint8_t *b, uint64_t n) {
  uint32_t any = 0;
  for (uint64_t i = 0; i < n; ++i) any |= (uint32_t)(b[i] == MRQ_P8_ESCAPE);
  return (int)any;
}

extern "C" __attribute__((visibility("hidden"))) void mrqi_pack8_row(const uint8_t *ty, const uint64_t *tm, const uint64_t *ix,
                                                                     const uint64_t *bi, const uint64_t *bt, uint64_t n, uint8_t *out) {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) {
    row_avx2(ty, tm, ix, bi, bt, n, out);
    return;
  }
#endif
  row_generic(ty, tm, ix, bi, bt, n, out);
}
