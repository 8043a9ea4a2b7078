"""The two arithmetic shortcuts of the tick kernel, proven here on CPU as properties, independent of
any GPU run:

  (1) upstream evaluates maybeCommit() after EVERY successful maybeUpdate; the kernel evaluates it once per tick —
      exact whenever no match exceeds lastIndex, and the kernel falls back to eager evaluation (its `strict` bit)
      in precisely the other case;
  (2) the q-th largest of R uint64 match values is computed on 32-bit deltas max(x - committed, 0) with a borrow
      chain and a min/max network, falling back to the 64-bit network only when some value is >= 2^32 ahead.

Both restatements below mirror the device code line by line (raftsql_b200/csrc/mrq_kernels.cuh: Group::flushCommit /
step<R_> leader branch, delta32 / quorum_index / quorum_index32) in plain Python integers.
"""
import random

import pytest

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


# ---- (2) the selection ---------------------------------------------------------------------------------------
def delta32(m, c):
    """device delta32(): (d, bad) via sub.cc / subc.cc / subc"""
    lo = ((m & M32) - (c & M32)) & M32
    borrow1 = 1 if (m & M32) < (c & M32) else 0
    hi_full = (m >> 32) - (c >> 32) - borrow1
    hi = hi_full & M32
    b = M32 if hi_full < 0 else 0  # subc 0,0: all-ones iff the 64-bit subtraction borrowed (m < c)
    d = lo if (b | hi) == 0 else 0
    bad = (~b & M32) & hi
    return d, bad


def quorum_index32(d):
    R = len(d)
    q = R // 2 + 1
    if R == 1:
        return d[0]
    if R == 2:
        return min(d)
    if R == 3:
        return max(min(d[0], d[1]), min(max(d[0], d[1]), d[2]))
    if R == 5:
        lo_ab, hi_ab = min(d[0], d[1]), max(d[0], d[1])
        lo_cd, hi_cd = min(d[2], d[3]), max(d[2], d[3])
        f, g = max(lo_ab, lo_cd), min(hi_ab, hi_cd)
        return max(min(d[4], f), min(max(d[4], f), g))
    v = list(d)
    for p in range(q):  # partial selection by bubbling maxima
        for i in range(R - 1, p, -1):
            a, b = v[i - 1], v[i]
            v[i - 1], v[i] = max(a, b), min(a, b)
    return v[q - 1]


def quorum_index_device(m, committed):
    ds, bad = [], 0
    for x in m:
        d, b = delta32(x, committed)
        ds.append(d)
        bad |= b
    if bad == 0:
        return (committed + quorum_index32(ds)) & M64
    return sorted(m, reverse=True)[len(m) // 2]  # the 64-bit network: q-th largest


def interesting_u64(rng, around):
    k = rng.randrange(8)
    if k == 0:
        return rng.randrange(0, 4)
    if k == 1:
        return M64 - rng.randrange(0, 3)
    if k == 2:
        return (around + rng.randrange(-5, 6)) & M64
    if k == 3:
        return (around + (1 << 32) + rng.randrange(-2, 3)) & M64  # right at the 32-bit fallback boundary
    if k == 4:
        return (around - (1 << 32) + rng.randrange(-2, 3)) & M64
    if k == 5:
        return rng.randrange(0, 1 << 64)
    return (around + rng.randrange(0, 1 << 20)) & M64


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8])
def test_delta_selection_equals_qth_largest_where_it_matters(R):
    """quorum_index() promises: the exact mci whenever mci > committed, some value <= committed otherwise."""
    rng = random.Random(R)
    for _ in range(20000):
        committed = interesting_u64(rng, 1 << 40)
        m = [interesting_u64(rng, committed) for _ in range(R)]
        true_mci = sorted(m, reverse=True)[R // 2]  # q = R/2+1 -> index q-1 = R//2
        got = quorum_index_device(m, committed)
        if true_mci > committed:
            assert got == true_mci, (m, committed)
        else:
            assert got <= committed, (m, committed)


def test_delta32_flags_exactly_the_out_of_range_cases():
    rng = random.Random(9)
    for _ in range(100000):
        c = interesting_u64(rng, 1 << 50)
        m = interesting_u64(rng, c)
        d, bad = delta32(m, c)
        if m < c:
            assert (d, bad) == (0, 0)
        elif m - c <= M32:
            assert (d, bad) == (m - c, 0)
        else:
            assert bad != 0


# ---- (1) lazy evaluation ---------------------------------------------------------------------------------------
def upstream_leader_tick(match, committed, gate, last_index, acks, nprop, self_slot):
    """upstream: Step every ack (maybeUpdate -> maybeCommit -> bcastAppend), then appendEntry."""
    match = list(match)
    bcast = False

    def maybe_commit():
        nonlocal committed
        mci = sorted(match, reverse=True)[len(match) // 2]
        if mci > committed and gate <= mci <= last_index:  # term(mci) == Term for a leader
            committed = mci
            return True
        return False

    for r, idx in acks:
        if match[r] < idx:
            match[r] = idx
            if maybe_commit():
                bcast = True
    if nprop:
        last_index += nprop
        match[self_slot] = max(match[self_slot], last_index)
        maybe_commit()
        bcast = True
    return match, committed, last_index, bcast


def kernel_leader_tick(match, committed, gate, last_index, acks, nprop, self_slot, strict):
    """the device's general path: deferred evaluation with the strict escape (Group::step / flushCommit / appendEntry)."""
    match = list(match)
    bcast = False
    pending = False

    def maybe_commit():
        nonlocal committed
        mci = quorum_index_device(match, committed)
        if mci > committed and gate <= mci <= last_index:
            committed = mci
            return True
        return False

    for r, idx in acks:
        if match[r] < idx:
            if idx > last_index and not strict:  # out-of-range ack: settle what is pending, then go eager
                if pending:
                    pending = False
                    if maybe_commit():
                        bcast = True
                strict = True
            match[r] = idx
            if strict:
                if maybe_commit():
                    bcast = True
            else:
                pending = True
    if nprop:
        last_index += nprop
        match[self_slot] = max(match[self_slot], last_index)
        pending = False
        maybe_commit()
        bcast = True
    if pending and maybe_commit():
        bcast = True
    return match, committed, last_index, bcast, strict


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7, 8])
def test_deferred_quorum_evaluation_equals_upstream(R):
    rng = random.Random(100 + R)
    for _ in range(8000):
        last_index = rng.randrange(10, 1 << 30)
        self_slot = rng.randrange(R)
        match = [max(0, last_index - rng.randrange(0, 12)) for _ in range(R)]
        match[self_slot] = last_index
        committed = max(0, min(match) - rng.randrange(0, 5))
        gate = rng.choice([0, committed, committed + 1, last_index - 1, last_index, last_index + 1])
        strict = False
        state_u = (match, committed, last_index)
        state_k = (match, committed, last_index)
        for _tick in range(4):  # several ticks, so a sticky strict bit is exercised across ticks too
            n_acks = rng.randrange(0, R + 1)
            acks = []
            for r in rng.sample(range(R), n_acks):
                if r == self_slot:
                    continue
                kind = rng.randrange(10)
                li = state_u[2]
                idx = li + rng.randrange(1, 50) if kind == 0 else max(0, li - rng.randrange(0, 6))  # 10%: beyond lastIndex
                acks.append((r, idx))
            acks.sort()  # sender order
            nprop = rng.choice([0, 0, 1, 3])
            mu, cu, lu, bu = upstream_leader_tick(state_u[0], state_u[1], gate, state_u[2], acks, nprop, self_slot)
            mk, ck, lk, bk, strict = kernel_leader_tick(state_k[0], state_k[1], gate, state_k[2], acks, nprop, self_slot, strict)
            assert (mu, cu, lu, bu) == (mk, ck, lk, bk), (acks, nprop, gate)
            state_u, state_k = (mu, cu, lu), (mk, ck, lk)
