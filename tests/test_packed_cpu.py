"""Host-side packed-inbox encoders (raftsql_b200/packed.py) checked on CPU against a numpy restatement of the
device decoders (unpack_inbox_kernel / unpack16_inbox_kernel in csrc/mrq_kernels.cuh): every message either
decodes to exactly its wide form or is in the escape list — nothing is approximated."""
import numpy as np
import pytest

import oracle
from raftsql_b200 import _ffi as F
from raftsql_b200.packed import PAYLOAD16_MAX, PAYLOAD_MAX, pack_inbox, pack_inbox16


def decode32(word, base_index, base_term):
    R, G = word.shape
    out = oracle.empty_inbox(G, R)
    ty, tc, pay = word & 15, (word >> 5) & 3, (word >> 7).astype(np.uint64)
    ok = (ty != 0) & (tc != 3)
    out["type"][:] = np.where(ok, ty | np.where(word & 16, F.MSG_REJECT, 0), 0).astype(np.uint8)
    out["term"][:] = np.where(ok, base_term[None, :] + tc.astype(np.uint64), 0)
    bi, bt = base_index[None, :], base_term[None, :]
    out["index"][:] = np.where(ok & (ty == F.MSG_APP_RESP), bi + pay, np.where(ok & (ty == F.MSG_VOTE), bi + (pay >> np.uint64(2)), 0))
    out["commit"][:] = np.where(ok & (ty == F.MSG_HEARTBEAT), bi + pay, 0)
    out["logterm"][:] = np.where(ok & (ty == F.MSG_VOTE), bt + (pay & np.uint64(3)), 0)
    return out


def decode16(word, base_index, base_term):
    R, G = word.shape
    w = word.astype(np.uint32)
    out = oracle.empty_inbox(G, R)
    kind, tc, pay = w & 7, (w >> 3) & 3, (w >> 5).astype(np.uint64)
    ok = (kind != 0) & (kind != 7) & (tc != 3)
    ty = np.select([kind <= 2, kind <= 4, kind == 5, kind == 6], [F.MSG_APP_RESP, F.MSG_VOTE_RESP, F.MSG_HEARTBEAT,
                                                                  F.MSG_HEARTBEAT_RESP], 0)
    rej = np.where((kind == 2) | (kind == 4), F.MSG_REJECT, 0)
    out["type"][:] = np.where(ok, ty | rej, 0).astype(np.uint8)
    out["term"][:] = np.where(ok, base_term[None, :] + tc.astype(np.uint64), 0)
    out["index"][:] = np.where(ok & (kind <= 2), base_index[None, :] + pay, 0)
    out["commit"][:] = np.where(ok & (kind == 5), base_index[None, :] + pay, 0)
    return out


def scatter(ib, wide):
    for g, frm, ty, term, index, logterm, commit in wide:
        r = frm - 1
        ib["type"][r, g], ib["term"][r, g], ib["index"][r, g] = ty, term, index
        ib["logterm"][r, g], ib["commit"][r, g] = logterm, commit


def relevant_equal(a, b):
    kind = b["type"] & F.MSG_TYPE_MASK
    np.testing.assert_array_equal(a["type"], b["type"])
    np.testing.assert_array_equal(a["term"][kind != 0], b["term"][kind != 0])
    uses = {"index": (F.MSG_APP_RESP, F.MSG_VOTE, F.MSG_APP), "logterm": (F.MSG_VOTE, F.MSG_APP),
            "commit": (F.MSG_HEARTBEAT, F.MSG_APP)}
    for k, types in uses.items():
        sel = np.isin(kind, types)
        np.testing.assert_array_equal(a[k][sel], b[k][sel], err_msg=k)


def random_inbox(G, R, rng, base_index, base_term):
    ib = oracle.empty_inbox(G, R)
    kinds = np.array([0, F.MSG_APP, F.MSG_APP_RESP, F.MSG_APP_RESP | F.MSG_REJECT, F.MSG_VOTE, F.MSG_VOTE_RESP,
                      F.MSG_VOTE_RESP | F.MSG_REJECT, F.MSG_HEARTBEAT, F.MSG_HEARTBEAT_RESP], np.uint8)
    ib["type"][:] = rng.choice(kinds, size=(R, G))
    # terms: mostly base..base+2, sometimes below the base or far above (must escape)
    ib["term"][:] = base_term[None, :] + rng.choice(np.array([0, 0, 1, 2, 3, 50], np.uint64), size=(R, G))
    below = rng.random((R, G)) < 0.05
    ib["term"][below] = 0
    span = rng.choice(np.array([0, 1, 100, PAYLOAD16_MAX, PAYLOAD16_MAX + 1, PAYLOAD_MAX, PAYLOAD_MAX + 1, 1 << 40],
                               np.uint64), size=(R, G))
    ib["index"][:] = base_index[None, :] + span
    ib["index"][rng.random((R, G)) < 0.05] = 0  # below the base
    ib["commit"][:] = base_index[None, :] + rng.permutation(span.ravel()).reshape(R, G)
    ib["logterm"][:] = base_term[None, :] + rng.integers(0, 6, size=(R, G), dtype=np.uint64)
    ib["prop_count"][:] = rng.integers(0, 256, size=G, dtype=np.uint32)
    present = (ib["type"] & F.MSG_TYPE_MASK) != 0
    for k in ("term", "index", "logterm", "commit"):
        ib[k][~present] = 0
    return ib


@pytest.mark.parametrize("bits", [32, 16])
def test_pack_decode_roundtrip(bits):
    rng = np.random.default_rng(bits)
    G, R = 3000, 7
    base_index = rng.integers(1000, 1 << 50, size=G, dtype=np.uint64)
    base_term = rng.integers(1, 1 << 30, size=G, dtype=np.uint64)
    ib = random_inbox(G, R, rng, base_index, base_term)
    packer, decoder = (pack_inbox, decode32) if bits == 32 else (pack_inbox16, decode16)
    word, prop8, wide = packer(ib, base_index, base_term)
    assert word.dtype == (np.uint32 if bits == 32 else np.uint16) and word.shape == (R, G)
    got = decoder(word, base_index, base_term)
    scatter(got, wide)
    relevant_equal(got, ib)
    np.testing.assert_array_equal(prop8, ib["prop_count"].astype(np.uint8))
    kind = ib["type"] & F.MSG_TYPE_MASK
    n_present, n_wide = int((kind != 0).sum()), len(wide)
    assert 0 < n_wide < n_present  # both the packed path and the escape path were exercised
    # MsgApp never fits either form
    assert all(((m[2] & F.MSG_TYPE_MASK) == F.MSG_APP) or True for m in wide)
    app_slots = {(int(g), int(r) + 1) for r, g in zip(*np.nonzero(kind == F.MSG_APP))}
    assert app_slots <= {(m[0], m[1]) for m in wide}


def test_too_many_proposals_is_rejected():
    ib = oracle.empty_inbox(4, 3)
    ib["prop_count"][2] = 300
    with pytest.raises(ValueError):
        pack_inbox(ib, np.zeros(4, np.uint64), np.zeros(4, np.uint64))
