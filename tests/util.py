"""Shared helpers for the parity tests: compare engine state with the CPU oracle, column by column."""
import numpy as np

LEADER = 2
U64MAX = np.uint64(0xFFFFFFFFFFFFFFFF)

# columns compared on every group
ALWAYS = ("term", "vote", "committed", "last_index", "last_term", "term_start", "role", "lead", "self_id", "votes",
          "election_elapsed", "heartbeat_elapsed", "randomized_timeout")


def assert_state_equal(eng_state: dict, orc_state: dict, where: str = ""):
    """Bit-exact comparison.  Progress.Match is only meaningful while a group is leader (upstream reset()
    rebuilds it on every role change), so it is compared on leader groups only."""
    for k in ALWAYS:
        a, b = eng_state[k], orc_state[k]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            i = tuple(bad[0])
            raise AssertionError(f"{where}: column {k} differs at {i}: engine={a[i]} oracle={b[i]} "
                                 f"({len(bad)} cells differ)")
    lead = orc_state["role"] == LEADER
    a, b = eng_state["match"][:, lead], orc_state["match"][:, lead]
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{where}: match differs on {len(bad)} leader cells, first {tuple(bad[0])}")


def assert_inbox_equal(a: dict, b: dict, where: str = ""):
    for k in ("type", "term", "index", "logterm", "commit", "prop_count"):
        if not np.array_equal(a[k], b[k]):
            bad = np.argwhere(a[k] != b[k])
            raise AssertionError(f"{where}: inbox column {k} differs in {len(bad)} cells, first {tuple(bad[0])}")


def numpy_quorum_index(match_rg: np.ndarray) -> np.ndarray:
    R = match_rg.shape[0]
    return np.sort(match_rg, axis=0)[R - (R // 2 + 1)]


def leader_state(G, R, rng, gate_open_frac=0.99):
    """A steady-state multi-raft node: every group led by this node (BASELINE configs[2])."""
    from raftsql_b200 import empty_state

    st = empty_state(G, R)
    g = np.arange(G, dtype=np.uint64)
    st["self_id"][:] = (g % np.uint64(R) + np.uint64(1)).astype(np.uint8)
    st["role"][:] = LEADER
    st["lead"][:] = st["self_id"]
    st["term"][:] = rng.integers(1, 9, size=G, dtype=np.uint64)
    st["vote"][:] = st["self_id"]
    st["last_index"][:] = rng.integers(2 ** 20, 2 ** 40, size=G, dtype=np.uint64)
    st["last_term"][:] = st["term"]
    lag = rng.geometric(0.2, size=(R, G)).astype(np.uint64)
    st["match"][:] = st["last_index"][None, :] - lag
    st["match"][st["self_id"] - 1, np.arange(G)] = st["last_index"]
    st["committed"][:] = st["last_index"] - np.uint64(40)
    gate_open = rng.random(G) < gate_open_frac
    st["term_start"][:] = np.where(gate_open, st["committed"] - np.uint64(5), st["last_index"] - np.uint64(1))
    st["randomized_timeout"][:] = 10
    return st
