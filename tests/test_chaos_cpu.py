"""Safety of the host node + consensus core under an adversarial network (CPU: the oracle is the core; the same
HostNode drives the GPU engine in tests/test_plumbing.py).  Messages are dropped, delayed and reordered, nodes are
stopped and restarted from their WAL, proposals arrive at random nodes.  Checked at the seam (what CommitC carries):

  * agreement  — every node's applied sequence is a prefix of the longest one (no divergence, no reordering);
  * integrity  — nothing is applied twice and nothing is applied that was never proposed;
  * durability — an entry a node has applied is still applied, in the same position, after that node restarts;
  * liveness   — once the network heals and everyone is up, every node catches up with the longest sequence.
"""
import os
import random

import numpy as np
import pytest

from oracle_core import make_oracle_core
from raftsql_b200.hostnode import HostNode, LocalTransport


class LossyTransport(LocalTransport):
    def __init__(self, rng, drop=0.0, delay=0.0):
        super().__init__()
        self.rng, self.drop, self.delay = rng, drop, delay
        self.limbo = []  # (deliver_at_round, message)
        self.round = 0

    def send(self, msgs):
        keep = []
        for m in msgs:
            if self.rng.random() < self.drop:
                continue
            if self.rng.random() < self.delay:
                self.limbo.append((self.round + self.rng.randint(1, 6), m))
            else:
                keep.append(m)
        self.rng.shuffle(keep)
        super().send(keep)

    def advance(self):
        self.round += 1
        due = [m for (t, m) in self.limbo if t <= self.round]
        self.limbo = [(t, m) for (t, m) in self.limbo if t > self.round]
        super().send(due)


class Cluster:
    def __init__(self, n, tmp, rng, drop, delay):
        self.n, self.tmp, self.rng = n, str(tmp), rng
        self.tr = LossyTransport(rng, drop, delay)
        self.nodes = [None] * n
        self.applied = [[] for _ in range(n)]
        self.generation = [0] * n
        for i in range(n):
            self.start(i)

    def start(self, i):
        core = make_oracle_core(self.n, i + 1, seed=1000 * self.generation[i] + i + 1)
        self.generation[i] += 1
        node = HostNode(core, i + 1, self.n, self.tr, os.path.join(self.tmp, f"raftsql-{i + 1}"))
        node.start()
        # a restart replays the committed prefix from the WAL: it must be exactly what this node had applied,
        # possibly shorter (commit index persisted a little behind), never different
        replay = [d.decode() for d in node.replay]
        assert replay == self.applied[i][: len(replay)], f"node {i}: replay differs from what it had applied"
        self.applied[i] = replay
        self.nodes[i] = node

    def stop(self, i):
        self.nodes[i].stop()
        self.nodes[i].core.close()
        self.nodes[i] = None

    def round(self):
        self.tr.advance()
        order = list(range(self.n))
        self.rng.shuffle(order)
        for i in order:
            if self.nodes[i] is not None:
                self.applied[i].extend(d.decode() for d in self.nodes[i].step_tick())

    def check_agreement(self):
        longest = max(self.applied, key=len)
        for i, a in enumerate(self.applied):
            assert a == longest[: len(a)], f"node {i} diverged: {a[-3:]} vs {longest[len(a) - 3: len(a)]}"
        assert len(set(longest)) == len(longest), "an entry was applied twice"
        return longest


@pytest.mark.parametrize("n,seed,drop,delay", [(3, 1, 0.10, 0.15), (3, 2, 0.25, 0.30), (5, 3, 0.10, 0.20), (5, 4, 0.20, 0.10)])
def test_agreement_under_drops_delays_and_restarts(tmp_path, n, seed, drop, delay):
    rng = random.Random(seed)
    clus = Cluster(n, tmp_path, rng, drop, delay)
    proposed = set()
    k = 0
    for rnd in range(1500):
        if rng.random() < 0.15:  # a client proposes at a random live node
            i = rng.randrange(n)
            if clus.nodes[i] is not None:
                k += 1
                p = f"p{k}-n{i}"
                proposed.add(p)
                clus.nodes[i].propose(p.encode())
        if rng.random() < 0.01:  # crash / restart (never more than a minority down, so progress stays possible)
            down = [i for i in range(n) if clus.nodes[i] is None]
            if down and rng.random() < 0.6:
                clus.start(rng.choice(down))
            elif len(down) < (n - 1) // 2:
                clus.stop(rng.choice([i for i in range(n) if clus.nodes[i] is not None]))
        clus.round()
        if rnd % 50 == 0:
            clus.check_agreement()
    # heal: everyone up, perfect network, then everything must converge
    clus.tr.drop = clus.tr.delay = 0.0
    for i in range(n):
        if clus.nodes[i] is None:
            clus.start(i)
    for _ in range(400):
        clus.round()
    longest = clus.check_agreement()
    assert set(longest) <= proposed, "something was applied that nobody proposed"
    assert len(longest) > 0.4 * len(proposed), f"too little progress: {len(longest)} of {len(proposed)}"
    for i, a in enumerate(clus.applied):
        assert a == longest, f"node {i} did not catch up ({len(a)} of {len(longest)})"
    roles = [nd.role for nd in clus.nodes]
    assert roles.count(2) == 1, f"exactly one leader after healing, got roles {roles}"
    terms = {nd.term for nd in clus.nodes}
    assert len(terms) == 1
    for i in range(n):
        clus.stop(i)
