"""tests/test_chaos_cpu.py for the multi-group node: MultiHostNode (G groups per node over one core, group-commit WAL,
sends deferred until the tick's fsync) under an adversarial network — messages dropped, delayed and reordered per
group, nodes stopped and restarted from their ONE shared WAL file, proposals arriving at random (node, group).
Checked per group at what the node publishes:

  * agreement  — every node's applied sequence of group g is a prefix of the longest one;
  * isolation  — nothing proposed to group g is ever applied in another group;
  * integrity  — nothing is applied twice, nothing is applied that was never proposed;
  * durability — what a node had applied in group g is what its WAL replays for group g after a restart;
  * liveness   — once the network heals and everyone is up, every group converges on every node, one leader each.
"""
import random

import pytest

from oracle_core import make_oracle_multicore
from raftsql_b200.multipipe import MultiHostNode, MultiLocalTransport, _GroupTransport


class LossyGroupTransport(_GroupTransport):
    def send(self, msgs):
        net = self.tr
        keep = []
        for m in msgs:
            if net.rng.random() < net.drop:
                continue
            if net.rng.random() < net.delay:
                net.limbo.append((net.round + net.rng.randint(1, 6), self.g, m))
            else:
                keep.append(m)
        net.rng.shuffle(keep)
        super().send(keep)


class LossyMultiTransport(MultiLocalTransport):
    def __init__(self, rng, drop, delay):
        super().__init__()
        self.rng, self.drop, self.delay = rng, drop, delay
        self.limbo, self.round = [], 0

    def group(self, g):
        return LossyGroupTransport(self, g)

    def advance(self):
        self.round += 1
        due = [(g, m) for (t, g, m) in self.limbo if t <= self.round]
        self.limbo = [x for x in self.limbo if x[0] > self.round]
        for g, m in due:
            _GroupTransport.send(_GroupTransport(self, g), [m])


class Cluster:
    def __init__(self, n, G, tmp, rng, drop, delay):
        self.n, self.G, self.tmp, self.rng = n, G, str(tmp), rng
        self.tr = LossyMultiTransport(rng, drop, delay)
        self.nodes = [None] * n
        self.applied = [[[] for _ in range(G)] for _ in range(n)]
        self.generation = [0] * n
        for i in range(n):
            self.start(i)

    def start(self, i):
        core = make_oracle_multicore(self.n, i + 1, self.G, seed=1000 * self.generation[i] + i + 1)
        self.generation[i] += 1
        node = MultiHostNode(core, i + 1, self.n, self.G, self.tr, f"{self.tmp}/raftsql-{i + 1}")
        for g, replay in enumerate(node.start()):
            # the shared WAL replays each group's committed prefix: exactly what this node had applied there,
            # possibly shorter (commit index persisted a little behind), never different
            replay = [d.decode() for d in replay]
            assert replay == self.applied[i][g][: len(replay)], f"node {i} group {g}: replay differs from what it had applied"
            self.applied[i][g] = replay
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
                for g, out in enumerate(self.nodes[i].step_tick()):
                    self.applied[i][g].extend(d.decode() for d in out)

    def check(self):
        longest = []
        for g in range(self.G):
            seqs = [self.applied[i][g] for i in range(self.n)]
            lg = max(seqs, key=len)
            for i, a in enumerate(seqs):
                assert a == lg[: len(a)], f"group {g}: node {i} diverged"
            assert len(set(lg)) == len(lg), f"group {g}: an entry was applied twice"
            assert all(p.startswith(f"g{g}-") for p in lg), f"group {g} applied another group's entry: {lg[-3:]}"
            longest.append(lg)
        return longest


@pytest.mark.parametrize("n,G,seed,drop,delay", [(3, 4, 11, 0.10, 0.15), (3, 3, 12, 0.25, 0.25), (5, 3, 13, 0.10, 0.20)])
def test_multi_group_agreement_isolation_durability_liveness(tmp_path, n, G, seed, drop, delay):
    rng = random.Random(seed)
    clus = Cluster(n, G, tmp_path, rng, drop, delay)
    proposed = [set() for _ in range(G)]
    k = 0
    for rnd in range(1200):
        if rng.random() < 0.25:  # a client proposes to a random group at a random live node
            i, g = rng.randrange(n), rng.randrange(G)
            if clus.nodes[i] is not None:
                k += 1
                p = f"g{g}-p{k}-n{i}"
                proposed[g].add(p)
                clus.nodes[i].propose(g, p.encode())
        if rng.random() < 0.01:  # crash / restart, never more than a minority down
            down = [i for i in range(n) if clus.nodes[i] is None]
            if down and rng.random() < 0.6:
                clus.start(rng.choice(down))
            elif len(down) < (n - 1) // 2:
                clus.stop(rng.choice([i for i in range(n) if clus.nodes[i] is not None]))
        clus.round()
        if rnd % 50 == 0:
            clus.check()
    # heal: everyone up, perfect network, then every group must converge
    clus.tr.drop = clus.tr.delay = 0.0
    for i in range(n):
        if clus.nodes[i] is None:
            clus.start(i)
    for _ in range(400):
        clus.round()
    longest = clus.check()
    for g in range(G):
        assert set(longest[g]) <= proposed[g], f"group {g}: something was applied that nobody proposed there"
        assert len(longest[g]) > 0.3 * len(proposed[g]), f"group {g}: too little progress ({len(longest[g])} of {len(proposed[g])})"
        for i in range(n):
            assert clus.applied[i][g] == longest[g], f"group {g}: node {i} did not catch up"
        roles = [int(clus.nodes[i].state["role"][g]) for i in range(n)]
        assert roles.count(2) == 1, f"group {g}: exactly one leader after healing, got {roles}"
        assert len({int(clus.nodes[i].state["term"][g]) for i in range(n)}) == 1
    for i in range(n):
        assert clus.nodes[i].wal.syncs > 0
        clus.stop(i)
