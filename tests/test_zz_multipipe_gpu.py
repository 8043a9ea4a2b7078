"""The multi-group seam over the REAL engine: three nodes in one process, each with one GPU engine of G groups on
cuda:0 (raftsql_b200.multipipe.make_engine_core), the scenarios of tests/test_multipipe_cpu.py.
The host code is also exercised by the CPU suite over the oracle core."""
import pytest

import test_multipipe_cpu as cpu
from raftsql_b200.multipipe import make_engine_core

pytestmark = [pytest.mark.gpu]


@pytest.fixture(autouse=True)
def engine_core(monkeypatch):
    monkeypatch.setattr(cpu, "make_oracle_multicore", make_engine_core)


def test_every_group_replicates_independently_and_in_order_gpu(tmp_path):
    cpu.test_every_group_replicates_independently_and_in_order(tmp_path)


def test_stopped_node_replays_every_groups_wal_and_catches_up_gpu(tmp_path):
    cpu.test_stopped_node_replays_every_groups_wal_and_catches_up(tmp_path)
