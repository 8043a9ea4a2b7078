"""The multi-group seam over the REAL engine: three nodes in one process, each with one GPU engine of G groups on
cuda:0 (raftsql_b200.multipipe.make_engine_core), the scenarios of tests/test_multipipe_cpu.py.

STATUS: written after round 1's GPU budget was spent.  It only uses engine entry points the GPU suite already
validates (sparse posts, proposals, tick, state export), and the host code is exercised by the CPU suite over the
oracle core — but this file itself has not yet run on hardware, so it is non-strict xfail until it has (a pass shows
as XPASS; a failure cannot take the validated suite down).  Round 2 removes the marker."""
import pytest

import test_multipipe_cpu as cpu
from raftsql_b200.multipipe import make_engine_core

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="multi-group seam on the GPU engine: first hardware run pending")]


@pytest.fixture(autouse=True)
def engine_core(monkeypatch):
    monkeypatch.setattr(cpu, "make_oracle_multicore", make_engine_core)


def test_every_group_replicates_independently_and_in_order_gpu(tmp_path):
    cpu.test_every_group_replicates_independently_and_in_order(tmp_path)


def test_stopped_node_replays_every_groups_wal_and_catches_up_gpu(tmp_path):
    cpu.test_stopped_node_replays_every_groups_wal_and_catches_up(tmp_path)
