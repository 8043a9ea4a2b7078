"""The C++ host side above the C-ABI (raftsql_b200/csrc/host/: Chan, HostNode, Wal, LocalTransport,
NewRaftPipe) — the seam of reference raftpipe.go:3-17 in the reference's own kind of language (compiled).
tests/cpp/raftpipe_test.cpp runs the scenarios (nil sentinel / order / Close protocol on one node; a 3-node
in-process cluster; stop + WAL replay + catch-up, after raftsql_test.go:92-171); here it is built and run with
the CPU oracle as the consensus core, and under `-m gpu` with the real engine through libmrq.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "raftpipe_test")


def _build():
    subprocess.check_call(["make", "-C", ROOT, "raftsql_b200/libraftpipe.so", "tests/cpp/raftpipe_test"],
                          stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)


def _run(core, tmp_path, env=None):
    _build()
    r = subprocess.run([BIN, core, str(tmp_path)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and "raftpipe_test: ok" in r.stdout, r.stdout + r.stderr


def test_host_library_exports_only_cxx_and_links_the_c_abi():
    _build()
    lib = os.path.join(ROOT, "raftsql_b200", "libraftpipe.so")
    needed = subprocess.check_output(["readelf", "-d", lib], text=True)
    assert "libmrq.so" in needed  # the host side sits ON the C-ABI library, it does not re-implement it
    syms = subprocess.check_output(["nm", "-D", "--undefined-only", lib], text=True)
    used = sorted({s.split()[-1] for s in syms.splitlines() if " mrq_" in s})
    assert {"mrq_create", "mrq_tick", "mrq_post_inbox_delta", "mrq_export_state", "mrq_sync_out"} <= set(used)
    assert not any("orc_" in s for s in syms.splitlines()), "the product host library must not touch the oracle"


def test_cpp_raftpipe_scenarios_oracle_core(tmp_path):
    _run("oracle", tmp_path)


@pytest.mark.gpu
def test_cpp_raftpipe_scenarios_gpu_engine(tmp_path):
    _run("engine", tmp_path)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="C++ multi-group seam over the GPU engine (EngineMultiCore): first hardware run pending; "
                                        "the same scenario passes over the oracle core in the CPU suite")
def test_cpp_multi_group_seam_gpu_engine(tmp_path):
    _run("engine", tmp_path, env={"MRQ_TEST_MULTI_GROUP": "1"})
