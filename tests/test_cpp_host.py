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


@pytest.mark.parametrize("san,env", [("asan", {"ASAN_OPTIONS": "detect_leaks=1"}), ("tsan", {"TSAN_OPTIONS": "halt_on_error=1"})])
def test_cpp_host_under_sanitizers(tmp_path, san, env):
    """The same scenarios with the C++ host compiled under AddressSanitizer + UBSan (+ leak check) and under
    ThreadSanitizer: the seam is threads and channels, so memory errors and data races are what to look for.
    (Oracle core: no device code runs in these builds.)"""
    cxx = subprocess.run(["make", "-s", "--no-print-directory", "-C", ROOT, "san_cxx"], capture_output=True, text=True).stdout.strip()
    if not cxx:
        pytest.skip("no C++ compiler with the sanitizer runtimes on this machine")
    exe = os.path.join(ROOT, "tests", "cpp", f"raftpipe_test_{san}")
    subprocess.check_call(["make", "-C", ROOT, f"tests/cpp/raftpipe_test_{san}"], stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, "oracle", str(tmp_path)], capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "raftpipe_test: ok" in out, out[-3000:]
    assert "ThreadSanitizer" not in out and "AddressSanitizer" not in out and "runtime error" not in out, out[-3000:]


@pytest.mark.gpu
def test_cpp_raftpipe_scenarios_gpu_engine(tmp_path):
    """every C++ scenario — single node, 3-node cluster + restart, and the multi-group seam — over the GPU engine"""
    _run("engine", tmp_path)
