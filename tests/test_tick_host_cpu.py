"""The DEVICE source of the tick, compiled for the host and run against the oracle (tests/cpp/tick_host_test.cpp).

raftsql_b200/csrc/mrq_kernels.cuh is included unchanged; tests/cpp/device_on_host.hpp supplies the CUDA vocabulary
and MRQ_HOST_EMULATION swaps the cache-hinted PTX accessors for plain loads/stores.  What runs is, line for line,
what one GPU thread runs for one group — `fast_group_tick<R>` and, for the groups it declines, `general_group_tick<R>`
— over host arrays in the engine's layout, for R = 1..8 on the election, lag + churn, heartbeat-heavy and steady
traces, every state column and the out word compared with the oracle after every tick.

It cannot see launch geometry, warp collectives or memory ordering (the `-m gpu` parity tests do); it checks the
arithmetic of both tick paths where there is no GPU, which is where device code gets written between GPU sessions.
Product builds never define MRQ_HOST_EMULATION: the library's SASS is byte-identical with and without these hooks
(checked with cuobjdump when they were added)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_headers():
    inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    return inc if os.path.exists(os.path.join(inc, "cuda_runtime.h")) else None


def _run(target, *argv):
    inc = _cuda_headers()
    if inc is None:
        pytest.skip("CUDA headers not found (set CUDA_INC)")
    subprocess.check_call(["make", "-C", ROOT, f"CUDA_INC={inc}", target], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, target), *argv], capture_output=True, text=True, timeout=1200)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "tick_host_test: ok" in out, out[-3000:]
    assert "runtime error" not in out and "AddressSanitizer" not in out, out[-3000:]
    return out


def test_device_tick_source_on_the_host_equals_the_oracle():
    out = _run("tests/cpp/tick_host_test")
    lines = [ln.split() for ln in out.splitlines() if ln.strip().startswith("G=")]
    assert len(lines) >= 20
    # both paths really ran: the fast path took the bulk, the general path the elections / churn / vote traffic
    fast = sum(int(ln[ln.index("fast") + 1]) for ln in lines)
    general = sum(int(ln[ln.index("general") + 1]) for ln in lines)
    assert fast > 1_000_000 and general > 100_000


def test_device_tick_source_under_asan_and_ubsan():
    """no undefined behaviour in the device arithmetic (shifts, wraps, conversions) and no access outside the
    [R][gs] columns by either tick function"""
    cxx = subprocess.run(["make", "-s", "--no-print-directory", "-C", ROOT, "san_cxx"], capture_output=True, text=True).stdout.strip()
    if not cxx:
        pytest.skip("no C++ compiler with the sanitizer runtimes on this machine")
    _run("tests/cpp/tick_host_test_san")
