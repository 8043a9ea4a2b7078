"""Contract checks that need no GPU: the ABI header is valid plain C (what cgo compiles), the bench's reference arm
prints the JSON line the driver expects, and the repo layout promised in DESIGN.md exists."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c99(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include "mrq.h"\n#include "mrq_trace.h"\n#include "mrq_packed8.h"\n'
                   "int main(void) { mrq_config c; mrq_msg m; mrq_state s; mrq_inbox_packed p; mrq_counters k;\n"
                   "  (void)c; (void)m; (void)s; (void)p; (void)k;\n"
                   "  mrq_p8_cell a = mrq_p8_decode(mrq_p8_encode(4u, 7u, 1020u, 0u, 1000u, 7u), 1000u);\n"
                   "  int bad = a.value != 1020u || mrq_p8_next_base(1000u, a.pay) != 1004u || mrq_p8_row(3u, 2u, 5u) != 2u;\n"
                   "  mrq_trace_params t = mrq_trace_preset(5);\n"
                   "  return (int)(sizeof(mrq_msg) != 48) + (t.lagging_pct != 20) + bad; }\n")
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-o", str(exe), str(src)])
    assert subprocess.run([str(exe)]).returncode == 0


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "raft_ticks_per_sec_1Mx5" and line["unit"] == "ticks/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["groups_total"] == 1 << 20 and line["config"]["replicas"] == 5


def test_bench_byte_accounting_and_the_traffic_guard(tmp_path, monkeypatch):
    """the roofline's algorithmic bytes per group-tick for every form bench.py can time, and the rule that an ncu traffic
    figure captured from another build of the kernel is refused instead of quoted"""
    sys.path.insert(0, ROOT)
    import bench

    assert bench.tick_bytes_per_group(5, "wide")["total"] == 217
    assert bench.tick_bytes_per_group(5, "bytes")["total"] == 173
    c1 = bench.tick_bytes_per_group(5, "compact", 1)
    assert c1 == {"read": 42, "write": 40, "total": 82}
    c20 = bench.tick_bytes_per_group(5, "compact", 20, True)
    assert abs(c20["read"] - (5 + 37 / 20)) < 0.01 and c20["write"] == 40  # frame + state/K read; write-through writes it all
    c20e = bench.tick_bytes_per_group(5, "compact", 20, False)
    assert abs(c20e["total"] - (5 + 5 + (37 + 35) / 20)) < 0.02
    assert bench.quorum_bytes_per_group(5) == 56 and bench.quorum_bytes_per_group(7) == 72
    # traffic guard
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    (tmp_path / "profiles" / "r02_traffic.json").write_text(json.dumps(
        {"k": {"dram_read_bytes": 100, "dram_write_bytes": 23, "sass_instructions": 1000}}))
    monkeypatch.setattr(bench, "sass_instruction_count", lambda sub: 1000)
    assert bench.ncu_traffic("k", "k") == 123
    monkeypatch.setattr(bench, "sass_instruction_count", lambda sub: 1001)
    assert bench.ncu_traffic("k", "k") is None, "a capture of another kernel build must not be quoted"
    assert bench.ncu_traffic("missing", "k") is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_oracle_selftest_under_asan_ubsan():
    """The checker itself is run under AddressSanitizer + UBSan (oracle/selftest.c)."""
    cc = subprocess.run(["make", "-s", "--no-print-directory", "-C", os.path.join(ROOT, "oracle"), "asan_cc"], capture_output=True,
                        text=True).stdout.strip()
    if not cc:
        pytest.skip("no C compiler with the sanitizer runtimes on this machine")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "selftest_asan"], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "oracle", "selftest_asan")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "oracle selftest: ok" in r.stdout, r.stdout + r.stderr


def test_layout():
    for p in ("include/mrq.h", "include/mrq_trace.h", "oracle/raft_oracle.c", "oracle/Makefile", "tests/golden/upstream_kats.json",
              "tests/golden/make_golden.py", "profiles/r01_launches.md", "profiles/r01_results.md", "DESIGN.md", "INTEGRATION.md",
              "bench.py", "__graft_entry__.py", "raftsql_b200/csrc/mrq_kernels.cuh", "raftsql_b200/csrc/host/raftpipe.hpp",
              "go/mrq/mrq.go", "go/raftpipe_mrq.go"):
        assert os.path.exists(os.path.join(ROOT, p)), p
    assert not os.path.exists(os.path.join(ROOT, "raftsql_b200", "models"))  # not an ML framework layout
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi and "*.so" in gi
