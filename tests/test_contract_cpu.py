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
    assert line["config"]["groups"] == 1 << 20 and line["config"]["replicas"] == 5


def test_byte_form_leg_can_only_add_to_the_bench_line():
    """bench.py measures the byte-form inbox in a child process and merges its outcome: adopted only when it
    verified itself and is faster, recorded otherwise, and nothing the child returns can break the line."""
    sys.path.insert(0, ROOT)
    import bench

    def e2e():
        return {"value": 4467.0, "unit": "ticks/s", "h2d_bytes_per_step": 11_534_336, "d2h_bytes_per_step": 1 << 20, "steps": 20,
                "api": "16-bit", "packed_equals_wide": True}

    good = {"value": 9000.0, "unit": "ticks/s", "steps": 20, "h2d_bytes_per_step": 5 << 20, "d2h_bytes_per_step": 1 << 20,
            "equals_wide_form": True, "escapes": 0, "api": "8-bit"}
    e = e2e()
    bench.merge_packed8(e, good)
    assert e["value"] == 9000.0 and e["api"] == "8-bit" and e["h2d_bytes_per_step"] == 5 << 20 and e["packed8"] is good
    assert e["packed_equals_wide"] is True
    e = e2e()
    bench.merge_packed8(e, dict(good, value=3000.0))  # verified but slower: recorded, not adopted
    assert e["value"] == 4467.0 and e["api"] == "16-bit" and e["packed8"]["value"] == 3000.0
    e = e2e()
    bench.merge_packed8(e, dict(good, equals_wide_form=False))  # wrong answers: never adopted, and flagged
    assert e["value"] == 4467.0 and e["packed_equals_wide"] is False
    for junk in ({"error": "exit 1: boom"}, {}, None, [1, 2], "text", {"equals_wide_form": True}):
        e = e2e()
        bench.merge_packed8(e, junk)
        assert e["value"] == 4467.0 and e["api"] == "16-bit" and "packed8" in e
        json.dumps(e)
    # and on a machine without a GPU the child itself fails cleanly
    r = bench.e2e8_from_child(4)
    assert set(r) == {"error"} and "no CUDA device" in r["error"]


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
