#!/usr/bin/env python
"""Does tests/cpp/tick_host_test.cpp (the device tick source compiled for the host, checked against the oracle) have
teeth?  Inject one bug at a time into a scratch copy of raftsql_b200/csrc/mrq_kernels.cuh and require the harness
to fail.  Takes a couple of minutes (one host compile per mutant), so it is a tool, not part of the suite:

    python tools/mutate_tick_host.py          # every mutant must be reported as CAUGHT
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = os.environ.get("HOST_CXX", "/usr/bin/g++")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")

MUTANTS = [
    ("fast path: commit gate off by one",
     "          if (mci > committed && mci >= gate && mci <= last_index) {\n            committed = mci;",
     "          if (mci > committed && mci > gate && mci <= last_index) {\n            committed = mci;"),
    ("32-bit median-of-5 network: one max turned into min",
     "    const uint32_t f = max(lo_ab, lo_cd), g = min(hi_ab, hi_cd);",
     "    const uint32_t f = min(lo_ab, lo_cd), g = min(hi_ab, hi_cd);"),
    ("fast path: follower ignores the heartbeat's commit index",
     "            if (committed < mx[r]) {",
     "            if (committed < mx[r] && false) {"),
    ("general path: vote tally drops the last voter",
     "const uint32_t granted = __popc(votes & 0x5555u);",
     "const uint32_t granted = __popc(votes & 0x1555u);"),
    # (forcing b = 0 in delta32 is an EQUIVALENT mutant: a match behind `committed` then trips `bad` and the exact
    #  64-bit network answers instead — slower, same result.  This one is not equivalent:)
    ("delta32: a match behind the commit index is not clamped to zero",
     "  d = ((b | hi) == 0u) ? lo : 0u;",
     "  d = lo;"),
    ("byte-form tick: an escaped sender is taken for 'no message' by the fast path",
     "  if (w == MRQ_P8_ESCAPE) ty = kTypeEscaped;",
     ""),
    ("byte-form tick: the fast path forgets to slide the window",
     "(nb != bi ? D_GATE : 0u)",
     "0u"),
    ("byte-form tick: the general path clobbers an escaped sender's wide message",
     "    if (w == MRQ_P8_ESCAPE) continue;  // the wide message is already in the slot",
     ""),
    ("leader timers: heartbeat flag never raised on the fast path",
     "          out |= MRQ_OUT_BCAST_HEARTBEAT;\n        }\n      }\n    } else if (m.role == MRQ_ROLE_FOLLOWER) {",
     "        }\n      }\n    } else if (m.role == MRQ_ROLE_FOLLOWER) {"),
]


def main():
    src = open(os.path.join(ROOT, "raftsql_b200", "csrc", "mrq_kernels.cuh")).read()
    tmp = tempfile.mkdtemp(prefix="mrq_mut_")
    try:
        for sub in ("tests/cpp", "raftsql_b200/csrc", "oracle", "include"):
            os.makedirs(os.path.join(tmp, sub))
        for f in ("tests/cpp/tick_host_test.cpp", "tests/cpp/device_on_host.hpp", "oracle/raft_oracle.c", "oracle/raft_oracle.h"):
            shutil.copy(os.path.join(ROOT, f), os.path.join(tmp, f))
        for f in os.listdir(os.path.join(ROOT, "include")):
            shutil.copy(os.path.join(ROOT, "include", f), os.path.join(tmp, "include", f))
        missed = 0
        for name, a, b in MUTANTS:
            if a not in src:
                print(f"STALE   {name}: the text to mutate is no longer in the header")
                missed += 1
                continue
            open(os.path.join(tmp, "raftsql_b200", "csrc", "mrq_kernels.cuh"), "w").write(src.replace(a, b, 1))
            exe = os.path.join(tmp, "t")
            c = subprocess.run([CXX, "-std=c++17", "-O1", f"-I{CUDA_INC}", "-pthread", "-o", exe, "tests/cpp/tick_host_test.cpp",
                                "oracle/raft_oracle.c"], cwd=tmp, capture_output=True, text=True)
            if c.returncode != 0:
                print(f"NOBUILD {name}\n{c.stderr[-500:]}")
                missed += 1
                continue
            r = subprocess.run([exe, "quick"], capture_output=True, text=True)
            caught = r.returncode != 0
            print(f"{'CAUGHT ' if caught else 'MISSED '} {name}: {r.stdout.strip().splitlines()[-1]}")
            missed += 0 if caught else 1
        return 1 if missed else 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
