"""The C-ABI boundary, checked without a GPU: the library loads, exports every symbol include/mrq.h
declares, the ctypes binding covers every one of them, and engine creation FAILS LOUDLY when there is no
CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from raftsql_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrq.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mrq_[a-z0-9_]+)\s*\(", src)
    # typedef'd struct names never appear followed by "(": what is left are function declarations
    return sorted(set(names))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_is_built_in_tree():
    assert os.path.exists(F.LIB_PATH), "run `make` (or __graft_entry__.build()) first"


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_functions()
    assert len(declared) >= 35
    nm = subprocess.check_output(["nm", "-D", "--defined-only", F.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT\s+(mrq_[a-z0-9_]+)", nm))
    missing = [n for n in declared if n not in exported]
    assert not missing, f"declared in mrq.h but not exported by libmrq.so: {missing}"
    unbound = [n for n in declared if n not in F.SIGNATURES]
    assert not unbound, f"declared in mrq.h but missing from the ctypes binding: {unbound}"
    extra = [n for n in F.SIGNATURES if n not in declared]
    assert not extra, f"bound but not declared in mrq.h: {extra}"
    # nothing but the C-ABI leaks out of the library
    leaked = [s for s in re.findall(r"\bT\s+(\S+)", nm) if not s.startswith("mrq_") and not s.startswith("_")]
    assert not leaked, leaked


def test_library_loads_and_reports_sm100():
    L = F.load()
    arch = C.c_uint32()
    assert L.mrq_version(C.byref(arch)) == F.MRQ_ABI_VERSION
    assert arch.value == 100


def test_struct_layouts_match_the_header():
    assert C.sizeof(F.Msg) == 48
    assert C.sizeof(F.TraceParams) == 40
    assert C.sizeof(F.Counters) == 64
    cfg = F.Config()
    F.load().mrq_config_default(C.byref(cfg))
    # the reference's raft.Config constants (raft.go:154-155)
    assert (cfg.abi_version, cfg.election_tick, cfg.heartbeat_tick) == (F.MRQ_ABI_VERSION, 10, 1)


def test_built_for_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", F.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_bad_config_is_rejected_with_text():
    L = F.load()
    cfg = F.Config()
    L.mrq_config_default(C.byref(cfg))
    cfg.n_replicas = 9
    h = C.c_void_p()
    assert L.mrq_create(C.byref(cfg), C.byref(h)) == F.MRQ_E_INVAL
    assert b"n_replicas" in L.mrq_last_error(None)
    assert not h.value


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_no_gpu_means_loud_failure_not_fallback():
    import raftsql_b200

    with pytest.raises(raftsql_b200.MrqError) as ei:
        raftsql_b200.Engine(16, 3)
    assert ei.value.code == F.MRQ_E_NODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "raftsql_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle-free", ""), f"{f} mentions the oracle"
