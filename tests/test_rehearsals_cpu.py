"""Rehearsals: host-side code that will next run on a GPU, run here first against tests/engine_double.py (the CPU
oracle behind the Engine methods / C-ABI calls involved, decoding posted byte-form frames with `mrq_unpack8`, the
same inline decode the device kernel runs):

  * bench.py's byte-form end-to-end leg (`run_e2e8_child`): the REAL orchestration code — frame building with
    `mrq_pack8` as the trace is generated, the pipelined post / tick / drain loop, the commit-advance accumulation
    and the equality verdict (including a deliberately wrong decode, to show the verdict has teeth);
  * bench.py's whole N = 1 main path (`run_ours`) and the driver-facing contract of the JSON line it prints;
  * the bodies of the GPU tests written after round 1's GPU budget was spent (tests/test_zz_packed8_gpu.py,
    tests/test_zz_kat_gpu.py), so that when they first meet hardware a failure implicates the device path alone,
    not the test code.

What this cannot cover is the device kernels themselves and real PCIe timing."""
import argparse
import json

import pytest

from engine_double import FakeEngine, FakePinned


def test_byte_form_leg_end_to_end_on_the_double(monkeypatch, capsys):
    import bench
    import raftsql_b200
    import raftsql_b200.packed as packed

    made = []

    def make(*a, **kw):
        made.append(FakeEngine(*a, **kw))
        return made[-1]

    monkeypatch.setattr(bench, "G_TOTAL", 6000)
    monkeypatch.setattr(raftsql_b200, "Engine", make)
    monkeypatch.setattr(packed, "PinnedArray", FakePinned)
    bench.run_e2e8_child(argparse.Namespace(steps=7))
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert res["equals_wide_form"] is True
    assert res["steps"] == 7 and res["escapes"] == 0 and res["value"] > 0
    assert res["h2d_bytes_per_step"] == 6000 * bench.R  # (R-1) sender bytes + 1 proposal byte per group
    assert res["d2h_bytes_per_step"] == 6000
    assert made[0].L.posts == 3 + 7 + 7  # the three runs of the leg really went through the byte-form post
    # and the verdict is not vacuous: the trace commits entries on every one of those ticks
    st = made[0].o.export()
    assert (st["committed"] > bench.steady_state(6000, bench.R, 0, bench.SEED)["committed"]).mean() > 0.9


def test_a_wrong_decode_is_caught_by_the_legs_own_verdict(monkeypatch, capsys):
    """the leg's run-time check has teeth: a decode that is off by one entry on one sender makes it say False"""
    import bench
    import raftsql_b200
    import raftsql_b200.packed as packed

    class OffByOne(FakeEngine):
        def post_inbox_packed(self, word, prop8=None, wide=(), slot=0):
            super().post_inbox_packed(word, prop8, wide, slot)
            ack = (self.slots[slot]["type"][1] & 0x0F) == 4
            self.slots[slot]["index"][1][ack] += 1

    monkeypatch.setattr(bench, "G_TOTAL", 3000)
    monkeypatch.setattr(raftsql_b200, "Engine", OffByOne)
    monkeypatch.setattr(packed, "PinnedArray", FakePinned)
    bench.run_e2e8_child(argparse.Namespace(steps=5))
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert res["equals_wide_form"] is False


@pytest.mark.parametrize("G,R,cfg", [(1500, 7, 5), (1001, 5, 3), (300, 2, 5), (64, 1, 2)])
def test_rehearse_gpu_test_byte_form_decode_and_ticks(monkeypatch, G, R, cfg):
    import test_zz_packed8_gpu as t

    monkeypatch.setattr(t, "Engine", FakeEngine)
    t.test_byte_form_decodes_like_the_host_and_ticks_like_the_oracle(G, R, cfg)


@pytest.mark.parametrize("G,R,cfg", [(1200, 7, 5), (900, 5, 3), (300, 2, 5), (800, 3, 2)])
def test_rehearse_gpu_test_tick_mode_3(monkeypatch, G, R, cfg):
    import test_zz_packed8_gpu as t

    monkeypatch.setattr(t, "Engine", FakeEngine)
    t.test_tick_mode_3_consumes_the_bytes_itself(G, R, cfg)


def test_rehearse_gpu_test_sliding_window(monkeypatch):
    import test_zz_packed8_gpu as t

    monkeypatch.setattr(t, "Engine", FakeEngine)
    t.test_device_window_slides_by_itself_for_hundreds_of_ticks()


def test_rehearse_bench_main_path_and_line_assembly(monkeypatch, capsys):
    """`python bench.py` at N = 1, rehearsed: the real run_ours() — dry run, rehearsal, warm-up, timed region, post-roll,
    roofline arithmetic, the legs' results merged, the engine closed BEFORE the byte-form child is consulted, one JSON
    line printed — over the engine double, with the GPU-only legs stubbed.  Guards the driver-facing contract of the
    line (keys, types) against edits made without a GPU at hand."""
    import torch

    import bench
    import raftsql_b200
    from engine_double import FakeBenchEngine

    order = []

    class Eng(FakeBenchEngine):
        def close(self):
            order.append("close")

    def child(steps):
        order.append("child")
        return {"value": 9000.0, "unit": "ticks/s", "steps": steps, "h2d_bytes_per_step": 5 << 20, "d2h_bytes_per_step": 1 << 20,
                "equals_wide_form": True, "escapes": 0, "api": "8-bit"}

    monkeypatch.setattr(bench, "G_TOTAL", 2048)
    monkeypatch.setattr(raftsql_b200, "Engine", Eng)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(bench, "bench_quorum_kernel", lambda *a: {"bound": "hbm", "achieved": 1.0, "peak": 2.0, "frac": 0.5})
    monkeypatch.setattr(bench, "bench_e2e", lambda *a, **k: {"value": 4467.0, "unit": "ticks/s", "h2d_bytes_per_step": 11 << 20,
                                                            "d2h_bytes_per_step": 1 << 20, "steps": 6, "api": "16-bit",
                                                            "packed_equals_wide": True})
    monkeypatch.setattr(bench, "cpu_reference_ticks", lambda *a, **k: (80.0, 8, 3, 0.04, None))
    monkeypatch.setattr(bench, "e2e8_from_child", child)
    monkeypatch.delenv("MRQ_BENCH_FAST", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    args = argparse.Namespace(gpus=1, steps=6, warmup=3, impl="ours", gather="fused", tick_mode=None, l2=None, graph="auto")
    bench.run_ours(args)
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 1, "exactly one JSON line"
    line = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in line, k
    assert line["metric"] == "raft_ticks_per_sec_1Mx5" and line["unit"] == "ticks/s" and line["n_gpus"] == 1
    assert line["steps"] == 6 and line["warmup"] == 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and abs(line["value"] - 1e3 / line["ms_per_step"]) < 1e-6 * line["value"]
    assert line["gpu_launches"] == 2 * 6  # the fast + slow kernel of each timed tick, counted by the engine
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["value"] == 9000.0 and line["e2e"]["api"] == "8-bit" and line["e2e"]["packed8"]["equals_wide_form"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"])
    assert order == ["close", "child"], "the engine must be gone before the byte-form child gets the GPU"


def test_rehearse_bench_inbox_bytes_replays_the_same_ticks(monkeypatch, capsys):
    """`bench.py --inbox bytes` (tick mode 3, byte frames kept in their slots and replayed after each rewind) must
    drive the engine through exactly the ticks of the default run: same commit indices at the end of the timed region,
    no escapes on the steady-state trace, and the byte-form byte count in the roofline."""
    import torch

    import bench
    import raftsql_b200
    from engine_double import FakeBenchEngine

    finals = {}
    for inbox in ("wide", "bytes"):
        made = []

        class Eng(FakeBenchEngine):
            def timer_stop(self, _inbox=inbox):  # the end of the timed region (the untimed post-roll ticks on after it)
                finals[_inbox] = self.o.export()["committed"].copy()
                return super().timer_stop()

        def make(*a, **kw):
            made.append(Eng(*a, **kw))
            return made[-1]

        monkeypatch.setattr(bench, "G_TOTAL", 2048)
        monkeypatch.setattr(raftsql_b200, "Engine", make)
        monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
        monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
        monkeypatch.setenv("MRQ_BENCH_FAST", "1")  # kernels only: no e2e / cpu legs, no child
        monkeypatch.setattr(bench, "bench_quorum_kernel", lambda *a: {"bound": "hbm"})
        args = argparse.Namespace(gpus=1, steps=6, warmup=3, impl="ours", gather="fused", tick_mode=None, l2=None, graph="auto",
                                  inbox=inbox)
        bench.run_ours(args)
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        if inbox == "bytes":
            assert "0 escaped" in line["config"]["inbox"] and "tick mode 3" in line["config"]["inbox"]
            assert line["roofline"]["algorithmic_bytes_per_group"]["total"] == 173
            assert "tick_fast8_kernel" in line["roofline"]["kernel"] and line["roofline"]["traffic"] is None
        else:
            assert line["roofline"]["algorithmic_bytes_per_group"]["total"] == 217
    import numpy as np

    assert np.array_equal(finals["wide"], finals["bytes"]), "the byte-form run must commit exactly what the wide run commits"
    assert (finals["wide"] > bench.steady_state(2048, bench.R, 0, bench.SEED)["committed"]).mean() > 0.9


def _engine_kat_cases():
    import test_zz_kat_gpu as t

    cases = []
    for name in sorted(n for n in dir(t) if n.startswith("test_")):
        fn = getattr(t, name)
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        if not marks:
            cases.append(pytest.param(name, (), id=name))
            continue
        # expand the parametrize marks the same way pytest does (outer product, innermost mark first)
        import itertools

        names, values = [], []
        for m in marks:
            ns = [x.strip() for x in m.args[0].split(",")]
            names.append(ns)
            values.append([v if isinstance(v, (tuple, list)) and len(ns) > 1 else (v,) for v in m.args[1]])
        for combo in itertools.product(*values):
            kw = {}
            for ns, vs in zip(names, combo):
                kw.update(dict(zip(ns, vs)))
            cases.append(pytest.param(name, tuple(sorted(kw.items())), id=f"{name}-{'-'.join(str(v) for _, v in sorted(kw.items()))}"))
    return cases


@pytest.mark.parametrize("name,kwargs", _engine_kat_cases())
def test_rehearse_engine_kats(monkeypatch, name, kwargs):
    """the bodies of tests/test_zz_kat_gpu.py (upstream's tables through the engine, one message per tick) on the
    engine double: the adapter's way of raising MsgHup / MsgProp through the ABI must reproduce every table row"""
    import test_zz_kat_gpu as t

    monkeypatch.setattr(t, "Engine", FakeEngine)
    getattr(t, name)(**dict(kwargs))
