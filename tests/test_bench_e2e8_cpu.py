"""Host-side code of the byte-form inbox rehearsed on the CPU against tests/engine_double.py (the CPU oracle
behind the Engine methods / C-ABI calls involved, decoding posted frames with `mrq_unpack8`, the same inline decode
the device kernel runs):

  * bench.py's byte-form end-to-end leg (`run_e2e8_child`): the REAL orchestration code — frame building with
    `mrq_pack8` as the trace is generated, the pipelined post / tick / drain loop, the commit-advance accumulation
    and the equality verdict;
  * the bodies of the byte-form GPU tests (tests/test_zz_packed8_gpu.py), so that when they first meet hardware a
    failure implicates the device path alone, not the test code.

What this cannot cover is the device kernel itself and real PCIe timing."""
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


def test_rehearse_gpu_test_sliding_window(monkeypatch):
    import test_zz_packed8_gpu as t

    monkeypatch.setattr(t, "Engine", FakeEngine)
    t.test_device_window_slides_by_itself_for_hundreds_of_ticks()
