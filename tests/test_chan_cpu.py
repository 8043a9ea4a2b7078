"""Go channel semantics of raftsql_b200.raftpipe.Chan that the seam relies on (reference raftpipe.go:3-17,
raft.go:65-66,89-93): unbuffered rendezvous, FIFO across blocked senders, close, and a select-with-stop send that
withdraws its value when the stop side wins.  (The C++ Chan<T> is held to the same list in tests/cpp.)"""
import threading
import time

import pytest

from raftsql_b200.raftpipe import Chan, ChanClosed


def test_aborted_send_withdraws_its_value():
    ch, stop, out = Chan(), threading.Event(), []
    th = threading.Thread(target=lambda: out.append(ch.send(7, stop)))
    th.start()
    time.sleep(0.05)
    stop.set()
    th.join(2)
    assert out == [False]
    ch.close()
    assert ch.recv() == (None, False)  # the receiver never sees the withdrawn value


def test_unbuffered_rendezvous_and_fifo_across_senders():
    ch, done = Chan(), []
    s1 = threading.Thread(target=lambda: (ch.send(1), done.append(1)))
    s1.start()
    time.sleep(0.03)
    s2 = threading.Thread(target=lambda: (ch.send(2), done.append(2)))
    s2.start()
    time.sleep(0.03)
    assert done == []  # nobody has received yet
    assert ch.recv(timeout=2) == (1, True)
    assert ch.recv(timeout=2) == (2, True)
    s1.join(2), s2.join(2)
    assert sorted(done) == [1, 2]


def test_close_fails_a_blocked_sender_and_later_sends_but_buffered_values_drain():
    ch, errs = Chan(), []

    def sender():
        try:
            ch.send(5)
        except ChanClosed:
            errs.append("closed")

    th = threading.Thread(target=sender)
    th.start()
    time.sleep(0.03)
    ch.close()
    th.join(2)
    assert errs == ["closed"]
    assert ch.recv() == (None, False)
    with pytest.raises(ChanClosed):
        ch.send(6)
    buf = Chan(buffered=2)
    assert buf.send(1) and buf.send(2)
    buf.close()
    assert list(buf) == [1, 2]


def test_recv_timeout():
    with pytest.raises(TimeoutError):
        Chan().recv(timeout=0.05)
