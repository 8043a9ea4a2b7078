"""Config 1 over real sockets (BASELINE configs[0]; reference Procfile:2-4, README.md:19-25): every node has its OWN
HttpTransport instance (as separate processes would), peers talk over loopback TCP, clients use the HTTP SQL API.

CPU: three nodes in one process, oracle core (exercises transport, host node, db, http).  GPU: the real thing —
three `python -m raftsql_b200.server` processes, one GPU engine each, driven with PUT / GET exactly like the
reference's README does with curl."""
import http.client
import os
import socket
import subprocess
import sys
import time

import pytest

from oracle_core import make_oracle_core
from raftsql_b200.db import NewDB
from raftsql_b200.httpapi import ServeHttpSqlAPI
from raftsql_b200.raftpipe import Chan, NewRaftPipe
from raftsql_b200.transport import HttpTransport, decode, encode
from raftsql_b200.hostnode import Message

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_ports(n):
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def req(port, method, body, timeout=30):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=timeout)
    c.request(method, "/", body=body)
    r = c.getresponse()
    data = r.read().decode()
    c.close()
    return r.status, data


def wait_until(pred, timeout=40.0):
    end = time.monotonic() + timeout
    while time.monotonic() < end:
        try:
            if pred():
                return True
        except Exception:
            pass
        time.sleep(0.05)
    return False


def test_wire_format_roundtrip():
    msgs = [Message(3, 2, 1, term=7, logterm=6, index=41, commit=40, entries=[(6, b"INSERT \x00\xff"), (7, b"")]),
            Message(4, 1, 2, term=7, index=41, reject=True, reject_hint=12), Message(6, 3, 1, term=9, reject=False)]
    back = decode(encode(msgs))
    assert [(m.type, m.to, m.frm, m.term, m.logterm, m.index, m.commit, m.reject, m.reject_hint, m.entries) for m in back] == \
           [(m.type, m.to, m.frm, m.term, m.logterm, m.index, m.commit, m.reject, m.reject_hint, m.entries) for m in msgs]


def test_a_dead_peer_never_stalls_the_sender():
    """rafthttp drops what it cannot deliver (ReportUnreachable is a no-op in the reference, raft.go:271-273): a node's
    tick loop must not block on a peer that is down, however long it stays down."""
    dead = free_ports(1)[0]  # nobody listens there
    tr = HttpTransport([f"http://127.0.0.1:{free_ports(1)[0]}", f"http://127.0.0.1:{dead}"])
    t0 = time.monotonic()
    for k in range(2000):  # far more batches than the per-peer queue holds
        tr.send([Message(8, 2, 1, term=1, commit=k)])
    assert time.monotonic() - t0 < 2.0
    tr.close()


def test_three_nodes_over_loopback_tcp_oracle_core(tmp_path):
    raft_ports, sql_ports = free_ports(3), free_ports(3)
    peers = [f"http://127.0.0.1:{p}" for p in raft_ports]
    trs, dbs, srvs = [], [], []
    try:
        for i in range(3):
            tr = HttpTransport(peers)  # one per node, as one per process
            rp = NewRaftPipe(i + 1, peers, Chan(), transport=tr, tick_seconds=0.02, core_factory=make_oracle_core,
                             waldir=os.path.join(str(tmp_path), f"raftsql-{i + 1}"))
            trs.append(tr)
            dbs.append(NewDB(os.path.join(str(tmp_path), f"raftsql-{i + 1}.db"), rp))
            srvs.append(ServeHttpSqlAPI(sql_ports[i], dbs[i], background=True))
        # README.md:19-25
        assert req(sql_ports[0], "PUT", "CREATE TABLE main.t (id int primary key asc, nodeid text)")[0] == 204
        for i in range(3):
            assert req(sql_ports[i], "PUT", f'INSERT INTO main.t (nodeid) VALUES ("{i}")')[0] == 204
        for p in sql_ports:
            assert wait_until(lambda p=p: req(p, "GET", "SELECT * from main.t")[1].count("\n") == 3)
            st, v = req(p, "GET", "SELECT * from main.t")
            assert st == 200 and all(f"||{i}|" in v for i in range(3)), v
        st, v = req(sql_ports[1], "GET", "SELECT * from main.x")
        assert st == 400 and "no such table" in v
    finally:
        for s in srvs:
            s.shutdown()
            s.server_close()
        for db in dbs:
            db.Close()
        for tr in trs:
            tr.close()


@pytest.mark.gpu
def test_procfile_cluster_three_processes_gpu(tmp_path):
    """The reference's Procfile, for real: three server processes (one GPU engine each), curl-style PUT / GET."""
    raft_ports, sql_ports = free_ports(3), free_ports(3)
    cluster = ",".join(f"http://127.0.0.1:{p}" for p in raft_ports)
    env = dict(os.environ, PYTHONPATH=ROOT)
    procs = []
    try:
        for i in range(3):
            procs.append(subprocess.Popen([sys.executable, "-m", "raftsql_b200.server", "--id", str(i + 1), "--cluster", cluster,
                                           "--port", str(sql_ports[i]), "--tick-ms", "20"], cwd=str(tmp_path), env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        assert wait_until(lambda: all(req(p, "GET", "SELECT 1", timeout=2)[0] == 200 for p in sql_ports), timeout=120), \
            "servers did not come up"
        assert req(sql_ports[0], "PUT", "CREATE TABLE main.t (id int primary key asc, nodeid text)", timeout=60)[0] == 204
        for i in range(3):
            assert req(sql_ports[i], "PUT", f'INSERT INTO main.t (nodeid) VALUES ("{i}")', timeout=60)[0] == 204
        for p in sql_ports:
            assert wait_until(lambda p=p: req(p, "GET", "SELECT * from main.t")[1].count("\n") == 3)
            v = req(p, "GET", "SELECT * from main.t")[1]
            assert all(f"||{i}|" in v for i in range(3)), v
        assert req(sql_ports[2], "POST", "x")[0] == 405
    finally:
        for p in procs:
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
