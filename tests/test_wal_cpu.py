"""The host's write-ahead log (raftsql_b200/hostnode.py Wal; stands in for etcd `wal`, reference raft.go:100-124,228):
what was saved is what is read back, conflict truncation replays, and a crash mid-append (torn tail) loses only
the torn record."""
import os
import struct

from raftsql_b200.hostnode import Wal


def _saved(tmp_path):
    w = Wal(str(tmp_path / "raftsql-1"))
    assert not Wal.exist(w.dir)
    w.open()
    w.save((1, 1, 0), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")], 1)
    w.save((1, 1, 2), [], 0)
    w.close()
    assert Wal.exist(w.dir)
    return w


def test_round_trip_and_last_hardstate_wins(tmp_path):
    w = _saved(tmp_path)
    hs, ents = Wal(w.dir).read_all()
    assert hs == (1, 1, 2)
    assert ents == [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")]


def test_truncate_then_append_replays_as_the_conflict_resolution_it_was(tmp_path):
    w = _saved(tmp_path)
    w.open()
    w.save((2, 2, 2), [(2, b"other")], 3, truncate_after=2)  # entry 3 replaced by a new leader's entry
    w.close()
    hs, ents = Wal(w.dir).read_all()
    assert hs == (2, 2, 2)
    assert ents == [(1, b""), (1, b"CREATE"), (2, b"other")]


def test_torn_tail_records_are_dropped_and_the_prefix_stands(tmp_path):
    w = _saved(tmp_path)
    good = os.path.getsize(w.path)
    # (a) a length prefix that outruns the file
    with open(w.path, "ab") as f:
        f.write(struct.pack("<I", 1 << 30) + b'{"e":[4,1,')
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])
    # (b) a complete length but bytes that are not a record (the sector never reached the disk)
    with open(w.path, "r+b") as f:
        f.truncate(good)
        f.seek(good)
        f.write(struct.pack("<I", 6) + b"\0\0\0\0\0\0")
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])
    # (c) only half a length prefix
    with open(w.path, "r+b") as f:
        f.truncate(good)
        f.seek(good)
        f.write(b"\x07\x00")
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])


def test_missing_wal_reads_as_empty(tmp_path):
    assert Wal(str(tmp_path / "nothing")).read_all() == (None, [])
