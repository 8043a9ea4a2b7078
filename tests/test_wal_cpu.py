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
        f.write(struct.pack("<II", 1 << 30, 0) + b'{"e":[4,1,')
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])
    # (b) a complete length but bytes that are not a record (the sector never reached the disk)
    with open(w.path, "r+b") as f:
        f.truncate(good)
        f.seek(good)
        f.write(struct.pack("<II", 6, 0) + b"\0\0\0\0\0\0")
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])
    # (c) only half a length prefix
    with open(w.path, "r+b") as f:
        f.truncate(good)
        f.seek(good)
        f.write(b"\x07\x00")
    assert Wal(w.dir).read_all() == ((1, 1, 2), [(1, b""), (1, b"CREATE"), (1, b"INSERT-0")])


def test_a_torn_tail_is_cut_off_on_open_so_later_records_survive_the_next_restart(tmp_path):
    """read_all stopped at a torn record but open() appended BEHIND it, hiding every later record (votes,
    acknowledged entries) from the next replay.  Now open() truncates to the valid prefix first."""
    w = _saved(tmp_path)
    good = os.path.getsize(w.path)
    for garbage in (struct.pack("<II", 1 << 30, 0) + b'{"e":[4,1,', struct.pack("<II", 6, 0) + b"\0\0\0\0\0\0", b"\x07\x00"):
        with open(w.path, "r+b") as f:
            f.truncate(good)
            f.seek(good)
            f.write(garbage)
        w2 = Wal(w.dir)                      # crash, restart #1
        assert w2.read_all()[0] == (1, 1, 2)
        w2.open()
        assert os.path.getsize(w.path) == good, "the torn tail must be cut off before anything is appended"
        w2.save((3, 2, 2), [(3, b"after-the-crash")], 4)
        w2.close()
        hs, ents = Wal(w.dir).read_all()     # restart #2 finds what restart #1 saved
        assert hs == (3, 2, 2) and ents[-1] == (3, b"after-the-crash") and len(ents) == 4
        with open(w.path, "r+b") as f:
            f.truncate(good)


def test_a_flipped_bit_inside_a_record_ends_the_valid_prefix(tmp_path):
    w = _saved(tmp_path)
    size = os.path.getsize(w.path)
    with open(w.path, "r+b") as f:  # corrupt one payload byte of the LAST record (the hardstate (1,1,2))
        f.seek(size - 3)
        b = f.read(1)
        f.seek(size - 3)
        f.write(bytes([b[0] ^ 0x01]))
    hs, ents = Wal(w.dir).read_all()
    assert hs == (1, 1, 0) and len(ents) == 3  # the CRC rejects the damaged record; the prefix stands


def test_missing_wal_reads_as_empty(tmp_path):
    assert Wal(str(tmp_path / "nothing")).read_all() == (None, [])
