"""raftsql server — the reference's CLI (reference server/main.go:24-38) over this repo's seam:

    python -m raftsql_b200.server --id 1 --cluster http://127.0.0.1:12379,http://127.0.0.1:22379,http://127.0.0.1:32379 --port 12380

Flags and defaults are the reference's (`--cluster` comma separated peer URLs, `--id` 1-based node id, `--port`
SQL port; server/main.go:25-27); the Procfile at the repo root is the reference's (Procfile:2-4).  Below the seam
sits one GPU engine (G = 1, R = len(cluster)) and the HTTP peer transport; the consensus arithmetic runs on
cuda:<--device>.  Without a GPU the process exits with the engine's error: there is no CPU fallback.
"""
from __future__ import annotations

import argparse

from .db import NewDB
from .httpapi import ServeHttpSqlAPI
from .raftpipe import Chan, NewRaftPipe
from .transport import HttpTransport


def main(argv=None):
    ap = argparse.ArgumentParser(description="raftsql on the mrq engine")
    ap.add_argument("--cluster", default="http://127.0.0.1:9021", help="comma separated cluster peers")  # main.go:25
    ap.add_argument("--id", type=int, default=1, help="node ID")  # main.go:26
    ap.add_argument("--port", type=int, default=9121, help="sql server port")  # main.go:27
    ap.add_argument("--device", type=int, default=0, help="CUDA device ordinal of this node's engine")
    ap.add_argument("--tick-ms", type=float, default=100.0, help="raft tick period (raft.go:207: 100 ms)")
    args = ap.parse_args(argv)
    peers = args.cluster.split(",")
    proposeC = Chan()  # main.go:30
    transport = HttpTransport(peers)
    rp = NewRaftPipe(args.id, peers, proposeC, transport=transport, tick_seconds=args.tick_ms / 1e3, device=args.device)  # main.go:34
    try:
        ServeHttpSqlAPI(args.port, NewDB(f"raftsql-{args.id}.db", rp))  # main.go:37
    finally:
        proposeC.close()  # main.go:31 (defer)
        transport.close()


if __name__ == "__main__":
    main()
