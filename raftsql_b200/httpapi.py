"""HTTP SQL API kept exactly as the reference's (reference httpapi.go:26-79): PUT body=SQL -> 204 or
400 + error text; GET body=SELECT -> rows as text or 400; anything else 405 with `Allow: PUT, GET`.
Out of the hot path; part of the config-1 plumbing.
"""
from __future__ import annotations

import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from .db import RaftDB


def _handler_for(rdb: RaftDB):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, *a):  # the reference only logs errors (httpapi.go:30-34)
            pass

        def _body(self) -> str:
            n = int(self.headers.get("Content-Length") or 0)
            return self.rfile.read(n).decode() if n else ""

        def _err(self, ex):  # dumpErr: 400 + error text (httpapi.go:30-34)
            msg = (str(ex) + "\n").encode()
            self.send_response(400)
            self.send_header("Content-Type", "text/plain; charset=utf-8")
            self.send_header("Content-Length", str(len(msg)))
            self.end_headers()
            self.wfile.write(msg)

        def do_PUT(self):  # httpapi.go:38-49
            try:
                errc = rdb.Propose(self._body())
                err, _ = errc.recv()
            except Exception as ex:
                err = ex
            if err is not None:
                self._err(err)
            else:
                self.send_response(204)
                self.send_header("Content-Length", "0")
                self.end_headers()

        def do_GET(self):  # httpapi.go:51-62 (the query travels in the request body)
            try:
                v = rdb.Query(self._body()).encode()
            except Exception as ex:
                self._err(ex)
                return
            self.send_response(200)
            self.send_header("Content-Length", str(len(v)))
            self.end_headers()
            self.wfile.write(v)

        def _not_allowed(self):  # httpapi.go:63-67
            msg = b"Method not allowed\n"
            self.send_response(405)
            self.send_header("Allow", "PUT")
            self.send_header("Allow", "GET")
            self.send_header("Content-Length", str(len(msg)))
            self.end_headers()
            self.wfile.write(msg)

        do_POST = do_DELETE = do_PATCH = do_HEAD = _not_allowed

    return Handler


def ServeHttpSqlAPI(port: int, rdb: RaftDB, *, background: bool = False):
    """reference httpapi.go:71-79.  background=True returns the server (tests) instead of blocking forever."""
    srv = ThreadingHTTPServer(("", port), _handler_for(rdb))  # Addr ":" + port (httpapi.go:73)
    if background:
        threading.Thread(target=srv.serve_forever, daemon=True).start()
        return srv
    srv.serve_forever()
