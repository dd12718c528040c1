"""The reference's microservice endpoint (src/http.rs) over the CUDA path: POST a render
description as `application/json` (<= 1 MiB), get `image/jpeg` (quality 90) back.

Mirrors HttpServer::{start, handle, raytrace} (http.rs:62-162): one thread per connection, one
Sampler per request (http.rs:138), the same status lines for the same faults — 505 (not HTTP/1.1),
405 (not POST), 400 (no Content-Type / length mismatch), 415 (not application/json), 411 (no
Content-Length).  A body that is not a valid description is answered 400 here; the reference only
logs it and drops the connection.  SURVEY.md §8(f) "next #3".
"""
from __future__ import annotations

import io
import json
import socketserver
import threading
import time
from typing import Optional, Tuple

from .sampler import MrtError
from .scene import SceneError, render_from_dict

MAX_REQUEST = 1024 * 1024  # http.rs:66: one read into a 1 MiB buffer


def render_jpeg(body: bytes, device: int = 0, log=None) -> bytes:
    """HttpServer::raytrace + the JPEG encode of handle(): http.rs:115-122, 136-148."""
    from PIL import Image

    from .sampler import Sampler
    render = render_from_dict(json.loads(body.decode("utf-8")))
    sampler = Sampler(24, 64, device=device)  # http.rs:138
    try:  # the context and its device buffers go away whatever happens below
        t0 = time.perf_counter()
        if render.rt.sample > 0:
            sampler.execute(render.scene, render.frame, render.rt, render.rt.sample)
        img = sampler.img(render.frame)
        if log:
            log(f"http:done: {time.perf_counter() - t0:.3f}s")
    finally:
        sampler.close()
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=90)
    return buf.getvalue()


class _Handler(socketserver.BaseRequestHandler):
    device = 0
    log = None

    def _status(self, line: str):
        self.request.sendall(f"HTTP/1.1 {line}\r\n".encode())

    def handle(self):
        data = b""
        self.request.settimeout(10.0)
        try:
            while b"\r\n\r\n" not in data and len(data) < MAX_REQUEST:
                chunk = self.request.recv(65536)
                if not chunk:
                    break
                data += chunk
            head, _, body = data.partition(b"\r\n\r\n")
            lines = head.decode("latin-1").split("\r\n")
            parts = lines[0].split(" ")
            if len(parts) < 3:
                return self._status("400 Bad Request")
            method, _uri, version = parts[0], parts[1], parts[2]
            headers = {}
            for ln in lines[1:]:
                k, sep, v = ln.partition(": ")
                if sep:
                    headers[k] = v
            if version != "HTTP/1.1":
                return self._status("505 HTTP Version Not Supported")
            if method != "POST":
                return self._status("405 Method Not Allowed")
            if "Content-Type" not in headers:
                return self._status("400 Bad Request")
            if not headers["Content-Type"].startswith("application/json"):
                return self._status("415 Unsupported Media Type")
            if "Content-Length" not in headers:
                return self._status("411 Length Required")
            try:
                n = int(headers["Content-Length"])
            except ValueError:
                return self._status("400 Bad Request")
            if n > MAX_REQUEST:
                return self._status("400 Bad Request")
            while len(body) < n:
                chunk = self.request.recv(65536)
                if not chunk:
                    break
                body += chunk
            if len(body) != n:
                return self._status("400 Bad Request")
            try:
                jpg = render_jpeg(body, self.device, self.log)
            except (SceneError, MrtError, ValueError, KeyError, TypeError) as e:  # MrtError: what the library rejects
                if self.log:
                    self.log(f"http: {e}")
                return self._status("400 Bad Request")
            self.request.sendall(b"HTTP/1.1 200 OK\r\nContent-Type: image/jpeg\r\nContent-Length: " + str(len(jpg)).encode()
                                 + b"\r\n\r\n" + jpg + b"\r\n")
        except OSError as e:
            if self.log:
                self.log(f"http: {e}")


class HttpServer(socketserver.ThreadingTCPServer):
    """≙ HttpServer{hlr: TcpListener}; serve_forever() ≙ start() (http.rs:150-162)."""
    allow_reuse_address = True
    daemon_threads = True

    def __init__(self, address: Tuple[str, int], device: int = 0, log=None):
        handler = type("Handler", (_Handler,), {"device": device, "log": staticmethod(log) if log else None})
        super().__init__(address, handler)

    def start_in_thread(self) -> threading.Thread:
        t = threading.Thread(target=self.serve_forever, daemon=True)
        t.start()
        return t


def parse_address(addr: str) -> Tuple[str, int]:
    host, _, port = addr.rpartition(":")
    return (host or "localhost", int(port))


def serve(addr: str, device: int = 0, log: Optional[callable] = print):
    srv = HttpServer(parse_address(addr), device, log)
    if log:
        log(f"http:listening: {addr}")
    srv.serve_forever()
