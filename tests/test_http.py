"""The microservice endpoint (micro_raytracer_b200/http.py ≙ src/http.rs): status lines of the
validation chain without a GPU; a real render round trip with one."""
import json
import socket

import numpy as np
import pytest

from micro_raytracer_b200.http import HttpServer


@pytest.fixture()
def server():
    srv = HttpServer(("127.0.0.1", 0))
    srv.start_in_thread()
    yield srv.server_address
    srv.shutdown()
    srv.server_close()


def _raw(addr, payload: bytes) -> bytes:
    with socket.create_connection(addr, timeout=30) as s:
        s.sendall(payload)
        s.shutdown(socket.SHUT_WR)
        out = b""
        while True:
            c = s.recv(1 << 20)
            if not c:
                return out
            out += c


@pytest.mark.parametrize("req,status", [
    (b"POST / HTTP/1.0\r\nContent-Type: application/json\r\nContent-Length: 2\r\n\r\n{}", b"505"),
    (b"GET / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 2\r\n\r\n{}", b"405"),
    (b"POST / HTTP/1.1\r\nContent-Length: 2\r\n\r\n{}", b"400"),
    (b"POST / HTTP/1.1\r\nContent-Type: text/plain\r\nContent-Length: 2\r\n\r\n{}", b"415"),
    (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\n\r\n{}", b"411"),
    (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 5\r\n\r\n{}", b"400"),
    (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 9\r\n\r\n{\"rt\": 1]", b"400"),
])
def test_validation_chain_matches_http_rs(server, req, status):
    """http.rs:73-113, in the reference's order."""
    assert _raw(server, req).startswith(b"HTTP/1.1 " + status)


@pytest.mark.gpu
def test_post_json_returns_the_rendered_jpeg(server):
    import io
    from PIL import Image
    body = json.dumps({"rt": {"sample": 4}, "frame": {"res": [96, 54]},
                       "scene": {"renderer": [{"type": "sphere", "r": 0.5}], "light": [{"type": "point", "pos": [-0.5, -1, 0.5]}]}}).encode()
    req = b"POST /render HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: " + str(len(body)).encode() + b"\r\n\r\n" + body
    res = _raw(server, req)
    head, _, payload = res.partition(b"\r\n\r\n")
    assert head.startswith(b"HTTP/1.1 200 OK") and b"Content-Type: image/jpeg" in head
    n = int([l for l in head.split(b"\r\n") if l.startswith(b"Content-Length")][0].split(b": ")[1])
    img = np.asarray(Image.open(io.BytesIO(payload[:n])).convert("RGB"))
    assert img.shape == (54, 96, 3) and img.max() > 100 and img[0, 0].max() < 10
