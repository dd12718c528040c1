"""The native C++ host (micro_raytracer_b200/host/ -> the `raytrace` binary): the reference's
front-end (src/bin/raytrace.rs, src/cli.rs, src/parser.rs, src/http.rs) in compiled code over the
C ABI.  CPU tests pin its description parsing to the Python mirror (which the golden scenes pin to
the reference's example files) byte for byte on the packed C-ABI arrays; GPU tests check that it
renders the same images through libmrt.so."""
import base64
import ctypes as C
import gzip
import io
import json
import os
import socket
import subprocess
import time

import numpy as np
import pytest

import micro_raytracer_b200 as mrt
from micro_raytracer_b200 import abi, cli
from util import ROOT, SCENES

BIN = os.path.join(ROOT, "micro_raytracer_b200", "raytrace")
ALL_SCENES = ["Default", "CornellBox", "CornellBox2", "Mesh", "Instance", "Minecraft", "dof"]


def run(*args, check=True, **kw):
    p = subprocess.run([BIN, *map(str, args)], capture_output=True, text=True, timeout=600, **kw)
    if check:
        assert p.returncode == 0, p.stderr
    return p


def packed_bytes(render) -> bytes:
    """The layout `raytrace --dump-packed` writes (host/render.cpp PackedScene::bytes + frame + rt)."""
    p = mrt.pack_scene(render.scene)
    s = p.c

    def raw(arr, n):
        return bytes(memoryview(arr))[: n * C.sizeof(arr._type_)]
    out = np.asarray([s.n_objects, s.n_instances, s.n_textures, p.n_texels, s.n_meshes, p.n_triangles, s.n_lights, 0], np.uint32).tobytes()
    out += raw(p.objects, s.n_objects) + raw(p.instances, s.n_instances) + raw(p.textures, s.n_textures)
    out += p.texels[: p.n_texels].astype(np.float32).tobytes()
    out += raw(p.meshes, s.n_meshes) + p.triangles[: p.n_triangles].astype(np.float32).tobytes() + raw(p.lights, s.n_lights)
    out += bytes(memoryview(s.sky_color)) + np.float32(s.sky_pwr).tobytes()
    out += bytes(memoryview(render.frame.pack()))
    out += np.asarray([render.rt.bounce, render.rt.sample], np.uint32).tobytes() + np.float32(render.rt.loss).tobytes()
    return out


def native_packed(tmp_path, *args) -> bytes:
    f = tmp_path / "packed.bin"
    run("--dry", "--dump-packed", f, *args)
    return f.read_bytes()


def test_binary_is_built():
    assert os.path.exists(BIN), "build it with __graft_entry__.build() (make -C micro_raytracer_b200/host)"


@pytest.mark.parametrize("name", ALL_SCENES)
def test_example_scenes_pack_identically_to_the_python_host(tmp_path, name):
    """JSON defaults, hex colours, instances, PNG texture files, .obj meshes: same bytes into mrt_set_scene."""
    path = os.path.join(SCENES, name + ".json")
    assert native_packed(tmp_path, path) == packed_bytes(mrt.load_render(path))


README_CORNELLBOX2 = """--bounce 8 --sample 512 --loss 0.15 --res 1080 1080 --ssaa 2
 --cam pos: 0 -1.25 0 fov: 60 gamma: 0.6 exp: 0.8
 --obj sphere pos: 0 0 -0.1 r: 0.15
 --obj box size: 0.25 0.25 0.25 pos: 0 0 -0.375 dir: 0 0.5 0.5 0
 --obj box size: 0.3 0.3 0.01 pos: 0 0 0.499 emit: 1
 --obj box size: 1 0.01 1 pos: 0 0.5 0
 --obj box size: 1 1 0.01 pos: 0 0 0.5
 --obj box size: 1 1 0.01 pos: 0 0 -0.5
 --obj box size: 0.01 1 1 pos: -0.5 0 0 albedo: #ff0000
 --obj box size: 0.01 1 1 pos: 0.5 0 0 albedo: #00ff00""".split()


def test_readme_command_is_cornellbox2_json(tmp_path):
    """README.md:14-27 -> example/CornellBox2.json: mini-grammar, reverse-order quirk (parser.rs:584-598)."""
    got = native_packed(tmp_path, *README_CORNELLBOX2)
    want = native_packed(tmp_path, os.path.join(SCENES, "CornellBox2.json"))
    assert got == want
    _, _, r = cli.parse_render(README_CORNELLBOX2)
    assert got == packed_bytes(r)


@pytest.mark.parametrize("argv", [
    "--obj sphere --light point: -0.5 -1 0.5",
    "--light dir: 0.3 0 -2 pwr: 0.7 col: #ff8000 --light pt: 1 2 3 col: 0.1 0.2 0.3 --obj pln n: 0 0.1 1 pos: 0 0 -1 rough: 0.3",
    "--obj tri vtx: 0 0 0 1 0 0 0 0 1 metal: 1 --obj mesh name: m glass: 0.1 opacity: 0.5 mesh: 0 0 0 1 0 0 0 0 1 0 1 0 1 1 0 0 1 1",
    "--sky 0.2 0.3 0.4 0.9 --obj box --cam pos: 1 -2 3 dir: -0.25 0 1 0 fov: 55 gamma: 0.5 exp: 0.1 aprt: 0.01 foc: 2 --res 64 48 --ssaa 1.5",
])
def test_mini_grammar_matches_the_python_front_end(tmp_path, argv):
    args = argv.split()
    _, _, r = cli.parse_render(args)
    assert native_packed(tmp_path, *args) == packed_bytes(r)


def test_precedence_and_replacement_rules(tmp_path):
    """CLI::parse_render, cli.rs:78-153: --cam replaces the camera, --scene the scene, --obj extends."""
    full = tmp_path / "full.json"
    full.write_text(json.dumps({"rt": {"bounce": 3, "sample": 9}, "frame": {"res": [64, 48], "ssaa": 2, "cam": {"fov": 50, "gamma": 0.5}},
                                "scene": {"renderer": [{"type": "plane", "n": [0, 0, 1]}], "sky": {"color": [0.1, 0.2, 0.3], "pwr": 0.9}}}))
    scene = tmp_path / "scene.json"
    scene.write_text(json.dumps({"renderer": [{"type": "sphere", "r": 1}]}))
    for args in ([str(full), "--sample", "4", "--ssaa", "1", "--cam", "pos:", "1", "2", "3", "--obj", "box", "--sky", "1", "1", "1", "0.25"],
                 [str(full), "--scene", str(scene)], [str(full), "--loss", "0.5", "--bounce", "2"]):
        _, _, r = cli.parse_render(args)
        assert native_packed(tmp_path, *args) == packed_bytes(r)


def test_inline_base64_gzip_assets(tmp_path):
    """parser.rs:620-628, 674-682: textures and meshes as base64(gzip(json))."""
    rng = np.random.default_rng(3)
    tex = {"w": 4, "h": 2, "dat": (rng.integers(0, 256, (8, 3)) / 255.0).round(6).tolist()}
    tris = rng.uniform(-1, 1, (5, 3, 3)).round(4).tolist()
    enc = lambda o: base64.b64encode(gzip.compress(json.dumps(o).encode())).decode()
    d = {"scene": {"renderer": [{"type": "mesh", "mesh": enc(tris), "mat": {"rough": 0.5}},
                                {"type": "box", "sizes": [1, 1, 1], "mat": {"tex": enc(tex), "emap": enc(tex), "omap": {"w": 1, "h": 1}}}]}}
    f = tmp_path / "inline.json"
    f.write_text(json.dumps(d))
    assert native_packed(tmp_path, f) == packed_bytes(mrt.load_render(str(f)))


def test_errors_use_the_reference_messages(tmp_path):
    for args, msg in (("--obj torus", "`torus` type is unxpected!"), ("--obj sphere size: 1 1 1", "param for `sphere` is unxpected"),
                      ("--obj sphere pos: 1 2", "unexpected ends!"), ("--obj sphere r: abc", "should be <f32>!"),
                      ("--obj box albedo: red", "should be <f32>!"), ("--light sun: 1 2 3", "type is unxpected")):
        p = run("--dry", *args.split(), check=False)
        assert p.returncode == 1 and p.stderr.startswith("cli: ") and msg in p.stderr, (args, p.stderr)
    bad = tmp_path / "bad.json"
    bad.write_text('{"scene": {"renderer": [{"type": "sphere", "r": 1, "mat": {"albedo": "ff0000"}}]}}')
    p = run("--dry", bad, check=False)
    assert p.returncode == 1 and "is not a hex color!" in p.stderr
    bad.write_text('{"rt": {"sample": 4,}}')
    assert run("--dry", bad, check=False).returncode == 1
    p = run("--dry", "-v", "--pretty", "--obj", "sphere")
    assert json.loads(p.stdout)["scene"]["renderer"][0]["type"] == "sphere"


def test_json_reader_and_verbose_dump_round_trip(tmp_path):
    """The host's own JSON reader (stands in for serde_json): escapes, exponents, nesting, explicit nulls (serde
    Option), key order kept; `-v` dumps the merged description, which must parse back to the same packing."""
    src = tmp_path / "odd.json"
    src.write_text('{ "rt" : {"bounce": 3, "sample": 2.0, "loss": 1.5e-1},\n "frame": {"res": [6.4e1, 48], "ssaa": 1, "cam": null},\n'
                   ' "scene": {"renderer": [{"type": "sphere", "r": 5E-1, "name": "q\\"uo\\u0074e\\n", "mat": {"albedo": "#FF8000", "tex": null}},'
                   ' {"type": "plane", "n": [0, 0, 1], "pos": [-0.0, 0, -1e0]}], "light": null, "sky": {"pwr": 0.25}}}')
    p = run("--dry", "-v", src)
    dumped = json.loads(p.stdout)
    assert dumped["scene"]["renderer"][0]["name"] == 'q"uote\n' and list(dumped) == ["rt", "frame", "scene"]
    back = tmp_path / "back.json"
    back.write_text(p.stdout)
    a, b = native_packed(tmp_path, src), native_packed(tmp_path, back)
    assert a == b == packed_bytes(mrt.load_render(str(src)))
    for bad in ('{"rt": {"bounce": 1} trailing', '{"rt": {"bounce": 01x}}', '{"rt": [1, 2,]}', '{"a": "\\q"}', ""):
        src.write_text(bad)
        assert run("--dry", src, check=False).returncode == 1, bad


def test_image_encoders_round_trip(tmp_path):
    """PNG / PPM are lossless, JPEG decodes close to the source (stand-ins for the image crate at cli.rs:168,174, http.rs:122)."""
    from PIL import Image
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:45, 0:70]
    img = np.stack([xx * 3, yy * 5, (xx + yy) * 2], axis=2).astype(np.uint8)
    img[10:20, 10:30] = rng.integers(0, 256, (10, 20, 3))
    src = tmp_path / "src.png"
    Image.fromarray(img).save(src)
    for ext in ("png", "ppm"):
        out = tmp_path / f"o.{ext}"
        run("--convert", src, out)
        assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), img)
    run("--convert", tmp_path / "o.ppm", tmp_path / "back.png")
    assert np.array_equal(np.asarray(Image.open(tmp_path / "back.png")), img)
    out = tmp_path / "o.jpg"
    run("--convert", src, out)
    jpg = np.asarray(Image.open(out).convert("RGB")).astype(np.float64)
    smooth = np.ones(img.shape[:2], bool)
    smooth[8:22, 8:32] = False
    assert np.abs(jpg - img)[smooth].mean() < 2.0
    gray = tmp_path / "gray.png"
    Image.fromarray(img[:, :, 0]).save(gray)
    p = run("--convert", gray, tmp_path / "x.png", check=False)
    assert p.returncode == 1 and "is not rgb888 image!" in p.stderr  # parser.rs:664


# ----------------------------------------------------------------------------- HTTP endpoint
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture()
def native_server():
    port = _free_port()
    proc = subprocess.Popen([BIN, "--http", f"127.0.0.1:{port}"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for _ in range(100):
        try:
            socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
            break
        except OSError:
            time.sleep(0.05)
    yield ("127.0.0.1", port)
    proc.kill()
    proc.wait()


def _raw(addr, payload: bytes) -> bytes:
    with socket.create_connection(addr, timeout=60) as s:
        s.sendall(payload)
        s.shutdown(socket.SHUT_WR)
        out = b""
        while True:
            c = s.recv(1 << 20)
            if not c:
                return out
            out += c


def test_http_validation_chain_matches_http_rs(native_server):
    """http.rs:73-113, in the reference's order."""
    for req, status in [
        (b"POST / HTTP/1.0\r\nContent-Type: application/json\r\nContent-Length: 2\r\n\r\n{}", b"505"),
        (b"GET / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 2\r\n\r\n{}", b"405"),
        (b"POST / HTTP/1.1\r\nContent-Length: 2\r\n\r\n{}", b"400"),
        (b"POST / HTTP/1.1\r\nContent-Type: text/plain\r\nContent-Length: 2\r\n\r\n{}", b"415"),
        (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\n\r\n{}", b"411"),
        (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 5\r\n\r\n{}", b"400"),
        (b"POST / HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: 9\r\n\r\n{\"rt\": 1]", b"400"),
    ]:
        assert _raw(native_server, req).startswith(b"HTTP/1.1 " + status), req


# ----------------------------------------------------------------------------- GPU: same pixels as the library
@pytest.mark.gpu
@pytest.mark.parametrize("name,extra", [("Default", []), ("dof", []), ("Mesh", []), ("CornellBox2", ["--ssaa", "2"])])
def test_native_front_end_renders_the_library_image(tmp_path, name, extra):
    from PIL import Image
    path = os.path.join(SCENES, name + ".json")
    args = [path, "--res", "96", "54", "--sample", "4", *extra]
    out = tmp_path / "o.png"
    p = run(*args, "-v", "-o", out)
    assert "cli:done" in p.stdout
    got = np.asarray(Image.open(out).convert("RGB"))
    _, _, r = cli.parse_render(args)
    s = mrt.Sampler(device=0)
    s.execute(r.scene, r.frame, r.rt, 4)
    assert np.array_equal(got, s.img(r.frame))
    run(*args, "--update", "-o", tmp_path / "u.ppm")  # one pass per call + a frame per pass (cli.rs:162-169)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "u.ppm").convert("RGB")), got)


@pytest.mark.gpu
def test_native_http_returns_the_rendered_jpeg(native_server):
    from PIL import Image
    d = {"rt": {"sample": 4}, "frame": {"res": [96, 54]},
         "scene": {"renderer": [{"type": "sphere", "r": 0.5}], "light": [{"type": "point", "pos": [-0.5, -1, 0.5]}]}}
    body = json.dumps(d).encode()
    req = b"POST /render HTTP/1.1\r\nContent-Type: application/json\r\nContent-Length: " + str(len(body)).encode() + b"\r\n\r\n" + body
    res = _raw(native_server, req)
    head, _, payload = res.partition(b"\r\n\r\n")
    assert head.startswith(b"HTTP/1.1 200 OK") and b"Content-Type: image/jpeg" in head
    n = int([l for l in head.split(b"\r\n") if l.startswith(b"Content-Length")][0].split(b": ")[1])
    img = np.asarray(Image.open(io.BytesIO(payload[:n])).convert("RGB")).astype(np.float64)
    r = mrt.render_from_dict(d)
    s = mrt.Sampler(device=0)
    s.execute(r.scene, r.frame, r.rt, 4)
    want = s.img(r.frame).astype(np.float64)
    assert img.shape == want.shape and np.abs(img - want).mean() < 2.0


# ----------------------------------------------------------------------------- GPU: several devices, one process
def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_native_multi_gpu_render_equals_the_single_gpu_one(tmp_path):
    """--gpus N: ONE Sampler over a device group (mrt_create_group): the library splits the samples over the
    N devices and img() gathers their films over peer mappings.  The counter-based RNG keys on the global sample
    index, so only the f32 summation order differs from the one-GPU image."""
    from PIL import Image
    path = os.path.join(SCENES, "CornellBox2.json")
    args = [path, "--res", "128", "128", "--ssaa", "2", "--sample", "16"]
    if _n_devices() < 2:
        p = run(*args, "--gpus", "2", "-o", tmp_path / "x.png", check=False)
        assert p.returncode == 1 and p.stderr.startswith("cli: ")  # no second device: a clean error, never a fallback
        pytest.skip("needs two GPUs")
    run(*args, "-o", tmp_path / "one.png")
    for g in (2, min(_n_devices(), 8)):
        p = run(*args, "--gpus", g, "-v", "-o", tmp_path / f"g{g}.png")
        assert f"on {g} gpus" in p.stdout
        a = np.asarray(Image.open(tmp_path / "one.png")).astype(int)
        b = np.asarray(Image.open(tmp_path / f"g{g}.png")).astype(int)
        assert np.abs(a - b).max() <= 1 and (a == b).mean() > 0.999


# ----------------------------------------------------------------------------- the host as a library
def _build_api_example(tmp_path):
    exe = tmp_path / "api_example"
    pkg = os.path.join(ROOT, "micro_raytracer_b200")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-pthread", os.path.join(ROOT, "tests", "native", "api_example.cpp"),
           "-I", os.path.join(pkg, "host"), "-L", pkg, "-lmrt_host", "-lmrt", "-lz", "-ldl", f"-Wl,-rpath,{pkg}", "-o", str(exe)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return exe


def test_host_library_links_and_fails_loudly_without_a_gpu(tmp_path):
    """libmrt_host.a + the headers of micro_raytracer_b200/host are a usable C++ API (mrt_host::Sampler ≙ src/sampler.rs).
    Without a device the program reports the error and exits 1 — there is no CPU path to fall back to."""
    exe = _build_api_example(tmp_path)
    from util import have_gpu
    if have_gpu():
        pytest.skip("GPU present: covered by test_host_library_api")
    p = subprocess.run([exe, os.path.join(SCENES, "Default.json"), "2", tmp_path / "o.ppm"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "no CUDA device" in p.stderr and not (tmp_path / "o.ppm").exists()


@pytest.mark.gpu
def test_host_library_api(tmp_path):
    """Pass-by-pass execute + img through mrt_host::Sampler give the image the Python mirror gives."""
    from PIL import Image
    exe = _build_api_example(tmp_path)
    out = tmp_path / "o.ppm"
    p = subprocess.run([exe, os.path.join(SCENES, "dof.json"), "3", out], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "passes 3" in p.stdout
    r = mrt.load_render(os.path.join(SCENES, "dof.json"))
    r.frame.res = (96, 54)
    s = mrt.Sampler(device=0)
    for _ in range(3):
        s.execute(r.scene, r.frame, r.rt)
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), s.img(r.frame))
