// parser.cpp — host-side description parsing (see parser.hpp).  Follows the reference's
// src/parser.rs: serde defaults (188-271), hex colours (713-733), inline base64+gzip assets
// (620-628, 674-682), texture files (660-672), .obj meshes (602-618), instance expansion
// (838-853), the CLI mini-grammar (274-582) with its reverse-order split (584-598), and the merge
// precedence of CLI::parse_render (src/cli.rs:78-153).
#include "parser.hpp"

#include <zlib.h>

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "image_io.hpp"

namespace mrt_host {

// ----------------------------------------------------------------------------- small helpers
std::string dirname_of(const std::string& path) {
    const size_t s = path.rfind('/');
    if (s == std::string::npos) return ".";
    return s == 0 ? "/" : path.substr(0, s);
}
static std::string join_path(const std::string& base, const std::string& p) {
    if (!p.empty() && p[0] == '/') return p;
    if (base.empty()) return p;
    return base + "/" + p;
}
static std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(path + ": No such file or directory (os error 2)");
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static float num(const Json& j) { return (float)j.as_num(); }
static uint32_t u32_of(const Json& j, const char* what) {  // serde: usize / u32 fields reject negatives and fractions
    const double v = j.as_num();
    if (!(v >= 0.0 && v <= 4294967295.0) || v != std::floor(v)) throw Error(std::string("invalid value for `") + what + "`: expected an unsigned integer");
    return (uint32_t)v;
}

template <size_t N>
static std::array<float, N> vec(const Json& j, const char* what) {
    if (!j.is_arr() || j.size() != N) throw Error(std::string(what) + ": expected " + std::to_string(N) + " numbers");
    std::array<float, N> v{};
    for (size_t i = 0; i < N; i++) v[i] = num(j.items()[i]);
    return v;
}
static void flatten(const Json& j, std::vector<float>& out) {
    if (j.is_num()) { out.push_back(num(j)); return; }
    if (!j.is_arr()) throw Error("invalid type: expected a number or an array");
    for (const Json& e : j.items()) flatten(e, out);
}

// ColorWrapper::unwrap, parser.rs:713-733: "#rrggbb" or [r, g, b]
static Vec3 color(const Json& j, const char* what) {
    if (j.is_str()) {
        const std::string& s = j.as_str();
        if (s.empty() || s[0] != '#') throw Error(s + " is not a hex color!");
        const std::string hex = s.substr(1, 6);
        char* end = nullptr;
        const unsigned long n = std::strtoul(hex.c_str(), &end, 16);
        if (hex.empty() || end != hex.c_str() + hex.size()) throw Error("invalid digit found in string");
        return {(float)((n >> 16) & 0xFF) / 255.0f, (float)((n >> 8) & 0xFF) / 255.0f, (float)(n & 0xFF) / 255.0f};
    }
    return vec<3>(j, what);
}

// ----------------------------------------------------------------------------- inline assets
static std::string base64_decode(const std::string& s) {
    std::string o;
    uint32_t acc = 0;
    int n = 0;
    for (unsigned char ch : s) {
        int v;
        if (ch >= 'A' && ch <= 'Z') v = ch - 'A';
        else if (ch >= 'a' && ch <= 'z') v = ch - 'a' + 26;
        else if (ch >= '0' && ch <= '9') v = ch - '0' + 52;
        else if (ch == '+') v = 62;
        else if (ch == '/') v = 63;
        else if (ch == '=' || ch == '\n' || ch == '\r' || ch == ' ') continue;
        else throw Error(std::string("inline asset: Invalid byte ") + std::to_string((int)ch));
        acc = (acc << 6) | (uint32_t)v;
        n += 6;
        if (n >= 8) { n -= 8; o += (char)((acc >> n) & 0xFF); }
    }
    return o;
}
static std::string gunzip(const std::string& z) {
    z_stream s{};
    if (inflateInit2(&s, 16 + MAX_WBITS) != Z_OK) throw Error("inline asset: zlib init failed");
    s.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(z.data()));
    s.avail_in = (uInt)z.size();
    std::string out;
    char buf[1 << 16];
    int rc;
    do {
        s.next_out = reinterpret_cast<Bytef*>(buf);
        s.avail_out = sizeof buf;
        rc = inflate(&s, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&s); throw Error("inline asset: invalid gzip header"); }
        out.append(buf, sizeof buf - s.avail_out);
    } while (rc != Z_STREAM_END);
    inflateEnd(&s);
    return out;
}
static Json inline_json(const std::string& s) { return Json::parse(gunzip(base64_decode(s))); }

// TextureWrapper (untagged: buffer | inline base64 | file), parser.rs:86-92, 660-696
static TexturePtr texture(const Json& v, const std::string& base_dir) {
    if (v.is_obj()) {
        auto t = std::make_shared<Texture>();
        if (const Json* w = v.find("w")) t->w = u32_of(*w, "w");
        if (const Json* h = v.find("h")) t->h = u32_of(*h, "h");
        if (const Json* d = v.find("dat")) { t->has_dat = true; flatten(*d, t->dat); }
        return t;
    }
    if (v.is_str()) {
        const std::string& s = v.as_str();
        if (s.find('.') == std::string::npos) return texture(inline_json(s), base_dir);
        const Image im = load_image_rgb8(join_path(base_dir, s));  // TextureWrapper::load: RGB8 / 255
        auto t = std::make_shared<Texture>();
        t->w = im.w; t->h = im.h; t->has_dat = true;
        t->dat.resize(im.rgb.size());
        for (size_t i = 0; i < im.rgb.size(); i++) t->dat[i] = (float)im.rgb[i] / 255.0f;
        return t;
    }
    throw Error("bad texture");
}

// MeshWrapper::load, parser.rs:602-618: first object, first group, position indices, the first
// three vertices of every polygon.
static std::vector<float> mesh_obj(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw Error(path + ": No such file or directory (os error 2)");
    std::vector<std::array<float, 3>> pos;
    std::vector<std::array<long, 3>> tris;
    bool group_open = true;
    int n_groups = 0;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") {
            std::array<float, 3> p{};
            std::string a, b, c;
            ss >> a >> b >> c;
            p[0] = (float)std::strtod(a.c_str(), nullptr); p[1] = (float)std::strtod(b.c_str(), nullptr); p[2] = (float)std::strtod(c.c_str(), nullptr);
            pos.push_back(p);
        } else if (tag == "o" || tag == "g") {
            n_groups++;
            group_open = n_groups <= 1 || tris.empty();
        } else if (tag == "f" && group_open) {
            std::array<long, 3> idx{};
            for (int k = 0; k < 3; k++) {
                std::string tok;
                if (!(ss >> tok)) throw Error(path + ": face with fewer than 3 vertices");
                const long i = std::strtol(tok.substr(0, tok.find('/')).c_str(), nullptr, 10);
                idx[k] = i > 0 ? i - 1 : (long)pos.size() + i;
            }
            tris.push_back(idx);
        }
    }
    std::vector<float> out;
    out.reserve(tris.size() * 9);
    for (const auto& t : tris)
        for (int k = 0; k < 3; k++) {
            if (t[k] < 0 || (size_t)t[k] >= pos.size()) throw Error(path + ": vertex index out of range");
            for (int c = 0; c < 3; c++) out.push_back(pos[(size_t)t[k]][c]);
        }
    return out;
}
static std::vector<float> mesh(const Json& v, const std::string& base_dir) {
    if (v.is_str()) {
        const std::string& s = v.as_str();
        if (s.find('.') != std::string::npos) return mesh_obj(join_path(base_dir, s));
        return mesh(inline_json(s), base_dir);
    }
    std::vector<float> out;
    flatten(v, out);
    if (out.size() % 9 != 0) throw Error("mesh: expected triangles of 3 x 3 numbers");
    return out;
}

// ----------------------------------------------------------------------------- JSON -> Render
static Material material(const Json* d, const std::string& base_dir) {
    Material m;
    if (!d) return m;
    if (const Json* a = d->find("albedo")) m.albedo = color(*a, "albedo");
    if (const Json* v = d->find("rough")) m.rough = num(*v);
    if (const Json* v = d->find("metal")) m.metal = num(*v);
    if (const Json* v = d->find("glass")) m.glass = num(*v);
    if (const Json* v = d->find("opacity")) m.opacity = num(*v);
    if (const Json* v = d->find("emit")) m.emit = num(*v);
    if (const Json* v = d->find("tex")) m.tex = texture(*v, base_dir);
    if (const Json* v = d->find("rmap")) m.rmap = texture(*v, base_dir);
    if (const Json* v = d->find("mmap")) m.mmap = texture(*v, base_dir);
    if (const Json* v = d->find("gmap")) m.gmap = texture(*v, base_dir);
    if (const Json* v = d->find("omap")) m.omap = texture(*v, base_dir);
    if (const Json* v = d->find("emap")) m.emap = texture(*v, base_dir);
    return m;
}
static const Json& need(const Json& d, const char* key) {
    const Json* v = d.find(key);
    if (!v) throw Error(std::string("missing field `") + key + "`");
    return *v;
}
// RendererWrapper + unwrap, parser.rs:130-150, 826-864
static Renderer renderer(const Json& d, const std::string& base_dir) {
    Renderer r;
    const Json* ty = d.find("type");
    const std::string kind = ty && ty->is_str() ? ty->as_str() : "";
    if (kind == "sphere") { r.kind = Kind::Sphere; r.r = num(need(d, "r")); }
    else if (kind == "plane") { r.kind = Kind::Plane; r.n = vec<3>(need(d, "n"), "n"); }
    else if (kind == "box") { r.kind = Kind::Box; r.sizes = vec<3>(need(d, "sizes"), "sizes"); }
    else if (kind == "triangle") {
        r.kind = Kind::Triangle;
        std::vector<float> v;
        flatten(need(d, "vtx"), v);
        if (v.size() != 9) throw Error("vtx: expected 3 x 3 numbers");
        for (int i = 0; i < 9; i++) r.vtx[(size_t)i] = v[(size_t)i];
    } else if (kind == "mesh") { r.kind = Kind::Mesh; r.mesh = mesh(need(d, "mesh"), base_dir); }
    else throw Error("unknown variant `" + kind + "`, expected one of `sphere`, `plane`, `box`, `triangle`, `mesh`");
    r.mat = material(d.find("mat"), base_dir);
    if (const Json* n = d.find("name")) r.name = n->as_str();
    const Json* pos = d.find("pos");
    const Json* dir = d.find("dir");
    auto own = [&]() {
        Instance in;
        if (pos) in.pos = vec<3>(*pos, "pos");
        if (dir) in.dir = vec<4>(*dir, "dir");
        return in;
    };
    if (const Json* inst = d.find("inst")) {
        if (pos || dir) r.instance.push_back(own());  // parser.rs:841-843: prepended
        for (const Json& e : inst->items()) {
            if (!e.is_arr() || e.size() != 2) throw Error("inst: expected [pos, dir] pairs");
            Instance in;
            in.pos = vec<3>(e.items()[0], "inst pos");
            in.dir = vec<4>(e.items()[1], "inst dir");
            r.instance.push_back(in);
        }
    } else {
        r.instance.push_back(own());
    }
    return r;
}
static Light light(const Json& d) {
    Light l;
    const Json* ty = d.find("type");
    const std::string kind = ty && ty->is_str() ? ty->as_str() : "";
    if (kind == "point") { l.kind = MRT_LIGHT_POINT; l.v = vec<3>(need(d, "pos"), "pos"); }
    else if (kind == "dir") { l.kind = MRT_LIGHT_DIR; l.v = vec<3>(need(d, "dir"), "dir"); }
    else throw Error("unknown variant `" + kind + "`, expected `point` or `dir`");
    if (const Json* v = d.find("pwr")) l.pwr = num(*v);
    if (const Json* v = d.find("color")) l.color = color(*v, "color");
    return l;
}

Render render_from_json(const Json& d, const std::string& base_dir) {
    Render out;
    if (!d.is_obj()) throw Error("invalid type: expected a render description object");
    if (const Json* rt = d.find("rt")) {
        if (const Json* v = rt->find("bounce")) out.rt.bounce = u32_of(*v, "bounce");
        if (const Json* v = rt->find("sample")) out.rt.sample = u32_of(*v, "sample");
        if (const Json* v = rt->find("loss")) out.rt.loss = num(*v);
    }
    if (const Json* fr = d.find("frame")) {
        if (const Json* cam = fr->find("cam")) {
            Camera& c = out.frame.cam;
            if (const Json* v = cam->find("pos")) c.pos = vec<3>(*v, "cam pos");
            if (const Json* v = cam->find("dir")) c.dir = vec<4>(*v, "cam dir");
            if (const Json* v = cam->find("fov")) c.fov = num(*v);
            if (const Json* v = cam->find("gamma")) c.gamma = num(*v);
            if (const Json* v = cam->find("exp")) c.exp = num(*v);
            if (const Json* v = cam->find("aprt")) c.aprt = num(*v);
            if (const Json* v = cam->find("foc")) c.foc = num(*v);
        }
        if (const Json* res = fr->find("res")) {
            if (!res->is_arr() || res->size() != 2) throw Error("res: expected 2 numbers");
            for (int i = 0; i < 2; i++) {
                const double v = res->items()[(size_t)i].as_num();
                if (!(v >= 0.0 && v <= 65535.0)) throw Error("res does not fit u16");
                out.frame.res[(size_t)i] = (uint16_t)v;
            }
        }
        if (const Json* v = fr->find("ssaa")) out.frame.ssaa = num(*v);
    }
    if (const Json* sc = d.find("scene")) {
        if (const Json* rs = sc->find("renderer")) {
            out.scene.renderer.emplace();
            for (const Json& o : rs->items()) out.scene.renderer->push_back(renderer(o, base_dir));
        }
        if (const Json* ls = sc->find("light")) {
            out.scene.light.emplace();
            for (const Json& l : ls->items()) out.scene.light->push_back(light(l));
        }
        if (const Json* sky = sc->find("sky")) {
            if (const Json* v = sky->find("color")) out.scene.sky.color = color(*v, "color");
            if (const Json* v = sky->find("pwr")) out.scene.sky.pwr = num(*v);
        }
    }
    return out;
}
Render load_render(const std::string& path) { return render_from_json(Json::parse(slurp(path)), dirname_of(path)); }

// ----------------------------------------------------------------------------- CLI mini-grammar
namespace {
struct Peek {
    const std::vector<std::string>& t;
    size_t i = 0;
    bool done() const { return i >= t.size(); }
    const std::string* peek() const { return i < t.size() ? &t[i] : nullptr; }
    const std::string& next() {
        if (i >= t.size()) throw Error("unexpected ends!");
        return t[i++];
    }
};
double f32_tok(Peek& it) {  // parser.rs:276-280
    const std::string& tok = it.next();
    char* end = nullptr;
    const double v = std::strtod(tok.c_str(), &end);
    if (tok.empty() || end != tok.c_str() + tok.size()) throw Error("should be <f32>!");
    return (double)(float)v;
}
Json vec_tok(Peek& it, int n) {
    Json a = Json::array();
    for (int i = 0; i < n; i++) a.push(Json::number(f32_tok(it)));
    return a;
}
Json color_tok(Peek& it) {  // parser.rs:312-323: '#rrggbb' or three floats
    const std::string* p = it.peek();
    if (!p) throw Error("unexpected ends!");
    if (!p->empty() && (*p)[0] == '#') return Json::string(it.next());
    return vec_tok(it, 3);
}
Json backward() { return Json::numbers({-0.0, -0.0, -1.0, -0.0}); }  // Vec4f::backward(), lin.rs:143
Json default_tri() {  // parser.rs:417-421
    Json t = Json::array();
    t.push(Json::numbers({0.5, 0.0, -0.25}));
    t.push(Json::numbers({0.0, 0.0, 0.5}));
    t.push(Json::numbers({-0.5, 0.0, -0.25}));
    return t;
}
bool is_one_of(const std::string& s, std::initializer_list<const char*> l) {
    for (const char* x : l) if (s == x) return true;
    return false;
}
}  // namespace

// ParseFromArgs::parse_args, parser.rs:584-598: the argument list is REVERSED, split after every
// type token, and each piece reversed back — the objects come out in reverse command-line order.
std::vector<std::vector<std::string>> split_args(const std::vector<std::string>& args, const std::vector<std::string>& pat) {
    std::vector<std::vector<std::string>> out;
    std::vector<std::string> cur;
    for (auto it = args.rbegin(); it != args.rend(); ++it) {
        cur.push_back(*it);
        bool hit = false;
        for (const auto& p : pat) hit |= (p == *it);
        if (hit) { out.emplace_back(cur.rbegin(), cur.rend()); cur.clear(); }
    }
    if (!cur.empty()) out.emplace_back(cur.rbegin(), cur.rend());
    return out;
}

Json camera_from_args(const std::vector<std::string>& args) {  // parser.rs:330-349
    Json cam = Json::object();
    Peek it{args};
    while (!it.done()) {
        const std::string p = it.next();
        if (p == "pos:") cam.set("pos", vec_tok(it, 3));
        else if (p == "dir:") cam.set("dir", vec_tok(it, 4));
        else if (is_one_of(p, {"fov:", "gamma:", "exp:", "aprt:", "foc:"})) cam.set(p.substr(0, p.size() - 1), Json::number(f32_tok(it)));
        else throw Error("`" + p + "` param for `cam` is unxpected!");
    }
    return cam;
}

Json light_from_args(const std::vector<std::string>& args) {  // parser.rs:352-403
    if (args.empty()) throw Error("unexpected ends!");
    const std::string& t = args[0];
    Json l = Json::object();
    bool point;
    if (t == "pt:" || t == "point:") { point = true; l.set("type", Json::string("point")); l.set("pos", Json::numbers({0.0, 0.0, 0.0})); }
    else if (t == "dir:") { point = false; l.set("type", Json::string("dir")); l.set("dir", Json::numbers({0.0, 1.0, 0.0})); }
    else throw Error("`" + t + "` type is unxpected!");
    Peek it{args};
    while (!it.done()) {
        const std::string p = it.next();
        if (point && (p == "pt:" || p == "point:")) l.set("pos", vec_tok(it, 3));
        else if (!point && p == "dir:") {
            const Json v = vec_tok(it, 3);
            const float x = num(v.items()[0]), y = num(v.items()[1]), z = num(v.items()[2]);
            const float r = 1.0f / std::sqrt(x * x + y * y + z * z);  // Vec3f::norm, lin.rs:60-66 (parser.rs:383)
            l.set("dir", Json::numbers({(double)(x * r), (double)(y * r), (double)(z * r)}));
        }
        else if (p == "col:") l.set("color", color_tok(it));
        else if (p == "pwr:") l.set("pwr", Json::number(f32_tok(it)));
        else throw Error("`" + p + "` param for `light` is unxpected!");
    }
    return l;
}

Json renderer_from_args(const std::vector<std::string>& args) {  // parser.rs:406-582
    if (args.empty()) throw Error("unexpected ends!");
    const std::string& t = args[0];
    Json o = Json::object();
    std::string kind;
    if (t == "sph" || t == "sphere") { kind = "sphere"; o.set("type", Json::string(kind)); o.set("r", Json::number(0.5)); }
    else if (t == "pln" || t == "plane") { kind = "plane"; o.set("type", Json::string(kind)); o.set("n", Json::numbers({0.0, 0.0, 1.0})); }
    else if (t == "box") { kind = "box"; o.set("type", Json::string(kind)); o.set("sizes", Json::numbers({0.5, 0.5, 0.5})); }
    else if (t == "tri" || t == "triangle") { kind = "triangle"; o.set("type", Json::string(kind)); o.set("vtx", default_tri()); }
    else if (t == "mesh") { kind = "mesh"; o.set("type", Json::string(kind)); Json m = Json::array(); m.push(default_tri()); o.set("mesh", m); }
    else throw Error("`" + t + "` type is unxpected!");
    o.set("pos", Json::numbers({0.0, 0.0, 0.0}));
    o.set("dir", backward());
    Json mat = Json::object();
    std::vector<std::string> rest(args.begin() + 1, args.end());
    Peek it{rest};
    auto tri_tok = [&]() { Json tr = Json::array(); for (int k = 0; k < 3; k++) tr.push(vec_tok(it, 3)); return tr; };
    while (!it.done()) {
        const std::string p = it.next();
        if (kind == "sphere" && p == "r:") o.set("r", Json::number(f32_tok(it)));
        else if (kind == "plane" && p == "n:") o.set("n", vec_tok(it, 3));
        else if (kind == "box" && p == "size:") o.set("sizes", vec_tok(it, 3));  // CLI spelling; the JSON key is `sizes`
        else if (kind == "triangle" && p == "vtx:") o.set("vtx", tri_tok());
        else if (kind == "mesh" && p == "mesh:") {
            Json m = Json::array();
            m.push(tri_tok());
            for (;;) {  // parser.rs:493-503: triangles until the numbers run out (consumed tokens stay consumed)
                try { m.push(tri_tok()); } catch (const Error&) { break; }
            }
            o.set("mesh", m);
        }
        else if (p == "name:") { if (!it.done()) o.set("name", Json::string(it.next())); }
        else if (p == "pos:") o.set("pos", vec_tok(it, 3));
        else if (p == "dir:") o.set("dir", vec_tok(it, 4));
        else if (p == "albedo:") mat.set("albedo", color_tok(it));
        else if (is_one_of(p, {"rough:", "metal:", "glass:", "opacity:", "emit:"})) mat.set(p.substr(0, p.size() - 1), Json::number(f32_tok(it)));
        else if (is_one_of(p, {"tex:", "rmap:", "mmap:", "gmap:", "omap:", "emap:"})) {
            if (it.done()) throw Error("unexpected ended!");
            mat.set(p.substr(0, p.size() - 1), Json::string(it.next()));  // with a '.': a file, else inline base64 (parser.rs:521-527)
        }
        else throw Error("`" + p + "` param for `" + t + "` is unxpected!");
    }
    o.set("mat", mat);
    return o;
}

// ----------------------------------------------------------------------------- command line
static bool looks_like_flag(const std::string& s) {
    if (s.size() < 2 || s[0] != '-') return false;
    char* end = nullptr;
    std::strtod(s.c_str(), &end);
    return end != s.c_str() + s.size();  // negative numbers are values (allow_negative_numbers, cli.rs:63-69)
}

std::string usage() {
    return "Tiny raytracing microservice (B200 path).\n\n"
           "Usage: raytrace [OPTIONS] [FILE.json]\n\n"
           "  [FILE.json]            Full render description json input filename\n"
           "  -v, --verbose          Enable logging\n"
           "      --pretty           Print full render info in json with prettifier\n"
           "  -d, --dry              Dry run (useful with verbose)\n"
           "  -o, --output FILE.EXT  Final image output filename (png, ppm, jpg)\n"
           "      --http address     Launch http server\n"
           "      --bounce N         Max ray bounce\n"
           "      --sample N         Max path-tracing samples\n"
           "      --loss F           Ray bounce energy loss\n"
           "  -u, --update           Save output on each sample\n"
           "  -w, --worker N         Parallel workers count (accepted, ignored: the CUDA grid replaces the pool)\n"
           "      --dim N            Parallel jobs count on each dimension (accepted, ignored)\n"
           "  -s, --scene FILE.json  Scene description json input filename\n"
           "  -f, --frame FILE.json  Frame description json input filename\n"
           "      --res w h          Frame output image resolution\n"
           "      --ssaa F           Output image SSAAx antialiasing\n"
           "      --cam ...          pos: <f32 x3> dir: <f32 x4> fov: gamma: exp: aprt: foc:\n"
           "      --obj ...          sphere|plane|box|triangle|mesh name: <param> pos: dir: albedo: rough: metal: glass: opacity: emit: tex: rmap: mmap: gmap: omap: emap:\n"
           "      --light ...        point: <f32 x3> | dir: <f32 x3>  pwr: col:\n"
           "      --sky r g b pwr    Scene sky color\n"
           "      --device N         CUDA device (extension)\n"
           "      --gpus N           one Sampler over N GPUs (devices --device .. +N-1): sample split, films gathered over NVLink (extension);\n"
           "                         with --http: requests go round-robin over the N GPUs\n"
           "      --seed N           RNG seed (extension: the reference is unseedable)\n";
}

CliArgs parse_cli(const std::vector<std::string>& argv) {
    CliArgs a;
    size_t i = 0;
    auto one = [&](const std::string& flag) -> const std::string& {
        if (i >= argv.size()) throw Error("a value is required for '" + flag + "' but none was supplied");
        return argv[i++];
    };
    auto many = [&](std::optional<std::vector<std::string>>& dst) {
        if (!dst) dst.emplace();
        while (i < argv.size() && !looks_like_flag(argv[i])) dst->push_back(argv[i++]);
    };
    auto integer = [&](const std::string& flag) {
        const std::string& s = one(flag);
        char* end = nullptr;
        const long v = std::strtol(s.c_str(), &end, 10);
        if (s.empty() || end != s.c_str() + s.size()) throw Error("invalid value '" + s + "' for '" + flag + "': invalid digit found in string");
        return v;
    };
    auto real = [&](const std::string& flag) {
        const std::string& s = one(flag);
        char* end = nullptr;
        const double v = std::strtod(s.c_str(), &end);
        if (s.empty() || end != s.c_str() + s.size()) throw Error("invalid value '" + s + "' for '" + flag + "': invalid float literal");
        return v;
    };
    while (i < argv.size()) {
        const std::string f = argv[i++];
        if (!looks_like_flag(f)) {
            if (a.full) throw Error("unexpected argument '" + f + "' found");
            a.full = f;
        }
        else if (f == "-v" || f == "--verbose") a.verbose = true;
        else if (f == "--pretty") a.pretty = true;
        else if (f == "-d" || f == "--dry") a.dry = true;
        else if (f == "-vd" || f == "-dv") a.verbose = a.dry = true;
        else if (f == "-u" || f == "--update") a.update = true;
        else if (f == "-o" || f == "--output") a.output = one(f);
        else if (f == "--http") a.http = one(f);
        else if (f == "--bounce") a.bounce = integer(f);
        else if (f == "--sample") a.sample = integer(f);
        else if (f == "--loss") a.loss = real(f);
        else if (f == "-w" || f == "--worker") a.worker = integer(f);
        else if (f == "--dim") a.dim = integer(f);
        else if (f == "-s" || f == "--scene") a.scene = one(f);
        else if (f == "-f" || f == "--frame") a.frame = one(f);
        else if (f == "--res") { a.res.emplace(); a.res->push_back(integer(f)); a.res->push_back(integer(f)); }
        else if (f == "--ssaa") a.ssaa = real(f);
        else if (f == "--cam") { many(a.cam); if (a.cam->empty()) throw Error("a value is required for '--cam' but none was supplied"); }
        else if (f == "--obj") many(a.obj);
        else if (f == "--light") many(a.light);
        else if (f == "--sky") { many(a.sky); if (a.sky->empty()) throw Error("a value is required for '--sky' but none was supplied"); }
        else if (f == "--device") a.device = (int)integer(f);
        else if (f == "--gpus") { a.gpus = (int)integer(f); if (a.gpus < 1) throw Error("invalid value for '--gpus': at least 1"); }
        else if (f == "--seed") a.seed = (uint64_t)std::strtoull(one(f).c_str(), nullptr, 0);
        else if (f == "--dump-packed") a.dump_packed = one(f);
        else if (f == "-h" || f == "--help") throw Error(usage());
        else throw Error("unexpected argument '" + f + "' found");
    }
    return a;
}

// CLI::parse_render, cli.rs:78-153 on the JSON form: full json -> --bounce/--sample/--loss ->
// --frame -> --res/--ssaa/--cam -> --scene -> --obj/--light -> --sky
Json merged_description(const CliArgs& a) {
    Json d = a.full ? Json::parse(slurp(*a.full)) : Json::object();
    if (!d.is_obj()) throw Error("invalid type: expected a render description object");
    auto sub = [&](const char* key) {
        const Json* v = d.find(key);
        return v && v->is_obj() ? *v : Json::object();
    };
    Json rt = sub("rt"), frame = sub("frame"), scene = sub("scene");
    if (a.bounce) rt.set("bounce", Json::number((double)*a.bounce));
    if (a.sample) rt.set("sample", Json::number((double)*a.sample));
    if (a.loss) rt.set("loss", Json::number(*a.loss));
    if (a.frame) frame = Json::parse(slurp(*a.frame));  // replaces the whole frame, cli.rs:101-104
    if (a.res) frame.set("res", Json::numbers({(double)(*a.res)[0], (double)(*a.res)[1]}));
    if (a.ssaa) frame.set("ssaa", Json::number(*a.ssaa));
    if (a.cam) frame.set("cam", camera_from_args(*a.cam));  // replaces the whole camera, cli.rs:117-119
    if (a.scene) scene = Json::parse(slurp(*a.scene));   // replaces the whole scene, cli.rs:122-125
    if (a.obj) {
        Json lst = Json::array();
        if (const Json* old = scene.find("renderer")) lst = *old;
        for (const auto& piece : split_args(*a.obj, {"sphere", "sph", "plane", "pln", "box", "tri", "triangle", "mesh"})) lst.push(renderer_from_args(piece));
        scene.set("renderer", lst);
    }
    if (a.light) {
        Json lst = Json::array();
        if (const Json* old = scene.find("light")) lst = *old;
        for (const auto& piece : split_args(*a.light, {"pt:", "point:", "dir:"})) lst.push(light_from_args(piece));
        scene.set("light", lst);
    }
    if (a.sky) {  // three floats then pwr; hex is not accepted here (cli.rs:146-150)
        Peek it{*a.sky};
        Json sky = Json::object();
        sky.set("color", vec_tok(it, 3));
        sky.set("pwr", Json::number(f32_tok(it)));
        scene.set("sky", sky);
    }
    Json out = Json::object();
    out.set("rt", rt);
    out.set("frame", frame);
    out.set("scene", scene);
    return out;
}

}  // namespace mrt_host
