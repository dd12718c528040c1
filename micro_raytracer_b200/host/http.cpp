// http.cpp — the reference's microservice endpoint (src/http.rs:14-164) over the CUDA path:
// POST a render description as `application/json` (<= 1 MiB), get `image/jpeg` (quality 90) back.
// One thread per connection and one Sampler per request (http.rs:138,155); the same status lines
// for the same faults, in the same order: 505 (not HTTP/1.1), 405 (not POST), 400 (no
// Content-Type / length mismatch), 415 (not application/json), 411 (no Content-Length).  A body
// that is not a valid description is answered 400 (the reference logs it and drops the connection).
#include "http.hpp"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <map>
#include <thread>

#include "image_io.hpp"
#include "parser.hpp"

namespace mrt_host {

static const size_t kMaxRequest = 1024 * 1024;  // http.rs:66: one read into a 1 MiB buffer

// HttpServer::raytrace + the JPEG encode of handle(): http.rs:115-122, 136-148
std::string render_jpeg(const std::string& body, int device, const Logger& log) {
    const Render render = render_from_json(Json::parse(body), ".");
    Sampler sampler(24, 64, device);  // http.rs:138
    const auto t0 = std::chrono::steady_clock::now();
    if (render.rt.sample > 0) sampler.execute(render.scene, render.frame, render.rt, render.rt.sample);
    const Image im = sampler.img(render.frame);
    if (log) log("http:done: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()) + "s");
    return encode_jpeg(im, 90);
}

static void send_all(int fd, const std::string& s) {
    size_t off = 0;
    while (off < s.size()) {
        const ssize_t n = ::send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
        if (n <= 0) return;
        off += (size_t)n;
    }
}
static void status(int fd, const char* line) { send_all(fd, std::string("HTTP/1.1 ") + line + "\r\n"); }

static void handle(int fd, int device, const Logger& log) {
    timeval tv{10, 0};
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    std::string data;
    char buf[65536];
    size_t head_end;
    while ((head_end = data.find("\r\n\r\n")) == std::string::npos && data.size() < kMaxRequest) {
        const ssize_t n = ::recv(fd, buf, sizeof buf, 0);
        if (n <= 0) break;
        data.append(buf, (size_t)n);
    }
    const std::string head = head_end == std::string::npos ? data : data.substr(0, head_end);
    std::string body = head_end == std::string::npos ? std::string() : data.substr(head_end + 4);
    std::vector<std::string> lines;
    for (size_t p = 0; p <= head.size();) {
        const size_t q = head.find("\r\n", p);
        lines.push_back(head.substr(p, q == std::string::npos ? std::string::npos : q - p));
        if (q == std::string::npos) break;
        p = q + 2;
    }
    std::vector<std::string> parts;
    {
        size_t p = 0;
        const std::string& l0 = lines[0];
        while (p <= l0.size()) {
            const size_t q = l0.find(' ', p);
            parts.push_back(l0.substr(p, q == std::string::npos ? std::string::npos : q - p));
            if (q == std::string::npos) break;
            p = q + 1;
        }
    }
    if (parts.size() < 3) return status(fd, "400 Bad Request");
    std::map<std::string, std::string> headers;
    for (size_t k = 1; k < lines.size(); k++) {
        const size_t c = lines[k].find(": ");
        if (c != std::string::npos) headers[lines[k].substr(0, c)] = lines[k].substr(c + 2);
    }
    if (parts[2] != "HTTP/1.1") return status(fd, "505 HTTP Version Not Supported");
    if (parts[0] != "POST") return status(fd, "405 Method Not Allowed");
    if (!headers.count("Content-Type")) return status(fd, "400 Bad Request");
    if (headers["Content-Type"].rfind("application/json", 0) != 0) return status(fd, "415 Unsupported Media Type");
    if (!headers.count("Content-Length")) return status(fd, "411 Length Required");
    char* end = nullptr;
    const std::string& cl = headers["Content-Length"];
    const unsigned long n = std::strtoul(cl.c_str(), &end, 10);
    if (cl.empty() || end != cl.c_str() + cl.size() || n > kMaxRequest) return status(fd, "400 Bad Request");
    while (body.size() < n) {
        const ssize_t got = ::recv(fd, buf, sizeof buf, 0);
        if (got <= 0) break;
        body.append(buf, (size_t)got);
    }
    if (body.size() != n) return status(fd, "400 Bad Request");
    std::string jpg;
    try {
        jpg = render_jpeg(body, device, log);
    } catch (const std::exception& e) {
        if (log) log(std::string("http: ") + e.what());
        return status(fd, "400 Bad Request");
    }
    send_all(fd, "HTTP/1.1 200 OK\r\nContent-Type: image/jpeg\r\nContent-Length: " + std::to_string(jpg.size()) + "\r\n\r\n" + jpg + "\r\n");
}

// HttpServer::start, http.rs:150-162: accept forever, one thread per connection
void serve(const std::string& address, int device, int n_gpus, const Logger& log) {
    const size_t c = address.rfind(':');
    std::string host = c == std::string::npos ? "localhost" : address.substr(0, c);
    const std::string port = c == std::string::npos ? address : address.substr(c + 1);
    if (host.empty()) host = "localhost";
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    hints.ai_flags = AI_PASSIVE;
    if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || !res) throw Error("http: cannot resolve " + address);
    const int srv = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
    int yes = 1;
    setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &yes, sizeof yes);
    if (srv < 0 || ::bind(srv, res->ai_addr, res->ai_addrlen) != 0 || ::listen(srv, 64) != 0) {
        freeaddrinfo(res);
        throw Error("http: cannot listen on " + address + ": " + std::strerror(errno));
    }
    freeaddrinfo(res);
    if (log) log("http:listening: " + address);
    unsigned n_conn = 0;
    for (;;) {
        const int fd = ::accept(srv, nullptr, nullptr);
        if (fd < 0) continue;
        const int dev = device + (int)(n_conn++ % (unsigned)(n_gpus > 0 ? n_gpus : 1));  // one Sampler per request, GPUs in turn
        std::thread([fd, dev, log]() {
            try { handle(fd, dev, log); } catch (const std::exception& e) { if (log) log(std::string("http: ") + e.what()); }
            ::shutdown(fd, SHUT_RDWR);
            ::close(fd);
        }).detach();
    }
}

}  // namespace mrt_host
