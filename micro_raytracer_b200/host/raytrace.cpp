// raytrace.cpp — the native `raytrace` front-end: the reference's binary (src/bin/raytrace.rs:12-57)
// and render loop (CLI::raytrace, src/cli.rs:155-177) over the CUDA path.  Same flags, same JSON,
// same `key: v v v` mini-grammar; the pixels come from libmrt.so (sm_100a kernels), never from
// the CPU.  SURVEY.md §8(f) "next #2" (front-end) and "#3" (--http).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "http.hpp"
#include "image_io.hpp"
#include "parser.hpp"

using namespace mrt_host;

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// CLI::raytrace, cli.rs:155-177, as the reference writes it: one Sampler, `rt.sample` one-pass calls, optional
// save per pass (--update), final save.  The library queues the one-pass calls and renders them in full-length
// launches, on every GPU of --gpus (one Sampler over a device group), so this loop IS the fast path.
static double raytrace(const CliArgs& a, const Render& render, const Logger& log) {
    const std::string out = a.output.value_or("out.png");
    std::vector<int> devices;
    for (int g = 0; g < a.gpus; g++) devices.push_back(a.device + g);
    const double t_new = now();
    Sampler sampler(devices, (uint32_t)a.worker.value_or(24), (uint32_t)a.dim.value_or(64), a.seed);  // cli.rs:157
    if (log) log("cli:sampler: on " + std::to_string(sampler.n_devices()) + " gpus, created in " + std::to_string(now() - t_new) + "s");
    const double t0 = now();
    for (uint32_t n = 0; n < render.rt.sample; n++) {  // cli.rs:162
        const double dt = sampler.execute(render.scene, render.frame, render.rt);
        if (log) log("cli:sample:" + std::to_string(n) + ": " + std::to_string(dt) + "s");
        if (a.update) save_image(sampler.img(render.frame), out);  // cli.rs:166-169
    }
    const Image im = sampler.img(render.frame);  // cli.rs:173: renders whatever is still queued
    const double t1 = now();
    if (log) log("cli:device: " + std::to_string(sampler.device_seconds()) + "s in path kernels; render (first execute -> image in host memory): " + std::to_string(t1 - t0) + "s");
    save_image(im, out);
    if (log) log("cli:save: " + std::to_string(now() - t1) + "s");
    return now() - t0;
}

int main(int argc, char** argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    try {
        if (args.size() == 3 && args[0] == "--convert") {  // test hook for the encoders: RGB8 image in, format by extension out
            save_image(load_image_rgb8(args[1]), args[2]);
            return 0;
        }
        const CliArgs a = parse_cli(args);
        Logger log;
        if (a.verbose) log = [](const std::string& m) { std::cout << m << std::endl; };
        if (a.http) {  // raytrace.rs:22-30: blocks forever
            serve(*a.http, a.device, a.gpus, log ? log : Logger([](const std::string& m) { std::cout << m << std::endl; }));
            return 0;
        }
        const Json d = merged_description(a);
        const std::string base = dirname_of(a.full ? *a.full : a.scene ? *a.scene : std::string("./x"));
        const Render render = render_from_json(d, base);
        if (a.verbose) std::cout << d.dump(a.pretty ? 2 : -1) << std::endl;  // raytrace.rs:36-40
        if (a.dump_packed) {  // test hook: the flat C-ABI arrays this description packs to
            const PackedScene p(render.scene);
            const mrt_frame f = render.frame.pack();
            std::ofstream o(*a.dump_packed, std::ios::binary);
            const std::string s = p.bytes();
            o.write(s.data(), (std::streamsize)s.size());
            o.write(reinterpret_cast<const char*>(&f), sizeof f);
            o.write(reinterpret_cast<const char*>(&render.rt.bounce), 4);
            o.write(reinterpret_cast<const char*>(&render.rt.sample), 4);
            o.write(reinterpret_cast<const char*>(&render.rt.loss), 4);
        }
        if (a.dry) return 0;  // raytrace.rs:42
        const double dt = raytrace(a, render, log);
        if (log) log("cli:done: " + std::to_string(dt) + "s");
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "cli: %s\n", e.what());  // raytrace.rs:55
        return 1;
    }
}
