// The native host used as a library, the way the reference's own callers use `Sampler`
// (CLI::raytrace, src/cli.rs:155-177): load a description, run the passes, read the image out.
// Built and run by tests/test_native_host.py::test_host_library_api (needs a GPU to run).
//   usage: api_example SCENE.json PASSES OUT.ppm
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "image_io.hpp"
#include "parser.hpp"
#include "render.hpp"

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    try {
        mrt_host::Render render = mrt_host::load_render(argv[1]);
        render.frame.res = {96, 54};                      // the caller owns the description and may edit it
        const uint32_t passes = (uint32_t)std::atoi(argv[2]);
        mrt_host::Sampler sampler(24, 64);                // ≙ Sampler::new(workers, n_dim)
        double device_s = 0.0;
        for (uint32_t n = 0; n < passes; n++)             // ≙ for sample in 0..rt.sample { sampler.execute(...) }
            device_s += sampler.execute(render.scene, render.frame, render.rt);
        const mrt_host::Image im = sampler.img(render.frame);  // ≙ sampler.img(&frame)
        mrt_host::save_image(im, argv[3]);
        std::printf("passes %u device_s %.6f image %ux%u\n", sampler.passes(), device_s, im.w, im.h);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "api_example: %s\n", e.what());
        return 1;
    }
}
