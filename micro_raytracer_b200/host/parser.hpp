// parser.hpp — description parsing of the native front-end: JSON -> Render with serde's
// defaults, inline / file assets, and the command line's `key: v v v` mini-grammar
// (src/parser.rs, src/cli.rs:78-153).  See parser.cpp.
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "json.hpp"
#include "render.hpp"

namespace mrt_host {

// RenderWrapper (every key optional) -> Render, parser.rs:160-166, 929-937.
// base_dir resolves relative asset file names (textures, .obj meshes).
Render render_from_json(const Json& d, const std::string& base_dir);
Render load_render(const std::string& path);

// the mini-grammar, parser.rs:274-598; each returns the JSON form of the description
Json camera_from_args(const std::vector<std::string>& args);
Json light_from_args(const std::vector<std::string>& args);
Json renderer_from_args(const std::vector<std::string>& args);
std::vector<std::vector<std::string>> split_args(const std::vector<std::string>& args, const std::vector<std::string>& pat);

// the command line of src/cli.rs:11-74 (+ --device / --seed / --dump-packed extensions)
struct CliArgs {
    std::optional<std::string> full, output, http, scene, frame, dump_packed;
    bool verbose = false, pretty = false, dry = false, update = false;
    std::optional<long> bounce, sample, worker, dim;
    std::optional<double> loss, ssaa;
    std::optional<std::vector<long>> res;
    std::optional<std::vector<std::string>> cam, obj, light, sky;
    int device = 0;
    int gpus = 1;          // --gpus N (extension): devices device .. device+N-1 render one image together
    uint64_t seed = 0x5EED;
};
CliArgs parse_cli(const std::vector<std::string>& argv);
// CLI::parse_render, cli.rs:78-153: the merged description as JSON
Json merged_description(const CliArgs& a);
std::string dirname_of(const std::string& path);
std::string usage();

}  // namespace mrt_host
