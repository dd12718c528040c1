// http.hpp — the HTTP endpoint of the native front-end (see http.cpp; reference: src/http.rs).
#pragma once
#include <functional>
#include <string>

#include "render.hpp"

namespace mrt_host {

using Logger = std::function<void(const std::string&)>;
std::string render_jpeg(const std::string& json_body, int device, const Logger& log);
void serve(const std::string& address, int device, int n_gpus, const Logger& log);  // requests go round-robin over devices device .. device+n_gpus-1;  // blocks forever (raytrace.rs:22-30)

}  // namespace mrt_host
