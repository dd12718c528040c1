// mrt_scene.cu — mrt_set_scene's worker: validation of a scene description and its packing into the device layout
// (SlimInst / BoxPair / BxfInst / FatInst tables, lights, textures, the flattened depth-3 mesh octree with its
// triangle BVH and candidacy lists, the scene-level BVH) plus the text of the scene for the run-time specialised
// kernel (mrt_jit.cu).  Host code only; reference semantics cited per step.
#include "mrt_ctx.h"

namespace {


// Flattened depth-3 octree of one mesh: the non-empty leaves in the reference's depth-first
// child order (rt.rs:631-689, parser.rs:805-824), each with the triangles that have a vertex
// inside it (rt.rs:227-248).
struct LeafBuild { H3 center, size; std::vector<uint32_t> idx; };
void build_leaves(const float* tris, uint32_t n_tri, std::vector<LeafBuild>* out, float root_half[3]) {
    static const float G[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
    float mx = 0, my = 0, mz = 0;  // Mesh::gen_aabb, rt.rs:261-270
    for (uint32_t t = 0; t < n_tri; t++)
        for (int v = 0; v < 3; v++) {
            const float* p = tris + 9 * (size_t)t + 3 * v;
            mx = std::fmax(mx, std::fabs(p[0])); my = std::fmax(my, std::fabs(p[1])); mz = std::fmax(mz, std::fabs(p[2]));
        }
    const H3 A0 = {2.0f * mx, 2.0f * my, 2.0f * mz};
    root_half[0] = 0.5f * A0.x; root_half[1] = 0.5f * A0.y; root_half[2] = 0.5f * A0.z;  // Box::intersect halves the size, rt.rs:318
    const H3 A1 = {0.5f * A0.x, 0.5f * A0.y, 0.5f * A0.z};
    const H3 A2 = {0.5f * A1.x, 0.5f * A1.y, 0.5f * A1.z};
    const H3 A3 = {0.5f * A2.x, 0.5f * A2.y, 0.5f * A2.z};
    for (int i0 = 0; i0 < 8; i0++) {
        const H3 r1 = {0.0f + A0.x * (G[i0][0] * 0.25f), 0.0f + A0.y * (G[i0][1] * 0.25f), 0.0f + A0.z * (G[i0][2] * 0.25f)};
        for (int i1 = 0; i1 < 8; i1++) {
            const H3 r2 = {r1.x + A1.x * (G[i1][0] * 0.25f), r1.y + A1.y * (G[i1][1] * 0.25f), r1.z + A1.z * (G[i1][2] * 0.25f)};
            for (int i2 = 0; i2 < 8; i2++) {
                const H3 r3 = {r2.x + A2.x * (G[i2][0] * 0.25f), r2.y + A2.y * (G[i2][1] * 0.25f), r2.z + A2.z * (G[i2][2] * 0.25f)};
                const H3 hi = {r3.x + 0.5f * A3.x, r3.y + 0.5f * A3.y, r3.z + 0.5f * A3.z};
                const H3 lo = {r3.x - 0.5f * A3.x, r3.y - 0.5f * A3.y, r3.z - 0.5f * A3.z};
                LeafBuild lb{r3, A3, {}};
                for (uint32_t t = 0; t < n_tri; t++) {
                    bool in = false;
                    for (int v = 0; v < 3 && !in; v++) {
                        const float* p = tris + 9 * (size_t)t + 3 * v;
                        in = !(p[0] > hi.x || p[1] > hi.y || p[2] > hi.z) && !(p[0] < lo.x || p[1] < lo.y || p[2] < lo.z);
                    }
                    if (in) lb.idx.push_back(t);
                }
                if (!lb.idx.empty()) out->push_back(std::move(lb));
            }
        }
    }
}

// float literal that round-trips exactly (C++17 hex float)
void lit(std::string* o, float v) {
    char b[48];
    std::snprintf(b, sizeof b, "%af", (double)v);
    *o += b;
}
void lits(std::string* o, const float* v, int n) {
    for (int i = 0; i < n; i++) { *o += ", "; lit(o, v[i]); }
}
bool all_finite(const float* v, int n) {
    for (int i = 0; i < n; i++) if (!std::isfinite(v[i])) return false;
    return true;
}

// A non-negative float as fp16 bits, rounded UP (towards +inf): the half extents of a BVH node may only grow.  Values above
// the fp16 range become +inf (a box that is never culled), NaN is not expected (extents of finite boxes).
uint32_t half_up(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v > 65504.0f) return 0x7c00u;
    uint32_t b;
    std::memcpy(&b, &v, 4);
    const int e = (int)(b >> 23) - 127;
    uint32_t h;
    if (e < -24) return 1u;  // below the smallest subnormal: round up to it
    if (e < -14) {           // subnormal half: value = m * 2^-24
        const float scaled = std::ldexp(v, 24);
        h = (uint32_t)std::ceil(scaled);  // <= 1024: 1024 is the smallest normal, the bit pattern carries over
    } else {
        const uint32_t mant = b & 0x7fffffu;
        h = ((uint32_t)(e + 15) << 10) | (mant >> 13);
        if (mant & 0x1fffu) h += 1u;  // any dropped bit: next representable (the carry into the exponent is correct)
    }
    return h > 0x7c00u ? 0x7c00u : h;
}
uint32_t half2_up(float lo, float hi) { return half_up(lo) | (half_up(hi) << 16); }


// ---- BVH builder (scene-level BVH over the finite instances, triangle BVHs of the meshes; mrt_device.cuh:
// BvhNode): median split of the centroids along the widest axis, one primitive per leaf (measured best).
struct PrimBox { float lo[3], hi[3]; uint32_t ref; };
// Returns the reference of the subtree over prims[begin, end): a leaf (MRT_BVH_LEAF | prims[begin].ref) for a
// single primitive, else the index of a node that holds the boxes and references of its two halves.
// sah = false: median splits only (MRT_BVH_SAH=0, and the fallback when SAH splits came out deeper than the traversal stack).
uint32_t bvh_build(std::vector<PrimBox>& prims, size_t begin, size_t end, std::vector<BvhNode>* nodes, bool sah, int depth = 0, int* max_depth = nullptr) {
    if (max_depth) *max_depth = std::max(*max_depth, depth);
    if (end - begin == 1) return MRT_BVH_LEAF | prims[begin].ref;
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = begin; i < end; i++)
        for (int a = 0; a < 3; a++) {
            const float cc = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
            clo[a] = std::fmin(clo[a], cc); chi[a] = std::fmax(chi[a], cc);
        }
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    size_t mid = begin + (end - begin) / 2;
    bool split_done = false;
    if (sah && end - begin > 4) {
        // binned surface-area heuristic over the three axes (16 bins of the centroid range); falls back to the
        // median when every centroid lands in one bin or the best split is lopsided beyond the stack's depth budget
        constexpr int NB = 16;
        float best_cost = INFINITY; int best_ax = -1, best_bin = -1;
        for (int a = 0; a < 3; a++) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.0f)) continue;
            struct Bin { float lo[3], hi[3]; size_t n; } bins[NB];
            for (auto& b : bins) { for (int k = 0; k < 3; k++) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; } b.n = 0; }
            const float scale = (float)NB / ext;
            for (size_t i = begin; i < end; i++) {
                const float cc = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
                const int bi = std::min(NB - 1, std::max(0, (int)((cc - clo[a]) * scale)));
                Bin& b = bins[bi];
                for (int k = 0; k < 3; k++) { b.lo[k] = std::fmin(b.lo[k], prims[i].lo[k]); b.hi[k] = std::fmax(b.hi[k], prims[i].hi[k]); }
                b.n++;
            }
            auto area = [](const float* lo, const float* hi) {
                const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                return dx * dy + dy * dz + dz * dx;
            };
            float la[NB], ra[NB]; size_t ln[NB], rn[NB];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; size_t n = 0;
            for (int b = 0; b < NB; b++) {
                if (bins[b].n) for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], bins[b].lo[k]); hi[k] = std::fmax(hi[k], bins[b].hi[k]); }
                n += bins[b].n; ln[b] = n; la[b] = n ? area(lo, hi) : 0.0f;
            }
            for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } n = 0;
            for (int b = NB - 1; b >= 0; b--) {
                if (bins[b].n) for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], bins[b].lo[k]); hi[k] = std::fmax(hi[k], bins[b].hi[k]); }
                n += bins[b].n; rn[b] = n; ra[b] = n ? area(lo, hi) : 0.0f;
            }
            for (int b = 0; b + 1 < NB; b++) {  // split after bin b
                if (ln[b] == 0 || rn[b + 1] == 0) continue;
                const float cost = la[b] * (float)ln[b] + ra[b + 1] * (float)rn[b + 1];
                if (cost < best_cost) { best_cost = cost; best_ax = a; best_bin = b; }
            }
        }
        if (best_ax >= 0) {
            const int a = best_ax;
            const float scale = 16.0f / (chi[a] - clo[a]), c0 = clo[a];
            auto it = std::partition(prims.begin() + begin, prims.begin() + end, [&](const PrimBox& p) {
                const float cc = 0.5f * (p.lo[a] + p.hi[a]);
                return std::min(15, std::max(0, (int)((cc - c0) * scale))) <= best_bin;
            });
            const size_t m = (size_t)(it - prims.begin());
            const size_t small = std::min(m - begin, end - m);
            if (m > begin && m < end && small * 16 >= (end - begin) / 4 + 1) { mid = m; split_done = true; }  // keep the depth bounded
        }
    }
    if (!split_done) {
        mid = begin + (end - begin) / 2;
        std::nth_element(prims.begin() + begin, prims.begin() + mid, prims.begin() + end, [ax](const PrimBox& a, const PrimBox& b) {
            return a.lo[ax] + a.hi[ax] < b.lo[ax] + b.hi[ax];
        });
    }
    auto bounds = [&](size_t b0, size_t e0, float* lo, float* hi) {
        for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        for (size_t i = b0; i < e0; i++)
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], prims[i].lo[a]); hi[a] = std::fmax(hi[a], prims[i].hi[a]); }
    };
    const size_t node = nodes->size();
    nodes->emplace_back();
    float llo[3], lhi[3], rlo[3], rhi[3];
    bounds(begin, mid, llo, lhi);
    bounds(mid, end, rlo, rhi);
    const uint32_t l = bvh_build(prims, begin, mid, nodes, sah, depth + 1, max_depth);
    const uint32_t r = bvh_build(prims, mid, end, nodes, sah, depth + 1, max_depth);
    // centre / half-extent form, left child in the low lane; the half extent is taken from the centre AS ROUNDED
    // and padded, so the stored box still covers [lo, hi]
    float cl[3], hl[3], cr[3], hr[3];
    auto centre_half = [](const float* lo, const float* hi, float* c, float* h) {
        for (int a = 0; a < 3; a++) {
            c[a] = 0.5f * (lo[a] + hi[a]);
            const float e = std::fmax(hi[a] - c[a], c[a] - lo[a]);
            h[a] = e * (1.0f + 4e-7f) + 1e-30f;
        }
    };
    centre_half(llo, lhi, cl, hl);
    centre_half(rlo, rhi, cr, hr);
    BvhNode& n = (*nodes)[node];
    n.w0 = make_float4(cl[0], cr[0], cl[1], cr[1]);
    n.cz = make_float2(cl[2], cr[2]);
    n.hx = half2_up(hl[0], hr[0]);
    n.hy = half2_up(hl[1], hr[1]);
    n.hz = half2_up(hl[2], hr[2]);
    n.pad = 0u;
    n.refl = l;
    n.refr = r;
    return (uint32_t)node;
}
// The traversal stacks hold MRT_BVH_STACK entries.  SAH splits refuse lopsided cuts, so they stay far below that for any
// realistic input; should a build come out deeper all the same, it is redone with median splits (depth = ceil(log2 n)).
// Returns false when even that is too deep (>= 2^30 primitives: never in practice) — the caller then does without a BVH.
bool bvh_build_bounded(std::vector<PrimBox>& prims, std::vector<BvhNode>* nodes, bool sah, uint32_t* root) {
    const size_t mark = nodes->size();
    for (int attempt = 0; attempt < 2; attempt++) {
        int depth = 0;
        *root = bvh_build(prims, 0, prims.size(), nodes, sah && attempt == 0, 0, &depth);
        if (depth <= 30) return true;
        nodes->resize(mark);
        if (!sah) break;
    }
    return false;
}
// world-space AABB of an object-space box of half extents h centred on pos, under world->object matrix M
// (object->world is M^T), padded so that rounding in the primitive tests cannot leave the node
void world_box(const HM& M, H3 pos, H3 h, PrimBox* b) {
    const float hw[3] = {std::fabs(M.m[0]) * h.x + std::fabs(M.m[3]) * h.y + std::fabs(M.m[6]) * h.z,
                         std::fabs(M.m[1]) * h.x + std::fabs(M.m[4]) * h.y + std::fabs(M.m[7]) * h.z,
                         std::fabs(M.m[2]) * h.x + std::fabs(M.m[5]) * h.y + std::fabs(M.m[8]) * h.z};
    const float p[3] = {pos.x, pos.y, pos.z};
    for (int a = 0; a < 3; a++) {
        const float pad = 1e-4f * (std::fabs(hw[a]) + std::fabs(p[a])) + 1e-5f;
        b->lo[a] = p[a] - std::fabs(hw[a]) - pad;
        b->hi[a] = p[a] + std::fabs(hw[a]) + pad;
    }
}

uint32_t pack_ids(int32_t lo, int32_t hi) { return ((uint32_t)(lo < 0 ? 0xffff : lo) & 0xffffu) | (((uint32_t)(hi < 0 ? 0xffff : hi) & 0xffffu) << 16); }


// 64-bit content hash, 8 bytes at a time (multiply / xor-shift mixing; not cryptographic — it only has to tell
// "the host handed me the same scene again" from "the scene changed")
uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v * 0x9E3779B97F4A7C15ull;
    h = (h << 27 | h >> 37) * 0xD6E8FEB86659FD93ull;
    return h ^ (h >> 29);
}
uint64_t hash_bytes(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    h = mix64(h, n);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v; std::memcpy(&v, b + i, 8); h = mix64(h, v); }
    if (i < n) { uint64_t v = 0; std::memcpy(&v, b + i, n - i); h = mix64(h, v); }
    return h;
}

}  // namespace

uint64_t mrt_scene_hash(const mrt_scene* s, uint32_t normal_space) {
    uint64_t h = 0x6d72745f62323030ull ^ normal_space;
    if (s->n_objects) h = hash_bytes(h, s->objects, (size_t)s->n_objects * sizeof(mrt_object));
    if (s->n_instances) h = hash_bytes(h, s->instances, (size_t)s->n_instances * sizeof(mrt_instance));
    for (uint32_t i = 0; i < s->n_textures; i++) {  // field by field: _pad is the caller's garbage
        const mrt_texture& t = s->textures[i];
        h = mix64(mix64(mix64(h, (uint64_t)t.w << 32 | t.h), t.first_texel), t.has_dat);
    }
    h = mix64(h, s->n_textures);
    if (s->n_texels) h = hash_bytes(h, s->texels, (size_t)s->n_texels * 3 * sizeof(float));
    if (s->n_meshes) h = hash_bytes(h, s->meshes, (size_t)s->n_meshes * sizeof(mrt_mesh));
    if (s->n_triangles) h = hash_bytes(h, s->triangles, (size_t)s->n_triangles * 9 * sizeof(float));
    if (s->n_lights) h = hash_bytes(h, s->lights, (size_t)s->n_lights * sizeof(mrt_light));
    h = hash_bytes(h, s->sky_color, sizeof s->sky_color);
    h = hash_bytes(h, &s->sky_pwr, sizeof s->sky_pwr);
    return h ? h : 1;
}

int mrt_scene_upload(mrt_ctx* c, const mrt_scene* s) {
    CK(cudaSetDevice(c->device));
    if (s->n_lights > MRT_MAX_LIGHTS) return fail(c, MRT_ERR_INVALID, "more than 16 lights are not supported");
    if (s->n_objects > 0xffffu) return fail(c, MRT_ERR_INVALID, "too many objects");
    if (s->n_textures >= 0xffffu) return fail(c, MRT_ERR_INVALID, "too many textures");
    // every array pointer may be NULL when its count is 0 (an empty Vec / std::vector); a description without
    // renderers is valid — the reference renders the sky for it
    if ((s->n_objects && !s->objects) || (s->n_instances && !s->instances) || (s->n_textures && !s->textures) ||
        (s->n_texels && !s->texels) || (s->n_meshes && !s->meshes) || (s->n_triangles && !s->triangles) || (s->n_lights && !s->lights))
        return fail(c, MRT_ERR_INVALID, "null array with a non-zero count");
    for (uint32_t i = 0; i < s->n_lights; i++)
        if (s->lights[i].kind > MRT_LIGHT_DIR) return fail(c, MRT_ERR_INVALID, "unknown light kind");

    uint32_t feat = s->n_lights ? F_LIGHTS : 0u;
    // textures -> float4 texels
    std::vector<DTex> tex(s->n_textures);
    std::vector<float4> texels;
    for (uint32_t i = 0; i < s->n_textures; i++) {
        const mrt_texture& t = s->textures[i];
        const uint64_t n = (uint64_t)t.w * t.h;
        tex[i] = {t.w, t.h, (uint32_t)texels.size(), (t.has_dat && n > 0) ? 1u : 0u};
        if (tex[i].has_dat) {
            if (t.first_texel + n > s->n_texels) return fail(c, MRT_ERR_INVALID, "texture texel range out of bounds");
            if (texels.size() + n > 0x7fffffffull) return fail(c, MRT_ERR_INVALID, "textures too large");
            for (uint64_t k = 0; k < n; k++) {
                const float* p = s->texels + 3 * (t.first_texel + k);
                texels.push_back(make_float4(p[0], p[1], p[2], 0.0f));
            }
        }
    }
    // meshes -> leaves + triangles
    std::vector<DMesh> meshes(s->n_meshes);
    std::vector<DMeshLeaf> leaves;
    std::vector<uint32_t> leaf_idx;
    std::vector<DTri> tris;
    std::vector<BvhNode> tbvh;
    std::vector<DTriLeaf> tri_leaf;
    const bool mesh_bvh = c->knobs.mesh_bvh;  // test knob MRT_NO_MESH_BVH: the sequential leaf walk instead
    for (uint32_t i = 0; i < s->n_meshes; i++) {
        const mrt_mesh& m = s->meshes[i];
        if ((uint64_t)m.first_tri + m.n_tri > s->n_triangles) return fail(c, MRT_ERR_INVALID, "mesh triangle range out of bounds");
        if (m.n_tri == 0) return fail(c, MRT_ERR_INVALID, "empty mesh");
        const float* tp = s->triangles + 9 * (size_t)m.first_tri;
        std::vector<LeafBuild> lb;
        float root_half[3];
        build_leaves(tp, m.n_tri, &lb, root_half);
        if (lb.empty()) return fail(c, MRT_ERR_INVALID, "mesh octree is empty (the reference would panic, rt.rs:717)");
        meshes[i] = {(uint32_t)leaves.size(), (uint32_t)lb.size(), (uint32_t)tris.size(), m.n_tri, {root_half[0], root_half[1], root_half[2]}, 0xffffffffu};
        for (const LeafBuild& l : lb) {
            DMeshLeaf dl;
            dl.lo = make_float4(l.center.x - 0.5f * l.size.x, l.center.y - 0.5f * l.size.y, l.center.z - 0.5f * l.size.z, u2f((uint32_t)leaf_idx.size()));
            dl.hi = make_float4(l.center.x + 0.5f * l.size.x, l.center.y + 0.5f * l.size.y, l.center.z + 0.5f * l.size.z, u2f((uint32_t)l.idx.size()));
            leaves.push_back(dl);
            leaf_idx.insert(leaf_idx.end(), l.idx.begin(), l.idx.end());
        }
        // per triangle: its occurrences in the candidate sequence (leaf order, then list order), ascending
        std::vector<std::vector<DTriLeaf>> occ(m.n_tri);
        {
            size_t listed = 0;
            for (const LeafBuild& l : lb) listed += l.idx.size();
            uint32_t rank = (uint32_t)(leaf_idx.size() - listed);  // = this mesh's first position in leaf_idx
            for (size_t l = 0; l < lb.size(); l++)
                for (uint32_t ti : lb[l].idx) occ[ti].push_back(DTriLeaf{(uint32_t)(meshes[i].first_leaf + l), rank++});
        }
        for (uint32_t t = 0; t < m.n_tri; t++) {
            const float* p = tp + 9 * (size_t)t;
            DTri d;
            d.v0 = make_float4(p[0], p[1], p[2], u2f((uint32_t)tri_leaf.size()));
            d.e0 = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], u2f((uint32_t)occ[t].size()));
            d.e1 = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f);
            d.pad = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            tris.push_back(d);
            tri_leaf.insert(tri_leaf.end(), occ[t].begin(), occ[t].end());
        }
        // triangle BVH (median split, padded boxes: rounding in tri_test must not be able to leave a node).
        // A triangle no leaf lists can never be a candidate and is left out.
        meshes[i].bvh_root = 0xffffffffu;
        if (mesh_bvh && m.n_tri < (1u << 26)) {
            std::vector<PrimBox> pb;
            pb.reserve(m.n_tri);
            bool finite = true;
            for (uint32_t t = 0; t < m.n_tri; t++) {
                if (occ[t].empty()) continue;
                const float* p = tp + 9 * (size_t)t;
                PrimBox b;
                for (int a = 0; a < 3; a++) {
                    const float lo = std::fmin(p[a], std::fmin(p[3 + a], p[6 + a])), hi = std::fmax(p[a], std::fmax(p[3 + a], p[6 + a]));
                    const float pad = 1e-4f * (std::fabs(lo) + std::fabs(hi)) + 1e-5f;
                    b.lo[a] = lo - pad; b.hi[a] = hi + pad;
                    finite &= std::isfinite(b.lo[a]) && std::isfinite(b.hi[a]);
                }
                b.ref = t;
                pb.push_back(b);
            }
            if (finite && !pb.empty()) {
                uint32_t root = 0;
                if (bvh_build_bounded(pb, &tbvh, c->knobs.bvh_sah, &root)) meshes[i].bvh_root = root;
            }
        }
    }
    // instances, grouped by kind (declaration order inside a kind)
    std::vector<SlimInst> by_kind[K_NKIND];
    std::vector<FatInst> fat_k[K_NKIND];
    std::vector<uint32_t> oi_k[K_NKIND];
    std::vector<Xf> mesh_m;
    std::vector<BxfInst> bxf;
    int rot_class = 0;  // MRT_JIT_ROT: 0 no rotated instance, 1 yaw-only, 2 general
    std::vector<PrimBox> prim_boxes;  // finite instances, for the scene-level BVH
    bool prim_boxes_ok = true;
    for (uint32_t oi = 0; oi < s->n_objects; oi++) {
        const mrt_object& o = s->objects[oi];
        const mrt_material& mt = o.mat;
        if (o.kind > MRT_MESH) return fail(c, MRT_ERR_INVALID, "unknown object kind");
        if (o.kind == MRT_TRIANGLE)
            return fail(c, MRT_ERR_INVALID, "top-level triangle objects panic in the reference (Triangle::gen_aabb is todo!(), rt.rs:224); use a mesh");
        if (o.kind == MRT_MESH && o.mesh >= s->n_meshes) return fail(c, MRT_ERR_INVALID, "mesh index out of range");
        if ((uint64_t)o.first_inst + o.n_inst > s->n_instances) return fail(c, MRT_ERR_INVALID, "instance range out of bounds");
        if (o.n_inst > 0xffffu) return fail(c, MRT_ERR_INVALID, "too many instances in one object");
        if (!(mt.emit >= 0.0f && mt.emit <= 1.0f)) return fail(c, MRT_ERR_INVALID, "material emit outside [0,1] (gen_bool panics, rt.rs:968)");
        if (!(mt.opacity >= 0.0f && mt.opacity <= 1.0f)) return fail(c, MRT_ERR_INVALID, "material opacity outside [0,1] (gen_bool panics, rt.rs:1054)");
        const int32_t ids[6] = {mt.tex, mt.rmap, mt.mmap, mt.gmap, mt.omap, mt.emap};
        bool textured = false;
        for (int32_t id : ids) {
            if (id >= (int32_t)s->n_textures) return fail(c, MRT_ERR_INVALID, "texture index out of range");
            textured |= id >= 0;
        }
        if (textured && o.kind == MRT_MESH) return fail(c, MRT_ERR_INVALID, "textured mesh: to_uv is todo!() in the reference (rt.rs:806)");
        if (textured) feat |= F_TEX;
        if (mt.opacity < 1.0f || mt.omap >= 0) feat |= F_TRANSMIT;
        if (o.kind == MRT_MESH) feat |= F_MESH;
        for (uint32_t k = 0; k < o.n_inst; k++) {
            const mrt_instance& in = s->instances[o.first_inst + k];
            const float nd[4] = {-in.dir[0], -in.dir[1], -in.dir[2], -in.dir[3]};  // rt.rs:726: -inst.dir
            const HM M = transform_of(nd);
            if (!finite_m(M)) return fail(c, MRT_ERR_INVALID, "instance dir gives a non-finite transform (|w| > 1, zero or vertical facing vector)");
            const bool ident = is_identity(M);
            if (!ident) {
                const bool yaw = M.m[2] == 0.0f && M.m[5] == 0.0f && M.m[6] == 0.0f && M.m[7] == 0.0f && M.m[8] == 1.0f;
                rot_class = std::max(rot_class, yaw ? 1 : 2);
            }
            const H3 pos = {in.pos[0], in.pos[1], in.pos[2]};
            SlimInst si{};
            FatInst fi{};
            uint32_t kind;
            Xf x{};
            for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) x.m[4 * r + cc] = M.m[3 * r + cc];
            PrimBox pb{};
            bool finite_prim = true;
            if (o.kind == MRT_SPHERE) {
                kind = K_SPHERE;
                const float r = o.param[0];
                world_box(M, pos, {std::fabs(r), std::fabs(r), std::fabs(r)}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                si.b = make_float4(r * r, 0.0f, 0.0f, 0.0f);
                fi.A = make_float4(1.0f / r, r, 0.0f, 0.0f);
            } else if (o.kind == MRT_PLANE) {
                kind = K_PLANE;
                finite_prim = false;
                const H3 nraw = {o.param[0], o.param[1], o.param[2]};
                const H3 nh = hnorm(nraw);  // Plane::intersect normalises, rt.rs:404
                // t = -((o_l - pos).n^)/(d_l.n^) with o_l - pos = M(o - pos), d_l = M d  =>  n_w = M^T n^
                const H3 nw = {M.m[0] * nh.x + M.m[3] * nh.y + M.m[6] * nh.z, M.m[1] * nh.x + M.m[4] * nh.y + M.m[7] * nh.z,
                               M.m[2] * nh.x + M.m[5] * nh.y + M.m[8] * nh.z};
                si.a = make_float4(nw.x, nw.y, nw.z, 0.0f);
                si.b = make_float4(hdot(pos, nw), 0.0f, 0.0f, 0.0f);
                const H3 ns = c->normal_space == MRT_NORMAL_OBJECT ? hnorm(nraw) : hnorm(hmul(M, nraw));  // Renderer::normal, rt.rs:786,792
                fi.A = make_float4(ns.x, ns.y, ns.z, 0.0f);
            } else if (o.kind == MRT_BOX) {
                kind = ident ? K_BOX : K_BOX_XF;
                world_box(M, pos, {0.5f * std::fabs(o.param[0]), 0.5f * std::fabs(o.param[1]), 0.5f * std::fabs(o.param[2])}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                if (ident) {
                    si.a.w = 0.5f * o.param[0];
                    si.b = make_float4(0.5f * o.param[1], 0.5f * o.param[2], 0.0f, 0.0f);
                } else {
                    si.b = make_float4(0.5f * o.param[0], 0.5f * o.param[1], 0.5f * o.param[2], 0.0f);
                    const H3 mp = hmul(M, pos);
                    bxf.push_back({make_float4(M.m[0], M.m[1], M.m[2], -mp.x), make_float4(M.m[3], M.m[4], M.m[5], -mp.y),
                                   make_float4(M.m[6], M.m[7], M.m[8], -mp.z), si.b});
                }
                fi.A = make_float4((1.0f / o.param[0]) * 2.0f, (1.0f / o.param[1]) * 2.0f, (1.0f / o.param[2]) * 2.0f, 0.0f);  // rt.rs:416
            } else {
                kind = K_MESH;
                world_box(M, pos, {meshes[o.mesh].half[0], meshes[o.mesh].half[1], meshes[o.mesh].half[2]}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                si.b = make_float4(u2f(ident ? 0u : 1u), u2f(o.mesh), 0.0f, 0.0f);
                mesh_m.push_back(x);
                fi.A = make_float4(u2f(meshes[o.mesh].first_tri), 0.0f, 0.0f, 0.0f);
            }
            fi.P = make_float4(pos.x, pos.y, pos.z, u2f(kind | (ident ? FAT_IDENT : 0u) | (textured ? FAT_TEX : 0u) |
                                                        ((!ident && c->normal_space == MRT_NORMAL_FORWARD_XF) ? FAT_NXF : 0u)));
            fi.m0 = make_float4(M.m[0], M.m[1], M.m[2], u2f(pack_ids(mt.tex, mt.rmap)));
            fi.m1 = make_float4(M.m[3], M.m[4], M.m[5], u2f(pack_ids(mt.mmap, mt.gmap)));
            fi.m2 = make_float4(M.m[6], M.m[7], M.m[8], u2f(pack_ids(mt.omap, mt.emap)));
            fi.C = make_float4(mt.albedo[0], mt.albedo[1], mt.albedo[2], mt.emit);
            fi.R = make_float4(mt.rough, mt.metal, mt.glass, mt.opacity);
            if (finite_prim) {
                pb.ref = (kind << 28) | (uint32_t)by_kind[kind].size();
                for (int a = 0; a < 3; a++) prim_boxes_ok &= std::isfinite(pb.lo[a]) && std::isfinite(pb.hi[a]);
                prim_boxes.push_back(pb);
            }
            by_kind[kind].push_back(si);
            fat_k[kind].push_back(fi);
            oi_k[kind].push_back(oi | (k << 16));
        }
    }
    std::vector<FatInst> fat;
    std::vector<uint32_t> obj_inst;
    uint32_t first[K_NKIND], cnt[K_NKIND];
    for (uint32_t k = 0; k < K_NKIND; k++) {
        first[k] = (uint32_t)fat.size();
        cnt[k] = (uint32_t)by_kind[k].size();
        fat.insert(fat.end(), fat_k[k].begin(), fat_k[k].end());
        obj_inst.insert(obj_inst.end(), oi_k[k].begin(), oi_k[k].end());
    }
    if (fat.size() > 0x7fffffffu) return fail(c, MRT_ERR_INVALID, "too many instances");

    // ---- everything above only read the description; from here on the context's device state is replaced.  A
    // failure midway (out of memory) leaves the context WITHOUT a scene — never with pointers into freed buffers.
    CK(cudaStreamSynchronize(c->stream));
    c->have_scene = false;
    c->pending = 0;  // queued passes belonged to the scene that goes away (its accumulated passes are dropped too)
    // axis-aligned boxes, two per BoxPair (see mrt_device.cuh)
    std::vector<BoxPair> boxp((by_kind[K_BOX].size() + 1) / 2);
    for (size_t k = 0; k < boxp.size(); k++) {
        const SlimInst& a = by_kind[K_BOX][2 * k];
        SlimInst b{};
        if (2 * k + 1 < by_kind[K_BOX].size()) b = by_kind[K_BOX][2 * k + 1];
        else { b.a = make_float4(0.0f, 0.0f, 0.0f, -1.0f); b.b = make_float4(-1.0f, -1.0f, 0.0f, 0.0f); }  // never hit
        boxp[k].q0 = make_float4(a.a.x, b.a.x, a.a.y, b.a.y);
        boxp[k].q1 = make_float4(a.a.z, b.a.z, a.a.w, b.a.w);
        boxp[k].q2 = make_float4(a.b.x, b.b.x, a.b.y, b.b.y);
    }
    for (uint32_t k = 0; k < K_NKIND; k++) CK(c->d_slim[k].upload(by_kind[k], c->stream));
    CK(c->d_boxp.upload(boxp, c->stream));
    CK(c->d_bxf.upload(bxf, c->stream));
    // scene-level BVH: only for scenes too large to unroll (the specialised kernel covers <= 128 primitives)
    std::vector<BvhNode> bvh_nodes;
    uint32_t bvh_root = 0;
    // BVH or brute force?  Measured on random scenes of N boxes / N spheres, both through their specialised kernels
    // (unrolled / BVH, Mpaths/s): boxes 48: 11 584 / 9 157, 56: 8 621 / 8 264, 64: 6 695 / 7 307; spheres 16: 26 240 /
    // 24 744, 24: 17 986 / 18 105, 32: 14 179 / 14 699, 40: 10 865 / 12 521, 64: 5 658 / 8 041 — the cross-over sits at
    // ~60 boxes or ~26 spheres, i.e. ~60 box-equivalents with a sphere at 2.3 (a rotated box 2.5, a mesh far more).
    // Minecraft.json (84 boxes): 4 144 unrolled, 5 343 through the BVH.
    const size_t bvh_min = c->knobs.bvh_min;  // 60 unless MRT_BVH_MIN says otherwise (experiment knob)
    const size_t brute_cost = (6 * by_kind[K_BOX].size() + 14 * by_kind[K_SPHERE].size() + 15 * bxf.size() + 36 * by_kind[K_MESH].size()) / 6;
    // Scenes with a mesh always go through the scene BVH, whatever their size: its flat walk (mrt_device.cuh: bvh_walk) runs
    // the triangle BVH of a mesh instance in the same loop as the scene's nodes (MRT_MESH_VIA_BVH=0: A/B knob).
    const bool want_bvh = brute_cost > bvh_min || (c->knobs.mesh_via_bvh && !by_kind[K_MESH].empty());
    bool use_bvh = want_bvh && !prim_boxes.empty() && prim_boxes_ok && prim_boxes.size() < (1u << 24) && !c->knobs.no_bvh;  // leaf references keep 24 bits of index (bvh_leaf)
    if (use_bvh) use_bvh = bvh_build_bounded(prim_boxes, &bvh_nodes, c->knobs.bvh_sah, &bvh_root);
    CK(c->d_bvh.upload(bvh_nodes, c->stream));
    CK(c->d_mesh_m.upload(mesh_m, c->stream));
    CK(c->d_fat.upload(fat, c->stream));
    CK(c->d_tex.upload(tex, c->stream));
    CK(c->d_texels.upload(texels, c->stream));
    CK(c->d_mesh.upload(meshes, c->stream));
    CK(c->d_leaf.upload(leaves, c->stream));
    CK(c->d_leaf_idx.upload(leaf_idx, c->stream));
    CK(c->d_tri.upload(tris, c->stream));
    CK(c->d_tbvh.upload(tbvh, c->stream));
    CK(c->d_tri_leaf.upload(tri_leaf, c->stream));
    CK(c->d_obj_inst.upload(obj_inst, c->stream));

    // Do sphere hits need the reference's own arithmetic (mrt_device.cuh: refine_sphere_hit)?  Where rays of the scene's
    // extent D give the hit distance of some sphere a rounding noise above 4e-5 (0.4 of the E = 1e-4 by which the next
    // ray is offset): 1.2e-7 D^2 / r > 4e-5.  D = diagonal of the finite instances' bounds, doubled for the camera
    // standing outside them (the frame is not known here): CornellBox2 1.2e-5, CornellBox 1.2e-5, Instance 1.7e-4.
    // MRT_REFINE_SPHERES=0 / 1 forces it (A/B knob).
    bool refine_spheres = false;
    if (!by_kind[K_SPHERE].empty() && !prim_boxes.empty()) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (const PrimBox& b : prim_boxes)
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], b.lo[a]); hi[a] = std::fmax(hi[a], b.hi[a]); }
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const double D2 = 4.0 * (dx * dx + dy * dy + dz * dz);
        for (const FatInst& f : fat_k[K_SPHERE]) refine_spheres |= 1.2e-7 * D2 > 4e-5 * std::fabs((double)f.A.y);
    }
    if (c->knobs.refine_spheres >= 0) refine_spheres = c->knobs.refine_spheres != 0;

    // ---- text of the scene for the run-time specialised kernel (mrt_jit.cu); small scenes only
    if (c->jit_requested && !c->jit_header.empty()) mrt_jit_wait(c->jit_header);  // never abandon a running compile
    c->jit_header.clear();
    c->jit_kernel = nullptr;
    c->jit_requested = c->jit_failed = c->jit_from_disk = false;
    c->jit_err.clear();
    {
        const size_t n_prim = 2 * boxp.size() + by_kind[K_SPHERE].size() + by_kind[K_PLANE].size() + bxf.size() + by_kind[K_MESH].size();
        // Scenes that go through the BVH get a specialised kernel too, but one that only folds what does not
        // depend on the instance tables (kinds present, material scalars, lights, sky, rotation class): their
        // header has empty tables, so scenes of the same shape share one kernel.
        bool ok = n_prim > 0 && (n_prim <= 128 || use_bvh);
        std::string h = "// generated by mrt_set_scene\n";
        if (use_bvh) h += "#define MRT_JIT_BVH 1\n";
        const bool tables = !use_bvh;
        auto tab = [&](const char* name, size_t n, auto&& row) {
            h += std::string("#define ") + name + "(X)";
            for (size_t k = 0; k < n; k++) { h += " X(" + std::to_string(k); row(k); h += ")"; }
            h += "\n";
        };
        // box pairs: X = packed FFMA2 pair, XS = the two boxes one at a time, X1 = single box (odd count).
        // A pair constant whose two lanes differ costs two uniform-register moves per use in the packed form
        // (only equal lanes are an immediate broadcast), so lopsided pairs are cheaper unpacked.
        h += "#define MRT_JIT_BOXPAIRS(X, XS, X1, CB, CE)";
        // scenes of many boxes: consecutive pairs are bracketed, four at a time, by their bounding box
        // (declaration order is kept, so the first-minimum rule is untouched)
        const size_t cluster = c->knobs.jit_cluster;  // 4 unless MRT_JIT_CLUSTER says otherwise (experiment knob, 0 = off)
        const bool clustered = cluster > 0 && boxp.size() >= 3 * cluster;
        for (size_t k = 0; tables && k < boxp.size(); k++) {
            const float* q = &boxp[k].q0.x;  // (cA.x,cB.x, cA.y,cB.y, cA.z,cB.z, hA.x,hB.x, hA.y,hB.y, hA.z,hB.z)
            ok &= all_finite(q, 12);
            const bool odd = 2 * k + 1 >= by_kind[K_BOX].size();
            if (clustered && k % cluster == 0) {
                float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (size_t j = k; j < std::min(boxp.size(), k + cluster); j++) {
                    const float* p = &boxp[j].q0.x;
                    const int lanes = (2 * j + 1 >= by_kind[K_BOX].size()) ? 1 : 2;
                    for (int l = 0; l < lanes; l++)
                        for (int a = 0; a < 3; a++) {
                            lo[a] = std::fmin(lo[a], p[2 * a + l] - std::fabs(p[6 + 2 * a + l]));
                            hi[a] = std::fmax(hi[a], p[2 * a + l] + std::fabs(p[6 + 2 * a + l]));
                        }
                }
                for (int a = 0; a < 3; a++) {  // the cluster test and the box tests round differently: keep a margin
                    const float pad = 1e-5f * (std::fabs(lo[a]) + std::fabs(hi[a])) + 1e-6f;
                    lo[a] -= pad; hi[a] += pad;
                }
                const float v[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
                ok &= all_finite(v, 6);
                std::string t;
                lits(&t, v, 6);
                h += " CB(" + t.substr(2) + ")";
            }
            int packed = 6, scalar = 12;
            for (int a = 0; a < 3; a++) {
                const float ca = q[2 * a], cb = q[2 * a + 1], ha = q[6 + 2 * a], hb = q[7 + 2 * a];
                if (ca != 0.0f || cb != 0.0f) packed += 1 + (ca != cb ? 2 : 0);
                if (ha != hb) packed += 4;  // +h and -h pairs
                scalar += (ca != 0.0f) + (cb != 0.0f);
            }
            h += odd ? " X1(" : (scalar < packed ? " XS(" : " X(");
            h += std::to_string(k);
            lits(&h, q, 12);
            h += ")";
            if (clustered && (k % cluster == cluster - 1 || k + 1 == boxp.size())) h += " CE";
        }
        h += "\n";
        tab("MRT_JIT_SPHERES", tables ? by_kind[K_SPHERE].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_SPHERE][k];
            const float v[4] = {e.a.x, e.a.y, e.a.z, e.b.x};
            ok &= all_finite(v, 4); lits(&h, v, 4); });
        tab("MRT_JIT_PLANES", tables ? by_kind[K_PLANE].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_PLANE][k];
            const float v[4] = {e.a.x, e.a.y, e.a.z, e.b.x};
            ok &= all_finite(v, 4); lits(&h, v, 4); });
        tab("MRT_JIT_BXFS", tables ? bxf.size() : 0, [&](size_t k) { ok &= all_finite(&bxf[k].r0.x, 15); lits(&h, &bxf[k].r0.x, 12); lits(&h, &bxf[k].h.x, 3); });
        tab("MRT_JIT_MESHES", tables ? by_kind[K_MESH].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_MESH][k];
            ok &= all_finite(&e.a.x, 3) && all_finite(mesh_m[k].m, 12);
            lits(&h, &e.a.x, 3);
            uint32_t rot, mid;
            std::memcpy(&rot, &e.b.x, 4); std::memcpy(&mid, &e.b.y, 4);
            h += ", " + std::to_string(rot) + "u, " + std::to_string(mid) + "u";
            lits(&h, mesh_m[k].m, 12); });
        // big unrolled scenes: without a register budget ptxas hoists every operand (254 registers, 2 blocks
        // per SM on Minecraft.json); 3 blocks (168 registers) measured best there: 2594 -> 2757 Mpaths/s
        // BVH kernels are latency bound (long_scoreboard): 64 registers / 8 blocks per SM measured best
        // (Minecraft.json 5 429 -> 5 628, Instance.json 2 378 -> 2 401 Mpaths/s against ptxas' own 96 / 64)
        if (use_bvh && !c->knobs.jit_minblocks_env) h += "#define MRT_JIT_MINBLOCKS 8\n";
        if (tables && n_prim > 48 && !c->knobs.jit_minblocks_env) h += "#define MRT_JIT_MINBLOCKS 3\n";
        {   // rough/metal/glass/opacity shared by every material (and no map overrides them): fold them in
            bool uni = s->n_objects > 0;
            float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (uint32_t oi = 0; oi < s->n_objects && uni; oi++) {
                const mrt_material& m = s->objects[oi].mat;
                if (oi == 0) { v[0] = m.rough; v[1] = m.metal; v[2] = m.glass; v[3] = m.opacity; }
                uni = m.rough == v[0] && m.metal == v[1] && m.glass == v[2] && m.opacity == v[3] &&
                      m.rmap < 0 && m.mmap < 0 && m.gmap < 0 && m.omap < 0;
            }
            if (uni && all_finite(v, 4)) {
                h += "#define MRT_JIT_UNIFORM_R ";
                std::string t;
                lits(&t, v, 4);
                h += t.substr(2) + "\n";
            }
        }
        {
            bool binary = true;
            for (uint32_t oi = 0; oi < s->n_objects; oi++) {
                const mrt_material& m = s->objects[oi].mat;
                binary &= (m.emit == 0.0f || m.emit == 1.0f) && m.emap < 0;
            }
            if (binary) h += "#define MRT_JIT_EMIT_BINARY 1\n";
        }
        if (s->sky_color[0] == 0.0f && s->sky_color[1] == 0.0f && s->sky_color[2] == 0.0f) h += "#define MRT_JIT_SKY_BLACK 1\n";
        h += std::string("#define MRT_JIT_REFINE_SPHERES ") + (refine_spheres ? "1\n" : "0\n");
        h += "#define MRT_JIT_ROT " + std::to_string(rot_class) + "\n";
        h += "#define MRT_JIT_N_BOX " + std::to_string(cnt[K_BOX] + cnt[K_BOX_XF]) + "\n";
        h += "#define MRT_JIT_N_SPHERE " + std::to_string(cnt[K_SPHERE]) + "\n";
        if (use_bvh) {  // what the BVH leaves and the loops around the traversal may assume
            h += "#define MRT_JIT_N_ABOX " + std::to_string(cnt[K_BOX]) + "\n";
            h += "#define MRT_JIT_N_BXF " + std::to_string(cnt[K_BOX_XF]) + "\n";
            h += "#define MRT_JIT_N_MESH " + std::to_string(cnt[K_MESH]) + "\n";
            h += "#define MRT_JIT_N_LIGHTS " + std::to_string(s->n_lights) + "\n";
        }
        h += "#define MRT_JIT_N_PLANE " + std::to_string(cnt[K_PLANE]) + "\n";
        h += "#define MRT_JIT_FIRST_SPHERE " + std::to_string(first[K_SPHERE]) + "\n";
        h += "#define MRT_JIT_FIRST_PLANE " + std::to_string(first[K_PLANE]) + "\n";
        h += "#define MRT_JIT_FIRST_BXF " + std::to_string(first[K_BOX_XF]) + "\n";
        h += "#define MRT_JIT_FIRST_MESH " + std::to_string(first[K_MESH]) + "\n";
        if (ok) c->jit_header = h;  // the feature mask is appended in mrt_set_scene's tail
    }

    SceneCommon sc{};
    sc.fat = c->d_fat.p; sc.tex = c->d_tex.p; sc.texels = c->d_texels.p;
    sc.mesh = c->d_mesh.p; sc.leaf = c->d_leaf.p; sc.leaf_idx = c->d_leaf_idx.p; sc.tri = c->d_tri.p;
    sc.tbvh = c->d_tbvh.p; sc.tri_leaf = c->d_tri_leaf.p;
    sc.n_inst = (uint32_t)fat.size();
    sc.refine_spheres = refine_spheres ? 1u : 0u;
    sc.n_tex = (uint32_t)tex.size(); sc.n_texels = (uint32_t)texels.size(); sc.n_tri = (uint32_t)tris.size();
    sc.n_leaf = (uint32_t)leaves.size(); sc.n_leaf_idx = (uint32_t)leaf_idx.size(); sc.n_tri_leaf = (uint32_t)tri_leaf.size();
    sc.n_tbvh = (uint32_t)tbvh.size(); sc.n_bvh = (uint32_t)bvh_nodes.size(); sc.n_mesh = (uint32_t)meshes.size();
    sc.n_lights = s->n_lights;
    for (uint32_t k = 0; k < K_NKIND; k++) { sc.first[k] = first[k]; sc.cnt[k] = cnt[k]; }
    for (int k = 0; k < 3; k++) { sc.sky[k] = s->sky_color[k]; sc.sky_tail[k] = s->sky_color[k] * s->sky_pwr; }
    for (uint32_t i = 0; i < s->n_lights; i++) {
        const mrt_light& l = s->lights[i];  // kinds were validated before any state was touched
        H3 v = {l.v[0], l.v[1], l.v[2]};
        if (l.kind == MRT_LIGHT_DIR) { const H3 n = hnorm(v); v = {-n.x, -n.y, -n.z}; }  // rt.rs:977,1031: -dir.norm()
        sc.light[i].v_kind = make_float4(v.x, v.y, v.z, u2f(l.kind));
        sc.light[i].color_pwr = make_float4(l.color[0], l.color[1], l.color[2], l.pwr);
    }
    c->gscene.c = sc;
    c->gscene.boxp = c->d_boxp.p; c->gscene.box = c->d_slim[K_BOX].p; c->gscene.sph = c->d_slim[K_SPHERE].p; c->gscene.pln = c->d_slim[K_PLANE].p;
    c->gscene.bxf = c->d_bxf.p;
    c->gscene.bvh = use_bvh ? c->d_bvh.p : nullptr;
    c->gscene.bvh_root = bvh_root;
    c->gscene.mesh = c->d_slim[K_MESH].p; c->gscene.mesh_m = c->d_mesh_m.p;
    c->in_param = cnt[K_BOX] <= MRT_PB && cnt[K_SPHERE] <= MRT_PS && cnt[K_PLANE] <= MRT_PP && cnt[K_BOX_XF] <= MRT_PX &&
                  cnt[K_MESH] <= MRT_PM && !use_bvh && !c->knobs.force_global;
    if (c->in_param) {
        ParamScene& ps = *c->pscene;
        ps.c = sc;
        std::copy(boxp.begin(), boxp.end(), ps.boxp);
        std::copy(by_kind[K_SPHERE].begin(), by_kind[K_SPHERE].end(), ps.sph);
        std::copy(by_kind[K_PLANE].begin(), by_kind[K_PLANE].end(), ps.pln);
        std::copy(bxf.begin(), bxf.end(), ps.bxf);
        std::copy(by_kind[K_MESH].begin(), by_kind[K_MESH].end(), ps.mesh);
        std::copy(mesh_m.begin(), mesh_m.end(), ps.mesh_m);
    }
    feat |= c->knobs.force_features;
    c->features = feat;
    if (!c->jit_header.empty()) c->jit_header += "#define MRT_JIT_F " + std::to_string(feat & F_ALL) + "u\n";
    CK(cudaStreamSynchronize(c->stream));  // the uploads above are asynchronous and read this function's local arrays
    c->scene_hash = mrt_scene_hash(s, c->normal_space);
    c->have_scene = true;
    return MRT_OK;
}

