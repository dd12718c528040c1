// tools/experiments/r2_rejected_kernels.cuh — NOT COMPILED, NOT ON THE PRODUCT PATH.
//
// Three re-cuts of the work of the BVH scenes (BASELINE configs 4a Mesh.json, 4b Instance.json, 5a Minecraft.json) that
// round 2 built, verified (bit-identical hits / same paths: tests/test_gpu_api.py at commit "wave kernel draft" and the
// two commits before it) and MEASURED on a B200 — and that all lost to the megakernel they were meant to replace.
// They are kept here, as they ran, because the numbers and ncu captures in profiles/ (r2_pool_*, r2_flatwalk_*,
// r2_wave_*) refer to this code.  Mpaths/s, 1920x1080 (Minecraft 7680x4320), 128 passes per launch, one B200:
//
//                          Mesh.json   Instance.json   Minecraft.json   active lanes / instruction (Mesh, Inst., Minecr.)
//   megakernel (shipped)      2 549         2 608           6 650           5.8     9.1    10.9   (round 1 captures)
//   1. pooled kernel          2 371         2 644           5 768           6.4     9.7    12.7
//   2. flat BVH walk            912         1 856           4 738            -      9.1     -
//   3. wave kernel              992         1 306           2 362            -     12.2    11.3
//
// 1. POOLED (path_body_pool): lanes take (pixel, sample) items from their warp's pool instead of owning a pixel.
//    Hypothesis (round 1's): lanes on cheap pixels run out of samples and idle.  Falsified: the lane count barely moves;
//    the lanes are lost INSIDE each BVH walk, not to uneven sample budgets.  Overhead: -5 .. -12 %.
// 2. FLAT WALK (bvh_walk): every turn of the loop is a node visit; primitives / triangles behind a node's children are
//    tested inline, one stack entry is popped per turn, a mesh instance is entered on the same stack.
//    75 % MORE warp-instructions: inline leaf tests still run with 2.5 lanes (twice per node: left, right child), and
//    "one pop per turn" makes a lane whose popped entry is pruned sit out a whole node visit.
// 3. WAVE (path_body_wave): 64 paths in flight per warp in shared memory, lanes take searches and shading jobs from
//    two queues.  Node visits do rise to 21.4 of 32 lanes (from 14.6) — but jobs are a few turns long, and the per-job
//    bookkeeping (job start 49 instructions at 7.7 lanes, job end + next shadow ray 98 instructions at 1.3 lanes) is
//    32 % of all issued instructions; 2.6x the megakernel's instruction count in all.
// Also measured: the scene-level traversal stack in shared memory instead of local memory (MRT_SMEM_STACK=24): no
// difference (Instance 2 626 vs 2 634, Minecraft 6 648 vs 6 654) — the pop loop waits on its own divergence, not on LDL.
// What did pay in round 2: computing the light visibility BEFORE loading the surface / material (shade_hit split out of
// path_segment): fewer values live across the shadow-ray search, Minecraft.json spills 157 -> 100 LDL/STL, +6 %.

#if 0
// ------------------------------------------------------------------------------------------------ 2. flat walk (mrt_device.cuh)
// ---- The same search as ONE flat loop (what the kernels run; bvh_traverse + mesh_test_bvh above stay as the form
// the bit-identity tests compare with, MRT_WALK_V1).  ncu on round 1's loop (profiles/r2_*): the node-visit block ran
// with 12 - 14 of 32 lanes, but the leaf block with 4, the stack push with 3.7 and the pop loop with 2.5 — lanes at a
// leaf, lanes at a node and lanes popping took turns, and a mesh instance behind a scene leaf ran its whole triangle
// walk as one leaf step while the rest of the warp waited.  Here every turn of the loop is a NODE VISIT:
//   * the primitives (scene level) or triangles (inside a mesh instance) behind the children of the node are tested
//     inline, under a predicate, in the very turn that found their box — no push, no extra turn, no pop for them;
//   * one stack entry is popped per turn, in the same turn for every lane that needs one (no inner pop loop);
//   * a mesh instance is ENTERED: a marker goes on the stack, the lane switches to the object-space ray and keeps
//     visiting nodes — now of the triangle BVH — in the same loop; popping the marker folds the mesh's entry / exit
//     candidates into the scene-level best and switches back.
// Candidates, tie rules (lexicographic (t0, instance index); first-min / last-max with candidate ranks inside a mesh)
// and every primitive test are those of the nested form: hit ids, t0, t1 and images are bit-identical.
#define MRT_WALK_EMPTY 0x7ffffffeu
#define MRT_WALK_MARKER 0x7fffffffu
#define MRT_WALK_STACK 64   // scene depth (<= 30) + marker + mesh depth (<= 30)
template <uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ void bvh_walk(Best& B, const GlobalScene& s, const RayPre& r, const RayPk& rp) {
    const SceneCommon& c = s.c;
    constexpr bool MESH = (F & F_MESH) != 0 && MRT_BVH_HAS_MESH;
    constexpr bool PRUNE_TRI = !ANY && !WANT_T1;  // inside a mesh the exit candidate may lie anywhere: no pruning when it is wanted
    const float INF = __int_as_float(0x7f800000);
    uint32_t stack[MRT_WALK_STACK];
    float stack_t[MRT_WALK_STACK];
    int sp = 0;
    NodeRay nr = r.n;
    const BvhNode* nodes = s.bvh;
    bool in_mesh = false;
    // state of the mesh instance being walked
    f3 mo = mk(0.f, 0.f, 0.f), md = mk(0.f, 0.f, 0.f), mm = mk(0.f, 0.f, 0.f), mom = mk(0.f, 0.f, 0.f);  // object-space ray, 1/d (fixed), o * 1/d
    uint32_t m_first_tri = 0u, m_inst = 0u;
    float b0 = INF, b1 = -INF;
    uint32_t r0 = 0xffffffffu, r1 = 0u;
    int k0 = -1, k1 = -1;

    auto tri_leaf = [&](uint32_t ti) {
        const DTri* tp = &c.tri[m_first_tri + ti];
        DTri tr;
        tr.v0 = __ldg(&tp->v0); tr.e0 = __ldg(&tp->e0); tr.e1 = __ldg(&tp->e1);
        float t;
        if (tri_test(tr, mo, md, &t) && !(PRUNE_TRI && !(t <= b0))) {
            uint32_t rf = 0xffffffffu, rl = 0u;
            if (tri_candidate<WANT_T1>(c, tr, mm, mom, &rf, &rl)) {
                if constexpr (ANY) { B.any = true; return; }
                if (t < b0 || (t == b0 && rf < r0)) { b0 = t; r0 = rf; k0 = (int)ti; }
                if constexpr (WANT_T1) { if (t > b1 || (t == b1 && rl >= r1)) { b1 = t; r1 = rl; k1 = (int)ti; } }
            }
        }
    };
    // a mesh instance behind a scene-level leaf: ray into object space, root AABB (rt.rs:708-710), then its triangle BVH.
    // Returns the reference to continue with (the mesh's root, or EMPTY when the instance is missed / was walked inline).
    auto enter_mesh = [&](uint32_t k) -> uint32_t {
        if constexpr (MESH) {
            const SlimInst e = ldg_slim(s.mesh + k);
            f3 ol = r.o - xyz(e.a), dl = r.d;
            if (__float_as_uint(e.b.x) != 0u) { ol = mulXf(s.mesh_m[k], ol); dl = mulXf(s.mesh_m[k], r.d); }
            const DMesh mh = c.mesh[__float_as_uint(e.b.y)];
            const f3 m = rcp_fixed3(dl);
            const f3 om = ol * m;
            if (!mesh_root_hit(mh, m, om)) return MRT_WALK_EMPTY;
            if (mh.bvh_root == 0xffffffffu) {  // no triangle BVH (MRT_NO_MESH_BVH): the sequential leaf walk, as one step
                float t0 = 0.f, t1 = 0.f;
                int tr0 = -1, tr1 = -1;
                const bool hit = mesh_leaf_walk<ANY, WANT_T1>(c, mh, ol, dl, m, om, &t0, &t1, &tr0, &tr1);
                best_update_lex<F, ANY, WANT_T1>(B, hit, t0, t1, (int)(c.first[K_MESH] + k), tr0, tr1);
                return MRT_WALK_EMPTY;
            }
            stack[sp] = MRT_WALK_MARKER; stack_t[sp] = 0.0f; sp++;
            mo = ol; md = dl; mm = m; mom = om;
            m_first_tri = mh.first_tri; m_inst = k;
            b0 = INF; b1 = -INF; r0 = 0xffffffffu; r1 = 0u; k0 = -1; k1 = -1;
            nr.bm = true_rcp3(dl, m);
            nr.bnom = mk(-ol.x * nr.bm.x, -ol.y * nr.bm.y, -ol.z * nr.bm.z);
            nr.bam = mk(fabsf(nr.bm.x), fabsf(nr.bm.y), fabsf(nr.bm.z));
            nodes = c.tbvh;
            in_mesh = true;
            return mh.bvh_root;
        } else {
            (void)k;
            return MRT_WALK_EMPTY;
        }
    };
    auto is_mesh_ref = [&](uint32_t ref) -> bool { return MESH && !in_mesh && ((ref >> 28) & 7u) == K_MESH; };
    auto leaf = [&](uint32_t ref) {  // a leaf that is NOT a mesh instance
        if (MESH && in_mesh) tri_leaf(ref & ~MRT_BVH_LEAF);
        else bvh_leaf<F, ANY, WANT_T1>(B, s, r, rp, ref & ~MRT_BVH_LEAF);
    };

    auto far_bound = [&]() -> float {
        // behind this nothing can win: '<=' because an equal t0 with a lower index (rank) must still be found; inside a
        // mesh the mesh's own entry candidate and the scene-level best both bound the search — unless the exit
        // candidate is wanted, which may lie anywhere
        return ANY ? INF : ((MESH && in_mesh) ? (PRUNE_TRI ? fminf(b0, B.t0) : INF) : B.t0);
    };
    uint32_t cur = s.bvh_root;
    for (;;) {
        if (cur != MRT_WALK_EMPTY) {
            uint32_t cl = cur, cr = MRT_WALK_EMPTY, next = MRT_WALK_EMPTY;
            float tl = 0.0f, tr = 0.0f;
            bool hl = true, hr = false, entered = false;
            if (cur & MRT_BVH_LEAF) {
                // a mesh instance to enter; or a root that is a single primitive / triangle: (cl, hl) as set above
                if (is_mesh_ref(cur)) { next = enter_mesh(cur & 0x0fffffffu); entered = true; hl = false; }
            } else {
                const float4 q0 = __ldg(&nodes[cur].q0), q1 = __ldg(&nodes[cur].q1), q2 = __ldg(&nodes[cur].q2);
                const uint2 ref = __ldg(reinterpret_cast<const uint2*>(&nodes[cur].ref));
                float tfl, tfr;
                node_slabs(nr, q0, q1, q2, &tl, &tfl, &tr, &tfr);
                const float bound = far_bound();
                hl = tl <= tfl && tfl >= 0.0f && tl <= bound;
                hr = tr <= tfr && tfr >= 0.0f && tr <= bound;
                cl = ref.x; cr = ref.y;
            }
            // primitives / triangles behind the children: tested now, in this turn (one copy of the test code: the
            // lanes whose left child is a leaf go first, then those whose right child is)
#pragma unroll 1
            for (int ch = 0; ch < 2; ch++) {
                const uint32_t ref = ch ? cr : cl;
                const bool h = ch ? hr : hl;
                if (h && (ref & MRT_BVH_LEAF) && !is_mesh_ref(ref)) {
                    leaf(ref);
                    if (ch) hr = false; else hl = false;
                }
            }
            if constexpr (ANY) { if (B.any) return; }
            if (!entered) {
                if constexpr (!ANY) {
                    const float bound = far_bound();  // may just have come closer
                    hl = hl && tl <= bound; hr = hr && tr <= bound;
                }
                if (hl && hr) {
                    const bool left_first = tl <= tr;
                    if (sp < MRT_WALK_STACK) { stack[sp] = left_first ? cr : cl; stack_t[sp] = left_first ? tr : tl; sp++; }
                    next = left_first ? cl : cr;
                } else next = hl ? cl : (hr ? cr : MRT_WALK_EMPTY);
            }
            cur = next;
        }
        if (cur == MRT_WALK_EMPTY) {  // one pop per turn
            if (sp == 0) return;
            --sp;
            const uint32_t ref = stack[sp];
            if (MESH && ref == MRT_WALK_MARKER) {  // the mesh instance is done: fold its candidates into the scene-level best
                if constexpr (!ANY) {
                    best_update_lex<F, ANY, WANT_T1>(B, k0 >= 0, b0, WANT_T1 ? b1 : b0, (int)(c.first[K_MESH] + m_inst), k0, WANT_T1 ? k1 : k0);
                }
                in_mesh = false;
                nr = r.n;
                nodes = s.bvh;
                continue;
            }
            // a subtree that starts behind the best found since it was pushed holds nothing closer
            if (stack_t[sp] <= far_bound()) cur = ref;
        }
    }
}


// ------------------------------------------------------------------------------------------------ 1. pooled kernel (mrt_path.cuh)
// The same kernel with the lanes UNBOUND from the pixels: a warp owns the (pixel, sample) items of its tile — 32
// pixels x n_samples — and a lane that finishes a path takes the next item from the warp's counter, whichever pixel
// it belongs to.  In scenes searched through a BVH the cost of a path varies by orders of magnitude between the
// pixels of a tile (sky next to a mesh silhouette): bound to its pixel a lane on a cheap pixel runs out of samples
// and idles for the rest of the launch (ncu, round 1: 5.8 - 10.9 of 32 lanes active); unbound, every lane works until
// the tile's pool is dry.  The RNG is keyed by (pixel, global sample), so the paths are the same paths; only the
// order in which a pixel's samples are summed changes.  Per warp in shared memory: the pixels' focus vectors and lens
// seeds, their radiance sums (float atomics: at most a handful of lanes finish in the same iteration), the counter.
template <class V, uint32_t F>
__device__ __forceinline__ void path_body_pool(const V sc, const FilmParams& fp) {
    __shared__ float4 s_q[MRT_POOL_BLOCK];        // focus vector of the pixel (xyz) + its film index (w, bits)
    __shared__ float s_acc[3][MRT_POOL_BLOCK];
    __shared__ uint32_t s_next[MRT_POOL_BLOCK / 32];
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    const bool inside = px < fp.nw && py < fp.nh;
    const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u, w = threadIdx.x >> 5;
    const uint32_t valid = __ballot_sync(0xffffffffu, inside);
    if (valid == 0u || fp.n_samples == 0u) return;  // warp-uniform
    const uint32_t pix = py * fp.nw + px;
    {
        const f3 q = inside ? pixel_focus_vec(fp, px, py) : mk(0.f, 1.f, 0.f);
        s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, __uint_as_float(pix));
        s_acc[0][threadIdx.x] = 0.0f; s_acc[1][threadIdx.x] = 0.0f; s_acc[2][threadIdx.x] = 0.0f;
        if (lane == 0u) s_next[w] = 0u;
    }
    __syncwarp();
    const uint32_t nvalid = (uint32_t)__popc(valid);
    const uint32_t total = nvalid * fp.n_samples;
    uint32_t tl = 0, sample = 0, tpix = 0;  // the item this lane is working on: tile lane, global sample, film index
    f3 acc = mk(0.f, 0.f, 0.f);
    PathState p;
    p.o = mk(0.f, 0.f, 0.f); p.d = mk(0.f, 1.f, 0.f); p.T = mk(1.f, 1.f, 1.f);
    p.pwr = 1.0f;
    p.bounce = MRT_NEED_PATH;
    bool have = false;
    for (;;) {
        if (p.bounce == MRT_NEED_PATH) {
            if (have) {  // hand the finished path's radiance to its pixel
                atomicAdd(&s_acc[0][wbase + tl], acc.x);
                atomicAdd(&s_acc[1][wbase + tl], acc.y);
                atomicAdd(&s_acc[2][wbase + tl], acc.z);
                have = false;
            }
            const uint32_t item = atomicAdd(&s_next[w], 1u);
            if (item >= total) break;
            // items run sample-major: all pixels of the tile for one sample, then the next sample
            const uint32_t jj = item / nvalid, k = item - jj * nvalid;
            tl = valid == 0xffffffffu ? k : (uint32_t)__fns(valid, 0u, (int)k + 1);
            sample = fp.sample0 + jj * fp.sample_stride;
            const float4 qs = s_q[wbase + tl];
            tpix = __float_as_uint(qs.w);
            const float2 u = rng_cam(cam_hash_seed(tpix, fp.key), sample);
            camera_ray(fp, xyz(qs), u.x, u.y, &p.o, &p.d);
            p.T = mk(1.f, 1.f, 1.f);
            p.pwr = 1.0f;
            p.bounce = 0;
            acc = mk(0.f, 0.f, 0.f);
            have = true;
        }
        if (path_segment<V, F>(sc, fp, tpix, sample, p, acc)) p.bounce = MRT_NEED_PATH;
    }
    __syncwarp();
    if (inside) {
        float4 a = fp.accum[pix];
        a.x += s_acc[0][threadIdx.x]; a.y += s_acc[1][threadIdx.x]; a.z += s_acc[2][threadIdx.x];
        fp.accum[pix] = a;
    }
}


// ------------------------------------------------------------------------------------------------ 3. wave kernel (mrt_path.cuh)
// =====================================================================================================================
// The WAVE kernel: the megakernel's work re-cut for scenes that are searched through a BVH.
//
// ncu on the megakernel (profiles/r2_*): on Mesh.json 5.8 - 6.4 of 32 lanes are active per issued instruction, on
// Instance.json 9.1 - 9.7, on Minecraft.json 10.9 - 12.7 — and unbinding the lanes from their pixels (path_body_pool)
// does not change that.  The lanes are lost INSIDE the searches: every lane starts its ray's walk together, the walks
// last anything from one step (the ray misses everything) to a hundred, and a lane that is done waits for the
// longest walk of its warp, segment after segment.  So here a lane is bound neither to a pixel nor to a path:
//   * a warp keeps MRT_WAVE_SLOTS paths in flight, their state (ray, throughput, radiance, hit record) in shared
//     memory, and two queues of slot numbers: paths whose ray still has to be searched, paths whose search is done;
//   * SEARCH: a lane without a job takes the next slot from the first queue and walks its ray through the BVH one
//     step per turn of the warp's loop — scene nodes, primitives, and, inside a mesh instance, the triangle BVH, all in
//     the same loop; when the walk ends it goes on with the segment's shadow rays (one walk per light), then hands the
//     slot to the second queue and takes the next job.  Lanes finish at different turns and never wait for each other;
//   * SHADE: whenever 32 searched paths wait (or the search queue runs dry), every lane shades one of them
//     (shade_hit: the megakernel's own code), which ends the path — its radiance goes to its pixel, the slot gets the
//     warp's next (pixel, sample) item — or gives the slot its next ray; either way the slot returns to the first queue.
// Same rays, same tests, same tie rules, same RNG keys as the megakernel: a pixel's samples are merely summed in
// another order.  Everything is per warp: no block-level synchronisation, no global-memory queues.
// =====================================================================================================================
#ifndef MRT_WAVE_SLOTS
#define MRT_WAVE_SLOTS 64
#endif
#ifndef MRT_WAVE_STARVE
#define MRT_WAVE_STARVE 8   // lanes without work (and nothing queued for them) at which the warp turns to shading
#endif
#define MRT_WAVE_IDLE 0xffffffffu

// One lane's walk of one ray (registers).  The stack lives in local memory next to it.
struct Walk {
    uint32_t slot;    // MRT_WAVE_IDLE: no job
    uint32_t phase;   // 0: closest hit of the path's ray; 1 + li: shadow ray towards light li (any hit)
    f3 o, d;          // the ray in the space being searched: world, or the object space of the mesh instance entered
    f3 m, nom;        // 1/d with Box::intersect's fix-up (rt.rs:303-316) and -o * m: primitive tests, octree-leaf candidacy
    f3 bm, bnom;      // the geometrically true reciprocal (true_rcp3) and -o * bm: BVH node slabs
    Best B;           // scene-level best so far
    uint32_t cur;     // node / leaf to visit next, MRT_WALK_EMPTY: pop
    int sp;
    // the mesh instance being walked
    bool in_mesh;
    uint32_t m_first_tri, m_inst;
    float b0, b1;
    uint32_t r0, r1;
    int k0, k1;
};

__device__ __forceinline__ void walk_set_ray(Walk& w, f3 o, f3 d) {
    w.o = o; w.d = d;
    w.m = rcp_fixed3(d);
    w.nom = mk(-o.x * w.m.x, -o.y * w.m.y, -o.z * w.m.z);
    w.bm = true_rcp3(d, w.m);
    w.bnom = mk(-o.x * w.bm.x, -o.y * w.bm.y, -o.z * w.bm.z);
}

// Start the walk of (o, d) through the scene: planes first (infinite, not in the BVH), then the root.
// Returns true when the walk is already over (any-hit query answered by a plane).
template <uint32_t F>
__device__ __forceinline__ bool walk_begin(Walk& w, const GlobalScene& s, f3 o, f3 d) {
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    const SceneCommon& c = s.c;
    walk_set_ray(w, o, d);
    w.B.t0 = __int_as_float(0x7f800000); w.B.t1 = 0.0f; w.B.bi = -1; w.B.tr0 = w.B.tr1 = -1; w.B.any = false;
    w.in_mesh = false;
    w.sp = 0;
    w.cur = s.bvh_root;
    for (uint32_t k = 0; k < MRT_N_PLANES(c); k++) {  // same arithmetic and tie rule as closest_hit's plane loop
        const SlimInst e = ldg_slim(s.pln + k);
        const float t0 = (e.b.x - dot(o, xyz(e.a))) * frcp(dot(d, xyz(e.a)));
        best_update_lex<F, false, WANT_T1>(w.B, t0 > 0.0f, t0, t0, (int)(c.first[K_PLANE] + k), -1, -1);
    }
    return w.phase != 0u && w.B.bi >= 0;
}

// One step of a walk.  Returns true when the walk is over (w.B holds the answer; for a shadow ray: B.bi >= 0 = occluded).
// `wo`, `wd`: the walk's world-space ray (needed again when a mesh instance is left).
template <uint32_t F>
__device__ __forceinline__ bool walk_turn(Walk& w, uint32_t* stack, float* stack_t, const GlobalScene& s, f3 wo, f3 wd) {
    constexpr bool MESH = (F & F_MESH) != 0 && MRT_BVH_HAS_MESH;
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    const SceneCommon& c = s.c;
    const float INF = __int_as_float(0x7f800000);
    const bool any = w.phase != 0u;
    // behind this nothing can win ('<=': an equal t0 with a lower index / rank must still be found); inside a mesh whose
    // exit candidate is wanted nothing may be pruned
    auto far_bound = [&]() -> float {
        if (any) return INF;
        if (MESH && w.in_mesh) return WANT_T1 ? INF : fminf(w.b0, w.B.t0);
        return w.B.t0;
    };
    if (w.cur & MRT_BVH_LEAF) {
        const uint32_t ref = w.cur & ~MRT_BVH_LEAF;
        w.cur = MRT_WALK_EMPTY;
        if (MESH && w.in_mesh) {  // a triangle: hit, then candidacy (is one of the octree leaves that list it pierced?)
            const DTri* tp = &c.tri[w.m_first_tri + ref];
            DTri tr;
            tr.v0 = __ldg(&tp->v0); tr.e0 = __ldg(&tp->e0); tr.e1 = __ldg(&tp->e1);
            float t;
            if (tri_test(tr, w.o, w.d, &t) && (any || WANT_T1 || t <= w.b0)) {
                uint32_t rf = 0xffffffffu, rl = 0u;
                const f3 om = mk(-w.nom.x, -w.nom.y, -w.nom.z);
                if (tri_candidate<WANT_T1>(c, tr, w.m, om, &rf, &rl)) {
                    if (any) { w.B.bi = 0; return true; }
                    if (t < w.b0 || (t == w.b0 && rf < w.r0)) { w.b0 = t; w.r0 = rf; w.k0 = (int)ref; }
                    if constexpr (WANT_T1) { if (t > w.b1 || (t == w.b1 && rl >= w.r1)) { w.b1 = t; w.r1 = rl; w.k1 = (int)ref; } }
                }
            }
        } else if (MESH && (ref >> 28) == K_MESH) {  // a mesh instance: ray into object space, root AABB, then its triangle BVH
            const uint32_t k = ref & 0x0fffffffu;
            const SlimInst e = ldg_slim(s.mesh + k);
            f3 ol = w.o - xyz(e.a), dl = w.d;
            if (__float_as_uint(e.b.x) != 0u) { ol = mulXf(s.mesh_m[k], ol); dl = mulXf(s.mesh_m[k], w.d); }
            const DMesh mh = c.mesh[__float_as_uint(e.b.y)];
            const f3 m = rcp_fixed3(dl);
            const f3 om = ol * m;
            if (mesh_root_hit(mh, m, om)) {
                if (mh.bvh_root == 0xffffffffu) {  // no triangle BVH (MRT_NO_MESH_BVH): the sequential leaf walk, as one step
                    float t0 = 0.f, t1 = 0.f;
                    int tr0 = -1, tr1 = -1;
                    const bool hit = mesh_leaf_walk<false, WANT_T1>(c, mh, ol, dl, m, om, &t0, &t1, &tr0, &tr1);
                    best_update_lex<F, false, WANT_T1>(w.B, hit, t0, t1, (int)(c.first[K_MESH] + k), tr0, tr1);
                } else {
                    stack[w.sp] = MRT_WALK_MARKER; stack_t[w.sp] = 0.0f; w.sp++;
                    walk_set_ray(w, ol, dl);
                    w.in_mesh = true;
                    w.m_first_tri = mh.first_tri; w.m_inst = k;
                    w.b0 = INF; w.b1 = -INF; w.r0 = 0xffffffffu; w.r1 = 0u; w.k0 = -1; w.k1 = -1;
                    w.cur = mh.bvh_root;
                }
            }
        } else {  // any other primitive
            RayPre r;
            r.o = w.o; r.d = w.d; r.m = w.m; r.nom = w.nom;
            r.am = mk(fabsf(w.m.x), fabsf(w.m.y), fabsf(w.m.z));
            r.nam = -r.am;
            const RayPk rp = {pk2(w.o.x, w.d.x), pk2(w.o.y, w.d.y), pk2(w.o.z, w.d.z)};
            bvh_leaf<F, false, WANT_T1>(w.B, s, r, rp, ref);
        }
        if (any && w.B.bi >= 0) return true;
    } else if (w.cur != MRT_WALK_EMPTY) {
        const BvhNode* nd = ((MESH && w.in_mesh) ? c.tbvh : s.bvh) + w.cur;
        const float4 q0 = __ldg(&nd->q0), q1 = __ldg(&nd->q1), q2 = __ldg(&nd->q2);
        const uint2 ref = __ldg(reinterpret_cast<const uint2*>(&nd->ref));
        NodeRay nr;
        nr.bm = w.bm; nr.bnom = w.bnom;
        nr.bam = mk(fabsf(w.bm.x), fabsf(w.bm.y), fabsf(w.bm.z));
        float tl, tfl, tr, tfr;
        node_slabs(nr, q0, q1, q2, &tl, &tfl, &tr, &tfr);
        const float bound = far_bound();
        const bool hl = tl <= tfl && tfl >= 0.0f && tl <= bound;
        const bool hr = tr <= tfr && tfr >= 0.0f && tr <= bound;
        if (hl && hr) {
            const bool left_first = tl <= tr;
            if (w.sp < MRT_WALK_STACK) { stack[w.sp] = left_first ? ref.y : ref.x; stack_t[w.sp] = left_first ? tr : tl; w.sp++; }
            w.cur = left_first ? ref.x : ref.y;
        } else w.cur = hl ? ref.x : (hr ? ref.y : MRT_WALK_EMPTY);
    }
    while (w.cur == MRT_WALK_EMPTY) {  // next subtree; one that starts behind the best found since it was pushed holds nothing closer
        if (w.sp == 0) return true;
        --w.sp;
        const uint32_t ref = stack[w.sp];
        if (MESH && ref == MRT_WALK_MARKER) {  // the mesh instance is done: fold its candidates into the scene-level best
            best_update_lex<F, false, WANT_T1>(w.B, w.k0 >= 0, w.b0, WANT_T1 ? w.b1 : w.b0, (int)(c.first[K_MESH] + w.m_inst), w.k0, WANT_T1 ? w.k1 : w.k0);
            w.in_mesh = false;
            walk_set_ray(w, wo, wd);
        } else if (stack_t[w.sp] <= far_bound()) w.cur = ref;
    }
    return false;
}

template <class V, uint32_t F>
__device__ __forceinline__ void path_body_wave(const V sc, const FilmParams& fp) {
    constexpr uint32_t S = MRT_WAVE_SLOTS, NW = MRT_POOL_BLOCK / 32u;
    static_assert(S >= 32u && S <= 256u && (S & (S - 1u)) == 0u, "slots per warp: a power of two, 32 .. 256");
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    constexpr bool MESH = (F & F_MESH) != 0 && MRT_BVH_HAS_MESH;
    // per pixel of the block's tile
    __shared__ float4 s_q[MRT_POOL_BLOCK];        // focus vector (xyz) + film index (w, bits)
    __shared__ float s_acc[3][MRT_POOL_BLOCK];
    // per path slot (structure of arrays: lane i touches word i of a row)
    __shared__ float s_ray[6][NW * S];            // o, d
    __shared__ float s_T[3][NW * S];              // throughput
    __shared__ float s_L[3][NW * S];              // radiance of the path so far
    __shared__ float s_pwr[NW * S];
    __shared__ uint32_t s_bounce[NW * S], s_tl[NW * S], s_sample[NW * S];
    __shared__ float s_t0[NW * S], s_t1[NW * S];  // hit record of the path's current ray
    __shared__ int s_inst[NW * S], s_tri0[NW * S], s_tri1[NW * S];
    __shared__ uint32_t s_vis[NW * S];
    __shared__ uint8_t s_qtrav[NW][S], s_qshade[NW][S];

    const GlobalScene& gs = sc.s;
    const SceneCommon& c = sc.c();
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    const bool inside = px < fp.nw && py < fp.nh;
    const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u, wi = threadIdx.x >> 5, sbase = wi * S;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t valid = __ballot_sync(0xffffffffu, inside);
    if (valid == 0u || fp.n_samples == 0u) return;  // warp-uniform
    const uint32_t pix = py * fp.nw + px;
    {
        const f3 q = inside ? pixel_focus_vec(fp, px, py) : mk(0.f, 1.f, 0.f);
        s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, __uint_as_float(pix));
        s_acc[0][threadIdx.x] = 0.0f; s_acc[1][threadIdx.x] = 0.0f; s_acc[2][threadIdx.x] = 0.0f;
    }
    __syncwarp();
    const uint32_t nvalid = (uint32_t)__popc(valid);
    const uint32_t total = nvalid * fp.n_samples;  // the warp's (pixel, sample) items, sample-major

    // a new camera path for `slot` from item `item`
    auto start_path = [&](uint32_t slot, uint32_t item) {
        const uint32_t jj = item / nvalid, k = item - jj * nvalid;
        const uint32_t tl = valid == 0xffffffffu ? k : (uint32_t)__fns(valid, 0u, (int)k + 1);
        const uint32_t sample = fp.sample0 + jj * fp.sample_stride;
        const float4 qs = s_q[wbase + tl];
        const float2 u = rng_cam(cam_hash_seed(__float_as_uint(qs.w), fp.key), sample);
        f3 o, d;
        camera_ray(fp, xyz(qs), u.x, u.y, &o, &d);
        const uint32_t i = sbase + slot;
        s_ray[0][i] = o.x; s_ray[1][i] = o.y; s_ray[2][i] = o.z; s_ray[3][i] = d.x; s_ray[4][i] = d.y; s_ray[5][i] = d.z;
        s_T[0][i] = 1.0f; s_T[1][i] = 1.0f; s_T[2][i] = 1.0f;
        s_L[0][i] = 0.0f; s_L[1][i] = 0.0f; s_L[2][i] = 0.0f;
        s_pwr[i] = 1.0f; s_bounce[i] = 0u; s_tl[i] = tl; s_sample[i] = sample;
    };
    // world-space ray of job (slot, phase): the path's ray, or the shadow ray from its entry point towards light phase-1
    auto job_ray = [&](uint32_t slot, uint32_t phase, f3* o, f3* d) {
        const uint32_t i = sbase + slot;
        const f3 po = mk(s_ray[0][i], s_ray[1][i], s_ray[2][i]), pd = mk(s_ray[3][i], s_ray[4][i], s_ray[5][i]);
        if (phase == 0u) { *o = po; *d = pd; return; }
        const f3 hp = fma3(pd, s_t0[i], po);
        const f3 l = light_vec(c, phase - 1u, hp);
        *o = fma3(l, MRT_E, hp);  // Ray::cast_default, rt.rs:555-557
        *d = l;
    };

    // uniform across the warp: the pool of items and the two rings of slot numbers
    uint32_t pool_next = 0u, trav_head = 0u, trav_n = 0u, shade_head = 0u, shade_n = 0u;
    for (uint32_t base = 0u; base < S && pool_next < total; base += 32u) {  // the first paths
        const uint32_t slot = base + lane, item = pool_next + lane;
        if (item < total) { start_path(slot, item); s_qtrav[wi][(trav_n + lane) & (S - 1u)] = (uint8_t)slot; }
        const uint32_t n = min(32u, total - pool_next);
        pool_next += n; trav_n += n;
    }
    __syncwarp();

    uint32_t stack[MRT_WALK_STACK];
    float stack_t[MRT_WALK_STACK];
    Walk w;
    w.slot = MRT_WAVE_IDLE; w.phase = 0u; w.cur = MRT_WALK_EMPTY; w.sp = 0; w.in_mesh = false;
    w.o = w.d = w.m = w.nom = w.bm = w.bnom = mk(0.f, 0.f, 0.f);
    w.B.t0 = 0.f; w.B.t1 = 0.f; w.B.bi = -1; w.B.tr0 = w.B.tr1 = -1; w.B.any = false;
    w.m_first_tri = w.m_inst = 0u; w.b0 = w.b1 = 0.f; w.r0 = w.r1 = 0u; w.k0 = w.k1 = -1;

    // a walk is over: record its answer; go on with the segment's next search if there is one.  Returns true when the
    // slot has no search left (it then goes to the shade queue).
    auto walk_over = [&]() -> bool {
        for (;;) {
            const uint32_t i = sbase + w.slot;
            if (w.phase == 0u) {
                s_t0[i] = w.B.t0; s_t1[i] = w.B.t1; s_inst[i] = w.B.bi; s_tri0[i] = w.B.tr0; s_tri1[i] = w.B.tr1;
                s_vis[i] = 0u;
                if (w.B.bi < 0 || (F & F_LIGHTS) == 0 || MRT_N_LIGHTS(c) == 0u) return true;
            } else {
                if (w.B.bi < 0) s_vis[i] |= 1u << (w.phase - 1u);  // nothing in the way (no distance limit, rt.rs:1034-1038)
                if (w.phase >= MRT_N_LIGHTS(c)) return true;
            }
            w.phase++;
            f3 o, d;
            job_ray(w.slot, w.phase, &o, &d);
            if (!walk_begin<F>(w, gs, o, d)) return false;
        }
    };

    for (;;) {
        // ---------------------------------------------------------------- search
        for (;;) {
            const uint32_t idle = __ballot_sync(0xffffffffu, w.slot == MRT_WAVE_IDLE);
            bool done = false;
            if (idle != 0u && trav_n != 0u) {  // hand the queued slots to the lanes without a job
                const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
                if (w.slot == MRT_WAVE_IDLE && rank < trav_n) {
                    w.slot = s_qtrav[wi][(trav_head + rank) & (S - 1u)];
                    w.phase = 0u;
                    f3 o, d;
                    job_ray(w.slot, 0u, &o, &d);
                    if (walk_begin<F>(w, gs, o, d)) done = walk_over();
                }
                const uint32_t take = min((uint32_t)__popc(idle), trav_n);
                trav_head = (trav_head + take) & (S - 1u); trav_n -= take;
            }
            const uint32_t busy = __ballot_sync(0xffffffffu, w.slot != MRT_WAVE_IDLE);
            if (busy == 0u) break;
            if (w.slot != MRT_WAVE_IDLE && !done) {
                f3 wo = w.o, wd = w.d;
                if (MESH && w.in_mesh) job_ray(w.slot, w.phase, &wo, &wd);
                if (walk_turn<F>(w, stack, stack_t, gs, wo, wd)) done = walk_over();
            }
            const uint32_t dm = __ballot_sync(0xffffffffu, done);
            if (dm != 0u) {  // searched: on to the shade queue
                if (done) {
                    s_qshade[wi][(shade_head + shade_n + (uint32_t)__popc(dm & lt_mask)) & (S - 1u)] = (uint8_t)w.slot;
                    w.slot = MRT_WAVE_IDLE;
                }
                shade_n += (uint32_t)__popc(dm);
            }
            if (shade_n >= 32u) break;
            if (trav_n == 0u && shade_n != 0u && 32u - (uint32_t)__popc(busy & ~dm) >= MRT_WAVE_STARVE) break;
        }
        __syncwarp();
        if (shade_n == 0u) {
            if (trav_n == 0u && __ballot_sync(0xffffffffu, w.slot != MRT_WAVE_IDLE) == 0u) break;  // everything rendered
            continue;
        }
        // ---------------------------------------------------------------- shade: one searched path per lane
        const uint32_t n = min(32u, shade_n);
        const bool mine = lane < n;
        uint32_t slot = 0u;
        bool ended = false;
        if (mine) {
            slot = s_qshade[wi][(shade_head + lane) & (S - 1u)];
            const uint32_t i = sbase + slot;
            PathState p;
            p.o = mk(s_ray[0][i], s_ray[1][i], s_ray[2][i]); p.d = mk(s_ray[3][i], s_ray[4][i], s_ray[5][i]);
            p.T = mk(s_T[0][i], s_T[1][i], s_T[2][i]);
            p.pwr = s_pwr[i]; p.bounce = s_bounce[i];
            f3 L = mk(s_L[0][i], s_L[1][i], s_L[2][i]);
            const uint32_t tl = s_tl[i];
            HitRec h;
            h.t0 = s_t0[i]; h.t1 = s_t1[i]; h.inst = s_inst[i]; h.tri0 = s_tri0[i]; h.tri1 = s_tri1[i];
            if (h.inst < 0) { path_miss(c, p, L); ended = true; }
            else ended = shade_hit<F>(c, fp, __float_as_uint(s_q[wbase + tl].w), s_sample[i], p, L, h, s_vis[i]);
            if (ended) {  // the path's radiance goes to its pixel
                atomicAdd(&s_acc[0][wbase + tl], L.x); atomicAdd(&s_acc[1][wbase + tl], L.y); atomicAdd(&s_acc[2][wbase + tl], L.z);
            } else {
                s_ray[0][i] = p.o.x; s_ray[1][i] = p.o.y; s_ray[2][i] = p.o.z; s_ray[3][i] = p.d.x; s_ray[4][i] = p.d.y; s_ray[5][i] = p.d.z;
                s_T[0][i] = p.T.x; s_T[1][i] = p.T.y; s_T[2][i] = p.T.z;
                s_L[0][i] = L.x; s_L[1][i] = L.y; s_L[2][i] = L.z;
                s_pwr[i] = p.pwr; s_bounce[i] = p.bounce;
            }
        }
        shade_head = (shade_head + n) & (S - 1u); shade_n -= n;
        // a slot whose path ended takes the warp's next item; slots with a ray to search go back to the search queue
        const uint32_t em = __ballot_sync(0xffffffffu, mine && ended);
        bool queue = mine && !ended;
        if (mine && ended) {
            const uint32_t item = pool_next + (uint32_t)__popc(em & lt_mask);
            if (item < total) { start_path(slot, item); queue = true; }
        }
        pool_next = min(total, pool_next + (uint32_t)__popc(em));
        const uint32_t qm = __ballot_sync(0xffffffffu, queue);
        if (queue) s_qtrav[wi][(trav_head + trav_n + (uint32_t)__popc(qm & lt_mask)) & (S - 1u)] = (uint8_t)slot;
        trav_n += (uint32_t)__popc(qm);
        __syncwarp();
    }
    __syncwarp();
    if (inside) {
        float4 a = fp.accum[pix];
        a.x += s_acc[0][threadIdx.x]; a.y += s_acc[1][threadIdx.x]; a.z += s_acc[2][threadIdx.x];
        fp.accum[pix] = a;
    }
}

#endif


// ======================================================================================================================
// 4. (round 2, second session) Thin-lens loop with pre-generated camera rays — measured on the headline scene, B200:
//    plain loop 13 941 Mpaths/s; cadence 2: 13 178, cadence 4: 13 683, cadence 8: 13 425.  Rejected: the pop, the cadence
//    test and the spill-free bookkeeping cost what the rarer path starts save (the loop's dynamic instruction count is
//    unchanged, ~340 per iteration), and the latency of the path-start chain was already hidden by the other warps.
//    A first version with an idle state (a lane without a spare ray waits for the next top-up) ran at HALF speed: the compiler
//    left out the reconvergence point behind the path-start block (a `break` inside it), so the lanes that had just popped
//    a ray ran the whole segment code apart from the others.  Code as it ran (path_body dispatched LENS_THIN to it):
// ======================================================================================================================
#if 0
// Thin-lens loop with PRE-GENERATED camera rays (MRT_RING = cadence in iterations, a power of two; experiment).
// In the plain loop below the ~50-instruction path start runs in 96 % of the iterations for the ~5 lanes that need it: 12.5 % of
// the issued instructions and 19 % of the stall samples of the headline kernel (long dependent chain: hash -> int-to-float ->
// normalise -> MUFU.RSQ).  Here every lane keeps up to two ready rays in shared memory; all lanes top their rings up TOGETHER,
// one ray per lane, every MRT_RING-th iteration (the iteration counter is the same in every lane of a warp: no vote), and
// starting a path is a pop: two shared-memory loads.  A lane that has used up both spares before the next top-up (a run
// of very short paths) makes its ray on the spot, as the plain loop does.  Rays are keyed by the sample index, every lane
// still renders its samples in order: the image does not change by a bit.
#ifdef MRT_RING
template <class V, uint32_t F>
__device__ __forceinline__ void path_body_ring(const V sc, const FilmParams& fp) {
    __shared__ float4 s_q[128];        // q.xyz (pixel_focus_vec), lens seed (bits)
    __shared__ float4 s_od[2][128];    // ring slot: o.xyz, d.x
    __shared__ float2 s_dd[2][128];    //            d.y, d.z
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    if (px >= fp.nw || py >= fp.nh || fp.n_samples == 0u) return;
    const uint32_t pix = py * fp.nw + px;
    {
        const f3 q = pixel_focus_vec(fp, px, py);
        s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, __uint_as_float(cam_hash_seed(pix, fp.key)));
    }
    f3 acc = mk(0.f, 0.f, 0.f);
    PathState p;
    p.o = mk(0.f, 0.f, 0.f); p.d = mk(0.f, 1.f, 0.f); p.T = mk(1.f, 1.f, 1.f);
    p.pwr = 1.0f;
    p.bounce = MRT_NEED_PATH;
    uint32_t g = 0;  // rays generated so far: sample g goes to slot g & 1
    uint32_t c = 0;  // paths started so far; the running path is sample c - 1
    for (uint32_t it = 0;; it++) {
        if ((it & (uint32_t)(MRT_RING - 1)) == 0u) {
            if (g - c < 2u && g < fp.n_samples) {
                uint32_t tid;
                asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
                const float4 qs = s_q[tid];
                const float2 u = rng_cam(__float_as_uint(qs.w), fp.sample0 + g * fp.sample_stride);
                f3 o, d;
                camera_ray(fp, mk(qs.x, qs.y, qs.z), u.x, u.y, &o, &d);
                s_od[g & 1u][tid] = make_float4(o.x, o.y, o.z, d.x);
                s_dd[g & 1u][tid] = make_float2(d.y, d.z);
                g++;
            }
        }
        // (no `break` inside this block: with one the compiler leaves out the reconvergence point behind it, and the lanes
        // that have just started a path run the whole segment code apart from the others)
        if (p.bounce == MRT_NEED_PATH && c < fp.n_samples) {
            uint32_t tid;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
            if (c < g) {
                const float4 od = s_od[c & 1u][tid];
                const float2 dd = s_dd[c & 1u][tid];
                p.o = mk(od.x, od.y, od.z); p.d = mk(od.w, dd.x, dd.y);
            } else {  // both spares used up since the last top-up (a run of very short paths): make the ray here
                const float4 qs = s_q[tid];
                const float2 u = rng_cam(__float_as_uint(qs.w), fp.sample0 + c * fp.sample_stride);
                camera_ray(fp, mk(qs.x, qs.y, qs.z), u.x, u.y, &p.o, &p.d);
                g = c + 1u;
            }
            p.T = mk(1.f, 1.f, 1.f);
            p.pwr = 1.0f;
            p.bounce = 0;
            c++;
        }
        if (p.bounce == MRT_NEED_PATH) break;  // all samples rendered
        // (every lane that is still in the loop runs a segment: no lane ever waits for a ray, and the loop keeps the shape
        // of the plain one — with an idle state the compiler let the lanes that had just popped a ray run the whole segment
        // code a second time, apart from the others: half the speed)
        if (path_segment<V, F>(sc, fp, pix, fp.sample0 + (c - 1u) * fp.sample_stride, p, acc)) p.bounce = MRT_NEED_PATH;
    }
    MRT_CHECK(pix < fp.nw * fp.nh);
    float4 a = fp.accum[pix];
    a.x += acc.x; a.y += acc.y; a.z += acc.z;
    fp.accum[pix] = a;
}
#endif

#endif
