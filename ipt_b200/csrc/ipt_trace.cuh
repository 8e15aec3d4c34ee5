// ipt_trace.cuh — Geometry::traceRay on the device: the ordered analytic primitive list followed by the
// triangle mesh through the LBVH (stack traversal, shared-memory short stack with local-memory overflow).
//
// Semantics are those of the reference's linear scans (GeometrySphereInBox.cpp:23-36, FractalSpheres.cpp:75-84):
// candidates are considered in primitive order and a later one wins only if STRICTLY nearer. BVH traversal
// visits triangles in another order, so ties between triangles are resolved explicitly towards the lowest
// original triangle index, and nodes are culled with `tnear <= best`, never `<`.
#pragma once
#include "ipt_kernels.cuh"

namespace iptd {

#ifndef IPT_STACK_SHORT
#define IPT_STACK_SHORT 16  // entries per thread in shared memory (8 / 12 / 16 / 20 / 32: C4 343 / 359 / 365 / 350 / 319 Mpaths/s: beyond 16 the stacks eat into the L1 carve-out)
#endif
#define IPT_STACK_LOCAL (104 - IPT_STACK_SHORT)  // overflow entries in local memory (tree height bound 63 key bits + 32 index bits, + postponed leaves)
#define IPT_BLOCK 256

struct TravStack {
    uint32_t* sm; // interleaved: entry k of this thread at sm[k * IPT_BLOCK]
    uint32_t local[IPT_STACK_LOCAL];
    int n;
    __device__ __forceinline__ void push(uint32_t v) {
        if (n < IPT_STACK_SHORT) sm[n * IPT_BLOCK] = v;
        else local[n - IPT_STACK_SHORT] = v;
        ++n;
    }
    __device__ __forceinline__ uint32_t pop() {
        --n;
        return n < IPT_STACK_SHORT ? sm[n * IPT_BLOCK] : local[n - IPT_STACK_SHORT];
    }
};

// conservative slab test against a (padded) box; NaN-safe via fminf/fmaxf; returns entry distance or +inf
__device__ __forceinline__ float slab(float lox, float loy, float loz, float hix, float hiy, float hiz, f3 o, f3 inv, float best) {
    float t0x = (lox - o.x) * inv.x, t1x = (hix - o.x) * inv.x;
    float t0y = (loy - o.y) * inv.y, t1y = (hiy - o.y) * inv.y;
    float t0z = (loz - o.z) * inv.z, t1z = (hiz - o.z) * inv.z;
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    // 2 ulp of slack on the far side (Ize 2013), `<=` on both comparisons so that ties are never culled
    return (tn <= tf * 1.0000003f && tn <= best) ? tn : IPT_INF;
}

// ---- traversal of the 32-byte nodes (BvhNodeQ) in grid space ------------------------------------------------------------
// A ray is mapped once: entry / exit distances of a slab at grid coordinate g are  t = g * inv + c  with inv = 1 / d' and
// c = -o' * inv. t is the world-space ray parameter (GridMap), so it compares directly with the best hit distance. The
// quantised boxes are a full grid cell (1.5e-5) larger than the float boxes on every side; the float error of the mapped
// ray and of the fused multiply-add is below 1e-6 grid units, so the test stays conservative. 1 / d' is clamped to +-1e37:
// with an infinite reciprocal (a direction component that is exactly 0, which sampled directions do hit at a 2^-23 rate)
// g * inf - o' * inf would be NaN for every box, the axis would stop culling and the ray would walk the whole tree; with
// the clamp a ray outside the slab still gets two huge distances of the same sign and is culled.
#ifndef IPT_BVH_WIDE_NODES
#define IPT_BVH_WIDE_NODES 0 // 1: traverse the 64-byte float nodes instead (tuning A/B only)
#endif
struct GridRay {
    f3 inv, c;
};
__device__ __forceinline__ GridRay grid_ray(const GridMap& G, f3 o, f3 d) {
    GridRay r;
#if IPT_BVH_WIDE_NODES
    r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.c = o;
    return r;
#endif
    float ox = __fmaf_rn(__fsub_rn(o.x, G.lo[0]), G.scale[0], IPT_GRID_BASE), oy = __fmaf_rn(__fsub_rn(o.y, G.lo[1]), G.scale[1], IPT_GRID_BASE),
          oz = __fmaf_rn(__fsub_rn(o.z, G.lo[2]), G.scale[2], IPT_GRID_BASE);
    float dx = d.x * G.scale[0], dy = d.y * G.scale[1], dz = d.z * G.scale[2];
    r.inv = mk3(copysignf(fminf(fabsf(1.0f / dx), 1e37f), dx), copysignf(fminf(fabsf(1.0f / dy), 1e37f), dy), copysignf(fminf(fabsf(1.0f / dz), 1e37f), dz));
    r.c = mk3(-ox * r.inv.x, -oy * r.inv.y, -oz * r.inv.z);
    return r;
}
__device__ __forceinline__ float slab_q(uint32_t wa, uint32_t wb, uint32_t wc, const GridRay& R, float best) {
    // wa = lo.x | lo.y, wb = lo.z | hi.x, wc = hi.y | hi.z
    float t0x = __fmaf_rn(grid_lo16(wa), R.inv.x, R.c.x), t1x = __fmaf_rn(grid_hi16(wb), R.inv.x, R.c.x);
    float t0y = __fmaf_rn(grid_hi16(wa), R.inv.y, R.c.y), t1y = __fmaf_rn(grid_lo16(wc), R.inv.y, R.c.y);
    float t0z = __fmaf_rn(grid_lo16(wb), R.inv.z, R.c.z), t1z = __fmaf_rn(grid_hi16(wc), R.inv.z, R.c.z);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    return (tn <= tf * 1.0000003f && tn <= best) ? tn : IPT_INF;
}
// second child: words 3..5 hold lo1.x | lo1.y, lo1.z | hi1.x, hi1.y | hi1.z — the same layout

__device__ __forceinline__ void test_triangle(const DevScene& S, uint32_t pos, f3 o, f3 d, float& best_t, uint32_t& best_orig,
                                              uint32_t& best_pos, bool best_is_tri, TraceCounters& tc) {
    f8 r0 = ldg256(&S.tris[4 * (size_t)pos]);
    f8 r1 = ldg256(&S.tris[4 * (size_t)pos + 2]);
    ++tc.tris;
    f3 rel;
    float t = isect_parallelogram(mk3(r0.v[0], r0.v[1], r0.v[2]), mk3(r0.v[3], r0.v[4], r0.v[5]), mk3(r0.v[6], r0.v[7], r1.v[0]),
                                  mk3(r1.v[1], r1.v[2], r1.v[3]), true, o, d, &rel);
    if (t == IPT_INF) return;
    uint32_t orig = __float_as_uint(r1.v[4]);
    if (t < best_t || (t == best_t && best_is_tri && orig < best_orig)) {
        best_t = t;
        best_orig = orig;
        best_pos = pos;
    }
}

// Closest triangle nearer than `limit` (a triangle must be strictly nearer than the analytic winner).
// Returns true if one was found; t/orig/pos describe it.
__device__ __forceinline__ bool bvh_closest(const DevScene& S, f3 o, f3 d, float limit, float& t_out, uint32_t& orig_out,
                                            uint32_t& pos_out, uint32_t* smem_stack, TraceCounters& tc) {
    // best_t starts at `limit` with no triangle owning it: the first triangle must be STRICTLY nearer
    float best_t = limit;
    uint32_t best_orig = IPT_NO_HIT, best_pos = IPT_NO_HIT;
    if (S.n_tris == 1) {
        test_triangle(S, 0, o, d, best_t, best_orig, best_pos, false, tc);
        if (best_orig == IPT_NO_HIT) return false;
        t_out = best_t; orig_out = best_orig; pos_out = best_pos;
        return true;
    }
    const GridRay R = grid_ray(S.grid, o, d);
    TravStack st;
    st.sm = smem_stack + threadIdx.x;
    st.n = 0;
    uint32_t node = 0;
    while (true) {
#if IPT_BVH_WIDE_NODES
        f8 n0 = ldg256(&S.nodes[node]);
        f8 n1 = ldg256(reinterpret_cast<const char*>(&S.nodes[node]) + 32);
        ++tc.nodes;
        uint32_t left = __float_as_uint(n0.v[3]), right = __float_as_uint(n0.v[7]);
        float tn0 = slab(n0.v[0], n0.v[1], n0.v[2], n0.v[4], n0.v[5], n0.v[6], R.c, R.inv, best_t);
        float tn1 = slab(n1.v[0], n1.v[1], n1.v[2], n1.v[4], n1.v[5], n1.v[6], R.c, R.inv, best_t);
#else
        u8x32 q = ldg256u(&S.qnodes[node]);
        ++tc.nodes;
        uint32_t left = q.v[6], right = q.v[7];
        float tn0 = slab_q(q.v[0], q.v[1], q.v[2], R, best_t);
        float tn1 = slab_q(q.v[3], q.v[4], q.v[5], R, best_t);
#endif
        uint32_t next = IPT_NO_HIT;
        bool h0 = tn0 != IPT_INF, h1 = tn1 != IPT_INF;
        // leaves are intersected immediately; inner children are visited nearer-first
        if (h0 && (left & 0x80000000u)) {
            test_triangle(S, left & 0x7FFFFFFFu, o, d, best_t, best_orig, best_pos, best_orig != IPT_NO_HIT, tc);
            h0 = false;
        }
        if (h1 && (right & 0x80000000u)) {
            test_triangle(S, right & 0x7FFFFFFFu, o, d, best_t, best_orig, best_pos, best_orig != IPT_NO_HIT, tc);
            h1 = false;
        }
        if (h0 && h1) {
            bool first0 = tn0 <= tn1;
            next = first0 ? left : right;
            st.push(first0 ? right : left);
        } else if (h0) next = left;
        else if (h1) next = right;
        if (next == IPT_NO_HIT) {
            if (st.n == 0) break;
            next = st.pop();
        }
        node = next;
    }
    if (best_orig == IPT_NO_HIT) return false;
    t_out = best_t; orig_out = best_orig; pos_out = best_pos;
    return true;
}

extern __shared__ uint32_t ipt_dyn_smem[];

// the ordered analytic primitive list: grouped planes + statically unrolled spheres when the scene allows, else the
// generic ordered scan
// GFAST: the caller knows at compile time that the scene takes the first branch all the way (grouped planes and
// inline spheres only), so none of the generic scans is instantiated
// X = false: contracted arithmetic in the grouped-plane / inline-sphere branch (secondary rays, see ipt_device.cuh); the
// generic ordered scans keep the exact routines
template <bool SMALLPT, bool GFAST = false, bool X = true>
__device__ __forceinline__ void analytic_closest(const DevScene& S, f3 o, f3 d, double& dist_d, float& dist_f, uint32_t& best, bool skip_planes = false) {
    if (GFAST || (!SMALLPT && S.planes_grouped)) {
        // skip_planes (warp-uniform): an occlusion test towards a light that lies strictly inside the convex box the planes
        // bound (DevScene::lights_inside_box): the segment to the light cannot cross a wall, only the other primitives matter
        if (S.n_planes && !skip_planes) {
            isect_axis_planes<X, 0>(S.plane_of[0], S.plane_of[1], o.x, d.x, o, d, dist_f, best);
            isect_axis_planes<X, 1>(S.plane_of[2], S.plane_of[3], o.y, d.y, o, d, dist_f, best);
            isect_axis_planes<X, 2>(S.plane_of[4], S.plane_of[5], o.z, d.z, o, d, dist_f, best);
        }
        if (GFAST || S.others_inline) {
            // static indices: every sphere constant is an immediate constant-bank operand, no indexed LDC
#pragma unroll
            for (int k = 0; k < IPT_INLINE_OTHERS; ++k) {
                if (k < (int)S.n_others) {
                    const DevSphere& sp = S.others[k];
                    float t = isect_sphere<X>(sp.r2, xsub3(o, mk3(sp.cx, sp.cy, sp.cz)), d);
                    if (t < dist_f || (t == dist_f && t != IPT_INF && sp.index < best)) { dist_f = t; best = sp.index; }
                }
            }
        } else if (!GFAST && S.n_prims > S.n_planes) {
            if (S.prim_inline) trace_prim_list<false, true>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims[i]; }, o, d, dist_d, dist_f, best);
            else trace_prim_list<false, true>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims_g[i]; }, o, d, dist_d, dist_f, best);
        }
    } else if (!GFAST) {
        if (S.prim_inline) trace_prim_list<SMALLPT, false>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims[i]; }, o, d, dist_d, dist_f, best);
        else trace_prim_list<SMALLPT, false>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims_g[i]; }, o, d, dist_d, dist_f, best);
    }
}

template <bool SMALLPT, bool MESH, bool GFAST, bool X>
__device__ __forceinline__ SurfHit trace_geometry(const DevScene& S, f3 o, f3 d, TraceCounters& tc, bool skip_planes) {
    double dist_d = (double)IPT_INF;
    float dist_f = IPT_INF;
    uint32_t best = IPT_NO_HIT;
    analytic_closest<SMALLPT, GFAST, X>(S, o, d, dist_d, dist_f, best, skip_planes);
    SurfHit r;
    r.prim = best;
    r.tri_pos = IPT_NO_HIT;
    r.t = SMALLPT ? __double2float_rn(dist_d) : dist_f;
    if (MESH) {
        // a float candidate t beats the analytic winner iff (double)t < dist_d, which for float winners is t < dist_f
        float limit = SMALLPT ? __double2float_ru(dist_d) : dist_f;
        float t;
        uint32_t orig, pos;
        if (bvh_closest(S, o, d, limit, t, orig, pos, ipt_dyn_smem, tc) && (!SMALLPT || (double)t < dist_d)) {
            r.prim = S.n_prims + orig;
            r.tri_pos = pos;
            r.t = t;
        }
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// K2 for mesh scenes: persistent warps with ray replenishment (Aila & Laine 2009). BVH traversal lengths differ by
// orders of magnitude between the rays of a warp (measured: 2.7-4.6 active lanes of 32 in the one-ray-per-thread
// kernel), so a lane that finishes its ray fetches the next one from the queue instead of idling until the slowest
// lane of its warp is done. Per-ray work is split into setup (analytic primitives, and the lights at the last depth),
// traversal steps (one BVH node each) and finalisation (decision, emission / compaction), exactly the arithmetic of
// trace_scene / trace_scene_last.
// ---------------------------------------------------------------------------------------------------------------
#ifndef IPT_REFILL_MIN
#define IPT_REFILL_MIN 4    // fetch new rays when at least this many lanes are idle
#endif
#ifndef IPT_TRAV_STEPS
#define IPT_TRAV_STEPS 16    // node visits between two refill checks
#endif
#ifndef IPT_VISITS_PER_ROUND
#define IPT_VISITS_PER_ROUND 3 // node visits per lane between two rounds of warp votes (finished lanes, waiting lanes, refill). With the pair queue: 1 / 2 / 3 / 4 give C3 436 / 459 / 464 / 410, C4 394 / 415 / 419 / 369 Mpaths/s
#endif
#ifndef IPT_LEAF_BATCH
#define IPT_LEAF_BATCH 6    // run the postponed triangle tests once this many lanes hold one
#endif

#ifndef IPT_MESH_MIN_BLOCKS
#define IPT_MESH_MIN_BLOCKS 4
#endif
#ifndef IPT_PAIR_QUEUE
#define IPT_PAIR_QUEUE 1 // 1: (owner lane, triangle) pairs are queued per warp and tested 32 at a time; 0: a lane tests its own postponed leaf
#endif
#define IPT_POOL 64 // entries per warp and pool: fewer than 32 wait when up to 32 more arrive
// dynamic shared memory of k_extend_mesh: the short stacks, then per warp a READY pool (ray index, analytic hit distance,
// analytic primitive) and a DONE pool (ray index, hit distance, primitive id, primitive id as queued)
// + the pair queue: owner lane and triangle of 64 pairs, and per lane a 64-bit best key, the winner's sorted position, a count
#define IPT_MESH_POOL_WORDS (7 * IPT_POOL + (IPT_PAIR_QUEUE ? 2 * IPT_POOL + 4 * 32 : 0))
#define IPT_MESH_SMEM_BYTES ((IPT_STACK_SHORT * IPT_BLOCK + (IPT_BLOCK / 32) * IPT_MESH_POOL_WORDS) * sizeof(uint32_t))

// SPEC (see SceneSpec in ipt_kernels.cuh): SPEC_BOX_SCENE = one inline area light, analytic part = grouped box planes +
// inline spheres — the kernel then carries no light-LBVH walk (and not its local stack), no sphere-light code and none of
// the generic primitive scans; SPEC_RUNTIME decides everything from the scene.
//
// Three phases per warp, each run by (nearly) all 32 lanes:
//   setup     32 rays are fetched from the queue together; every lane intersects the analytic primitives for its ray (at the
//             last traced depth: the lights first — a ray that reaches none is finished) and parks (ray, analytic hit) in
//             the warp's READY pool;
//   traverse  a lane without a ray takes one from the READY pool and walks the LBVH; a leaf it reaches is not tested by the
//             lane but queued as an (owner lane, triangle) pair, and whenever 32 pairs wait the warp tests them together
//             (IPT_PAIR_QUEUE: the exact triangle test at 32 lanes instead of the 6.5 that happened to hold a leaf); a
//             finished lane parks (ray, closest hit) in the DONE pool and takes the next ray at once;
//   finalise  whenever 32 results wait, the warp runs the light test / decision / emission / hit append of main.cpp:111-128
//             for them.
// Round 1's kernel ran setup and finalisation on the 4-8 lanes that happened to be idle or finished (profiles/tuning_r02.md:
// c3_tree +8.7 % for the pools); the arithmetic per ray is unchanged.
template <bool LAST, int SPEC = SPEC_RUNTIME>
__global__ void __launch_bounds__(IPT_BLOCK, IPT_MESH_MIN_BLOCKS) k_extend_mesh(const __grid_constant__ DevScene S, const __grid_constant__ RenderCtx C, uint32_t depth) {
    const uint32_t n = C.cnt[2 * depth];
    uint32_t* next = &C.fetch[depth];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool shadow = LAST && !(C.flags & IPT_FLAG_RESOLVE_LAST_LEVEL); // last traced depth: only rays that reach a light matter
    uint32_t* pool = ipt_dyn_smem + IPT_STACK_SHORT * IPT_BLOCK + (threadIdx.x >> 5) * IPT_MESH_POOL_WORDS;
    uint32_t* rd_idx = pool;                 float* rd_at = reinterpret_cast<float*>(pool + IPT_POOL); uint32_t* rd_ap = pool + 2 * IPT_POOL;
    uint32_t* dn_idx = pool + 3 * IPT_POOL;  float* dn_t = reinterpret_cast<float*>(pool + 4 * IPT_POOL);
    uint32_t* dn_prim = pool + 5 * IPT_POOL; uint32_t* dn_iprim = pool + 6 * IPT_POOL;
#if IPT_PAIR_QUEUE
    uint32_t* pq_own = pool + 7 * IPT_POOL;  uint32_t* pq_tri = pool + 8 * IPT_POOL;
    unsigned long long* bk = reinterpret_cast<unsigned long long*>(pool + 9 * IPT_POOL); // 32 x u64 (8-byte aligned: pool offsets are even)
    uint32_t* bp = pool + 9 * IPT_POOL + 64;  uint32_t* dc = pool + 9 * IPT_POOL + 96;
    uint32_t pq_n = 0;    // warp-uniform: pairs waiting
    uint32_t pending = 0; // pairs of THIS lane's ray that wait in the queue
#endif
    uint32_t n_ready = 0, n_done = 0; // warp-uniform
    bool exhausted = false;           // warp-uniform
    uint32_t n_surface = 0, n_light = 0, n_miss = 0, n_dropped = 0;
    TraceCounters tc{0, 0, 0, 0};
    // the ray this lane is walking the LBVH for
    bool have = false, trav = false;
    uint32_t ridx = 0, a_prim = IPT_NO_HIT, best_orig = IPT_NO_HIT, best_pos = IPT_NO_HIT, node = 0, pend = IPT_NO_HIT;
    float a_t = IPT_INF, best_t = IPT_INF;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 1);
    GridRay R;
    R.inv = mk3(0, 0, 0); R.c = mk3(0, 0, 0);
    TravStack st;
    st.sm = ipt_dyn_smem + threadIdx.x;
    st.n = 0;
    while (true) {
        // ---- setup: keep at least 32 rays ready
        while (!exhausted && n_ready < 32) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next, 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + 32 >= n) exhausted = true;
            uint32_t i = base + lane;
            bool go = false;
            float at = IPT_INF;
            uint32_t ap = IPT_NO_HIT;
            if (i < n) {
                float4 ro = C.ray_o[i], rdv = C.ray_d[i];
                f3 so = mk3(ro.x, ro.y, ro.z), sd = mk3(rdv.x, rdv.y, rdv.z);
                go = true;
                if (shadow) { // nothing to do unless a light lies along the ray
                    uint32_t lw = IPT_NO_HIT;
                    f3 lp;
                    float lpdf, ld;
                    go = trace_lights<false, SPEC>(S, so, sd, lw, lp, lpdf, ld, tc);
                }
                if (go) {
                    double dd = (double)IPT_INF;
                    analytic_closest<false, IPT_SPEC_FAST_GEOMETRY(SPEC)>(S, so, sd, dd, at, ap);
                }
            }
            uint32_t gb = __ballot_sync(0xffffffffu, go);
            if (go) {
                uint32_t k = n_ready + __popc(gb & lt_mask);
                rd_idx[k] = i; rd_at[k] = at; rd_ap[k] = ap;
            }
            n_ready += __popc(gb);
            __syncwarp();
        }
        // ---- lanes without a ray take one from the READY pool
        {
            uint32_t idle = __ballot_sync(0xffffffffu, !have);
            if (idle && n_ready) {
                uint32_t k = __popc(idle & lt_mask);
                if (!have && k < n_ready) {
                    uint32_t e = n_ready - 1 - k;
                    ridx = rd_idx[e]; a_t = rd_at[e]; a_prim = rd_ap[e];
                    float4 ro = C.ray_o[ridx], rdv = C.ray_d[ridx];
                    o = mk3(ro.x, ro.y, ro.z); d = mk3(rdv.x, rdv.y, rdv.z);
                    best_t = a_t; best_orig = IPT_NO_HIT; best_pos = IPT_NO_HIT;
                    have = true;
                    if (S.n_tris == 1) {
                        test_triangle(S, 0, o, d, best_t, best_orig, best_pos, false, tc);
                        trav = false;
                    } else {
                        R = grid_ray(S.grid, o, d);
                        node = 0; pend = IPT_NO_HIT; st.n = 0; trav = true;
                    }
                }
                n_ready -= min(n_ready, (uint32_t)__popc(idle));
                __syncwarp();
            }
        }
        const bool drain = exhausted && n_ready == 0 && __ballot_sync(0xffffffffu, have) == 0;
        // ---- traverse: a few BVH node visits. Triangle tests are POSTPONED (speculative while-while traversal, Aila & Laine
        // 2009): a hit leaf is parked in `pend` (further ones go on the stack) and lanes keep walking inner nodes; the exact
        // triangle test then runs for many lanes at once instead of for the 1-2 lanes that reach a leaf in the same step.
        if (!drain) {
            uint32_t idle_mask = __ballot_sync(0xffffffffu, !have); // warp-uniform, maintained through the rounds below
#if IPT_PAIR_QUEUE
            // 32 queued (owner lane, triangle) pairs at a time: every lane fetches its pair's ray from the owner by shuffle, runs
            // the exact triangle test and folds a hit into the owner's best (distance, original index) with a 64-bit
            // shared-memory atomicMin — the tie-break of the linear scan (equal distance: lowest original index) is the order
            // of those keys; a triangle must be STRICTLY nearer than the analytic winner, whose key carries index 0.
            auto test_pairs = [&](uint32_t cnt) {
                const unsigned long long key0 = ((unsigned long long)__float_as_uint(best_t) << 32) | (best_orig == IPT_NO_HIT ? 0u : best_orig);
                bk[lane] = have ? key0 : ~0ull;
                dc[lane] = 0;
                __syncwarp();
                const bool valid = lane < cnt;
                const uint32_t e = pq_n - cnt + lane;
                const uint32_t own = valid ? pq_own[e] : lane, pos = valid ? pq_tri[e] : 0u;
                f3 po = mk3(__shfl_sync(0xffffffffu, o.x, own), __shfl_sync(0xffffffffu, o.y, own), __shfl_sync(0xffffffffu, o.z, own));
                f3 pd = mk3(__shfl_sync(0xffffffffu, d.x, own), __shfl_sync(0xffffffffu, d.y, own), __shfl_sync(0xffffffffu, d.z, own));
                unsigned long long key = ~0ull;
                if (valid) {
                    f8 r0 = ldg256(&S.tris[4 * (size_t)pos]);
                    f8 r1 = ldg256(&S.tris[4 * (size_t)pos + 2]);
                    ++tc.tris;
                    f3 rel;
                    float t = isect_parallelogram(mk3(r0.v[0], r0.v[1], r0.v[2]), mk3(r0.v[3], r0.v[4], r0.v[5]), mk3(r0.v[6], r0.v[7], r1.v[0]),
                                                  mk3(r1.v[1], r1.v[2], r1.v[3]), true, po, pd, &rel);
                    if (t != IPT_INF) {
                        key = ((unsigned long long)__float_as_uint(t) << 32) | __float_as_uint(r1.v[4]);
                        atomicMin(&bk[own], key);
                    }
                    atomicAdd(&dc[own], 1u);
                }
                __syncwarp();
                if (key != ~0ull && bk[own] == key) bp[own] = pos; // the winner leaves its sorted position
                __syncwarp();
                const unsigned long long k = bk[lane];
                if (have && k != key0) { best_t = __uint_as_float((uint32_t)(k >> 32)); best_orig = (uint32_t)k; best_pos = bp[lane]; }
                pending -= dc[lane];
                pq_n -= cnt;
                __syncwarp();
            };
            auto queue_leaf = [&](uint32_t leaf) { // warp-wide: lanes with leaf != IPT_NO_HIT append (lane, sorted triangle position)
                const uint32_t m = __ballot_sync(0xffffffffu, leaf != IPT_NO_HIT);
                if (m) {
                    if (leaf != IPT_NO_HIT) {
                        const uint32_t k = pq_n + __popc(m & lt_mask);
                        pq_own[k] = lane; pq_tri[k] = leaf;
                        ++pending;
                    }
                    pq_n += __popc(m);
                    __syncwarp();
                    if (pq_n >= 32) test_pairs(32);
                }
            };
            for (int step = 0; step < IPT_TRAV_STEPS; ++step) {
#pragma unroll
                for (int visit = 0; visit < IPT_VISITS_PER_ROUND; ++visit) {
                    uint32_t leaf_a = IPT_NO_HIT, leaf_b = IPT_NO_HIT;
                    if (have && trav && node != IPT_NO_HIT) {
#if IPT_BVH_WIDE_NODES
                        f8 n0 = ldg256(&S.nodes[node]);
                        f8 n1 = ldg256(reinterpret_cast<const char*>(&S.nodes[node]) + 32);
                        ++tc.nodes;
                        uint32_t left = __float_as_uint(n0.v[3]), right = __float_as_uint(n0.v[7]);
                        float tn0 = slab(n0.v[0], n0.v[1], n0.v[2], n0.v[4], n0.v[5], n0.v[6], R.c, R.inv, best_t);
                        float tn1 = slab(n1.v[0], n1.v[1], n1.v[2], n1.v[4], n1.v[5], n1.v[6], R.c, R.inv, best_t);
#else
                        u8x32 q = ldg256u(&S.qnodes[node]);
                        ++tc.nodes;
                        uint32_t left = q.v[6], right = q.v[7];
                        float tn0 = slab_q(q.v[0], q.v[1], q.v[2], R, best_t);
                        float tn1 = slab_q(q.v[3], q.v[4], q.v[5], R, best_t);
#endif
                        bool h0 = tn0 != IPT_INF, h1 = tn1 != IPT_INF;
                        if (h0 && (left & 0x80000000u)) { leaf_a = left & 0x7FFFFFFFu; h0 = false; }
                        if (h1 && (right & 0x80000000u)) { leaf_b = right & 0x7FFFFFFFu; h1 = false; }
                        uint32_t nxt = IPT_NO_HIT;
                        if (h0 && h1) {
                            bool first0 = tn0 <= tn1;
                            nxt = first0 ? left : right;
                            st.push(first0 ? right : left);
                        } else if (h0) nxt = left;
                        else if (h1) nxt = right;
                        if (nxt == IPT_NO_HIT && st.n) nxt = st.pop(); // the stack holds inner nodes only
                        node = nxt;
                    }
                    queue_leaf(leaf_a);
                    queue_leaf(leaf_b);
                }
                // a lane whose walk is over waits for its queued pairs; they are tested when enough lanes wait, when nobody walks
                // any more, and at the end of every block of rounds
                {
                    const uint32_t walking = __ballot_sync(0xffffffffu, have && trav && node != IPT_NO_HIT);
                    const uint32_t waiting = __ballot_sync(0xffffffffu, have && trav && node == IPT_NO_HIT && pending != 0);
                    if (pq_n && (walking == 0 || __popc(waiting) >= IPT_LEAF_BATCH || step == IPT_TRAV_STEPS - 1))
                        while (pq_n) test_pairs(min(pq_n, 32u));
                    if (have && trav && node == IPT_NO_HIT && pending == 0) trav = false;
                }
#else
            for (int step = 0; step < IPT_TRAV_STEPS; ++step) {
#pragma unroll
                for (int visit = 0; visit < IPT_VISITS_PER_ROUND; ++visit)
                if (have && trav) {
                    if (node != IPT_NO_HIT && pend == IPT_NO_HIT) {
#if IPT_BVH_WIDE_NODES
                        f8 n0 = ldg256(&S.nodes[node]);
                        f8 n1 = ldg256(reinterpret_cast<const char*>(&S.nodes[node]) + 32);
                        ++tc.nodes;
                        uint32_t left = __float_as_uint(n0.v[3]), right = __float_as_uint(n0.v[7]);
                        float tn0 = slab(n0.v[0], n0.v[1], n0.v[2], n0.v[4], n0.v[5], n0.v[6], R.c, R.inv, best_t);
                        float tn1 = slab(n1.v[0], n1.v[1], n1.v[2], n1.v[4], n1.v[5], n1.v[6], R.c, R.inv, best_t);
#else
                        u8x32 q = ldg256u(&S.qnodes[node]);
                        ++tc.nodes;
                        uint32_t left = q.v[6], right = q.v[7];
                        float tn0 = slab_q(q.v[0], q.v[1], q.v[2], R, best_t);
                        float tn1 = slab_q(q.v[3], q.v[4], q.v[5], R, best_t);
#endif
                        bool h0 = tn0 != IPT_INF, h1 = tn1 != IPT_INF;
                        if (h0 && (left & 0x80000000u)) { pend = left; h0 = false; }
                        if (h1 && (right & 0x80000000u)) {
                            if (pend == IPT_NO_HIT) pend = right; else st.push(right);
                            h1 = false;
                        }
                        uint32_t nxt = IPT_NO_HIT;
                        if (h0 && h1) {
                            bool first0 = tn0 <= tn1;
                            nxt = first0 ? left : right;
                            st.push(first0 ? right : left);
                        } else if (h0) nxt = left;
                        else if (h1) nxt = right;
                        node = nxt;
                    }
                    if (node == IPT_NO_HIT && pend == IPT_NO_HIT) {
                        if (st.n == 0) trav = false;
                        else {
                            uint32_t x = st.pop();
                            if (x & 0x80000000u) pend = x; else node = x;
                        }
                    }
                }
                // two votes per round: lanes still traversing, and those of them that hold a postponed leaf
                const uint32_t busy = __ballot_sync(0xffffffffu, have && trav);
                const uint32_t pm = __ballot_sync(0xffffffffu, have && trav && pend != IPT_NO_HIT);
                if (pm && (__popc(pm) >= IPT_LEAF_BATCH || pm == busy || step == IPT_TRAV_STEPS - 1)) {
                    if (have && trav && pend != IPT_NO_HIT) {
                        test_triangle(S, pend & 0x7FFFFFFFu, o, d, best_t, best_orig, best_pos, best_orig != IPT_NO_HIT, tc);
                        pend = IPT_NO_HIT;
                        if (node == IPT_NO_HIT && st.n == 0) trav = false;
                    }
                }
#endif
                // finished lanes park their result and become idle
                bool fin = have && !trav;
                uint32_t fb = __ballot_sync(0xffffffffu, fin);
                if (fb) {
                    if (fin) {
                        uint32_t k = n_done + __popc(fb & lt_mask);
                        bool tri = best_orig != IPT_NO_HIT;
                        dn_idx[k] = ridx; dn_t[k] = tri ? best_t : a_t;
                        dn_prim[k] = tri ? S.n_prims + best_orig : a_prim;
                        dn_iprim[k] = tri ? S.n_prims + best_pos : a_prim;
                        have = false;
                    }
                    n_done += __popc(fb);
                    __syncwarp();
                }
                idle_mask |= fb; // (the lanes that were idle when the round began, plus the ones that just finished)
                const uint32_t idle = idle_mask;
                if (n_done >= 32 || idle == 0xffffffffu) break;
                if (__popc(idle) >= IPT_REFILL_MIN && (n_ready || !exhausted)) break;
            }
        }
        // ---- finalise 32 finished rays (main.cpp:111-128), compact the survivors
        while (n_done >= 32 || (drain && n_done)) {
            const uint32_t cnt = min(n_done, 32u);
            const uint32_t e = n_done - cnt + lane;
            const bool valid = lane < cnt;
            n_done -= cnt;
            bool emit = false;
            float4 ro = make_float4(0, 0, 0, 0), rdv = make_float4(0, 0, 1, 0);
            float t = IPT_INF;
            uint32_t iprim = IPT_NO_HIT;
            if (valid) {
                const uint32_t i = dn_idx[e];
                t = dn_t[e];
                const uint32_t prim = dn_prim[e];
                iprim = dn_iprim[e];
                ro = C.ray_o[i]; rdv = C.ray_d[i];
                const float sv = C.ray_x[i];
                f3 fo = mk3(ro.x, ro.y, ro.z), fd = mk3(rdv.x, rdv.y, rdv.z);
                uint32_t lwhich = IPT_NO_HIT;
                f3 lpos = mk3(0, 0, 0);
                float lpdf = 0.0f, ldist = 0.0f;
                bool lh;
                // (at the last traced depth this repeats the light test of the setup phase — the ray is here because it hit —
                // and is not counted a second time)
                TraceCounters uncounted{0, 0, 0, 0};
                TraceCounters& tcl = shadow ? uncounted : tc;
                if (lights_inline<SPEC>(S)) lh = trace_lights<false, SPEC>(S, fo, fd, lwhich, lpos, lpdf, ldist, tcl);
                else lh = trace_lights<true, SPEC>(S, fo, fd, lwhich, lpos, lpdf, ldist, tcl);
                // single inline light: its density follows from the hit position that is kept anyway
                static_assert(IPT_INLINE_LIGHTS == 1, "the lights[0] shortcut below assumes that 'inline lights' means exactly one light");
                if (lights_inline<SPEC>(S) && lh && S.n_lights) lpdf = S.lights[0].weight * light_pdf_at<IPT_SPEC_AREA_LIGHTS(SPEC)>(S.lights[0], fo, lpos);
                const bool sh = prim != IPT_NO_HIT;
#ifdef IPT_DEBUG_PRINT
                float K = ro.w;
#endif
                ro.w = resolve_weight(S, ro.w, sv, lpdf);
#ifdef IPT_DEBUG_PRINT
                if (C.flags & IPT_FLAG_DEBUG_PRINT)
                    printf("GPU mesh extend d=%u o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) K=%.9g sv=%.9g thr=%.9g lh=%d lpos=(%.9g %.9g %.9g) prim=%u t=%.9g\n", depth, fo.x, fo.y, fo.z,
                           fd.x, fd.y, fd.z, K, sv, ro.w, (int)lh, lpos.x, lpos.y, lpos.z, prim, t);
#endif
                if (!isfinite(ro.w)) {
                    ++n_dropped; // non-finite multiplier (main.cpp:175): drop this sample
                } else {
                    bool light_wins = false;
                    if (lh) {
                        float len_surf = sh ? xlength3(xsub3(xpoint(fo, fd, t), fo)) : IPT_INF;
                        light_wins = !sh || len_surf > ldist; // main.cpp:113 (ldist = length(light position - origin))
                    }
                    if (light_wins) {
                        ++n_light;
                        atomicAdd(&C.pathval[__float_as_uint(rdv.w) & C.slot_mask], ro.w * light_power<SPEC>(S, lwhich));
                    } else if (sh) {
                        ++n_surface;
                        emit = !LAST;
                    } else {
                        ++n_miss;
                    }
                }
            }
            if (!LAST) {
                uint32_t ballot = __ballot_sync(0xffffffffu, emit);
                if (ballot) {
                    uint32_t basepos = 0;
                    if (lane == 0) basepos = atomicAdd(&C.cnt[2 * depth + 1], (uint32_t)__popc(ballot));
                    basepos = __shfl_sync(0xffffffffu, basepos, 0);
                    if (emit) {
                        uint32_t j = basepos + __popc(ballot & lt_mask);
                        f3 p = xpoint(mk3(ro.x, ro.y, ro.z), mk3(rdv.x, rdv.y, rdv.z), t);
                        float2 oct = oct_encode(mk3(rdv.x, rdv.y, rdv.z));
                        if (IPT_BOUNDS_OK(j, C.hit_cap, C.stats)) {
                            C.hit_a[depth & 1][j] = make_float4(p.x, p.y, p.z, ro.w);
                            C.hit_b[depth & 1][j] = make_uint4(__float_as_uint(rdv.w), iprim, __float_as_uint(oct.x), __float_as_uint(oct.y));
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (drain) break;
    }
    flush_stat(C.stats, ST_SURFACE, n_surface);
    flush_stat(C.stats, ST_LIGHT, n_light);
    flush_stat(C.stats, ST_MISS, n_miss);
    flush_stat(C.stats, ST_DROPPED, n_dropped);
    flush_stat(C.stats, ST_NODES, tc.nodes);
    flush_stat(C.stats, ST_TRIS, tc.tris);
    flush_stat(C.stats, ST_LIGHTS, tc.lights);
    flush_stat(C.stats, ST_LIGHT_NODES, tc.light_nodes);
}

} // namespace iptd
