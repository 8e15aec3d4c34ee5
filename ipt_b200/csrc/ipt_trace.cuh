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

#define IPT_STACK_SHORT 12  // entries per thread in shared memory
#define IPT_STACK_LOCAL 84  // overflow entries in local memory (tree height bound: 63 key bits + 32 index bits)
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

__device__ __forceinline__ void test_triangle(const DevScene& S, uint32_t pos, f3 o, f3 d, float& best_t, uint32_t& best_orig,
                                              uint32_t& best_pos, bool best_is_tri, TraceCounters& tc) {
    float4 a = __ldg(&S.tris[3 * (size_t)pos]);
    float4 b = __ldg(&S.tris[3 * (size_t)pos + 1]);
    float4 c = __ldg(&S.tris[3 * (size_t)pos + 2]);
    ++tc.tris;
    f3 rel;
    float t = isect_parallelogram(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), mk3(c.y, c.z, c.w), true, o, d, &rel);
    if (t == IPT_INF) return;
    uint32_t orig = __ldg(&S.tri_id[pos]);
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
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    TravStack st;
    st.sm = smem_stack + threadIdx.x;
    st.n = 0;
    uint32_t node = 0;
    while (true) {
        const float4* np = reinterpret_cast<const float4*>(&S.nodes[node]);
        float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
        ++tc.nodes;
        uint32_t left = __float_as_uint(q0.w), right = __float_as_uint(q1.w);
        float tn0 = slab(q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, o, inv, best_t);
        float tn1 = slab(q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, o, inv, best_t);
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

template <bool SMALLPT, bool MESH>
__device__ __forceinline__ SurfHit trace_geometry(const DevScene& S, f3 o, f3 d, TraceCounters& tc) {
    double dist_d = (double)IPT_INF;
    float dist_f = IPT_INF;
    uint32_t best = IPT_NO_HIT;
    if (!SMALLPT && S.planes_grouped) {
        if (S.n_planes) {
            isect_axis_planes(S.plane_of[0], S.plane_of[1], o.x, d.x, o, d, dist_f, best);
            isect_axis_planes(S.plane_of[2], S.plane_of[3], o.y, d.y, o, d, dist_f, best);
            isect_axis_planes(S.plane_of[4], S.plane_of[5], o.z, d.z, o, d, dist_f, best);
        }
        if (S.others_inline) {
            // static indices: every sphere constant is an immediate constant-bank operand, no indexed LDC
#pragma unroll
            for (int k = 0; k < IPT_INLINE_OTHERS; ++k) {
                if (k < (int)S.n_others) {
                    const DevSphere& sp = S.others[k];
                    float t = isect_sphere(sp.r2, xsub3(o, mk3(sp.cx, sp.cy, sp.cz)), d);
                    if (t < dist_f || (t == dist_f && t != IPT_INF && sp.index < best)) { dist_f = t; best = sp.index; }
                }
            }
        } else if (S.n_prims > S.n_planes) {
            if (S.prim_inline) trace_prim_list<false, true>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims[i]; }, o, d, dist_d, dist_f, best);
            else trace_prim_list<false, true>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims_g[i]; }, o, d, dist_d, dist_f, best);
        }
    } else {
        if (S.prim_inline) trace_prim_list<SMALLPT, false>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims[i]; }, o, d, dist_d, dist_f, best);
        else trace_prim_list<SMALLPT, false>(S.n_prims, [&S](uint32_t i) -> const DevPrim& { return S.prims_g[i]; }, o, d, dist_d, dist_f, best);
    }
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

} // namespace iptd
