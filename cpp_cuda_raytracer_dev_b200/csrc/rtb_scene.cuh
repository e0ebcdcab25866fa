// rtb_scene.cuh -- the scene kernel: what the reference leaves dormant around its hot path (SURVEY.md section 8(f) items
// 3-4): several objects per camera, a light list, the shadow test of Camera.cu:33, sample_rate^2 rays per pixel.
//
// The reference draws one object with one light and one ray per pixel, and that case is the persistent kernel of
// rtb_render.cuh.  Its code carries the stubs of more -- Camera::object_list (Camera.cpp:118-130) with a second object
// registered by WinMain.cpp:153-156 and its render commented out (:214), the commented light loop and shadow test of
// color_cam_cuda (Camera.cu:28-34, 54), Camera::render_properites::sample_rate (Camera.h:46) -- without any behaviour to be
// identical to.  DESIGN.md section 11 defines the behaviour (the smallest completion of those stubs that leaves the
// default output untouched), oracle/rtb_oracle.c (orc_render_scene) restates it on the CPU, and this kernel is compared
// with that oracle bit for bit:
//   objects     : traversed in registration order with the closest distance carried over; the strict `w < best` of
//                 Trixel.cu:127 makes the closest hit win and the first registered object win ties; id = id_base + triangle
//   lights      : radiance summed over the lights in order (the commented loop), then the max-channel normalise
//   shadows     : a light counts only if the segment hit point -> light (direction light - pnt, unnormalised, parameter
//                 1e-4 < t < 1) hits no other triangle of the object that was hit
//   sample_rate : n x n rays per pixel on a regular sub-pixel grid, per-channel integer mean; id of sample (n/2, n/2)
// Every decision value is formed with the same explicitly rounded fp32 operations as in the hot kernel (same helpers).
//
// Shape: one thread per pixel, a warp = an 8x4 pixel block, classic while-while traversal with a per-thread stack.  This
// path is the general one, not the headline one: the persistent kernel's lane refill, work stealing and frame batching
// are not repeated here.
#pragma once
#include "rtb_render.cuh"

namespace rtb {

constexpr int kMaxSceneObjects = 8;
constexpr int kMaxSceneLights = 8;
constexpr float kShadowTMin = 1e-4f;

struct SceneObject {
    const float4* __restrict__ nodes;
    const float4* __restrict__ tris;
    const float4* __restrict__ rad;  // may be null
    float uniform_rad[3];
    float root_box[6];
    int root_ref;
    int id_base;   // added to the object's triangle ids
    float m[12];   // the object's matrix for this frame: rows x,y,z = (i,j,k,w)
};

struct SceneParams {
    int W, H;
    float n_mod[3], u_mod[3], v_mod[3];
    float draw_distance;
    uint32_t background;
    int num_objects;
    SceneObject obj[kMaxSceneObjects];
    int num_lights;
    float lights[kMaxSceneLights][3];
    int shadows;
    int samples;  // n: n x n rays per pixel (>= 1)
    int cull;
    float cull_rel;
    uint32_t* __restrict__ out_bgra;
    int32_t* __restrict__ out_ids;
};

// One object, closest hit: the visit sequence of Trixel.cu:70-170 (same decisions, same order as render_stream_kernel),
// `best` carried in and out.  Returns the triangle of a closer hit or -1.
__device__ __forceinline__ int trace_closest(const SceneParams& P, const SceneObject& O, const Ray& r, float& best) {
    int stk_ref[kStackDepth];
    float stk_tmin[kStackDepth], stk_tmax[kStackDepth];
    int sp = 0, id = -1;
    int cur = O.root_ref;
    float cur_tmin = 0.0f, cur_tmax = 0.0f;
    float slack_abs = 0.0f;
    if (P.cull) {
        const float bx = fmaxf(fabsf(O.root_box[0]), fabsf(O.root_box[3])), by = fmaxf(fabsf(O.root_box[1]), fabsf(O.root_box[4]));
        const float bz = fmaxf(fabsf(O.root_box[2]), fabsf(O.root_box[5]));
        const float e = fmaxf(fmaxf(bx * fabsf(r.ix) + fabsf(r.fx), by * fabsf(r.iy) + fabsf(r.fy)), bz * fabsf(r.iz) + fabsf(r.fz));
        slack_abs = e * 9.5367431640625e-07f;  // 8 * 2^-23, as in the hot kernel
    }
    auto culled = [&](float tmin) -> bool {
        return P.cull && (tmin > __fmaf_rn(P.cull_rel, fabsf(tmin), best + (slack_abs + P.cull_rel * fabsf(best))));
    };
    if (O.root_ref >= 0) {
        slab(r, O.root_box[0], O.root_box[1], O.root_box[2], O.root_box[3], O.root_box[4], O.root_box[5], cur_tmin, cur_tmax);
        if (!box_entered_exact(cur_tmin, cur_tmax)) return -1;
    }
    for (;;) {
        if (cur < 0) {  // leaf: always intersected when popped (Trixel.cu:98)
            moller_trumbore(r, O.tris, (int)((unsigned)cur & kRefIndexMask), best, id);
        } else {
            const float4* rec = O.nodes + 4ll * (long long)((unsigned)cur & kRefIndexMask);
            const float4 q0 = ldg4(rec), q1 = ldg4(rec + 1), q2 = ldg4(rec + 2), q3 = ldg4(rec + 3);
            const int lref = __float_as_int(q3.x), rref = __float_as_int(q3.y);
            const float S1 = q3.z, S2 = q3.w;
            const int axis = (lref >> kRefAxisShift) & 3;
            const float dir = axis == 0 ? r.dx : (axis == 1 ? r.dy : r.dz), ds = axis == 0 ? r.ox : (axis == 1 ? r.oy : r.oz);
            const float a = __fmul_rn(cur_tmin, dir), b = __fmul_rn(cur_tmax, dir);
            const float s2 = __fadd_rn(S2, ds);
            float ltmin, ltmax, rtmin, rtmax;
            slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ltmin, ltmax);
            slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, rtmin, rtmax);
            bool decidable, left_first, visit_second, l_in, r_in;
            interior_decisions_fast(a, b, s2, S1, ds, ltmin, ltmax, rtmin, rtmax, decidable, left_first, visit_second, l_in, r_in);
            if (!decidable) {
                const int ex = interior_decisions_exact(a, b, s2, S1, ds, ltmin, ltmax, rtmin, rtmax);
                left_first = (ex & 1) != 0; visit_second = (ex & 2) != 0; l_in = (ex & 4) != 0; r_in = (ex & 8) != 0;
            }
            const int first = left_first ? lref : rref, second = left_first ? rref : lref;
            const float f_tmin = left_first ? ltmin : rtmin, f_tmax = left_first ? ltmax : rtmax;
            const float s_tmin = left_first ? rtmin : ltmin, s_tmax = left_first ? rtmax : ltmax;
            const bool f_in = left_first ? l_in : r_in, s_in = left_first ? r_in : l_in;
            const bool go_first = ((first < 0) | f_in) & !culled(f_tmin);
            const bool go_second = visit_second & ((second < 0) | s_in) & !culled(s_tmin);
            if (go_second) { stk_ref[sp] = second; stk_tmin[sp] = s_tmin; stk_tmax[sp] = s_tmax; sp++; }
            if (go_first) { cur = first; cur_tmin = f_tmin; cur_tmax = f_tmax; continue; }
        }
        // pop the next entry that can still hold a closer hit
        bool found = false;
        while (sp > 0) {
            sp--;
            if (!culled(stk_tmin[sp])) { cur = stk_ref[sp]; cur_tmin = stk_tmin[sp]; cur_tmax = stk_tmax[sp]; found = true; break; }
        }
        if (!found) break;
    }
    return id;
}

// Any-hit over one object: is the segment of `r` (r.o = MINUS its origin, as for a primary ray) blocked at a parameter
// kShadowTMin < t < 1 by a triangle other than `skip`?  A child box is entered iff tmax >= tmin, tmax >= 0, tmin <= 1
// (plain fp32 compares, false for NaN); leaf children are tested when their parent is entered.
__device__ __forceinline__ bool segment_blocked(const SceneObject& O, const Ray& r, int skip) {
    int stk[kStackDepth];
    int sp = 0;
    int cur = O.root_ref;
    for (;;) {
        if (cur < 0) {
            const int tri = (int)((unsigned)cur & kRefIndexMask);
            if (tri != skip) {
                const float4 a = ldg4(O.tris + 3ll * tri), b = ldg4(O.tris + 3ll * tri + 1), c = ldg4(O.tris + 3ll * tri + 2);
                const float px = __fsub_rn(__fmul_rn(r.dy, b.z), __fmul_rn(r.dz, b.y));
                const float py = __fsub_rn(__fmul_rn(r.dz, b.x), __fmul_rn(r.dx, b.z));
                const float pz = __fsub_rn(__fmul_rn(r.dx, b.y), __fmul_rn(r.dy, b.x));
                const float f = dot3(px, py, pz, a.x, a.y, a.z);
                if (!(f < RTB_EPS_UP && f > -RTB_EPS_UP)) {
                    const float pe1 = __frcp_rn(f);
                    const float tx = __fsub_rn(c.x, r.ox), ty = __fsub_rn(c.y, r.oy), tz = __fsub_rn(c.z, r.oz);
                    const float u = __fmul_rn(pe1, dot3(px, py, pz, tx, ty, tz));
                    const float qx = __fsub_rn(__fmul_rn(ty, a.z), __fmul_rn(tz, a.y));
                    const float qy = __fsub_rn(__fmul_rn(tz, a.x), __fmul_rn(tx, a.z));
                    const float qz = __fsub_rn(__fmul_rn(tx, a.y), __fmul_rn(ty, a.x));
                    const float v = __fmul_rn(pe1, dot3(r.dx, r.dy, r.dz, qx, qy, qz));
                    const float w = __fmul_rn(pe1, dot3(b.x, b.y, b.z, qx, qy, qz));
                    const bool reject = (u < RTB_EPS_UP) || (v < RTB_EPS_UP) || (__fadd_rn(u, v) > 1.0f) || (w < RTB_EPS_UP);
                    if (!reject && w > kShadowTMin && w < 1.0f) return true;
                }
            }
        } else {
            const float4* rec = O.nodes + 4ll * (long long)((unsigned)cur & kRefIndexMask);
            const float4 q0 = ldg4(rec), q1 = ldg4(rec + 1), q2 = ldg4(rec + 2), q3 = ldg4(rec + 3);
            const int lref = __float_as_int(q3.x) & ~(3 << kRefAxisShift), rref = __float_as_int(q3.y);
            float ltmin, ltmax, rtmin, rtmax;
            slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ltmin, ltmax);
            slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, rtmin, rtmax);
            const bool go_l = (lref < 0) || (ltmax >= ltmin && ltmax >= 0.0f && ltmin <= 1.0f);
            const bool go_r = (rref < 0) || (rtmax >= rtmin && rtmax >= 0.0f && rtmin <= 1.0f);
            if (go_l && go_r) { stk[sp++] = rref; cur = lref; continue; }
            if (go_l) { cur = lref; continue; }
            if (go_r) { cur = rref; continue; }
        }
        if (sp == 0) return false;
        cur = stk[--sp];
    }
}

// Trixel.cu:60-66 + the slab constants of Trixel.cu:94-95 for one object's matrix
__device__ __forceinline__ void object_ray(const SceneObject& O, float cmx, float cmy, float cmz, Ray& r) {
    const float* m = O.m;
    r.ox = m[3]; r.oy = m[7]; r.oz = m[11];
    r.dx = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m[0], -cmx), __fmul_rn(m[1], -cmy)), __fmul_rn(m[2], -cmz)));
    r.dy = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m[4], -cmx), __fmul_rn(m[5], -cmy)), __fmul_rn(m[6], -cmz)));
    r.dz = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m[8], -cmx), __fmul_rn(m[9], -cmy)), __fmul_rn(m[10], -cmz)));
    r.ix = __frcp_rn(r.dx); r.iy = __frcp_rn(r.dy); r.iz = __frcp_rn(r.dz);
    r.fx = __fdiv_rn(r.ox, r.dx); r.fy = __fdiv_rn(r.oy, r.dy); r.fz = __fdiv_rn(r.oz, r.dz);
}

// color_cam_cuda with the light loop and the shadow test restored (Camera.cu:27-61); returns 0x00RRGGBB
__device__ __forceinline__ uint32_t shade_lights(const SceneParams& P, const SceneObject& O, const Ray& r, float best, int tri, float cmx, float cmy,
                                                 float cmz) {
    const float* m = O.m;
    const float pntx = __fadd_rn(__fmul_rn(best, r.dx), r.ox), pnty = __fadd_rn(__fmul_rn(best, r.dy), r.oy), pntz = __fadd_rn(__fmul_rn(best, r.dz), r.oz);
    const float n0 = ldg4(O.tris + 3ll * tri).w, n1 = ldg4(O.tris + 3ll * tri + 1).w, n2 = ldg4(O.tris + 3ll * tri + 2).w;
    const float ax = __fmul_rn(-1.0f, n0), ay = __fmul_rn(-1.0f, n1), az = __fmul_rn(-1.0f, n2);
    float nx = __fadd_rn(__fadd_rn(__fmul_rn(ax, m[0]), __fmul_rn(ay, m[1])), __fmul_rn(az, m[2]));
    float ny = __fadd_rn(__fadd_rn(__fmul_rn(ax, m[4]), __fmul_rn(ay, m[5])), __fmul_rn(az, m[6]));
    float nz = __fadd_rn(__fadd_rn(__fmul_rn(ax, m[8]), __fmul_rn(ay, m[9])), __fmul_rn(az, m[10]));
    nx = __fmul_rn(nx, -1.0f); ny = __fmul_rn(ny, -1.0f); nz = __fmul_rn(nz, -1.0f);
    float cr, cg, cb;
    if (O.rad) { const float4 c = ldg4(O.rad + tri); cr = c.x; cg = c.y; cb = c.z; }
    else { cr = O.uniform_rad[0]; cg = O.uniform_rad[1]; cb = O.uniform_rad[2]; }
    float pr = 0.0f, pg = 0.0f, pb = 0.0f;
    for (int l = 0; l < P.num_lights; l++) {
        float sx = __fsub_rn(P.lights[l][0], pntx), sy = __fsub_rn(P.lights[l][1], pnty), sz = __fsub_rn(P.lights[l][2], pntz);
        if (P.shadows) {
            Ray s;  // from the hit point X = w*d - od towards the light: r.o = -X = od - w*d
            s.dx = sx; s.dy = sy; s.dz = sz;
            s.ox = __fsub_rn(r.ox, __fmul_rn(best, r.dx)); s.oy = __fsub_rn(r.oy, __fmul_rn(best, r.dy)); s.oz = __fsub_rn(r.oz, __fmul_rn(best, r.dz));
            s.ix = __frcp_rn(s.dx); s.iy = __frcp_rn(s.dy); s.iz = __frcp_rn(s.dz);
            s.fx = __fdiv_rn(s.ox, s.dx); s.fy = __fdiv_rn(s.oy, s.dy); s.fz = __fdiv_rn(s.oz, s.dz);
            if (segment_blocked(O, s, tri)) continue;
        }
        normalize21(sx, sy, sz);
        const float k = dot3(sx, sy, sz, nx, nx, nz);  // norm.x twice, Camera.cu:38
        const float k2 = __fmul_rn(2.0f, k);
        const float rx = __fmul_rn(__fsub_rn(sx, __fmul_rn(k2, nx)), cmx);
        const float ry = __fmul_rn(__fsub_rn(sy, __fmul_rn(k2, ny)), cmy);
        const float rz = __fmul_rn(__fsub_rn(sz, __fmul_rn(k2, nz)), cmz);
        const float dif = __double2float_rn(__dmul_rn(.6, (double)fabsf(k)));
        const double x = (double)fabsf(__fadd_rn(__fadd_rn(rx, ry), rz));
        const double x2 = __dmul_rn(x, x);
        const float p5 = __double2float_rn(__dmul_rn(__dmul_rn(x2, x2), x));
        const float spc = __double2float_rn(__dmul_rn((double)p5, .3));
        pr = __fadd_rn(pr, __fadd_rn(__fmul_rn(cr, dif), spc));
        pg = __fadd_rn(pg, __fadd_rn(__fmul_rn(cg, dif), spc));
        pb = __fadd_rn(pb, __fadd_rn(__fmul_rn(cb, dif), spc));
    }
    const float mx = fmaxf(fmaxf(pr, pg), pb);
    const uint32_t r8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pr, mx), 255.0f)) & 0xffu;
    const uint32_t g8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pg, mx), 255.0f)) & 0xffu;
    const uint32_t b8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pb, mx), 255.0f)) & 0xffu;
    return (r8 << 16) | (g8 << 8) | b8;
}

// grid = (ceil(W/16), ceil(H/8)), block = 128: warp w of a block owns the 8x4 pixel block (2*bx + (w & 1), 2*by + (w >> 1))
__global__ void __launch_bounds__(kBlockThreads) render_scene_kernel(const SceneParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = (blockIdx.x * 2 + (warp & 1)) * 8 + (lane & 7), py = (blockIdx.y * 2 + (warp >> 1)) * 4 + (lane >> 3);
    if (px >= P.W || py >= P.H) return;
    const int n = P.samples;
    uint32_t sum_r = 0, sum_g = 0, sum_b = 0;
    int centre_id = -1;
    for (int sb = 0; sb < n; sb++) {
        for (int sa = 0; sa < n; sa++) {
            // primary ray, Camera.cu:103-104; sub-pixel offsets (a + .5)/n - .5 for n >= 2
            float fxp = (float)px, fyp = (float)py;
            if (n > 1) {
                fxp = __fadd_rn(fxp, __fsub_rn(__fdiv_rn(__fadd_rn((float)sa, 0.5f), (float)n), 0.5f));
                fyp = __fadd_rn(fyp, __fsub_rn(__fdiv_rn(__fadd_rn((float)sb, 0.5f), (float)n), 0.5f));
            }
            float cmx = __fadd_rn(__fadd_rn(P.n_mod[0], __fmul_rn(P.u_mod[0], fxp)), __fmul_rn(P.v_mod[0], fyp));
            float cmy = __fadd_rn(__fadd_rn(P.n_mod[1], __fmul_rn(P.u_mod[1], fxp)), __fmul_rn(P.v_mod[1], fyp));
            float cmz = __fadd_rn(__fadd_rn(P.n_mod[2], __fmul_rn(P.u_mod[2], fxp)), __fmul_rn(P.v_mod[2], fyp));
            normalize21(cmx, cmy, cmz);
            float best = P.draw_distance;
            int hit_obj = -1, hit_tri = -1;
            for (int k = 0; k < P.num_objects; k++) {
                Ray r;
                object_ray(P.obj[k], cmx, cmy, cmz, r);
                const int t = trace_closest(P, P.obj[k], r, best);
                if (t >= 0) { hit_obj = k; hit_tri = t; }
            }
            uint32_t c = P.background;
            if (hit_obj >= 0) {
                Ray r;
                object_ray(P.obj[hit_obj], cmx, cmy, cmz, r);
                c = shade_lights(P, P.obj[hit_obj], r, best, hit_tri, cmx, cmy, cmz);
            }
            sum_r += (c >> 16) & 0xffu; sum_g += (c >> 8) & 0xffu; sum_b += c & 0xffu;
            if (sa == n / 2 && sb == n / 2) centre_id = hit_obj >= 0 ? P.obj[hit_obj].id_base + hit_tri : -1;
        }
    }
    const long long o = (long long)py * P.W + px;
    const uint32_t nn = (uint32_t)(n * n);
    if (P.out_bgra) P.out_bgra[o] = ((sum_r / nn) << 16) | ((sum_g / nn) << 8) | (sum_b / nn) | (P.background & 0xff000000u);
    if (P.out_ids) P.out_ids[o] = centre_id;
}

}  // namespace rtb
