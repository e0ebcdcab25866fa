// rtb_kernels.cuh -- device side of the ray-cast path for sm_100a.
//
// Layout in HBM (all camera-relative, built once per (camera, mesh) by the pack kernels below;
// SURVEY.md Appendix E explains why the reference's 6+9 scattered arrays are replaced):
//
//   nodes : one 64-byte record (4 x float4, 64-byte aligned = two 32-byte sectors) per INTERIOR
//           node, holding the boxes of BOTH children; leaves have no record, a child reference
//           < 0 is ~triangle.  The reference's per-node split planes are redundant with the child
//           boxes (s1 == left child's max, s2 == right child's min on the split axis,
//           Trixel.h:353-376) and are read from them.
//             q0 = L.t0x L.t0y L.t0z L.t1x      q1 = L.t1y L.t1z R.t0x R.t0y
//             q2 = R.t0z R.t1x R.t1y R.t1z      q3 = left_ref right_ref axis pad   (as int bits)
//   tris  : one 48-byte record (3 x float4) per triangle
//             t0 = e1.xyz n.x    t1 = e2.xyz n.y    t2 = (cam_pos - p1).xyz n.z
//   rad   : float4 per triangle (r,g,b,-) or absent when the mesh has one colour.
//
// Numerical contract: every value that feeds a DECISION (slab tests, plane compares,
// Moller-Trumbore u/v/w) is computed with explicitly rounded fp32 operations in the reference's
// association order -- the translation unit is compiled with -fmad=false and the code below uses
// __fmul_rn/__fadd_rn/__fsub_rn so nothing can be contracted into an FMA.  Comparisons that the
// reference performs in double precision (because its epsilons are double literals,
// vector.cuh:10-11) are reproduced exactly; see the cmp_* helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

constexpr int kTile = 32;        // multi-GPU interleave granularity: 32x32 pixel tiles
constexpr int kWarpTileW = 8;    // one warp renders 8x4 pixels
constexpr int kWarpTileH = 4;
constexpr int kItemsPerTile = (kTile / kWarpTileW) * (kTile / kWarpTileH);  // 16
constexpr int kStackDepth = 40;  // >= tree height + 2 (median split: height = ceil(log2 n))
constexpr int kBlockThreads = 128;

struct RenderParams {
    int W, H;
    float n_mod[3], u_mod[3], v_mod[3];
    float root_box[6];  // t0x t0y t0z t1x t1y t1z of node 0, camera-relative
    float draw_distance;
    uint32_t background;  // 0x00RRGGBB
    int root_ref;         // 0 = interior root record 0; < 0: ~triangle (single-triangle mesh)
    const float4* __restrict__ nodes;
    const float4* __restrict__ tris;
    const float4* __restrict__ rad;  // may be null
    float uniform_rad[3];
    const float* __restrict__ frames;  // 12 floats per frame: rows x,y,z = (i,j,k,w)
    int num_frames;
    int tiles_x;           // tiles per image row
    int tile_first, tile_stride;
    int my_tiles;          // number of 32x32 tiles of one frame rendered by this launch
    int chunk;             // consecutive warp items taken per work fetch
    long long total_items; // num_frames * my_tiles * kItemsPerTile
    uint32_t* __restrict__ out_bgra;
    int32_t* __restrict__ out_ids;
    unsigned long long* work_counter;
    unsigned long long* counters;  // [0] rays [1] interior nodes entered [2] nodes popped (reference sense) [3] triangle tests [4] hits
    float cull_rel;
};

// ---------------------------------------------------------------------------------------------
// exact comparison helpers.  EPS = 1e-16 (double).  kEpsUp is the smallest float >= 1e-16, so for a
// float x:  (double)x < 1e-16  <=>  x < kEpsUp   and   (double)x > -1e-16  <=>  x > -kEpsUp.
// ---------------------------------------------------------------------------------------------
#define RTB_EPS_UP __int_as_float(0x24e69595)
#define RTB_TINY 3.7252902984619140625e-09f /* 2^-28: above this, neighbouring floats are > 1e-16 apart */

// The double-precision forms are only reachable for |x| < 2^-28, NaN or exact ties; they live in
// separate functions so the common path stays branch-light fp32.
__device__ __noinline__ bool slow_ge_minus_eps(float hi, float lo) { return (double)hi >= (double)lo - 1e-16; }
__device__ __noinline__ bool slow_lt_plus_eps(float a, float s) { return (double)a < (double)s + 1e-16; }
__device__ __noinline__ bool slow_gt_minus_eps(float b, float s) { return (double)b > (double)s - 1e-16; }

// (double)hi >= (double)lo - 1e-16          (Trixel.cu:146, first clause)
// hi >= lo implies it; hi < lo with |lo| >= 2^-28 refutes it (lo - 1e-16 rounds above prev(lo)).
__device__ __forceinline__ bool cmp_ge_minus_eps(float hi, float lo) {
    bool r = hi >= lo;
    if (!r && !(fabsf(lo) >= RTB_TINY)) r = slow_ge_minus_eps(hi, lo);
    return r;
}
// (double)a < (double)s + 1e-16             (Trixel.cu:155)
__device__ __forceinline__ bool cmp_lt_plus_eps(float a, float s) {
    bool r = a < s;
    if (!r && !(a > s && fabsf(s) >= RTB_TINY)) r = slow_lt_plus_eps(a, s);
    return r;
}
// (double)b > (double)s - 1e-16             (Trixel.cu:156)
__device__ __forceinline__ bool cmp_gt_minus_eps(float b, float s) {
    bool r = b > s;
    if (!r && !(b < s && fabsf(s) >= RTB_TINY)) r = slow_gt_minus_eps(b, s);
    return r;
}
// (float)(((double)S1 + 1e-16) + (double)ds)   (Trixel.cu:150: `s1 = cvm->s1[cni] + EPS + ds`)
__device__ __forceinline__ float s1_plus_eps_plus_ds(float S1, float ds) {
    return __double2float_rn(__dadd_rn(__dadd_rn((double)S1, 1e-16), (double)ds));
}

// vector.cuh:79-95 + 117-120: Quake start value, 21 Newton steps, then scale
__device__ __forceinline__ void normalize21(float& x, float& y, float& z) {
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const float half = __fmul_rn(0.5f, s);
    float g = __int_as_float(0x5f375a86 - (__float_as_int(half) >> 1));
#pragma unroll
    for (int k = 0; k < 21; k++) g = __fmul_rn(g, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(half, g), g)));
    x = __fmul_rn(x, g); y = __fmul_rn(y, g); z = __fmul_rn(z, g);
}
// vector.cuh:122-124
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

struct Ray {
    float dx, dy, dz;     // object-space direction (Trixel.cu:64-66)
    float ix, iy, iz;     // 1 / d
    float fx, fy, fz;     // od / d  (Trixel.cu:94-95)
    float ox, oy, oz;     // od = translation column (Trixel.cu:60-62)
};

// slab entry/exit of a camera-relative box for this ray, Trixel.cu:76-95
__device__ __forceinline__ void slab(const Ray& r, float b0x, float b0y, float b0z, float b1x, float b1y, float b1z,
                                     float& tmin, float& tmax) {
    const bool px = r.dx > 0.0f, py = r.dy > 0.0f, pz = r.dz > 0.0f;
    const float t0x = __fadd_rn(__fmul_rn(px ? b0x : b1x, r.ix), r.fx);
    const float t1x = __fadd_rn(__fmul_rn(px ? b1x : b0x, r.ix), r.fx);
    const float t0y = __fadd_rn(__fmul_rn(py ? b0y : b1y, r.iy), r.fy);
    const float t1y = __fadd_rn(__fmul_rn(py ? b1y : b0y, r.iy), r.fy);
    const float t0z = __fadd_rn(__fmul_rn(pz ? b0z : b1z, r.iz), r.fz);
    const float t1z = __fadd_rn(__fmul_rn(pz ? b1z : b0z, r.iz), r.fz);
    tmin = fmaxf(t0z, fmaxf(t0x, t0y));
    tmax = fminf(t1z, fminf(t1x, t1y));
}
// Trixel.cu:146: enter iff tmax >= tmin - EPS && tmin > -EPS (both in double)
__device__ __forceinline__ bool box_entered(float tmin, float tmax) {
    return cmp_ge_minus_eps(tmax, tmin) && (tmin > -RTB_EPS_UP);
}

// ---------------------------------------------------------------------------------------------
// pack kernels (the reference's init kernels)
// ---------------------------------------------------------------------------------------------

// Trixel.cu:11-27 init_tri_mem_cuda (e1, e2, unit normal) + Trixel.cu:29-36 init_cam_tri_mem_cuda
// (T = cam_pos - p1), one 48-byte record per triangle.
__global__ void pack_triangles_kernel(const float* __restrict__ points9, long long n, float cx, float cy, float cz,
                                      float4* __restrict__ tris) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = points9 + 9 * i;
    const float p0 = p[0], p1 = p[1], p2 = p[2];
    const float e1x = __fsub_rn(p[3], p0), e1y = __fsub_rn(p[4], p1), e1z = __fsub_rn(p[5], p2);
    const float e2x = __fsub_rn(p[6], p0), e2y = __fsub_rn(p[7], p1), e2z = __fsub_rn(p[8], p2);
    float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
    float ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
    float nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    normalize21(nx, ny, nz);
    tris[3 * i + 0] = make_float4(e1x, e1y, e1z, nx);
    tris[3 * i + 1] = make_float4(e2x, e2y, e2z, ny);
    tris[3 * i + 2] = make_float4(__fsub_rn(cx, p0), __fsub_rn(cy, p1), __fsub_rn(cz, p2), nz);
}

// Camera.cu:137-162 init_cam_voxel_mem_cuda: camera-relative boxes `bound - cam + obj_center`
// (obj_center == 0, Camera.cpp:167-170).  One thread per INTERIOR node; record_of[node] is its
// record index (-1 for leaves).
__global__ void pack_nodes_kernel(const float* __restrict__ bounds6, const int* __restrict__ left,
                                  const int* __restrict__ tri, const unsigned char* __restrict__ cut_flag,
                                  const int* __restrict__ record_of, long long num_nodes, float cx, float cy, float cz,
                                  float4* __restrict__ nodes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int rec = record_of[i];
    if (rec < 0) return;
    const int l = left[i], r = l + 1;
    const float* bl = bounds6 + 6ll * l;  // x0,x1,y0,y1,z0,z1
    const float* br = bounds6 + 6ll * r;
    float L[6], R[6];  // t0x t0y t0z t1x t1y t1z
    L[0] = __fadd_rn(__fsub_rn(bl[0], cx), 0.0f); L[3] = __fadd_rn(__fsub_rn(bl[1], cx), 0.0f);
    L[1] = __fadd_rn(__fsub_rn(bl[2], cy), 0.0f); L[4] = __fadd_rn(__fsub_rn(bl[3], cy), 0.0f);
    L[2] = __fadd_rn(__fsub_rn(bl[4], cz), 0.0f); L[5] = __fadd_rn(__fsub_rn(bl[5], cz), 0.0f);
    R[0] = __fadd_rn(__fsub_rn(br[0], cx), 0.0f); R[3] = __fadd_rn(__fsub_rn(br[1], cx), 0.0f);
    R[1] = __fadd_rn(__fsub_rn(br[2], cy), 0.0f); R[4] = __fadd_rn(__fsub_rn(br[3], cy), 0.0f);
    R[2] = __fadd_rn(__fsub_rn(br[4], cz), 0.0f); R[5] = __fadd_rn(__fsub_rn(br[5], cz), 0.0f);
    const int lref = record_of[l] >= 0 ? record_of[l] : ~tri[l];
    const int rref = record_of[r] >= 0 ? record_of[r] : ~tri[r];
    const int axis = cut_flag[i] % 3;
    float4* o = nodes + 4ll * rec;
    o[0] = make_float4(L[0], L[1], L[2], L[3]);
    o[1] = make_float4(L[4], L[5], R[0], R[1]);
    o[2] = make_float4(R[2], R[3], R[4], R[5]);
    o[3] = make_float4(__int_as_float(lref), __int_as_float(rref), __int_as_float(axis), 0.0f);
}

__global__ void fill_kernel(uint32_t* __restrict__ out, long long n, uint32_t value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = value;
}
__global__ void fill_ids_kernel(int32_t* __restrict__ out, long long n, int32_t value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = value;
}

// ---------------------------------------------------------------------------------------------
// the hot kernel
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Moller-Trumbore, Trixel.cu:98-145.  Returns true and updates best/id on acceptance.
__device__ __forceinline__ bool moller_trumbore(const Ray& r, const float4* __restrict__ tris, int tri, float& best, int& id) {
    const float4 a = ldg4(tris + 3ll * tri), b = ldg4(tris + 3ll * tri + 1), c = ldg4(tris + 3ll * tri + 2);
    // p = d x e2
    const float px = __fsub_rn(__fmul_rn(r.dy, b.z), __fmul_rn(r.dz, b.y));
    const float py = __fsub_rn(__fmul_rn(r.dz, b.x), __fmul_rn(r.dx, b.z));
    const float pz = __fsub_rn(__fmul_rn(r.dx, b.y), __fmul_rn(r.dy, b.x));
    const float f = dot3(px, py, pz, a.x, a.y, a.z);
    if (f < RTB_EPS_UP && f > -RTB_EPS_UP) return false;  // (double)f < EPS && (double)f > -EPS
    // pe1 = (float)(1.0 / (double)f): the double quotient rounded to float equals the correctly
    // rounded float reciprocal for every finite f (no float midpoint is within 2^-53 of 1/f)
    const float pe1 = __frcp_rn(f);
    const float tx = __fsub_rn(c.x, r.ox), ty = __fsub_rn(c.y, r.oy), tz = __fsub_rn(c.z, r.oz);
    const float u = __fmul_rn(pe1, dot3(px, py, pz, tx, ty, tz));
    // q = T x e1
    const float qx = __fsub_rn(__fmul_rn(ty, a.z), __fmul_rn(tz, a.y));
    const float qy = __fsub_rn(__fmul_rn(tz, a.x), __fmul_rn(tx, a.z));
    const float qz = __fsub_rn(__fmul_rn(tx, a.y), __fmul_rn(ty, a.x));
    const float v = __fmul_rn(pe1, dot3(r.dx, r.dy, r.dz, qx, qy, qz));
    const float w = __fmul_rn(pe1, dot3(b.x, b.y, b.z, qx, qy, qz));
    // (w < d) && !(u < EPS || v < EPS || (u+v) > 1+EPS || w < EPS); 1+1e-16 == 1.0 in double
    const bool reject = (u < RTB_EPS_UP) || (v < RTB_EPS_UP) || (__fadd_rn(u, v) > 1.0f) || (w < RTB_EPS_UP);
    if ((w < best) && !reject) { best = w; id = tri; return true; }
    return false;
}

// Camera.cu:19-69 color_cam_cuda for a hit pixel; returns 0x00RRGGBB.
__device__ __forceinline__ uint32_t phong(const RenderParams& P, const float* __restrict__ M, const Ray& r, float best, int id,
                                          float cmx, float cmy, float cmz) {
    // hit record as written at Trixel.cu:134-140
    const float pntx = __fadd_rn(__fmul_rn(best, r.dx), r.ox);
    const float pnty = __fadd_rn(__fmul_rn(best, r.dy), r.oy);
    const float pntz = __fadd_rn(__fmul_rn(best, r.dz), r.oz);
    const float n0 = ldg4(P.tris + 3ll * id).w, n1 = ldg4(P.tris + 3ll * id + 1).w, n2 = ldg4(P.tris + 3ll * id + 2).w;
    // VEC3_CUDA::device_rotate(rot_m, i, -1), vector.cuh:25-33
    const float ax = __fmul_rn(-1.0f, n0), ay = __fmul_rn(-1.0f, n1), az = __fmul_rn(-1.0f, n2);
    float nx = __fadd_rn(__fadd_rn(__fmul_rn(ax, M[0]), __fmul_rn(ay, M[1])), __fmul_rn(az, M[2]));
    float ny = __fadd_rn(__fadd_rn(__fmul_rn(ax, M[4]), __fmul_rn(ay, M[5])), __fmul_rn(az, M[6]));
    float nz = __fadd_rn(__fadd_rn(__fmul_rn(ax, M[8]), __fmul_rn(ay, M[9])), __fmul_rn(az, M[10]));
    nx = __fmul_rn(nx, -1.0f); ny = __fmul_rn(ny, -1.0f); nz = __fmul_rn(nz, -1.0f);
    float cr, cg, cb;
    if (P.rad) { const float4 c = ldg4(P.rad + id); cr = c.x; cg = c.y; cb = c.z; }
    else { cr = P.uniform_rad[0]; cg = P.uniform_rad[1]; cb = P.uniform_rad[2]; }
    // light at (2,2,2), Camera.cu:32
    float sx = __fsub_rn(2.0f, pntx), sy = __fsub_rn(2.0f, pnty), sz = __fsub_rn(2.0f, pntz);
    normalize21(sx, sy, sz);
    const float k = dot3(sx, sy, sz, nx, nx, nz);  // norm.x twice, Camera.cu:38
    const float k2 = __fmul_rn(2.0f, k);
    const float rx = __fmul_rn(__fsub_rn(sx, __fmul_rn(k2, nx)), cmx);
    const float ry = __fmul_rn(__fsub_rn(sy, __fmul_rn(k2, ny)), cmy);
    const float rz = __fmul_rn(__fsub_rn(sz, __fmul_rn(k2, nz)), cmz);
    const float dif = __double2float_rn(__dmul_rn(.6, (double)fabsf(k)));  // Camera.cu:44
    // powf(x, 5) * .3 (Camera.cu:45): x^5 through three double products, rounded once to float --
    // the correctly rounded value, which is what glibc's powf returns (<= 1 ulp otherwise)
    const double x = (double)fabsf(__fadd_rn(__fadd_rn(rx, ry), rz));
    const double x2 = __dmul_rn(x, x);
    const float p5 = __double2float_rn(__dmul_rn(__dmul_rn(x2, x2), x));
    const float spc = __double2float_rn(__dmul_rn((double)p5, .3));
    const float pr = __fadd_rn(0.0f, __fadd_rn(__fmul_rn(cr, dif), spc));
    const float pg = __fadd_rn(0.0f, __fadd_rn(__fmul_rn(cg, dif), spc));
    const float pb = __fadd_rn(0.0f, __fadd_rn(__fmul_rn(cb, dif), spc));
    const float mx = fmaxf(fmaxf(pr, pg), pb);
    // (u8)((c / max) * 255): truncation; cvt.rzi.u8 saturates and maps NaN to 0
    const uint32_t r8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pr, mx), 255.0f)) & 0xffu;
    const uint32_t g8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pg, mx), 255.0f)) & 0xffu;
    const uint32_t b8 = (uint32_t)__float2uint_rz(__fmul_rn(__fdiv_rn(pb, mx), 255.0f)) & 0xffu;
    return (r8 << 16) | (g8 << 8) | b8;
}

template <bool CULL, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads) render_kernel(const RenderParams P) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long c_nodes = 0, c_boxes = 0, c_tris = 0, c_rays = 0, c_hits = 0;
    const long long items_per_frame = (long long)P.my_tiles * kItemsPerTile;

    int stk_ref[kStackDepth];
    float stk_tmin[kStackDepth], stk_tmax[kStackDepth];

    for (;;) {
        // ---- warp-level work fetch: `chunk` consecutive 8x4 warp items per atomic -------------
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(P.work_counter, (unsigned long long)P.chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= P.total_items) break;
        const long long stop = base + P.chunk < P.total_items ? base + P.chunk : P.total_items;
        for (long long item = base; item < stop; item++) {
            const int frame = (int)(item / items_per_frame);
            const int rem = (int)(item - (long long)frame * items_per_frame);
            const int tile = P.tile_first + (rem / kItemsPerTile) * P.tile_stride;
            const int sub = rem % kItemsPerTile;
            const int px = (tile % P.tiles_x) * kTile + (sub % (kTile / kWarpTileW)) * kWarpTileW + (int)(lane % kWarpTileW);
            const int py = (tile / P.tiles_x) * kTile + (sub / (kTile / kWarpTileW)) * kWarpTileH + (int)(lane / kWarpTileW);
            if (px >= P.W || py >= P.H) continue;
            const float* __restrict__ M = P.frames + 12ll * frame;
            const long long pix = (long long)py * P.W + px;

            // ---- primary ray, Camera.cu:103-104 (row 0 = bottom) ------------------------------
            const float fxp = (float)px, fyp = (float)py;
            float cmx = __fadd_rn(__fadd_rn(P.n_mod[0], __fmul_rn(P.u_mod[0], fxp)), __fmul_rn(P.v_mod[0], fyp));
            float cmy = __fadd_rn(__fadd_rn(P.n_mod[1], __fmul_rn(P.u_mod[1], fxp)), __fmul_rn(P.v_mod[1], fyp));
            float cmz = __fadd_rn(__fadd_rn(P.n_mod[2], __fmul_rn(P.u_mod[2], fxp)), __fmul_rn(P.v_mod[2], fyp));
            normalize21(cmx, cmy, cmz);
            // ---- into object space, Trixel.cu:60-66 (sign dance kept for -0 fidelity) ----------
            Ray r;
            {
                const float m0 = __ldg(M + 0), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3);
                const float m4 = __ldg(M + 4), m5 = __ldg(M + 5), m6 = __ldg(M + 6), m7 = __ldg(M + 7);
                const float m8 = __ldg(M + 8), m9 = __ldg(M + 9), m10 = __ldg(M + 10), m11 = __ldg(M + 11);
                r.ox = m3; r.oy = m7; r.oz = m11;
                r.dx = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m0, -cmx), __fmul_rn(m1, -cmy)), __fmul_rn(m2, -cmz)));
                r.dy = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m4, -cmx), __fmul_rn(m5, -cmy)), __fmul_rn(m6, -cmz)));
                r.dz = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m8, -cmx), __fmul_rn(m9, -cmy)), __fmul_rn(m10, -cmz)));
            }
            r.ix = __frcp_rn(r.dx); r.iy = __frcp_rn(r.dy); r.iz = __frcp_rn(r.dz);
            r.fx = __fdiv_rn(r.ox, r.dx); r.fy = __fdiv_rn(r.oy, r.dy); r.fz = __fdiv_rn(r.oz, r.dz);

            float best = P.draw_distance;  // Trixel.cu:47
            int id = -1;
            if (COUNT) c_rays++;

            // conservative culling slack: rounding error bound of any slab value of this ray
            float slack_abs = 0.0f;
            if (CULL) {
                const float bx = fmaxf(fabsf(P.root_box[0]), fabsf(P.root_box[3]));
                const float by = fmaxf(fabsf(P.root_box[1]), fabsf(P.root_box[4]));
                const float bz = fmaxf(fabsf(P.root_box[2]), fabsf(P.root_box[5]));
                const float e = fmaxf(fmaxf(bx * fabsf(r.ix) + fabsf(r.fx), by * fabsf(r.iy) + fabsf(r.fy)), bz * fabsf(r.iz) + fabsf(r.fz));
                slack_abs = e * 9.5367431640625e-07f;  // 8 * 2^-23
            }
            auto culled = [&](float tmin) -> bool {
                // never true for NaN/inf slack; keeps every node that could hold a closer hit
                if (!CULL) return false;
                const float lim = best + (slack_abs + P.cull_rel * (fabsf(tmin) + fabsf(best)));
                return tmin > lim;
            };

            int sp = 0;
            int cur;
            float cur_tmin, cur_tmax;
            bool have = false;
            if (P.root_ref < 0) {
                // single-triangle mesh: the root is a leaf, tested unconditionally (Trixel.cu:98)
                if (COUNT) { c_tris++; c_boxes++; }
                moller_trumbore(r, P.tris, ~P.root_ref, best, id);
                cur = 0; cur_tmin = 0.0f; cur_tmax = 0.0f;
            } else {
                slab(r, P.root_box[0], P.root_box[1], P.root_box[2], P.root_box[3], P.root_box[4], P.root_box[5], cur_tmin, cur_tmax);
                if (COUNT) c_boxes++;
                have = box_entered(cur_tmin, cur_tmax);
                cur = 0;
            }

            while (have) {
                // ---- interior descent: `cur` is an interior node whose box test passed -----------
                while (cur >= 0) {
                    if (COUNT) c_nodes++;
                    const float4* rec = P.nodes + 4ll * cur;
                    const float4 q0 = ldg4(rec), q1 = ldg4(rec + 1), q2 = ldg4(rec + 2), q3 = ldg4(rec + 3);
                    const int lref = __float_as_int(q3.x), rref = __float_as_int(q3.y), axis = __float_as_int(q3.z);
                    // split-axis components, Trixel.cu:88-90 (three-term sums with 0/1 flags)
                    const float fxa = axis == 0 ? 1.0f : 0.0f, fya = axis == 1 ? 1.0f : 0.0f, fza = axis == 2 ? 1.0f : 0.0f;
                    const float dir = __fadd_rn(__fadd_rn(__fmul_rn(r.dx, fxa), __fmul_rn(r.dy, fya)), __fmul_rn(r.dz, fza));
                    const float ds = __fadd_rn(__fadd_rn(__fmul_rn(r.ox, fxa), __fmul_rn(r.oy, fya)), __fmul_rn(r.oz, fza));
                    // s1 = left child's max, s2 = right child's min on the split axis
                    const float S1 = axis == 0 ? q0.w : (axis == 1 ? q1.x : q1.y);
                    const float S2 = axis == 0 ? q1.z : (axis == 1 ? q1.w : q2.x);
                    const float a = __fmul_rn(cur_tmin, dir), b = __fmul_rn(cur_tmax, dir);  // Trixel.cu:149
                    const float s2 = __fadd_rn(S2, ds);                                      // Trixel.cu:151
                    bool visit_l, visit_r, left_first;
                    if (cmp_lt_plus_eps(a, s2)) {  // Trixel.cu:155-161
                        visit_l = true; left_first = true;
                        visit_r = cmp_gt_minus_eps(b, s2);
                    } else {  // Trixel.cu:162-168
                        const float s1 = s1_plus_eps_plus_ds(S1, ds);
                        visit_r = true; left_first = false;
                        visit_l = (b < s1) || (a < s1);
                    }
                    // children that the reference would pop: leaves are always intersected, interior
                    // nodes only after their own box test (Trixel.cu:98,146)
                    float ltmin, ltmax, rtmin, rtmax;
                    slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ltmin, ltmax);
                    slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, rtmin, rtmax);
                    if (COUNT) c_boxes += (int)visit_l + (int)visit_r;  // children the reference would pop
                    const bool l_in = box_entered(ltmin, ltmax), r_in = box_entered(rtmin, rtmax);
                    const bool l_cull = culled(ltmin), r_cull = culled(rtmin);
                    const bool go_l = visit_l & ((lref < 0) | l_in) & !l_cull;
                    const bool go_r = visit_r & ((rref < 0) | r_in) & !r_cull;
                    const int first = left_first ? lref : rref, second = left_first ? rref : lref;
                    const float f_tmin = left_first ? ltmin : rtmin, f_tmax = left_first ? ltmax : rtmax;
                    const float s_tmin = left_first ? rtmin : ltmin, s_tmax = left_first ? rtmax : ltmax;
                    const bool go_first = left_first ? go_l : go_r, go_second = left_first ? go_r : go_l;
                    if (go_second) { stk_ref[sp] = second; stk_tmin[sp] = s_tmin; stk_tmax[sp] = s_tmax; sp++; }
                    if (go_first) { cur = first; cur_tmin = f_tmin; cur_tmax = f_tmax; }
                    else {
                        bool got = false;
                        while (sp > 0) {
                            sp--;
                            if (!culled(stk_tmin[sp])) { cur = stk_ref[sp]; cur_tmin = stk_tmin[sp]; cur_tmax = stk_tmax[sp]; got = true; break; }
                        }
                        if (!got) { have = false; break; }
                    }
                }
                if (!have) break;
                // ---- leaf: cur == ~triangle --------------------------------------------------------
                if (COUNT) c_tris++;
                moller_trumbore(r, P.tris, ~cur, best, id);
                bool got = false;
                while (sp > 0) {
                    sp--;
                    if (!culled(stk_tmin[sp])) { cur = stk_ref[sp]; cur_tmin = stk_tmin[sp]; cur_tmax = stk_tmax[sp]; got = true; break; }
                }
                if (!got) have = false;
            }

            // ---- shade + write ------------------------------------------------------------------
            uint32_t color = P.background;
            if (id >= 0) {
                color = phong(P, M, r, best, id, cmx, cmy, cmz);
                if (COUNT) c_hits++;
            }
            const long long o = (long long)frame * P.W * P.H + pix;
            // frames are write-once streams: keep them from displacing the scene in L2
            if (P.out_bgra) __stcs(P.out_bgra + o, color);
            if (P.out_ids) __stcs(P.out_ids + o, id);
        }
    }
    if (COUNT) {
        // warp-reduce then one atomic per counter per warp
        for (int s = 16; s > 0; s >>= 1) {
            c_rays += __shfl_down_sync(0xffffffffu, c_rays, s);
            c_nodes += __shfl_down_sync(0xffffffffu, c_nodes, s);
            c_boxes += __shfl_down_sync(0xffffffffu, c_boxes, s);
            c_tris += __shfl_down_sync(0xffffffffu, c_tris, s);
            c_hits += __shfl_down_sync(0xffffffffu, c_hits, s);
        }
        if (lane == 0) {
            atomicAdd(P.counters + 0, c_rays); atomicAdd(P.counters + 1, c_nodes); atomicAdd(P.counters + 2, c_boxes);
            atomicAdd(P.counters + 3, c_tris); atomicAdd(P.counters + 4, c_hits);
        }
    }
}

}  // namespace rtb
