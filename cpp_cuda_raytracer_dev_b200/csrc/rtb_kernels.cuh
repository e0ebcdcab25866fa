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
//             q2 = R.t0z R.t1x R.t1y R.t1z      q3 = left_ref right_ref S1 S2
//           S1/S2 are copies of L.t1[axis] / R.t0[axis] so the hot loop needs no axis-indexed
//           select; a child reference is a 32-bit word: bit 31 = leaf, bits 29-30 = split axis of
//           THIS node (stored in left_ref only), bits 0-28 = record index or triangle id.
//   tris  : one 48-byte record (3 x float4) per triangle
//             t0 = e1.xyz n.x    t1 = e2.xyz n.y    t2 = (cam_pos - p1).xyz n.z
//   rad   : float4 per triangle (r,g,b,-) or absent when the mesh has one colour.
//
// Numerical contract: every value that feeds a DECISION (slab tests, plane compares,
// Moller-Trumbore u/v/w) is computed with explicitly rounded fp32 operations in the reference's
// association order -- the translation unit is compiled with -fmad=false and the code below uses
// __fmul_rn/__fadd_rn/__fsub_rn so nothing can be contracted into an FMA.  Comparisons that the
// reference performs in double precision (because its epsilons are double literals,
// vector.cuh:10-11) are reproduced exactly; see the exact_* helpers and their fp32 shortcuts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

constexpr int kTile = 32;        // multi-GPU interleave granularity: 32x32 pixel tiles
constexpr int kStackDepth = 40;  // >= tree height + 2 (median split: height = ceil(log2 n))
constexpr int kBlockThreads = 128;
constexpr int kFrameStride = 16;
constexpr int kPushSlots = 8;    // work units a warp may have open at once in the peer-push variant
constexpr int kMaxPushOwners = 8;  // GPUs that may own final frames of one pushed launch (one NVSwitch domain)
constexpr unsigned kRefLeaf = 0x80000000u, kRefIndexMask = 0x1fffffffu;
constexpr int kRefAxisShift = 29;

struct RenderParams {
    int W, H;
    float n_mod[3], u_mod[3], v_mod[3];
    float root_box[6];  // t0x t0y t0z t1x t1y t1z of node 0, camera-relative
    float draw_distance;
    uint32_t background;  // 0x00RRGGBB
    int root_ref;         // encoded reference of node 0 (record 0, or leaf|triangle for a single-triangle mesh)
    const float4* __restrict__ nodes;
    const float4* __restrict__ tris;
    const float4* __restrict__ rad;  // may be null
    float uniform_rad[3];
    const float* __restrict__ frames;  // kFrameStride floats per frame: rows x,y,z = (i,j,k,w), then the pixel rectangle
                                       // x0,y0,x1,y1 (int bits, inclusive) outside which no ray can reach the root box
    const int* __restrict__ frame_order;  // optional: the k-th frame PROCESSED is frame_order[k] (views sorted for cache reuse);
                                          // null = in index order.  Output placement is by the frame's own index either way.
    float frame0[kFrameStride];        // render_stream_kernel<.., INLINE = true>: the record of a single-frame launch travels
                                       // with the kernel parameters (constant bank) instead of through device memory
    int num_frames;
    int tiles_x;           // tiles per image row
    // the launch covers the tiles of the window [win_tx0, win_tx0 + win_tw) x [win_ty0, ...) only (the whole frame unless the
    // caller knows that nothing outside can change, as the single-frame path into the host frame does); tile indices below
    // count inside the window, row by row
    int win_tx0, win_ty0, win_tw;
    int tile_first, tile_stride;
    int my_tiles;          // number of 32x32 tiles of one frame rendered by this launch
    int unit_shift;        // log2 pixels per work unit: 5..10, a Morton block of a 32x32 tile (first segment)
    // The frames of a launch form up to three consecutive segments with decreasing unit size, so that the last units
    // handed out are small and the launch does not end with a few warps grinding through large units while the rest
    // of the GPU idles.  seg_items[k] = seg_frames[k] * my_tiles * (1024 >> seg_shift[k]).
    int seg_frames[3], seg_shift[3];
    long long seg_items[3];
    int t_active, t_leaf;  // refill when <= t_active lanes still traverse; leaf step when >= t_leaf lanes wait at a leaf
    int steal_mask;        // single-frame launches: a draining warp looks for lanes to share rays with every steal_mask + 1 iterations
    int prefetch;          // single-frame launches: 1 = ask for both children's records as soon as a record arrives
    long long total_items; // work units: num_frames * my_tiles * (1024 >> unit_shift)
    long long frame_stride;  // output elements per frame: W*H (row-major) or my_tile_slots*1024 (tile-major)
    int tile_major;          // 1: compact tile-major output (the multi-GPU exchange format), 0: final row-major position
    uint32_t* __restrict__ out_bgra;
    int32_t* __restrict__ out_ids;
    // peer push (render_stream_kernel<.., PUSH = true>): out_* is this GPU's tile-major staging, and every finished
    // work unit is copied by its warp, as 16-byte row segments, to its final row-major place in these full-frame
    // buffers (W*H elements per frame) -- which may be another GPU's memory mapped over NVLink
    // frame f of the launch belongs to owner f % push_owners and is frame f / push_owners of that owner's buffers (striped
    // ownership: every GPU's NVLink ingest and HBM take 1/N of the finished frames; one owner = everything to one GPU)
    uint32_t* push_bgra[kMaxPushOwners];
    int32_t* push_ids[kMaxPushOwners];
    int push_owners;
    int push_skip_background;  // 1: the owner of push_* pre-filled the frames with background / -1: background-only units are not
                               // sent.  2: the owner's frame holds an earlier frame of this camera whose content lies inside
                               // push_prev_rect and is background outside it: background-only units are sent only where they
                               // overlap that rectangle (single-frame path into the camera's host frame, rtb_object_render)
    int push_prev_rect[4];     // x0, y0, x1, y1 (inclusive)
    unsigned long long* work_counter;  // never reset: a launch hands out units work_base, work_base + 1, ... and every warp
    unsigned long long work_base;      // overshoots exactly once, so the host knows the counter's value after the launch
    unsigned long long* counters;  // [0] rays [1] interior nodes entered [2] nodes popped (reference sense) [3] triangle tests [4] hits
                                   // [5] sum over traced rays of their deepest stack (entries incl. the register top) [6] deepest stack of any ray
    float cull_rel;
};

// ---------------------------------------------------------------------------------------------
// exact comparison helpers.  EPS = 1e-16 (double).  kEpsUp is the smallest float >= 1e-16, so for a
// float x:  (double)x < 1e-16  <=>  x < kEpsUp   and   (double)x > -1e-16  <=>  x > -kEpsUp.
// ---------------------------------------------------------------------------------------------
#define RTB_EPS_UP __int_as_float(0x24e69595)
#define RTB_TINY 3.7252902984619140625e-09f /* 2^-28: above this, neighbouring floats are > 1e-16 apart */

// Exact double-precision forms of the reference's epsilon comparisons.  The hot loop uses fp32
// shortcuts that are provably equal to these whenever they can decide (see child_order / box_entered
// in rtb_render.cuh) and calls the functions below only for the rare undecidable inputs: operands
// below 2^-28 in magnitude, exact ties, NaN.
__device__ __forceinline__ bool exact_ge_minus_eps(float hi, float lo) { return (double)hi >= (double)lo - 1e-16; }   // Trixel.cu:146
__device__ __forceinline__ bool exact_lt_plus_eps(float a, float s) { return (double)a < (double)s + 1e-16; }         // Trixel.cu:155
__device__ __forceinline__ bool exact_gt_minus_eps(float b, float s) { return (double)b > (double)s - 1e-16; }        // Trixel.cu:156
// (float)(((double)S1 + 1e-16) + (double)ds)   (Trixel.cu:150: `s1 = cvm->s1[cni] + EPS + ds`)
__device__ __forceinline__ float exact_s1(float S1, float ds) {
    return __double2float_rn(__dadd_rn(__dadd_rn((double)S1, 1e-16), (double)ds));
}

// sm_100a three-input min/max (FMNMX3).  NaN operands are ignored exactly like nested fmax/fmin.
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// vector.cuh:79-95 + 117-120: Quake start value, 21 Newton steps, then scale.
// Each step is a pure function of the previous iterate, so once the sequence reaches a fixed point
// (or a 2-cycle) the value after exactly 21 steps is known without running them: the result is
// bit-identical to the full loop (a NaN iterate never compares equal and runs all 21 steps).
__device__ __forceinline__ float rsqrt21(float s) {
    const float half = __fmul_rn(0.5f, s);
    float g0 = __int_as_float(0x5f375a86 - (__float_as_int(half) >> 1));
    float g1 = __fmul_rn(g0, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(half, g0), g0)));
    int k = 1;
#pragma unroll 1
    while (k < 21) {
        const float g2 = __fmul_rn(g1, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(half, g1), g1)));
        k++;
        if (g2 == g1) break;                                        // fixed point
        if (g2 == g0) { g1 = ((21 - k) & 1) ? g1 : g2; break; }      // 2-cycle: parity of the remaining steps
        g0 = g1; g1 = g2;
    }
    return g1;
}
__device__ __forceinline__ void normalize21(float& x, float& y, float& z) {
    const float g = rsqrt21(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    x = __fmul_rn(x, g); y = __fmul_rn(y, g); z = __fmul_rn(z, g);
}
// the literal 21-step loop (pack kernels, and the reference for the shortcut above in tests)
__device__ __forceinline__ float rsqrt21_literal(float s) {
    const float half = __fmul_rn(0.5f, s);
    float g = __int_as_float(0x5f375a86 - (__float_as_int(half) >> 1));
#pragma unroll
    for (int k = 0; k < 21; k++) g = __fmul_rn(g, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(half, g), g)));
    return g;
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
    tmin = fmax3(t0z, t0x, t0y);  // fmax(t0z, fmax(t0x, t0y)), Trixel.cu:94
    tmax = fmin3(t1z, t1x, t1y);  // Trixel.cu:95
}
// Trixel.cu:146: enter iff (double)tmax >= (double)tmin - EPS && (double)tmin > -EPS.
// fp32 shortcut: tmax >= tmin proves the first clause; tmax < tmin with |tmin| >= 2^-28 refutes it
// (tmin - 1e-16 then rounds to a double strictly above the float below tmin).  `unsure` is raised
// for everything else (tiny |tmin|, NaN) and the caller re-evaluates with box_entered_exact.
__device__ __forceinline__ bool box_entered_fast(float tmin, float tmax, bool& unsure) {
    const bool ge = tmax >= tmin;
    unsure |= !ge & !(fabsf(tmin) >= RTB_TINY);
    return ge & (tmin > -RTB_EPS_UP);
}
__device__ __forceinline__ bool box_entered_exact(float tmin, float tmax) {
    return exact_ge_minus_eps(tmax, tmin) && (tmin > -RTB_EPS_UP);
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
    const int axis = cut_flag[i] % 3;
    const unsigned lref = (record_of[l] >= 0 ? (unsigned)record_of[l] : (kRefLeaf | (unsigned)tri[l])) | ((unsigned)axis << kRefAxisShift);
    const unsigned rref = record_of[r] >= 0 ? (unsigned)record_of[r] : (kRefLeaf | (unsigned)tri[r]);
    float4* o = nodes + 4ll * rec;
    o[0] = make_float4(L[0], L[1], L[2], L[3]);
    o[1] = make_float4(L[4], L[5], R[0], R[1]);
    o[2] = make_float4(R[2], R[3], R[4], R[5]);
    o[3] = make_float4(__uint_as_float(lref), __uint_as_float(rref), L[3 + axis], R[axis]);
}

// Reassembly of a frame from per-rank tile-major buffers (the receiving side of the multi-GPU
// exchange): tile t of the image was rendered by rank t % world into slot t / world of that rank's
// buffer; every rank's buffer has `slots` 32x32 slots per frame.
struct ComposeParts { const uint32_t* part[8]; };
// grid = (ceil(W/4 / 128), H, frames), block = 128: one thread moves 4 horizontally adjacent pixels
// (they always lie in one tile row), 16-byte loads and stores when the frame width allows it.
__global__ void compose_tiles_kernel(ComposeParts parts, int world, int W, int H, int tiles_x, long long slots, uint32_t* __restrict__ out) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= W) return;
    const long long frame = blockIdx.z;
    const int t = (y / kTile) * tiles_x + (x / kTile);
    const int rank = t % world, slot = t / world;
    const uint32_t* src = parts.part[rank] + (frame * slots + slot) * (kTile * kTile) + (y % kTile) * kTile + (x % kTile);
    uint32_t* dst = out + (frame * H + y) * (long long)W + x;
    if ((W & 3) == 0) {
        __stcs(reinterpret_cast<uint4*>(dst), __ldcs(reinterpret_cast<const uint4*>(src)));
    } else {
        for (int k = 0; k < 4 && x + k < W; k++) __stcs(dst + k, __ldcs(src + k));
    }
}

// Both frame arrays of a block of frames in one grid-stride kernel with 16-byte stores (either array may be null): meant to
// run BESIDE the persistent render kernel on a few blocks per SM (rtb_fill_frames_device_async), not to own the GPU.
__global__ void fill_frames_kernel(uint32_t* __restrict__ bgra, int32_t* __restrict__ ids, long long n, uint32_t background) {
    const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(bgra) | reinterpret_cast<uintptr_t>(ids)) & 15u) == 0;
    if (vec) {
        const uint4 b4 = make_uint4(background, background, background, background);
        const int4 i4 = make_int4(-1, -1, -1, -1);
        for (long long i = first; i < (n >> 2); i += stride) {
            if (bgra) reinterpret_cast<uint4*>(bgra)[i] = b4;
            if (ids) reinterpret_cast<int4*>(ids)[i] = i4;
        }
    } else {
        for (long long i = first; i < n; i += stride) {
            if (bgra) bgra[i] = background;
            if (ids) ids[i] = -1;
        }
    }
}
__global__ void fill_kernel(uint32_t* __restrict__ out, long long n, uint32_t value) {  // grid-stride
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = value;
}
__global__ void fill_ids_kernel(int32_t* __restrict__ out, long long n, int32_t value) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = value;
}

// L2 read-bandwidth probe for the roofline ("L2_BW measured on the box by an L2-resident read
// microbenchmark"): every thread streams 16-byte loads that bypass L1 (ld.global.cg) over a buffer small enough to
// stay in L2, `iters` times; the XOR of everything read is stored so that nothing can be optimised away.
__global__ void l2_read_kernel(const uint4* __restrict__ buf, long long n16, int iters, uint4* __restrict__ sink) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) sink[0] = acc;  // practically never true: keeps the loads alive
}

// ---------------------------------------------------------------------------------------------
// the hot kernel
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Moller-Trumbore, Trixel.cu:98-145.  Returns true and updates best/id on acceptance.
// `tie` (optional): raised when the test yields exactly the current best distance for a triangle other than the current
// best one -- the reference keeps whichever it visited first, which only matters to callers that split a walk (STEAL).
__device__ __forceinline__ bool moller_trumbore(const Ray& r, const float4* __restrict__ tris, int tri, float& best, int& id, bool* tie = nullptr) {
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
    if (tie && (w == best) && !reject && id >= 0 && id != tri) *tie = true;
    return false;
}

// Camera.cu:19-69 color_cam_cuda for a hit pixel; returns 0x00RRGGBB.
// M = the rotation part of the object's matrix: rows (m0 m1 m2), (m4 m5 m6), (m8 m9 m10)
__device__ __forceinline__ uint32_t phong(const RenderParams& P, float m0, float m1, float m2, float m4, float m5, float m6, float m8, float m9,
                                          float m10, const Ray& r, float best, int id, float cmx, float cmy, float cmz) {
    // hit record as written at Trixel.cu:134-140
    const float pntx = __fadd_rn(__fmul_rn(best, r.dx), r.ox);
    const float pnty = __fadd_rn(__fmul_rn(best, r.dy), r.oy);
    const float pntz = __fadd_rn(__fmul_rn(best, r.dz), r.oz);
    const float n0 = ldg4(P.tris + 3ll * id).w, n1 = ldg4(P.tris + 3ll * id + 1).w, n2 = ldg4(P.tris + 3ll * id + 2).w;
    // VEC3_CUDA::device_rotate(rot_m, i, -1), vector.cuh:25-33
    const float ax = __fmul_rn(-1.0f, n0), ay = __fmul_rn(-1.0f, n1), az = __fmul_rn(-1.0f, n2);
    float nx = __fadd_rn(__fadd_rn(__fmul_rn(ax, m0), __fmul_rn(ay, m1)), __fmul_rn(az, m2));
    float ny = __fadd_rn(__fadd_rn(__fmul_rn(ax, m4), __fmul_rn(ay, m5)), __fmul_rn(az, m6));
    float nz = __fadd_rn(__fadd_rn(__fmul_rn(ax, m8), __fmul_rn(ay, m9)), __fmul_rn(az, m10));
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

}  // namespace rtb
