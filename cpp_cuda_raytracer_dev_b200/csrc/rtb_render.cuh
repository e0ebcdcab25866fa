// rtb_render.cuh -- the hot kernel: persistent warps that stream pixels through a lane-level
// state machine (ray generation -> traversal -> Moller-Trumbore -> Phong -> frame store).
//
// Why this shape (ncu evidence in profiles/): the first version rendered one 8x4 pixel tile per
// warp and ran at 13-14 of 32 active threads per instruction -- rays of one tile need very
// different numbers of node visits (the meshes have ~14 triangles per pixel), lanes waiting at a
// leaf idled through the other lanes' interior steps, and finished lanes idled until the slowest
// ray of the tile was done.  Here a warp owns no tile: every lane that runs out of work is refilled
// with the next pixel of the warp's current work unit (a Morton-ordered block of a 32x32 image
// tile, so refills stay spatially coherent), units are fetched with one atomic per warp, and the
// warp alternates between warp-uniform phases chosen by ballot:
//     interior step  (fetch one 64-byte node record, test both child boxes, push / descend)
//     leaf step      (one Moller-Trumbore test)
//     retire+refill  (Phong + streaming store for finished rays, ray setup for new pixels)
// The visit ORDER of every ray is exactly the reference's (Trixel.cu:70-170): first-visited wins
// ties, so hit ids stay bit-identical; distance culling only drops subtrees that cannot contain a
// closer hit (see `cull_base`).
#pragma once
#include "rtb_kernels.cuh"

namespace rtb {

constexpr int kStateEmpty = 0, kStateTraverse = 1, kStateDone = 2;
// single-frame launches only (STEAL, see render_stream_kernel): a lane that finished a subtree it took over from another
// lane and has not reported yet / a lane whose own walk is over while helpers still hold parts of its ray's stack
constexpr int kStateHelperDone = 3, kStateWait = 4;

__device__ __forceinline__ uint32_t compact_even_bits(uint32_t x) {
    x &= 0x55555555u;
    x = (x ^ (x >> 1)) & 0x33333333u;
    x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
    x = (x ^ (x >> 4)) & 0x00ff00ffu;
    x = (x ^ (x >> 8)) & 0x0000ffffu;
    return x;
}

// Order of the 1024 pixels of a 32x32 tile in which a warp consumes them, and with it the shape of a work unit (2^s
// consecutive ordinals).  RTB_TILE_ORDER 0: Morton (units 8x4, 8x8, 16x8, 16x16, ...).  1: Morton inside 4x4 blocks, then
// the blocks of a row of blocks, then the rows: ordinal bits x0 y0 x1 y1 x2 x3 x4 y2 y3 y4 -- units are 8x4, 16x4, 32x4,
// 32x8, ...: a 128-pixel unit is four whole 128-byte rows of its tile, which is what the peer push wants to send.
#ifndef RTB_TILE_ORDER
#define RTB_TILE_ORDER 1
#endif
__device__ __forceinline__ void tile_xy(uint32_t k, int& x, int& y) {
#if RTB_TILE_ORDER == 0
    x = (int)compact_even_bits(k); y = (int)compact_even_bits(k >> 1);
#else
    x = (int)((k & 1u) | ((k >> 1) & 2u) | ((k >> 2) & 0x1cu));
    y = (int)(((k >> 1) & 1u) | ((k >> 2) & 2u) | ((k >> 5) & 0x1cu));
#endif
}
// log2 of the width in pixels of a unit of 2^shift pixels (shift 5..10); its height is 2^(shift - this)
__device__ __forceinline__ int unit_wshift(int shift) {
#if RTB_TILE_ORDER == 0
    return (shift + 1) >> 1;
#else
    return shift - 2 < 5 ? shift - 2 : 5;
#endif
}

// ordered "not equal": false when either operand is NaN
__device__ __forceinline__ bool ordered_ne(float x, float y) { return (x < y) | (x > y); }

// Exact evaluation of everything the fp32 shortcuts of one interior step decide (Trixel.cu:146,
// 149-168), for the rare inputs they cannot: operands below 2^-28 in magnitude, exact ties, NaN.
// Returns bit 0 = left child first, bit 1 = second child scheduled, bit 2 = left box entered,
// bit 3 = right box entered.
__device__ __forceinline__ int interior_decisions_exact(float a, float b, float s2, float S1, float ds, float ltmin, float ltmax,
                                                     float rtmin, float rtmax) {
    int out = 0;
    if (exact_lt_plus_eps(a, s2)) {  // Trixel.cu:155-161: left is popped first, right only if the exit lies beyond s2
        out |= 1;
        if (exact_gt_minus_eps(b, s2)) out |= 2;
    } else {  // Trixel.cu:162-168
        const float s1 = exact_s1(S1, ds);
        if ((b < s1) || (a < s1)) out |= 2;
    }
    if (box_entered_exact(ltmin, ltmax)) out |= 4;
    if (box_entered_exact(rtmin, rtmax)) out |= 8;
    return out;
}

// The fp32 shortcuts of the same decisions (same bit layout), valid whenever `decidable` comes back true:
//   (double)a < (double)s2 + EPS  ==  a < s2        } when a != s2, b != s2 and |s2| >= 2^-28:
//   (double)b > (double)s2 - EPS  ==  b > s2        } s2 +- 1e-16 stays strictly between s2's float neighbours
//   (b < s1) || (a < s1)          ==  min(a,b) < sf   when min(a,b) != sf and |sf| >= 2^-27, sf = fl(S1 + ds):
//        s1 = (float)(((double)S1 + EPS) + (double)ds) is sf or its upper neighbour, never below sf
//   (double)tmax >= (double)tmin - EPS  ==  tmax >= tmin   when |tmin| >= 2^-28
// `decidable` is false for ties, tiny operands and NaN; those lanes take the exact path.
__device__ __forceinline__ void interior_decisions_fast(float a, float b, float s2, float S1, float ds, float ltmin, float ltmax, float rtmin,
                                                        float rtmax, bool& decidable, bool& left_first, bool& visit_second, bool& l_in,
                                                        bool& r_in) {
    const float sf = __fadd_rn(S1, ds);
    const float mab = fminf(a, b);
    decidable = (fabsf(s2) >= RTB_TINY) & (fabsf(sf) >= 2.0f * RTB_TINY) & ordered_ne(a, s2) & ordered_ne(b, s2) & ordered_ne(mab, sf) &
                (fabsf(ltmin) >= RTB_TINY) & (fabsf(rtmin) >= RTB_TINY);
    left_first = a < s2;
    visit_second = left_first ? (b > s2) : (mab < sf);
    l_in = (ltmax >= ltmin) & (ltmin > -RTB_EPS_UP);
    r_in = (rtmax >= rtmin) & (rtmin > -RTB_EPS_UP);
}

// Self-test of the exactness arguments (rtb_selftest_exact): every thread draws operands -- raw random
// bit patterns over all exponents, plus neighbours within a few ulps of each other to provoke ties --
// and checks (0) the early-exit Newton rsqrt against the literal 21-step loop, (1) __frcp_rn against
// the reference's (float)(1.0 / (double)f), (2) the fp32 decision shortcuts against the exact
// double-precision evaluation wherever they claim to be decidable.  out[3] counts decidable samples.
__device__ __forceinline__ uint32_t selftest_hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}
__global__ void selftest_exact_kernel(uint64_t seed, long long count, unsigned long long* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float v[9];
    const uint32_t mode = selftest_hash(seed + 977 * (uint64_t)i) % 4u;
    for (int k = 0; k < 9; k++) {
        uint32_t bits = selftest_hash(seed * 31 + (uint64_t)i * 16 + k);
        if (mode >= 1) bits = (bits & 0x807fffffu) | ((96u + (bits >> 23) % 64u) << 23);  // magnitudes 2^-31 .. 2^32
        v[k] = __uint_as_float(bits);
    }
    if (mode >= 2) {  // provoke ties: derive operands from each other within a few ulps
        const int d = (int)(selftest_hash(seed + 5 * (uint64_t)i) % 5u) - 2;
        v[2] = __uint_as_float(__float_as_uint(v[0]) + d);                         // s2 ~ a
        v[1] = mode == 3 ? v[2] : v[1];                                            // b == s2
        v[6] = __uint_as_float(__float_as_uint(v[5]) + (d & 1));                   // ltmax ~ ltmin
        v[3] = __fsub_rn(v[0], v[4]);                                              // S1 + ds ~ a
    }
    unsigned long long bad0 = 0, bad1 = 0, bad2 = 0, dec = 0;
    const float s = fabsf(v[0]);
    if (__float_as_uint(rsqrt21(s)) != __float_as_uint(rsqrt21_literal(s)) && !(rsqrt21(s) != rsqrt21(s) && rsqrt21_literal(s) != rsqrt21_literal(s))) bad0 = 1;
    const float f = v[1];
    if (f == f && f != 0.0f && fabsf(f) <= 3.0e38f) {
        const float ref = (float)(1.0 / (double)f);
        if (__float_as_uint(__frcp_rn(f)) != __float_as_uint(ref)) bad1 = 1;
    }
    bool decidable, lf, vs, li, ri;
    interior_decisions_fast(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], decidable, lf, vs, li, ri);
    if (decidable) {
        dec = 1;
        const int ex = interior_decisions_exact(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
        const int fast = (int)lf | ((int)vs << 1) | ((int)li << 2) | ((int)ri << 3);
        if (ex != fast) bad2 = 1;
    }
    if (bad0) atomicAdd(out + 0, 1ull);
    if (bad1) atomicAdd(out + 1, 1ull);
    if (bad2) atomicAdd(out + 2, 1ull);
    if (dec) atomicAdd(out + 3, 1ull);
}

// PUSH = the multi-GPU tile exchange fused into the render: the warp writes its pixels into a local tile-major
// staging buffer as usual, keeps a count of the pixels each of its open work units still owes, and the moment a
// unit is complete copies it -- 16-byte row segments, one load and one store per lane and array for a 128-pixel
// unit -- to its final row-major place in P.push_*, typically rank 0's frame mapped over NVLink.  The transfer
// therefore overlaps the rendering unit by unit, and there is no gather, no receive buffer and no reassembly pass.
// A unit belongs to exactly one warp, so the bookkeeping is warp-local (shared-memory counters, __syncwarp).
// single-pixel stores of retiring rays (4 bytes, scattered over the sectors of an 8x4 pixel block in time)
// RTB_PIXEL_STORE_MODE: 0 = streaming (st.global.cs), 1 = write-back (plain store), 2 = cache-global (st.global.cg)
#ifndef RTB_PIXEL_STORE_MODE
#define RTB_PIXEL_STORE_MODE 0
#endif
#if RTB_PIXEL_STORE_MODE == 1
#define RTB_PIXEL_STORE(ptr, value) (*(ptr) = (value))
#elif RTB_PIXEL_STORE_MODE == 2
#define RTB_PIXEL_STORE(ptr, value) __stcg(ptr, value)
#else
#define RTB_PIXEL_STORE(ptr, value) __stcs(ptr, value)
#endif
#ifndef RTB_MIN_BLOCKS
#define RTB_MIN_BLOCKS 8
#endif
#ifndef RTB_MIN_BLOCKS_PUSH
#define RTB_MIN_BLOCKS_PUSH 7  // the push variant carries more warp state: 72 registers spill nothing
#endif
// INLINE: single-frame launch whose frame record is P.frame0 (kernel parameter space) -- the per-frame path of the
// reference's loop (Object::render, WinMain.cpp:212) then needs no upload before the launch.
//
// INLINE launches also share long rays between the lanes of a warp (STEAL).  A single frame has fewer rays than the GPU
// has lanes, so it lasts as long as its longest ray: 300-400 dependent steps along the silhouette against ~40 for a ray
// that hits the object squarely (measured, rtb_camera_counters_ex[7]).  Once a warp has no pixels left to fetch, its idle
// lanes take over the BOTTOM entry of a busy lane's traversal stack -- the subtree that lane would have visited last --
// together with a copy of the ray, walk it with the donor's current best distance as their bound, and report what they
// found to the lane that owns the pixel.  The result is the reference's: the closest hit wins; the reference breaks exact
// ties between different triangles by visit order (Trixel.cu:127: strict `<`), which a split walk does not know, so a
// pixel that sees such a tie while its walk may be split is traced again by its owner alone, ties ignored (3_walls:
// every hit is a three-way tie; elsewhere it practically never happens).  Ties are only looked for once the warp has
// run out of pixels, because nothing is shared before that.
#ifndef RTB_STEAL
#define RTB_STEAL 1
#endif
#ifndef RTB_STEAL_ALL
#define RTB_STEAL_ALL 0  // 1: multi-frame launches share long rays too, when their queue has run dry (the tail of a launch)
#endif
#ifndef RTB_MIN_BLOCKS_INLINE
#define RTB_MIN_BLOCKS_INLINE 6  // a single frame never fills the GPU: registers matter more than resident warps (80 instead of 64)
#endif
template <bool CULL, bool COUNT, bool PUSH, bool INLINE = false>
__global__ void __launch_bounds__(kBlockThreads, INLINE ? RTB_MIN_BLOCKS_INLINE : PUSH ? RTB_MIN_BLOCKS_PUSH : RTB_MIN_BLOCKS)
render_stream_kernel(const RenderParams P) {
    constexpr bool STEAL = (INLINE || (RTB_STEAL_ALL != 0 && !COUNT)) && (RTB_STEAL != 0);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    __shared__ int s_helpers[STEAL ? kBlockThreads / 32 : 1][STEAL ? 32 : 1];  // per lane: helpers still out with parts of its ray
    int sbase = 0;            // STEAL: index of the bottom entry of this lane's stack (entries below it were given away)
    int owner = (int)lane;    // STEAL: the lane whose pixel this lane is working for
    bool tie = false, solo = false;  // STEAL: owner saw an exact tie between lanes / traces its pixel again alone
    int spin = 0, rounds = 0;
    volatile int* const helpers = s_helpers[STEAL ? threadIdx.x >> 5 : 0];
    if (STEAL) { helpers[lane] = 0; __syncwarp(); }
#ifdef RTB_WARP_LOG  // development builds only (tools/warp_log.py): what every warp did and when, in nanoseconds
    unsigned long long wl_t0, wl_first_work = 0, wl_exhausted = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(wl_t0));
    unsigned wl_units = 0, wl_bg_units = 0, wl_iters = 0, wl_rays = 0, wl_steals = 0;
#endif
    __shared__ int s_owed[PUSH ? kBlockThreads / 32 : 1][PUSH ? kPushSlots : 1];   // pixels of an open unit not yet written
    __shared__ int4 s_unit[PUSH ? kBlockThreads / 32 : 1][PUSH ? kPushSlots : 1];  // frame, tile slot, x and y offset inside the tile
    const int wib = threadIdx.x >> 5;
    unsigned open_mask = 0u;   // warp-uniform: slots of s_owed/s_unit in use
    int my_pslot = 0, u_pslot = 0;
    bool blocked = false;      // warp-uniform: no free slot for the next unit -> drain in-flight rays first
    unsigned long long c_nodes = 0, c_boxes = 0, c_tris = 0, c_rays = 0, c_hits = 0, c_depth = 0;
    int ray_depth = 0, c_depth_max = 0;  // COUNT: deepest stack of the lane's current ray / of any ray of this lane
    int ray_steps = 0, c_steps_max = 0;  // COUNT: interior + leaf steps of the lane's current ray / the longest ray of this lane

    // Traversal stack: the top entry lives in registers (top_*), deeper entries in local memory.
    // A pop hands out the register copy at once and re-loads the new top in the background, so the
    // memory latency of the stack is off the dependent chain.
    int stk_ref[kStackDepth];
    float stk_tmin[kStackDepth], stk_tmax[kStackDepth];
    int top_ref = 0;
    float top_tmin = 0.0f, top_tmax = 0.0f;

    // ---- per-lane ray state ---------------------------------------------------------------------
    Ray r;
    r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = r.fx = r.fy = r.fz = r.ox = r.oy = r.oz = 0.0f;
    float cmx = 0.0f, cmy = 0.0f, cmz = 0.0f;  // camera-space ray (Camera::pixel_memory::rmd), needed again by Phong
    float best = 0.0f, slack_abs = 0.0f, cull_base = 0.0f, cur_tmin = 0.0f, cur_tmax = 0.0f;
    int id = -1, cur = 0, sp = 0, frame = 0, pix = 0;
    int state = kStateEmpty;
    bool want_pop = false;

    // ---- warp-uniform unit state ----------------------------------------------------------------
    int u_shift = P.seg_shift[0];  // log2 pixels of the warp's current unit
    int u_next = 1 << u_shift;     // nothing loaded yet
    int u_frame = 0, u_x0 = 0, u_y0 = 0, u_base = 0, u_slot = 0;
    int u_rx0 = 0, u_ry0 = 0, u_rx1 = -1, u_ry1 = -1;  // pixel rectangle of the unit's frame that can reach the root box
    bool exhausted = false;

    // Distance culling that cannot change the answer: a subtree (or leaf) is skipped only if its box
    // entry lies beyond the current best hit by more than the rounding error the slab arithmetic of
    // this ray can have (slack_abs) plus a relative margin: tmin > best + slack_abs + rel*(|tmin|+|best|).
    // cull_base caches the terms that only change with `best`.  Never true for NaN/inf operands.
    auto set_cull_base = [&]() { cull_base = best + (slack_abs + P.cull_rel * fabsf(best)); };
    auto culled = [&](float tmin) -> bool { return CULL && (tmin > __fmaf_rn(P.cull_rel, fabsf(tmin), cull_base)); };

    for (;;) {
        unsigned m_trav = __ballot_sync(0xffffffffu, state == kStateTraverse);

        // ================= traversal: steps until enough lanes have run dry ==========================
        // (while pixels remain, fall out to retire + refill as soon as no more than t_active lanes are
        // still traversing; once the work is exhausted, drain)
        const int keep_active = (exhausted | (PUSH && blocked)) ? 0 : P.t_active;
        while (__popc(m_trav) > keep_active) {
#ifdef RTB_WARP_LOG
            wl_iters++;
#endif
            // ---- lanes that finished a node or leaf take the stack top ------------------------------
            if (want_pop) {
                if (sp == (STEAL ? sbase : 0)) {
                    want_pop = false;
                    if (!STEAL) state = kStateDone;
                    else state = owner != (int)lane ? kStateHelperDone : (helpers[lane] != 0 ? kStateWait : kStateDone);
                } else {
                    if (!culled(top_tmin)) { cur = top_ref; cur_tmin = top_tmin; cur_tmax = top_tmax; want_pop = false; }
                    sp--;
                    if (sp > (STEAL ? sbase : 0)) { top_ref = stk_ref[sp - 1]; top_tmin = stk_tmin[sp - 1]; top_tmax = stk_tmax[sp - 1]; }
                }
            }
            const bool ready = (state == kStateTraverse) & !want_pop;
            const bool at_leaf = ready & (cur < 0);
            const unsigned m_ready = __ballot_sync(0xffffffffu, ready);
            const unsigned m_leaf = __ballot_sync(0xffffffffu, at_leaf);
            if (m_leaf != 0u && (__popc(m_leaf) >= P.t_leaf || m_leaf == m_ready)) {
                // ---- leaf step: always intersected when popped (Trixel.cu:98) -----------------------
                if (at_leaf) {
                    if (COUNT) { c_tris++; ray_steps++; }
                    if (moller_trumbore(r, P.tris, (int)((unsigned)cur & kRefIndexMask), best, id, (STEAL && exhausted && !solo) ? &tie : nullptr)) set_cull_base();
                    want_pop = true;
                }
            } else if (ready & !at_leaf) {
                // ---- interior step: `cur` is a node whose own box test passed ------------------------
                if (COUNT) { c_nodes++; ray_steps++; }
                const float4* rec;  // P.nodes + 64 bytes * record index, as one IMAD.WIDE
                asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(rec) : "r"((unsigned)cur & kRefIndexMask), "l"(P.nodes));
                const float4 q0 = ldg4(rec), q1 = ldg4(rec + 1), q2 = ldg4(rec + 2), q3 = ldg4(rec + 3);
                const int lref = __float_as_int(q3.x), rref = __float_as_int(q3.y);
                if (INLINE && P.prefetch) {
                    // A single frame is bound by the LATENCY of its longest rays, not by throughput: a step is a record fetch
                    // (L2 or DRAM) followed by a dependent chain of ~150 instructions that decides which record comes next.
                    // Asking for both candidates now overlaps the next fetch with that chain.  (In the multi-frame kernel,
                    // which is bound by issue slots, the same requests cost more than they return: DESIGN.md section 4.)
                    const char* lp = reinterpret_cast<const char*>(lref < 0 ? (const void*)(P.tris + 3ll * ((unsigned)lref & kRefIndexMask)) : (const void*)(P.nodes + 4ll * ((unsigned)lref & kRefIndexMask)));
                    const char* rp = reinterpret_cast<const char*>(rref < 0 ? (const void*)(P.tris + 3ll * ((unsigned)rref & kRefIndexMask)) : (const void*)(P.nodes + 4ll * ((unsigned)rref & kRefIndexMask)));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(lp));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rp));
                }
                const float S1 = q3.z, S2 = q3.w;  // left child's max / right child's min on the split axis (Trixel.h:353-376)
                const int axis = (lref >> kRefAxisShift) & 3;
                // Split-axis components.  Trixel.cu:88-90 forms them as three-term sums with 0/1 flags,
                // which for finite operands is the selected component up to the sign of a zero; neither
                // the products below nor the comparisons can see that sign.
                const bool ax0 = axis == 0, ax1 = axis == 1;
                float dir = r.dz, ds = r.oz;
                dir = ax1 ? r.dy : dir; ds = ax1 ? r.oy : ds;
                dir = ax0 ? r.dx : dir; ds = ax0 ? r.ox : ds;
                const float a = __fmul_rn(cur_tmin, dir), b = __fmul_rn(cur_tmax, dir);  // Trixel.cu:149
                const float s2 = __fadd_rn(S2, ds);                                      // Trixel.cu:151
                // ---- both child boxes (children the reference would pop: leaves are always intersected,
                // interior nodes only after their own box test, Trixel.cu:98,146) -----------------------
                float ltmin, ltmax, rtmin, rtmax;
                slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ltmin, ltmax);
                slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, rtmin, rtmax);
                // ---- decisions: fp32 shortcuts of the reference's double-precision comparisons (see
                // interior_decisions_fast), exact evaluation for the inputs they cannot decide ------------
                bool decidable, left_first, visit_second, l_in, r_in;
                interior_decisions_fast(a, b, s2, S1, ds, ltmin, ltmax, rtmin, rtmax, decidable, left_first, visit_second, l_in, r_in);
                // The tail is instantiated once per path (not merged) so that on the fast path the four
                // decisions stay in predicate registers instead of being materialised for a join.
                auto descend = [&](bool left_first, bool visit_second, bool l_in, bool r_in) {
                    if (COUNT) c_boxes += 1 + (int)visit_second;
                    if (CULL) {
                        // A child that may not be entered gets a huge entry distance, so that one culling
                        // comparison per child decides everything: go <=> !(tmin_eff > cull limit).
                        // (NaN entries are never "culled", exactly as with the explicit flags.)
                        const float l_eff = ((lref < 0) | l_in) ? ltmin : 3.0e38f, r_eff = ((rref < 0) | r_in) ? rtmin : 3.0e38f;
                        const float f_eff = left_first ? l_eff : r_eff, s_eff = left_first ? r_eff : l_eff;
                        const bool go_first = !culled(f_eff);
                        const bool go_second = visit_second & !culled(s_eff);
                        if (go_second) {
                            if (sp > (STEAL ? sbase : 0)) { stk_ref[sp - 1] = top_ref; stk_tmin[sp - 1] = top_tmin; stk_tmax[sp - 1] = top_tmax; }
                            top_ref = left_first ? rref : lref; top_tmin = s_eff; top_tmax = left_first ? rtmax : ltmax;
                            sp++;
                            if (COUNT) ray_depth = max(ray_depth, sp);
                        }
                        cur = left_first ? lref : rref; cur_tmin = f_eff; cur_tmax = left_first ? ltmax : rtmax;
                        want_pop = !go_first;
                        return;
                    }
                    // order the two children first, then judge them: fewer live predicates
                    const int first = left_first ? lref : rref, second = left_first ? rref : lref;
                    const float f_tmin = left_first ? ltmin : rtmin, f_tmax = left_first ? ltmax : rtmax;
                    const float s_tmin = left_first ? rtmin : ltmin, s_tmax = left_first ? rtmax : ltmax;
                    const bool f_in = left_first ? l_in : r_in, s_in = left_first ? r_in : l_in;
                    const bool go_first = ((first < 0) | f_in) & !culled(f_tmin);
                    const bool go_second = visit_second & ((second < 0) | s_in) & !culled(s_tmin);
                    if (go_second) {
                        if (sp > (STEAL ? sbase : 0)) { stk_ref[sp - 1] = top_ref; stk_tmin[sp - 1] = top_tmin; stk_tmax[sp - 1] = top_tmax; }
                        top_ref = second; top_tmin = s_tmin; top_tmax = s_tmax;
                        sp++;
                        if (COUNT) ray_depth = max(ray_depth, sp);
                    }
                    cur = first; cur_tmin = f_tmin; cur_tmax = f_tmax;
                    want_pop = !go_first;
                };
                if (decidable) {
                    descend(left_first, visit_second, l_in, r_in);
                } else {
                    const int ex = interior_decisions_exact(a, b, s2, S1, ds, ltmin, ltmax, rtmin, rtmax);
                    descend((ex & 1) != 0, (ex & 2) != 0, (ex & 4) != 0, (ex & 8) != 0);
                }
            }
            m_trav = __ballot_sync(0xffffffffu, state == kStateTraverse);
            if (STEAL && exhausted && m_trav != 0xffffffffu && (++spin & P.steal_mask) == 0) break;  // idle lanes: see below
        }

        // ================= retire: Phong + store for finished rays ===================================
        if (STEAL && state == kStateDone && tie && !solo) {
            // two lanes found different triangles at exactly the same distance: which one the reference keeps depends
            // on its visit order, so the owner walks the whole ray again, alone
            tie = false; solo = true;
            best = P.draw_distance; id = -1; sp = 0; sbase = 0; want_pop = false;
            set_cull_base();
            cur = P.root_ref;
            if (P.root_ref < 0) { cur_tmin = 0.0f; cur_tmax = 0.0f; }
            else slab(r, P.root_box[0], P.root_box[1], P.root_box[2], P.root_box[3], P.root_box[4], P.root_box[5], cur_tmin, cur_tmax);
            state = kStateTraverse;  // (it entered the root box the first time, so it does again)
        }
        if (state == kStateDone) {
            uint32_t color = P.background;
            if (id >= 0) {
                if (INLINE) {
                    color = phong(P, P.frame0[0], P.frame0[1], P.frame0[2], P.frame0[4], P.frame0[5], P.frame0[6], P.frame0[8], P.frame0[9], P.frame0[10],
                                  r, best, id, cmx, cmy, cmz);
                } else {
                    const float* __restrict__ M = P.frames + (long long)kFrameStride * frame;
                    color = phong(P, M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10], r, best, id, cmx, cmy, cmz);
                }
                if (COUNT) c_hits++;
            }
            if (COUNT) {
                c_depth += (unsigned long long)ray_depth; c_depth_max = max(c_depth_max, ray_depth); ray_depth = 0;
                c_steps_max = max(c_steps_max, ray_steps); ray_steps = 0;
            }
            const long long o = (long long)frame * P.frame_stride + pix;
            // frames are write-once streams: keep them from displacing the scene in L2
            if (P.out_bgra) RTB_PIXEL_STORE(P.out_bgra + o, color);
            if (P.out_ids) RTB_PIXEL_STORE(P.out_ids + o, id);
            if (PUSH) atomicSub(&s_owed[wib][my_pslot], 1);
            state = kStateEmpty;
            if (STEAL) { solo = false; tie = false; }
        }
        const unsigned m_empty = ~m_trav;

        // ================= refill: next pixels of the warp's unit ====================================
        if (PUSH && !exhausted && u_next >= (1 << u_shift)) blocked = open_mask == (1u << kPushSlots) - 1u;
        if (!exhausted && u_next >= (1 << u_shift) && !(PUSH && blocked)) {
            for (;;) {
                unsigned long long uid = 0;
                if (lane == 0) uid = atomicAdd(P.work_counter, 1ull) - P.work_base;
                uid = __shfl_sync(0xffffffffu, uid, 0);
                if ((long long)uid >= P.total_items) {
                    exhausted = true;
#ifdef RTB_WARP_LOG
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(wl_exhausted));
#endif
                    break;
                }
#ifdef RTB_WARP_LOG
                wl_units++;
#endif
                // which segment of the launch (unit size), which frame, which tile, which block of the tile
                long long item = (long long)uid;
                int frame0 = 0;
                u_shift = P.seg_shift[0];
                if (item >= P.seg_items[0]) {
                    item -= P.seg_items[0]; frame0 = P.seg_frames[0]; u_shift = P.seg_shift[1];
                    if (item >= P.seg_items[1]) { item -= P.seg_items[1]; frame0 += P.seg_frames[1]; u_shift = P.seg_shift[2]; }
                }
                const int unit_pixels = 1 << u_shift;
                const int units_per_tile = (kTile * kTile) >> u_shift;
                const long long per_frame = (long long)P.my_tiles * units_per_tile;
                const int fseg = (int)(item / per_frame);
                u_frame = frame0 + fseg;
                if (!INLINE && P.frame_order) u_frame = __ldg(P.frame_order + u_frame);
                const int rem = (int)(item - (long long)fseg * per_frame);
                u_slot = rem / units_per_tile;
                const int tile = P.tile_first + u_slot * P.tile_stride;
                u_base = (rem % units_per_tile) << u_shift;
                u_x0 = (P.win_tx0 + tile % P.win_tw) * kTile;
                u_y0 = (P.win_ty0 + tile / P.win_tw) * kTile;
                if (INLINE) {
                    u_rx0 = __float_as_int(P.frame0[12]); u_ry0 = __float_as_int(P.frame0[13]);
                    u_rx1 = __float_as_int(P.frame0[14]); u_ry1 = __float_as_int(P.frame0[15]);
                } else {
                    const float* __restrict__ F = P.frames + (long long)kFrameStride * u_frame;
                    u_rx0 = __float_as_int(__ldg(F + 12)); u_ry0 = __float_as_int(__ldg(F + 13));
                    u_rx1 = __float_as_int(__ldg(F + 14)); u_ry1 = __float_as_int(__ldg(F + 15));
                }
                int xoff, yoff;
                tile_xy((uint32_t)u_base, xoff, yoff);
                if (!COUNT) {
                    // ---- a unit that lies entirely outside the frame's root-box rectangle is background: the warp
                    // writes it with 16-byte stores at its final place and goes for the next unit -----------------
                    const int wshift = unit_wshift(u_shift);  // a unit is a (1 << wshift) x (unit_pixels >> wshift) pixel block
                    const int bx = u_x0 + xoff, by = u_y0 + yoff;
                    if (bx > u_rx1 || bx + (1 << wshift) - 1 < u_rx0 || by > u_ry1 || by + (unit_pixels >> wshift) - 1 < u_ry0) {
#ifdef RTB_WARP_LOG
                        wl_bg_units++;
#endif
                        if (PUSH && P.push_skip_background == 1) continue;  // the frame's owner has pre-filled it: nothing to send
                        if (PUSH && P.push_skip_background == 2 &&
                            (bx > P.push_prev_rect[2] || bx + (1 << wshift) - 1 < P.push_prev_rect[0] || by > P.push_prev_rect[3] ||
                             by + (unit_pixels >> wshift) - 1 < P.push_prev_rect[1]))
                            continue;  // background before, background now: the owner's frame already says so
                        const int owner = PUSH ? u_frame % P.push_owners : 0;
                        uint32_t* __restrict__ dc = PUSH ? P.push_bgra[owner] : P.out_bgra;
                        int32_t* __restrict__ di = PUSH ? P.push_ids[owner] : P.out_ids;
                        const bool row_major = PUSH || !P.tile_major;
                        const long long fbase = PUSH ? (long long)(u_frame / P.push_owners) * P.W * P.H : (long long)u_frame * P.frame_stride;
                        const bool vec = !row_major || (P.W & 3) == 0;
                        for (int i = (int)lane; i < (unit_pixels >> 2); i += 32) {
                            const int row = i >> (wshift - 2), col = (i - (row << (wshift - 2))) << 2;
                            const int px = bx + col, py = by + row;
                            if (py >= P.H || px >= P.W) continue;
                            const long long o = fbase + (row_major ? (long long)py * P.W + px : ((long long)u_slot * kTile + yoff + row) * kTile + xoff + col);
                            if (vec && px + 3 < P.W) {
                                if (dc) __stcs(reinterpret_cast<uint4*>(dc + o), make_uint4(P.background, P.background, P.background, P.background));
                                if (di) __stcs(reinterpret_cast<int4*>(di + o), make_int4(-1, -1, -1, -1));
                            } else {
                                for (int k = 0; k < 4 && px + k < P.W; k++) {
                                    if (dc) __stcs(dc + o + k, P.background);
                                    if (di) __stcs(di + o + k, -1);
                                }
                            }
                        }
                        continue;
                    }
                }
                u_next = 0;
#ifdef RTB_WARP_LOG
                if (wl_first_work == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(wl_first_work));
#endif
                if (PUSH) {
                    u_pslot = __ffs(~open_mask) - 1;
                    open_mask |= 1u << u_pslot;
                    if (lane == 0) {
                        s_owed[wib][u_pslot] = unit_pixels;
                        s_unit[wib][u_pslot] = make_int4(u_frame, u_slot, xoff | (u_shift << 8), yoff);
                    }
                    __syncwarp();
                }
                break;
            }
        }
        if (!exhausted) {
            const int slot = __popc(m_empty & lanemask_lt);
            const int avail = (1 << u_shift) - u_next;
            bool settled = false;  // PUSH: this lane took a pixel that needs no ray (outside the image, or background)
            if (((m_empty >> lane) & 1u) && slot < avail) {
                settled = true;
                const uint32_t K = (uint32_t)(u_base + u_next + slot);  // ordinal inside the 32x32 tile
                int lx, ly;
                tile_xy(K, lx, ly);
                const int px = u_x0 + lx, py = u_y0 + ly;
                if (px < P.W && py < P.H) {
                    const int opix = P.tile_major ? (u_slot * kTile + ly) * kTile + lx : py * P.W + px;
                    if (px < u_rx0 || px > u_rx1 || py < u_ry0 || py > u_ry1) {
                        // outside the (conservatively enlarged) projection of the root box: the ray cannot
                        // pass the root's box test (Trixel.cu:146), the pixel is background
                        const long long o = (long long)u_frame * P.frame_stride + opix;
                        if (P.out_bgra) RTB_PIXEL_STORE(P.out_bgra + o, P.background);
                        if (P.out_ids) RTB_PIXEL_STORE(P.out_ids + o, -1);
                        if (COUNT) { c_rays++; c_boxes++; }
                    } else {
                        frame = u_frame;
                        pix = opix;
                        settled = false;
                        my_pslot = u_pslot;
                        const float* __restrict__ M = P.frames + (long long)kFrameStride * frame;
                        auto mword = [&](int k) -> float { return INLINE ? P.frame0[k] : __ldg(M + k); };
                        // ---- primary ray, Camera.cu:103-104 (row 0 = bottom) --------------------------
                        const float fxp = (float)px, fyp = (float)py;
                        cmx = __fadd_rn(__fadd_rn(P.n_mod[0], __fmul_rn(P.u_mod[0], fxp)), __fmul_rn(P.v_mod[0], fyp));
                        cmy = __fadd_rn(__fadd_rn(P.n_mod[1], __fmul_rn(P.u_mod[1], fxp)), __fmul_rn(P.v_mod[1], fyp));
                        cmz = __fadd_rn(__fadd_rn(P.n_mod[2], __fmul_rn(P.u_mod[2], fxp)), __fmul_rn(P.v_mod[2], fyp));
                        normalize21(cmx, cmy, cmz);
                        // ---- into object space, Trixel.cu:60-66 (sign dance kept for -0 fidelity) ------
                        const float m0 = mword(0), m1 = mword(1), m2 = mword(2), m3 = mword(3);
                        const float m4 = mword(4), m5 = mword(5), m6 = mword(6), m7 = mword(7);
                        const float m8 = mword(8), m9 = mword(9), m10 = mword(10), m11 = mword(11);
                        r.ox = m3; r.oy = m7; r.oz = m11;
                        r.dx = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m0, -cmx), __fmul_rn(m1, -cmy)), __fmul_rn(m2, -cmz)));
                        r.dy = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m4, -cmx), __fmul_rn(m5, -cmy)), __fmul_rn(m6, -cmz)));
                        r.dz = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(m8, -cmx), __fmul_rn(m9, -cmy)), __fmul_rn(m10, -cmz)));
                        r.ix = __frcp_rn(r.dx); r.iy = __frcp_rn(r.dy); r.iz = __frcp_rn(r.dz);
                        r.fx = __fdiv_rn(r.ox, r.dx); r.fy = __fdiv_rn(r.oy, r.dy); r.fz = __fdiv_rn(r.oz, r.dz);
                        best = P.draw_distance;  // Trixel.cu:47
                        id = -1;
                        sp = 0;
                        want_pop = false;
                        if (COUNT) c_rays++;
#ifdef RTB_WARP_LOG
                        wl_rays++;
#endif
                        if (CULL) {
                            const float bx = fmaxf(fabsf(P.root_box[0]), fabsf(P.root_box[3]));
                            const float by = fmaxf(fabsf(P.root_box[1]), fabsf(P.root_box[4]));
                            const float bz = fmaxf(fabsf(P.root_box[2]), fabsf(P.root_box[5]));
                            const float e = fmaxf(fmaxf(bx * fabsf(r.ix) + fabsf(r.fx), by * fabsf(r.iy) + fabsf(r.fy)), bz * fabsf(r.iz) + fabsf(r.fz));
                            slack_abs = e * 9.5367431640625e-07f;  // 8 * 2^-23
                            set_cull_base();
                        }
                        cur = P.root_ref;
                        if (P.root_ref < 0) {
                            // single-triangle mesh: the root is a leaf, tested unconditionally (Trixel.cu:98)
                            cur_tmin = 0.0f; cur_tmax = 0.0f;
                            state = kStateTraverse;
                            if (COUNT) c_boxes++;
                        } else {
                            slab(r, P.root_box[0], P.root_box[1], P.root_box[2], P.root_box[3], P.root_box[4], P.root_box[5], cur_tmin, cur_tmax);
                            if (COUNT) c_boxes++;
                            state = box_entered_exact(cur_tmin, cur_tmax) ? kStateTraverse : kStateDone;
                        }
                    }
                }
            }
            const int want = __popc(m_empty);
            u_next += want < avail ? want : avail;
            if (PUSH) {
                const int n_settled = __popc(__ballot_sync(0xffffffffu, settled));
                if (lane == 0 && n_settled) atomicSub(&s_owed[wib][u_pslot], n_settled);
            }
        }
        if (PUSH) {
            // ---- complete units leave for their final place (possibly another GPU's frame) ---------------
            __syncwarp();  // orders this warp's staged pixel stores and counter updates before the reads below
            unsigned m_flush = __ballot_sync(0xffffffffu, lane < (unsigned)kPushSlots && ((open_mask >> lane) & 1u) && s_owed[wib][lane & (kPushSlots - 1)] == 0);
            open_mask &= ~m_flush;
            while (m_flush) {
                int4 d = s_unit[wib][__ffs(m_flush) - 1];
                m_flush &= m_flush - 1u;
                const int dshift = d.z >> 8, unit_pixels = 1 << dshift;
                d.z &= 0xff;
                const int wshift = unit_wshift(dshift);  // a unit is a (1 << wshift) x (unit_pixels >> wshift) block of its tile
                const int tile = P.tile_first + d.y * P.tile_stride;
                const int bx = (P.win_tx0 + tile % P.win_tw) * kTile + d.z, by = (P.win_ty0 + tile / P.win_tw) * kTile + d.w;
                const long long lbase = (long long)d.x * P.frame_stride + ((long long)d.y * kTile + d.w) * kTile + d.z;
                const int owner = d.x % P.push_owners;
                uint32_t* const pc = P.push_bgra[owner];
                int32_t* const pi = P.push_ids[owner];
                const long long rbase = (long long)(d.x / P.push_owners) * P.W * P.H;
                for (int i = (int)lane; i < (unit_pixels >> 2); i += 32) {
                    const int row = i >> (wshift - 2), col = (i - (row << (wshift - 2))) << 2;
                    const int px = bx + col, py = by + row;
                    if (py >= P.H || px >= P.W) continue;
                    const long long lo = lbase + row * kTile + col, ro = rbase + (long long)py * P.W + px;
                    if (((P.W & 3) == 0)) {  // whole 16-byte segments (px is a multiple of 4, so px + 3 < W as well)
                        if (pc) *reinterpret_cast<uint4*>(pc + ro) = __ldcg(reinterpret_cast<const uint4*>(P.out_bgra + lo));
                        if (pi) *reinterpret_cast<int4*>(pi + ro) = __ldcg(reinterpret_cast<const int4*>(P.out_ids + lo));
                    } else {
                        for (int k = 0; k < 4 && px + k < P.W; k++) {
                            if (pc) pc[ro + k] = __ldcg(P.out_bgra + lo + k);
                            if (pi) pi[ro + k] = __ldcg(P.out_ids + lo + k);
                        }
                    }
                }
            }
            blocked = blocked && open_mask == (1u << kPushSlots) - 1u;
        }
        if (STEAL && exhausted) {
            // ---- helpers report: the pixel's owner keeps the closer hit ------------------------------------------
            unsigned m_rep = __ballot_sync(0xffffffffu, state == kStateHelperDone);
            while (m_rep) {
                const int src = __ffs(m_rep) - 1;
                m_rep &= m_rep - 1u;
                const int o = __shfl_sync(0xffffffffu, owner, src);
                const float w = __shfl_sync(0xffffffffu, best, src);
                const int t = __shfl_sync(0xffffffffu, id, src);
                const bool t_tie = __shfl_sync(0xffffffffu, (int)tie, src) != 0;
                if ((int)lane == o) {
                    if (t >= 0) {
                        if (w < best) { best = w; id = t; if (CULL) set_cull_base(); }
                        else if (w == best && id >= 0 && id != t) tie = true;
                    }
                    tie |= t_tie;
                    const int left = helpers[lane] - 1;
                    helpers[lane] = left;
                    if (state == kStateWait && left == 0) state = kStateDone;  // retired on the next round
                }
            }
            if (state == kStateHelperDone) { state = kStateEmpty; owner = (int)lane; tie = false; }
            // ---- idle lanes take over the bottom stack entries of busy lanes -------------------------------------
            const unsigned m_idle = __ballot_sync(0xffffffffu, state == kStateEmpty);
            const unsigned m_donor = __ballot_sync(0xffffffffu, state == kStateTraverse && sp > sbase && !solo);
            if (m_idle != 0u && m_donor != 0u) {
                const int pairs = min(__popc(m_idle), __popc(m_donor));
                const int my_rank = state == kStateEmpty ? __popc(m_idle & lanemask_lt) : __popc(m_donor & lanemask_lt);
                const bool helper = state == kStateEmpty && my_rank < pairs;
                const bool donor = state == kStateTraverse && sp > sbase && !solo && my_rank < pairs;
                // the entry a donor gives away: the bottom of its stack (in memory), or its only entry (the register top)
                int e_ref = 0; float e_tmin = 0.0f, e_tmax = 0.0f;
                if (donor) {
                    while (CULL && sp - sbase >= 2 && culled(stk_tmin[sbase])) sbase++;  // dead entries: a pop would drop them too
                    if (sp - sbase >= 2) { e_ref = stk_ref[sbase]; e_tmin = stk_tmin[sbase]; e_tmax = stk_tmax[sbase]; sbase++; }
                    else { e_ref = top_ref; e_tmin = top_tmin; e_tmax = top_tmax; sp--; }
                }
                const int src = helper ? (int)__fns(m_donor, 0, my_rank + 1) : (int)lane;
                const int n_ref = __shfl_sync(0xffffffffu, e_ref, src);
                const float n_tmin = __shfl_sync(0xffffffffu, e_tmin, src), n_tmax = __shfl_sync(0xffffffffu, e_tmax, src);
                const int n_owner = __shfl_sync(0xffffffffu, owner, src);
                const float n_best = __shfl_sync(0xffffffffu, best, src), n_slack = __shfl_sync(0xffffffffu, slack_abs, src);
                const int n_id = __shfl_sync(0xffffffffu, id, src);  // (kept to recognise ties with the donor's hit, see moller_trumbore)
                Ray q;
                q.dx = __shfl_sync(0xffffffffu, r.dx, src); q.dy = __shfl_sync(0xffffffffu, r.dy, src); q.dz = __shfl_sync(0xffffffffu, r.dz, src);
                q.ix = __shfl_sync(0xffffffffu, r.ix, src); q.iy = __shfl_sync(0xffffffffu, r.iy, src); q.iz = __shfl_sync(0xffffffffu, r.iz, src);
                q.fx = __shfl_sync(0xffffffffu, r.fx, src); q.fy = __shfl_sync(0xffffffffu, r.fy, src); q.fz = __shfl_sync(0xffffffffu, r.fz, src);
                q.ox = __shfl_sync(0xffffffffu, r.ox, src); q.oy = __shfl_sync(0xffffffffu, r.oy, src); q.oz = __shfl_sync(0xffffffffu, r.oz, src);
                if (helper) {
                    r = q;
                    best = n_best; slack_abs = n_slack; id = n_id; tie = false;
                    set_cull_base();
                    owner = n_owner;
                    sp = 0; sbase = 0;
                    cur = n_ref; cur_tmin = n_tmin; cur_tmax = n_tmax;
                    want_pop = culled(n_tmin);  // the donor's best may have improved since the entry was pushed
                    state = kStateTraverse;
                    atomicAdd(&s_helpers[wib][n_owner], 1);
#ifdef RTB_WARP_LOG
                    wl_steals++;
#endif
                }
                __syncwarp();
            }
            if (__ballot_sync(0xffffffffu, state != kStateEmpty) == 0u) break;  // nothing in flight, nothing left to fetch
            // (a bound on the rounds of a draining warp -- orders of magnitude above any real walk -- so that a mistake in the
            // sharing logic shows up as a wrong frame in the tests instead of a kernel that never ends)
            if (++rounds > (1 << 22)) break;
        } else if (exhausted && m_trav == 0u) break;  // nothing in flight, nothing left to fetch
    }
#ifdef RTB_WARP_LOG
    {
        unsigned long long wl_t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(wl_t1));
        wl_rays = __reduce_add_sync(0xffffffffu, wl_rays);
        wl_steals = __reduce_add_sync(0xffffffffu, wl_steals);
        if (lane == 0 && P.counters) {  // log lives behind the counters: 8 words per warp from counters[16]
            unsigned long long* L = P.counters + 16 + 8ull * ((unsigned long long)blockIdx.x * (kBlockThreads / 32) + wib);
            unsigned smid;
            asm("mov.u32 %0, %smid;" : "=r"(smid));
            L[0] = wl_t0; L[1] = wl_t1; L[2] = wl_first_work; L[3] = wl_exhausted;
            L[4] = ((unsigned long long)wl_units << 32) | wl_bg_units; L[5] = ((unsigned long long)wl_iters << 32) | wl_rays; L[6] = smid; L[7] = ((unsigned long long)(unsigned)rounds << 32) | wl_steals;
        }
    }
#endif

    if (COUNT) {
        for (int s = 16; s > 0; s >>= 1) {
            c_rays += __shfl_down_sync(0xffffffffu, c_rays, s);
            c_nodes += __shfl_down_sync(0xffffffffu, c_nodes, s);
            c_boxes += __shfl_down_sync(0xffffffffu, c_boxes, s);
            c_tris += __shfl_down_sync(0xffffffffu, c_tris, s);
            c_hits += __shfl_down_sync(0xffffffffu, c_hits, s);
            c_depth += __shfl_down_sync(0xffffffffu, c_depth, s);
            c_depth_max = max(c_depth_max, __shfl_down_sync(0xffffffffu, c_depth_max, s));
            c_steps_max = max(c_steps_max, __shfl_down_sync(0xffffffffu, c_steps_max, s));
        }
        if (lane == 0) {
            atomicAdd(P.counters + 0, c_rays); atomicAdd(P.counters + 1, c_nodes); atomicAdd(P.counters + 2, c_boxes);
            atomicAdd(P.counters + 3, c_tris); atomicAdd(P.counters + 4, c_hits);
            atomicAdd(P.counters + 5, c_depth); atomicMax(P.counters + 6, (unsigned long long)c_depth_max);
            atomicMax(P.counters + 7, (unsigned long long)c_steps_max);
        }
    }
}

}  // namespace rtb
