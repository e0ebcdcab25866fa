// rtb_api.cu -- the C ABI declared in include/rtb.h: handles, device memory, launches.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtb.h"
#include "rtb_host.hpp"
#include "rtb_kernels.cuh"
#include "rtb_render.cuh"
#include "rtb_scene.cuh"

namespace {

std::atomic<uint64_t> g_launches{0};
#ifdef RTB_WARP_LOG
constexpr size_t kCounterWords = 16 + 8 * 8192;  // development builds: 8 words per warp behind the counters (rtb_camera_warp_log)
#else
constexpr size_t kCounterWords = 8;
#endif
thread_local std::string g_error;
thread_local int g_device = 0;

int fail(rtb_status code, const std::string& what) {
    g_error = what;
    return (int)code;
}
#define RTB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) return fail(RTB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// include/rtb.h promises that no entry point throws: the ones that grow host containers from caller- or file-controlled
// sizes run their body through this guard.
template <class F>
int guarded(const char* who, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(RTB_ERR_NOMEM, std::string(who) + ": out of host memory");
    } catch (const std::exception& e) {
        return fail(RTB_ERR_ARG, std::string(who) + ": " + e.what());
    } catch (...) {
        return fail(RTB_ERR_ARG, std::string(who) + ": unexpected exception");
    }
}

template <class T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

}  // namespace

struct rtb_mesh {
    int device = 0;
    int64_t n = 0;
    std::vector<float> points;  // 9 per triangle (Trixel::h_points_init_data)
    std::vector<float> rad;     // 3 per triangle, or empty when uniform
    float uniform_rgb[3] = {0.1f, 0.55f, 0.2f};
    float* d_points = nullptr;  // Trixel::d_points_init_data
    rtb::HostTree tree;     // host copy; for a device build it is filled on first use (ensure_host_tree)
    rtb::DeviceTree dtree;  // device build: the tree stays in HBM and feeds the pack kernels directly
    bool built = false;
    bool built_on_device = false;
    bool host_tree_valid = false;
    uint64_t generation = 0;  // bumped by every build / load: camera-side arrays made from an older tree are not reused
    double seconds_sort = 0, seconds_partition = 0;
};

// Camera::voxel_memory + Camera::trixel_memory (Camera.h:64-84) of one (camera, mesh, tree): the camera-relative node
// and triangle records.  Objects that instance the same mesh on the same camera (WinMain.cpp:152-156 registers two)
// share one copy -- in the reference the second add_object overwrites the first one's identical arrays (Camera.cpp:156,206).
struct SceneArrays {
    const rtb_mesh* mesh = nullptr;  // identity only (never dereferenced after creation)
    uint64_t generation = 0;
    int refs = 0;
    float4* d_scene = nullptr;  // nodes then triangles in ONE allocation
    float4* d_nodes = nullptr;
    float4* d_tris = nullptr;
    float4* d_rad = nullptr;
    size_t scene_bytes = 0, node_bytes = 0;
    void* l2_window_base = nullptr;  // persisting access-policy window attached to every render launch (0 bytes = off)
    size_t l2_window_bytes = 0;
    float l2_hit_ratio = 0.0f;
    float root_box[6] = {0, 0, 0, 0, 0, 0};
    int root_ref = 0;
    float uniform_rgb[3] = {0, 0, 0};
    int64_t num_tri = 0;
};

// Single-frame path (rtb_object_render -> rtb_camera_color_pixels): a frame is rendered into a SLOT -- a pinned host frame the
// kernel stores into directly, a device staging area for its work units, a work counter, a stream and an event of its own --
// so that frames can be in flight side by side.  A single 960x540 frame cannot fill a B200: it lasts as long as its longest
// rays while most of the GPU idles.  When the caller's motion repeats (the same transform steps before
// every render: the reference's key-held orbit, WinMain.cpp:186-213) the frames after the current one are rendered ahead, on
// other slots' streams, in the shadow of the current frame's tail; a render call whose matrix equals the predicted one
// bit for bit finds its frame already on the way.  A wrong guess costs idle-GPU work and is simply never looked at.
constexpr int kFrameSlots = 4;
struct FrameSlot {
    uint32_t* h_bgra = nullptr;      // slot 0 uses the camera's own host frame
    int32_t* h_ids = nullptr;
    uint32_t* stage_bgra = nullptr;  // tile-major staging of one frame (device)
    int32_t* stage_ids = nullptr;
    unsigned long long* d_work = nullptr;
    unsigned long long work_base = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool ready = false;        // allocated
    bool in_flight = false;    // a launch into this slot has not been waited for yet
    bool rect_valid = false;   // the host frame is background / -1 outside rect (the frame last stored there)
    int rect[4] = {0, 0, -1, -1};
    float m12[12] = {0};
    uint32_t flags = 0;
};

struct rtb_camera {
    int device = 0;
    rtb::CameraBasis basis;
    long long pixels = 0;
    uint32_t* d_bgra = nullptr;  // Camera::h_mem.d_color.c
    int32_t* d_ids = nullptr;    // Camera::pixel_memory::d_rmi
    uint32_t* h_bgra = nullptr;  // Camera::h_mem.h_color.c (pinned)
    int32_t* h_ids = nullptr;    // pinned
    unsigned long long* d_counters = nullptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_render[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    // sweep ring: two chunks of device frames + pinned staging
    uint32_t* ring_bgra[2] = {nullptr, nullptr};
    int32_t* ring_ids[2] = {nullptr, nullptr};
    uint32_t* stage_bgra[2] = {nullptr, nullptr};
    int32_t* stage_ids[2] = {nullptr, nullptr};
    int ring_frames = 0;
    // peer push: this GPU's tile-major staging of the frames in flight (rtb_render_frames_push_async)
    uint32_t* push_bgra = nullptr;
    int32_t* push_ids = nullptr;
    size_t push_elements = 0;
    std::vector<SceneArrays*> scenes;   // one per (mesh, tree) in use on this camera
    std::vector<rtb_object*> objects;   // every object currently added to this camera (Camera::object_list)
    int sm_count = 0;
    int blocks_per_sm[16] = {0};  // occupancy of the render kernel variants, asked once
    // scene extension (rtb_camera_render_scene): the light list, the shadow test and sample_rate the reference leaves dormant
    int num_lights = 1;
    float lights[rtb::kMaxSceneLights][3] = {{2.0f, 2.0f, 2.0f}};  // Camera.cu:32
    bool shadows = false;
    int sample_rate = 0;          // Camera::render_properites::sample_rate (Camera.h:46, Camera.cpp:71)
    bool frame_rendered = false;  // the device frame holds a render (background + shaded hits of every pixel)
    bool frame_on_host = false;   // ... and h_bgra / h_ids hold that very frame already
    // Single-frame path (rtb_object_render): the kernel stores the frame straight into h_bgra / h_ids (pinned, mapped),
    // and only where it can differ from what they hold -- host_rect is the pixel rectangle outside which the host frame
    // is known to be background / -1 (the root-box rectangle of the frame last stored there this way).
    FrameSlot sweep_lane[2];          // rtb_render_sweep into pinned buffers: two streams with their own staging and work counter, so
    size_t sweep_lane_elements = 0;   // that the tail of one chunk's launch overlaps the start of the next (only stream, d_work, stage_* used)
    FrameSlot slots[kFrameSlots];
    int cur_slot = 0;                 // the slot whose host frame rtb_camera_host_color / _ids name
    std::deque<int> ahead;            // slots rendering predicted frames, oldest first
    rtb_object* ahead_obj = nullptr;  // ... of this object,
    uint32_t ahead_flags = 0;         // ... with these render flags;
    rtb::Transform ahead_xf;          // the recurrence's state after the last predicted frame
    int ahead_misses = 0;             // predictions dropped in a row (a caller that alternates objects or jitters: stop guessing)
    int ahead_pause = 0;              // renders left before guessing is tried again
    bool device_frame_stale = false;  // the last frame went to the host only: d_bgra / d_ids hold an older one
};

struct rtb_object {
    rtb_mesh* mesh = nullptr;
    rtb_camera* cam = nullptr;
    int device = 0;
    rtb::Transform xf;
    bool xf_ready = false;
    bool xf_overridden = false;  // rtb_object_set_matrix replaced the rows: the recurrence's quaternion / faces no longer describe them
    SceneArrays* scene = nullptr;
    // per-object launch scratch: frame records (matrices + pixel rectangles) of the frames in flight and the work
    // counter of the persistent kernel.  Launches of one object are ordered on the device through ev_launch, whatever
    // streams they are issued on; the pinned staging is reused only after ev_upload (the previous upload) has passed.
    float* d_frames = nullptr;
    float* h_frames = nullptr;  // pinned
    int frames_capacity = 0;
    unsigned long long* d_work = nullptr;
    unsigned long long work_base = 0;  // value of *d_work once everything issued so far has run
    cudaEvent_t ev_upload = nullptr, ev_launch = nullptr;
    cudaStream_t last_stream = nullptr;
    bool upload_pending = false, launch_pending = false;
    // the transform steps applied since the last single-frame render and before the one before it: when they repeat, the
    // next frames are predictable (rtb_camera::ahead)
    struct Step { uint8_t select; float v[4]; };
    std::vector<Step> steps_now, steps_frame;
    bool steps_repeat = false;
};

namespace {

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// Scheduling knobs of the render kernel (rtb_set_knob).  Initialised once from the environment (RTB_UNIT_SHIFT, ...),
// so that a launch costs no getenv; tests and the tuning tools change them through rtb_set_knob.
struct Knobs {
    int unit_shift = 0;   // log2 pixels per work unit (5..10), 0 = chosen from the launch size
    int t_active = 4;     // refill when no more than this many lanes still traverse   } re-tuned in round 2 with view-ordered frames
    int t_leaf = 4;       // leaf step when at least this many lanes wait at a leaf    } (12 / 8 before: DESIGN.md section 4)
    int tail5 = -1, tail6 = -1;  // trailing frames of a launch in 32- / 64-pixel units, -1 = automatic
    int reserve_sms = 0;  // SMs left free beside the persistent kernel
    int l2_window = 1;    // persisting L2 window: 0 off, 1 node records, 2 nodes + triangles (takes effect at add_object)
    int no_rect = 0;      // 1: frame records carry the whole frame as the root-box rectangle
    int frame_order = 1;  // multi-frame launches work through their frames sorted by viewing direction (order_frames)
    int host_direct = 1;  // rtb_object_render stores its frame straight into the camera's host buffers (0: device frame + copy)
    int sweep_direct = 1; // rtb_render_sweep into pinned caller buffers: host threads pre-fill the background, the kernel stores the rest
    int host_fill_threads = 0;  // threads of that pre-fill, 0 = min(12, CPUs this process may run on)
    int t_active_inline = 30;   // single-frame launches: t_active (a warp that sees the empty queue early shares its long rays early)
    int steal_spin = 8;         // ... and how many iterations a draining warp runs between two looks for lanes to share with (power of two)
    int inline_prefetch = 1;    // ... and whether it asks for both children's records ahead of the decision
    int fill_ctas_per_sm = 2;   // rtb_fill_frames_device_async: thread blocks per SM of the fill kernel (it is meant to run beside a render)
    int lookahead = 2;          // single-frame path: predicted frames kept in flight behind the current one (0 = off)
    int sweep_chunk_mb = 256;   // ... and the largest size of the chunks (MB of host frames) in which fill and render alternate
    int l2_carve_mb = 0;  // persisting L2 carve-out in MB, 0 = the size of the window
};
Knobs& knobs() {
    static Knobs k = [] {
        Knobs v;
        v.unit_shift = env_int("RTB_UNIT_SHIFT", v.unit_shift); v.t_active = env_int("RTB_T_ACTIVE", v.t_active);
        v.t_leaf = env_int("RTB_T_LEAF", v.t_leaf); v.tail5 = env_int("RTB_TAIL5", v.tail5); v.tail6 = env_int("RTB_TAIL6", v.tail6);
        v.reserve_sms = env_int("RTB_RESERVE_SMS", v.reserve_sms); v.l2_window = env_int("RTB_L2_WINDOW", v.l2_window);
        v.no_rect = env_int("RTB_NO_RECT", v.no_rect); v.frame_order = env_int("RTB_FRAME_ORDER", v.frame_order);
        v.host_direct = env_int("RTB_HOST_DIRECT", v.host_direct); v.sweep_direct = env_int("RTB_SWEEP_DIRECT", v.sweep_direct);
        v.host_fill_threads = env_int("RTB_HOST_FILL_THREADS", v.host_fill_threads); v.sweep_chunk_mb = env_int("RTB_SWEEP_CHUNK_MB", v.sweep_chunk_mb);
        v.t_active_inline = env_int("RTB_T_ACTIVE_INLINE", v.t_active_inline); v.steal_spin = env_int("RTB_STEAL_SPIN", v.steal_spin);
        v.inline_prefetch = env_int("RTB_INLINE_PREFETCH", v.inline_prefetch); v.lookahead = env_int("RTB_LOOKAHEAD", v.lookahead);
        v.fill_ctas_per_sm = env_int("RTB_FILL_CTAS_PER_SM", v.fill_ctas_per_sm); v.l2_carve_mb = env_int("RTB_L2_CARVE_MB", v.l2_carve_mb);
        return v;
    }();
    return k;
}

int host_fill_threads() {
    int t = knobs().host_fill_threads;
    if (t > 0) return t;
    cpu_set_t set;
    CPU_ZERO(&set);
    t = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : 1;
    return std::max(1, std::min(12, t));  // measured on the pool's hosts: 8 threads already reach the memory system's limit
}

int ensure_host_tree(rtb_mesh* m) {
    if (m->host_tree_valid) return RTB_OK;
    if (!m->built_on_device) return fail(RTB_ERR_STATE, "tree not built");
    RTB_CUDA(cudaSetDevice(m->device));
    const std::string err = rtb::download_tree(m->dtree, m->tree);
    if (!err.empty()) return fail(RTB_ERR_CUDA, err);
    m->host_tree_valid = true;
    return RTB_OK;
}

// the object's launch scratch exists (work counter, events)
int ensure_object_scratch(rtb_object* o) {
    if (!o->d_work) {
        RTB_CUDA(cudaMalloc(&o->d_work, sizeof(unsigned long long)));
        RTB_CUDA(cudaMemset(o->d_work, 0, sizeof(unsigned long long)));
        o->work_base = 0;
    }
    if (!o->ev_upload) RTB_CUDA(cudaEventCreateWithFlags(&o->ev_upload, cudaEventDisableTiming));
    if (!o->ev_launch) RTB_CUDA(cudaEventCreateWithFlags(&o->ev_launch, cudaEventDisableTiming));
    return RTB_OK;
}

// Room for `frames` frame records; returns with the pinned staging free to be overwritten.
int ensure_frames(rtb_object* o, int frames) {
    if (o->upload_pending) {  // the previous upload still reads h_frames until its event has passed (a few microseconds)
        RTB_CUDA(cudaEventSynchronize(o->ev_upload));
        o->upload_pending = false;
    }
    if (frames <= o->frames_capacity) return RTB_OK;
    if (o->launch_pending) {  // a kernel in flight may still read the old d_frames
        RTB_CUDA(cudaEventSynchronize(o->ev_launch));
        o->launch_pending = false;
    }
    int cap = std::max(frames, 64);
    if (o->d_frames) cudaFree(o->d_frames);
    if (o->h_frames) cudaFreeHost(o->h_frames);
    o->d_frames = nullptr; o->h_frames = nullptr; o->frames_capacity = 0;
    // (one more word per frame: the processing order of a multi-frame launch follows the records, see order_frames)
    RTB_CUDA(cudaMalloc(&o->d_frames, sizeof(float) * (rtb::kFrameStride + 1) * (size_t)cap));
    RTB_CUDA(cudaMallocHost(&o->h_frames, sizeof(float) * (rtb::kFrameStride + 1) * (size_t)cap));
    o->frames_capacity = cap;
    return RTB_OK;
}

// Device-side order between launches of one object that arrive on different streams: they share d_frames, the work
// counter and (push variant) the camera's tile staging.  Same stream: nothing to do.
int order_after_previous(rtb_object* o, cudaStream_t s) {
    if (o->launch_pending && o->last_stream != s) RTB_CUDA(cudaStreamWaitEvent(s, o->ev_launch, 0));
    return RTB_OK;
}

int upload_frames(rtb_object* o, int num_frames, cudaStream_t s, bool with_order = false) {
    RTB_CUDA(cudaMemcpyAsync(o->d_frames, o->h_frames, sizeof(float) * (rtb::kFrameStride + (with_order ? 1 : 0)) * (size_t)num_frames,
                             cudaMemcpyHostToDevice, s));
    RTB_CUDA(cudaEventRecord(o->ev_upload, s));
    o->upload_pending = true;
    return RTB_OK;
}

// One frame record for the kernel: the object's 3x4 matrix followed by the pixel rectangle outside
// which no primary ray can reach the root box.  The rectangle is the projection of the eight corners
// of the (camera-relative, translated) root box through the inverse rotation onto the pixel grid,
// enlarged by 2 pixels -- three orders of magnitude more than the rounding error of the slab test
// (Trixel.cu:76-95,146) -- so pixels outside it are background exactly as in the reference.  If a
// corner is not in front of the camera, or the matrix is not invertible, the rectangle is the frame.
void fill_frame_record(const SceneArrays* sc, const rtb_camera* c, const float m12[12], float* rec) {
    std::memcpy(rec, m12, sizeof(float) * 12);
    const rtb::CameraBasis& b = c->basis;
    int rect[4] = {0, 0, b.W - 1, b.H - 1};
    auto dot = [](const double* p, const float* q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
    auto dotf = [](const float* p, const float* q) { return (double)p[0] * q[0] + (double)p[1] * q[1] + (double)p[2] * q[2]; };
    const double R[3][3] = {{m12[0], m12[1], m12[2]}, {m12[4], m12[5], m12[6]}, {m12[8], m12[9], m12[10]}};
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    const double pw = dotf(b.u_mod, b.u), ph = dotf(b.v_mod, b.v), f = dotf(b.n_mod, b.n);
    bool ok = std::isfinite(det) && std::fabs(det) > 1e-6 && pw > 0 && ph > 0 && f > 0 && sc->root_ref >= 0;
    if (ok) {
        double inv[3][3];
        inv[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det; inv[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det; inv[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det;
        inv[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det; inv[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det; inv[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
        inv[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det; inv[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det; inv[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
        const double ax = -dotf(b.n_mod, b.u) / pw, ay = -dotf(b.n_mod, b.v) / ph;
        double lo_x = 1e300, hi_x = -1e300, lo_y = 1e300, hi_y = -1e300;
        for (int k = 0; k < 8 && ok; k++) {
            const double cx = (double)sc->root_box[(k & 1) ? 3 : 0] + m12[3];
            const double cy = (double)sc->root_box[(k & 2) ? 4 : 1] + m12[7];
            const double cz = (double)sc->root_box[(k & 4) ? 5 : 2] + m12[11];
            const double p[3] = {inv[0][0] * cx + inv[0][1] * cy + inv[0][2] * cz, inv[1][0] * cx + inv[1][1] * cy + inv[1][2] * cz,
                                 inv[2][0] * cx + inv[2][1] * cy + inv[2][2] * cz};
            const double depth = dot(p, b.n);
            const double scale = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            if (!(depth > 1e-3 * scale) || !std::isfinite(depth)) { ok = false; break; }
            const double ix = ax + f * dot(p, b.u) / (depth * pw), iy = ay + f * dot(p, b.v) / (depth * ph);
            if (!std::isfinite(ix) || !std::isfinite(iy)) { ok = false; break; }
            lo_x = std::min(lo_x, ix); hi_x = std::max(hi_x, ix); lo_y = std::min(lo_y, iy); hi_y = std::max(hi_y, iy);
        }
        if (ok) {
            const double margin = 2.0;
            auto clampi = [](double v, int lo, int hi) { return (int)std::max((double)lo, std::min((double)hi, v)); };
            rect[0] = clampi(std::floor(lo_x - margin), 0, b.W); rect[1] = clampi(std::floor(lo_y - margin), 0, b.H);
            rect[2] = clampi(std::ceil(hi_x + margin), -1, b.W - 1); rect[3] = clampi(std::ceil(hi_y + margin), -1, b.H - 1);
        }
    }
    if (knobs().no_rect) { rect[0] = 0; rect[1] = 0; rect[2] = b.W - 1; rect[3] = b.H - 1; }
    std::memcpy(rec + 12, rect, sizeof rect);
}

// Processing order of the frames of one launch.  The frames of a launch are independent, so the order in which the
// persistent kernel works through them is free; what is in flight at any moment is about one frame's worth of work
// units, and the node and triangle records a frame touches are re-used by the next one only if it shows the object from
// nearly the same side.  An orbit in index order turns the object by the step angle from frame to frame and comes
// back to a view only a revolution (31 frames, ~100 MB of records) later; sorted by viewing direction, neighbours in
// the order differ by a fraction of a degree and find each other's records in L1/L2.  Key: direction of the camera
// axis in object space (R^T n), as azimuth inside eight elevation bands walked in alternating direction.
// Written behind the records in the pinned staging (h_frames + kFrameStride * num_frames); returns false (and
// writes nothing) when the launch is too small to gain or the knob is off.
bool order_frames(const rtb_object* o, const rtb_camera* c, int num_frames) {
    if (!knobs().frame_order || num_frames < 4) return false;
    const float* rec = o->h_frames;
    int* order = reinterpret_cast<int*>(o->h_frames + rtb::kFrameStride * (size_t)num_frames);
    const float* n = c->basis.n;
    std::vector<std::pair<double, int>> key((size_t)num_frames);
    const double kPi = 3.14159265358979323846;
    for (int f = 0; f < num_frames; f++) {
        const float* m = rec + rtb::kFrameStride * (size_t)f;
        const double dx = (double)m[0] * n[0] + (double)m[4] * n[1] + (double)m[8] * n[2];
        const double dy = (double)m[1] * n[0] + (double)m[5] * n[1] + (double)m[9] * n[2];
        const double dz = (double)m[2] * n[0] + (double)m[6] * n[1] + (double)m[10] * n[2];
        const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
        double az = 0.0, el = 0.0;
        if (std::isfinite(len) && len > 0.0) { az = std::atan2(dz, dx); el = std::asin(std::max(-1.0, std::min(1.0, dy / len))); }
        const int band = std::max(0, std::min(7, (int)((el + kPi / 2) / kPi * 8.0)));
        key[(size_t)f] = {band * 8.0 + ((band & 1) ? -az : az), f};  // |az| <= pi < 4: bands never overlap
    }
    std::sort(key.begin(), key.end());
    for (int f = 0; f < num_frames; f++) order[f] = key[(size_t)f].second;
    return true;
}

// Launch the persistent render kernel over `num_frames` frame records: resident in device memory at `d_frames`, or --
// single frame -- handed over as `inline_record` and carried in the kernel's parameters (no upload, d_frames == nullptr).
int launch_render(rtb_object* o, rtb_camera* c, const float* d_frames, const float* inline_record, int num_frames, int tile_first,
                  int tile_stride, uint32_t flags, uint32_t* d_bgra, int32_t* d_ids, cudaStream_t stream, int push_owners = 0,
                  uint32_t* const* push_bgra = nullptr, int32_t* const* push_ids = nullptr, const int* d_frame_order = nullptr,
                  const int* push_prev_rect = nullptr, const int* tile_window = nullptr, FrameSlot* slot = nullptr) {
    using namespace rtb;
    const SceneArrays* sc = o->scene;
    RenderParams P;
    std::memset(&P, 0, sizeof P);
    const CameraBasis& b = c->basis;
    P.W = b.W; P.H = b.H;
    for (int k = 0; k < 3; k++) { P.n_mod[k] = b.n_mod[k]; P.u_mod[k] = b.u_mod[k]; P.v_mod[k] = b.v_mod[k]; }
    for (int k = 0; k < 6; k++) P.root_box[k] = sc->root_box[k];
    P.draw_distance = b.draw_distance;
    P.background = ((uint32_t)b.background[3] << 24) | ((uint32_t)b.background[0] << 16) | ((uint32_t)b.background[1] << 8) | b.background[2];
    P.root_ref = sc->root_ref;
    P.nodes = sc->d_nodes; P.tris = sc->d_tris; P.rad = sc->d_rad;
    for (int k = 0; k < 3; k++) P.uniform_rad[k] = sc->uniform_rgb[k];
    P.frames = d_frames;
    P.frame_order = d_frame_order;
    if (inline_record) std::memcpy(P.frame0, inline_record, sizeof P.frame0);
    P.num_frames = num_frames;
    P.tiles_x = (b.W + kTile - 1) / kTile;
    const int tiles_y = (b.H + kTile - 1) / kTile;
    P.win_tx0 = 0; P.win_ty0 = 0; P.win_tw = P.tiles_x;
    int win_th = tiles_y;
    if (tile_window) {  // tx0, ty0, tx1, ty1 (inclusive); empty when tx1 < tx0
        P.win_tx0 = std::max(0, tile_window[0]); P.win_ty0 = std::max(0, tile_window[1]);
        P.win_tw = std::max(0, std::min(P.tiles_x - 1, tile_window[2]) - P.win_tx0 + 1);
        win_th = std::max(0, std::min(tiles_y - 1, tile_window[3]) - P.win_ty0 + 1);
        if (P.win_tw == 0 || win_th == 0) return RTB_OK;  // nothing can change
    }
    const int tiles = P.win_tw * win_th;
    if (tile_stride < 1 || tile_first < 0 || tile_first >= tile_stride) return fail(RTB_ERR_ARG, "render: bad tile_first/tile_stride");
    P.tile_first = tile_first; P.tile_stride = tile_stride;
    P.my_tiles = tile_first < tiles ? (tiles - tile_first + tile_stride - 1) / tile_stride : 0;
    P.total_items = (long long)num_frames * P.my_tiles;
    P.out_bgra = d_bgra; P.out_ids = d_ids;
    const bool push = push_owners > 0;
    P.push_owners = std::max(push_owners, 1);
    for (int k = 0; k < push_owners; k++) { P.push_bgra[k] = push_bgra ? push_bgra[k] : nullptr; P.push_ids[k] = push_ids ? push_ids[k] : nullptr; }
    P.push_skip_background = (push && (flags & RTB_RENDER_PUSH_PREFILLED)) ? 1 : 0;
    if (push && push_prev_rect) {
        P.push_skip_background = 2;
        std::memcpy(P.push_prev_rect, push_prev_rect, sizeof P.push_prev_rect);
    }
    P.tile_major = ((flags & RTB_RENDER_TILE_MAJOR) || push) ? 1 : 0;
    P.frame_stride = P.tile_major ? (long long)((tiles + tile_stride - 1) / tile_stride) * kTile * kTile : (long long)b.W * b.H;
    P.work_counter = slot ? slot->d_work : o->d_work;  // a slot's launches share nothing with the object's other launches
    P.work_base = slot ? slot->work_base : o->work_base;
    P.counters = c->d_counters;
    P.cull_rel = 1e-5f;
    if (P.total_items == 0) return RTB_OK;

    const bool cull = !(flags & RTB_RENDER_NO_CULL), count = (flags & RTB_RENDER_COUNTERS) != 0;
    const bool inl = inline_record != nullptr && !count;
    if (inline_record && !inl) return fail(RTB_ERR_ARG, "render: inline frame records are for single-frame launches without counters");
    void (*kern)(const RenderParams);
    int variant;
    if (push && inl) { kern = cull ? render_stream_kernel<true, false, true, true> : render_stream_kernel<false, false, true, true>; variant = 8 + (cull ? 1 : 0); }
    else if (push) { kern = cull ? render_stream_kernel<true, false, true> : render_stream_kernel<false, false, true>; variant = 4 + (cull ? 1 : 0); }
    else if (inl) { kern = cull ? render_stream_kernel<true, false, false, true> : render_stream_kernel<false, false, false, true>; variant = 6 + (cull ? 1 : 0); }
    else {
        kern = cull ? (count ? render_stream_kernel<true, true, false> : render_stream_kernel<true, false, false>)
                    : (count ? render_stream_kernel<false, true, false> : render_stream_kernel<false, false, false>);
        variant = (cull ? 1 : 0) + (count ? 2 : 0);
    }
    if (c->blocks_per_sm[variant] == 0) {
        int per_sm = 0;
        RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, 0));
        c->blocks_per_sm[variant] = std::max(per_sm, 1);
    }
    const int per_sm = c->blocks_per_sm[variant];
    const long long warps_total = (long long)c->sm_count * per_sm * (kBlockThreads / 32);
    // work unit: Morton block of 2^shift pixels; small launches get small units so every warp has work
    const long long pixels = (long long)num_frames * P.my_tiles * kTile * kTile;
    // measured on the dragon stand-in: 256-pixel units for launches that give every warp a hundred of them or more (600 frames at
    // 960x540, 36 at 4K), 128-pixel units below that (60-frame launches: 1.41 vs 1.59 ms), smaller ones for small launches
    // (the same holds for the push variant in tile mode: tools/push_tune.py)
    int shift = (pixels >> 8) >= warps_total * 128 ? 8 : 7;
    while (shift > 5 && (pixels >> shift) < warps_total * 4) shift--;
    const Knobs& K = knobs();
    P.unit_shift = std::min(10, std::max(5, K.unit_shift > 0 ? K.unit_shift : shift));
    P.t_active = std::min(31, std::max(0, inl ? K.t_active_inline : K.t_active));
    {
        int spin = 1;
        while (spin < K.steal_spin && spin < 64) spin <<= 1;
        P.steal_mask = spin - 1;
    }
    P.prefetch = K.inline_prefetch ? 1 : 0;
    P.t_leaf = std::max(1, K.t_leaf);
    // Segments of decreasing unit size towards the end of the launch (see RenderParams::seg_*): a unit handed out late
    // is worked through by ONE warp while the queue is already empty, so late units must be small.  Measured on the
    // dragon stand-in (60 frames): see DESIGN.md.  RTB_TAIL6 / RTB_TAIL5 = number of trailing frames in 64- / 32-pixel units.
    {
        int t5 = 0, t6 = 0;
        if (P.unit_shift > 5 && num_frames >= 8) t5 = std::min(8, std::max(1, num_frames / 20));  // measured: 3 of 60 frames, 0-16 of 600 alike; a 64-pixel segment adds nothing
        const int e5 = K.tail5, e6 = K.tail6;
        t5 = std::min(num_frames, std::max(0, e5 >= 0 ? e5 : t5));
        t6 = std::min(num_frames - t5, std::max(0, e6 >= 0 ? e6 : t6));
        P.seg_frames[0] = num_frames - t6 - t5; P.seg_shift[0] = P.unit_shift;
        P.seg_frames[1] = t6;                   P.seg_shift[1] = std::min(P.unit_shift, 6);
        P.seg_frames[2] = t5;                   P.seg_shift[2] = 5;
        P.total_items = 0;
        for (int k = 0; k < 3; k++) {
            P.seg_items[k] = (long long)P.seg_frames[k] * P.my_tiles * ((kTile * kTile) >> P.seg_shift[k]);
            P.total_items += P.seg_items[k];
        }
        if (P.seg_frames[0] == 0) {  // keep segment 0 non-empty: the kernel starts from its unit size
            P.seg_frames[0] = P.seg_frames[1]; P.seg_shift[0] = P.seg_shift[1]; P.seg_items[0] = P.seg_items[1];
            P.seg_frames[1] = 0; P.seg_items[1] = 0;
        }
    }
    const long long fetches = P.total_items;
    const long long blocks_needed = (fetches + (kBlockThreads / 32) - 1) / (kBlockThreads / 32);
    // RTB_RESERVE_SMS leaves SMs free for kernels that must run beside this persistent one (the NCCL
    // gather of the multi-GPU tile exchange); the default keeps the whole GPU.
    const int sms = std::max(1, c->sm_count - std::max(0, K.reserve_sms));
    const int grid = (int)std::max(1ll, std::min<long long>((long long)sms * per_sm, blocks_needed));
    // per-launch L2 policy: the scene's node records persist, everything else streams.  Set on the launch so that it
    // also holds on streams the caller owns.
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kBlockThreads); cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (sc->l2_window_bytes > 0) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = sc->l2_window_base;
        attr[0].val.accessPolicyWindow.num_bytes = sc->l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = sc->l2_hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    if (!slot) {
        const int rc = order_after_previous(o, stream);
        if (rc) return rc;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RTB_ERR_CUDA, std::string("render launch: ") + cudaGetErrorString(e));
    }
    g_launches++;
    // every warp of the grid fetches until its first miss: the counter ends at base + units + warps
    const unsigned long long handed_out = (unsigned long long)P.total_items + (unsigned long long)grid * (kBlockThreads / 32);
    if (slot) {
        slot->work_base += handed_out;
        return RTB_OK;
    }
    o->work_base += handed_out;
    RTB_CUDA(cudaEventRecord(o->ev_launch, stream));
    o->launch_pending = true;
    o->last_stream = stream;
    return RTB_OK;
}

// ---- frame slots of the single-frame path ---------------------------------------------------------------------------
int ensure_slot(rtb_camera* c, int k) {
    FrameSlot& s = c->slots[k];
    if (s.ready) return RTB_OK;
    const size_t P = (size_t)c->pixels, stage = (size_t)rtb_tile_major_elements(c, 1);
    if (k == 0) { s.h_bgra = c->h_bgra; s.h_ids = c->h_ids; }
    else {
        if (!s.h_bgra) RTB_CUDA(cudaMallocHost(&s.h_bgra, 4 * P));
        if (!s.h_ids) RTB_CUDA(cudaMallocHost(&s.h_ids, 4 * P));
    }
    if (!s.stage_bgra) RTB_CUDA(cudaMalloc(&s.stage_bgra, 4 * stage));
    if (!s.stage_ids) RTB_CUDA(cudaMalloc(&s.stage_ids, 4 * stage));
    if (!s.d_work) {
        RTB_CUDA(cudaMalloc(&s.d_work, sizeof(unsigned long long)));
        RTB_CUDA(cudaMemset(s.d_work, 0, sizeof(unsigned long long)));
        s.work_base = 0;
    }
    if (!s.stream) RTB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.done) RTB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    s.rect_valid = false;  // nothing is known about a fresh host frame
    s.ready = true;
    return RTB_OK;
}
int wait_slot(FrameSlot& s) {
    if (s.in_flight) {
        RTB_CUDA(cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    return RTB_OK;
}
// every frame in flight on the slots has landed and no prediction is outstanding
int drain_slots(rtb_camera* c) {
    c->ahead.clear();
    c->ahead_obj = nullptr;
    for (FrameSlot& s : c->slots) { const int rc = wait_slot(s); if (rc) return rc; }
    return RTB_OK;
}
void free_slots(rtb_camera* c) {
    for (int k = 0; k < kFrameSlots; k++) {
        FrameSlot& s = c->slots[k];
        if (s.in_flight && s.done) cudaEventSynchronize(s.done);
        if (k != 0) { if (s.h_bgra) cudaFreeHost(s.h_bgra); if (s.h_ids) cudaFreeHost(s.h_ids); }
        if (s.stage_bgra) cudaFree(s.stage_bgra);
        if (s.stage_ids) cudaFree(s.stage_ids);
        if (s.d_work) cudaFree(s.d_work);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
        s = FrameSlot();
    }
    for (FrameSlot& s : c->sweep_lane) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.stage_bgra) cudaFree(s.stage_bgra);
        if (s.stage_ids) cudaFree(s.stage_ids);
        if (s.d_work) cudaFree(s.d_work);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = FrameSlot();
    }
    c->sweep_lane_elements = 0;
    c->ahead.clear();
    c->cur_slot = 0;
}
// the two lanes of a host-direct sweep hold at least `elements` staging elements per array
int ensure_sweep_lanes(rtb_camera* c, size_t elements) {
    for (FrameSlot& s : c->sweep_lane) {
        if (!s.stream) RTB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (!s.d_work) {
            RTB_CUDA(cudaMalloc(&s.d_work, sizeof(unsigned long long)));
            RTB_CUDA(cudaMemset(s.d_work, 0, sizeof(unsigned long long)));
            s.work_base = 0;
        }
    }
    if (c->sweep_lane_elements >= elements) return RTB_OK;
    for (FrameSlot& s : c->sweep_lane) {
        RTB_CUDA(cudaStreamSynchronize(s.stream));
        if (s.stage_bgra) cudaFree(s.stage_bgra);
        if (s.stage_ids) cudaFree(s.stage_ids);
        s.stage_bgra = nullptr; s.stage_ids = nullptr;
    }
    c->sweep_lane_elements = 0;
    for (FrameSlot& s : c->sweep_lane) {
        RTB_CUDA(cudaMalloc(&s.stage_bgra, 4 * elements));
        RTB_CUDA(cudaMalloc(&s.stage_ids, 4 * elements));
    }
    c->sweep_lane_elements = elements;
    return RTB_OK;
}
// a slot that is neither the current frame's nor holds a predicted frame; one whose kernel has finished if there is one
int pick_slot(rtb_camera* c, int keep) {
    int pick = -1;
    for (int k = 0; k < kFrameSlots; k++) {
        if (k == keep || std::find(c->ahead.begin(), c->ahead.end(), k) != c->ahead.end()) continue;
        FrameSlot& s = c->slots[k];
        if (s.in_flight && s.done && cudaEventQuery(s.done) == cudaSuccess) s.in_flight = false;
        if (!s.in_flight) return k;
        if (pick < 0) pick = k;
    }
    cudaGetLastError();  // (cudaEventQuery's cudaErrorNotReady)
    return pick;
}
// One frame with matrix m12 into slot k, on the slot's stream: the kernel stores finished work units straight into the
// slot's pinned host frame, and only inside the window of tiles where that frame can differ from what it holds.
int launch_into_slot(rtb_object* obj, rtb_camera* cam, int k, const float m12[12], uint32_t flags) {
    int rc = ensure_slot(cam, k);
    if (rc) return rc;
    FrameSlot& s = cam->slots[k];
    rc = wait_slot(s);  // (an abandoned prediction may still be running there)
    if (rc) return rc;
    float record[rtb::kFrameStride];
    fill_frame_record(obj->scene, cam, m12, record);
    int now[4], prev[4] = {0, 0, cam->basis.W - 1, cam->basis.H - 1};
    std::memcpy(now, record + 12, sizeof now);
    if (s.rect_valid) std::memcpy(prev, s.rect, sizeof prev);
    // window = the tiles touched by either rectangle (an empty rectangle has x1 < x0)
    const bool now_empty = now[2] < now[0] || now[3] < now[1], prev_empty = prev[2] < prev[0] || prev[3] < prev[1];
    int win[4] = {0, 0, -1, -1};
    if (!now_empty || !prev_empty) {
        const int x0 = now_empty ? prev[0] : prev_empty ? now[0] : std::min(now[0], prev[0]), y0 = now_empty ? prev[1] : prev_empty ? now[1] : std::min(now[1], prev[1]);
        const int x1 = now_empty ? prev[2] : prev_empty ? now[2] : std::max(now[2], prev[2]), y1 = now_empty ? prev[3] : prev_empty ? now[3] : std::max(now[3], prev[3]);
        win[0] = x0 / rtb::kTile; win[1] = y0 / rtb::kTile; win[2] = x1 / rtb::kTile; win[3] = y1 / rtb::kTile;
    }
    uint32_t* owner_bgra = s.h_bgra;  // unified addressing: a pinned allocation has the same address on the device
    int32_t* owner_ids = s.h_ids;
    s.rect_valid = false;
    rc = launch_render(obj, cam, nullptr, record, 1, 0, 1, flags, s.stage_bgra, s.stage_ids, s.stream, 1, &owner_bgra, &owner_ids, nullptr, prev, win, &s);
    if (rc) return rc;
    RTB_CUDA(cudaEventRecord(s.done, s.stream));
    s.in_flight = true;
    std::memcpy(s.rect, now, sizeof now);
    s.rect_valid = true;
    std::memcpy(s.m12, m12, sizeof s.m12);
    s.flags = flags;
    return RTB_OK;
}

void release_scene(rtb_camera* c, SceneArrays* sc) {
    if (!sc || --sc->refs > 0) return;
    cudaFree(sc->d_scene);
    cudaFree(sc->d_rad);
    c->scenes.erase(std::remove(c->scenes.begin(), c->scenes.end(), sc), c->scenes.end());
    delete sc;
}

// the object leaves its camera (Camera::object_list loses it; its share of the camera-side arrays is released)
void detach_object(rtb_object* o) {
    rtb_camera* c = o->cam;
    if (!c) return;
    cudaSetDevice(c->device);
    if (o->launch_pending) { cudaEventSynchronize(o->ev_launch); o->launch_pending = false; }  // kernels still read the arrays
    drain_slots(c);  // ... and so may frames in flight on the camera's slots (predictions included)
    c->objects.erase(std::remove(c->objects.begin(), c->objects.end(), o), c->objects.end());
    release_scene(c, o->scene);
    o->scene = nullptr;
    o->cam = nullptr;
}

int check_bound(rtb_object* o, rtb_camera* c, const char* who) {
    if (!o || !c) return fail(RTB_ERR_ARG, std::string(who) + ": null handle");
    if (o->cam != c || !o->scene) return fail(RTB_ERR_STATE, std::string(who) + ": object was not added to this camera (rtb_camera_add_object)");
    return RTB_OK;
}

}  // namespace

extern "C" {

const char* rtb_last_error(void) { return g_error.c_str(); }
const char* rtb_version(void) { return "rtb 0.1 (sm_100a)"; }

int rtb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int rtb_set_device(int device) {
    RTB_CUDA(cudaSetDevice(device));
    g_device = device;
    return RTB_OK;
}

int rtb_set_knob(const char* name, int value) {
    if (!name) return fail(RTB_ERR_ARG, "set_knob: null name");
    Knobs& k = knobs();
    const std::string n(name);
    if (n == "unit_shift") k.unit_shift = value;
    else if (n == "t_active") k.t_active = value;
    else if (n == "t_leaf") k.t_leaf = value;
    else if (n == "tail5") k.tail5 = value;
    else if (n == "tail6") k.tail6 = value;
    else if (n == "reserve_sms") k.reserve_sms = value;
    else if (n == "l2_window") k.l2_window = value;
    else if (n == "no_rect") k.no_rect = value;
    else if (n == "frame_order") k.frame_order = value;
    else if (n == "host_direct") k.host_direct = value;
    else if (n == "sweep_direct") k.sweep_direct = value;
    else if (n == "host_fill_threads") k.host_fill_threads = value;
    else if (n == "sweep_chunk_mb") k.sweep_chunk_mb = value;
    else if (n == "t_active_inline") k.t_active_inline = value;
    else if (n == "steal_spin") k.steal_spin = value;
    else if (n == "inline_prefetch") k.inline_prefetch = value;
    else if (n == "lookahead") k.lookahead = value;
    else if (n == "fill_ctas_per_sm") k.fill_ctas_per_sm = value;
    else if (n == "l2_carve_mb") k.l2_carve_mb = value;
    else return fail(RTB_ERR_ARG, "set_knob: unknown knob " + n);
    return RTB_OK;
}

int rtb_read_ply(const char* file_name, int mode, float** points9, uint32_t* num_tri) {
    return guarded("read_ply", [&]() -> int {
    if (!file_name || !points9 || !num_tri) return fail(RTB_ERR_ARG, "read_ply: null argument");
    std::vector<float> pts;
    std::string err = rtb::load_ply(file_name, mode, pts);
    if (!err.empty()) return fail(RTB_ERR_IO, err);
    float* out = (float*)std::malloc(sizeof(float) * std::max<size_t>(pts.size(), 1));
    if (!out) return fail(RTB_ERR_NOMEM, "read_ply: out of memory");
    std::memcpy(out, pts.data(), sizeof(float) * pts.size());
    *points9 = out;
    *num_tri = (uint32_t)(pts.size() / 9);
    return RTB_OK;
    });
}
void rtb_free(void* p) { std::free(p); }

int rtb_write_ply(const char* file_name, const float* points9, uint32_t num_tri) {
    return guarded("write_ply", [&]() -> int {
    if (!file_name || !points9) return fail(RTB_ERR_ARG, "write_ply: null argument");
    std::string err = rtb::save_ply(file_name, points9, num_tri);
    if (!err.empty()) return fail(RTB_ERR_IO, err);
    return RTB_OK;
    });
}

int rtb_mesh_geodesic(int nu, float radius, const float center[3], float displacement, uint32_t seed, float** points9,
                      uint32_t* num_tri) {
    return guarded("mesh_geodesic", [&]() -> int {
    if (nu < 1 || nu > 4000 || !center || !points9 || !num_tri) return fail(RTB_ERR_ARG, "mesh_geodesic: bad argument");
    std::vector<float> pts;
    rtb::make_geodesic(nu, radius, center, displacement, seed, pts);
    float* out = (float*)std::malloc(sizeof(float) * pts.size());
    if (!out) return fail(RTB_ERR_NOMEM, "mesh_geodesic: out of memory");
    std::memcpy(out, pts.data(), sizeof(float) * pts.size());
    *points9 = out;
    *num_tri = (uint32_t)(pts.size() / 9);
    return RTB_OK;
    });
}

// ---- mesh ----------------------------------------------------------------------------------------

int rtb_mesh_create(const float* points9, int64_t num_tri, const float* rad3, const float uniform_rgb[3], rtb_mesh** out) {
    if (!points9 || num_tri <= 0 || num_tri > 0x1fffffff || !out) return fail(RTB_ERR_ARG, "mesh_create: bad argument");
    rtb_mesh* m = nullptr;
    try {
        m = new rtb_mesh();
        m->points.assign(points9, points9 + 9 * (size_t)num_tri);
        if (rad3) m->rad.assign(rad3, rad3 + 3 * (size_t)num_tri);
    } catch (const std::bad_alloc&) {
        delete m;
        return fail(RTB_ERR_NOMEM, "mesh_create: out of host memory");
    }
    m->device = g_device;
    m->n = num_tri;
    if (uniform_rgb) std::memcpy(m->uniform_rgb, uniform_rgb, 12);
    cudaError_t e = cudaSetDevice(m->device);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_points, sizeof(float) * 9 * (size_t)num_tri);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_points, points9, sizeof(float) * 9 * (size_t)num_tri, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        // like the reference, the object exists but carries the error: the tree can still be built
        // and inspected on the host; anything that renders will fail with RTB_ERR_CUDA
        cudaGetLastError();
        m->d_points = nullptr;
        g_error = std::string("mesh_create: ") + cudaGetErrorString(e);
    }
    *out = m;
    return e == cudaSuccess ? RTB_OK : RTB_ERR_CUDA;
}

int rtb_mesh_build_tree_on(rtb_mesh* mesh, int where) {
    return guarded("build_tree", [&]() -> int {
    if (!mesh) return fail(RTB_ERR_ARG, "build_tree: null mesh");
    if (where < 0 || where > 2) return fail(RTB_ERR_ARG, "build_tree: where must be 0 (auto), 1 (host) or 2 (device)");
    const bool device = where == 2 || (where == 0 && mesh->d_points != nullptr);
    if (device) {
        if (!mesh->d_points) return fail(RTB_ERR_CUDA, "build_tree: the mesh has no device copy (no usable GPU)");
        RTB_CUDA(cudaSetDevice(mesh->device));
        rtb::free_device_tree(mesh->dtree);
        const std::string err = rtb::build_tree_gpu(mesh->d_points, mesh->n, mesh->dtree);
        if (!err.empty()) return fail(RTB_ERR_CUDA, err);
        g_launches += (uint64_t)mesh->dtree.launches;
        mesh->host_tree_valid = false;
        mesh->seconds_sort = mesh->dtree.seconds_sort; mesh->seconds_partition = mesh->dtree.seconds_partition;
    } else {
        if (mesh->d_points) { cudaSetDevice(mesh->device); rtb::free_device_tree(mesh->dtree); }
        rtb::build_tree(mesh->points.data(), mesh->n, mesh->tree, 0);
        mesh->host_tree_valid = true;
        mesh->seconds_sort = mesh->tree.seconds_sort; mesh->seconds_partition = mesh->tree.seconds_partition;
    }
    mesh->built = true;
    mesh->built_on_device = device;
    mesh->generation++;
    return RTB_OK;
    });
}
int rtb_mesh_build_tree(rtb_mesh* mesh) { return rtb_mesh_build_tree_on(mesh, 0); }
int64_t rtb_mesh_num_triangles(const rtb_mesh* mesh) { return mesh ? mesh->n : 0; }
int64_t rtb_mesh_num_nodes(const rtb_mesh* mesh) { return mesh ? 2 * mesh->n - 1 : 0; }

int rtb_mesh_get_tree(const rtb_mesh* mesh_c, int32_t* left, int32_t* right, int32_t* tri, int32_t* cut_flag, float* bounds6,
                      float* s1, float* s2) {
    return guarded("get_tree", [&]() -> int {
    rtb_mesh* mesh = const_cast<rtb_mesh*>(mesh_c);  // a device-built tree is downloaded on first use
    if (!mesh || !mesh->built) return fail(RTB_ERR_STATE, "get_tree: tree not built");
    const int rc = ensure_host_tree(mesh);
    if (rc) return rc;
    const rtb::HostTree& T = mesh->tree;
    for (int64_t i = 0; i < T.num_nodes; i++) {
        if (left) left[i] = T.left[i];
        if (right) right[i] = T.left[i] < 0 ? -1 : T.left[i] + 1;
        if (tri) tri[i] = T.tri[i];
        if (cut_flag) cut_flag[i] = T.cut_flag[i];
        if (s1) s1[i] = T.s1[i];
        if (s2) s2[i] = T.s2[i];
    }
    if (bounds6) std::memcpy(bounds6, T.bounds.data(), sizeof(float) * 6 * (size_t)T.num_nodes);
    return RTB_OK;
    });
}
int rtb_mesh_save_tree(const rtb_mesh* mesh_c, const char* file_name) {
    return guarded("save_tree", [&]() -> int {
    rtb_mesh* mesh = const_cast<rtb_mesh*>(mesh_c);
    if (!mesh || !file_name || !mesh->built) return fail(RTB_ERR_STATE, "save_tree: tree not built");
    const int rc = ensure_host_tree(mesh);
    if (rc) return rc;
    const std::string err = rtb::save_tree(file_name, mesh->tree, mesh->points.data());
    return err.empty() ? RTB_OK : fail(RTB_ERR_IO, err);
    });
}
int rtb_mesh_load_tree(rtb_mesh* mesh, const char* file_name) {
    return guarded("load_tree", [&]() -> int {
    if (!mesh || !file_name) return fail(RTB_ERR_ARG, "load_tree: null argument");
    rtb::HostTree T;
    const std::string err = rtb::load_tree(file_name, mesh->points.data(), mesh->n, T);
    if (!err.empty()) return fail(RTB_ERR_IO, err);
    if (mesh->d_points) { cudaSetDevice(mesh->device); rtb::free_device_tree(mesh->dtree); }
    mesh->tree = std::move(T);
    mesh->built = true; mesh->built_on_device = false; mesh->host_tree_valid = true;
    mesh->generation++;
    mesh->seconds_sort = mesh->seconds_partition = 0.0;
    return RTB_OK;
    });
}
int rtb_write_frame(const char* file_name, const uint32_t* bgra, int32_t width, int32_t height) {
    return guarded("write_frame", [&]() -> int {
    if (!file_name || !bgra || width <= 0 || height <= 0) return fail(RTB_ERR_ARG, "write_frame: bad argument");
    const std::string err = rtb::save_frame(file_name, bgra, width, height);
    return err.empty() ? RTB_OK : fail(RTB_ERR_IO, err);
    });
}
int rtb_mesh_build_seconds(const rtb_mesh* mesh, double out3[3]) {
    if (!mesh || !mesh->built || !out3) return fail(RTB_ERR_STATE, "build_seconds: tree not built");
    out3[0] = mesh->seconds_sort; out3[1] = mesh->seconds_partition; out3[2] = out3[0] + out3[1];
    return RTB_OK;
}
void rtb_mesh_destroy(rtb_mesh* mesh) {
    if (!mesh) return;
    if (mesh->d_points) { cudaSetDevice(mesh->device); cudaFree(mesh->d_points); rtb::free_device_tree(mesh->dtree); }
    delete mesh;
}

// ---- camera --------------------------------------------------------------------------------------

int rtb_camera_create(int32_t r_w, int32_t r_h, float f_w, float f_h, float fclen, const float pos[3], const float look_at[3],
                      const float up[3], rtb_camera** out) {
    if (r_w <= 0 || r_h <= 0 || !pos || !look_at || !up || !out) return fail(RTB_ERR_ARG, "camera_create: bad argument");
    rtb_camera* c = new rtb_camera();
    c->device = g_device;
    rtb::camera_basis(r_w, r_h, f_w, f_h, fclen, pos, look_at, up, c->basis);
    c->pixels = (long long)r_w * r_h;
    *out = c;
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_bgra, sizeof(uint32_t) * (size_t)c->pixels);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ids, sizeof(int32_t) * (size_t)c->pixels);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_bgra, sizeof(uint32_t) * (size_t)c->pixels);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_ids, sizeof(int32_t) * (size_t)c->pixels);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counters, sizeof(unsigned long long) * kCounterWords);
    if (e == cudaSuccess) e = cudaMemset(c->d_counters, 0, sizeof(unsigned long long) * kCounterWords);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&c->ev_render[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
    if (e == cudaSuccess) {
        // init_cam_mem_cuda zeroes the frame (Camera.cu:98) and sets ids to -1 (Camera.cu:110)
        std::memset(c->h_bgra, 0, sizeof(uint32_t) * (size_t)c->pixels);
        for (long long i = 0; i < c->pixels; i++) c->h_ids[i] = -1;
        e = cudaMemset(c->d_bgra, 0, sizeof(uint32_t) * (size_t)c->pixels);
        if (e == cudaSuccess) e = cudaMemset(c->d_ids, 0xff, sizeof(int32_t) * (size_t)c->pixels);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_error = std::string("camera_create: ") + cudaGetErrorString(e);
        return RTB_ERR_CUDA;  // the handle stays valid for host-side queries (basis)
    }
    return RTB_OK;
}

int rtb_camera_get_basis(const rtb_camera* cam, float out18[18]) {
    if (!cam || !out18) return fail(RTB_ERR_ARG, "camera_get_basis: null argument");
    const rtb::CameraBasis& b = cam->basis;
    const float* src[6] = {b.n, b.v, b.u, b.n_mod, b.v_mod, b.u_mod};
    for (int k = 0; k < 6; k++) std::memcpy(out18 + 3 * k, src[k], 12);
    return RTB_OK;
}

namespace {
// L2 residency (the 800k-triangle dragon fits the B200's L2).  A persisting access-policy window
// over the NODE records is attached to every render launch (launch_render), and the device's persisting carve-out is set
// to exactly their size.  Measured on the dragon stand-in (873 620 triangles: 56 MB of nodes + 42 MB of triangles,
// 600-frame launches, 126 MB L2 of which at most 79 MB may persist):
//     no window                                            12.89 ms
//     nodes + triangles, carve-out 79 MB, hit ratio 0.81   12.95 ms   (round 1)
//     nodes only,        carve-out 79 MB                   12.94 ms   (the carve-out starves triangles, stack and frames)
//     nodes only,        carve-out 56 MB = the window      11.90 ms
// Knob l2_window: 0 off, 1 node records (default), 2 nodes + triangles; l2_carve_mb > 0 forces the carve-out size.
void choose_l2_window(const rtb_camera* cam, SceneArrays* sc) {
    sc->l2_window_bytes = 0;
    const int mode = knobs().l2_window;
    if (mode <= 0) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cam->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, cam->device);
    if (max_persist > 0 && max_window > 0) {
        const size_t span = std::min<size_t>(mode == 2 ? sc->scene_bytes : sc->node_bytes, (size_t)max_window);
        size_t carve = std::min<size_t>(span, (size_t)max_persist);
        if (knobs().l2_carve_mb > 0) carve = std::min<size_t>((size_t)knobs().l2_carve_mb << 20, (size_t)max_persist);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
            sc->l2_window_base = sc->d_scene;
            sc->l2_window_bytes = span;
            sc->l2_hit_ratio = (float)std::min(1.0, (double)carve / (double)span);
        }
    }
    cudaGetLastError();
}

// The camera-relative device arrays of one mesh for one camera (init_camera_trixel_device_memory Trixel.cu:244 +
// init_camera_voxel_device_memory Camera.cu:163).  Every temporary is released on every path.
int make_scene_arrays(rtb_camera* cam, rtb_mesh* m, SceneArrays** out) {
    const bool dev_tree = m->built_on_device;
    const int64_t n = m->n, N = 2 * n - 1, interior = n - 1;
    SceneArrays* sc = new SceneArrays();
    sc->mesh = m; sc->generation = m->generation; sc->num_tri = n;
    std::memcpy(sc->uniform_rgb, m->uniform_rgb, sizeof sc->uniform_rgb);
    sc->node_bytes = sizeof(float4) * 4 * (size_t)std::max<int64_t>(interior, 1);
    const size_t tri_bytes = sizeof(float4) * 3 * (size_t)n;
    sc->scene_bytes = sc->node_bytes + tri_bytes;
    // Record index of every interior node.  The tree is numbered breadth-first (Trixel.h:143); the
    // device records are laid out depth-first (pre-order, left child first) so that a descent walks
    // forward through memory: a node and its left child share a 128-byte line, and a subtree is one
    // contiguous range (better L1/L2 locality for neighbouring rays).  Node identity never reaches
    // the output, so the order is free.  A device-built tree brings the pre-order ranks with it
    // (DeviceTree::rec) and never visits the host; a host-built tree is numbered and uploaded here.
    // RTB_NODE_ORDER=bfs keeps the reference numbering (host-built trees only).
    float* d_bounds = nullptr; int* d_left = nullptr; int* d_tri = nullptr; unsigned char* d_cut = nullptr; int* d_rec = nullptr;
    const float* root_bounds = nullptr;
    int root_tri = -1;
    cudaError_t e = cudaMalloc(&sc->d_scene, sc->scene_bytes);
    if (e == cudaSuccess) {
        sc->d_nodes = sc->d_scene;
        sc->d_tris = sc->d_scene + sc->node_bytes / sizeof(float4);
    }
    if (dev_tree) {
        d_bounds = m->dtree.bounds; d_left = m->dtree.left; d_tri = m->dtree.tri; d_cut = m->dtree.cut; d_rec = m->dtree.rec;
        root_bounds = m->dtree.root_bounds; root_tri = m->dtree.root_tri;
    } else {
        const rtb::HostTree& T = m->tree;
        std::vector<int32_t> record_of((size_t)N, -1);
        const char* order = std::getenv("RTB_NODE_ORDER");
        int32_t next = 0;
        if (order && std::strcmp(order, "bfs") == 0) {
            for (int64_t i = 0; i < N; i++) record_of[(size_t)i] = T.left[(size_t)i] >= 0 ? next++ : -1;
        } else {
            std::vector<int32_t> stack;
            stack.push_back(0);
            while (!stack.empty()) {
                const int32_t node = stack.back();
                stack.pop_back();
                const int32_t l = T.left[(size_t)node];
                if (l < 0) continue;
                record_of[(size_t)node] = next++;
                stack.push_back(l + 1);  // right child after the whole left subtree
                stack.push_back(l);
            }
        }
        if (e == cudaSuccess) e = cudaMalloc(&d_bounds, sizeof(float) * 6 * (size_t)N);
        if (e == cudaSuccess) e = cudaMalloc(&d_left, sizeof(int) * (size_t)N);
        if (e == cudaSuccess) e = cudaMalloc(&d_tri, sizeof(int) * (size_t)N);
        if (e == cudaSuccess) e = cudaMalloc(&d_cut, (size_t)N);
        if (e == cudaSuccess) e = cudaMalloc(&d_rec, sizeof(int) * (size_t)N);
        if (e == cudaSuccess) e = cudaMemcpy(d_bounds, T.bounds.data(), sizeof(float) * 6 * (size_t)N, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_left, T.left.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_tri, T.tri.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_cut, T.cut_flag.data(), (size_t)N, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_rec, record_of.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice);
        root_bounds = T.bounds.data(); root_tri = T.tri[0];
    }
    const float cx = cam->basis.pos[0], cy = cam->basis.pos[1], cz = cam->basis.pos[2];
    if (e == cudaSuccess) {
        rtb::pack_triangles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cam->stream>>>(m->d_points, n, cx, cy, cz, sc->d_tris);
        if (interior > 0)
            rtb::pack_nodes_kernel<<<(unsigned)((N + 255) / 256), 256, 0, cam->stream>>>(d_bounds, d_left, d_tri, d_cut, d_rec, N, cx, cy, cz, sc->d_nodes);
        g_launches += interior > 0 ? 2 : 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && !m->rad.empty()) {
        std::vector<float4> rad4((size_t)n);
        for (int64_t i = 0; i < n; i++) rad4[(size_t)i] = make_float4(m->rad[3 * i], m->rad[3 * i + 1], m->rad[3 * i + 2], 0.0f);
        e = cudaMalloc(&sc->d_rad, sizeof(float4) * (size_t)n);
        if (e == cudaSuccess) e = cudaMemcpy(sc->d_rad, rad4.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(cam->stream);
    if (!dev_tree) { cudaFree(d_bounds); cudaFree(d_left); cudaFree(d_tri); cudaFree(d_cut); cudaFree(d_rec); }
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(sc->d_scene); cudaFree(sc->d_rad);
        delete sc;
        return fail(RTB_ERR_CUDA, std::string("add_object: ") + cudaGetErrorString(e));
    }
    // root box, camera-relative (same arithmetic as pack_nodes_kernel; host is compiled without FMA)
    const float* rb = root_bounds;
    sc->root_box[0] = (rb[0] - cx) + 0.0f; sc->root_box[3] = (rb[1] - cx) + 0.0f;
    sc->root_box[1] = (rb[2] - cy) + 0.0f; sc->root_box[4] = (rb[3] - cy) + 0.0f;
    sc->root_box[2] = (rb[4] - cz) + 0.0f; sc->root_box[5] = (rb[5] - cz) + 0.0f;
    sc->root_ref = n == 1 ? (int)(rtb::kRefLeaf | (unsigned)root_tri) : 0;

    *out = sc;
    return RTB_OK;
}
}  // namespace

int rtb_camera_add_object(rtb_camera* cam, rtb_object* obj) {
    return guarded("add_object", [&]() -> int {
    if (!cam || !obj) return fail(RTB_ERR_ARG, "add_object: null handle");
    rtb_mesh* m = obj->mesh;
    if (!m->built) return fail(RTB_ERR_STATE, "add_object: rtb_mesh_build_tree has not been called");
    if (!m->d_points || !cam->d_bgra) return fail(RTB_ERR_CUDA, "add_object: mesh or camera has no device memory (no usable GPU)");
    if (m->device != cam->device) return fail(RTB_ERR_ARG, "add_object: mesh and camera live on different devices");
    RTB_CUDA(cudaSetDevice(cam->device));
    // the camera-side arrays of this (mesh, tree): shared with the other objects that instance it on this camera
    SceneArrays* sc = nullptr;
    for (SceneArrays* have : cam->scenes)
        if (have->mesh == m && have->generation == m->generation) sc = have;
    if (!sc) {
        const int rc = make_scene_arrays(cam, m, &sc);
        if (rc) return rc;
        cam->scenes.push_back(sc);
    }
    choose_l2_window(cam, sc);
    sc->refs++;
    detach_object(obj);  // re-added, or moved over from another camera: leaves the old list, drops the old arrays
    RTB_CUDA(cudaSetDevice(cam->device));
    obj->device = cam->device;
    const int rc = ensure_object_scratch(obj);
    if (rc) { release_scene(cam, sc); return rc; }
    obj->scene = sc;
    obj->cam = cam;
    cam->objects.push_back(obj);
    // Camera.cpp:131-134: faces = -camera position, identity quaternion
    obj->xf.reset(cam->basis.pos);
    obj->xf_ready = true;
    obj->xf_overridden = false;
    obj->steps_now.clear(); obj->steps_frame.clear(); obj->steps_repeat = false;  // the recurrence starts over: nothing to predict from
    return RTB_OK;
    });
}

int rtb_camera_color_pixels(rtb_camera* cam, uint8_t tag) {
    if (!cam || !cam->d_bgra) return fail(RTB_ERR_CUDA, "color_pixels: camera has no device memory");
    RTB_CUDA(cudaSetDevice(cam->device));
    bool enqueued = false;  // something was put on the camera's own stream by this call
    if (tag == RTB_SET_COLOR_TAG) {
        // Camera.cu:77-82: the reference's SET case fills the frame with the background and FALLS THROUGH into the
        // Phong pass, so what reaches the host is "background + shaded hits of the last render" -- the call exists to
        // clear pixels that were hit in an earlier frame, because its Phong pass only touches hit pixels.  Here every
        // render already writes every pixel, so after a render the device frame IS that result; before the first
        // render there are no hits and the frame is the plain background.
        if (!cam->frame_rendered) {
            const rtb::CameraBasis& b = cam->basis;
            const uint32_t bg = ((uint32_t)b.background[3] << 24) | ((uint32_t)b.background[0] << 16) | ((uint32_t)b.background[1] << 8) | b.background[2];
            const unsigned grid = (unsigned)std::min<long long>((cam->pixels + 255) / 256, 148ll * 64);
            rtb::fill_kernel<<<grid, 256, 0, cam->stream>>>(cam->d_bgra, cam->pixels, bg);
            rtb::fill_ids_kernel<<<grid, 256, 0, cam->stream>>>(cam->d_ids, cam->pixels, -1);
            g_launches += 2;
            RTB_CUDA(cudaGetLastError());
            cam->frame_on_host = false;
            enqueued = true;
        }
    } else if (tag != RTB_PHONG_COLOR_TAG) {
        return fail(RTB_ERR_ARG, "color_pixels: unknown tag");
    }
    if (!cam->frame_on_host) {
        // the host frame becomes a copy of the device frame: slot 0's buffer, about whose background nothing is known afterwards
        const int rc = drain_slots(cam);
        if (rc) return rc;
        cam->cur_slot = 0;
        cam->slots[0].rect_valid = false;
        RTB_CUDA(cudaMemcpyAsync(cam->h_bgra, cam->d_bgra, sizeof(uint32_t) * (size_t)cam->pixels, cudaMemcpyDeviceToHost, cam->stream));
        RTB_CUDA(cudaMemcpyAsync(cam->h_ids, cam->d_ids, sizeof(int32_t) * (size_t)cam->pixels, cudaMemcpyDeviceToHost, cam->stream));
        enqueued = true;
    }
    // the one synchronisation of a frame: the reference's wrappers synchronise after every launch (Trixel.cu:234,
    // Camera.cu:83), but nothing can observe the frame before this call returns it
    if (enqueued) RTB_CUDA(cudaStreamSynchronize(cam->stream));
    {
        const int rc = wait_slot(cam->slots[cam->cur_slot]);  // a frame stored by its kernel: complete when the slot's event has passed
        if (rc) return rc;
    }
    cam->frame_on_host = true;
    return RTB_OK;
}
// (the frame of the last render lives in the current slot's pinned buffer; slot 0 is the camera's own)
const uint32_t* rtb_camera_host_color(const rtb_camera* cam) {
    if (!cam) return nullptr;
    const FrameSlot& s = cam->slots[cam->cur_slot];
    return s.h_bgra ? s.h_bgra : cam->h_bgra;
}
const int32_t* rtb_camera_host_ids(const rtb_camera* cam) {
    if (!cam) return nullptr;
    const FrameSlot& s = cam->slots[cam->cur_slot];
    return s.h_ids ? s.h_ids : cam->h_ids;
}

static int read_counters(rtb_camera* cam, uint64_t* out, int count, int reset) {
    if (!cam || !cam->d_counters || !out) return fail(RTB_ERR_ARG, "counters: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    unsigned long long h[8];
    RTB_CUDA(cudaStreamSynchronize(cam->stream));
    RTB_CUDA(cudaMemcpy(h, cam->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    for (int k = 0; k < count; k++) out[k] = h[k];
    if (reset) RTB_CUDA(cudaMemset(cam->d_counters, 0, sizeof h));
    return RTB_OK;
}
#ifdef RTB_WARP_LOG
// development builds only: the per-warp log of the last launch (8 words per warp, see rtb_render.cuh)
int rtb_camera_warp_log(rtb_camera* cam, uint64_t* out, int warps) {
    if (!cam || !out || warps < 1 || warps > 8192) return fail(RTB_ERR_ARG, "warp_log: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(out, cam->d_counters + 16, sizeof(unsigned long long) * 8 * (size_t)warps, cudaMemcpyDeviceToHost));
    return RTB_OK;
}
#endif
int rtb_camera_counters(rtb_camera* cam, uint64_t out5[5], int reset) { return read_counters(cam, out5, 5, reset); }
int rtb_camera_counters_ex(rtb_camera* cam, uint64_t out8[8], int reset) { return read_counters(cam, out8, 8, reset); }

void rtb_camera_destroy(rtb_camera* cam) {
    if (!cam) return;
    cudaSetDevice(cam->device);
    if (cam->stream) cudaStreamSynchronize(cam->stream);
    if (cam->copy_stream) cudaStreamSynchronize(cam->copy_stream);
    free_slots(cam);
    dfree(cam->d_bgra); dfree(cam->d_ids); dfree(cam->d_counters); dfree(cam->push_bgra); dfree(cam->push_ids);
    if (cam->h_bgra) cudaFreeHost(cam->h_bgra);
    if (cam->h_ids) cudaFreeHost(cam->h_ids);
    for (int k = 0; k < 2; k++) {
        dfree(cam->ring_bgra[k]); dfree(cam->ring_ids[k]);
        if (cam->stage_bgra[k]) cudaFreeHost(cam->stage_bgra[k]);
        if (cam->stage_ids[k]) cudaFreeHost(cam->stage_ids[k]);
        if (cam->ev_render[k]) cudaEventDestroy(cam->ev_render[k]);
        if (cam->ev_copy[k]) cudaEventDestroy(cam->ev_copy[k]);
    }
    // every object still on this camera loses it (and its share of the camera-side arrays) but stays a valid handle
    while (!cam->objects.empty()) detach_object(cam->objects.back());
    for (SceneArrays* sc : cam->scenes) { cudaFree(sc->d_scene); cudaFree(sc->d_rad); delete sc; }  // (none left unless refs leaked)
    if (cam->stream) cudaStreamDestroy(cam->stream);
    if (cam->copy_stream) cudaStreamDestroy(cam->copy_stream);
    delete cam;
}

// ---- object --------------------------------------------------------------------------------------

int rtb_object_create(rtb_mesh* mesh, rtb_object** out) {
    if (!mesh || !out) return fail(RTB_ERR_ARG, "object_create: null argument");
    rtb_object* o = new rtb_object();
    o->mesh = mesh;
    *out = o;
    return RTB_OK;
}

int rtb_object_transform_host(rtb_object* obj, const float xyzw[4], uint8_t select, float m12_out[12]) {
    if (!obj || !xyzw) return fail(RTB_ERR_ARG, "transform: null argument");
    if (!obj->xf_ready) return fail(RTB_ERR_STATE, "transform: object was not added to a camera");
    if (obj->xf_overridden)
        return fail(RTB_ERR_STATE, "transform: the matrix was set with rtb_object_set_matrix; rtb_camera_add_object restarts the recurrence");
    if (!obj->xf.apply(select, xyzw[0], xyzw[1], xyzw[2], xyzw[3])) return fail(RTB_ERR_ARG, "transform: unknown selector");
    if (obj->steps_now.size() < 16) obj->steps_now.push_back({select, {xyzw[0], xyzw[1], xyzw[2], xyzw[3]}});
    else obj->steps_frame.clear();  // too many steps per frame to be worth predicting
    if (m12_out) obj->xf.matrix(m12_out);
    return RTB_OK;
}
int rtb_object_transform(rtb_object* obj, const float xyzw[4], uint8_t select) {
    // the matrix reaches the device with the next render (it travels with the frame list), which
    // replaces the reference's two 1-thread kernels per transform (Quaternion.cu:21, Camera.cu:279,322)
    return rtb_object_transform_host(obj, xyzw, select, nullptr);
}
int rtb_transform_sequence_host(const float cam_pos[3], int32_t count, const float* ops5, float* m12_out) {
    if (!cam_pos || count < 0 || (count > 0 && (!ops5 || !m12_out))) return fail(RTB_ERR_ARG, "transform_sequence: bad argument");
    rtb::Transform xf;
    xf.reset(cam_pos);
    for (int32_t k = 0; k < count; k++) {
        const float* op = ops5 + 5 * (size_t)k;
        const int select = (int)op[0];
        if (select != 0 && !xf.apply((uint8_t)select, op[1], op[2], op[3], op[4])) return fail(RTB_ERR_ARG, "transform_sequence: unknown selector");
        xf.matrix(m12_out + 12 * (size_t)k);
    }
    return RTB_OK;
}
int rtb_object_get_matrix(const rtb_object* obj, float m12[12]) {
    if (!obj || !m12 || !obj->xf_ready) return fail(RTB_ERR_STATE, "get_matrix: object was not added to a camera");
    obj->xf.matrix(m12);
    return RTB_OK;
}
int rtb_object_set_matrix(rtb_object* obj, const float m12[12]) {
    if (!obj || !m12 || !obj->xf_ready) return fail(RTB_ERR_STATE, "set_matrix: object was not added to a camera");
    obj->xf.set_matrix(m12);
    obj->xf_overridden = true;
    return RTB_OK;
}
void rtb_object_destroy(rtb_object* obj) {
    if (!obj) return;
    detach_object(obj);
    cudaSetDevice(obj->device);
    if (obj->launch_pending) cudaEventSynchronize(obj->ev_launch);
    if (obj->upload_pending) cudaEventSynchronize(obj->ev_upload);
    dfree(obj->d_frames); dfree(obj->d_work);
    if (obj->h_frames) cudaFreeHost(obj->h_frames);
    if (obj->ev_upload) cudaEventDestroy(obj->ev_upload);
    if (obj->ev_launch) cudaEventDestroy(obj->ev_launch);
    delete obj;
}

// ---- render --------------------------------------------------------------------------------------

namespace {
// the camera's tile-major staging of the push variants holds at least `need` elements per array
int ensure_push_staging(rtb_camera* cam, size_t need) {
    if (cam->push_elements >= need) return RTB_OK;
    // the staging grows: kernels of any object of this camera that still stage into the old one must finish
    for (rtb_object* other : cam->objects)
        if (other->launch_pending) { RTB_CUDA(cudaEventSynchronize(other->ev_launch)); other->launch_pending = false; }
    dfree(cam->push_bgra); dfree(cam->push_ids);
    cam->push_elements = 0;
    RTB_CUDA(cudaMalloc(&cam->push_bgra, 4 * need));
    RTB_CUDA(cudaMalloc(&cam->push_ids, 4 * need));
    cam->push_elements = need;
    return RTB_OK;
}

// One frame of the object's current transform into the camera's device frame, no synchronisation: the record travels
// in the kernel parameters, the work counter is never reset -- the launch is the only stream operation.
int render_current(rtb_object* obj, rtb_camera* cam, uint32_t flags) {
    float m12[12], record[rtb::kFrameStride];
    obj->xf.matrix(m12);
    // do the transform steps since the last render repeat the ones before it?  (the prediction of the frames ahead)
    {
        auto same = [](const std::vector<rtb_object::Step>& x, const std::vector<rtb_object::Step>& y) {
            if (x.size() != y.size()) return false;
            for (size_t k = 0; k < x.size(); k++)
                if (x[k].select != y[k].select || std::memcmp(x[k].v, y[k].v, sizeof x[k].v) != 0) return false;
            return true;
        };
        obj->steps_repeat = same(obj->steps_now, obj->steps_frame);
        obj->steps_frame.swap(obj->steps_now);
        obj->steps_now.clear();
    }
    int rc;
    if ((flags & RTB_RENDER_COUNTERS) || !knobs().host_direct) {
        rc = drain_slots(cam);  // the frame goes through the device frame and the copy into slot 0's host frame
        if (rc) return rc;
        fill_frame_record(obj->scene, cam, m12, record);
    }
    if (flags & RTB_RENDER_COUNTERS) {  // the counting variants read their records from device memory
        rc = ensure_frames(obj, 1);
        if (rc) return rc;
        std::memcpy(obj->h_frames, record, sizeof record);
        rc = order_after_previous(obj, cam->stream);
        if (!rc) rc = upload_frames(obj, 1, cam->stream);
        if (!rc) rc = launch_render(obj, cam, obj->d_frames, nullptr, 1, 0, 1, flags, cam->d_bgra, cam->d_ids, cam->stream);
    } else if (knobs().host_direct) {
        // Straight into a host frame: the warps store their finished work units as row segments into a pinned, mapped buffer
        // (the peer-push variant with the host as the owner), so the frame crosses PCIe while the rest of it is still being
        // traced and no copy follows the kernel.  Work units that hold nothing but background are stored only where the
        // host frame may hold something else: inside the rectangle of the frame stored there before.
        const bool repeat = obj->steps_repeat;
        int slot = -1;
        if (!cam->ahead.empty() && cam->ahead_obj == obj && cam->ahead_flags == flags &&
            std::memcmp(cam->slots[cam->ahead.front()].m12, m12, sizeof m12) == 0) {
            slot = cam->ahead.front();  // predicted: this very frame is already on its way
            cam->ahead.pop_front();
            cam->ahead_misses = 0;
        } else {
            if (!cam->ahead.empty() && ++cam->ahead_misses >= 2) { cam->ahead_pause = 32; cam->ahead_misses = 0; }
            cam->ahead.clear();  // (wrong guesses finish on their own and are never looked at)
            slot = knobs().lookahead > 0 ? pick_slot(cam, cam->cur_slot) : 0;
            rc = launch_into_slot(obj, cam, slot, m12, flags);
            if (rc) return rc;
        }
        cam->cur_slot = slot;
        // frames ahead: only while the steps between two renders repeat, and never more than the free slots allow
        const int depth = std::min(knobs().lookahead, kFrameSlots - 2);
        if (cam->ahead_pause > 0) cam->ahead_pause--;
        if (depth > 0 && repeat && cam->ahead_pause == 0 && !obj->steps_frame.empty() && !obj->xf_overridden) {
            rtb::Transform t = cam->ahead.empty() ? obj->xf : cam->ahead_xf;
            while ((int)cam->ahead.size() < depth) {
                for (const rtb_object::Step& st : obj->steps_frame) t.apply(st.select, st.v[0], st.v[1], st.v[2], st.v[3]);
                float next12[12];
                t.matrix(next12);
                const int k = pick_slot(cam, cam->cur_slot);
                if (k < 0) break;
                rc = launch_into_slot(obj, cam, k, next12, flags);
                if (rc) return rc;
                cam->ahead.push_back(k);
            }
            cam->ahead_xf = t; cam->ahead_obj = obj; cam->ahead_flags = flags;
        }
        cam->frame_rendered = true;
        cam->frame_on_host = true;  // once the slot's event has passed, which is what rtb_camera_color_pixels waits for
        cam->device_frame_stale = true;
        return RTB_OK;
    } else {
        rc = launch_render(obj, cam, nullptr, record, 1, 0, 1, flags, cam->d_bgra, cam->d_ids, cam->stream);
    }
    if (rc) return rc;
    cam->frame_rendered = true;
    cam->frame_on_host = false;
    cam->device_frame_stale = false;
    return RTB_OK;
}
}  // namespace

int rtb_object_render(rtb_object* obj, rtb_camera* cam, uint32_t flags) {
    int rc = check_bound(obj, cam, "render");
    if (rc) return rc;
    RTB_CUDA(cudaSetDevice(cam->device));
    // The reference's wrapper synchronises here (Trixel.cu:234); the frame it waits for cannot be observed before
    // Camera::color_pixels brings it to the host, which is where this library synchronises -- once per frame.
    return render_current(obj, cam, flags);
}

// ---- scene extension: SURVEY.md section 8(f) items 3-4 (csrc/rtb_scene.cuh) ------------------------------
int rtb_camera_set_lights(rtb_camera* cam, int32_t num_lights, const float* xyz3) {
    if (!cam || num_lights < 1 || num_lights > rtb::kMaxSceneLights || !xyz3) return fail(RTB_ERR_ARG, "set_lights: 1..8 lights");
    cam->num_lights = num_lights;
    std::memcpy(cam->lights, xyz3, sizeof(float) * 3 * (size_t)num_lights);
    return RTB_OK;
}
int rtb_camera_set_shadows(rtb_camera* cam, int32_t enable) {
    if (!cam) return fail(RTB_ERR_ARG, "set_shadows: null camera");
    cam->shadows = enable != 0;
    return RTB_OK;
}
int rtb_camera_set_sample_rate(rtb_camera* cam, int32_t n) {
    if (!cam || n < 0 || n > 16) return fail(RTB_ERR_ARG, "set_sample_rate: 0..16");
    cam->sample_rate = n;
    return RTB_OK;
}
int64_t rtb_camera_object_id_base(const rtb_camera* cam, const rtb_object* obj) {
    if (!cam || !obj) return -1;
    int64_t base = 0;
    for (const rtb_object* o : cam->objects) {
        if (o == obj) return base;
        base += o->scene->num_tri;
    }
    return -1;
}

int rtb_camera_render_scene_device_async(rtb_camera* cam, uint32_t flags, uint32_t* d_bgra, int32_t* d_ids, void* stream) {
    if (!cam || !cam->d_bgra) return fail(RTB_ERR_CUDA, "render_scene: camera has no device memory");
    if (cam->objects.empty()) return fail(RTB_ERR_STATE, "render_scene: no object was added to this camera (rtb_camera_add_object)");
    if ((int)cam->objects.size() > rtb::kMaxSceneObjects) return fail(RTB_ERR_ARG, "render_scene: at most 8 objects per camera");
    if (flags & ~(uint32_t)RTB_RENDER_NO_CULL) return fail(RTB_ERR_ARG, "render_scene: only RTB_RENDER_NO_CULL is understood here");
    RTB_CUDA(cudaSetDevice(cam->device));
    using namespace rtb;
    SceneParams P;
    std::memset(&P, 0, sizeof P);
    const CameraBasis& b = cam->basis;
    P.W = b.W; P.H = b.H;
    for (int k = 0; k < 3; k++) { P.n_mod[k] = b.n_mod[k]; P.u_mod[k] = b.u_mod[k]; P.v_mod[k] = b.v_mod[k]; }
    P.draw_distance = b.draw_distance;
    P.background = ((uint32_t)b.background[3] << 24) | ((uint32_t)b.background[0] << 16) | ((uint32_t)b.background[1] << 8) | b.background[2];
    P.num_objects = (int)cam->objects.size();
    int64_t base = 0;
    for (int k = 0; k < P.num_objects; k++) {
        const rtb_object* o = cam->objects[(size_t)k];
        const SceneArrays* sc = o->scene;
        SceneObject& O = P.obj[k];
        O.nodes = sc->d_nodes; O.tris = sc->d_tris; O.rad = sc->d_rad;
        for (int c = 0; c < 3; c++) O.uniform_rad[c] = sc->uniform_rgb[c];
        for (int c = 0; c < 6; c++) O.root_box[c] = sc->root_box[c];
        O.root_ref = sc->root_ref;
        if (base + sc->num_tri > 0x7fffffff) return fail(RTB_ERR_ARG, "render_scene: more than 2^31 triangles in the scene");
        O.id_base = (int)base;
        base += sc->num_tri;
        o->xf.matrix(O.m);
    }
    P.num_lights = cam->num_lights;
    std::memcpy(P.lights, cam->lights, sizeof P.lights);
    P.shadows = cam->shadows ? 1 : 0;
    P.samples = cam->sample_rate >= 2 ? cam->sample_rate : 1;
    P.cull = (flags & RTB_RENDER_NO_CULL) ? 0 : 1;
    P.cull_rel = 1e-5f;
    const bool own = !d_bgra && !d_ids;
    P.out_bgra = own ? cam->d_bgra : d_bgra;
    P.out_ids = own ? cam->d_ids : d_ids;
    cudaStream_t s = stream ? (cudaStream_t)stream : cam->stream;
    // the objects' own launches (rtb_object_render ...) may still be writing the camera's frame on other streams
    for (rtb_object* o : cam->objects) { const int rc = order_after_previous(o, s); if (rc) return rc; }
    const dim3 grid((unsigned)((b.W + 15) / 16), (unsigned)((b.H + 7) / 8));
    render_scene_kernel<<<grid, kBlockThreads, 0, s>>>(P);
    g_launches++;
    RTB_CUDA(cudaGetLastError());
    for (rtb_object* o : cam->objects) {  // later launches of these objects are ordered behind this one
        RTB_CUDA(cudaEventRecord(o->ev_launch, s));
        o->launch_pending = true;
        o->last_stream = s;
    }
    if (own) { cam->frame_rendered = true; cam->frame_on_host = false; }
    return RTB_OK;
}
int rtb_camera_render_scene(rtb_camera* cam, uint32_t flags) { return rtb_camera_render_scene_device_async(cam, flags, nullptr, nullptr, nullptr); }

int rtb_render_frame(rtb_object* obj, rtb_camera* cam, uint32_t flags) {
    const int rc = rtb_object_render(obj, cam, flags);
    return rc ? rc : rtb_camera_color_pixels(cam, RTB_PHONG_COLOR_TAG);
}

int rtb_render_frames_device_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12, int32_t tile_first,
                                   int32_t tile_stride, uint32_t flags, uint32_t* d_bgra, int32_t* d_ids, void* stream) {
    int rc = check_bound(obj, cam, "render_frames_device");
    if (rc) return rc;
    if (num_frames <= 0 || !m12) return fail(RTB_ERR_ARG, "render_frames_device: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : cam->stream;
    if (num_frames == 1 && !(flags & RTB_RENDER_COUNTERS)) {
        float record[rtb::kFrameStride];
        fill_frame_record(obj->scene, cam, m12, record);
        return launch_render(obj, cam, nullptr, record, 1, tile_first, tile_stride, flags, d_bgra, d_ids, s);
    }
    rc = ensure_frames(obj, num_frames);  // (waits for the previous upload from the pinned staging, never for a kernel)
    if (rc) return rc;
    for (int f = 0; f < num_frames; f++) fill_frame_record(obj->scene, cam, m12 + 12 * (size_t)f, obj->h_frames + rtb::kFrameStride * (size_t)f);
    const bool ordered = order_frames(obj, cam, num_frames);
    rc = order_after_previous(obj, s);
    if (!rc) rc = upload_frames(obj, num_frames, s, ordered);
    if (rc) return rc;
    return launch_render(obj, cam, obj->d_frames, nullptr, num_frames, tile_first, tile_stride, flags, d_bgra, d_ids, s, 0, nullptr, nullptr,
                         ordered ? reinterpret_cast<const int*>(obj->d_frames + rtb::kFrameStride * (size_t)num_frames) : nullptr);
}

int rtb_render_frames_push_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12, int32_t tile_first,
                                 int32_t tile_stride, uint32_t flags, uint32_t* d_frame_bgra, int32_t* d_frame_ids, void* stream) {
    if (!d_frame_bgra && !d_frame_ids) return fail(RTB_ERR_ARG, "render_frames_push: bad argument");
    return rtb_render_frames_push_striped_async(obj, cam, num_frames, m12, tile_first, tile_stride, flags, 1, d_frame_bgra ? &d_frame_bgra : nullptr,
                                                d_frame_ids ? &d_frame_ids : nullptr, stream);
}

int rtb_render_frames_push_striped_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12, int32_t tile_first,
                                         int32_t tile_stride, uint32_t flags, int32_t owners, uint32_t* const* d_frame_bgra,
                                         int32_t* const* d_frame_ids, void* stream) {
    int rc = check_bound(obj, cam, "render_frames_push");
    if (rc) return rc;
    if (num_frames <= 0 || !m12 || (!d_frame_bgra && !d_frame_ids) || owners < 1 || owners > rtb::kMaxPushOwners)
        return fail(RTB_ERR_ARG, "render_frames_push: bad argument");
    for (int k = 0; k < owners; k++)
        if ((d_frame_bgra && !d_frame_bgra[k]) || (d_frame_ids && !d_frame_ids[k])) return fail(RTB_ERR_ARG, "render_frames_push: null owner buffer");
    if (flags & RTB_RENDER_COUNTERS) return fail(RTB_ERR_ARG, "render_frames_push: counters are not available in the push variant");
    RTB_CUDA(cudaSetDevice(cam->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : cam->stream;
    rc = ensure_frames(obj, num_frames);
    if (rc) return rc;
    rc = ensure_push_staging(cam, (size_t)num_frames * (size_t)rtb_tile_major_elements(cam, tile_stride));
    if (rc) return rc;
    for (int f = 0; f < num_frames; f++) fill_frame_record(obj->scene, cam, m12 + 12 * (size_t)f, obj->h_frames + rtb::kFrameStride * (size_t)f);
    const bool ordered = order_frames(obj, cam, num_frames);
    rc = order_after_previous(obj, s);
    if (!rc) rc = upload_frames(obj, num_frames, s, ordered);
    if (rc) return rc;
    return launch_render(obj, cam, obj->d_frames, nullptr, num_frames, tile_first, tile_stride, flags, d_frame_bgra ? cam->push_bgra : nullptr,
                         d_frame_ids ? cam->push_ids : nullptr, s, owners, d_frame_bgra, d_frame_ids,
                         ordered ? reinterpret_cast<const int*>(obj->d_frames + rtb::kFrameStride * (size_t)num_frames) : nullptr);
}

int rtb_fill_frames_device_async(rtb_camera* cam, int32_t num_frames, uint32_t* d_frame_bgra, int32_t* d_frame_ids, void* stream) {
    if (!cam || num_frames <= 0 || (!d_frame_bgra && !d_frame_ids)) return fail(RTB_ERR_ARG, "fill_frames: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : cam->stream;
    const rtb::CameraBasis& b = cam->basis;
    const uint32_t bg = ((uint32_t)b.background[3] << 24) | ((uint32_t)b.background[0] << 16) | ((uint32_t)b.background[1] << 8) | b.background[2];
    const long long count = (long long)num_frames * cam->pixels;
    // A few blocks per SM (knob fill_ctas_per_sm, default 2): queued in front of a persistent render launch on another stream,
    // the fill then runs beside it -- it takes two of the SM's thread-block slots for its duration instead of the whole GPU
    // for half a millisecond -- which is how bench.py's tile mode pre-fills the next step's frames.
    const unsigned grid = (unsigned)std::min<long long>((count / 4 + 511) / 512, (long long)std::max(1, cam->sm_count) * std::max(1, knobs().fill_ctas_per_sm));
    rtb::fill_frames_kernel<<<std::max(1u, grid), 512, 0, s>>>(d_frame_bgra, d_frame_ids, count, bg);
    g_launches++;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

// ---- peer memory: frames other processes' GPUs can write into over NVLink ---------------------------
int rtb_peer_alloc(size_t bytes, void** d_ptr) {
    if (!d_ptr || bytes == 0) return fail(RTB_ERR_ARG, "peer_alloc: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    RTB_CUDA(cudaMalloc(d_ptr, bytes));  // a whole allocation of its own: that is what an IPC handle names
    return RTB_OK;
}
int rtb_peer_free(void* d_ptr) {
    if (d_ptr) RTB_CUDA(cudaFree(d_ptr));
    return RTB_OK;
}
int rtb_peer_export(void* d_ptr, uint8_t handle64[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle64) return fail(RTB_ERR_ARG, "peer_export: bad argument");
    cudaIpcMemHandle_t h;
    RTB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle64, &h, 64);
    return RTB_OK;
}
int rtb_peer_open(const uint8_t handle64[64], void** d_ptr) {
    if (!d_ptr || !handle64) return fail(RTB_ERR_ARG, "peer_open: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    RTB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RTB_OK;
}
int rtb_peer_read(void* host_dst, const void* d_ptr, size_t bytes) {
    if (!host_dst || !d_ptr) return fail(RTB_ERR_ARG, "peer_read: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(host_dst, d_ptr, bytes, cudaMemcpyDeviceToHost));
    return RTB_OK;
}
int rtb_peer_close(void* d_ptr) {
    if (d_ptr) RTB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RTB_OK;
}

int rtb_render_sweep(rtb_object* obj, rtb_camera* cam, int32_t num_frames, int32_t steps_per_frame, const float* ops5,
                     uint32_t flags, uint32_t* bgra_out, int32_t* ids_out) {
    int rc = check_bound(obj, cam, "render_sweep");
    if (rc) return rc;
    if (num_frames <= 0 || steps_per_frame < 0 || (steps_per_frame > 0 && !ops5)) return fail(RTB_ERR_ARG, "render_sweep: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    RTB_CUDA(cudaStreamSynchronize(cam->stream));
    rc = ensure_frames(obj, num_frames);
    if (rc) return rc;
    rc = order_after_previous(obj, cam->stream);
    if (rc) return rc;
    // host recurrence for every frame (Camera.cu:254-335), exactly as if the calls were made one by one
    for (int f = 0; f < num_frames; f++) {
        for (int s = 0; s < steps_per_frame; s++) {
            const float* op = ops5 + 5 * ((size_t)f * steps_per_frame + s);
            const int select = (int)op[0];
            if (select == 0) continue;
            if (obj->xf_overridden)
                return fail(RTB_ERR_STATE, "render_sweep: the matrix was set with rtb_object_set_matrix; rtb_camera_add_object restarts the recurrence");
            if (!obj->xf.apply((uint8_t)select, op[1], op[2], op[3], op[4])) return fail(RTB_ERR_ARG, "render_sweep: unknown selector");
        }
        float m12[12];
        obj->xf.matrix(m12);
        fill_frame_record(obj->scene, cam, m12, obj->h_frames + rtb::kFrameStride * (size_t)f);
    }
    rc = upload_frames(obj, num_frames, cam->stream);
    if (rc) return rc;

    const size_t P = (size_t)cam->pixels;
    // caller buffers that are already pinned can be written by the device directly
    auto pinned_device_pointer = [](const void* p) -> void* {
        if (!p) return nullptr;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    };
    void* const dev_bgra = pinned_device_pointer(bgra_out);
    void* const dev_ids = pinned_device_pointer(ids_out);
    const bool direct_bgra = dev_bgra != nullptr, direct_ids = dev_ids != nullptr;

    // ---- pinned caller buffers: no copy engine at all -------------------------------------------------------------------
    // Most of a frame is background (94 % of the pixels of the default view), and background is a constant: the HOST writes
    // it, with streaming stores on a few threads (measured ~190 GB/s on the pool's hosts, against the ~57 GB/s a device-to-host
    // copy gets over PCIe), chunk by chunk, while the GPU traces the chunk before and stores only the work units that contain
    // something else straight into the caller's buffers (the peer-push variant of the kernel with the host as the owner of
    // the frames: 128-byte row stores over PCIe).  What crosses the bus is the picture, not the wallpaper.
    if (knobs().sweep_direct && !(flags & RTB_RENDER_COUNTERS) && (bgra_out || ids_out) && (!bgra_out || direct_bgra) && (!ids_out || direct_ids)) {
        const size_t frame_bytes = 4 * P * ((bgra_out ? 1u : 0u) + (ids_out ? 1u : 0u));
        // Chunks grow from 32 MB to the knob's size (256 MB): the first fill is the only one the GPU cannot hide behind, so it
        // is small; afterwards fill k+1 and launch k run side by side and a launch's fixed cost wants large chunks.
        const size_t cap_bytes = (size_t)std::max(1, knobs().sweep_chunk_mb) << 20;
        const int cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)num_frames, cap_bytes / frame_bytes));
        int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)cap, std::min<size_t>(cap_bytes / 8, (size_t)32 << 20) / frame_bytes));
        rc = ensure_sweep_lanes(cam, (size_t)cap * (size_t)rtb_tile_major_elements(cam, 1));
        if (rc) return rc;
        for (FrameSlot& lane : cam->sweep_lane) RTB_CUDA(cudaStreamWaitEvent(lane.stream, obj->ev_upload, 0));  // the frame records
        const rtb::CameraBasis& b = cam->basis;
        const uint32_t bg = ((uint32_t)b.background[3] << 24) | ((uint32_t)b.background[0] << 16) | ((uint32_t)b.background[1] << 8) | b.background[2];
        const int threads = host_fill_threads();
        int turn = 0;
        const int first_chunk = chunk;
        for (int f0 = 0; f0 < num_frames;) {
            FrameSlot& lane = cam->sweep_lane[turn];  // chunks alternate between two streams: a launch's tail (the latency of its
            turn ^= 1;                                // longest rays) runs beside the next chunk instead of in front of it
            // ... and they shrink again towards the end: what the call waits for after its last fill is the last launch
            const int left = num_frames - f0;
            const int nf = std::min(chunk, left > 2 * first_chunk ? (left + 1) / 2 : left);
            rtb::fill_words2(bgra_out ? bgra_out + (size_t)f0 * P : nullptr, bg, ids_out ? reinterpret_cast<uint32_t*>(ids_out) + (size_t)f0 * P : nullptr,
                             0xffffffffu, (size_t)nf * P, threads);
            uint32_t* owner_bgra = bgra_out ? static_cast<uint32_t*>(dev_bgra) + (size_t)f0 * P : nullptr;
            int32_t* owner_ids = ids_out ? static_cast<int32_t*>(dev_ids) + (size_t)f0 * P : nullptr;
            rc = launch_render(obj, cam, obj->d_frames + rtb::kFrameStride * (size_t)f0, nullptr, nf, 0, 1, flags | RTB_RENDER_PUSH_PREFILLED,
                               bgra_out ? lane.stage_bgra : nullptr, ids_out ? lane.stage_ids : nullptr, lane.stream, 1,
                               bgra_out ? &owner_bgra : nullptr, ids_out ? &owner_ids : nullptr, nullptr, nullptr, nullptr, &lane);
            if (rc) return rc;
            f0 += nf;
            chunk = std::min(cap, chunk * 2);
        }
        for (FrameSlot& lane : cam->sweep_lane) RTB_CUDA(cudaStreamSynchronize(lane.stream));
        RTB_CUDA(cudaStreamSynchronize(cam->stream));
        return RTB_OK;
    }

    // ---- pageable caller buffers: ring of two device chunks; chunk k+1 renders while chunk k streams to the host ------------
    int chunk_frames = (int)std::max<size_t>(1, std::min<size_t>((size_t)num_frames, (size_t)(16u << 20) / (P * 4)));
    if (cam->ring_frames < chunk_frames) {
        for (int k = 0; k < 2; k++) {
            dfree(cam->ring_bgra[k]); dfree(cam->ring_ids[k]);
            if (cam->stage_bgra[k]) { cudaFreeHost(cam->stage_bgra[k]); cam->stage_bgra[k] = nullptr; }
            if (cam->stage_ids[k]) { cudaFreeHost(cam->stage_ids[k]); cam->stage_ids[k] = nullptr; }
            RTB_CUDA(cudaMalloc(&cam->ring_bgra[k], 4 * P * (size_t)chunk_frames));
            RTB_CUDA(cudaMalloc(&cam->ring_ids[k], 4 * P * (size_t)chunk_frames));
            RTB_CUDA(cudaMallocHost(&cam->stage_bgra[k], 4 * P * (size_t)chunk_frames));
            RTB_CUDA(cudaMallocHost(&cam->stage_ids[k], 4 * P * (size_t)chunk_frames));
        }
        cam->ring_frames = chunk_frames;
    }
    chunk_frames = std::min(chunk_frames, cam->ring_frames);
    const int chunks = (num_frames + chunk_frames - 1) / chunk_frames;
    auto drain = [&](int k) -> int {  // staging -> caller memory for chunk k
        const int slot = k & 1, f0 = k * chunk_frames, nf = std::min(chunk_frames, num_frames - f0);
        RTB_CUDA(cudaEventSynchronize(cam->ev_copy[slot]));
        if (bgra_out && !direct_bgra) std::memcpy(bgra_out + (size_t)f0 * P, cam->stage_bgra[slot], 4 * P * (size_t)nf);
        if (ids_out && !direct_ids) std::memcpy(ids_out + (size_t)f0 * P, cam->stage_ids[slot], 4 * P * (size_t)nf);
        return RTB_OK;
    };
    for (int k = 0; k < chunks; k++) {
        const int slot = k & 1, f0 = k * chunk_frames, nf = std::min(chunk_frames, num_frames - f0);
        if (k >= 2) {
            // slot reuse: the device chunk must have been copied out, the staging chunk drained
            rc = drain(k - 2);
            if (rc) return rc;
            RTB_CUDA(cudaStreamWaitEvent(cam->stream, cam->ev_copy[slot], 0));
        }
        rc = launch_render(obj, cam, obj->d_frames + rtb::kFrameStride * (size_t)f0, nullptr, nf, 0, 1, flags, bgra_out ? cam->ring_bgra[slot] : nullptr,
                           ids_out ? cam->ring_ids[slot] : nullptr, cam->stream);
        if (rc) return rc;
        RTB_CUDA(cudaEventRecord(cam->ev_render[slot], cam->stream));
        RTB_CUDA(cudaStreamWaitEvent(cam->copy_stream, cam->ev_render[slot], 0));
        if (bgra_out)
            RTB_CUDA(cudaMemcpyAsync(direct_bgra ? bgra_out + (size_t)f0 * P : cam->stage_bgra[slot], cam->ring_bgra[slot], 4 * P * (size_t)nf, cudaMemcpyDeviceToHost, cam->copy_stream));
        if (ids_out)
            RTB_CUDA(cudaMemcpyAsync(direct_ids ? ids_out + (size_t)f0 * P : cam->stage_ids[slot], cam->ring_ids[slot], 4 * P * (size_t)nf, cudaMemcpyDeviceToHost, cam->copy_stream));
        RTB_CUDA(cudaEventRecord(cam->ev_copy[slot], cam->copy_stream));
    }
    for (int k = std::max(0, chunks - 2); k < chunks; k++) { rc = drain(k); if (rc) return rc; }
    RTB_CUDA(cudaStreamSynchronize(cam->stream));
    // (the camera's single-frame host buffers are not touched by a sweep: copying the last frame into
    // them would cost more host time than rendering it)
    return RTB_OK;
}

int rtb_host_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return fail(RTB_ERR_ARG, "host_alloc: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    RTB_CUDA(cudaMallocHost(ptr, bytes));
    return RTB_OK;
}
int rtb_host_free(void* ptr) {
    if (ptr) RTB_CUDA(cudaFreeHost(ptr));
    return RTB_OK;
}

int64_t rtb_tile_major_elements(const rtb_camera* cam, int32_t tile_stride) {
    if (!cam || tile_stride < 1) return 0;
    const int tiles = ((cam->basis.W + rtb::kTile - 1) / rtb::kTile) * ((cam->basis.H + rtb::kTile - 1) / rtb::kTile);
    return (int64_t)((tiles + tile_stride - 1) / tile_stride) * rtb::kTile * rtb::kTile;
}

int rtb_compose_tiles_device_async(rtb_camera* cam, int32_t num_frames, int32_t world, const void* const* d_parts, void* d_out, void* stream) {
    if (!cam || num_frames <= 0 || world < 1 || world > 8 || !d_parts || !d_out) return fail(RTB_ERR_ARG, "compose_tiles: bad argument");
    RTB_CUDA(cudaSetDevice(cam->device));
    rtb::ComposeParts parts;
    std::memset(&parts, 0, sizeof parts);
    for (int r = 0; r < world; r++) {
        if (!d_parts[r]) return fail(RTB_ERR_ARG, "compose_tiles: null part");
        parts.part[r] = (const uint32_t*)d_parts[r];
    }
    const int tiles_x = (cam->basis.W + rtb::kTile - 1) / rtb::kTile;
    const long long slots = rtb_tile_major_elements(cam, world) / (rtb::kTile * rtb::kTile);
    if (num_frames > 65535 || cam->basis.H > 65535) return fail(RTB_ERR_ARG, "compose_tiles: too many frames or rows for one launch");
    cudaStream_t s = stream ? (cudaStream_t)stream : cam->stream;
    const dim3 grid((unsigned)(((cam->basis.W + 3) / 4 + 127) / 128), (unsigned)cam->basis.H, (unsigned)num_frames);
    rtb::compose_tiles_kernel<<<grid, 128, 0, s>>>(parts, world, cam->basis.W, cam->basis.H, tiles_x, slots, (uint32_t*)d_out);
    g_launches++;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_selftest_exact(uint64_t seed, int64_t count, uint64_t out4[4]) {
    if (count <= 0 || !out4) return fail(RTB_ERR_ARG, "selftest_exact: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    unsigned long long* d = nullptr;
    RTB_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * 4));
    cudaError_t e = cudaMemset(d, 0, sizeof(unsigned long long) * 4);
    if (e == cudaSuccess) {
        rtb::selftest_exact_kernel<<<(unsigned)((count + 255) / 256), 256>>>(seed, count, d);
        g_launches++;
        e = cudaGetLastError();
    }
    unsigned long long h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RTB_ERR_CUDA, std::string("selftest_exact: ") + cudaGetErrorString(e));
    for (int k = 0; k < 4; k++) out4[k] = h[k];
    return RTB_OK;
}

int rtb_measure_l2_read_bandwidth(size_t bytes, int iters, double* gb_per_s) {
    if (!gb_per_s || bytes < (1u << 20) || iters < 1) return fail(RTB_ERR_ARG, "measure_l2: bad argument");
    RTB_CUDA(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    RTB_CUDA(cudaGetDeviceProperties(&prop, g_device));
    const long long n16 = (long long)(bytes / 16);
    uint4* buf = nullptr;
    uint4* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)n16 * 16);
    if (e == cudaSuccess) e = cudaMalloc(&sink, 16);
    if (e == cudaSuccess) e = cudaMemset(buf, 0x5a, (size_t)n16 * 16);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    float ms = 0.0f;
    if (e == cudaSuccess) {
        const int grid = prop.multiProcessorCount * 8;
        rtb::l2_read_kernel<<<grid, 256>>>(buf, n16, 2, sink);  // warm-up: brings the buffer into L2
        cudaEventRecord(e0);
        rtb::l2_read_kernel<<<grid, 256>>>(buf, n16, iters, sink);
        cudaEventRecord(e1);
        g_launches += 2;
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return fail(RTB_ERR_CUDA, std::string("measure_l2: ") + cudaGetErrorString(e));
    *gb_per_s = (double)n16 * 16.0 * iters / ((double)ms * 1e-3) / 1e9;
    return RTB_OK;
}

int rtb_measure_host_fill_bandwidth(size_t bytes, int threads, double* gb_per_s) {
    if (!gb_per_s || bytes < (1u << 20)) return fail(RTB_ERR_ARG, "measure_host_fill: bad argument");
    // pinned memory where a device exists (that is what the sweep fills); plain memory otherwise, so that the host-side
    // machinery can be exercised without a GPU
    uint32_t* buf = nullptr;
    bool pinned = cudaSetDevice(g_device) == cudaSuccess && cudaMallocHost(&buf, bytes) == cudaSuccess;
    if (!pinned) {
        cudaGetLastError();
        buf = static_cast<uint32_t*>(std::malloc(bytes));
        if (!buf) return fail(RTB_ERR_NOMEM, "measure_host_fill: out of host memory");
    }
    double best = 0.0;
    bool ok = true;
    const size_t words = bytes / 4;
    for (int rep = 0; rep < 4; rep++) {  // the first pass also touches the pages
        const uint32_t value = 0x00f08200u + (uint32_t)rep;
        const auto t0 = std::chrono::steady_clock::now();
        rtb::fill_words(buf + (rep & 1), words - 3, value, threads);  // (odd starts: the unaligned head and tail paths)
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep) best = std::max(best, (double)bytes / s / 1e9);
        for (size_t k = (size_t)(rep & 1); k < (size_t)(rep & 1) + words - 3; k += 4099) ok = ok && buf[k] == value;
        ok = ok && buf[(rep & 1) + words - 4] == value;
    }
    if (pinned) cudaFreeHost(buf); else std::free(buf);
    if (!ok) return fail(RTB_ERR_STATE, "measure_host_fill: the filled buffer does not hold the value");
    *gb_per_s = best;
    return RTB_OK;
}

uint64_t rtb_launch_count(void) { return g_launches.load(); }

int rtb_device_props(int64_t out7[7]) {
    if (!out7) return fail(RTB_ERR_ARG, "device_props: null argument");
    cudaDeviceProp p;
    RTB_CUDA(cudaGetDeviceProperties(&p, g_device));
    int clock_khz = 0, mem_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, g_device);
    cudaDeviceGetAttribute(&mem_khz, cudaDevAttrMemoryClockRate, g_device);
    out7[0] = p.multiProcessorCount; out7[1] = p.l2CacheSize; out7[2] = p.persistingL2CacheMaxSize;
    out7[3] = clock_khz; out7[4] = mem_khz; out7[5] = p.memoryBusWidth; out7[6] = p.major * 10 + p.minor;
    return RTB_OK;
}

}  // extern "C"
