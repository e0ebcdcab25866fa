// rtb_seam.cpp -- the translation unit a maintainer of ams3878/cpp_cuda_raytracer_dev adds to the project INSTEAD OF
// Trixel.cu, Camera.cu and Quaternion.cu.  It defines the reference's seven host<->CUDA seam functions
// and the two Quaternion device methods (Trixel.h:13-14, Camera.h:99-103, Quaternion.h:8,20) on top of librtb.so, so the
// reference's own host classes (Camera, Trixel, Object, Quaternion, Input) drive the B200 path unchanged.
//
// This file is compiled and RUN by the test-suite: oracle/build_ref.py (build_seam) compiles it together with the
// reference's own host sources, read in place from /root/reference, into oracle/_ref/libref_seam.so, and
// tests/test_gpu_parity.py::test_reference_host_classes_over_librtb renders through the reference's classes.
// It contains no reference code: only calls into include/rtb.h, keyed by the reference objects' addresses.
#include <unordered_map>  // (before framework.h: <windows.h> defines min/max macros)
#include <vector>

#include "framework.h"
#include "rtb.h"

namespace {
struct MeshEntry { rtb_mesh* mesh = nullptr; bool tree_built = false; };
std::unordered_map<Trixel*, MeshEntry> g_mesh;
std::unordered_map<Camera*, rtb_camera*> g_cam;
std::unordered_map<Object*, rtb_object*> g_obj;
// Object::render hands the seam the object's OWN quaternion (Object.cpp:10-12 -> Trixel.h:474-476 -> Trixel.cu:210), which
// Camera::add_object created just before it called init_camera_voxel_device_memory (Camera.cpp:134,139): the quaternion's
// address identifies the object to draw.  WinMain registers two objects and renders / transforms the first
// (WinMain.cpp:152-156,186-189,212).
std::unordered_map<Quaternion*, rtb_object*> g_obj_of_quat;

cudaError_t report(int rc, const char* who) {  // the reference prints and carries on (vector.cuh:15-18); its callers ignore the value
    if (rc) printf("%s failed: %s\n", who, rtb_last_error());
    return (cudaError_t)rc;
}

// Camera::Camera keeps only the pixel size (f_w / r_w, Camera.cpp:16-17; film.h is never set, :14-15).  Recover a film
// size whose quotient is exactly that pixel size, so that the camera basis comes out bit-identical.
float film_from_pixel(float pix, unsigned n) {
    const float guess = pix * (float)n;
    for (int k = 0; k < 9; k++) {
        const int step = (k + 1) / 2 * ((k & 1) ? 1 : -1);  // 0, +1, -1, +2, -2, ... ulps
        union { float f; int i; } u;
        u.f = guess;
        u.i += step;
        if (u.f / (float)n == pix) return u.f;
    }
    return guess;
}
}  // namespace

extern "C" cudaError_t init_trixels_device_memory(Trixel* t) {  // Trixel.cu:266, called by the Trixel constructor (Trixel.h:132)
    // the constructor has copied the colours to t->h_mem.d_color.rad (Trixel.h:130); the points have a host copy (:124-126)
    std::vector<Color::radiance> rad((size_t)t->num_trixels);
    cudaMemcpy(rad.data(), t->h_mem.d_color.rad, sizeof(Color::radiance) * (size_t)t->num_trixels, cudaMemcpyDeviceToHost);
    rtb_mesh* m = nullptr;
    const int rc = rtb_mesh_create(t->h_points_init_data, t->num_trixels, &rad[0].r, nullptr, &m);
    g_mesh[t].mesh = m;
    g_mesh[t].tree_built = false;
    return report(rc, "init_trixels_device_memory");
}

extern "C" cudaError_t init_camera_device_memory(Camera* c) {  // Camera.cu:112, called by the Camera constructor (Camera.cpp:116)
    rtb_camera* h = nullptr;
    const float pos[3] = {c->o_prop.pos.x, c->o_prop.pos.y, c->o_prop.pos.z}, la[3] = {c->o_prop.la.x, c->o_prop.la.y, c->o_prop.la.z},
                up[3] = {c->o_prop.up.x, c->o_prop.up.y, c->o_prop.up.z};
    const int rc = rtb_camera_create((int32_t)c->f_prop.res.w, (int32_t)c->f_prop.res.h, film_from_pixel(c->f_prop.pix.w, c->f_prop.res.w),
                                     film_from_pixel(c->f_prop.pix.h, c->f_prop.res.h), c->l_prop.focal_length, pos, la, up, &h);
    g_cam[c] = h;
    if (h) {
        free(c->h_mem.h_color.c);                                   // Camera.cpp:79
        c->h_mem.h_color.c = (u32*)rtb_camera_host_color(h);        // the window blit reads this pointer (WinMain.cpp:217)
    }
    return report(rc, "init_camera_device_memory");
}

extern "C" cudaError_t init_camera_trixel_device_memory(Trixel*, Camera*) { return cudaSuccess; }  // Trixel.cu:244: done by the next one

extern "C" cudaError_t init_camera_voxel_device_memory(Trixel* t, Camera* c) {  // Camera.cu:163, via Camera::add_object (Camera.cpp:139,208)
    Object* o = c->object_list[c->num_objects - 1];                              // the object being added (Camera.cpp:118-130)
    MeshEntry& e = g_mesh[t];
    // Trixel::create_kd has built the reference's host tree by now; the identical tree is rebuilt on the GPU in
    // milliseconds, once per Trixel however many objects instance it (a maintainer may instead turn create_kd itself
    // into rtb_mesh_build_tree and skip the host build)
    int rc = e.tree_built ? 0 : rtb_mesh_build_tree(e.mesh);
    e.tree_built = e.tree_built || rc == 0;
    rtb_object* h = nullptr;
    if (!rc) rc = rtb_object_create(e.mesh, &h);
    if (!rc) rc = rtb_camera_add_object(g_cam[c], h);  // objects of one mesh share the camera-side arrays inside librtb
    g_obj[o] = h;
    g_obj_of_quat[o->quat] = h;                        // assigned at Camera.cpp:134, five lines before this call
    return report(rc, "init_camera_voxel_device_memory");
}

extern "C" cudaError_t transform_camera_voxel_device_memory(Object* o, VEC4<T_fp>* tv, Quaternion*, u8 select) {  // Camera.cu:254
    const float v[4] = {tv->x, tv->y, tv->z, tv->w};  // Input::set_quat stored the step here (Input.cpp:7-20)
    return report(rtb_object_transform(g_obj[o], v, select), "transform_camera_voxel_device_memory");
}

cudaError_t intersect_trixels_device(Trixel*, Camera* c, Quaternion* q, u32) {  // Trixel.cu:210 (C++ linkage, Trixel.h:13)
    // the reference renders with the matrix of the quaternion it is handed (Trixel.cu:222 `q->d_rot_m`): that names the object
    const auto it = g_obj_of_quat.find(q);
    if (it == g_obj_of_quat.end()) {
        printf("intersect_trixels_device failed: the quaternion does not belong to an object added to a camera\n");
        return (cudaError_t)RTB_ERR_STATE;
    }
    return report(rtb_object_render(it->second, g_cam[c], RTB_RENDER_DEFAULT), "intersect_trixels_device");
}

extern "C" cudaError_t color_camera_device(Camera* c, u8 tag) {  // Camera.cu:70
    const int rc = rtb_camera_color_pixels(g_cam[c], tag);
    c->h_mem.h_color.c = (u32*)rtb_camera_host_color(g_cam[c]);
    return report(rc, "color_camera_device");
}

cudaError_t Quaternion::set_device_rotation(VEC4<VEC4<T_fp>*>*) { return cudaSuccess; }                    // Quaternion.cu:21
cudaError_t Quaternion::initialize_CUDA(VEC4<T_fp>*, VEC4<T_fp>*, VEC4<T_fp>*) { return cudaSuccess; }     // Quaternion.cu:27

// (not part of the reference's seam) what the new path offers beyond it, for code that wants it
extern "C" const int32_t* rtb_seam_host_ids(Camera* c) { return rtb_camera_host_ids(g_cam[c]); }
extern "C" int rtb_seam_object_matrix(Object* o, float m12[12]) { return rtb_object_get_matrix(g_obj[o], m12); }
