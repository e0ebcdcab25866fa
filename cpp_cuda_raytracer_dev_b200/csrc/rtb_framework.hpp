// rtb_framework.hpp -- the reference's host classes, re-created over the C ABI (include/rtb.h).
//
// Same class names, method names, argument meaning and call order as the reference application
// (TEST_Dungeonrun/framework.h), so that the call sequence of WinMain.cpp:69-237 ports line for line
// to a headless Linux driver (rtb_render_main.cpp).  Header-only; links against librtb.so.  Methods
// return the library's status code where the reference returns cudaError_t (0 = success) and never
// throw, exactly like the reference.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtb.h"

namespace rtbfw {

typedef float T_fp;        // typedefs.h:15
typedef uint32_t T_uint;   // typedefs.h:16
typedef uint8_t u8;

constexpr u8 SET_COLOR_TAG = RTB_SET_COLOR_TAG;      // Camera.h:13
constexpr u8 PHONG_COLOR_TAG = RTB_PHONG_COLOR_TAG;  // Camera.h:14
constexpr int TRANSLATE_XYZ = RTB_TRANSLATE_XYZ, TRANSLATE_X = RTB_TRANSLATE_X, TRANSLATE_Z = RTB_TRANSLATE_Z;  // platform_common.h:16-18
constexpr int ROTATE_TRI_PY = RTB_ROTATE_TRI_PY, ROTATE_TRI_NY = RTB_ROTATE_TRI_NY;                              // platform_common.h:20-21

template <typename T>
struct VEC3 { T x, y, z; };  // Vector.h:29
template <typename T>
struct VEC4 { T x, y, z, w; };  // Vector.h:64

// Color.h:4 -- per-triangle radiance table handed to Trixel
struct Color {
    struct radiance { T_fp r, g, b; };
    uint32_t* c = nullptr;
    radiance* rad = nullptr;
};

// read_ply.cpp:13.  kd_leaf_sort / kd_vertex outputs of the reference are internal to the build here
// (the per-triangle AABB lists are derived from the points), so only points and the count are returned.
inline int read_ply(const char* file_name, T_fp** points_list, T_uint* num_tri, u8 mode) {
    return rtb_read_ply(file_name, mode == 255 ? -1 : (int)mode, points_list, num_tri);
}

// Input.h:2 -- carries the pending transform (quaternion step or direction + distance)
class Input {
public:
    VEC4<T_fp> t_vec{0, 0, 0, 0};
    void set_vec(T_fp x, T_fp y, T_fp z, T_fp w) { t_vec = VEC4<T_fp>{x, y, z, w}; }
    void set_quat(T_fp x, T_fp y, T_fp z, T_fp w) { set_vec(x, y, z, w); }  // Input.cpp:15
};

class Camera;

// Trixel.h:39 -- mesh + tree
class Trixel {
public:
    rtb_mesh* handle = nullptr;
    int64_t num_trixels = 0, num_voxels = 0;
    Trixel(int64_t num_t, const T_fp* points_data, const Color* color_data) {  // Trixel.h:87
        num_trixels = num_t;
        num_voxels = 2 * num_t - 1;
        rtb_mesh_create(points_data, num_t, color_data && color_data->rad ? &color_data->rad[0].r : nullptr, nullptr, &handle);
    }
    ~Trixel() { rtb_mesh_destroy(handle); }
    // Trixel.h:386 -- the six sorted lists are produced inside create_kd(); kept for call-order parity
    int set_sorted_voxels(const void* /*kd_leaf_list*/, T_uint /*num*/) { return 0; }
    int create_kd() { return rtb_mesh_build_tree(handle); }  // Trixel.h:135
};

// Object.h:6 -- a mesh instance with its transform
class Object {
public:
    rtb_object* handle = nullptr;
    Trixel* trixel_list = nullptr;
    explicit Object(Trixel* t) : trixel_list(t) { rtb_object_create(t->handle, &handle); }  // Object.cpp:4
    ~Object() { rtb_object_destroy(handle); }
    int transform(Input* in, u8 transform_select) {  // Object.cpp:14
        const float q[4] = {in->t_vec.x, in->t_vec.y, in->t_vec.z, in->t_vec.w};
        return rtb_object_transform(handle, q, transform_select);
    }
    int render(Camera* c);  // Object.cpp:10
};

// Camera.h:15
class Camera {
public:
    rtb_camera* handle = nullptr;
    struct { struct { uint32_t w, h; uint64_t count; } res; } f_prop;                  // Camera.h:22-27
    struct { VEC3<T_fp> pos, n, v, u, n_mod, v_mod, u_mod; } o_prop;                    // Camera.h:32-42
    struct { const uint32_t* c = nullptr; const int32_t* id = nullptr; } h_color;       // h_mem.h_color.c (+ hit ids)
    Camera(int32_t r_w, int32_t r_h, T_fp f_w, T_fp f_h, T_fp fclen, T_fp p_x, T_fp p_y, T_fp p_z, T_fp la_x, T_fp la_y, T_fp la_z,
           T_fp up_x, T_fp up_y, T_fp up_z) {  // Camera.h:86
        const float pos[3] = {p_x, p_y, p_z}, la[3] = {la_x, la_y, la_z}, up[3] = {up_x, up_y, up_z};
        f_prop.res.w = (uint32_t)r_w; f_prop.res.h = (uint32_t)r_h; f_prop.res.count = (uint64_t)r_w * (uint64_t)r_h;
        rtb_camera_create(r_w, r_h, f_w, f_h, fclen, pos, la, up, &handle);
        float b[18];
        if (handle && rtb_camera_get_basis(handle, b) == 0) {
            o_prop.pos = VEC3<T_fp>{p_x, p_y, p_z};
            VEC3<T_fp>* dst[6] = {&o_prop.n, &o_prop.v, &o_prop.u, &o_prop.n_mod, &o_prop.v_mod, &o_prop.u_mod};
            for (int k = 0; k < 6; k++) *dst[k] = VEC3<T_fp>{b[3 * k], b[3 * k + 1], b[3 * k + 2]};
        }
        refresh();
    }
    ~Camera() { rtb_camera_destroy(handle); }
    int add_object(Object* o) { return rtb_camera_add_object(handle, o->handle); }  // Camera.cpp:118
    // Camera::render() (Camera.cpp:160-163, a stub in the reference): every object of object_list in one frame, with the
    // light list, shadow test and sample_rate the reference leaves dormant (include/rtb.h, scene extension)
    int render() { return rtb_camera_render_scene(handle, RTB_RENDER_DEFAULT); }
    struct render_properites { int sample_rate = 0; } r_prop;                          // Camera.h:44-48
    int set_sample_rate(int n) { r_prop.sample_rate = n; return rtb_camera_set_sample_rate(handle, n); }
    int set_lights(int n, const float* xyz3) { return rtb_camera_set_lights(handle, n, xyz3); }
    int set_shadows(bool on) { return rtb_camera_set_shadows(handle, on ? 1 : 0); }
    int color_pixels(u8 tag) {                                                       // Camera.cpp:229
        const int rc = rtb_camera_color_pixels(handle, tag);
        refresh();
        return rc;
    }
private:
    void refresh() { h_color.c = rtb_camera_host_color(handle); h_color.id = rtb_camera_host_ids(handle); }
};

inline int Object::render(Camera* c) { return rtb_object_render(handle, c->handle, RTB_RENDER_DEFAULT); }

// Headless replacement of the GDI blit (WinMain.cpp:217): the frame is bottom-up 0x00RRGGBB words.
// path ending in .png -> PNG, otherwise binary PPM (rtb_write_frame)
inline bool write_ppm(const std::string& path, const uint32_t* frame, uint32_t w, uint32_t h) {
    return rtb_write_frame(path.c_str(), frame, (int32_t)w, (int32_t)h) == 0;
}
inline bool write_raw_ids(const std::string& path, const int32_t* ids, uint64_t count) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(ids, sizeof(int32_t), (size_t)count, f);
    std::fclose(f);
    return true;
}

}  // namespace rtbfw
