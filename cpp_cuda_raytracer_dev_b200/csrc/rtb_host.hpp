// rtb_host.hpp -- host-side operator surface of the ray-cast path (no CUDA in this header).
//
// Mesh load, the n log n tree build, the camera basis and the object transform recurrence: the
// pieces of the reference that run on the host and feed the per-pixel kernels.  Everything here is
// fp32 with contraction off (the file is compiled with -ffp-contract=off) so that the values handed
// to the GPU are bit-identical to the reference's.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rtb {

struct Vec4 {
    float x = 0, y = 0, z = 0, w = 0;
};

// ---- mesh input (read_ply.cpp:13-152) --------------------------------------------------------
// Returns an empty string on success, otherwise the error text.  points9: 9 floats per triangle.
std::string load_ply(const char* path, int mode, std::vector<float>& points9);
std::string save_ply(const char* path, const float* points9, uint32_t num_tri);
// frame file: .ppm (binary P6) or .png (8-bit RGB, stored deflate); bgra = W*H words 0x00RRGGBB, row 0 at the bottom
std::string save_frame(const char* path, const uint32_t* bgra, int W, int H);
void make_geodesic(int nu, float radius, const float center[3], float displacement, uint32_t seed,
                   std::vector<float>& points9);

// ---- tree (sort.h:11-60, Trixel.h:135-473) ---------------------------------------------------
// Struct-of-arrays form of Trixel::kd_tree::kd_tree_node (Trixel.h:68-79), 2n-1 nodes, BFS order,
// right child = left child + 1.
struct HostTree {
    int64_t num_tri = 0, num_nodes = 0;
    std::vector<float> bounds;     // 6 per node: x0,x1,y0,y1,z0,z1 (kd_leaf order, Trixel.h:31-37)
    std::vector<int32_t> left;     // -1 for leaves
    std::vector<int32_t> tri;      // leaf triangle, -1 for interior nodes
    std::vector<uint8_t> cut_flag; // 0=x1 1=y1 2=z1 3=x0 4=y0 5=z0 (list that was split); leaves inherit
    std::vector<float> s1, s2;     // left child's max / right child's min on the split axis
    double seconds_sort = 0, seconds_partition = 0;
};
// tree cache files (tied to the mesh by a hash of its points); both return the error text or ""
std::string save_tree(const char* path, const HostTree& tree, const float* points9);
std::string load_tree(const char* path, const float* points9, int64_t num_tri, HostTree& out);
// threads <= 0: use all hardware threads
void build_tree(const float* points9, int64_t num_tri, HostTree& out, int threads = 0);
// The same tree built on the current CUDA device from device-resident points (rtb_build.cu) and LEFT THERE: the arrays
// feed the pack kernels of rtb_camera_add_object directly, a host copy is made only when somebody asks for one
// (rtb_mesh_get_tree).  `rec` is the pre-order rank of every interior node among the interior nodes (-1 for leaves):
// the depth-first record index of the render layout, computed during the build from the range sizes.
struct DeviceTree {
    int64_t num_tri = 0, num_nodes = 0;
    float* bounds = nullptr;      // 6 per node
    int* left = nullptr;
    int* tri = nullptr;
    unsigned char* cut = nullptr;
    float* s1 = nullptr;
    float* s2 = nullptr;
    int* rec = nullptr;
    void* arena = nullptr;        // the one allocation all seven arrays live in
    float root_bounds[6] = {0, 0, 0, 0, 0, 0};
    int root_tri = -1;
    double seconds_sort = 0, seconds_partition = 0;
    int launches = 0;             // kernels the build launched (counted into rtb_launch_count)
};
// all three return the error text or ""
std::string build_tree_gpu(const float* d_points9, int64_t num_tri, DeviceTree& out);
std::string download_tree(const DeviceTree& in, HostTree& out);
void free_device_tree(DeviceTree& t);

// host-side fill of a frame buffer (background pre-fill of the sweep's host frames), threads <= 0: all hardware threads
void fill_words(uint32_t* dst, size_t count, uint32_t value, int threads);
void fill_words2(uint32_t* a, uint32_t value_a, uint32_t* b, uint32_t value_b, size_t count, int threads);  // either may be null

// ---- camera (Camera.cpp:5-67) ----------------------------------------------------------------
struct CameraBasis {
    int32_t W = 0, H = 0;
    float pos[3], n[3], v[3], u[3], n_mod[3], v_mod[3], u_mod[3];
    float draw_distance = 400.0f;           // Camera.cpp:70
    uint8_t background[4] = {240, 130, 0, 0}; // r,g,b,a  Camera.cpp:72
};
void camera_basis(int32_t W, int32_t H, float f_w, float f_h, float fclen, const float pos[3], const float la[3],
                  const float up[3], CameraBasis& out);

// ---- object transform (Quaternion.cpp, vector.cpp:38-65, Camera.cu:254-335) -------------------
struct Transform {
    Vec4 quat;       // Quaternion::vec, never renormalised
    Vec4 row[3];     // Quaternion::rot_m x,y,z; .w = translation column
    Vec4 init_face;  // Object::init_face (Camera.cpp:131)
    Vec4 cur_face;   // Object::cur_face  (Camera.cpp:132)
    void reset(const float cam_pos[3]);
    // select: 10/11 rotate by step quaternion (x,y,z,w); 30/31/32 translate along (x,y,z) by w
    bool apply(uint8_t select, float x, float y, float z, float w);
    void matrix(float m12[12]) const;
    void set_matrix(const float m12[12]);
};

}  // namespace rtb
