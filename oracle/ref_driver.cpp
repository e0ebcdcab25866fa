// TEST INFRASTRUCTURE ONLY -- headless driver around the UNMODIFIED reference classes.
//
// This file is compiled together with the reference's own sources (read in place from
// /root/reference/TEST_Dungeonrun by oracle/build_ref.py, through the host emulation shim
// in oracle/shim/) into oracle/_ref/libref_emu.so.  It replays the reference's own call
// sequence (WinMain.cpp:69-74 camera, :93 read_ply, :114-121 colours, :134-135 Trixel +
// set_sorted_voxels, :144 create_kd, :152-155 Object + add_object, :187-208 set_quat +
// transform, :212 render, :237 color_pixels) without the Win32 window, and hands the
// reference's own buffers back through a flat C interface for tests/ and bench.py's
// reference arm.  It is the ground truth the C restatement (oracle/rtb_oracle.c) and the
// CUDA path are pinned against.  Nothing in the product links or loads it.
//
// The same file also serves the second build of oracle/build_ref.py, RTB_REF_CUDA: the reference's
// three .cu files compiled by nvcc for sm_100a and run on the B200 itself (oracle/_ref/libref_cuda*.so),
// which gives the reference's own GPU kernels as a second ground truth and as "the kernel to beat".
#include "framework.h"
#include "sort.h"
#ifdef RTB_REF_CUDA
#include <cuda_runtime.h>
#else
#include <omp.h>

thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif
// Third build, RTB_REF_SEAM (oracle/build_ref.py, build_seam): the reference's HOST classes only, with its three .cu
// files replaced by integration/rtb_seam.cpp on top of librtb.so -- the drop-in a maintainer would make.
#ifdef RTB_REF_SEAM
extern "C" const int* rtb_seam_host_ids(Camera* c);
extern "C" int rtb_seam_object_matrix(Object* o, float m12[12]);
#endif

void read_ply(const char* file_name, T_fp** points_list, T_uint* num_tri, kd_leaf_sort** leaf_list,
              kd_vertex** vertex_list, T_uint* num_vert, u8 mode);

namespace {
struct RefScene {
    Camera* cam = nullptr;
    Trixel* trixels = nullptr;
    Object* obj = nullptr;   // WinMain's obj1: the one it transforms and renders (WinMain.cpp:188-208,212)
    Object* obj2 = nullptr;  // WinMain's obj2 (WinMain.cpp:153,156): registered, never drawn (:214 is commented out)
    Input* input = nullptr;
    T_fp* points = nullptr;
    T_uint ntri = 0;
};
int g_winmain_objects = 1;  // ref_set_objects: 2 = register two objects over the mesh like WinMain.cpp:152-156
float fmin3(float a, float b, float c) { return min(a, min(b, c)); }
float fmax3(float a, float b, float c) { return max(a, max(b, c)); }
}  // namespace

extern "C" {

struct ref_node {
    long long left, right, tri, parent;
    int cut_flag, is_leaf;
    float x0, x1, y0, y1, z0, z1, s1, s2;
};

// cam14 = f_w, f_h, fclen, pos.xyz, la.xyz, up.xyz, (2 unused)   (WinMain.cpp:69-74)
// Either ply_path != NULL (reference read_ply with `mode`) or points9/ntri (already
// triangulated soup, 9 floats per triangle in the loader's output order).
void* ref_open(const char* ply_path, int mode, const float* points9, long ntri_in, int W, int H, const float* cam14,
               const float* rgb) {
    RefScene* s = new RefScene();
    s->cam = new Camera(W, H, cam14[0], cam14[1], cam14[2], cam14[3], cam14[4], cam14[5], cam14[6], cam14[7], cam14[8],
                        cam14[9], cam14[10], cam14[11]);
    kd_leaf_sort* leafs = NULL;
    kd_vertex* verts = NULL;
    T_uint ntri = 0, nvert = 0;
    if (ply_path) {
        read_ply(ply_path, &s->points, &ntri, &leafs, &verts, &nvert, (u8)mode);
    } else {
        ntri = (T_uint)ntri_in;
        s->points = (T_fp*)malloc(sizeof(T_fp) * 9 * ntri);
        memcpy(s->points, points9, sizeof(T_fp) * 9 * ntri);
        leafs = (kd_leaf_sort*)calloc(ntri, sizeof(kd_leaf_sort));
        for (T_uint i = 0; i < ntri; i++) {  // same rule as read_ply.cpp:127-134
            const float* p = points9 + 9 * (size_t)i;
            leafs[i].x0 = fmin3(p[0], p[3], p[6]); leafs[i].x1 = fmax3(p[0], p[3], p[6]);
            leafs[i].y0 = fmin3(p[1], p[4], p[7]); leafs[i].y1 = fmax3(p[1], p[4], p[7]);
            leafs[i].z0 = fmin3(p[2], p[5], p[8]); leafs[i].z1 = fmax3(p[2], p[5], p[8]);
            leafs[i].tri_list_index = i;
        }
    }
    s->ntri = ntri;
    Color colors;
    colors.c = (u32*)calloc(ntri, sizeof(u32));
    colors.rad = (Color::radiance*)malloc(sizeof(Color::radiance) * ntri);
    for (T_uint i = 0; i < ntri; i++) { colors.rad[i].r = rgb[0]; colors.rad[i].g = rgb[1]; colors.rad[i].b = rgb[2]; }
    s->trixels = new Trixel(ntri, s->points, &colors);
    s->trixels->set_sorted_voxels(leafs, ntri);
    s->trixels->create_kd();
    s->obj = new Object(s->trixels);               // WinMain.cpp:152
    if (g_winmain_objects >= 2) s->obj2 = new Object(s->trixels);  // :153
    s->cam->add_object(s->obj);                    // :155
    if (s->obj2) s->cam->add_object(s->obj2);      // :156
    s->input = new Input();
    return s;
}

// 1 (default): one object, the smallest scene; 2: the literal WinMain.cpp:152-156 sequence -- two objects over one
// mesh, both added to the camera, the first one transformed and rendered.  Applies to scenes opened afterwards.
void ref_set_objects(int n) { g_winmain_objects = n; }

long ref_num_tris(void* h) { return (long)((RefScene*)h)->ntri; }
long ref_num_nodes(void* h) { return (long)((RefScene*)h)->trixels->num_voxels; }

void ref_get_points(void* h, float* out) {
    RefScene* s = (RefScene*)h;
    memcpy(out, s->points, sizeof(float) * 9 * (size_t)s->ntri);
}

void ref_get_nodes(void* h, ref_node* out) {
    RefScene* s = (RefScene*)h;
    for (long long i = 0; i < s->trixels->num_voxels; i++) {
        const Trixel::kd_tree::kd_tree_node& n = s->trixels->h_tree.h_nodes[i];
        out[i].left = n.left_node; out[i].right = n.right_node; out[i].tri = n.tri_index; out[i].parent = n.parent;
        out[i].cut_flag = n.cut_flag; out[i].is_leaf = n.is_leaf;
        out[i].x0 = n.h_bound.x0; out[i].x1 = n.h_bound.x1; out[i].y0 = n.h_bound.y0; out[i].y1 = n.h_bound.y1;
        out[i].z0 = n.h_bound.z0; out[i].z1 = n.h_bound.z1; out[i].s1 = n.s1; out[i].s2 = n.s2;
    }
}

// out18 = n, v, u, n_mod, v_mod, u_mod (Camera.cpp:32-67)
void ref_get_camera(void* h, float* out18) {
    Camera* c = ((RefScene*)h)->cam;
    VEC3<T_fp>* v[6] = {&c->o_prop.n, &c->o_prop.v, &c->o_prop.u, &c->o_prop.n_mod, &c->o_prop.v_mod, &c->o_prop.u_mod};
    for (int i = 0; i < 6; i++) { out18[3 * i] = v[i]->x; out18[3 * i + 1] = v[i]->y; out18[3 * i + 2] = v[i]->z; }
}

// per-pixel primary ray table written by init_cam_mem_cuda (Camera.cu:89-111), 3 floats/pixel
void ref_get_rays(void* h, float* out3p) {
    Camera* c = ((RefScene*)h)->cam;
    const u64 P = c->f_prop.res.count;
    float* tmp = (float*)malloc(sizeof(float) * 3 * P);  // the arrays live in "device" memory
    cudaMemcpy(tmp, c->h_mem.rmd.x, sizeof(float) * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(tmp + P, c->h_mem.rmd.y, sizeof(float) * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(tmp + 2 * P, c->h_mem.rmd.z, sizeof(float) * P, cudaMemcpyDeviceToHost);
    for (u64 i = 0; i < P; i++) { out3p[3 * i] = tmp[i]; out3p[3 * i + 1] = tmp[P + i]; out3p[3 * i + 2] = tmp[2 * P + i]; }
    free(tmp);
}

// Input::set_quat(x,y,z,w) then Object::transform(input, select)  (WinMain.cpp:186-209)
void ref_transform(void* h, int select, float x, float y, float z, float w) {
    RefScene* s = (RefScene*)h;
    s->input->set_quat(x, y, z, w);
    s->obj->transform(s->input, (u8)select);
}

// 12 floats: rows x,y,z of the object matrix, each (i, j, k, w=translation)  (Quaternion.h:12)
void ref_get_matrix(void* h, float* m12) {
#ifdef RTB_REF_SEAM
    rtb_seam_object_matrix(((RefScene*)h)->obj, m12);  // the recurrence runs inside librtb.so (rtb_object_transform)
    return;
#endif
    Quaternion* q = ((RefScene*)h)->obj->quat;
    VEC4<T_fp>* rows[3] = {q->rot_m->x, q->rot_m->y, q->rot_m->z};
    for (int r = 0; r < 3; r++) { m12[4 * r] = rows[r]->i; m12[4 * r + 1] = rows[r]->j; m12[4 * r + 2] = rows[r]->k; m12[4 * r + 3] = rows[r]->w; }
}

// One clean frame: Object::render (intersect_voxel_cuda), then color_pixels(SET_COLOR_TAG), i.e.
// set_cam_cuda background fill FOLLOWED BY color_cam_cuda (the reference's own fall-through,
// Camera.cu:77-82), so the buffer is background + Phong of exactly this frame's hits.
void ref_render(void* h, long long* ids, unsigned* bgra) {
    RefScene* s = (RefScene*)h;
    s->obj->render(s->cam);
    s->cam->color_pixels(SET_COLOR_TAG);
    const u64 P = s->cam->f_prop.res.count;
#ifdef RTB_REF_SEAM
    if (ids) { const int* src = rtb_seam_host_ids(s->cam); for (u64 i = 0; i < P; i++) ids[i] = src[i]; }
#else
    if (ids) cudaMemcpy(ids, s->cam->h_mem.d_rmi.index, sizeof(long long) * P, cudaMemcpyDeviceToHost);
#endif
    if (bgra) memcpy(bgra, s->cam->h_mem.h_color.c, sizeof(unsigned) * P);
}

// traversal only (timed by bench.py --impl reference together with shade)
void ref_render_nocopy(void* h) {
    RefScene* s = (RefScene*)h;
    s->obj->render(s->cam);
    s->cam->color_pixels(SET_COLOR_TAG);
}

#ifdef RTB_REF_CUDA
int ref_threads(void) { return 0; }  // runs on the GPU
void ref_set_threads(int) {}
int ref_is_cuda(void) { return 1; }
#else
int ref_threads(void) { return omp_get_max_threads(); }
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_is_cuda(void) { return 0; }
#endif
#ifdef RTB_REF_SEAM
int ref_is_seam(void) { return 1; }
#else
int ref_is_seam(void) { return 0; }
#endif
}
