/* rtb.h -- C ABI of the B200-native ray-cast path (librtb.so).
 *
 * Drop-in boundary for the per-pixel ray-cast hot path of ams3878/cpp_cuda_raytracer_dev
 * (primary-ray generation -> tree traversal -> Moller-Trumbore -> Phong -> framebuffer) and the
 * operator surface that feeds it (read_ply, the n log n tree build, the camera/quaternion
 * transform API, render-frame-to-buffer).  The reference's own host<->CUDA seam is seven free
 * functions plus two Quaternion methods that take pointers to C++ objects and read their fields
 * (C linkage, not C layout).  This header exports the same operations over
 * opaque handles with plain pointers and sizes.  Every entry point cites the reference interface
 * (file:line under TEST_Dungeonrun/) it replaces.
 *
 * Conventions kept from the reference: every call returns a status code and never throws
 * (0 = success, otherwise an rtb_status; rtb_last_error() gives the text that the reference
 * would have printf'd); caller-owned inputs are copied; rendered frames live in LIBRARY-OWNED
 * pinned host buffers that stay valid until the next render on the same camera
 * (Camera::h_mem.h_color.c, Camera.cpp:79); one host thread per handle, calls are synchronous
 * unless the name says _async.  Unlike the reference every handle has an explicit destroy.
 *
 * There is no CPU fallback: every call that needs the GPU fails with RTB_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef RTB_H
#define RTB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtb_mesh rtb_mesh;     /* = class Trixel   (Trixel.h:39)  : triangles + tree      */
typedef struct rtb_camera rtb_camera; /* = class Camera   (Camera.h:15)  : film, rays, frame      */
typedef struct rtb_object rtb_object; /* = class Object   (Object.h:6)   : mesh instance + its
                                         Quaternion (Quaternion.h:5) and the per-(camera, mesh)
                                         device arrays Camera::trixel_memory / voxel_memory      */

typedef enum {
    RTB_OK = 0,
    RTB_ERR_ARG = 1,    /* bad argument / handle state */
    RTB_ERR_IO = 2,     /* file could not be read or parsed */
    RTB_ERR_CUDA = 3,   /* CUDA runtime error, or no usable device */
    RTB_ERR_NOMEM = 4,
    RTB_ERR_STATE = 5   /* call order violated (e.g. render before build_tree / add_object) */
} rtb_status;

/* selectors, same values as the reference */
#define RTB_SET_COLOR_TAG 1   /* Camera.h:13 */
#define RTB_PHONG_COLOR_TAG 2 /* Camera.h:14 */
#define RTB_TRANSLATE_XYZ 30  /* platform_common.h:16 */
#define RTB_TRANSLATE_X 31    /* platform_common.h:17 */
#define RTB_TRANSLATE_Z 32    /* platform_common.h:18 */
#define RTB_ROTATE_TRI_PY 10  /* platform_common.h:20 */
#define RTB_ROTATE_TRI_NY 11  /* platform_common.h:21 */

/* render flags */
#define RTB_RENDER_DEFAULT 0u
#define RTB_RENDER_NO_CULL 1u   /* visit exactly the node sequence of Trixel.cu:70-170 (no distance culling) */
#define RTB_RENDER_COUNTERS 2u  /* accumulate per-launch work counters (rtb_camera_counters) */
#define RTB_RENDER_PUSH_PREFILLED 8u /* rtb_render_frames_push_async: the owner pre-filled the frames (rtb_fill_frames_device_async): background-only work units are not sent */
#define RTB_RENDER_TILE_MAJOR 4u /* rtb_render_frames_device_async: compact tile-major output (multi-GPU exchange format) */

const char* rtb_last_error(void);
const char* rtb_version(void);
/* number of CUDA devices visible / select the device used by handles created afterwards on this
 * thread.  Replaces the hard-coded cudaSetDevice(0) of Trixel.cu:213,247,269, Camera.cu:73,115,166. */
int rtb_device_count(void);
int rtb_set_device(int device);
/* Scheduling knobs of the render kernel, process-wide (development / tuning; the defaults are the measured best and are
 * also read once from the environment as RTB_<NAME IN CAPITALS>): "unit_shift" log2 pixels per work unit (5..10, 0 = from
 * the launch size), "t_active" (refill when no more lanes than this still traverse), "t_leaf" (leaf step when at least
 * this many lanes wait at a leaf), "tail5" / "tail6" (trailing frames of a launch handed out in 32- / 64-pixel units,
 * -1 = automatic), "reserve_sms" (SMs left free beside the persistent kernel), "l2_window" (persisting L2 window:
 * 0 off, 1 node records, 2 nodes + triangles; applied at the next rtb_camera_add_object), "l2_carve_mb" (persisting
 * carve-out in MB, 0 = the size of the window), "no_rect" (1: no root-box rectangle, every pixel is traced),
 * "frame_order" (1: a multi-frame launch works through its frames sorted by viewing direction), "host_direct" (1:
 * rtb_object_render stores its frame straight into the camera's host buffers; 0: device frame + copy), "lookahead" (frames
 * rendered ahead of the current one while the transform steps between two renders repeat, 0..2), "steal_spin" /
 * "t_active_inline" / "inline_prefetch" (single-frame launches: iterations between two looks for lanes to share a long ray
 * with, their t_active, child-record prefetch), "sweep_direct" (1: rtb_render_sweep into pinned buffers pre-fills the
 * background on the host and lets the kernel store the rest), "host_fill_threads", "sweep_chunk_mb" (threads and chunk
 * size of that pre-fill). */
int rtb_set_knob(const char* name, int value);

/* ---- mesh input -------------------------------------------------------------------------- */

/* read_ply (read_ply.cpp:13): ASCII PLY subset; mode 0/1/2 = 3/5/6 numbers per vertex
 * (read_ply.cpp:52-65); "3 a b c" faces are stored (c,a,b), "4 a b c d" faces become (a,b,c),(a,c,d).
 * mode -1 (extension) additionally accepts `format binary_little_endian` files such as the
 * reference's 3_walls.ply, applying the same face rules.  *points9 (9 floats per triangle, malloc'd,
 * release with rtb_free) and *num_tri are outputs; per-triangle AABBs (the reference's kd_leaf_sort
 * list) are derived inside rtb_mesh_create. */
int rtb_read_ply(const char* file_name, int mode, float** points9, uint32_t* num_tri);
void rtb_free(void* p);
/* utility (not in the reference): write a triangle soup as an ASCII PLY that rtb_read_ply(mode 0)
 * and the reference loader read back to the same triangles (faces are written rotated so that the
 * loader's (c,a,b) storage restores the input order). */
int rtb_write_ply(const char* file_name, const float* points9, uint32_t num_tri);
/* The headless replacement of the reference's window blit (StretchDIBits, WinMain.cpp:217): write a frame buffer
 * (W*H words 0x00RRGGBB, row 0 at the bottom, Camera.h:35) as .ppm (binary P6) or .png (8-bit RGB), chosen by the
 * file name's extension; rows are flipped so the image is upright. */
int rtb_write_frame(const char* file_name, const uint32_t* bgra, int32_t width, int32_t height);
/* utility (not in the reference): displaced geodesic icosphere with 20*nu*nu triangles, the
 * labelled stand-in for the Stanford meshes that are absent from the reference checkout
 *  and the synthetic 10M-triangle mesh of BASELINE.json configs[4]. */
int rtb_mesh_geodesic(int nu, float radius, const float center[3], float displacement, uint32_t seed,
                      float** points9, uint32_t* num_tri);

/* ---- mesh + tree ------------------------------------------------------------------------- */

/* Trixel::Trixel(num_t, points_data, color_data) + init_trixels_device_memory (Trixel.h:87,
 * Trixel.cu:266).  rad3 = per-triangle radiance r,g,b (3 floats each, Color::radiance) or NULL
 * with uniform_rgb (the reference app uses (.1,.55,.2) for all, WinMain.cpp:118-120). */
int rtb_mesh_create(const float* points9, int64_t num_tri, const float* rad3, const float uniform_rgb[3], rtb_mesh** out);
/* Trixel::set_sorted_voxels + Trixel::create_kd (Trixel.h:386, Trixel.h:135; sort.h:11): six sorted
 * AABB lists, object-median split, BFS numbering, 2n-1 nodes.  Same tree as the reference, bit for bit. */
int rtb_mesh_build_tree(rtb_mesh* mesh);
/* same, choosing where the build runs: 0 = on the GPU when the mesh has a device copy (default), 1 = host threads,
 * 2 = GPU (radix-sorted lists + level-synchronous partition, csrc/rtb_build.cu).  Both produce the identical tree. */
int rtb_mesh_build_tree_on(rtb_mesh* mesh, int where);
/* Tree cache (not in the reference, which rebuilds on every start, WinMain.cpp:134-144): the built tree as a flat
 * little-endian file tied to the mesh by a hash of its points.  rtb_mesh_load_tree replaces rtb_mesh_build_tree and
 * fails with RTB_ERR_IO if the file belongs to another mesh or is damaged. */
int rtb_mesh_save_tree(const rtb_mesh* mesh, const char* file_name);
int rtb_mesh_load_tree(rtb_mesh* mesh, const char* file_name);
int64_t rtb_mesh_num_triangles(const rtb_mesh* mesh);
int64_t rtb_mesh_num_nodes(const rtb_mesh* mesh); /* Trixel::num_voxels */
/* copy the host tree out in the reference's kd_tree_node terms (Trixel.h:68-79): per node
 * left, right (-1 for leaves), tri (leaf triangle or -1), cut_flag (0..5), bounds x0,x1,y0,y1,z0,z1, s1, s2.
 * Any output pointer may be NULL. */
int rtb_mesh_get_tree(const rtb_mesh* mesh, int32_t* left, int32_t* right, int32_t* tri, int32_t* cut_flag,
                      float* bounds6, float* s1, float* s2);
/* seconds spent in the last rtb_mesh_build_tree: [0] sort, [1] partition, [2] total */
int rtb_mesh_build_seconds(const rtb_mesh* mesh, double out3[3]);
void rtb_mesh_destroy(rtb_mesh* mesh);

/* ---- camera ------------------------------------------------------------------------------ */

/* Camera::Camera(r_w, r_h, f_w, f_h, fclen, pos, look-at, up) + init_camera_device_memory
 * (Camera.h:86, Camera.cpp:5, Camera.cu:112).  The per-pixel ray table of init_cam_mem_cuda is not
 * materialised: rays are regenerated per pixel with the same arithmetic. */
int rtb_camera_create(int32_t r_w, int32_t r_h, float f_w, float f_h, float fclen, const float pos[3],
                      const float look_at[3], const float up[3], rtb_camera** out);
/* n, v, u, n_mod, v_mod, u_mod (Camera::orientation_properties, Camera.h:32-42), 18 floats */
int rtb_camera_get_basis(const rtb_camera* cam, float out18[18]);
/* Camera::add_object (Camera.cpp:118): binds the object's mesh to this camera and builds the
 * camera-relative device arrays (init_camera_trixel_device_memory Trixel.cu:244,
 * init_camera_voxel_device_memory Camera.cu:163).  Requires rtb_mesh_build_tree. */
int rtb_camera_add_object(rtb_camera* cam, rtb_object* obj);
/* Camera::color_pixels(tag) -> color_camera_device (Camera.cpp:229, Camera.cu:70):
 * PHONG: bring the shaded frame of the last render to the host buffer.  SET: in the reference this case fills the
 * frame with the background colour (240,130,0) (Camera.cpp:72, Camera.cu:12-18) and FALLS THROUGH into the Phong
 * pass (Camera.cu:77-82), i.e. the host receives "background + shaded hits of the last render"; same here: the last
 * rendered frame (every render writes every pixel), or the plain background / ids -1 before the first render. */
int rtb_camera_color_pixels(rtb_camera* cam, uint8_t color_tag_select);
/* Camera::h_mem.h_color.c (Camera.cpp:79): W*H little-endian 0x00RRGGBB words, row 0 = bottom.
 * Library-owned pinned memory holding the frame rtb_camera_color_pixels delivered last, valid until the next render on
 * this camera.  The library rotates between a few such buffers (frames are stored into them by the render kernel itself,
 * and predicted frames may be in flight behind the current one): read the pointer again after every color_pixels, as
 * integration/rtb_seam.cpp does for Camera::h_mem.h_color.c. */
const uint32_t* rtb_camera_host_color(const rtb_camera* cam);
/* hit triangle id per pixel (-1 = miss): the reference keeps these in Camera::pixel_memory::d_rmi
 * (Camera.h:57) and never reads them back; here they come with every frame. */
const int32_t* rtb_camera_host_ids(const rtb_camera* cam);
/* work counters of renders issued with RTB_RENDER_COUNTERS since the last reset:
 * [0] rays, [1] interior nodes entered (64-byte record fetches), [2] nodes popped in the reference's
 * sense (root + every child the parent scheduled), [3] triangle tests, [4] hits */
int rtb_camera_counters(rtb_camera* cam, uint64_t out5[5], int reset);
/* the same plus [5] sum over the traced rays of their deepest traversal stack (entries, counting the one kept in
 * registers), [6] the deepest stack any ray reached, [7] the most steps (interior records + triangle tests) any single ray took:
 * the length of the longest dependent chain of the launch */
int rtb_camera_counters_ex(rtb_camera* cam, uint64_t out8[8], int reset);
void rtb_camera_destroy(rtb_camera* cam);

/* ---- object + transform ------------------------------------------------------------------ */

/* Object::Object(Trixel*) (Object.cpp:4) */
int rtb_object_create(rtb_mesh* mesh, rtb_object** out);
/* Input::set_quat(x,y,z,w) + Object::transform(input, select) -> transform_camera_voxel_device_memory
 * (Input.cpp:15, Object.cpp:14, Camera.cu:254) incl. Quaternion::set_device_rotation (Quaternion.cu:21)
 * and update_voxel_transform_m_translate_cuda (Camera.cu:188).  ROTATE_TRI_*: xyzw = step quaternion;
 * TRANSLATE_*: xyz = direction, w = distance.  Requires rtb_camera_add_object. */
int rtb_object_transform(rtb_object* obj, const float xyzw[4], uint8_t transform_select);
/* the object's 3x4 matrix: rows x,y,z of Quaternion::rot_m, each (i,j,k,w=translation) */
int rtb_object_get_matrix(const rtb_object* obj, float m12[12]);
/* Render-only override (no reference counterpart): the next renders use m12 as they would use the recurrence's
 * matrix.  The quaternion and face bookkeeping of the recurrence (Quaternion::vec, Object::init_face / cur_face) are
 * NOT derived from it, so rtb_object_transform / rtb_object_transform_host / rtb_render_sweep steps return
 * RTB_ERR_STATE afterwards, until rtb_camera_add_object restarts the recurrence (Camera.cpp:131-134). */
int rtb_object_set_matrix(rtb_object* obj, const float m12[12]);
void rtb_object_destroy(rtb_object* obj);

/* ---- render ------------------------------------------------------------------------------ */

/* Object::render(Camera*) -> Trixel::intersect_trixels -> intersect_trixels_device
 * (Object.cpp:10, Trixel.h:474, Trixel.cu:210).  One fused launch: ray generation, traversal,
 * Moller-Trumbore, Phong and background; the kernel stores the finished frame straight into one of the camera's pinned host
 * buffers (with RTB_RENDER_COUNTERS, or the knob "host_direct" off: into the device frame, copied by color_pixels); follow
 * with rtb_camera_color_pixels(PHONG) to obtain it, exactly like the reference's frame loop (WinMain.cpp:212-213).  While
 * the transform steps between two renders repeat (a key held down), the next frames are rendered ahead on other streams
 * and a render call whose matrix equals the predicted one bit for bit finds its frame already on the way.  The call
 * returns as soon as the kernel is queued (the reference's cudaDeviceSynchronize at Trixel.cu:234 cannot be observed
 * before color_pixels delivers the frame); rtb_camera_color_pixels is the one
 * synchronisation of a frame, and launch errors surface there at the latest. */
int rtb_object_render(rtb_object* obj, rtb_camera* cam, uint32_t flags);
/* = rtb_object_render + rtb_camera_color_pixels(PHONG) with one synchronisation; the frame and the
 * id buffer are then readable through rtb_camera_host_color / rtb_camera_host_ids. */
int rtb_render_frame(rtb_object* obj, rtb_camera* cam, uint32_t flags);

/* ---- scene extension: what the reference leaves dormant around its hot path ----------
 * The reference draws ONE object with ONE light and ONE ray per pixel.  Its code carries the stubs of more: the light loop
 * and the shadow test of color_cam_cuda, commented out (Camera.cu:28-34, 54); Camera::render_properites::sample_rate
 * (Camera.h:46, Camera.cpp:71), never read; a second object, created and registered but never rendered
 * (WinMain.cpp:153,156,214-215; Camera::object_list, Camera.cpp:118-130).  There is no reference behaviour to be
 * identical to, so DESIGN.md section 11 defines it -- the smallest completion of those stubs that leaves the default
 * output untouched -- oracle/rtb_oracle.c (orc_render_scene) restates the definition on the CPU, and the kernel
 * (csrc/rtb_scene.cuh) is held to it bit for bit.  With one object, the default light, no shadows and sample_rate <= 1,
 * rtb_camera_render_scene produces exactly the frame of rtb_object_render.
 *   lights      : 1..8 point lights, default one at (2,2,2) (Camera.cu:32); radiance is summed over them in order
 *   shadows     : a light counts only if the segment from the hit point to it hits no other triangle of the object that
 *                 was hit (the commented call names one triangle list: objects do not shadow one another)
 *   sample_rate : n >= 2 casts n x n rays per pixel on a regular sub-pixel grid; the pixel is the per-channel integer
 *                 mean of the shaded samples, its hit id that of sample (n/2, n/2)
 *   objects     : every object added to the camera, in the order of rtb_camera_add_object; the closest hit wins, the
 *                 first added object wins ties; hit id = rtb_camera_object_id_base(object) + triangle id.  At most 8. */
int rtb_camera_set_lights(rtb_camera* cam, int32_t num_lights, const float* xyz3);
int rtb_camera_set_shadows(rtb_camera* cam, int32_t enable);
int rtb_camera_set_sample_rate(rtb_camera* cam, int32_t n);
int64_t rtb_camera_object_id_base(const rtb_camera* cam, const rtb_object* obj);
/* Camera::render() (Camera.cpp:160-163, a stub in the reference): all objects of the camera, each with its current
 * transform, into the camera's device frame; follow with rtb_camera_color_pixels(PHONG).  flags: RTB_RENDER_NO_CULL. */
int rtb_camera_render_scene(rtb_camera* cam, uint32_t flags);
/* the same into caller-owned device buffers (W*H elements each, either may be NULL) on a caller stream */
int rtb_camera_render_scene_device_async(rtb_camera* cam, uint32_t flags, uint32_t* d_bgra, int32_t* d_ids, void* stream);

/* Batched animation sweep (WinMain.cpp:174-239 with a key held down): for frame k = 0..num_frames-1
 * apply `steps_per_frame` transforms ops[k*steps_per_frame ...] (each 5 floats: select, x, y, z, w;
 * select 0 = no-op) and render.  Pinned caller buffers: host threads pre-fill a chunk of frames (256 MB) with the background
 * while the kernel traces the chunk before and stores every work unit that holds anything else straight into the buffers
 * over PCIe -- no copy engine.  Pageable buffers: chunks of about 16 MB of output per buffer, one persistent launch per
 * chunk; a finished chunk streams to the host while the next one renders (two chunks in flight).  The call
 * returns when every frame is in the caller's buffers.  bgra_out / ids_out are caller buffers of
 * num_frames*W*H elements (pageable or pinned; either may be NULL) -- the only interface here that
 * writes into caller memory.  The object's transform state advances as if the calls had been made
 * one by one; the camera's single-frame buffers (rtb_camera_host_color / _ids) are left untouched. */
int rtb_render_sweep(rtb_object* obj, rtb_camera* cam, int32_t num_frames, int32_t steps_per_frame,
                     const float* ops5, uint32_t flags, uint32_t* bgra_out, int32_t* ids_out);
/* Pinned (page-locked, device-mapped) host memory for rtb_render_sweep's output buffers, for callers without the CUDA
 * headers: buffers from here take the sweep's fast path (background pre-filled by host threads, the rest stored by the
 * kernel); pageable buffers take the copy-engine path.  Release with rtb_host_free. */
int rtb_host_alloc(size_t bytes, void** ptr);
int rtb_host_free(void* ptr);


/* Device-resident variant for callers that own device memory and a stream (multi-GPU gather,
 * kernel-only timing): renders frames m12[f] (12 floats each), f in [0,num_frames), restricted to
 * image tiles t with t % tile_stride == tile_first (tiles are 32x32 pixels, row-major over the
 * image; stride 1 = whole frame).  Output is written at its final row-major position in
 * d_bgra/d_ids (num_frames*W*H elements each, device pointers; either may be NULL).  `stream` is a
 * cudaStream_t (NULL = the library's own stream; pass cudaStreamLegacy for the default stream).  Asynchronous with
 * respect to the host: the call queues an upload of the frame records (none for a single frame: its record travels in
 * the kernel parameters) and the launch, and returns; it may wait a few microseconds for the PREVIOUS call's upload on
 * this object, never for a kernel.  Launches of one object are ordered on the device in call order whatever streams
 * they are issued on (they share the object's frame records and work counter); different objects render concurrently. */
int rtb_render_frames_device_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12,
                                   int32_t tile_first, int32_t tile_stride, uint32_t flags, uint32_t* d_bgra,
                                   int32_t* d_ids, void* stream);
/* Multi-GPU exchange format.  With RTB_RENDER_TILE_MAJOR, rtb_render_frames_device_async writes only
 * the tiles of this rank, compactly: frame f, owned tile slot j (image tile tile_first + j*tile_stride),
 * pixel (lx,ly) of the tile -> element (f*slots + j)*1024 + ly*32 + lx, slots = ceil(tiles/tile_stride).
 * rtb_tile_major_elements returns slots*1024, the per-frame element count of such a buffer (equal on
 * all ranks, so the buffers can be gathered with one equal-sized collective).  After the gather,
 * rtb_compose_tiles_device_async scatters `world` such buffers (d_parts[r] = rank r's, device
 * pointers to 32-bit elements: colours or ids alike) into the final row-major frames d_out. */
int64_t rtb_tile_major_elements(const rtb_camera* cam, int32_t tile_stride);
int rtb_compose_tiles_device_async(rtb_camera* cam, int32_t num_frames, int32_t world, const void* const* d_parts, void* d_out, void* stream);

/* Multi-GPU tile exchange fused into the render kernel (no collective, no reassembly pass).  Like
 * rtb_render_frames_device_async restricted to tiles t % tile_stride == tile_first, but d_frame_bgra / d_frame_ids are
 * the FINAL frames (num_frames*W*H elements, row-major; either may be NULL) and may live on ANOTHER GPU: every warp of the
 * persistent kernel copies each of its work units, the moment its last pixel is shaded, to its place in those frames
 * with 16-byte stores -- over NVLink when the pointer is peer-mapped (rtb_peer_open) -- so the transfer overlaps the
 * rendering unit by unit.  The frames are complete once the kernels of all ranks have finished (order a barrier or any
 * collective after them on `stream`).  This replaces the reference's single-GPU frame buffer (Camera.cpp:78) for the
 * multi-GPU case; RTB_RENDER_COUNTERS is not available here. */
int rtb_render_frames_push_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12,
                                 int32_t tile_first, int32_t tile_stride, uint32_t flags, uint32_t* d_frame_bgra,
                                 int32_t* d_frame_ids, void* stream);
/* The same with STRIPED frame ownership: frame f of the launch is frame f / owners of owner f % owners, whose final-frame
 * buffers are d_frame_bgra[f % owners] / d_frame_ids[f % owners] (each holding ceil(num_frames / owners) * W*H elements;
 * either array may be NULL).  With one owner per rank every GPU takes in 1/N of the pushed pixels instead of one GPU
 * taking all of them (that GPU's NVLink ingest bounded the single-owner exchange at N = 8), and every rank ends up
 * holding -- and delivering -- 1/N of the finished frames.  owners <= 8. */
int rtb_render_frames_push_striped_async(rtb_object* obj, rtb_camera* cam, int32_t num_frames, const float* m12,
                                         int32_t tile_first, int32_t tile_stride, uint32_t flags, int32_t owners,
                                         uint32_t* const* d_frame_bgra, int32_t* const* d_frame_ids, void* stream);
/* set_cam_cuda (Camera.cu:12-18, the SET_COLOR_TAG fill) for `num_frames` device frames at once: background colour into
 * d_frame_bgra, -1 into d_frame_ids (either may be NULL).  With RTB_RENDER_PUSH_PREFILLED the owner of the final frames
 * runs this before the ranks push, and work units that contain nothing but background never cross NVLink. */
int rtb_fill_frames_device_async(rtb_camera* cam, int32_t num_frames, uint32_t* d_frame_bgra, int32_t* d_frame_ids, void* stream);
/* Frames that other processes' GPUs can write: rtb_peer_alloc = one device allocation on the current device (plays the
 * role of the cudaMalloc of Camera.cpp:78), rtb_peer_export = its 64-byte inter-process handle (send it to the other
 * ranks by any means, e.g. torch.distributed.all_gather_object), rtb_peer_open = map a peer's allocation into this
 * process on the current device (NVLink peer access is enabled on first use), rtb_peer_close / rtb_peer_free undo them. */
int rtb_peer_alloc(size_t bytes, void** d_ptr);
int rtb_peer_free(void* d_ptr);
int rtb_peer_export(void* d_ptr, uint8_t handle64[64]);
int rtb_peer_open(const uint8_t handle64[64], void** d_ptr);
int rtb_peer_close(void* d_ptr);
/* synchronous copy of a (peer) device buffer to host memory after a device-wide synchronisation: the read-out of the
 * assembled frames (the role of the cudaMemcpy of Camera.cu:84) */
int rtb_peer_read(void* host_dst, const void* d_ptr, size_t bytes);

/* host-side transform recurrence only (no GPU): advance `obj` by one op and return its matrix;
 * lets callers precompute the m12 array for rtb_render_frames_device_async. */
int rtb_object_transform_host(rtb_object* obj, const float xyzw[4], uint8_t transform_select, float m12_out[12]);

/* The same recurrence with no object and no GPU: start from the state Camera::add_object gives a new object for a camera
 * at cam_pos (Camera.cpp:131-134), apply `count` ops (5 floats each: select, x, y, z, w; select 0 = no-op) one by one and
 * write the matrix after each into m12_out (12 floats per op).  What a caller uses to precompute the matrices of a sweep,
 * and what the tests pin against the reference's Object::transform over the whole 600-step orbit. */
int rtb_transform_sequence_host(const float cam_pos[3], int32_t count, const float* ops5, float* m12_out);

/* Device self-test of the exactness arguments the kernel relies on: `count` pseudo-random
 * operand sets (random bit patterns, near-ties, tiny values, NaN).  out4[0] = early-exit Newton rsqrt != literal 21
 * steps, out4[1] = __frcp_rn(f) != (float)(1.0/(double)f), out4[2] = fp32 decision shortcut != exact double
 * evaluation on an input it declared decidable, out4[3] = number of decidable samples.  All mismatch counts must be 0. */
int rtb_selftest_exact(uint64_t seed, int64_t count, uint64_t out4[4]);

/* L2 read bandwidth of the current device for the roofline: `iters` passes of 16-byte L1-bypassing loads over a buffer of
 * `bytes` (choose it well below the L2 size, e.g. 32 MB) by 8 blocks of 256 threads per SM; result in GB/s. */
int rtb_measure_l2_read_bandwidth(size_t bytes, int iters, double* gb_per_s);

/* What the host can take when IT writes a frame buffer: `bytes` of pinned memory filled by `threads` host threads
 * (<= 0: all) with streaming stores, best of three passes.  The sweep's background pre-fill is bound by this, the frames
 * that cross PCIe by the copy bandwidth (bench.py reports both). */
int rtb_measure_host_fill_bandwidth(size_t bytes, int threads, double* gb_per_s);
/* number of kernels this library has launched in this process (render, pack and fill kernels) */
uint64_t rtb_launch_count(void);

/* device properties the roofline uses: [0] SM count, [1] L2 bytes, [2] max persisting L2 bytes,
 * [3] SM clock kHz, [4] memory clock kHz, [5] memory bus width bits, [6] compute capability*10 */
int rtb_device_props(int64_t out7[7]);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
