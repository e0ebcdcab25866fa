/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-pixel ray-cast path.
 *
 * This is the parity ORACLE for the CUDA path: a plain-C restatement of the arithmetic of
 * ams3878/cpp_cuda_raytracer_dev (TEST_Dungeonrun/), fp32 with FMA contraction OFF
 * (build: gcc -O2 -ffp-contract=off -fopenmp, SSE2 scalar floats; see oracle/Makefile).
 * Every function cites the reference file:line it restates.  It is pinned bit-for-bit against
 * the reference's own kernels compiled for the host (oracle/_ref/libref_emu.so, built by
 * oracle/build_ref.py from /root/reference) and against the fixtures in tests/golden/ that
 * were generated from that library (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (cpp_cuda_raytracer_dev_b200/csrc) never links or calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t s64;
typedef int32_t s32;
typedef uint32_t u32;
typedef uint8_t u8;

/* vector.cuh:10-11 -- both epsilons are DOUBLE literals, so every comparison against them is
 * carried out in double precision. */
#define MT_EPS 1e-16
#define DEV_EPS 1e-16

/* ------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* vector.cuh:122-124 device_dot: (ax*bx) + (ay*by) + (az*bz), left to right */
static inline float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ((ax * bx) + (ay * by)) + (az * bz);
}
/* vector.cuh:73-77 device_cross */
static inline void cross3(float* cx, float* cy, float* cz, float ax, float ay, float az, float bx, float by, float bz) {
    *cx = ay * bz - az * by;
    *cy = az * bx - ax * bz;
    *cz = ax * by - ay * bx;
}
/* vector.cuh:79-95 device_inverse_sqrt: Quake start value, 1 + 20 Newton steps */
static inline float device_inverse_sqrt(float x, float y, float z) {
    union { float x; s32 i; } u;
    u.x = ((x * x) + (y * y)) + (z * z);
    float half = 0.5f * u.x;
    u.x = half;
    u.i = 0x5f375a86 - (u.i >> 1);
    for (int k = 0; k < 21; k++) u.x = u.x * (1.5f - half * u.x * u.x);
    return u.x;
}
/* vector.cuh:117-120 */
static inline void device_normalize(float* x, float* y, float* z) {
    float s = device_inverse_sqrt(*x, *y, *z);
    *x *= s; *y *= s; *z *= s;
}
/* vector.cpp:13-26 vector_norm (host): same start value, 8 Newton steps.  The original unions the
 * float with an s64 whose upper half is never written; the intended (and MSVC-observed) behaviour
 * is the 32-bit one. */
static inline float vector_norm(float s) {
    float half = 0.5f * s;
    union { float x; s32 i; } u;
    u.x = half;
    u.i = 0x5f375a86 - (u.i >> 1);
    for (int k = 0; k < 8; k++) u.x = u.x * (1.5f - half * u.x * u.x);
    return u.x;
}
typedef struct { float x, y, z, w; } vec4;
/* Vector.h:116-124 normalize_Vector(VEC4*) : length goes to w */
static inline void normalize_vec4(vec4* v) {
    float s = v->x * v->x + v->y * v->y + v->z * v->z;
    s = vector_norm(s);
    v->x *= s; v->y *= s; v->z *= s;
    v->w = 1 / s;
}
/* vector.cpp:31-36 VEC4::cross (in place: this = this x b) */
static inline void cross_vec4(vec4* a, vec4 b) {
    float t0 = a->y * b.z - a->z * b.y;
    float t1 = a->z * b.x - a->x * b.z;
    float t2 = a->x * b.y - a->y * b.x;
    a->x = t0; a->y = t1; a->z = t2;
}

/* ------------------------------------------------------------------------------------------ */
/* camera basis                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* Camera.cpp:5-67.  out18 = n, v, u, n_mod, v_mod, u_mod. */
void orc_camera_basis(int W, int H, float f_w, float f_h, float fclen, const float* pos, const float* la, const float* up,
                      float* out18) {
    float pix_w = f_w / (float)W;
    float pix_h = f_h / (float)H;
    vec4 tn = {la[0] - pos[0], la[1] - pos[1], la[2] - pos[2], 1.0f};
    normalize_vec4(&tn);
    float n[3] = {tn.x, tn.y, tn.z};
    vec4 tu = {up[0], up[1], up[2], 1.0f};
    normalize_vec4(&tu);
    cross_vec4(&tu, tn);  /* up x n */
    cross_vec4(&tn, tu);  /* n x (up x n) */
    vec4 tv = tn;
    normalize_vec4(&tv);
    float v[3] = {tv.x, tv.y, tv.z};
    float v_mod[3] = {v[0] * pix_h, v[1] * pix_h, v[2] * pix_h};
    vec4 t2 = {la[0] - pos[0], la[1] - pos[1], la[2] - pos[2], 1.0f};
    normalize_vec4(&t2);
    cross_vec4(&tv, t2);  /* v x n */
    float u[3] = {tv.x, tv.y, tv.z};
    float u_mod[3] = {u[0] * pix_w, u[1] * pix_w, u[2] * pix_w};
    float adjust_y = (float)((u32)H >> 1), adjust_x = (float)((u32)W >> 1);
    if (!(H & 1)) adjust_y -= .5;
    if (!(W & 1)) adjust_x -= .5;
    float n_mod[3];
    for (int c = 0; c < 3; c++) n_mod[c] = (n[c] * fclen) - (v_mod[c] * adjust_y) - (u_mod[c] * adjust_x);
    memcpy(out18 + 0, n, 12); memcpy(out18 + 3, v, 12); memcpy(out18 + 6, u, 12);
    memcpy(out18 + 9, n_mod, 12); memcpy(out18 + 12, v_mod, 12); memcpy(out18 + 15, u_mod, 12);
}

/* Camera.cu:89-111 init_cam_mem_cuda: camera-space primary ray of pixel i (row 0 = bottom). */
static inline void primary_ray(const float* n_mod, const float* u_mod, const float* v_mod, int W, s64 i, float* r) {
    uint64_t i_y = (uint64_t)i / (uint64_t)W, i_x = (uint64_t)i % (uint64_t)W;
    r[0] = n_mod[0] + u_mod[0] * i_x + v_mod[0] * i_y;
    r[1] = n_mod[1] + u_mod[1] * i_x + v_mod[1] * i_y;
    r[2] = n_mod[2] + u_mod[2] * i_x + v_mod[2] * i_y;
    device_normalize(&r[0], &r[1], &r[2]);
}
void orc_rays(const float* basis18, int W, int H, float* out3p) {
    for (s64 i = 0; i < (s64)W * H; i++) primary_ray(basis18 + 9, basis18 + 15, basis18 + 12, W, i, out3p + 3 * i);
}

/* ------------------------------------------------------------------------------------------ */
/* object transform recurrence                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* Quaternion.cpp:9-16 + Camera.cpp:131-134: identity quaternion/matrix, faces = -camera position */
typedef struct {
    vec4 q;          /* Quaternion::vec (i,j,k,w) */
    vec4 rx, ry, rz; /* Quaternion::rot_m rows; .w = translation */
    vec4 init_face, cur_face;
} orc_xform;

void orc_xform_init(orc_xform* t, const float* cam_pos) {
    t->q = (vec4){0, 0, 0, 1};
    t->rx = (vec4){1, 0, 0, 0}; t->ry = (vec4){0, 1, 0, 0}; t->rz = (vec4){0, 0, 1, 0};
    t->init_face = (vec4){-cam_pos[0], -cam_pos[1], -cam_pos[2], 1.0f};
    t->cur_face = t->init_face;
}
/* vector.cpp:38-65 VEC4::rotate */
static void vec4_rotate(vec4* v, orc_xform* t, const vec4* nq, int reverse) {
    if (nq) {
        vec4* c = &t->q;
        float t_i = c->x, t_j = c->y, t_k = c->z, t_w = c->w;
        c->x = t_j * nq->z - t_k * nq->y + t_i * nq->w + t_w * nq->x;
        c->y = t_k * nq->x - t_i * nq->z + t_j * nq->w + t_w * nq->y;
        c->z = t_i * nq->y - t_j * nq->x + t_k * nq->w + t_w * nq->z;
        c->w = t_w * nq->w - t_i * nq->x - t_j * nq->y - t_k * nq->z;
        float i = c->x, j = c->y, k = c->z, w = c->w;
        t->rx.x = (1 - 2 * j * j - 2 * k * k);
        t->rx.y = (2 * i * j - 2 * k * w);
        t->rx.z = (2 * i * k + 2 * j * w);
        t->ry.x = (2 * i * j + 2 * k * w);
        t->ry.y = (1 - 2 * i * i - 2 * k * k);
        t->ry.z = (2 * j * k - 2 * i * w);
        t->rz.x = (2 * i * k - 2 * j * w);
        t->rz.y = (2 * j * k + 2 * i * w);
        t->rz.z = (1 - 2 * i * i - 2 * j * j);
    }
    float tx = v->x * reverse, ty = v->y * reverse, tz = v->z * reverse;
    v->x = (tx * t->rx.x + ty * t->rx.y + tz * t->rx.z);
    v->y = (tx * t->ry.x + ty * t->ry.y + tz * t->ry.z);
    v->z = (tx * t->rz.x + ty * t->rz.y + tz * t->rz.z);
}
/* Vector.h:89-100 VEC4 -= / += : operands are first scaled by their w, result w = 1 */
static inline void vec4_sub(vec4* a, vec4 b) {
    a->x = (a->x * a->w) - (b.x * b.w); a->y = (a->y * a->w) - (b.y * b.w); a->z = (a->z * a->w) - (b.z * b.w); a->w = 1.0f;
}
static inline void vec4_add(vec4* a, vec4 b) {
    a->x = (a->x * a->w) + (b.x * b.w); a->y = (a->y * a->w) + (b.y * b.w); a->z = (a->z * a->w) + (b.z * b.w); a->w = 1.0f;
}
/* Camera.cu:254-335 transform_camera_voxel_device_memory (+ Input::set_quat, Input.cpp:6-19).
 * select: 10/11 = ROTATE_TRI_PY/NY with step quaternion (x,y,z,w); 30/31/32 = TRANSLATE with
 * direction (x,y,z) and distance w (platform_common.h:16-21).  The device copy of the matrix
 * (Quaternion.cu:4-10, Camera.cu:188-192) receives the same values (contraction off). */
void orc_xform_apply(orc_xform* t, int select, float x, float y, float z, float w) {
    vec4 tv = {x, y, z, w};
    switch (select) {
    case 30: case 31: case 32: {
        vec4_sub(&t->init_face, tv);
        vec4_rotate(&tv, t, NULL, -1);
        t->rx.w += tv.w * tv.x;
        t->ry.w += tv.w * tv.y;
        t->rz.w += tv.w * tv.z;
        normalize_vec4(&t->init_face);
        t->cur_face = t->init_face;
        vec4_rotate(&t->cur_face, t, NULL, -1);
        t->cur_face.x = -t->cur_face.x; t->cur_face.y = -t->cur_face.y; t->cur_face.z = -t->cur_face.z;
        break;
    }
    case 10: case 11: {
        vec4 tmp = t->init_face;
        vec4_rotate(&tmp, t, &tv, -1);
        vec4_add(&tmp, t->cur_face);
        t->rx.w -= tmp.x * tmp.w;
        t->ry.w -= tmp.y * tmp.w;
        t->rz.w -= tmp.z * tmp.w;
        vec4_sub(&tmp, t->cur_face);
        t->cur_face = tmp;
        normalize_vec4(&t->cur_face);
        t->cur_face.x = -t->cur_face.x; t->cur_face.y = -t->cur_face.y; t->cur_face.z = -t->cur_face.z;
        break;
    }
    default: break;
    }
}
void orc_xform_matrix(const orc_xform* t, float* m12) {
    const vec4* rows[3] = {&t->rx, &t->ry, &t->rz};
    for (int r = 0; r < 3; r++) { m12[4 * r] = rows[r]->x; m12[4 * r + 1] = rows[r]->y; m12[4 * r + 2] = rows[r]->z; m12[4 * r + 3] = rows[r]->w; }
}
size_t orc_xform_sizeof(void) { return sizeof(orc_xform); }


/* ------------------------------------------------------------------------------------------ */
/* PLY loader                                                                                  */
/* ------------------------------------------------------------------------------------------ */

static inline float min3f(float a, float b, float c) { float m = b < c ? b : c; return a < m ? a : m; }
static inline float max3f(float a, float b, float c) { float m = b > c ? b : c; return a > m ? a : m; }

/* append one triangle (9 floats) in the loader's output order */
static void put_tri(float* pts, long* nt, const float* a, const float* b, const float* c) {
    float* p = pts + 9 * (size_t)(*nt);
    memcpy(p, a, 12); memcpy(p + 3, b, 12); memcpy(p + 6, c, 12);
    (*nt)++;
}

/* read_ply.cpp:13-152.  The header scan only honours `element vertex|face <n>` (:19-44).
 * mode 0/1/2 = 3/5/6 whitespace-separated numbers per vertex (:52-65), parsed like
 * `istream >> float` (one correctly-rounded decimal->float conversion = strtof).
 * Faces (:66-150): "3 a b c" is stored as (c,a,b) (:138-148); "4 a b c d" becomes the two
 * triangles (a,b,c),(a,c,d) (:92-124); triangle id = running output index (:90,112,136).
 * mode -1 is NOT in the reference: a conforming reader for `format binary_little_endian`
 * files (3_walls.ply: N float properties per vertex of which the first three are x,y,z; faces =
 * uchar count + uint32 indices) that applies the same two face rules.
 * Returns 0 and a malloc'd 9-floats-per-triangle soup. */
int orc_read_ply(const char* path, int mode, float** points9, long* ntri_out) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
    fclose(f);
    buf[sz] = 0;
    long nvert = 0, nface = 0;
    int nprops = 0, in_vertex = 0, binary = 0;
    char* p = buf;
    char* body = NULL;
    while (p < buf + sz) {
        char* e = memchr(p, '\n', (size_t)(buf + sz - p));
        if (!e) e = buf + sz;
        size_t L = (size_t)(e - p);
        while (L && (p[L - 1] == '\r' || p[L - 1] == ' ')) L--;
        if (L == 10 && !memcmp(p, "end_header", 10)) { body = e + 1; break; }
        if (L >= 27 && !memcmp(p, "format binary_little_endian", 27)) binary = 1;
        if (L > 15 && !memcmp(p, "element vertex ", 15)) { nvert = atol(p + 15); in_vertex = 1; }
        else if (L > 13 && !memcmp(p, "element face ", 13)) { nface = atol(p + 13); in_vertex = 0; }
        else if (L > 8 && !memcmp(p, "element ", 8)) in_vertex = 0;
        else if (in_vertex && L > 9 && !memcmp(p, "property ", 9)) nprops++;
        p = e + 1;
    }
    if (!body || nvert <= 0 || nface <= 0) { free(buf); return -2; }
    float* vx = (float*)malloc(sizeof(float) * 3 * (size_t)nvert);
    float* pts = (float*)malloc(sizeof(float) * 9 * 2 * (size_t)nface);
    long nt = 0;
    int rc = 0;
    if (mode == -1 && binary) {
        const unsigned char* b = (const unsigned char*)body;
        const unsigned char* end = (const unsigned char*)buf + sz;
        for (long i = 0; i < nvert && rc == 0; i++) {
            if (b + 4 * nprops > end) { rc = -3; break; }
            memcpy(vx + 3 * i, b, 12);
            b += 4 * nprops;
        }
        for (long fi = 0; fi < nface && rc == 0; fi++) {
            if (b + 1 > end) { rc = -3; break; }
            int c = *b++;
            u32 idx[4];
            if ((c != 3 && c != 4) || b + 4 * c > end) { rc = -4; break; }
            memcpy(idx, b, 4 * (size_t)c);
            b += 4 * c;
            for (int k = 0; k < c; k++) if ((long)idx[k] >= nvert) rc = -5;
            if (rc) break;
            if (c == 4) {
                put_tri(pts, &nt, vx + 3 * idx[0], vx + 3 * idx[1], vx + 3 * idx[2]);
                put_tri(pts, &nt, vx + 3 * idx[0], vx + 3 * idx[2], vx + 3 * idx[3]);
            } else {
                put_tri(pts, &nt, vx + 3 * idx[2], vx + 3 * idx[0], vx + 3 * idx[1]);
            }
        }
    } else {
        int extra = mode == 1 ? 2 : mode == 2 ? 3 : mode == -1 ? nprops - 3 : 0;
        char* q = body;
        char* e2;
        for (long i = 0; i < nvert && rc == 0; i++) {
            for (int k = 0; k < 3 + extra; k++) {
                float v = strtof(q, &e2);
                if (e2 == q) { rc = -3; break; }
                q = e2;
                if (k < 3) vx[3 * i + k] = v;
            }
        }
        /* the reference loops while (points written) < num_tri*3, num_tri growing per quad (:68-72):
         * i.e. exactly the `nface` face records of the file are consumed */
        for (long fi = 0; fi < nface && rc == 0; fi++) {
            long c = strtol(q, &e2, 10);
            if (e2 == q) { rc = -3; break; }
            q = e2;
            long idx[4];
            if (c != 3 && c != 4) { rc = -4; break; }
            for (int k = 0; k < c; k++) {
                idx[k] = strtol(q, &e2, 10);
                if (e2 == q || idx[k] < 0 || idx[k] >= nvert) { rc = -5; break; }
                q = e2;
            }
            if (rc) break;
            if (c == 4) {
                put_tri(pts, &nt, vx + 3 * idx[0], vx + 3 * idx[1], vx + 3 * idx[2]);
                put_tri(pts, &nt, vx + 3 * idx[0], vx + 3 * idx[2], vx + 3 * idx[3]);
            } else {
                put_tri(pts, &nt, vx + 3 * idx[2], vx + 3 * idx[0], vx + 3 * idx[1]);
            }
        }
    }
    free(vx);
    free(buf);
    if (rc) { free(pts); return rc; }
    *points9 = pts;
    *ntri_out = nt;
    return 0;
}
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* tree build                                                                                  */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    s64 left, right, tri, parent;
    int cut_flag, is_leaf;
    float x0, x1, y0, y1, z0, z1, s1, s2;
} orc_node;

/* list numbering of Trixel.h:217-236: 0=x1 1=y1 2=z1 3=x0 4=y0 5=z0 */
static const float* g_sort_key;
static int cmp_key_desc_index(const void* a, const void* b) {
    s32 ia = *(const s32*)a, ib = *(const s32*)b;
    float ka = g_sort_key[ia], kb = g_sort_key[ib];
    if (ka < kb) return -1;
    if (kb < ka) return 1;
    return ia > ib ? -1 : ia < ib ? 1 : 0; /* ties: higher original index first */
}

/* sort.h:11-60 merge_sort + Trixel.h:386-473 set_sorted_voxels + Trixel.h:135-385 create_kd.
 *  - per-triangle AABB = min/max of its three vertices (read_ply.cpp:127-134);
 *  - six lists sorted by x1,y1,z1,x0,y0,z0; the top-down merge takes the RIGHT run on ties
 *    (sort.h:31-54: `left < right ? left : right`), which is the order (key asc, index desc);
 *  - BFS over nodes (read_index/write_index): node = range [l,r] holding the same triangle set
 *    in all six lists; bounds from list ends (:155-160, :345-350); split list = first strict
 *    maximum of key[r]-key[l] in the order x1,x0,y1,y0,z1,z0 (:172-193); m = l+(r-l)/2, left =
 *    [l,m], right = [m+1,r]; other five lists stably partitioned by "position in split list <= m"
 *    (:214-327); leaf when r==l, triangle = x1-list[l], leaf inherits parent's cut_flag (:194-205);
 *    s1 = left child's max, s2 = right child's min on the split axis (:353-376).
 * nodes must hold 2n-1 entries. */
int orc_build_tree(const float* points9, long n, orc_node* nodes) {
    if (n <= 0) return -1;
    float* key[6];
    s32* ord[6];
    s32* pos[6];
    for (int k = 0; k < 6; k++) {
        key[k] = (float*)malloc(sizeof(float) * (size_t)n);
        ord[k] = (s32*)malloc(sizeof(s32) * (size_t)n);
        pos[k] = (s32*)malloc(sizeof(s32) * (size_t)n);
    }
    for (long i = 0; i < n; i++) {
        const float* p = points9 + 9 * (size_t)i;
        key[0][i] = max3f(p[0], p[3], p[6]); key[3][i] = min3f(p[0], p[3], p[6]);
        key[1][i] = max3f(p[1], p[4], p[7]); key[4][i] = min3f(p[1], p[4], p[7]);
        key[2][i] = max3f(p[2], p[5], p[8]); key[5][i] = min3f(p[2], p[5], p[8]);
    }
    for (int k = 0; k < 6; k++) {
        for (long i = 0; i < n; i++) ord[k][i] = (s32)i;
        g_sort_key = key[k];
        qsort(ord[k], (size_t)n, sizeof(s32), cmp_key_desc_index);
        for (long i = 0; i < n; i++) pos[k][ord[k][i]] = (s32)i;
    }
    s32* tmp = (s32*)malloc(sizeof(s32) * (size_t)n);
    s64* nl = (s64*)malloc(sizeof(s64) * (size_t)(2 * n - 1));
    s64* nr = (s64*)malloc(sizeof(s64) * (size_t)(2 * n - 1));
    s64 read_index = 0, write_index = 1;
    nl[0] = 0; nr[0] = n - 1;
    nodes[0].parent = 0;
    nodes[0].cut_flag = 5;
    nodes[0].x0 = key[3][ord[3][0]]; nodes[0].x1 = key[0][ord[0][n - 1]];
    nodes[0].y0 = key[4][ord[4][0]]; nodes[0].y1 = key[1][ord[1][n - 1]];
    nodes[0].z0 = key[5][ord[5][0]]; nodes[0].z1 = key[2][ord[2][n - 1]];
    static const int scan[6] = {0, 3, 1, 4, 2, 5};
    while (read_index < write_index) {
        orc_node* cur = &nodes[read_index];
        s64 l = nl[read_index], r = nr[read_index], m = l + (r - l) / 2;
        cur->s1 = 0; cur->s2 = 0; cur->tri = -1; cur->is_leaf = 0;
        float max_split = key[0][ord[0][r]] - key[0][ord[0][l]];
        int max_cut = 0;
        for (int si = 1; si < 6; si++) {
            int k = scan[si];
            float d = key[k][ord[k][r]] - key[k][ord[k][l]];
            if (d > max_split) { max_split = d; max_cut = k; }
        }
        if (r == l) {
            cur->cut_flag = nodes[cur->parent].cut_flag;
            cur->left = -1; cur->right = -1; cur->is_leaf = 1;
            cur->tri = ord[0][l];
            read_index++;
            continue;
        }
        cur->cut_flag = max_cut;
        const s32* cpos = pos[max_cut];
        for (int k = 0; k < 6; k++) {
            if (k == max_cut) continue;
            s64 li = l, ri = m + 1;
            for (s64 i = l; i <= r; i++) {
                s32 t = ord[k][i];
                if (cpos[t] <= m) tmp[li++] = t; else tmp[ri++] = t;
            }
            for (s64 i = l; i <= r; i++) { ord[k][i] = tmp[i]; pos[k][tmp[i]] = (s32)i; }
        }
        for (int branch = 0; branch < 2; branch++) {
            s64 a = branch == 0 ? l : m + 1, b = branch == 0 ? m : r;
            orc_node* ch = &nodes[write_index];
            ch->parent = read_index;
            nl[write_index] = a; nr[write_index] = b;
            ch->x1 = key[0][ord[0][b]]; ch->x0 = key[3][ord[3][a]];
            ch->y1 = key[1][ord[1][b]]; ch->y0 = key[4][ord[4][a]];
            ch->z1 = key[2][ord[2][b]]; ch->z0 = key[5][ord[5][a]];
            if (branch == 0) cur->left = write_index; else cur->right = write_index;
            write_index++;
        }
        const orc_node* L = &nodes[write_index - 2];
        const orc_node* R = &nodes[write_index - 1];
        switch (max_cut % 3) {
        case 0: cur->s2 = R->x0; cur->s1 = L->x1; break;
        case 1: cur->s2 = R->y0; cur->s1 = L->y1; break;
        case 2: cur->s2 = R->z0; cur->s1 = L->z1; break;
        }
        read_index++;
    }
    for (int k = 0; k < 6; k++) { free(key[k]); free(ord[k]); free(pos[k]); }
    free(tmp); free(nl); free(nr);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* scene precompute (the reference's init kernels)                                             */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    long n, nn;
    int W, H;
    float cam_pos[3], n_mod[3], u_mod[3], v_mod[3];
    float draw_distance;
    u8 bg[4]; /* r,g,b,a  Camera.cpp:72 */
    /* per triangle: Trixel.cu:11-27 (e1,e2,n) + Trixel.cu:29-36 (d_t) + colour */
    float *e1, *e2, *nrm, *dt, *rad;
    /* per node: Camera.cu:137-162 */
    float *Bo, *S1, *S2; /* Bo = t0x,t0y,t0z,t1x,t1y,t1z */
    s64 *left, *right, *tri;
    u8 *is_leaf, *flags; /* flags = x,y,z one-hot */
} orc_scene;

orc_scene* orc_scene_create(const float* points9, long n, const float* rad3, int rad_per_tri, const orc_node* nodes,
                            int W, int H, const float* cam_pos, const float* basis18) {
    orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
    s->n = n; s->nn = 2 * n - 1; s->W = W; s->H = H;
    memcpy(s->cam_pos, cam_pos, 12);
    memcpy(s->n_mod, basis18 + 9, 12); memcpy(s->v_mod, basis18 + 12, 12); memcpy(s->u_mod, basis18 + 15, 12);
    s->draw_distance = 400;
    s->bg[0] = 240; s->bg[1] = 130; s->bg[2] = 0; s->bg[3] = 0;
    s->e1 = (float*)malloc(12 * (size_t)n); s->e2 = (float*)malloc(12 * (size_t)n);
    s->nrm = (float*)malloc(12 * (size_t)n); s->dt = (float*)malloc(12 * (size_t)n); s->rad = (float*)malloc(12 * (size_t)n);
    for (long i = 0; i < n; i++) {
        const float* p = points9 + 9 * (size_t)i;
        float* e1 = s->e1 + 3 * i; float* e2 = s->e2 + 3 * i; float* nn = s->nrm + 3 * i; float* dt = s->dt + 3 * i;
        e1[0] = p[3] - p[0]; e1[1] = p[4] - p[1]; e1[2] = p[5] - p[2];
        e2[0] = p[6] - p[0]; e2[1] = p[7] - p[1]; e2[2] = p[8] - p[2];
        cross3(&nn[0], &nn[1], &nn[2], e1[0], e1[1], e1[2], e2[0], e2[1], e2[2]);
        device_normalize(&nn[0], &nn[1], &nn[2]);
        dt[0] = cam_pos[0] - p[0]; dt[1] = cam_pos[1] - p[1]; dt[2] = cam_pos[2] - p[2];
        const float* c = rad_per_tri ? rad3 + 3 * i : rad3;
        s->rad[3 * i] = c[0]; s->rad[3 * i + 1] = c[1]; s->rad[3 * i + 2] = c[2];
    }
    long nn = s->nn;
    s->Bo = (float*)malloc(24 * (size_t)nn); s->S1 = (float*)malloc(4 * (size_t)nn); s->S2 = (float*)malloc(4 * (size_t)nn);
    s->left = (s64*)malloc(8 * (size_t)nn); s->right = (s64*)malloc(8 * (size_t)nn); s->tri = (s64*)malloc(8 * (size_t)nn);
    s->is_leaf = (u8*)malloc((size_t)nn); s->flags = (u8*)malloc(3 * (size_t)nn);
    const float ocx = 0.0f, ocy = 0.0f, ocz = 0.0f; /* obj_center, Camera.cpp:167-170 */
    for (long i = 0; i < nn; i++) {
        const orc_node* k = &nodes[i];
        float* B = s->Bo + 6 * i;
        B[0] = k->x0 - cam_pos[0] + ocx; B[3] = k->x1 - cam_pos[0] + ocx;
        B[1] = k->y0 - cam_pos[1] + ocy; B[4] = k->y1 - cam_pos[1] + ocy;
        B[2] = k->z0 - cam_pos[2] + ocz; B[5] = k->z1 - cam_pos[2] + ocz;
        s->is_leaf[i] = (u8)k->is_leaf;
        s->left[i] = k->left; s->right[i] = k->right;
        s->tri[i] = k->is_leaf == 0 ? -1 : k->tri;
        int cd = k->cut_flag;
        u8 fx = (cd == 0 || cd == 3) ? 1 : 0, fy = (cd == 1 || cd == 4) ? 1 : 0, fz = (cd == 2 || cd == 5) ? 1 : 0;
        s->flags[3 * i] = fx; s->flags[3 * i + 1] = fy; s->flags[3 * i + 2] = fz;
        s->S1[i] = k->s1 - (((cam_pos[0] + ocx) * (float)fx) + ((cam_pos[1] + ocy) * (float)fy) + ((cam_pos[2] + ocz) * (float)fz));
        s->S2[i] = k->s2 - (((cam_pos[0] + ocx) * (float)fx) + ((cam_pos[1] + ocy) * (float)fy) + ((cam_pos[2] + ocx) * (float)fz));
    }
    return s;
}
void orc_scene_destroy(orc_scene* s) {
    if (!s) return;
    free(s->e1); free(s->e2); free(s->nrm); free(s->dt); free(s->rad); free(s->Bo); free(s->S1); free(s->S2);
    free(s->left); free(s->right); free(s->tri); free(s->is_leaf); free(s->flags); free(s);
}

/* ------------------------------------------------------------------------------------------ */
/* the hot path: traversal + Moller-Trumbore + Phong                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct { float dist, rad[3], pnt[3], norm[3]; s64 id; } orc_hit;

/* Moller-Trumbore of Trixel.cu:98-145 for triangle t; updates *best / *h on acceptance. */
static inline void mt_test(const orc_scene* s, s64 t, const float* m12, float rx, float ry, float rz, float odx, float ody,
                           float odz, float* best, orc_hit* h) {
    const float* e1 = s->e1 + 3 * t; const float* e2 = s->e2 + 3 * t; const float* dt = s->dt + 3 * t;
    float px, py, pz, qx, qy, qz;
    cross3(&px, &py, &pz, rx, ry, rz, e2[0], e2[1], e2[2]);
    float f = dot3(px, py, pz, e1[0], e1[1], e1[2]);
    if (!(f < MT_EPS && f > -MT_EPS)) {
        float pe1 = 1.0 / f;
        float u = pe1 * dot3(px, py, pz, dt[0] - odx, dt[1] - ody, dt[2] - odz);
        cross3(&qx, &qy, &qz, dt[0] - odx, dt[1] - ody, dt[2] - odz, e1[0], e1[1], e1[2]);
        float v = pe1 * dot3(rx, ry, rz, qx, qy, qz);
        float w = pe1 * dot3(e2[0], e2[1], e2[2], qx, qy, qz);
        if ((w < *best) && !((u < MT_EPS) || (v < MT_EPS) || ((u + v) > 1 + MT_EPS) || (w < MT_EPS))) {
            *best = w;
            h->id = t;
            h->dist = w;
            h->rad[0] = s->rad[3 * t]; h->rad[1] = s->rad[3 * t + 1]; h->rad[2] = s->rad[3 * t + 2];
            h->pnt[0] = *best * rx + odx; h->pnt[1] = *best * ry + ody; h->pnt[2] = *best * rz + odz;
            /* VEC3_CUDA::device_rotate(rot_m, i, -1), vector.cuh:25-33 */
            const int reverse = -1;
            float tx = reverse * s->nrm[3 * t], ty = reverse * s->nrm[3 * t + 1], tz = reverse * s->nrm[3 * t + 2];
            h->norm[0] = tx * m12[0] + ty * m12[1] + tz * m12[2];
            h->norm[1] = tx * m12[4] + ty * m12[5] + tz * m12[6];
            h->norm[2] = tx * m12[8] + ty * m12[9] + tz * m12[10];
            h->norm[0] *= reverse; h->norm[1] *= reverse; h->norm[2] *= reverse;
        }
    }
}

/* Trixel.cu:41-172 intersect_voxel_cuda for one pixel.  cnt[0] += node pops, cnt[1] += MT tests,
 * cnt[2] = max(stack depth). */
static void trace_pixel(const orc_scene* s, const float* m12, const float* rmd, orc_hit* h, uint64_t* cnt) {
    float d = s->draw_distance;
    s32 stack[128];
    int front = 0;
    stack[0] = 0;
    h->id = -1;
    float odx = m12[3], ody = m12[7], odz = m12[11];
    float rx = -1 * (m12[0] * -rmd[0] + m12[1] * -rmd[1] + m12[2] * -rmd[2]);
    float ry = -1 * (m12[4] * -rmd[0] + m12[5] * -rmd[1] + m12[6] * -rmd[2]);
    float rz = -1 * (m12[8] * -rmd[0] + m12[9] * -rmd[1] + m12[10] * -rmd[2]);
    while (front >= 0) {
        s32 c = stack[front--];
        cnt[0]++;
        const float* B = s->Bo + 6 * (size_t)c;
        float t0x = rx > 0 ? B[0] * (1 / rx) : B[3] * (1 / rx);
        float t1x = rx > 0 ? B[3] * (1 / rx) : B[0] * (1 / rx);
        float t0y = ry > 0 ? B[1] * (1 / ry) : B[4] * (1 / ry);
        float t1y = ry > 0 ? B[4] * (1 / ry) : B[1] * (1 / ry);
        float t0z = rz > 0 ? B[2] * (1 / rz) : B[5] * (1 / rz);
        float t1z = rz > 0 ? B[5] * (1 / rz) : B[2] * (1 / rz);
        const u8* fl = s->flags + 3 * (size_t)c;
        float dir = ((rx * fl[0]) + (ry * fl[1]) + (rz * fl[2]));
        float ds = ((odx * fl[0]) + (ody * fl[1]) + (odz * fl[2]));
        float maxt0 = fmax(t0z + odz / rz, fmax(t0x + odx / rx, t0y + ody / ry));
        float mint1 = fmin(t1z + odz / rz, fmin(t1x + odx / rx, t1y + ody / ry));
        if (s->is_leaf[c]) {
            cnt[1]++;
            mt_test(s, s->tri[c], m12, rx, ry, rz, odx, ody, odz, &d, h);
            continue;
        }
        if (mint1 >= maxt0 - DEV_EPS && maxt0 > -DEV_EPS) {
            maxt0 *= dir; mint1 *= dir;
            float s1 = s->S1[c] + DEV_EPS + ds;
            float s2 = s->S2[c] + ds;
            if (maxt0 < s2 + DEV_EPS) {
                if (mint1 > s2 - DEV_EPS) stack[++front] = (s32)s->right[c];
                stack[++front] = (s32)s->left[c];
            } else {
                if (mint1 < s1 || maxt0 < s1) stack[++front] = (s32)s->left[c];
                stack[++front] = (s32)s->right[c];
            }
            if ((uint64_t)(front + 1) > cnt[2]) cnt[2] = (uint64_t)(front + 1);
        }
    }
}

/* Camera.cu:19-69 color_cam_cuda for a hit pixel (light at (2,2,2); `norm.x` used for the y term, :38;
 * reflection multiplied component-wise by the CAMERA-space ray, :39-41).  Returns 0x00RRGGBB. */
static inline u32 shade_pixel(const orc_hit* h, const float* rmd) {
    float sdx = 2 - h->pnt[0], sdy = 2 - h->pnt[1], sdz = 2 - h->pnt[2];
    device_normalize(&sdx, &sdy, &sdz);
    float dot_r_n = dot3(sdx, sdy, sdz, h->norm[0], h->norm[0], h->norm[2]);
    float rx = (sdx - (2 * dot_r_n * h->norm[0])) * rmd[0];
    float ry = (sdy - (2 * dot_r_n * h->norm[1])) * rmd[1];
    float rz = (sdz - (2 * dot_r_n * h->norm[2])) * rmd[2];
    float dif = .6 * fabsf(dot_r_n);
    float spc = powf(fabsf((rx + ry + rz)), 5) * .3;
    float pr = 0.0f, pg = 0.0f, pb = 0.0f;
    pr += (h->rad[0] * dif) + (1 * spc);
    pg += (h->rad[1] * dif) + (1 * spc);
    pb += (h->rad[2] * dif) + (1 * spc);
    float mx = fmaxf(fmaxf(pr, pg), pb);
    float cr = (pr / mx) * 255, cg = (pg / mx) * 255, cb = (pb / mx) * 255;
    /* float -> u8 truncation; NaN (all-zero radiance) is defined as 0 here */
    u32 r8 = cr == cr ? (u32)(u8)(int)cr : 0, g8 = cg == cg ? (u32)(u8)(int)cg : 0, b8 = cb == cb ? (u32)(u8)(int)cb : 0;
    return (r8 << 16) | (g8 << 8) | b8;
}

/* One frame over pixel rows [y0,y1): ids (s64 per pixel, -1 = miss) and 0x00RRGGBB colours
 * (set_cam_cuda background first, Camera.cu:12-18, then Phong on hit pixels).  ids/bgra are
 * full-frame buffers indexed by pixel i = y*W + x (row 0 = bottom).  counters[0..2] as above. */
void orc_render(const orc_scene* s, const float* m12, int y0, int y1, s64* ids, u32* bgra, uint64_t* counters) {
    uint64_t tot0 = 0, tot1 = 0, mx2 = 0;
    const u32 bg = ((u32)s->bg[3] << 24) | ((u32)s->bg[0] << 16) | ((u32)s->bg[1] << 8) | (u32)s->bg[2];
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot0, tot1) reduction(max : mx2)
    for (int y = y0; y < y1; y++) {
        uint64_t cnt[3] = {0, 0, 0};
        for (int x = 0; x < s->W; x++) {
            s64 i = (s64)y * s->W + x;
            float rmd[3];
            primary_ray(s->n_mod, s->u_mod, s->v_mod, s->W, i, rmd);
            orc_hit h;
            trace_pixel(s, m12, rmd, &h, cnt);
            if (ids) ids[i] = h.id;
            if (bgra) bgra[i] = h.id >= 0 ? shade_pixel(&h, rmd) : bg;
        }
        tot0 += cnt[0]; tot1 += cnt[1];
        if (cnt[2] > mx2) mx2 = cnt[2];
    }
    if (counters) { counters[0] += tot0; counters[1] += tot1; if (mx2 > counters[2]) counters[2] = mx2; }
}

/* Second oracle: closest hit by testing EVERY triangle with the same Moller-Trumbore (the idea of
 * the dormant kernel Trixel.cu:173-209, but with the object transform applied like the live one).
 * Ties resolve to the lowest triangle index here, so only hit/miss and distance are comparable. */
void orc_render_bruteforce(const orc_scene* s, const float* m12, int y0, int y1, s64* ids, float* dist) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < s->W; x++) {
            s64 i = (s64)y * s->W + x;
            float rmd[3];
            primary_ray(s->n_mod, s->u_mod, s->v_mod, s->W, i, rmd);
            float odx = m12[3], ody = m12[7], odz = m12[11];
            float rx = -1 * (m12[0] * -rmd[0] + m12[1] * -rmd[1] + m12[2] * -rmd[2]);
            float ry = -1 * (m12[4] * -rmd[0] + m12[5] * -rmd[1] + m12[6] * -rmd[2]);
            float rz = -1 * (m12[8] * -rmd[0] + m12[9] * -rmd[1] + m12[10] * -rmd[2]);
            float d = s->draw_distance;
            orc_hit h;
            h.id = -1;
            for (s64 t = 0; t < s->n; t++) mt_test(s, t, m12, rx, ry, rz, odx, ody, odz, &d, &h);
            ids[i] = h.id;
            if (dist) dist[i] = h.id >= 0 ? h.dist : 0.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* EXTENSIONS -- SURVEY.md section 8(f) items 3 and 4: what the reference leaves dormant        */
/* ------------------------------------------------------------------------------------------ */
/* The reference draws ONE object with ONE light and ONE ray per pixel; its code carries the stubs of more:
 *   - the light loop and the shadow test of color_cam_cuda, commented out (Camera.cu:28-34, 54):
 *         for each light { sd = light - pnt; in_shadow = (-1 != device_moller_trumbore(..., sd, ...)); if (!in_shadow) phong }
 *   - Camera::render_properites::sample_rate (Camera.h:46), never read;
 *   - a second object, created and registered but never rendered (WinMain.cpp:153,156,214-215), for which
 *     Camera::add_object keeps an object_list (Camera.cpp:118-130).
 * There is no reference BEHAVIOUR to be identical to, so this section DEFINES it, as the smallest
 * completion of those stubs that leaves today's output untouched: with one object, one light at (2,2,2), shadows off and
 * sample_rate <= 1, orc_render_scene() is orc_render() bit for bit (tests/test_oracle_cpu.py).  "parity unpinned" applies
 * to everything beyond that default: the GPU path is compared with THIS definition only.
 *
 *   objects     : every object of the camera's list is traversed in registration order with the running closest distance
 *                 carried over; the strict `w < best` of Trixel.cu:127 then makes the closest hit win and the first
 *                 registered object win ties.  Hit id = id_base[object] + triangle (id_base = triangles of the objects
 *                 registered before it).  Shading uses the winning object's matrix and colours.
 *   lights      : point_rad += phong(light) for each light in order (the commented loop), then the max-channel normalise.
 *   shadows     : a light contributes only if the segment from the hit point to it is free: a ray from the hit point
 *                 X = w*d - od (object space of the object that was hit) with the UNNORMALISED direction sd = light - pnt
 *                 (the very vector the commented call passes) hits no triangle OF THAT OBJECT other than the hit triangle
 *                 at a parameter 1e-4 < t < 1 (Moller-Trumbore with the u/v tests of Trixel.cu:127).  Objects do not
 *                 shadow one another (the commented call names one triangle list).  All lights shadowed: radiance 0,
 *                 0/0 in the normalise, pixel black (NaN -> 0).
 *   sample_rate : n >= 2 casts n x n rays per pixel through (ix + (a + .5)/n - .5, iy + (b + .5)/n - .5); the pixel is the
 *                 per-channel integer mean (floor) of the n*n shaded samples (background included); the hit id is that of
 *                 sample (n/2, n/2), the one nearest the pixel centre. */

/* Trixel.cu:41-172 for one object with the closest distance carried in/out: *best is updated and *h overwritten only when
 * this object has a closer hit (h->id >= 0 afterwards says so; the caller resets h->id = -1 before). */
static void trace_object(const orc_scene* s, const float* m12, const float* rmd, float* best, orc_hit* h, uint64_t* cnt) {
    s32 stack[128];
    int front = 0;
    stack[0] = 0;
    float odx = m12[3], ody = m12[7], odz = m12[11];
    float rx = -1 * (m12[0] * -rmd[0] + m12[1] * -rmd[1] + m12[2] * -rmd[2]);
    float ry = -1 * (m12[4] * -rmd[0] + m12[5] * -rmd[1] + m12[6] * -rmd[2]);
    float rz = -1 * (m12[8] * -rmd[0] + m12[9] * -rmd[1] + m12[10] * -rmd[2]);
    while (front >= 0) {
        s32 c = stack[front--];
        cnt[0]++;
        const float* B = s->Bo + 6 * (size_t)c;
        float t0x = rx > 0 ? B[0] * (1 / rx) : B[3] * (1 / rx);
        float t1x = rx > 0 ? B[3] * (1 / rx) : B[0] * (1 / rx);
        float t0y = ry > 0 ? B[1] * (1 / ry) : B[4] * (1 / ry);
        float t1y = ry > 0 ? B[4] * (1 / ry) : B[1] * (1 / ry);
        float t0z = rz > 0 ? B[2] * (1 / rz) : B[5] * (1 / rz);
        float t1z = rz > 0 ? B[5] * (1 / rz) : B[2] * (1 / rz);
        const u8* fl = s->flags + 3 * (size_t)c;
        float dir = ((rx * fl[0]) + (ry * fl[1]) + (rz * fl[2]));
        float ds = ((odx * fl[0]) + (ody * fl[1]) + (odz * fl[2]));
        float maxt0 = fmax(t0z + odz / rz, fmax(t0x + odx / rx, t0y + ody / ry));
        float mint1 = fmin(t1z + odz / rz, fmin(t1x + odx / rx, t1y + ody / ry));
        if (s->is_leaf[c]) {
            cnt[1]++;
            mt_test(s, s->tri[c], m12, rx, ry, rz, odx, ody, odz, best, h);
            continue;
        }
        if (mint1 >= maxt0 - DEV_EPS && maxt0 > -DEV_EPS) {
            maxt0 *= dir; mint1 *= dir;
            float s1 = s->S1[c] + DEV_EPS + ds;
            float s2 = s->S2[c] + ds;
            if (maxt0 < s2 + DEV_EPS) {
                if (mint1 > s2 - DEV_EPS) stack[++front] = (s32)s->right[c];
                stack[++front] = (s32)s->left[c];
            } else {
                if (mint1 < s1 || maxt0 < s1) stack[++front] = (s32)s->left[c];
                stack[++front] = (s32)s->right[c];
            }
        }
    }
}

/* Is the segment o' + t*d, 1e-4 < t < 1, blocked by a triangle of `s` other than `skip`?  The ray is handed over the way
 * the primary ray is (Trixel.cu:60-66,112): `o` = MINUS its origin in camera-relative object coordinates, so that
 * T = d_t - o and the slab offsets o/d are formed exactly as in the primary traversal.  A node is entered iff the segment
 * overlaps its box: tmax >= tmin, tmax >= 0, tmin <= 1 (plain float comparisons, false for NaN); leaves are tested when
 * reached.  Any-hit: the answer does not depend on the visit order. */
#define SHADOW_T_MIN 1e-4f
static int segment_blocked(const orc_scene* s, s64 skip, const float* o, const float* d) {
    s32 stack[128];
    int front = 0;
    stack[0] = 0;
    const float rx = d[0], ry = d[1], rz = d[2], odx = o[0], ody = o[1], odz = o[2];
    while (front >= 0) {
        s32 c = stack[front--];
        if (s->is_leaf[c]) {
            const s64 t = s->tri[c];
            if (t == skip) continue;
            const float* e1 = s->e1 + 3 * t; const float* e2 = s->e2 + 3 * t; const float* dt = s->dt + 3 * t;
            float px, py, pz, qx, qy, qz;
            cross3(&px, &py, &pz, rx, ry, rz, e2[0], e2[1], e2[2]);
            float f = dot3(px, py, pz, e1[0], e1[1], e1[2]);
            if (f < MT_EPS && f > -MT_EPS) continue;
            float pe1 = 1.0 / f;
            float u = pe1 * dot3(px, py, pz, dt[0] - odx, dt[1] - ody, dt[2] - odz);
            cross3(&qx, &qy, &qz, dt[0] - odx, dt[1] - ody, dt[2] - odz, e1[0], e1[1], e1[2]);
            float v = pe1 * dot3(rx, ry, rz, qx, qy, qz);
            float w = pe1 * dot3(e2[0], e2[1], e2[2], qx, qy, qz);
            if (!((u < MT_EPS) || (v < MT_EPS) || ((u + v) > 1 + MT_EPS) || (w < MT_EPS)) && w > SHADOW_T_MIN && w < 1.0f) return 1;
            continue;
        }
        for (int side = 0; side < 2; side++) {
            const s32 k = (s32)(side ? s->right[c] : s->left[c]);
            if (!s->is_leaf[k]) {
                const float* B = s->Bo + 6 * (size_t)k;
                float t0x = rx > 0 ? B[0] * (1 / rx) : B[3] * (1 / rx);
                float t1x = rx > 0 ? B[3] * (1 / rx) : B[0] * (1 / rx);
                float t0y = ry > 0 ? B[1] * (1 / ry) : B[4] * (1 / ry);
                float t1y = ry > 0 ? B[4] * (1 / ry) : B[1] * (1 / ry);
                float t0z = rz > 0 ? B[2] * (1 / rz) : B[5] * (1 / rz);
                float t1z = rz > 0 ? B[5] * (1 / rz) : B[2] * (1 / rz);
                float tmin = fmax(t0z + odz / rz, fmax(t0x + odx / rx, t0y + ody / ry));
                float tmax = fmin(t1z + odz / rz, fmin(t1x + odx / rx, t1y + ody / ry));
                if (!(tmax >= tmin && tmax >= 0.0f && tmin <= 1.0f)) continue;
            }
            stack[++front] = k;
        }
    }
    return 0;
}

/* color_cam_cuda with the light loop and the shadow test restored (Camera.cu:27-61).  d = object-space direction of the
 * primary ray, od = the object's translation column (what Trixel.cu:134-136 built pnt from). */
static inline u32 shade_lights(const orc_scene* s, const orc_hit* h, const float* rmd, const float* d, const float* od, int nlights,
                               const float* lights3, int shadows) {
    float pr = 0.0f, pg = 0.0f, pb = 0.0f;
    for (int l = 0; l < nlights; l++) {
        float sdx = lights3[3 * l] - h->pnt[0], sdy = lights3[3 * l + 1] - h->pnt[1], sdz = lights3[3 * l + 2] - h->pnt[2];
        if (shadows) {
            /* minus the hit point X = w*d - od, i.e. od - w*d */
            const float o[3] = {od[0] - h->dist * d[0], od[1] - h->dist * d[1], od[2] - h->dist * d[2]};
            const float sd[3] = {sdx, sdy, sdz};
            if (segment_blocked(s, h->id, o, sd)) continue;
        }
        device_normalize(&sdx, &sdy, &sdz);
        float dot_r_n = dot3(sdx, sdy, sdz, h->norm[0], h->norm[0], h->norm[2]);
        float rx = (sdx - (2 * dot_r_n * h->norm[0])) * rmd[0];
        float ry = (sdy - (2 * dot_r_n * h->norm[1])) * rmd[1];
        float rz = (sdz - (2 * dot_r_n * h->norm[2])) * rmd[2];
        float dif = .6 * fabsf(dot_r_n);
        float spc = powf(fabsf((rx + ry + rz)), 5) * .3;
        pr += (h->rad[0] * dif) + (1 * spc);
        pg += (h->rad[1] * dif) + (1 * spc);
        pb += (h->rad[2] * dif) + (1 * spc);
    }
    float mx = fmaxf(fmaxf(pr, pg), pb);
    float cr = (pr / mx) * 255, cg = (pg / mx) * 255, cb = (pb / mx) * 255;
    u32 r8 = cr == cr ? (u32)(u8)(int)cr : 0, g8 = cg == cg ? (u32)(u8)(int)cg : 0, b8 = cb == cb ? (u32)(u8)(int)cb : 0;
    return (r8 << 16) | (g8 << 8) | b8;
}

/* One frame of a scene of `nobj` objects (scenes[k] = the camera-side arrays of object k's mesh, all made for the same
 * camera; m12s + 12*k = its matrix; id_base[k] added to its triangle ids) with `nlights` lights, optional shadows and
 * sample_rate^2 rays per pixel, over pixel rows [y0,y1). */
void orc_render_scene(int nobj, const orc_scene* const* scenes, const float* m12s, const s64* id_base, int nlights, const float* lights3,
                      int shadows, int sample_rate, int y0, int y1, s64* ids, u32* bgra) {
    const orc_scene* s0 = scenes[0];
    const int n = sample_rate >= 2 ? sample_rate : 1;
    const u32 bg = ((u32)s0->bg[3] << 24) | ((u32)s0->bg[0] << 16) | ((u32)s0->bg[1] << 8) | (u32)s0->bg[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = y0; y < y1; y++) {
        uint64_t cnt[3] = {0, 0, 0};
        for (int x = 0; x < s0->W; x++) {
            const s64 i = (s64)y * s0->W + x;
            u32 sum_r = 0, sum_g = 0, sum_b = 0;
            s64 centre_id = -1;
            for (int b = 0; b < n; b++) {
                for (int a = 0; a < n; a++) {
                    float rmd[3];
                    if (n == 1) {
                        primary_ray(s0->n_mod, s0->u_mod, s0->v_mod, s0->W, i, rmd);
                    } else {
                        const float fx = (float)x + (((float)a + 0.5f) / (float)n - 0.5f), fy = (float)y + (((float)b + 0.5f) / (float)n - 0.5f);
                        rmd[0] = s0->n_mod[0] + s0->u_mod[0] * fx + s0->v_mod[0] * fy;
                        rmd[1] = s0->n_mod[1] + s0->u_mod[1] * fx + s0->v_mod[1] * fy;
                        rmd[2] = s0->n_mod[2] + s0->u_mod[2] * fx + s0->v_mod[2] * fy;
                        device_normalize(&rmd[0], &rmd[1], &rmd[2]);
                    }
                    float best = s0->draw_distance;
                    orc_hit hit;
                    int hit_obj = -1;
                    hit.id = -1;
                    for (int k = 0; k < nobj; k++) {
                        orc_hit hk;
                        hk.id = -1;
                        trace_object(scenes[k], m12s + 12 * k, rmd, &best, &hk, cnt);
                        if (hk.id >= 0) { hit = hk; hit_obj = k; }
                    }
                    u32 c = bg;
                    if (hit_obj >= 0) {
                        const float* m = m12s + 12 * hit_obj;
                        const float d[3] = {-1 * (m[0] * -rmd[0] + m[1] * -rmd[1] + m[2] * -rmd[2]), -1 * (m[4] * -rmd[0] + m[5] * -rmd[1] + m[6] * -rmd[2]),
                                            -1 * (m[8] * -rmd[0] + m[9] * -rmd[1] + m[10] * -rmd[2])};
                        const float od[3] = {m[3], m[7], m[11]};
                        c = shade_lights(scenes[hit_obj], &hit, rmd, d, od, nlights, lights3, shadows);
                    }
                    sum_r += (c >> 16) & 0xff; sum_g += (c >> 8) & 0xff; sum_b += c & 0xff;
                    if (a == n / 2 && b == n / 2) centre_id = hit_obj >= 0 ? id_base[hit_obj] + hit.id : -1;
                }
            }
            if (ids) ids[i] = centre_id;
            if (bgra) bgra[i] = n == 1 ? ((sum_r << 16) | (sum_g << 8) | sum_b) | (bg & 0xff000000u)
                                       : ((sum_r / (u32)(n * n)) << 16) | ((sum_g / (u32)(n * n)) << 8) | (sum_b / (u32)(n * n)) | (bg & 0xff000000u);
        }
    }
}

/* 64-bit FNV-1a over raw bytes (golden fixtures store these) */
uint64_t orc_fnv1a64(const void* data, size_t nbytes) {
    const u8* p = (const u8*)data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
