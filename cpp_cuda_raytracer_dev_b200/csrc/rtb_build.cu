// rtb_build.cu -- the n log n tree build on the GPU.
//
// Same tree as Trixel::set_sorted_voxels + Trixel::create_kd (Trixel.h:386-473, 135-385; sort.h:11-60),
// bit for bit -- the host builder in rtb_host.cpp is the reference for it in the tests:
//   * six lists of triangle ids sorted by the per-triangle AABB keys x1,y1,z1,x0,y0,z0; the reference's
//     merge takes the right run on ties, i.e. equal keys are ordered by DESCENDING original index: a
//     stable LSD radix sort (cub::DeviceRadixSort) over the ids fed in descending order gives exactly that;
//   * breadth-first, one level per iteration: every node owns the same position range [l,r] in all six
//     lists; split list = first strict maximum of key[r]-key[l] in the order x1,x0,y1,y0,z1,z0; the other
//     five lists are stably partitioned by "position in the split list <= m" -- here one flag array, one
//     device-wide exclusive scan (cub::DeviceScan) and one scatter per list and level, for all nodes of the
//     level at once (positions of different nodes never mix because offsets are taken relative to l);
//   * children are numbered in node order (exclusive scan over the level's interior flags), bounds come
//     from the list ends, leaves take the x1 list's triangle.
// cub's radix sort and scan are library primitives used for the build only; the ray-cast hot path does not
// touch them.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <chrono>
#include <string>
#include <vector>

#include "rtb_host.hpp"

namespace rtb {
namespace {

#define RTB_BUILD_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); goto done; } \
    } while (0)

__device__ __forceinline__ float min3f(float a, float b, float c) { const float m = b < c ? b : c; return a < m ? a : m; }
__device__ __forceinline__ float max3f(float a, float b, float c) { const float m = b > c ? b : c; return a > m ? a : m; }

// list numbering of the reference (Trixel.h:217-236): 0=x1 1=y1 2=z1 3=x0 4=y0 5=z0
__global__ void keys_kernel(const float* __restrict__ points9, int n, float* __restrict__ key /* 6 x n */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = points9 + 9ll * i;
    key[0ll * n + i] = max3f(p[0], p[3], p[6]); key[3ll * n + i] = min3f(p[0], p[3], p[6]);
    key[1ll * n + i] = max3f(p[1], p[4], p[7]); key[4ll * n + i] = min3f(p[1], p[4], p[7]);
    key[2ll * n + i] = max3f(p[2], p[5], p[8]); key[5ll * n + i] = min3f(p[2], p[5], p[8]);
}
// sort input: ids in descending order with their keys mapped to order-preserving unsigned integers
// (-0 is first folded onto +0: the reference compares floats, for which they are equal)
__global__ void sort_input_kernel(const float* __restrict__ key, int n, unsigned* __restrict__ ukey, int* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = n - 1 - i;
    unsigned u = __float_as_uint(__fadd_rn(key[t], 0.0f));
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
    ukey[i] = u;
    ids[i] = t;
}
__global__ void rank_kernel(const int* __restrict__ order, int n, int* __restrict__ rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[order[i]] = i;
}

struct Lists {
    const float* key;  // 6 x n
    int* order;        // 6 x n (current)
    int* rank;         // 6 x n
    int n;
};
__device__ __forceinline__ float key_at(const Lists& L, int k, int pos) { return L.key[(long long)k * L.n + L.order[(long long)k * L.n + pos]]; }

// one thread per node of the level: leaf or split choice (Trixel.h:172-205)
__global__ void level_nodes_kernel(Lists L, int level_begin, int count, const int* __restrict__ lo, const int* __restrict__ hi,
                                   const int* __restrict__ parent, unsigned char* __restrict__ cut, int* __restrict__ tri,
                                   int* __restrict__ interior /* count */) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const int node = level_begin + q, l = lo[node], r = hi[node];
    if (r == l) {
        cut[node] = node == 0 ? 5 : cut[parent[node]];  // Trixel.h:152,194
        tri[node] = L.order[l];                          // x1 list, Trixel.h:202
        interior[q] = 0;
        return;
    }
    const int scan[6] = {0, 3, 1, 4, 2, 5};
    float best = __fsub_rn(key_at(L, 0, r), key_at(L, 0, l));
    int c = 0;
    for (int s = 1; s < 6; s++) {
        const int k = scan[s];
        const float spread = __fsub_rn(key_at(L, k, r), key_at(L, k, l));
        if (spread > best) { best = spread; c = k; }
    }
    cut[node] = (unsigned char)c;
    tri[node] = -1;
    interior[q] = 1;
}
// one thread per list position: does the element go to the left child?  (Trixel.h:237-259)
__global__ void flags_kernel(Lists L, int k, const int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                             const unsigned char* __restrict__ cut, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n) return;
    const int node = node_of_pos[i], l = lo[node], r = hi[node];
    int f = 0;
    if (r > l && cut[node] != k) {
        const int m = l + (r - l) / 2;
        f = L.rank[(long long)cut[node] * L.n + L.order[(long long)k * L.n + i]] <= m;
    }
    flag[i] = f;
}
// stable partition of list k inside every node range, using the device-wide exclusive scan of the flags
__global__ void scatter_kernel(Lists L, int k, const int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                               const unsigned char* __restrict__ cut, const int* __restrict__ flag, const int* __restrict__ scan,
                               int* __restrict__ order_out /* n, list k */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n) return;
    const int node = node_of_pos[i], l = lo[node], r = hi[node];
    const int t = L.order[(long long)k * L.n + i];
    int dst = i;
    if (r > l && cut[node] != k) {
        const int m = l + (r - l) / 2;
        const int left_before = scan[i] - scan[l];
        dst = flag[i] ? l + left_before : (m + 1) + ((i - l) - left_before);
    }
    order_out[dst] = t;
    L.rank[(long long)k * L.n + t] = dst;
}
// one thread per node of the level: create the two children (Trixel.h:329-352)
__global__ void children_kernel(Lists L, int level_begin, int count, int next_begin, const int* __restrict__ child_scan, int* __restrict__ lo,
                                int* __restrict__ hi, int* __restrict__ parent, int* __restrict__ left, float* __restrict__ bounds,
                                int* __restrict__ rec) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const int node = level_begin + q, l = lo[node], r = hi[node];
    if (r == l) { left[node] = -1; return; }
    const int m = l + (r - l) / 2;
    const int cl = next_begin + 2 * child_scan[q], cr = cl + 1;
    left[node] = cl;
    lo[cl] = l; hi[cl] = m; parent[cl] = node;
    lo[cr] = m + 1; hi[cr] = r; parent[cr] = node;
    // pre-order rank among interior nodes: the left subtree [l,m] holds m-l interior nodes and follows its parent directly
    rec[cl] = m > l ? rec[node] + 1 : -1;
    rec[cr] = r > m + 1 ? rec[node] + 1 + (m - l) : -1;
    for (int c = 0; c < 2; c++) {
        const int a = c ? m + 1 : l, b = c ? r : m;
        float* B = bounds + 6ll * (c ? cr : cl);  // x0,x1,y0,y1,z0,z1 (Trixel.h:345-350)
        B[0] = key_at(L, 3, a); B[1] = key_at(L, 0, b);
        B[2] = key_at(L, 4, a); B[3] = key_at(L, 1, b);
        B[4] = key_at(L, 5, a); B[5] = key_at(L, 2, b);
    }
}
__global__ void reassign_kernel(int n, int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                                const int* __restrict__ left) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int node = node_of_pos[i], cl = left[node];
    if (cl < 0) return;
    const int l = lo[node], r = hi[node], m = l + (r - l) / 2;
    node_of_pos[i] = i <= m ? cl : cl + 1;
}
__global__ void root_bounds_kernel(Lists L, float* __restrict__ bounds) {
    const int n = L.n;
    bounds[0] = key_at(L, 3, 0); bounds[1] = key_at(L, 0, n - 1);
    bounds[2] = key_at(L, 4, 0); bounds[3] = key_at(L, 1, n - 1);
    bounds[4] = key_at(L, 5, 0); bounds[5] = key_at(L, 2, n - 1);
}
// s1 = left child's max, s2 = right child's min on the split axis (Trixel.h:353-376)
__global__ void split_planes_kernel(int num_nodes, const int* __restrict__ left, const unsigned char* __restrict__ cut,
                                    const float* __restrict__ bounds, float* __restrict__ s1, float* __restrict__ s2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int cl = left[i];
    if (cl < 0) { s1[i] = 0.0f; s2[i] = 0.0f; return; }
    const int axis = cut[i] % 3;
    s1[i] = bounds[6ll * cl + 2 * axis + 1];
    s2[i] = bounds[6ll * (cl + 1) + 2 * axis];
}

inline unsigned blocks(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void free_device_tree(DeviceTree& t) {
    cudaFree(t.arena);
    t = DeviceTree();
}

namespace {
// bump allocator over one cudaMalloc: the build needs 22 arrays, and 22 cudaMalloc/cudaFree pairs cost more than the
// kernels of a million-triangle build
struct Arena {
    char* base = nullptr;
    size_t used = 0, size = 0;
    template <class T> void reserve(size_t count) { size += (sizeof(T) * count + 255) & ~(size_t)255; }
    template <class T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += (sizeof(T) * count + 255) & ~(size_t)255;
        return p;
    }
};
}  // namespace

std::string download_tree(const DeviceTree& D, HostTree& T) {
    std::string err;
    const size_t N = (size_t)D.num_nodes;
    T.num_tri = D.num_tri; T.num_nodes = D.num_nodes;
    T.bounds.resize(N * 6); T.left.resize(N); T.tri.resize(N); T.cut_flag.resize(N); T.s1.resize(N); T.s2.resize(N);
    RTB_BUILD_CUDA(cudaMemcpy(T.bounds.data(), D.bounds, sizeof(float) * 6 * N, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(T.left.data(), D.left, sizeof(int) * N, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(T.tri.data(), D.tri, sizeof(int) * N, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(T.cut_flag.data(), D.cut, N, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(T.s1.data(), D.s1, sizeof(float) * N, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(T.s2.data(), D.s2, sizeof(float) * N, cudaMemcpyDeviceToHost));
    T.seconds_sort = D.seconds_sort; T.seconds_partition = D.seconds_partition;
done:
    if (!err.empty()) cudaGetLastError();
    return err;
}

// Builds the tree of the triangle soup `d_points9` (device, 9 floats per triangle) on the current device and leaves
// it there (`T`; release with free_device_tree).  Returns an empty string on success.
std::string build_tree_gpu(const float* d_points9, int64_t n64, DeviceTree& T) {
    using clock = std::chrono::steady_clock;
    std::string err;
    if (n64 <= 0 || n64 > 0x1fffffff) return "build_tree_gpu: bad triangle count";
    const int n = (int)n64;
    const int N = 2 * n - 1;
    const auto t_begin = clock::now();
    auto t_sorted = t_begin;

    float *key = nullptr, *bounds = nullptr, *s1 = nullptr, *s2 = nullptr;
    unsigned *ukey_in = nullptr, *ukey_out = nullptr;
    int *ids_in = nullptr, *order = nullptr, *order_tmp = nullptr, *rank = nullptr, *node_of_pos = nullptr, *flag = nullptr, *scan = nullptr;
    int *lo = nullptr, *hi = nullptr, *parent = nullptr, *left = nullptr, *tri = nullptr, *interior = nullptr, *child_scan = nullptr, *rec = nullptr;
    unsigned char* cut = nullptr;
    void* temp = nullptr;
    void *keep_base = nullptr, *work_base = nullptr;
    size_t temp_bytes = 0, need = 0;
    Lists L{};
    int level_begin = 0, level_end = 1;
    int launches = 0;

    // two allocations: the tree itself (kept, DeviceTree::arena) and the work arrays (freed at the end)
    Arena keep, work;
    keep.reserve<float>(6 * (size_t)N); keep.reserve<int>((size_t)N); keep.reserve<int>((size_t)N); keep.reserve<int>((size_t)N);
    keep.reserve<unsigned char>((size_t)N); keep.reserve<float>((size_t)N); keep.reserve<float>((size_t)N);
    cub::DeviceRadixSort::SortPairs(nullptr, need, ukey_in, ukey_out, ids_in, order, n);
    temp_bytes = need;
    cub::DeviceScan::ExclusiveSum(nullptr, need, flag, scan, n);
    temp_bytes = temp_bytes > need ? temp_bytes : need;
    work.reserve<float>(6 * (size_t)n); work.reserve<unsigned>((size_t)n); work.reserve<unsigned>((size_t)n); work.reserve<int>((size_t)n);
    work.reserve<int>(6 * (size_t)n); work.reserve<int>((size_t)n); work.reserve<int>(6 * (size_t)n);
    for (int k = 0; k < 5; k++) work.reserve<int>((size_t)n);  // node_of_pos, flag, scan, interior, child_scan
    for (int k = 0; k < 3; k++) work.reserve<int>((size_t)N);  // lo, hi, parent
    work.reserve<char>(temp_bytes);
    RTB_BUILD_CUDA(cudaMalloc(&keep_base, keep.size));
    {
        // The work arrays come from the device's stream-ordered pool, which is told to keep what it is given back:
        // mapping and unmapping gigabytes costs more than a ten-million-triangle build (measured: up to 0.5 s).
        int device = 0;
        cudaMemPool_t pool = nullptr;
        RTB_BUILD_CUDA(cudaGetDevice(&device));
        RTB_BUILD_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep_all = ~0ull;
        RTB_BUILD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
        RTB_BUILD_CUDA(cudaMallocAsync(&work_base, work.size, (cudaStream_t)0));
    }
    keep.base = (char*)keep_base; work.base = (char*)work_base;
    bounds = keep.take<float>(6 * (size_t)N); left = keep.take<int>((size_t)N); tri = keep.take<int>((size_t)N); rec = keep.take<int>((size_t)N);
    cut = keep.take<unsigned char>((size_t)N); s1 = keep.take<float>((size_t)N); s2 = keep.take<float>((size_t)N);
    key = work.take<float>(6 * (size_t)n); ukey_in = work.take<unsigned>((size_t)n); ukey_out = work.take<unsigned>((size_t)n);
    ids_in = work.take<int>((size_t)n); order = work.take<int>(6 * (size_t)n); order_tmp = work.take<int>((size_t)n);
    rank = work.take<int>(6 * (size_t)n); node_of_pos = work.take<int>((size_t)n); flag = work.take<int>((size_t)n); scan = work.take<int>((size_t)n);
    interior = work.take<int>((size_t)n); child_scan = work.take<int>((size_t)n);
    lo = work.take<int>((size_t)N); hi = work.take<int>((size_t)N); parent = work.take<int>((size_t)N);
    temp = work.take<char>(temp_bytes);

    // ---- six sorted lists -----------------------------------------------------------------------------
    keys_kernel<<<blocks(n), 256>>>(d_points9, n, key); launches++;
    for (int k = 0; k < 6; k++) {
        sort_input_kernel<<<blocks(n), 256>>>(key + (size_t)k * n, n, ukey_in, ids_in); launches++;
        size_t tb = temp_bytes;
        RTB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, ukey_in, ukey_out, ids_in, order + (size_t)k * n, n)); launches += 5;  // histogram + four onesweep passes
        rank_kernel<<<blocks(n), 256>>>(order + (size_t)k * n, n, rank + (size_t)k * n); launches++;
    }
    RTB_BUILD_CUDA(cudaDeviceSynchronize());
    t_sorted = clock::now();

    // ---- level-synchronous partition --------------------------------------------------------------------
    L.key = key; L.order = order; L.rank = rank; L.n = n;
    RTB_BUILD_CUDA(cudaMemset(node_of_pos, 0, sizeof(int) * (size_t)n));
    {
        const int zero = 0, last = n - 1;
        RTB_BUILD_CUDA(cudaMemcpy(lo, &zero, sizeof(int), cudaMemcpyHostToDevice));
        RTB_BUILD_CUDA(cudaMemcpy(hi, &last, sizeof(int), cudaMemcpyHostToDevice));
        RTB_BUILD_CUDA(cudaMemcpy(parent, &zero, sizeof(int), cudaMemcpyHostToDevice));
        const int root_rec = n > 1 ? 0 : -1;
        RTB_BUILD_CUDA(cudaMemcpy(rec, &root_rec, sizeof(int), cudaMemcpyHostToDevice));
    }
    root_bounds_kernel<<<1, 1>>>(L, bounds); launches++;
    while (level_begin < level_end) {
        const int count = level_end - level_begin;
        level_nodes_kernel<<<blocks(count), 256>>>(L, level_begin, count, lo, hi, parent, cut, tri, interior); launches++;
        size_t tb = temp_bytes;
        RTB_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, interior, child_scan, count)); launches += 2;
        int last_flag = 0, last_scan = 0;
        RTB_BUILD_CUDA(cudaMemcpy(&last_flag, interior + (count - 1), sizeof(int), cudaMemcpyDeviceToHost));
        RTB_BUILD_CUDA(cudaMemcpy(&last_scan, child_scan + (count - 1), sizeof(int), cudaMemcpyDeviceToHost));
        const int num_interior = last_flag + last_scan;
        if (num_interior > 0) {
            for (int k = 0; k < 6; k++) {
                flags_kernel<<<blocks(n), 256>>>(L, k, node_of_pos, lo, hi, cut, flag); launches++;
                tb = temp_bytes;
                RTB_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, flag, scan, n)); launches += 2;
                scatter_kernel<<<blocks(n), 256>>>(L, k, node_of_pos, lo, hi, cut, flag, scan, order_tmp); launches++;
                RTB_BUILD_CUDA(cudaMemcpyAsync(order + (size_t)k * n, order_tmp, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice));
            }
        }
        children_kernel<<<blocks(count), 256>>>(L, level_begin, count, level_end, child_scan, lo, hi, parent, left, bounds, rec); launches++;
        reassign_kernel<<<blocks(n), 256>>>(n, node_of_pos, lo, hi, left); launches++;
        level_begin = level_end;
        level_end += 2 * num_interior;
    }
    split_planes_kernel<<<blocks(N), 256>>>(N, left, cut, bounds, s1, s2); launches++;
    RTB_BUILD_CUDA(cudaDeviceSynchronize());
    if (level_end != N) { err = "build_tree_gpu: node count mismatch"; goto done; }

    // ---- the tree stays on the device ------------------------------------------------------------------
    RTB_BUILD_CUDA(cudaMemcpy(T.root_bounds, bounds, sizeof(float) * 6, cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaMemcpy(&T.root_tri, tri, sizeof(int), cudaMemcpyDeviceToHost));
    T.num_tri = n; T.num_nodes = N;
    T.bounds = bounds; T.left = left; T.tri = tri; T.cut = cut; T.s1 = s1; T.s2 = s2; T.rec = rec;
    T.arena = keep_base;
    keep_base = nullptr;
    T.seconds_sort = std::chrono::duration<double>(t_sorted - t_begin).count();
    T.seconds_partition = std::chrono::duration<double>(clock::now() - t_sorted).count();
    T.launches = launches;

done:
    if (work_base) cudaFreeAsync(work_base, (cudaStream_t)0);
    cudaFree(keep_base);
    if (!err.empty()) cudaGetLastError();
    return err;
}

}  // namespace rtb
