// rtb_build.cu -- the n log n tree build on the GPU.
//
// Same tree as Trixel::set_sorted_voxels + Trixel::create_kd (Trixel.h:386-473, 135-385; sort.h:11-60),
// bit for bit -- the host builder in rtb_host.cpp is the reference for it in the tests:
//   * six lists of triangle ids sorted by the per-triangle AABB keys x1,y1,z1,x0,y0,z0; the reference's
//     merge takes the right run on ties, i.e. equal keys are ordered by DESCENDING original index: a
//     stable LSD radix sort (cub::DeviceRadixSort) over the ids fed in descending order gives exactly that;
//   * breadth-first, one level per iteration: every node owns the same position range [l,r] in all six
//     lists; split list = first strict maximum of key[r]-key[l] in the order x1,x0,y1,y0,z1,z0; the other
//     five lists are stably partitioned by "position in the split list <= m".
// The split position is always the middle of the range, so the SHAPE of the tree -- how many nodes a level has, which
// of them are leaves -- follows from n alone (range sizes of one level differ by at most one); only the choice of the
// split list and the order inside the lists depend on the data.  The host therefore knows every level's extent in
// advance and never reads anything back during the build.  A level is five launches for ALL its nodes and ALL six
// lists at once:
//     level_nodes_kernel   one thread per node: own bounds from the list ends, leaf or split list, children created;
//     side_kernel          one thread per position of the split lists: one byte per triangle, "goes to the left child";
//     flags_kernel         one thread per position: its six "goes left" flags and their exclusive prefixes inside a
//                          512-position block, packed into one 64-bit word; the block's six totals;
//     block_scan_kernel    the totals become block bases (one 1024-thread block per list);
//     scatter_kernel       one thread per list position: the six stable partitions (offsets = prefix differences
//                          relative to the node's range start, so ranges never mix), node of position.
// (Round 1: flag / scan / scatter / copy per list and two blocking 4-byte read-backs per level, ~35 launches per level.)
// cub's radix sort and scan are library primitives used for the build only; the ray-cast hot path does not
// touch them.
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

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
struct Lists {
    const float* key;  // 6 x n
    const int* order;  // 6 x n (the level's input order)
    int n;
};
__device__ __forceinline__ float key_at(const Lists& L, int k, int pos) { return L.key[(long long)k * L.n + L.order[(long long)k * L.n + pos]]; }

// one thread per node of the level: the node's own bounds from the ends of its range in the six lists (Trixel.h:150,
// 345-350), then leaf (Trixel.h:194-202) or split choice (Trixel.h:172-205) and the two children (Trixel.h:329-352).
// child_scan == nullptr: every node of the level is interior, the children of the q-th node are 2q, 2q+1 of the next level.
__global__ void level_nodes_kernel(Lists L, int level_begin, int count, int next_begin, const int* __restrict__ child_scan, int* __restrict__ lo,
                                   int* __restrict__ hi, unsigned char* __restrict__ cut, int* __restrict__ tri, int* __restrict__ left,
                                   float* __restrict__ bounds, int* __restrict__ rec) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const int node = level_begin + q, l = lo[node], r = hi[node];
    float a[6], b[6];
    for (int k = 0; k < 6; k++) { a[k] = key_at(L, k, l); b[k] = key_at(L, k, r); }
    float* B = bounds + 6ll * node;  // x0,x1,y0,y1,z0,z1
    B[0] = a[3]; B[1] = b[0]; B[2] = a[4]; B[3] = b[1]; B[4] = a[5]; B[5] = b[2];
    if (r == l) {
        tri[node] = L.order[l];  // x1 list, Trixel.h:202 (the split list is inherited: written by the parent)
        left[node] = -1;
        return;
    }
    const int scan[6] = {0, 3, 1, 4, 2, 5};
    float best = __fsub_rn(b[0], a[0]);
    int c = 0;
    for (int s = 1; s < 6; s++) {
        const int k = scan[s];
        const float spread = __fsub_rn(b[k], a[k]);
        if (spread > best) { best = spread; c = k; }
    }
    cut[node] = (unsigned char)c;
    tri[node] = -1;
    const int m = l + (r - l) / 2;
    const int cl = next_begin + 2 * (child_scan ? child_scan[q] : q), cr = cl + 1;
    left[node] = cl;
    lo[cl] = l; hi[cl] = m; cut[cl] = (unsigned char)c;
    lo[cr] = m + 1; hi[cr] = r; cut[cr] = (unsigned char)c;
    // pre-order rank among interior nodes: the left subtree [l,m] holds m-l interior nodes and follows its parent directly
    rec[cl] = m > l ? rec[node] + 1 : -1;
    rec[cr] = r > m + 1 ? rec[node] + 1 + (m - l) : -1;
}

// ---- the level's six stable partitions ------------------------------------------------------------------------------------
// What position i needs in order to move is, per list k, (a) whether its element goes left and (b) how many elements of its
// node's range before it go left = prefix_k(i) - prefix_k(l), l = the range's start.  A device-wide scan of six-component
// vectors gave that (first round-2 version: cub::DeviceScan over a 24-byte type with the gathers inside its load phase, half
// of the build's time).  Here the prefix is split at 512-position blocks: flags_kernel packs, for every position, its six
// flags and its six IN-BLOCK exclusive prefixes (6 x 9 bits) into one 64-bit word and leaves the six block totals behind;
// block_scan_kernel turns the totals into block bases (a few thousand numbers); scatter_kernel reads two packed words (its
// own and its range start's) and two bases per list.  8 bytes per position and level instead of 24 + 72.
constexpr int kPartBlock = 512;  // positions per block: in-block prefixes fit 9 bits
// Which child every triangle of an interior node goes to (Trixel.h:237-259): its position in the node's split list is
// <= m.  One byte per triangle, written from the split list's side -- a 1-byte scatter into an array that stays in L2 --
// so that the five other lists can ask "left?" with a 1-byte gather instead of a 4-byte gather into a rank array per
// list (the round-1 builder kept six rank arrays up to date: 6 random reads and 6 random writes per triangle and level).
__global__ void side_kernel(int n, const int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                            const unsigned char* __restrict__ cut, const int* __restrict__ order, unsigned char* __restrict__ side) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int node = node_of_pos[i], l = lo[node], r = hi[node];
    if (r == l) return;
    side[order[(long long)cut[node] * n + i]] = i <= l + (r - l) / 2 ? 1 : 0;
}
struct IsInterior {
    const int* lo; const int* hi; int level_begin;
    __host__ __device__ __forceinline__ int operator()(int q) const { return hi[level_begin + q] > lo[level_begin + q] ? 1 : 0; }
};
__global__ void __launch_bounds__(kPartBlock) flags_kernel(int n, const int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                                                           const unsigned char* __restrict__ cut, const int* __restrict__ order,
                                                           const unsigned char* __restrict__ side, unsigned long long* __restrict__ packed,
                                                           int* __restrict__ block_sums /* 6 x nblocks */, int nblocks) {
    __shared__ int wsum[6][kPartBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kPartBlock + tid;
    unsigned f = 0;  // bit k: the element of list k at position i goes to the left child (Trixel.h:237-259)
    if (i < n) {
        const int node = node_of_pos[i], l = lo[node], r = hi[node];
        if (r > l) {
            const int c = cut[node];
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (k != c) f |= (unsigned)side[order[(long long)k * n + i]] << k;
        }
    }
    unsigned long long word = (unsigned long long)f << 54;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const unsigned b = __ballot_sync(0xffffffffu, (f >> k) & 1u);
        word |= (unsigned long long)__popc(b & ((1u << lane) - 1u)) << (9 * k);
        if (lane == 0) wsum[k][warp] = __popc(b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int before = 0;
        for (int w = 0; w < warp; w++) before += wsum[k][w];
        word += (unsigned long long)before << (9 * k);  // (<= 511: no carry into the next field)
    }
    if (i < n) packed[i] = word;
    if (tid < 6) {
        int total = 0;
        for (int w = 0; w < kPartBlock / 32; w++) total += wsum[tid][w];
        block_sums[(long long)tid * nblocks + blockIdx.x] = total;
    }
}
// exclusive scan of each list's block totals, in place: one block of 1024 threads per list
__global__ void __launch_bounds__(1024) block_scan_kernel(int* __restrict__ block_sums, int nblocks) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage temp;
    int* a = block_sums + (long long)blockIdx.x * nblocks;
    const int per = (nblocks + 1023) / 1024;
    const int begin = min(nblocks, (int)threadIdx.x * per), end = min(nblocks, begin + per);
    int sum = 0;
    for (int j = begin; j < end; j++) sum += a[j];
    int base;
    Scan(temp).ExclusiveSum(sum, base);
    for (int j = begin; j < end; j++) { const int v = a[j]; a[j] = base; base += v; }
}
// one thread per list position: stable partition of all six lists inside every node range; positions of leaves stay where
// they are.  Also moves every position to its child node.
__global__ void scatter_kernel(int n, const int* __restrict__ node_of_pos, const int* __restrict__ lo, const int* __restrict__ hi,
                               const unsigned char* __restrict__ cut, const int* __restrict__ left, const unsigned long long* __restrict__ packed,
                               const int* __restrict__ block_base, int nblocks, const int* __restrict__ order, int* __restrict__ order_out,
                               int* __restrict__ node_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int node = node_of_pos[i], l = lo[node], r = hi[node];
    if (r == l) {
#pragma unroll
        for (int k = 0; k < 6; k++) order_out[(long long)k * n + i] = order[(long long)k * n + i];
        node_out[i] = node;
        return;
    }
    const int c = cut[node], m = l + (r - l) / 2, cl = left[node];
    const unsigned long long wi = packed[i], wl = packed[l];
    const int bi = i / kPartBlock, bl = l / kPartBlock;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int t = order[(long long)k * n + i];
        int dst = i;
        if (k != c) {
            const int before_i = block_base[(long long)k * nblocks + bi] + (int)((wi >> (9 * k)) & 511u);
            const int before_l = block_base[(long long)k * nblocks + bl] + (int)((wl >> (9 * k)) & 511u);
            const int left_before = before_i - before_l;
            dst = ((wi >> (54 + k)) & 1u) ? l + left_before : (m + 1) + ((i - l) - left_before);
        }
        order_out[(long long)k * n + dst] = t;
    }
    node_out[i] = i <= m ? cl : cl + 1;
}
__global__ void root_init_kernel(int n, int* __restrict__ lo, int* __restrict__ hi, unsigned char* __restrict__ cut, int* __restrict__ rec) {
    lo[0] = 0; hi[0] = n - 1;
    cut[0] = 5;               // a single-triangle mesh: the root is a leaf with Trixel.h:152's initial value
    rec[0] = n > 1 ? 0 : -1;
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

namespace {
// The shape of the tree from n alone.  The split position is the middle of the range (Trixel.h:206), so the range sizes of
// one level take at most two consecutive values; a range of size k >= 2 leaves (k+1)/2 elements to its left child and
// k/2 to its right one.  Levels are numbered breadth-first, children in the order of their parents (Trixel.h:329).
struct LevelShape { int begin, count, interior; };
std::vector<LevelShape> tree_shape(int n) {
    std::vector<LevelShape> levels;
    long long size[2] = {n, 0}, many[2] = {1, 0};  // two (range size, number of ranges) classes
    int begin = 0;
    for (;;) {
        const long long count = many[0] + many[1];
        long long interior = 0;
        for (int k = 0; k < 2; k++) if (size[k] >= 2) interior += many[k];
        levels.push_back({begin, (int)count, (int)interior});
        if (interior == 0) break;
        begin += (int)count;
        long long nsize[4], nmany[4]; int classes = 0;
        for (int k = 0; k < 2; k++) {
            if (many[k] == 0 || size[k] < 2) continue;
            const long long parts[2] = {(size[k] + 1) / 2, size[k] / 2};
            for (long long part : parts) {
                int at = 0;
                while (at < classes && nsize[at] != part) at++;
                if (at == classes) { nsize[classes] = part; nmany[classes] = 0; classes++; }
                nmany[at] += many[k];
            }
        }
        // (classes <= 2: sizes k and k+1 halve into at most two consecutive values)
        size[0] = nsize[0]; many[0] = nmany[0];
        size[1] = classes > 1 ? nsize[1] : 0; many[1] = classes > 1 ? nmany[1] : 0;
        if (classes > 2) { levels.clear(); return levels; }  // cannot happen; the caller reports it
    }
    return levels;
}

// The work arrays (about 130 bytes per triangle) come from a stream-ordered pool OWNED BY THIS LIBRARY, one per device,
// which keeps up to 2 GB between builds: mapping and unmapping gigabytes costs more than a ten-million-triangle build
// (measured: up to 0.5 s).  The process's default pool and its release threshold are not touched.
cudaMemPool_t build_pool(int device, std::string& err) {
    static std::mutex lock;
    static cudaMemPool_t pools[64] = {};
    std::lock_guard<std::mutex> guard(lock);
    if (device < 0 || device >= 64) { err = "build_tree_gpu: device index out of range"; return nullptr; }
    if (!pools[device]) {
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&pools[device], &props);
        unsigned long long keep = 2ull << 30;
        if (e == cudaSuccess) e = cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) { err = std::string("build_tree_gpu: memory pool: ") + cudaGetErrorString(e); cudaGetLastError(); pools[device] = nullptr; }
    }
    return pools[device];
}
}  // namespace

// Builds the tree of the triangle soup `d_points9` (device, 9 floats per triangle) on the current device and leaves
// it there (`T`; release with free_device_tree).  Returns an empty string on success.
std::string build_tree_gpu(const float* d_points9, int64_t n64, DeviceTree& T) {
    using clock = std::chrono::steady_clock;
    std::string err;
    if (n64 <= 0 || n64 > 0x1fffffff) return "build_tree_gpu: bad triangle count";
    const int n = (int)n64;
    const int N = 2 * n - 1;
    const int nblocks = (n + kPartBlock - 1) / kPartBlock;
    const auto t_begin = clock::now();
    const std::vector<LevelShape> levels = tree_shape(n);
    if (levels.empty() || levels.back().begin + levels.back().count != N) return "build_tree_gpu: level shape does not add up";

    float *key = nullptr, *bounds = nullptr, *s1 = nullptr, *s2 = nullptr;
    unsigned *ukey_in = nullptr, *ukey_out = nullptr;
    int *ids_in = nullptr, *order[2] = {nullptr, nullptr}, *node_of_pos[2] = {nullptr, nullptr}, *child_scan = nullptr;
    int *lo = nullptr, *hi = nullptr, *left = nullptr, *tri = nullptr, *rec = nullptr;
    unsigned long long* packed = nullptr;
    int* block_sums = nullptr;
    unsigned char *cut = nullptr, *side = nullptr;
    void* temp = nullptr;
    void *keep_base = nullptr, *work_base = nullptr;
    size_t temp_bytes = 0, need = 0;
    int launches = 0, cur = 0, device = 0, max_count = 1;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float sort_ms = 0.0f;
    cudaMemPool_t pool = nullptr;
    thrust::counting_iterator<int> counting(0);
    double seconds_before_sort = 0.0;

    // two allocations: the tree itself (kept, DeviceTree::arena) and the work arrays (returned to the pool at the end)
    Arena keep, work;
    keep.reserve<float>(6 * (size_t)N); keep.reserve<int>((size_t)N); keep.reserve<int>((size_t)N); keep.reserve<int>((size_t)N);
    keep.reserve<unsigned char>((size_t)N); keep.reserve<float>((size_t)N); keep.reserve<float>((size_t)N);
    for (const LevelShape& lv : levels) max_count = lv.count > max_count ? lv.count : max_count;
    cub::DeviceRadixSort::SortPairs(nullptr, need, ukey_in, ukey_out, ids_in, ids_in, n);
    temp_bytes = need;
    {
        IsInterior isrc{};
        thrust::transform_iterator<IsInterior, thrust::counting_iterator<int>, int> iin(counting, isrc);
        cub::DeviceScan::ExclusiveSum(nullptr, need, iin, child_scan, max_count);
        temp_bytes = temp_bytes > need ? temp_bytes : need;
    }
    work.reserve<float>(6 * (size_t)n); work.reserve<unsigned>((size_t)n); work.reserve<unsigned>((size_t)n); work.reserve<int>((size_t)n);
    work.reserve<int>(6 * (size_t)n); work.reserve<int>(6 * (size_t)n); work.reserve<unsigned char>((size_t)n);  // order x2, side
    work.reserve<int>((size_t)n); work.reserve<int>((size_t)n); work.reserve<int>((size_t)max_count);       // node_of_pos x2, child_scan
    work.reserve<int>((size_t)N); work.reserve<int>((size_t)N);                                             // lo, hi
    work.reserve<unsigned long long>((size_t)n); work.reserve<int>(6 * (size_t)nblocks);
    work.reserve<char>(temp_bytes);
    RTB_BUILD_CUDA(cudaGetDevice(&device));
    pool = build_pool(device, err);
    if (!pool) goto done;
    RTB_BUILD_CUDA(cudaMalloc(&keep_base, keep.size));
    RTB_BUILD_CUDA(cudaMallocFromPoolAsync(&work_base, work.size, pool, (cudaStream_t)0));
    RTB_BUILD_CUDA(cudaEventCreate(&ev[0]));
    RTB_BUILD_CUDA(cudaEventCreate(&ev[1]));
    keep.base = (char*)keep_base; work.base = (char*)work_base;
    bounds = keep.take<float>(6 * (size_t)N); left = keep.take<int>((size_t)N); tri = keep.take<int>((size_t)N); rec = keep.take<int>((size_t)N);
    cut = keep.take<unsigned char>((size_t)N); s1 = keep.take<float>((size_t)N); s2 = keep.take<float>((size_t)N);
    key = work.take<float>(6 * (size_t)n); ukey_in = work.take<unsigned>((size_t)n); ukey_out = work.take<unsigned>((size_t)n);
    ids_in = work.take<int>((size_t)n); order[0] = work.take<int>(6 * (size_t)n); order[1] = work.take<int>(6 * (size_t)n);
    side = work.take<unsigned char>((size_t)n); node_of_pos[0] = work.take<int>((size_t)n); node_of_pos[1] = work.take<int>((size_t)n);
    child_scan = work.take<int>((size_t)max_count);
    lo = work.take<int>((size_t)N); hi = work.take<int>((size_t)N);
    packed = work.take<unsigned long long>((size_t)n); block_sums = work.take<int>(6 * (size_t)nblocks);
    temp = work.take<char>(temp_bytes);
    seconds_before_sort = std::chrono::duration<double>(clock::now() - t_begin).count();

    // ---- six sorted lists -----------------------------------------------------------------------------
    RTB_BUILD_CUDA(cudaEventRecord(ev[0], (cudaStream_t)0));
    keys_kernel<<<blocks(n), 256>>>(d_points9, n, key); launches++;
    for (int k = 0; k < 6; k++) {
        sort_input_kernel<<<blocks(n), 256>>>(key + (size_t)k * n, n, ukey_in, ids_in); launches++;
        size_t tb = temp_bytes;
        RTB_BUILD_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, ukey_in, ukey_out, ids_in, order[0] + (size_t)k * n, n)); launches += 5;  // histogram + four onesweep passes
    }
    RTB_BUILD_CUDA(cudaEventRecord(ev[1], (cudaStream_t)0));

    // ---- level-synchronous partition: nothing is read back, the host knows every level's extent ---------------
    RTB_BUILD_CUDA(cudaMemsetAsync(node_of_pos[0], 0, sizeof(int) * (size_t)n, (cudaStream_t)0));
    root_init_kernel<<<1, 1>>>(n, lo, hi, cut, rec); launches++;
    for (const LevelShape& lv : levels) {
        const Lists L{key, order[cur], n};
        const int next_begin = lv.begin + lv.count;
        const bool mixed = lv.interior > 0 && lv.interior < lv.count;  // leaves among the level's nodes: children need a scan
        if (mixed) {
            thrust::transform_iterator<IsInterior, thrust::counting_iterator<int>, int> iin(counting, IsInterior{lo, hi, lv.begin});
            size_t tb = temp_bytes;
            RTB_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, iin, child_scan, lv.count)); launches += 2;
        }
        level_nodes_kernel<<<blocks(lv.count), 256>>>(L, lv.begin, lv.count, next_begin, mixed ? child_scan : nullptr, lo, hi, cut, tri, left, bounds, rec);
        launches++;
        if (lv.interior == 0) break;
        side_kernel<<<blocks(n), 256>>>(n, node_of_pos[cur], lo, hi, cut, order[cur], side); launches++;
        flags_kernel<<<nblocks, kPartBlock>>>(n, node_of_pos[cur], lo, hi, cut, order[cur], side, packed, block_sums, nblocks); launches++;
        block_scan_kernel<<<6, 1024>>>(block_sums, nblocks); launches++;
        scatter_kernel<<<blocks(n), 256>>>(n, node_of_pos[cur], lo, hi, cut, left, packed, block_sums, nblocks, order[cur], order[cur ^ 1], node_of_pos[cur ^ 1]);
        launches++;
        cur ^= 1;
    }
    split_planes_kernel<<<blocks(N), 256>>>(N, left, cut, bounds, s1, s2); launches++;
    RTB_BUILD_CUDA(cudaGetLastError());
    // ---- the tree stays on the device; the root's box and triangle are all the host needs ----------------------------
    RTB_BUILD_CUDA(cudaMemcpy(T.root_bounds, bounds, sizeof(float) * 6, cudaMemcpyDeviceToHost));  // (synchronises with everything above)
    RTB_BUILD_CUDA(cudaMemcpy(&T.root_tri, tri, sizeof(int), cudaMemcpyDeviceToHost));
    RTB_BUILD_CUDA(cudaEventElapsedTime(&sort_ms, ev[0], ev[1]));
    T.num_tri = n; T.num_nodes = N;
    T.bounds = bounds; T.left = left; T.tri = tri; T.cut = cut; T.s1 = s1; T.s2 = s2; T.rec = rec;
    T.arena = keep_base;
    keep_base = nullptr;
    T.seconds_sort = seconds_before_sort + (double)sort_ms * 1e-3;  // allocation + keys + six sorts (device time)
    T.seconds_partition = std::chrono::duration<double>(clock::now() - t_begin).count() - T.seconds_sort;
    T.launches = launches;

done:
    if (work_base) cudaFreeAsync(work_base, (cudaStream_t)0);
    cudaFree(keep_base);
    for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    if (!err.empty()) cudaGetLastError();
    return err;
}

}  // namespace rtb
