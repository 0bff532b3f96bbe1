// Distributed setup kernels: DFS box order for the work
// partition, box masks, local-particle compaction, box -> user-rank CSR and the local
// target flags.  They restate the OpenCL kernels of boxtree/distributed/partition.py and
// boxtree/distributed/local_tree.py (cited per kernel).
#include "common.cuh"
#include "scan.cuh"
#include "../../include/boxtree_b200.h"

namespace bt {

// ---- c1: get_box_ids_dfs_order, partition.py:38-57 -----------------------------------
// The reference pops an explicit stack after pushing children in Morton order, i.e. it
// visits the HIGHEST Morton child first: pre-order with reversed child order.
template <int DIM>
__global__ void dist_subtree_size_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                         const int* __restrict__ child_ids, int* __restrict__ size)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        int sz = 1;
#pragma unroll
        for (int m = 0; m < NB; ++m) { const int c = child_ids[m * aligned + b]; if (c > 0) sz += size[c]; }
        size[b] = sz;
    }
}
template <int DIM>
__global__ void dist_dfs_order_kernel(const int* __restrict__ level_start, int lev, int aligned,
                                      const int* __restrict__ child_ids, const int* __restrict__ size,
                                      int* __restrict__ rank, int* __restrict__ dfs_order)
{
    constexpr int NB = 1 << DIM;
    const int lo = level_start[lev], hi = level_start[lev + 1];
    for (int b = lo + blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += gridDim.x * blockDim.x) {
        const int my = (b == 0) ? 0 : rank[b];
        if (b == 0) rank[0] = 0;
        dfs_order[my] = b;
        int r = my + 1;
#pragma unroll
        for (int m = NB - 1; m >= 0; --m) {
            const int c = child_ids[m * aligned + b];
            if (c > 0) { rank[c] = r; r += size[c]; }
        }
    }
}

// ---- c2: box masks, partition.py:124-357 ------------------------------------------------
__global__ void dist_mask_from_list_kernel(int n, const int* __restrict__ list, signed char* __restrict__ mask)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mask[list[i]] = 1;
}

// ancestors of the responsible boxes: the fixpoint of add_parent_boxes (partition.py:164-194)
// equals "every strict ancestor of a responsible box"
__global__ void dist_ancestor_mask_kernel(int nboxes, const signed char* __restrict__ responsible,
                                          const int* __restrict__ parent_ids, signed char* __restrict__ anc)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        if (!responsible[b]) continue;
        int cur = b;
        while (cur != 0) {
            cur = parent_ids[cur];
            if (anc[cur]) break;          // everything above is already marked (or being marked)
            anc[cur] = 1;
        }
    }
}

// add_interaction_list_boxes, partition.py:135-162 (mask_b, optional, is OR-ed in).  Rows with
// ~1e6 entries exist, so the work is cut by ENTRIES: every warp takes a span of kSpan consecutive
// list entries, finds the row of its first entry by one binary search in `starts` and then walks
// the rows of its span, 32 coalesced entries at a time.
constexpr int kMaskSpan = 2048;
__global__ void __launch_bounds__(256)
dist_add_list_boxes_kernel(int nrows, const int* __restrict__ box_list,
                           const signed char* __restrict__ mask_a, const signed char* __restrict__ mask_b,
                           const int* __restrict__ starts, const int* __restrict__ lists,
                           signed char* __restrict__ out_mask)
{
    const int nentries = starts[nrows];
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sp = w0 * kMaskSpan; sp < nentries; sp += nw * kMaskSpan) {
        const int first = (int)sp, last = (int)((sp + kMaskSpan < nentries) ? sp + kMaskSpan : nentries);
        int lo = 0, hi = nrows;                // last row with starts[row] <= first
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= first) lo = mid; else hi = mid; }
        int row = lo, k = first;
        while (k < last) {
            int rend = starts[row + 1];
            while (rend <= k) { ++row; rend = starts[row + 1]; }       // skip empty rows
            const int e = rend < last ? rend : last;
            const int box = box_list[row];
            if (mask_a[box] || (mask_b && mask_b[box]))
                for (int q = k + lane; q < e; q += 32) out_mask[lists[q]] = 1;
            k = e;
        }
    }
}

// every entry of a list marks its box (all rows of the list are rows the mask reads)
__global__ void __launch_bounds__(256)
dist_mark_list_boxes_kernel(int64_t n, const int* __restrict__ lists, signed char* __restrict__ out_mask)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t head = ((16 - ((uintptr_t)lists & 15)) & 15) / 4;      // entries before 16-byte alignment
    const int64_t h = head < n ? head : n;
    if (tid < h) out_mask[lists[tid]] = 1;
    const int4* v = reinterpret_cast<const int4*>(lists + h);
    const int64_t nv = (n - h) / 4;
    for (int64_t i = tid; i < nv; i += stride) {
        const int4 x = v[i];
        out_mask[x.x] = 1; out_mask[x.y] = 1; out_mask[x.z] = 1; out_mask[x.w] = 1;
    }
    const int64_t tail = h + nv * 4;
    if (tid < n - tail) out_mask[lists[tail + tid]] = 1;
}

// distributed build: target bits that the rank's local flags lack although the global flags
// have them, on the rows the masks read (responsible boxes and their ancestors) -- the rows of
// the global traversal that the local traversal does not contain (partition.py:197-297 reads
// them); source bits are kept.  any_out[0] != 0 if any such row exists.
__global__ void dist_corner_flags_kernel(int nboxes, const unsigned char* __restrict__ gflags,
                                         const unsigned char* __restrict__ lflags,
                                         const signed char* __restrict__ mask_a,
                                         const signed char* __restrict__ mask_b,
                                         unsigned char* __restrict__ out_flags, int* __restrict__ any_out)
{
    const unsigned char tbits = BT_BOX_IS_TARGET_BOX | BT_BOX_HAS_TARGET_CHILD_BOXES;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const unsigned char g = gflags[b], l = lflags[b];
        unsigned char keep = 0;
        if (mask_a[b] || mask_b[b]) keep = (unsigned char)(g & ~l & tbits);
        out_flags[b] = (unsigned char)((g & ~tbits) | keep);
        if (keep) *any_out = 1;
    }
}

// ---- c3: local particles, local_tree.py:70-151, 198-284 -------------------------------
__global__ void dist_particle_mask_kernel(int nboxes, const signed char* __restrict__ box_mask,
                                          const int* __restrict__ starts, const int* __restrict__ counts_nonchild,
                                          int* __restrict__ particle_mask)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int b = w; b < nboxes; b += nw) {
        if (!box_mask[b]) continue;
        const int s = starts[b], e = s + counts_nonchild[b];
        for (int p = s + lane; p < e; p += 32) particle_mask[p] = 1;
    }
}

struct MaskScanIn {
    const int* mask;
    __device__ int operator()(int64_t i) const { return mask[i]; }
};
struct MaskScanOut {   // mask_scan_kernel, local_tree.py:94-107: scan[i + 1] = inclusive sum
    int* g2l; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { g2l[i] = (int)excl; }
    __device__ void total(long long t) const { g2l[n] = (int)t; }
};

// fetch_local_particles, local_tree.py:124-151 (+ the global index of every local particle)
template <typename T>
__global__ void __launch_bounds__(256)
dist_fetch_kernel(int dim, int64_t n, const int* __restrict__ mask, const int* __restrict__ g2l,
                  const T* p0, const T* p1, const T* p2, const T* __restrict__ radii, T* o0, T* o1, T* o2,
                  T* __restrict__ oradii, long long* __restrict__ idx_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!mask[i]) continue;
        const int des = g2l[i];
        o0[des] = p0[i];
        if (dim > 1) o1[des] = p1[i];
        if (dim > 2) o2[des] = p2[i];
        if (radii) oradii[des] = radii[i];
        idx_out[des] = i;
    }
}

__global__ void dist_local_lists_kernel(int nboxes, const signed char* __restrict__ box_mask,
                                        const int* __restrict__ g2l, const int* __restrict__ starts,
                                        const int* __restrict__ counts_nonchild,
                                        const int* __restrict__ counts_cumul, int* __restrict__ lstarts,
                                        int* __restrict__ lnonchild, int* __restrict__ lcumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int s = starts[b];
        lstarts[b] = g2l[s];
        lnonchild[b] = box_mask[b] ? counts_nonchild[b] : 0;
        lcumul[b] = g2l[s + counts_cumul[b]] - g2l[s];
    }
}

// modify_target_flags, local_tree.py:163-185
__global__ void dist_modify_target_flags_kernel(int nboxes, const int* __restrict__ tgt_nonchild,
                                                const int* __restrict__ tgt_cumul,
                                                unsigned char* __restrict__ flags)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        unsigned char f = flags[b];
        f &= (unsigned char)~BT_BOX_IS_TARGET_BOX;
        f &= (unsigned char)~BT_BOX_HAS_TARGET_CHILD_BOXES;
        if (tgt_nonchild[b]) f |= BT_BOX_IS_TARGET_BOX;
        if (tgt_nonchild[b] < tgt_cumul[b]) f |= BT_BOX_HAS_TARGET_CHILD_BOXES;
        flags[b] = f;
    }
}

// flags with the target bits kept only for boxes of mask_a | mask_b (sharded setup: the
// rows of the partial traversal); need_mask = mask_a | mask_b
__global__ void dist_restrict_target_flags_kernel(int nboxes, const unsigned char* __restrict__ flags,
                                                  const signed char* __restrict__ mask_a,
                                                  const signed char* __restrict__ mask_b,
                                                  unsigned char* __restrict__ out_flags,
                                                  signed char* __restrict__ need_mask)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const bool need = mask_a[b] || mask_b[b];
        unsigned char f = flags[b];
        if (!need) f &= (unsigned char)~(BT_BOX_IS_TARGET_BOX | BT_BOX_HAS_TARGET_CHILD_BOXES);
        out_flags[b] = f;
        need_mask[b] = need ? 1 : 0;
    }
}

// MaskCompressorKernel, 2-D case (tools.py:647-740): masks[nranks][nboxes] -> per box the
// ascending list of ranks whose mask is set
struct RankCountIn {
    const signed char* masks; int nranks; int64_t nboxes;
    __device__ int operator()(int64_t b) const
    {
        int c = 0;
        for (int r = 0; r < nranks; ++r) c += masks[(int64_t)r * nboxes + b] ? 1 : 0;
        return c;
    }
};
struct RankCountOut {
    int* starts; int64_t n; long long* total_out;
    __device__ void operator()(int64_t b, long long excl) const { starts[b] = (int)excl; }
    __device__ void total(long long t) const { starts[n] = (int)t; *total_out = t; }
};
__global__ void dist_rank_fill_kernel(int nboxes, int nranks, const signed char* __restrict__ masks,
                                      const int* __restrict__ starts, int* __restrict__ lists)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        int k = starts[b];
        for (int r = 0; r < nranks; ++r) if (masks[(int64_t)r * nboxes + b]) lists[k++] = r;
    }
}

template <typename T>
static int fetch_impl(int dim, int64_t n, const int* mask, const int* g2l, void* const* parts,
                      const void* radii, void* const* outs, void* out_radii, long long* idx_out,
                      cudaStream_t s)
{
    if (n <= 0) return BT_OK;
    dist_fetch_kernel<T><<<grid_for(n, 256, 8), 256, 0, s>>>(
        dim, n, mask, g2l, (const T*)parts[0], dim > 1 ? (const T*)parts[1] : nullptr,
        dim > 2 ? (const T*)parts[2] : nullptr, (const T*)radii, (T*)outs[0],
        dim > 1 ? (T*)outs[1] : nullptr, dim > 2 ? (T*)outs[2] : nullptr, (T*)out_radii, idx_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}


// ---- (e): particle exchange of the distributed build ---------------------------------------
// After the distributed tree build every rank holds the GLOBAL box arrays and its OWN input
// particles in tree order.  A rank's local tree (local_tree.py:198-284) needs the sources of
// its point-source boxes and the targets of its responsible boxes, wherever they live: each
// owner packs, per destination rank, one record per needed particle
//     [coords (dim) | radius (optional) | box id (i32) | index inside the box's own range (i32)]
// the records travel in ONE all_to_all, and the receiver scatters them to
//     local_start[box] + index,
// which is the particle's place in the global tree order restricted to the rank's boxes (the
// order construct_local_particles_and_lists produces).

// dest_bits[b] = OR over ranks r with masks[r][b] != 0 of (1 << r)
__global__ void dist_mask_bits_kernel(int nboxes, int nranks, const signed char* __restrict__ masks,
                                      unsigned* __restrict__ bits)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        unsigned v = 0;
        for (int r = 0; r < nranks; ++r) if (masks[(int64_t)r * nboxes + b]) v |= 1u << r;
        bits[b] = v;
    }
}

// record offsets per (destination, box): exclusive scan, destination-major, of the own counts
// of the boxes whose bit is set
struct PackScanIn {
    const unsigned* dest_bits; const int* lown; int nboxes;
    __device__ int operator()(int64_t k) const
    {
        const int d = (int)(k / nboxes), b = (int)(k - (int64_t)d * nboxes);
        return ((dest_bits[b] >> d) & 1u) ? lown[b] : 0;
    }
};
struct PackScanOut {
    int* offs; long long* dest_offsets; /* [nranks + 1] */ int nboxes; int nranks;
    __device__ void operator()(int64_t k, long long excl) const
    {
        offs[k] = (int)excl;
        if (k % nboxes == 0) dest_offsets[k / nboxes] = excl;
    }
    __device__ void total(long long t) const { dest_offsets[nranks] = t; }
};

// 8 lanes per box copy the box's own particles into the record range of every destination
template <typename T>
__global__ void __launch_bounds__(256)
dist_pack_kernel(int nboxes, int nranks, int dim, int recbytes, const unsigned* __restrict__ dest_bits,
                 const int* __restrict__ lstart, const int* __restrict__ lown,
                 const int* __restrict__ rank_excl, const int* __restrict__ offs,
                 const T* c0, const T* c1, const T* c2, const T* __restrict__ radii,
                 unsigned char* __restrict__ sendbuf, long long cap)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    const int ng = (gridDim.x * blockDim.x) >> 3;
    for (int b = g; b < nboxes; b += ng) {
        unsigned bits = dest_bits[b];
        const int own = lown[b];
        if (!bits || !own) continue;
        const int s0 = lstart[b], rel0 = rank_excl[b];
        while (bits) {
            const int d = __ffs(bits) - 1;
            bits &= bits - 1;
            const long long base = offs[(int64_t)d * nboxes + b];
            for (int k = gl; k < own; k += 8) {
                if (base + k >= cap) break;
                unsigned char* rec = sendbuf + (base + k) * recbytes;
                T* c = reinterpret_cast<T*>(rec);
                const int p = s0 + k;
                c[0] = c0[p];
                if (dim > 1) c[1] = c1[p];
                if (dim > 2) c[2] = c2[p];
                int q = dim;
                if (radii) c[q++] = radii[p];
                int* tail = reinterpret_cast<int*>(c + q);
                tail[0] = b;
                tail[1] = rel0 + k;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
dist_unpack_kernel(int64_t nrec, int dim, int recbytes, int has_radii, const unsigned char* __restrict__ recv,
                   const int* __restrict__ dst_start, const int* __restrict__ gstart,
                   T* o0, T* o1, T* o2, T* __restrict__ oradii, long long* __restrict__ idx_out)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nrec;
         k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char* rec = recv + k * recbytes;
        const T* c = reinterpret_cast<const T*>(rec);
        int q = dim + (has_radii ? 1 : 0);
        const int* tail = reinterpret_cast<const int*>(c + q);
        const int b = tail[0], rel = tail[1];
        const int64_t pos = (int64_t)dst_start[b] + rel;
        o0[pos] = c[0];
        if (dim > 1) o1[pos] = c[1];
        if (dim > 2) o2[pos] = c[2];
        if (has_radii) oradii[pos] = c[dim];
        idx_out[pos] = (long long)gstart[b] + rel;
    }
}

// ranges of the rank's local particle arrays (local_tree.py:249-284) without a global particle
// array: the global tree order is the boxes' pre-order (own particles, then the children in
// Morton order), so the local start of a box is the prefix, in pre-order, of the own counts of
// the masked boxes before it
struct OwnScanIn {
    const signed char* mask; const int* own; const int* order;
    __device__ int operator()(int64_t i) const { const int b = order[i]; return mask[b] ? own[b] : 0; }
};
struct OwnScanOut {
    int* prefix; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { prefix[i] = (int)excl; }
    __device__ void total(long long t) const { prefix[n] = (int)t; }
};
__global__ void dist_local_ranges_kernel(int nboxes, const signed char* __restrict__ mask,
                                         const int* __restrict__ own, const int* __restrict__ rank,
                                         const int* __restrict__ subtree, const int* __restrict__ prefix,
                                         int* __restrict__ lstarts, int* __restrict__ lnonchild,
                                         int* __restrict__ lcumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int r = rank[b];
        lstarts[b] = prefix[r];
        lnonchild[b] = mask[b] ? own[b] : 0;
        lcumul[b] = prefix[r + subtree[b]] - prefix[r];
    }
}

template <typename T>
static int pack_impl(int nranks, int dim, int nboxes, const unsigned* dest_bits, void* const* parts,
                     const void* radii, const int* lstart, const int* lown, const int* rank_excl,
                     void* sendbuf, long long* dest_offsets, long long cap, cudaStream_t s)
{
    const int recbytes = (int)sizeof(T) * (dim + (radii ? 1 : 0)) + 8;
    int* offs = nullptr;
    BT_CHECK(temp_alloc((void**)&offs, sizeof(int) * (size_t)nranks * nboxes, s));
    PackScanIn in{dest_bits, lown, nboxes};
    PackScanOut out{offs, dest_offsets, nboxes, nranks};
    BT_TRY(scan_exclusive((int64_t)nranks * nboxes, nullptr, in, out, s));
    dist_pack_kernel<T><<<grid_for((int64_t)nboxes * 8, 256, 8), 256, 0, s>>>(
        nboxes, nranks, dim, recbytes, dest_bits, lstart, lown, rank_excl, offs, (const T*)parts[0],
        dim > 1 ? (const T*)parts[1] : nullptr, dim > 2 ? (const T*)parts[2] : nullptr,
        (const T*)radii, (unsigned char*)sendbuf, cap);
    BT_LAUNCH_CHECK();
    BT_CHECK(cudaFreeAsync(offs, s));
    return BT_OK;
}

}  // namespace bt

extern "C" {

int bt_dist_dfs_order(int dim, int nboxes, int aligned_nboxes, int nlevels,
                      const int32_t* level_start_box_nrs, const int32_t* box_child_ids,
                      int32_t* subtree_size, int32_t* rank_tmp, int32_t* dfs_order, void* stream)
{
    BT_PROF("bt_dist_dfs_order", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    const int grid = bt::grid_for(nboxes, 256, 4);
    for (int lev = nlevels - 1; lev >= 0; --lev) {
        if (dim == 1) bt::dist_subtree_size_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else if (dim == 2) bt::dist_subtree_size_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        else bt::dist_subtree_size_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size);
        BT_LAUNCH_CHECK();
    }
    for (int lev = 0; lev < nlevels; ++lev) {
        if (dim == 1) bt::dist_dfs_order_kernel<1><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        else if (dim == 2) bt::dist_dfs_order_kernel<2><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        else bt::dist_dfs_order_kernel<3><<<grid, 256, 0, s>>>(level_start_box_nrs, lev, aligned_nboxes, box_child_ids, subtree_size, rank_tmp, dfs_order);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

int bt_dist_mask_from_list(int n, const int32_t* list, int8_t* mask, void* stream)
{
    BT_PROF("bt_dist_mask_from_list", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::dist_mask_from_list_kernel<<<bt::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, list, (signed char*)mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_ancestor_mask(int nboxes, const int8_t* responsible, const int32_t* box_parent_ids,
                          int8_t* ancestor_mask, void* stream)
{
    BT_PROF("bt_dist_ancestor_mask", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_ancestor_mask_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)responsible, box_parent_ids, (signed char*)ancestor_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_add_list_boxes(int nrows, const int32_t* box_list, const int8_t* mask_a, const int8_t* mask_b,
                           const int32_t* starts, const int32_t* lists, int8_t* out_mask, void* stream)
{
    BT_PROF("bt_dist_add_list_boxes", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    bt::dist_add_list_boxes_kernel<<<bt::kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(
        nrows, box_list, (const signed char*)mask_a, (const signed char*)mask_b, starts, lists,
        (signed char*)out_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_particle_mask(int nboxes, const int8_t* box_mask, const int32_t* starts,
                          const int32_t* counts_nonchild, int32_t* particle_mask, void* stream)
{
    BT_PROF("bt_dist_particle_mask", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_particle_mask_kernel<<<bt::grid_for((int64_t)nboxes * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)box_mask, starts, counts_nonchild, particle_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mask_scan(int64_t n, const int32_t* particle_mask, int32_t* global_to_local, void* stream)
{
    BT_PROF("bt_dist_mask_scan", (cudaStream_t)stream);
    bt::MaskScanIn in{particle_mask};
    bt::MaskScanOut out{global_to_local, n};
    return bt::scan_exclusive(n, nullptr, in, out, (cudaStream_t)stream);
}

int bt_dist_fetch_local_particles(int dtype, int dim, int64_t n, const int32_t* particle_mask,
                                  const int32_t* global_to_local /*[n+1]*/, void* const* particles,
                                  const void* radii, void* const* local_particles, void* local_radii,
                                  int64_t* particle_idx, void* stream)
{
    BT_PROF("bt_dist_fetch_local_particles", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == BT_F32)
        return bt::fetch_impl<float>(dim, n, particle_mask, global_to_local, particles, radii,
                                     local_particles, local_radii, (long long*)particle_idx, s);
    if (dtype == BT_F64)
        return bt::fetch_impl<double>(dim, n, particle_mask, global_to_local, particles, radii,
                                      local_particles, local_radii, (long long*)particle_idx, s);
    return BT_ERR_BAD_ARG;
}

int bt_dist_local_lists(int nboxes, const int8_t* box_mask, const int32_t* global_to_local,
                        const int32_t* starts, const int32_t* counts_nonchild, const int32_t* counts_cumul,
                        int32_t* local_starts, int32_t* local_nonchild, int32_t* local_cumul, void* stream)
{
    BT_PROF("bt_dist_local_lists", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_local_lists_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, (const signed char*)box_mask, global_to_local, starts, counts_nonchild, counts_cumul,
        local_starts, local_nonchild, local_cumul);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_modify_target_flags(int nboxes, const int32_t* tgt_nonchild, const int32_t* tgt_cumul,
                                uint8_t* box_flags, void* stream)
{
    BT_PROF("bt_dist_modify_target_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_modify_target_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, tgt_nonchild, tgt_cumul, box_flags);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_restrict_target_flags(int nboxes, const uint8_t* box_flags, const int8_t* mask_a,
                                  const int8_t* mask_b, uint8_t* out_flags, int8_t* need_mask, void* stream)
{
    BT_PROF("bt_dist_restrict_target_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_restrict_target_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, box_flags, (const signed char*)mask_a, (const signed char*)mask_b, out_flags,
        (signed char*)need_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mark_list_boxes(int64_t nentries, const int32_t* lists, int8_t* out_mask, void* stream)
{
    BT_PROF("bt_dist_mark_list_boxes", (cudaStream_t)stream);
    if (nentries <= 0) return BT_OK;
    bt::dist_mark_list_boxes_kernel<<<bt::grid_for(nentries / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nentries, lists, (signed char*)out_mask);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_corner_flags(int nboxes, const uint8_t* global_flags, const uint8_t* local_flags,
                         const int8_t* mask_a, const int8_t* mask_b, uint8_t* out_flags, int32_t* any_out,
                         void* stream)
{
    BT_PROF("bt_dist_corner_flags", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    bt::dist_corner_flags_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, global_flags, local_flags, (const signed char*)mask_a, (const signed char*)mask_b,
        out_flags, any_out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_box_to_user_rank(int phase, int nboxes, int nranks, const int8_t* masks_all_ranks,
                             int32_t* starts, int32_t* lists, int64_t* total_dev, void* stream)
{
    BT_PROF("bt_dist_box_to_user_rank", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::RankCountIn in{(const signed char*)masks_all_ranks, nranks, nboxes};
        bt::RankCountOut out{starts, nboxes, (long long*)total_dev};
        return bt::scan_exclusive(nboxes, nullptr, in, out, s);
    }
    if (nboxes <= 0) return BT_OK;
    bt::dist_rank_fill_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
        nboxes, nranks, (const signed char*)masks_all_ranks, starts, lists);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_mask_bits(int nboxes, int nranks, const int8_t* masks_all_ranks, uint32_t* dest_bits, void* stream)
{
    BT_PROF("bt_dist_mask_bits", (cudaStream_t)stream);
    if (nboxes <= 0) return BT_OK;
    if (nranks > 32) return BT_ERR_UNSUPPORTED;
    bt::dist_mask_bits_kernel<<<bt::grid_for(nboxes, 256), 256, 0, (cudaStream_t)stream>>>(
        nboxes, nranks, (const signed char*)masks_all_ranks, dest_bits);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_pack_records(int dtype, int nranks, int dim, int nboxes, const uint32_t* dest_bits,
                         void* const* particles, const void* radii, const int32_t* local_start,
                         const int32_t* local_own, const int32_t* rank_excl, void* sendbuf,
                         int64_t* dest_offsets, void* stream, int64_t capacity)
{
    BT_PROF("bt_dist_pack_records", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nranks > 32) return BT_ERR_UNSUPPORTED;
    if (nboxes <= 0) return (int)cudaMemsetAsync(dest_offsets, 0, sizeof(int64_t) * (nranks + 1), s);
    if (dtype == BT_F32)
        return bt::pack_impl<float>(nranks, dim, nboxes, dest_bits, particles, radii, local_start,
                                    local_own, rank_excl, sendbuf, (long long*)dest_offsets,
                                    (long long)capacity, s);
    if (dtype == BT_F64)
        return bt::pack_impl<double>(nranks, dim, nboxes, dest_bits, particles, radii, local_start,
                                     local_own, rank_excl, sendbuf, (long long*)dest_offsets,
                                     (long long)capacity, s);
    return BT_ERR_BAD_ARG;
}

int bt_dist_unpack_records(int dtype, int dim, int64_t nrec, int has_radii, const void* recvbuf,
                           const int32_t* dst_start, const int32_t* box_global_start,
                           void* const* local_particles, void* local_radii, int64_t* particle_idx,
                           void* stream)
{
    BT_PROF("bt_dist_unpack_records", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nrec <= 0) return BT_OK;
    const int grid = bt::grid_for(nrec, 256, 8);
    if (dtype == BT_F32) {
        const int rb = 4 * (dim + (has_radii ? 1 : 0)) + 8;
        bt::dist_unpack_kernel<float><<<grid, 256, 0, s>>>(
            nrec, dim, rb, has_radii, (const unsigned char*)recvbuf, dst_start, box_global_start,
            (float*)local_particles[0], dim > 1 ? (float*)local_particles[1] : nullptr,
            dim > 2 ? (float*)local_particles[2] : nullptr, (float*)local_radii, (long long*)particle_idx);
    } else if (dtype == BT_F64) {
        const int rb = 8 * (dim + (has_radii ? 1 : 0)) + 8;
        bt::dist_unpack_kernel<double><<<grid, 256, 0, s>>>(
            nrec, dim, rb, has_radii, (const unsigned char*)recvbuf, dst_start, box_global_start,
            (double*)local_particles[0], dim > 1 ? (double*)local_particles[1] : nullptr,
            dim > 2 ? (double*)local_particles[2] : nullptr, (double*)local_radii, (long long*)particle_idx);
    } else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_dist_local_ranges(int nboxes, const int8_t* box_mask, const int32_t* own_counts,
                         const int32_t* preorder_rank, const int32_t* preorder_boxes,
                         const int32_t* subtree_size, int32_t* prefix_tmp /*[nboxes+1]*/,
                         int32_t* local_starts, int32_t* local_nonchild, int32_t* local_cumul, void* stream)
{
    BT_PROF("bt_dist_local_ranges", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes <= 0) return BT_OK;
    bt::OwnScanIn in{(const signed char*)box_mask, own_counts, preorder_boxes};
    bt::OwnScanOut out{prefix_tmp, nboxes};
    BT_TRY(bt::scan_exclusive(nboxes, nullptr, in, out, s));
    bt::dist_local_ranges_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
        nboxes, (const signed char*)box_mask, own_counts, preorder_rank, subtree_size, prefix_tmp,
        local_starts, local_nonchild, local_cumul);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // extern "C"
