// Kernels of the rows that surround the hot path (N1, N2, N4):
// the constant-one FMM (boxtree/fmm.py:342-532 with boxtree/constant_one.py:50-237), the
// particle-list filters and point-source linking (boxtree/tree.py:772-1239 with the kernel
// templates of tree_build_kernels.py:1872-2021), the translation-class finder
// (boxtree/translation_classes.py:60-196) and the segmented sums of the cost model
// (boxtree/cost.py:445-525, 715-1262).  All of them consume Tree / FMMTraversalInfo arrays.
#include "common.cuh"
#include "scan.cuh"
#include "../../include/boxtree_b200.h"

namespace bt {

// ---- CSR row sums: out[i] (+)= sum_k values[lists[k]], k in [starts[i], starts[i+1]) -----------
// one warp per row; rows are a few hundred entries (list 2) up to ~1e6 (close lists of
// upper-level boxes), so long rows are additionally split over the warps of a second launch
template <typename V>
__global__ void __launch_bounds__(256)
csr_row_sums_kernel(int nrows, const int* __restrict__ starts, const int* __restrict__ lists,
                    const V* __restrict__ values, const int* __restrict__ out_index, V* __restrict__ out,
                    int accumulate, V scale)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = w; r < nrows; r += nw) {
        const int s = starts[r], e = starts[r + 1];
        V acc = 0;
        for (int k = s + lane; k < e; k += 32) acc += values[lists[k]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const int dst = out_index ? out_index[r] : r;
            out[dst] = (accumulate ? out[dst] : (V)0) + acc * scale;
        }
    }
}

// out[b] = sum of values[starts[b] .. starts[b] + counts[b])
template <typename V>
__global__ void __launch_bounds__(256)
range_sums_kernel(int n, const int* __restrict__ starts, const int* __restrict__ counts,
                  const V* __restrict__ values, V* __restrict__ out)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    const int ng = (gridDim.x * blockDim.x) >> 3;
    const unsigned gm = 0xffu << ((threadIdx.x & 31) - gl);
    for (int b = g; b < n; b += ng) {
        const int s = starts[b], e = s + counts[b];
        V acc = 0;
        for (int k = s + gl; k < e; k += 8) acc += values[k];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(gm, acc, o);
        if (gl == 0) out[b] = acc;
    }
}

// pot[j] += vals[i] for every j in the own range of boxes[i] (ranges of distinct boxes are disjoint)
template <typename V>
__global__ void __launch_bounds__(256)
add_to_ranges_kernel(int nrows, const int* __restrict__ boxes, const V* __restrict__ vals,
                     const int* __restrict__ starts, const int* __restrict__ counts, V* __restrict__ pot)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    const int ng = (gridDim.x * blockDim.x) >> 3;
    for (int i = g; i < nrows; i += ng) {
        const int b = boxes ? boxes[i] : i;
        const V v = vals[i];
        const int s = starts[b], e = s + counts[b];
        for (int k = s + gl; k < e; k += 8) pot[k] += v;
    }
}

// constant-one multipole-to-multipole: mpoles[b] += sum over children (constant_one.py:118-152)
__global__ void fmm_upward_kernel(int nb_children, int nrows, const int* __restrict__ boxes,
                                  const int* __restrict__ child_ids, int aligned, long long* __restrict__ mpoles)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
        const int b = boxes[i];
        long long acc = 0;
        for (int m = 0; m < nb_children; ++m) {
            const int c = child_ids[(int64_t)m * aligned + b];
            if (c) acc += mpoles[c];
        }
        mpoles[b] += acc;
    }
}
// local-to-local: local[b] += local[parent[b]] (constant_one.py:208-224)
__global__ void fmm_downward_kernel(int nrows, const int* __restrict__ boxes, const int* __restrict__ parents,
                                    long long* __restrict__ local)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
        const int b = boxes[i];
        local[b] += local[parents[b]];
    }
}

template <typename V>
__global__ void gather_kernel(int64_t n, const V* __restrict__ src, const int* __restrict__ idx, V* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = src[idx[i]];
}
template <typename V>
__global__ void widen_i32_kernel(int64_t n, const int* __restrict__ src, V* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (V)src[i];
}

// ---- N2: particle list filters (tree.py:1057-1239) ---------------------------------------------
// user order (tree.py:1097-1130, ListOfListsBuilder over boxes): row b = the user ids of b's own
// targets whose flag is set, in tree order
__global__ void __launch_bounds__(256)
filter_user_order_kernel(int nboxes, int fill, const int* __restrict__ tstart, const int* __restrict__ tcount,
                         const int* __restrict__ user_target_ids, const signed char* __restrict__ flags_user,
                         int* __restrict__ starts, int* __restrict__ lists)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int b = w; b < nboxes; b += nw) {
        const int s = tstart[b], n = tcount[b];
        int run = fill ? starts[b] : 0;
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            int uid = 0;
            bool keep = false;
            if (k < n) { uid = user_target_ids[s + k]; keep = flags_user[uid] != 0; }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (fill && keep) lists[run + __popc(m & ((1u << lane) - 1u))] = uid;
            run += __popc(m);
        }
        if (!fill && lane == 0) starts[b] = run;
    }
}

struct CountsIn { const int* a; __device__ int operator()(int64_t i) const { return a[i]; } };
struct CountsOut {
    int* a; long long* t; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { a[i] = (int)excl; }
    __device__ void total(long long v) const { a[n] = (int)v; *t = v; }
};

// tree order (TREE_ORDER_TARGET_FILTER_SCAN_TPL / _INDEX_TPL, tree_build_kernels.py:1954-2021)
struct FlagScanIn {
    const signed char* flags_user; const int* user_target_ids;
    __device__ int operator()(int64_t i) const { return flags_user[user_target_ids[i]] != 0 ? 1 : 0; }
};
struct FlagScanOut {
    const signed char* flags_user; const int* user_target_ids;
    int* filtered_from_unfiltered; int* unfiltered_from_filtered; int64_t n; int* total_out;
    __device__ void operator()(int64_t i, long long excl) const
    {
        filtered_from_unfiltered[i] = (int)excl;
        if (flags_user[user_target_ids[i]] != 0) unfiltered_from_filtered[excl] = (int)i;
    }
    __device__ void total(long long t) const { filtered_from_unfiltered[n] = (int)t; *total_out = (int)t; }
};
__global__ void filter_box_ranges_kernel(int nboxes, int ntargets, const int* __restrict__ tstart,
                                         const int* __restrict__ tcount, const int* __restrict__ ffu,
                                         int* __restrict__ fstart, int* __restrict__ fcount)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int s = tstart[b], c = tcount[b];
        const int a = ffu[s < ntargets ? s : ntargets];
        const int e = ffu[(s + c) < ntargets ? (s + c) : ntargets];
        fstart[b] = a;
        fcount[b] = c > 0 ? e - a : 0;
    }
}

// ---- N2: point sources (POINT_SOURCE_LINKING_*, tree_build_kernels.py:1872-1950) ---------------
struct PsScanIn {
    const int* pss_user; const int* user_source_ids;
    __device__ int operator()(int64_t i) const { const int u = user_source_ids[i]; return pss_user[u + 1] - pss_user[u]; }
};
struct PsScanOut {
    const int* pss_user; const int* user_source_ids; int* tree_starts; int* counts; int64_t n; int* total_out;
    __device__ void operator()(int64_t i, long long excl) const
    {
        const int u = user_source_ids[i];
        tree_starts[i] = (int)excl; counts[i] = pss_user[u + 1] - pss_user[u];
    }
    __device__ void total(long long t) const { tree_starts[n] = (int)t; *total_out = (int)t; }
};
// ids of the point sources of tree-order source i: its user-order range, ascending
__global__ void __launch_bounds__(256)
ps_ids_kernel(int nsources, const int* __restrict__ pss_user, const int* __restrict__ user_source_ids,
              const int* __restrict__ tree_starts, int* __restrict__ user_point_source_ids)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    const int ng = (gridDim.x * blockDim.x) >> 3;
    for (int i = g; i < nsources; i += ng) {
        const int u = user_source_ids[i];
        const int first = pss_user[u], n = pss_user[u + 1] - first, dst = tree_starts[i];
        for (int k = gl; k < n; k += 8) user_point_source_ids[dst + k] = first + k;
    }
}
__global__ void ps_box_ranges_kernel(int nboxes, int nsources, const int* __restrict__ sstart,
                                     const int* __restrict__ snonchild, const int* __restrict__ scumul,
                                     const int* __restrict__ tree_starts /*[nsources+1]*/,
                                     int* __restrict__ ps_start, int* __restrict__ ps_nonchild,
                                     int* __restrict__ ps_cumul)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += gridDim.x * blockDim.x) {
        const int s = sstart[b];
        const int a = tree_starts[s < nsources ? s : nsources];
        ps_start[b] = a;
        const int c1 = snonchild[b], c2 = scumul[b];
        ps_nonchild[b] = c1 > 0 ? tree_starts[s + c1] - a : 0;
        ps_cumul[b] = c2 > 0 ? tree_starts[s + c2] - a : 0;
    }
}

// ---- N4: translation classes (translation_classes.py:60-196) -----------------------------------
// one thread per list-2 ENTRY (its row by binary search in starts); error flag instead of an
// exception; used[class] = 1
template <typename T>
__global__ void __launch_bounds__(256)
translation_classes_kernel(int dim, int nrows, const int* __restrict__ row_boxes, const int* __restrict__ starts,
                           const int* __restrict__ lists, const T* __restrict__ centers, int aligned,
                           const unsigned char* __restrict__ levels, T root_extent, int n_away, int per_level,
                           int nper, int* __restrict__ classes, int* __restrict__ used, int* __restrict__ error)
{
    const int npairs = starts[nrows];
    const int bound = 2 * n_away + 1, base = 4 * n_away + 3;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < npairs; k += gridDim.x * blockDim.x) {
        int lo = 0, hi = nrows;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= k) lo = mid; else hi = mid; }
        const int tgt = row_boxes[lo], src = lists[k];
        const int lev = levels[src];
        bool bad = lev != levels[tgt];
        // LEVEL_TO_RAD(level) = root_extent * 1 / (coord_t)(1 << (level + 1)); diameter = 2 * rad
        const T diam = 2 * (root_extent * 1 / (T)(1 << (lev + 1)));
        int cls = 0, mult = 1;
        for (int a = 0; a < dim; ++a) {
            const T q = (centers[(int64_t)a * aligned + tgt] - centers[(int64_t)a * aligned + src]) / diam;
            const int vec = (int)rint((double)q);
            bad = bad || vec < -bound || vec > bound;
            cls += (bound + vec) * mult;
            mult *= base;
        }
        if (per_level) cls += lev * nper;
        if (bad) { *error = 1; cls = 0; }
        classes[k] = cls;
        used[cls] = 1;
    }
}
__global__ void remap_classes_kernel(int64_t n, const int* __restrict__ used_map, int* __restrict__ classes)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        classes[i] = used_map[classes[i]];
}

}  // namespace bt

extern "C" {

int bt_csr_row_sums(int value_kind, int nrows, const int32_t* starts, const int32_t* lists, const void* values,
                    const int32_t* out_index, void* out, int accumulate, double scale, void* stream)
{
    BT_PROF("bt_csr_row_sums", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nrows <= 0) return BT_OK;
    const int grid = bt::grid_for((int64_t)nrows * 32, 256, 8);
    if (value_kind == 0)
        bt::csr_row_sums_kernel<long long><<<grid, 256, 0, s>>>(nrows, starts, lists, (const long long*)values,
                                                                out_index, (long long*)out, accumulate,
                                                                (long long)scale);
    else if (value_kind == 1)
        bt::csr_row_sums_kernel<double><<<grid, 256, 0, s>>>(nrows, starts, lists, (const double*)values,
                                                             out_index, (double*)out, accumulate, scale);
    else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_range_sums_i64(int n, const int32_t* starts, const int32_t* counts, const int64_t* values, int64_t* out,
                      void* stream)
{
    BT_PROF("bt_range_sums", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::range_sums_kernel<long long><<<bt::grid_for((int64_t)n * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        n, starts, counts, (const long long*)values, (long long*)out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_add_to_ranges_i64(int nrows, const int32_t* boxes, const int64_t* vals, const int32_t* starts,
                         const int32_t* counts, int64_t* pot, void* stream)
{
    BT_PROF("bt_add_to_ranges", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    bt::add_to_ranges_kernel<long long><<<bt::grid_for((int64_t)nrows * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nrows, boxes, (const long long*)vals, starts, counts, (long long*)pot);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_fmm_upward_i64(int dim, int nrows, const int32_t* boxes, const int32_t* box_child_ids, int aligned_nboxes,
                      int64_t* mpoles, void* stream)
{
    BT_PROF("bt_fmm_upward", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    bt::fmm_upward_kernel<<<bt::grid_for(nrows, 256), 256, 0, (cudaStream_t)stream>>>(
        1 << dim, nrows, boxes, box_child_ids, aligned_nboxes, (long long*)mpoles);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_fmm_downward_i64(int nrows, const int32_t* boxes, const int32_t* box_parent_ids, int64_t* local,
                        void* stream)
{
    BT_PROF("bt_fmm_downward", (cudaStream_t)stream);
    if (nrows <= 0) return BT_OK;
    bt::fmm_downward_kernel<<<bt::grid_for(nrows, 256), 256, 0, (cudaStream_t)stream>>>(
        nrows, boxes, box_parent_ids, (long long*)local);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_gather_i64(int64_t n, const int64_t* src, const int32_t* idx, int64_t* out, void* stream)
{
    BT_PROF("bt_gather_i64", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::gather_kernel<long long><<<bt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        n, (const long long*)src, idx, (long long*)out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_gather_coords(int dtype, int64_t n, const void* src, const int32_t* idx, void* out, void* stream)
{
    BT_PROF("bt_gather_coords", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    const int grid = bt::grid_for(n, 256, 8);
    if (dtype == BT_F32)
        bt::gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(n, (const float*)src, idx, (float*)out);
    else if (dtype == BT_F64)
        bt::gather_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(n, (const double*)src, idx, (double*)out);
    else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_widen_i32(int value_kind, int64_t n, const int32_t* src, void* out, void* stream)
{
    BT_PROF("bt_widen_i32", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    const int grid = bt::grid_for(n, 256, 8);
    if (value_kind == 0)
        bt::widen_i32_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>(n, src, (long long*)out);
    else if (value_kind == 1)
        bt::widen_i32_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(n, src, (double*)out);
    else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_filter_targets_user_order(int phase, int nboxes, const int32_t* box_target_starts,
                                 const int32_t* box_target_counts_nonchild, const int32_t* user_target_ids,
                                 const int8_t* flags_user, int32_t* starts, int32_t* lists,
                                 int64_t* total_dev, void* stream)
{
    BT_PROF("bt_filter_targets_user_order", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nboxes > 0) {
        bt::filter_user_order_kernel<<<bt::grid_for((int64_t)nboxes * 32, 256, 8), 256, 0, s>>>(
            nboxes, phase, box_target_starts, box_target_counts_nonchild, user_target_ids,
            (const signed char*)flags_user, starts, lists);
        BT_LAUNCH_CHECK();
    }
    if (phase == 0) {       // counts -> starts
        bt::CountsIn in{starts};
        bt::CountsOut out{starts, (long long*)total_dev, nboxes};
        return bt::scan_exclusive(nboxes, nullptr, in, out, s);
    }
    return BT_OK;
}

int bt_filter_targets_tree_order(int nboxes, int64_t ntargets, const int32_t* box_target_starts,
                                 const int32_t* box_target_counts_nonchild, const int32_t* user_target_ids,
                                 const int8_t* flags_user, int32_t* filtered_from_unfiltered /*[n+1]*/,
                                 int32_t* unfiltered_from_filtered /*[n]*/, int32_t* nfiltered_dev,
                                 int32_t* filtered_starts, int32_t* filtered_counts, void* stream)
{
    BT_PROF("bt_filter_targets_tree_order", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    bt::FlagScanIn in{(const signed char*)flags_user, user_target_ids};
    bt::FlagScanOut out{(const signed char*)flags_user, user_target_ids, filtered_from_unfiltered,
                        unfiltered_from_filtered, ntargets, nfiltered_dev};
    BT_TRY(bt::scan_exclusive(ntargets, nullptr, in, out, s));
    if (nboxes > 0) {
        bt::filter_box_ranges_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
            nboxes, (int)ntargets, box_target_starts, box_target_counts_nonchild, filtered_from_unfiltered,
            filtered_starts, filtered_counts);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

int bt_link_point_sources(int phase, int nboxes, int64_t nsources, const int32_t* point_source_starts_user,
                          const int32_t* user_source_ids, int32_t* tree_order_starts /*[nsources+1]*/,
                          int32_t* point_source_counts, int32_t* npoint_sources_dev,
                          int32_t* user_point_source_ids, const int32_t* box_source_starts,
                          const int32_t* box_source_counts_nonchild, const int32_t* box_source_counts_cumul,
                          int32_t* box_ps_starts, int32_t* box_ps_nonchild, int32_t* box_ps_cumul, void* stream)
{
    BT_PROF("bt_link_point_sources", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::PsScanIn in{point_source_starts_user, user_source_ids};
        bt::PsScanOut out{point_source_starts_user, user_source_ids, tree_order_starts, point_source_counts,
                          nsources, npoint_sources_dev};
        return bt::scan_exclusive(nsources, nullptr, in, out, s);
    }
    if (nsources > 0) {
        bt::ps_ids_kernel<<<bt::grid_for(nsources * 8, 256, 8), 256, 0, s>>>(
            (int)nsources, point_source_starts_user, user_source_ids, tree_order_starts, user_point_source_ids);
        BT_LAUNCH_CHECK();
    }
    if (nboxes > 0) {
        bt::ps_box_ranges_kernel<<<bt::grid_for(nboxes, 256), 256, 0, s>>>(
            nboxes, (int)nsources, box_source_starts, box_source_counts_nonchild, box_source_counts_cumul,
            tree_order_starts, box_ps_starts, box_ps_nonchild, box_ps_cumul);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

int bt_translation_classes(int dtype, int dim, int nrows, const int32_t* row_boxes, const int32_t* starts,
                           const int32_t* lists, const void* box_centers, int aligned_nboxes,
                           const uint8_t* box_levels, double root_extent, int well_sep_is_n_away,
                           int per_level, int nclasses_per_level, int64_t npairs, int32_t* classes,
                           int32_t* used, int32_t* error_dev, void* stream)
{
    BT_PROF("bt_translation_classes", (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (nrows <= 0 || npairs <= 0) return BT_OK;
    const int grid = bt::grid_for(npairs, 256, 8);
    if (dtype == BT_F32)
        bt::translation_classes_kernel<float><<<grid, 256, 0, s>>>(
            dim, nrows, row_boxes, starts, lists, (const float*)box_centers, aligned_nboxes, box_levels,
            (float)root_extent, well_sep_is_n_away, per_level, nclasses_per_level, classes, used, error_dev);
    else if (dtype == BT_F64)
        bt::translation_classes_kernel<double><<<grid, 256, 0, s>>>(
            dim, nrows, row_boxes, starts, lists, (const double*)box_centers, aligned_nboxes, box_levels,
            root_extent, well_sep_is_n_away, per_level, nclasses_per_level, classes, used, error_dev);
    else return BT_ERR_BAD_ARG;
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_remap_classes(int64_t n, const int32_t* used_map, int32_t* classes, void* stream)
{
    BT_PROF("bt_remap_classes", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt::remap_classes_kernel<<<bt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, used_map, classes);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

}  // extern "C"
