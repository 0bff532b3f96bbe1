// FMM traversal list builders for sm_100a.
//
// Every builder follows the reference's ListOfListsBuilder protocol (count
// pass, exclusive scan to `starts`, write pass; row content in the APPEND order
// of the row's depth-first, Morton-ordered walk) so the CSR arrays are
// identical to the reference's.  The walks restate boxtree/traversal.py:98-160
// (walk_init / walk_advance / walk_push) and the per-list generate() bodies
// cited at each kernel.  List 3 is produced for all source levels by ONE walk
// that bins appends by the level of the appended box (the reference repeats the
// walk nlevels(+1) times, traversal.py:2203-2228).
#include "common.cuh"
#include "scan.cuh"
#include "../../include/boxtree_b200.h"

namespace bt {

constexpr int kMaxWalkLevels = 40;
constexpr int kTravBlock = 128;

template <typename T, int DIM>
struct TreeView {
    const T* centers; const unsigned char* levels; const int* child_ids;
    const unsigned char* flags; const int* parents;
    int aligned; int nboxes; int nlevels; T root_extent; int n_away;
    __device__ __forceinline__ void center(int b, T* c) const
    {
#pragma unroll
        for (int a = 0; a < DIM; ++a) c[a] = centers[aligned * a + b];
    }
    __device__ __forceinline__ int child(int parent, int mnr) const
    { return child_ids[mnr * aligned + parent]; }
};

template <typename T, int DIM>
static TreeView<T, DIM> make_view(const bt_tree_view* v)
{
    TreeView<T, DIM> t;
    t.centers = (const T*)v->box_centers; t.levels = v->box_levels; t.child_ids = v->box_child_ids;
    t.flags = v->box_flags; t.parents = v->box_parent_ids; t.aligned = v->aligned_nboxes;
    t.nboxes = v->nboxes; t.nlevels = v->nlevels; t.root_extent = (T)v->root_extent;
    t.n_away = v->well_sep_is_n_away;
    return t;
}

// LEVEL_TO_RAD(level) = root_extent * 1 / (coord_t)(1 << (level + 1)), traversal.py:234-235
template <typename T>
__device__ __forceinline__ void fill_rad_table(T* rad, T root_extent)
{
    for (int l = threadIdx.x; l < kMaxWalkLevels; l += blockDim.x)
        rad[l] = (root_extent * 1 / (T)(1 << ((l + 1) & 31)));
    __syncthreads();
}

// is_adjacent_or_overlapping_with_neighborhood, traversal.py:279-305
template <typename T, int DIM>
__device__ __forceinline__ bool adj_nbhd(const T* rad, const T* tc, int tl, T nbhd, const T* sc, int sl)
{
    const T target_rad = rad[tl], source_rad = rad[sl];
    const T rad_sum = ((2 * (nbhd - 1) + 1) * target_rad + source_rad);
    const T slack = rad_sum + fmin(target_rad, source_rad);
    T l_inf_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) l_inf_dist = fmax(l_inf_dist, fabs(tc[a] - sc[a]));
    return l_inf_dist <= slack;
}

struct Walk {
    int box_stack[kMaxWalkLevels]; signed char mnr_stack[kMaxWalkLevels];
    int ssize, parent, mnr; bool cont;
    __device__ __forceinline__ void init(int start) { ssize = 0; parent = start; mnr = 0; cont = true; }
    __device__ __forceinline__ void push(int nb)
    { box_stack[ssize] = parent; mnr_stack[ssize] = (signed char)mnr; ++ssize; parent = nb; mnr = 0; }
    template <int NB> __device__ __forceinline__ void advance()
    {
        while (true) {
            ++mnr;
            if (mnr < NB) break;
            cont = (ssize > 0);
            if (cont) { --ssize; parent = box_stack[ssize]; mnr = mnr_stack[ssize]; }
            else break;
        }
    }
};

// emitters
struct CountEmit {
    int c0 = 0, c1 = 0;
    __device__ __forceinline__ void e0(int) { ++c0; }
    __device__ __forceinline__ void e1(int) { ++c1; }
};
struct FillEmit {
    int* p0; int* p1;
    __device__ __forceinline__ void e0(int v) { *p0++ = v; }
    __device__ __forceinline__ void e1(int v) { if (p1) *p1++ = v; }
};

// ---- b3 colleagues: traversal.py:398-464 -----------------------------------
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_colleagues(const TreeView<T, DIM>& t, const T* rad, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    if (box_id == 0) return;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    const T nbhd = (T)t.n_away;
    Walk w; w.init(0);
    while (w.cont) {
        const int wb = t.child(w.parent, w.mnr);
        if (wb) {
            T wc[DIM]; t.center(wb, wc);
            // The reference also descends into box_id itself (traversal.py:438-452) and
            // walks its whole subtree, which can never append (deeper levels only):
            // skipping that descent leaves the output unchanged and removes a serial
            // walk over up to nboxes/2^d boxes by a single thread.
            if (wb != box_id && adj_nbhd<T, DIM>(rad, center, level, nbhd, wc, t.levels[wb])) {
                if (w.ssize + 1 == level) e.e0(wb);
                else { w.push(wb); continue; }
            }
        }
        w.template advance<NB>();
    }
}

// ---- b4 list 1: traversal.py:470-550 ---------------------------------------
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list1(const TreeView<T, DIM>& t, const T* rad, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    if (t.flags[0] & BT_BOX_IS_SOURCE_BOX) e.e0(0);
    Walk w; w.init(0);
    while (w.cont) {
        const int wb = t.child(w.parent, w.mnr);
        if (wb) {
            T wc[DIM]; t.center(wb, wc);
            if (adj_nbhd<T, DIM>(rad, center, level, (T)1, wc, t.levels[wb])) {
                const unsigned char fl = t.flags[wb];
                if (fl & BT_BOX_IS_SOURCE_BOX) e.e0(wb);
                if (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES) { w.push(wb); continue; }
            }
        }
        w.template advance<NB>();
    }
}

// ---- b5 list 2: traversal.py:556-601 ---------------------------------------
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list2(const TreeView<T, DIM>& t, const T* rad, const int* coll_starts,
                                          const int* coll_lists, int box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    T center[DIM]; t.center(box_id, center);
    const int level = t.levels[box_id];
    const int parent = t.parents[box_id];
    if (parent == box_id) return;
    const T nbhd = (T)t.n_away;
    const int s = coll_starts[parent], en = coll_starts[parent + 1];
    for (int i = s; i < en; ++i) {
        const int parent_nf = coll_lists[i];
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const int sib = t.child(parent_nf, m);
            if (sib == 0) continue;
            T sc[DIM]; t.center(sib, sc);
            if (!adj_nbhd<T, DIM>(rad, center, level, nbhd, sc, t.levels[sib])) e.e0(sib);
        }
    }
}

// ---- b7 list 4 (+close): traversal.py:931-1146 -----------------------------
template <typename T, int DIM>
__device__ __forceinline__ bool meets_sep_bigger(const T* rad, const T* tc, int tl, const T* sc, int sl,
                                                 T stick_out_factor)
{   // traversal.py:933-972
    const T target_rad = rad[tl], source_rad = rad[sl];
    const T max_allowed = (3 * (1 + stick_out_factor) * target_rad + source_rad);
    T l_inf_dist = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) l_inf_dist = fmax(l_inf_dist, fabs(tc[a] - sc[a]));
    return l_inf_dist >= max_allowed * (1 - 8 * CoordTraits<T>::eps());
}

template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list4(const TreeView<T, DIM>& t, const T* rad, const int* coll_starts,
                                          const int* coll_lists, int with_extent, T stick_out_factor,
                                          int tgt_ibox, E& e)
{
    T tc[DIM]; t.center(tgt_ibox, tc);
    const int tgt_level = t.levels[tgt_ibox];
    if (tgt_level == 0) return;
    const int tgt_parent = t.parents[tgt_ibox];
    const int tgt_parent_level = tgt_level - 1;
    T pc[DIM]; t.center(tgt_parent, pc);
    const unsigned char tgt_flags = t.flags[tgt_ibox];
    int walk_level, cur;
    if (t.n_away == 1) { walk_level = tgt_level - 1; cur = tgt_parent; }
    else { walk_level = tgt_level; cur = tgt_ibox; }
    for (; walk_level != 0; --walk_level, cur = t.parents[cur]) {
        const int s = coll_starts[cur], en = coll_starts[cur + 1];
        for (int i = s; i < en; ++i) {
            const int sb = coll_lists[i];
            if (!(t.flags[sb] & BT_BOX_IS_SOURCE_BOX)) continue;
            T sc[DIM]; t.center(sb, sc);
            if (adj_nbhd<T, DIM>(rad, tc, tgt_level, (T)1, sc, walk_level)) continue;
            if (with_extent) {
                if (!meets_sep_bigger<T, DIM>(rad, tc, tgt_level, sc, walk_level, stick_out_factor)) {
                    if (tgt_flags & BT_BOX_IS_TARGET_BOX) e.e1(sb);
                    continue;
                }
            }
            const bool in_parent_list_1 = adj_nbhd<T, DIM>(rad, pc, tgt_parent_level, (T)1, sc, walk_level);
            bool would_be_in_parent_list_4 = !in_parent_list_1;
            if (t.n_away > 1) would_be_in_parent_list_4 = would_be_in_parent_list_4 && (walk_level < tgt_level);
            if (would_be_in_parent_list_4) {
                if (with_extent &&
                    !meets_sep_bigger<T, DIM>(rad, pc, tgt_parent_level, sc, walk_level, stick_out_factor))
                    e.e0(sb);
            } else e.e0(sb);
        }
    }
}

// generic row kernel; KIND: 0 colleagues, 1 list1, 2 list2, 4 list4
template <typename T, int DIM, int KIND, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list_kernel(TreeView<T, DIM> t, const int* __restrict__ row_boxes, const int* __restrict__ coll_starts,
            const int* __restrict__ coll_lists, int with_extent, T stick_out_factor, int nrows,
            int* __restrict__ starts, int* __restrict__ lists, int* __restrict__ close_starts,
            int* __restrict__ close_lists)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
        const int box = row_boxes ? row_boxes[r] : r;
        if (FILL) {
            FillEmit e{lists + starts[r], close_lists ? close_lists + close_starts[r] : nullptr};
            if (KIND == 0) gen_colleagues<T, DIM>(t, rad, box, e);
            if (KIND == 1) gen_list1<T, DIM>(t, rad, box, e);
            if (KIND == 2) gen_list2<T, DIM>(t, rad, coll_starts, coll_lists, box, e);
            if (KIND == 4) gen_list4<T, DIM>(t, rad, coll_starts, coll_lists, with_extent, stick_out_factor, box, e);
        } else {
            CountEmit e;
            if (KIND == 0) gen_colleagues<T, DIM>(t, rad, box, e);
            if (KIND == 1) gen_list1<T, DIM>(t, rad, box, e);
            if (KIND == 2) gen_list2<T, DIM>(t, rad, coll_starts, coll_lists, box, e);
            if (KIND == 4) gen_list4<T, DIM>(t, rad, coll_starts, coll_lists, with_extent, stick_out_factor, box, e);
            starts[r] = e.c0;
            if (close_starts) close_starts[r] = e.c1;
        }
    }
}

// counts -> starts (in place), total to starts[n] and totals[slot]
struct InPlaceIn {
    const int* a;
    __device__ int operator()(int64_t i) const { return a[i]; }
};
struct InPlaceOut {
    int* a; long long* total_out; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { a[i] = (int)excl; }
    __device__ void total(long long t) const { a[n] = (int)t; if (total_out) *total_out = t; }
};

static int counts_to_starts(int* a, int64_t n, long long* total_out, cudaStream_t s)
{
    InPlaceIn in{a};
    InPlaceOut out{a, total_out, n};
    return scan_exclusive(n, nullptr, in, out, s);
}

template <typename T, int DIM>
static int build_list_impl(int kind, int phase, const bt_tree_view* tv, const bt_list_args* a, int nrows,
                           int* starts, int* lists, int* close_starts, int* close_lists,
                           long long* totals, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    const int grid = grid_for(nrows, kTravBlock, 16);
    const T sof = (T)a->stick_out_factor;
#define BT_LAUNCH_LIST(KIND, FILL)                                                              \
    list_kernel<T, DIM, KIND, FILL><<<grid, kTravBlock, 0, s>>>(                                \
        t, a->row_boxes, a->coll_starts, a->coll_lists, a->with_extent, sof, nrows, starts, lists, \
        close_starts, close_lists)
    if (nrows > 0) {
        if (phase == 0) {
            switch (kind) {
            case 0: BT_LAUNCH_LIST(0, false); break;
            case 1: BT_LAUNCH_LIST(1, false); break;
            case 2: BT_LAUNCH_LIST(2, false); break;
            case 4: BT_LAUNCH_LIST(4, false); break;
            default: return BT_ERR_BAD_ARG;
            }
        } else {
            switch (kind) {
            case 0: BT_LAUNCH_LIST(0, true); break;
            case 1: BT_LAUNCH_LIST(1, true); break;
            case 2: BT_LAUNCH_LIST(2, true); break;
            case 4: BT_LAUNCH_LIST(4, true); break;
            default: return BT_ERR_BAD_ARG;
            }
        }
        BT_LAUNCH_CHECK();
    }
#undef BT_LAUNCH_LIST
    if (phase == 0) {
        BT_TRY(counts_to_starts(starts, nrows, totals, s));
        if (close_starts) BT_TRY(counts_to_starts(close_starts, nrows, totals + 1, s));
    }
    return BT_OK;
}

// ---- b6 list 3: traversal.py:607-875 ---------------------------------------
template <typename T, int DIM>
struct List3Args {
    const int* target_boxes; const int* coll_starts; const int* coll_lists;
    T stick_out_factor; int targets_have_extent, sources_have_extent, crit;
    const T* bb_min; const T* bb_max; const int* box_source_counts_cumul; int min_nsources_cumul;
};

// E must provide append(level, box) and close(box)
template <typename T, int DIM, class E>
__device__ __forceinline__ void gen_list3(const TreeView<T, DIM>& t, const T* rad, const List3Args<T, DIM>& x,
                                          int tgt_box_id, E& e)
{
    constexpr int NB = 1 << DIM;
    T tc[DIM]; t.center(tgt_box_id, tc);
    const int tgt_level = t.levels[tgt_box_id];
    T tgt_stickout_l_inf_rad = 0, ext_center[DIM], radii_vec[DIM];
    if (x.targets_have_extent) {
        if (x.crit == 0 || x.crit == 2)
            tgt_stickout_l_inf_rad = (1 + x.stick_out_factor) * rad[tgt_level];
        else {   // load_true_box_extent, traversal.py:177-198
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const T mn = x.bb_min[a * t.aligned + tgt_box_id], mx = x.bb_max[a * t.aligned + tgt_box_id];
                ext_center[a] = ((T)0.5) * (mn + mx);
                radii_vec[a] = ((T)0.5) * (mx - mn);
            }
        }
    }
    const bool close_lists_exist = x.sources_have_extent || x.targets_have_extent;
    const T two_minus = (2 - 8 * CoordTraits<T>::eps());
    const int s = x.coll_starts[tgt_box_id], en = x.coll_starts[tgt_box_id + 1];
    for (int i = s; i < en; ++i) {
        const int same_lev_nws_box = x.coll_lists[i];
        if (same_lev_nws_box == tgt_box_id) continue;
        Walk w; w.init(same_lev_nws_box);
        while (w.cont) {
            const int wb = t.child(w.parent, w.mnr);
            const unsigned char cfl = t.flags[wb];
            if (wb && (cfl & (BT_BOX_IS_SOURCE_BOX | BT_BOX_HAS_SOURCE_CHILD_BOXES))) {
                T wc[DIM]; t.center(wb, wc);
                const int walk_level = t.levels[wb];
                if (adj_nbhd<T, DIM>(rad, tc, tgt_level, (T)1, wc, walk_level)) {
                    // single walk for all source levels: always descend
                    if (cfl & BT_BOX_HAS_SOURCE_CHILD_BOXES) { w.push(wb); continue; }
                } else {
                    bool meets;
                    if (!x.targets_have_extent) meets = true;
                    else if (x.crit == 0) {
                        const T source_rad = rad[walk_level];
                        T d = 0;
#pragma unroll
                        for (int a = 0; a < DIM; ++a)
                            d = fmax(d, fabs(tc[a] - wc[a]) - tgt_stickout_l_inf_rad - source_rad);
                        meets = d >= two_minus * source_rad;
                    } else if (x.crit == 1) {
                        const T source_rad = rad[walk_level];
                        T d = 0;
#pragma unroll
                        for (int a = 0; a < DIM; ++a)
                            d = fmax(d, fabs(ext_center[a] - wc[a]) - radii_vec[a] - source_rad);
                        meets = d >= two_minus * source_rad;
                    } else {
                        const T source_rad = rad[walk_level];
                        T l2sq = 0;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) l2sq = l2sq + (tc[a] - wc[a]) * (tc[a] - wc[a]);
                        const T rhs = sqrt(l2sq) - sqrt((T)DIM) * tgt_stickout_l_inf_rad - source_rad;
                        meets = (two_minus * source_rad <= rhs);
                    }
                    const bool force_close = close_lists_exist &&
                        (x.box_source_counts_cumul[wb] < x.min_nsources_cumul);
                    if (meets && !force_close) e.append(walk_level, wb);
                    else if (close_lists_exist) {
                        if (cfl & BT_BOX_IS_SOURCE_BOX) e.close(wb);
                        if (cfl & BT_BOX_HAS_SOURCE_CHILD_BOXES) { w.push(wb); continue; }
                    }
                }
            }
            w.template advance<NB>();
        }
    }
}

struct L3Count {
    int cnt[kMaxWalkLevels + 1]; int nlevels;
    __device__ __forceinline__ void append(int level, int) { ++cnt[level]; }
    __device__ __forceinline__ void close(int) { ++cnt[nlevels]; }
};
struct L3Fill {
    int cur[kMaxWalkLevels + 1]; int nlevels; int* lists;
    __device__ __forceinline__ void append(int level, int v) { lists[cur[level]++] = v; }
    __device__ __forceinline__ void close(int v) { lists[cur[nlevels]++] = v; }
};

template <typename T, int DIM, bool FILL>
__global__ void __launch_bounds__(kTravBlock)
list3_kernel(TreeView<T, DIM> t, List3Args<T, DIM> x, int ntgt, int* __restrict__ G /*[nlevels+1][ntgt+1]*/,
             int* __restrict__ lists)
{
    __shared__ T rad[kMaxWalkLevels];
    fill_rad_table(rad, t.root_extent);
    const int nl = t.nlevels;
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int stride = gridDim.x * blockDim.x;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < ntgt; r += stride) {
        const int box = x.target_boxes[r];
        if (FILL) {
            L3Fill e; e.nlevels = nl; e.lists = lists;
            for (int l = 0; l <= nl; ++l) e.cur[l] = G[l * rowlen + r];
            gen_list3<T, DIM>(t, rad, x, box, e);
        } else {
            L3Count e; e.nlevels = nl;
            for (int l = 0; l <= nl; ++l) e.cnt[l] = 0;
            gen_list3<T, DIM>(t, rad, x, box, e);
            for (int l = 0; l <= nl; ++l) G[l * rowlen + r] = e.cnt[l];
        }
    }
}

// after the flattened scan of G: C = flattened exclusive scan of row-nonempty flags
struct NonemptyIn {
    const int* G; int64_t total_len; int64_t rowlen;
    __device__ int operator()(int64_t i) const
    {
        if ((i % rowlen) == rowlen - 1) return 0;          // padding slot of each row
        const int nxt = G[i + 1];
        return (nxt - G[i]) > 0 ? 1 : 0;
    }
};
struct PlainOut {
    int* a; long long* total_out; int64_t n;
    __device__ void operator()(int64_t i, long long excl) const { a[i] = (int)excl; }
    __device__ void total(long long t) const { a[n] = (int)t; if (total_out) *total_out = t; }
};

__global__ void list3_summary_kernel(const int* __restrict__ G, const int* __restrict__ C, int nrows,
                                     int64_t rowlen, long long* __restrict__ summary)
{   // summary[l] = G[l][0] (l = 0..nrows, last = total), summary[nrows+1+l] = C[l][0]
    const int l = threadIdx.x;
    if (l <= nrows) {
        summary[l] = G[(int64_t)l * rowlen];
        summary[nrows + 1 + l] = C[(int64_t)l * rowlen];
    }
}

template <typename T, int DIM>
static int list3_impl(int phase, const bt_tree_view* tv, const bt_list3_args* a, int ntgt, int* G, int* C,
                      int* lists, long long* summary, cudaStream_t s)
{
    TreeView<T, DIM> t = make_view<T, DIM>(tv);
    if (t.nlevels > kMaxWalkLevels) return BT_ERR_UNSUPPORTED;
    List3Args<T, DIM> x{a->target_boxes, a->coll_starts, a->coll_lists, (T)a->stick_out_factor,
                        a->targets_have_extent, a->sources_have_extent, a->crit,
                        (const T*)a->box_target_bounding_box_min, (const T*)a->box_target_bounding_box_max,
                        a->box_source_counts_cumul, a->min_nsources_cumul};
    const int nrows = t.nlevels + 1;
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int64_t total_len = rowlen * nrows;
    const int grid = grid_for(ntgt, kTravBlock, 16);
    if (phase == 0) {
        BT_CHECK(cudaMemsetAsync(G, 0, sizeof(int) * (total_len + 1), s));
        if (ntgt > 0) {
            list3_kernel<T, DIM, false><<<grid, kTravBlock, 0, s>>>(t, x, ntgt, G, nullptr);
            BT_LAUNCH_CHECK();
        }
        InPlaceIn in{G};
        PlainOut out{G, nullptr, total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, in, out, s));
        NonemptyIn nin{G, total_len, rowlen};
        PlainOut nout{C, nullptr, total_len};
        BT_TRY(scan_exclusive(total_len, nullptr, nin, nout, s));
        list3_summary_kernel<<<1, 64, 0, s>>>(G, C, nrows, rowlen, summary);
        BT_LAUNCH_CHECK();
    } else if (ntgt > 0) {
        list3_kernel<T, DIM, true><<<grid, kTravBlock, 0, s>>>(t, x, ntgt, G, lists);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

// per-level compressed CSR (eliminate_empty_output_lists) -- one kernel for all levels
__global__ void __launch_bounds__(256)
list3_compress_kernel(int nlevels, int ntgt, const int* __restrict__ G, const int* __restrict__ C,
                      const int* __restrict__ target_boxes, int* __restrict__ cstarts,
                      int* __restrict__ nonempty_indices, int* __restrict__ tb_nonempty,
                      int* __restrict__ compressed_indices /*[nlevels][ntgt+1]*/,
                      int* __restrict__ close_starts /*[ntgt+1] or null*/)
{
    const int64_t rowlen = (int64_t)ntgt + 1;
    const int64_t total = rowlen * (nlevels + 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int l = (int)(i / rowlen), tt = (int)(i % rowlen);
        const int g0 = G[(int64_t)l * rowlen];
        const int local_start = G[i] - g0;
        if (l == nlevels) { if (close_starts) close_starts[tt] = local_start; continue; }
        const int c0 = C[(int64_t)l * rowlen];
        const int ci = C[i] - c0;
        compressed_indices[i] = ci;
        const int soff = c0 + l;             // sum over previous levels of (nonempty + 1)
        if (tt < ntgt) {
            if (G[i + 1] - G[i] > 0) {
                nonempty_indices[c0 + ci] = tt;
                tb_nonempty[c0 + ci] = target_boxes[tt];
                cstarts[soff + ci] = local_start;
            }
        } else cstarts[soff + ci] = local_start;   // total of the level
    }
}

// ---- box lists, level starts -------------------------------------------------
struct BoxListIn {
    const unsigned char* flags; const signed char* mask; int which;
    __device__ int operator()(int64_t b) const
    {
        const unsigned char fl = flags[b];
        bool k;
        if (which == 0) k = (fl & BT_BOX_HAS_SOURCE_CHILD_BOXES) && (!mask || mask[b]);
        else if (which == 1) k = (fl & BT_BOX_IS_SOURCE_BOX) && (!mask || mask[b]);
        else if (which == 2) k = (fl & (BT_BOX_HAS_TARGET_CHILD_BOXES | BT_BOX_IS_TARGET_BOX));
        else k = (fl & BT_BOX_IS_TARGET_BOX);
        return k ? 1 : 0;
    }
};
struct BoxListOut {
    BoxListIn in; int* out; int* count;
    __device__ void operator()(int64_t b, long long excl) const { if (in(b)) out[excl] = (int)b; }
    __device__ void total(long long t) const { *count = (int)t; }
};

__global__ void level_starts_kernel(int nlevels, const int* __restrict__ level_start_box_nrs,
                                    const int* __restrict__ list, int nlist, int* __restrict__ out)
{
    const int lev = blockIdx.x * blockDim.x + threadIdx.x;
    if (lev > nlevels) return;
    if (lev == nlevels) { out[lev] = nlist; return; }
    const int key = level_start_box_nrs[lev];
    int lo = 0, hi = nlist;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] < key) lo = mid + 1; else hi = mid; }
    out[lev] = lo;
}

// ---- list merger: traversal.py:1153-1214 ------------------------------------
struct MergeArgs { const int* starts[3]; const int* lists[3]; int nlists; };

struct MergeIn {
    MergeArgs m; const int* o2i;
    __device__ int operator()(int64_t i) const
    {
        const int ibox = o2i ? o2i[i] : (int)i;
        int tot = 0;
        for (int l = 0; l < m.nlists; ++l) tot += m.starts[l][ibox + 1] - m.starts[l][ibox];
        return tot;
    }
};

__global__ void __launch_bounds__(256)
merge_write_kernel(MergeArgs m, const int* __restrict__ o2i, int nout, const int* __restrict__ new_starts,
                   int* __restrict__ new_lists)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += stride) {
        const int ibox = o2i ? o2i[i] : i;
        int cur = new_starts[i];
        for (int l = 0; l < m.nlists; ++l) {
            const int s = m.starts[l][ibox], c = m.starts[l][ibox + 1] - s;
            for (int j = 0; j < c; ++j) new_lists[cur++] = m.lists[l][s + j];
        }
    }
}

}  // namespace bt

#define BT_DISPATCH(dtype, dim, FN, ...)                                             \
    do {                                                                             \
        if ((dtype) == BT_F32) {                                                     \
            if ((dim) == 1) return bt::FN<float, 1>(__VA_ARGS__);                     \
            if ((dim) == 2) return bt::FN<float, 2>(__VA_ARGS__);                     \
            if ((dim) == 3) return bt::FN<float, 3>(__VA_ARGS__);                     \
        } else if ((dtype) == BT_F64) {                                              \
            if ((dim) == 1) return bt::FN<double, 1>(__VA_ARGS__);                    \
            if ((dim) == 2) return bt::FN<double, 2>(__VA_ARGS__);                    \
            if ((dim) == 3) return bt::FN<double, 3>(__VA_ARGS__);                    \
        }                                                                            \
        return BT_ERR_BAD_ARG;                                                       \
    } while (0)

extern "C" {

int bt_trav_box_list(int which, int nboxes, const uint8_t* box_flags, const int8_t* mask,
                     int32_t* out_list, int32_t* count_dev, void* stream)
{
    BT_PROF("bt_trav_box_list", (cudaStream_t)stream);
    bt::BoxListIn in{box_flags, (const signed char*)mask, which};
    bt::BoxListOut out{in, out_list, count_dev};
    return bt::scan_exclusive(nboxes, nullptr, in, out, (cudaStream_t)stream);
}

int bt_trav_level_starts(int nlevels, const int32_t* level_start_box_nrs, const int32_t* box_list,
                         int nlist, int32_t* out, void* stream)
{
    BT_PROF("bt_trav_level_starts", (cudaStream_t)stream);
    bt::level_starts_kernel<<<(nlevels + 1 + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        nlevels, level_start_box_nrs, box_list, nlist, out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_trav_build_list(int dtype, int kind, int phase, const bt_tree_view* tree, const bt_list_args* args,
                       int nrows, int32_t* starts, int32_t* lists, int32_t* close_starts,
                       int32_t* close_lists, int64_t* totals_dev, void* stream)
{
    static const char* const kNames[2][5] = {
        {"trav_colleagues_count", "trav_list1_count", "trav_list2_count", "?", "trav_list4_count"},
        {"trav_colleagues_fill", "trav_list1_fill", "trav_list2_fill", "?", "trav_list4_fill"}};
    BT_PROF(kNames[phase ? 1 : 0][(kind >= 0 && kind <= 4) ? kind : 3], (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, build_list_impl, kind, phase, tree, args, nrows, starts, lists,
                close_starts, close_lists, (long long*)totals_dev, (cudaStream_t)stream);
}

int bt_trav_list3(int dtype, int phase, const bt_tree_view* tree, const bt_list3_args* args,
                  int ntarget_boxes, int32_t* G, int32_t* C, int32_t* lists, int64_t* summary_dev,
                  void* stream)
{
    BT_PROF(phase ? "trav_list3_fill" : "trav_list3_count", (cudaStream_t)stream);
    BT_DISPATCH(dtype, tree->dim, list3_impl, phase, tree, args, ntarget_boxes, G, C, lists,
                (long long*)summary_dev, (cudaStream_t)stream);
}

int bt_trav_list3_compress(int nlevels, int ntarget_boxes, const int32_t* G, const int32_t* C,
                           const int32_t* target_boxes, int32_t* compressed_starts,
                           int32_t* nonempty_indices, int32_t* target_boxes_nonempty,
                           int32_t* compressed_indices, int32_t* close_starts, void* stream)
{
    BT_PROF("bt_trav_list3_compress", (cudaStream_t)stream);
    const int64_t total = ((int64_t)ntarget_boxes + 1) * (nlevels + 1);
    bt::list3_compress_kernel<<<bt::grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        nlevels, ntarget_boxes, G, C, target_boxes, compressed_starts, nonempty_indices,
        target_boxes_nonempty, compressed_indices, close_starts);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

__global__ void bt_gather_i32_kernel(int64_t n, const int* __restrict__ src, const int* __restrict__ idx,
                                     int* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

int bt_gather_i32(int64_t n, const int32_t* src, const int32_t* idx, int32_t* out, void* stream)
{
    BT_PROF("bt_gather_i32", (cudaStream_t)stream);
    if (n <= 0) return BT_OK;
    bt_gather_i32_kernel<<<bt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, src, idx, out);
    BT_LAUNCH_CHECK();
    return BT_OK;
}

int bt_trav_merge_lists(int phase, int noutput, const int32_t* output_to_input_box, int nlists,
                        const int32_t* const* starts, const int32_t* const* lists, int32_t* new_starts,
                        int32_t* new_lists, int64_t* totals_dev, void* stream)
{
    BT_PROF("bt_trav_merge_lists", (cudaStream_t)stream);
    if (nlists < 1 || nlists > 3) return BT_ERR_BAD_ARG;
    bt::MergeArgs m;
    m.nlists = nlists;
    for (int l = 0; l < 3; ++l) { m.starts[l] = l < nlists ? starts[l] : nullptr; m.lists[l] = l < nlists ? lists[l] : nullptr; }
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0) {
        bt::MergeIn in{m, output_to_input_box};
        bt::PlainOut out{new_starts, (long long*)totals_dev, noutput};
        return bt::scan_exclusive(noutput, nullptr, in, out, s);
    }
    if (noutput > 0) {
        bt::merge_write_kernel<<<bt::grid_for(noutput, 256, 8), 256, 0, s>>>(m, output_to_input_box, noutput,
                                                                             new_starts, new_lists);
        BT_LAUNCH_CHECK();
    }
    return BT_OK;
}

}  // extern "C"
