/*
 * oracle/oracle_tree.c -- CPU restatement of the reference tree-build kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product path (boxtree_b200/); only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY PINNED against the reference itself.  The reference's PyOpenCL path cannot
 * run as shipped in this image (no pyopencl / OpenCL ICD / mako), so tests/refexec
 * supplies stand-ins for those third-party modules that execute every kernel the
 * reference renders, one work item at a time, compiled with g++; the reference's
 * TreeBuilder.__call__ host code and kernel text run unmodified from
 * /root/reference.  This restatement reproduces every Tree field (dtype, shape and
 * bytes) of those runs: 136 + 196 sweep cases (1-3 D, fp32/fp64, all tree kinds,
 * weights, extents, user bounding box, MaxLevelsExceeded), the BASELINE
 * configurations including config 3 and uniform at 1e7 points
 * (tests/test_refexec.py, tests/golden/make_*golden.py).  Also pinned by reference
 * code run in place: drive_fmm + ConstantOneExpansionWrangler on the oracle's
 * lists, tree_of_boxes.py inputs (tests/test_reference_consumer.py,
 * tests/test_tree_of_boxes.py).  This file restates, kernel by kernel, the OpenCL
 * kernels the reference generates; every function cites the reference
 * file:line it follows (paths relative to /root/reference/).
 *
 * Floating point rules: every expression is evaluated in coord_t (float or
 * double, chosen at compile time with -DCOORD_F32 / -DCOORD_F64), IEEE
 * division and sqrt, no FMA contraction (-ffp-contract=off), C truncation for
 * float->unsigned casts, and OpenCL shift semantics (shift count taken modulo
 * the operand width).
 *
 * Compiled twice (f32/f64) by oracle/build.py into oracle/_build/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <math.h>
#include <float.h>
#include <limits.h>

#if defined(COORD_F32)
typedef float coord_t;
#define COORD_MAX FLT_MAX
#define COORD_SQRT sqrtf
#define COORD_FABS fabsf
#define COORD_FMAX fmaxf
#define COORD_FMIN fminf
#define COORD_EPS FLT_EPSILON
#elif defined(COORD_F64)
typedef double coord_t;
#define COORD_MAX DBL_MAX
#define COORD_SQRT sqrt
#define COORD_FABS fabs
#define COORD_FMAX fmax
#define COORD_FMIN fmin
#define COORD_EPS DBL_EPSILON
#else
#error "define COORD_F32 or COORD_F64"
#endif

typedef int32_t box_id_t;
typedef int32_t particle_id_t;
typedef int32_t refine_weight_t;
typedef uint8_t box_level_t;
typedef int8_t morton_nr_t;
typedef uint8_t box_flags_t;

#define MAXDIM 3

/* OpenCL: shift counts are reduced modulo the width of the promoted operand */
static inline unsigned ocl_shl_u(unsigned v, int s) { return v << (s & 31); }
static inline int ocl_shl_i(int v, int s) { return (int)((unsigned)v << (s & 31)); }

/* ------------------------------------------------------------------------
 * a1: bounding box reduction -- boxtree/bounding_box.py:54-122
 * out_min/out_max: [d]
 * ---------------------------------------------------------------------- */
void orc_bounding_box(int d, int64_t n, const coord_t *const *coords,
                      const coord_t *radii, coord_t *out_min, coord_t *out_max)
{
    for (int a = 0; a < d; ++a) { out_min[a] = COORD_MAX; out_max[a] = -COORD_MAX; }
    for (int64_t i = 0; i < n; ++i) {
        coord_t r = radii ? radii[i] : (coord_t)0;
        for (int a = 0; a < d; ++a) {
            coord_t lo = coords[a][i] - r, hi = coords[a][i] + r;
            /* OpenCL min/max: (y < x) ? y : x  and  (x < y) ? y : x */
            out_min[a] = (lo < out_min[a]) ? lo : out_min[a];
            out_max[a] = (out_max[a] < hi) ? hi : out_max[a];
        }
    }
}

/* ------------------------------------------------------------------------
 * Morton bin count struct, as a row of int32:
 *   [nonchild (only with extents)] [pcnt x 2^d] [pwt x 2^d]
 * boxtree/tree_build_kernels.py:158-189
 * ---------------------------------------------------------------------- */
static inline int mbc_width(int d, int have_extent) { return (have_extent ? 1 : 0) + 2 * (1 << d); }

static inline int add_sat_i32(int a, int b)
{ /* my_add_sat, tree_build_kernels.py:270-274 (same result as OpenCL add_sat) */
    long long r = (long long)a + b;
    if (r > INT_MAX) return INT_MAX;
    if (r < INT_MIN) return INT_MIN;
    return (int)r;
}

/* extent_norm: 0 = none, 1 = linf, 2 = l2 */
/* scan_t_from_particle -- tree_build_kernels.py:308-470 */
static int morton_nr_of_particle(
    int d, int extent_norm, int particle_level,
    const coord_t *bbox_min, const coord_t *bbox_max,
    particle_id_t user_id, const coord_t *const *coords,
    const coord_t *radii, coord_t stick_out_factor)
{
    coord_t next_level_box_size_factor =
        ((coord_t)1) / ((coord_t)ocl_shl_u(1U, 1 + particle_level));
    int stop = 0;
    coord_t radius = extent_norm ? radii[user_id] : (coord_t)0;
    const coord_t one_half = ((coord_t)1) / 2;
    /* "(1. + stick_out_factor) * one_half" is evaluated in double (the literal
       1. is a double in OpenCL C) and rounded on assignment to coord_t */
    const coord_t box_radius_factor = (coord_t)(
        (1. + (double)(extent_norm ? stick_out_factor : (coord_t)0)) * (double)one_half);

    unsigned bits[MAXDIM];
    coord_t center[MAXDIM], gext[MAXDIM], pos[MAXDIM];
    for (int a = 0; a < d; ++a) {
        coord_t gmin = bbox_min[a];
        gext[a] = bbox_max[a] - gmin;
        pos[a] = coords[a][user_id];
        bits[a] = (unsigned)(((pos[a] - gmin) / gext[a])
                             * (coord_t)ocl_shl_u(1U, 1 + particle_level));
        center[a] = gmin + gext[a] * ((coord_t)bits[a] + one_half) * next_level_box_size_factor;
    }

    if (extent_norm == 1) {
        for (int a = 0; a < d; ++a) {
            coord_t so_rad = box_radius_factor * gext[a] * next_level_box_size_factor;
            stop = stop || (pos[a] + radius >= center[a] + so_rad);
            stop = stop || (pos[a] - radius < center[a] - so_rad);
        }
    } else if (extent_norm == 2) {
        coord_t so_rad = box_radius_factor * gext[0] * next_level_box_size_factor;
        coord_t acc = 0;
        for (int a = 0; a < d; ++a)
            acc = acc + (pos[a] - center[a]) * (pos[a] - center[a]);
        coord_t dist = COORD_SQRT(acc) + radius;
        stop = stop || (dist * dist >= d * so_rad * so_rad);
    }

    int mnr = 0;
    for (int a = 0; a < d; ++a)
        mnr |= (int)(bits[a] & 1U) << (d - 1 - a);
    if (extent_norm && stop) mnr = -1;
    return mnr;
}

/* Test helper: the Morton number (or -1 = "stops here") scan_t_from_particle computes for every
 * particle sitting in a box of level particle_level -- the per-level digit the reference's
 * morton_count_scan bins by (tree_build_kernels.py:308-470).  tests/test_gpu_morton_keys.py
 * holds the digits packed into the CUDA path's sort keys to these. */
void orc_particle_morton_nrs(int d, int extent_norm, int particle_level, int64_t n,
                             const coord_t *bbox_min, const coord_t *bbox_max,
                             const coord_t *const *coords, const coord_t *radii,
                             coord_t stick_out_factor, morton_nr_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        out[i] = (morton_nr_t)morton_nr_of_particle(d, extent_norm, particle_level, bbox_min, bbox_max,
                                                    (particle_id_t)i, coords, radii, stick_out_factor);
}

/* morton_count_scan: segmented inclusive scan + output statement
 * tree_build_kernels.py:247-508, 1555-1572; called from tree_build.py:732 */
void orc_morton_count_scan(
    int d, int extent_norm, int64_t n,
    int32_t *morton_bin_counts /* [n, w] */, morton_nr_t *morton_nrs,
    const int8_t *box_start_flags, const box_id_t *srcntgt_box_ids,
    int32_t *box_morton_bin_counts /* [nboxes, w] */,
    const refine_weight_t *refine_weights,
    const particle_id_t *box_srcntgt_counts_cumul, const box_level_t *box_levels,
    const coord_t *bbox_min, const coord_t *bbox_max,
    const particle_id_t *user_srcntgt_ids, const coord_t *const *coords,
    const coord_t *radii, coord_t stick_out_factor)
{
    const int have_ext = extent_norm != 0;
    const int w = mbc_width(d, have_ext), nb = 1 << d, off = have_ext ? 1 : 0;
    /* The reference's GenericScanKernel is a parallel segmented scan; here the particle range is
       cut into one chunk per thread: (1) every chunk scans its own particles, (2) the value
       carried into each chunk is chained serially over the few chunks, (3) the carry is added to
       the chunk's particles before its first segment start, then the output statement runs.
       scan_t_add (:277-302) is associative (int adds; saturating adds of non-negative weights),
       so the result equals the sequential scan's. */
    enum { W = 1 + 2 * 8 };
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if (n < 65536) nthreads = 1;
    int32_t *chunk_end = (int32_t *)calloc((size_t)nthreads * W, sizeof(int32_t));
    int32_t *carry = (int32_t *)calloc((size_t)nthreads * W, sizeof(int32_t));
    int64_t *first_flag = (int64_t *)malloc(sizeof(int64_t) * nthreads);
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        int32_t acc[W], item[W];
        memset(acc, 0, sizeof acc);
        int64_t ff = hi;               /* first segment start inside the chunk */
        for (int64_t i = lo; i < hi; ++i) {
            particle_id_t uid = user_srcntgt_ids[i];
            int mnr = morton_nr_of_particle(d, extent_norm,
                box_levels[srcntgt_box_ids[i]], bbox_min, bbox_max, uid, coords,
                radii, stick_out_factor);
            morton_nrs[i] = (morton_nr_t)mnr;
            memset(item, 0, sizeof(int32_t) * w);
            if (have_ext) item[0] = (mnr == -1);
            for (int m = 0; m < nb; ++m) {
                item[off + m] = (mnr == m);
                item[off + nb + m] = (mnr == m) ? refine_weights[uid] : 0;
            }
            const int seg = (i == 0 || box_start_flags[i]);
            if (seg && ff == hi) ff = i;
            if (seg || i == lo) {
                memcpy(acc, item, sizeof(int32_t) * w);
            } else { /* scan_t_add, :277-302 */
                if (have_ext) acc[0] += item[0];
                for (int m = 0; m < nb; ++m) {
                    acc[off + m] = acc[off + m] + item[off + m];
                    acc[off + nb + m] = add_sat_i32(acc[off + nb + m], item[off + nb + m]);
                }
            }
            memcpy(morton_bin_counts + i * w, acc, sizeof(int32_t) * w);
        }
        first_flag[t] = ff;
        memcpy(chunk_end + (size_t)t * W, acc, sizeof(int32_t) * w);
#pragma omp barrier
#pragma omp single
        {
            /* carry[t]: scan value just before chunk t, within the segment open at its start */
            for (int c = 1; c < nthreads; ++c) {
                const int64_t clo = n * (c - 1) / nthreads, chi = n * c / nthreads;
                if (chi == clo) { memcpy(carry + (size_t)c * W, carry + (size_t)(c - 1) * W, sizeof(int32_t) * w); continue; }
                int32_t *dst = carry + (size_t)c * W;
                memcpy(dst, chunk_end + (size_t)(c - 1) * W, sizeof(int32_t) * w);
                if (first_flag[c - 1] == chi) {     /* no segment start in chunk c-1: chain */
                    const int32_t *pc = carry + (size_t)(c - 1) * W;
                    if (have_ext) dst[0] += pc[0];
                    for (int m = 0; m < nb; ++m) {
                        dst[off + m] += pc[off + m];
                        dst[off + nb + m] = add_sat_i32(dst[off + nb + m], pc[off + nb + m]);
                    }
                }
            }
        }
        /* (implicit barrier after single) */
        const int32_t *cin = carry + (size_t)t * W;
        for (int64_t i = lo; i < hi; ++i) {
            int32_t *a = morton_bin_counts + i * w;
            if (t > 0 && i < ff) {
                if (have_ext) a[0] += cin[0];
                for (int m = 0; m < nb; ++m) {
                    a[off + m] += cin[off + m];
                    a[off + nb + m] = add_sat_i32(a[off + nb + m], cin[off + nb + m]);
                }
            }
            /* output statement, :480-508 */
            particle_id_t my_id_in_my_box = -1;
            for (int k = 0; k < off + nb; ++k) my_id_in_my_box += a[k];
            box_id_t cur = srcntgt_box_ids[i];
            if (my_id_in_my_box + 1 == box_srcntgt_counts_cumul[cur])
                memcpy(box_morton_bin_counts + (int64_t)cur * w, a, sizeof(int32_t) * w);
        }
    }
    free(chunk_end); free(carry); free(first_flag);
}

/* split_box_id_scan -- tree_build_kernels.py:514-640; tree_build.py:740-759
 * scan over boxes [0, size), segmented by level. */
void orc_split_box_id_scan(
    int d, int have_extent, int adaptive, int level_restrict, int64_t size,
    const particle_id_t *box_srcntgt_counts_cumul,
    const int32_t *box_morton_bin_counts, refine_weight_t max_leaf_refine_weight,
    const box_level_t *box_levels, const box_id_t *level_start_box_ids,
    const box_id_t *level_used_box_counts, const int32_t *box_force_split,
    int last_level,
    int32_t *box_has_children, box_id_t *split_box_ids, int32_t *have_oversize_split_box)
{
    const int w = mbc_width(d, have_extent), nb = 1 << d, off = have_extent ? 1 : 0;
    box_id_t acc = 0;
    for (int64_t i = 0; i < size; ++i) {
        /* count_new_boxes_needed, :535-614 */
        box_id_t result = 0;
        int level = box_levels[i];
        if ((box_id_t)i == level_start_box_ids[level]) {
            result += level_start_box_ids[level + 1];
            result += level_used_box_counts[level + 1];
        }
        const int32_t *bins = box_morton_bin_counts + i * w;
        particle_id_t nonchild = have_extent ? bins[0] : 0;
        refine_weight_t box_refine_weight = 0;
        for (int m = 0; m < nb; ++m)
            box_refine_weight = add_sat_i32(box_refine_weight, bins[off + nb + m]);
        int cond = (level + 1 == last_level) &&
            (adaptive ? (box_refine_weight > max_leaf_refine_weight)
                      : (box_srcntgt_counts_cumul[i] - nonchild >= 0));
        if (level_restrict) cond = cond || box_force_split[i];
        if (cond) {
            result += nb;
            box_has_children[i] = 1;
            refine_weight_t mx = 0;
            for (int m = 0; m < nb; ++m)
                mx = (bins[off + nb + m] > mx) ? bins[off + nb + m] : mx;
            if (mx > max_leaf_refine_weight) *have_oversize_split_box = 1;
        }
        /* scan_expr "across_seg_boundary ? b : a + b", segment start when the
           level changes (:632-634) */
        int seg_start = (i == 0) || (box_levels[i] != box_levels[i - 1]);
        acc = seg_start ? result : acc + result;
        split_box_ids[i] = acc;
    }
}

/* box_splitter -- tree_build_kernels.py:646-711; tree_build.py:1064-1076 */
void orc_box_splitter(
    int d, int have_extent, int level_restrict, int64_t nboxes_range, int level,
    const int32_t *box_morton_bin_counts, int8_t *box_start_flags,
    const box_id_t *split_box_ids, particle_id_t *box_srcntgt_starts,
    particle_id_t *box_srcntgt_counts_cumul, box_id_t *box_parent_ids,
    box_level_t *box_levels, const int32_t *box_has_children,
    const int32_t *box_force_split, coord_t root_extent,
    box_id_t *const *box_child_ids /* [2^d] */, coord_t *const *box_centers /* [d] */)
{
    const int w = mbc_width(d, have_extent), nb = 1 << d, off = have_extent ? 1 : 0;
    for (int64_t ibox = 0; ibox < nboxes_range; ++ibox) {
        int do_split = (box_has_children[ibox] && box_levels[ibox] + 1 == level);
        if (level_restrict) do_split = do_split || box_force_split[ibox];
        if (!do_split) continue;
        const int32_t *bins = box_morton_bin_counts + ibox * w;
        for (int mnr = 0; mnr < nb; ++mnr) {
            box_id_t nb_id = split_box_ids[ibox] - nb + mnr;
            box_parent_ids[nb_id] = (box_id_t)ibox;
            box_child_ids[mnr][ibox] = nb_id;
            box_level_t new_level = box_levels[ibox] + 1;
            box_levels[nb_id] = new_level;
            particle_id_t new_count = bins[off + mnr];
            box_srcntgt_counts_cumul[nb_id] = new_count;
            if (new_count > 0) {
                particle_id_t st = box_srcntgt_starts[ibox];
                if (have_extent) st += bins[0];
                for (int s = 0; s < mnr; ++s) st += bins[off + s];
                box_start_flags[st] = 1;
                box_srcntgt_starts[nb_id] = st;
            }
            coord_t radius = (root_extent * 1 / (coord_t)ocl_shl_i(1, 1 + new_level));
            for (int a = 0; a < d; ++a) {
                int has_bit = mnr & (1 << (d - 1 - a));
                box_centers[a][nb_id] = has_bit ? box_centers[a][ibox] + radius
                                                : box_centers[a][ibox] - radius;
            }
        }
    }
}

/* particle renumberer -- tree_build_kernels.py:717-819; tree_build.py:1101-1121 */
void orc_particle_renumberer(
    int d, int have_extent, int level_restrict, int64_t n, int level,
    const int32_t *morton_bin_counts, const morton_nr_t *morton_nrs,
    const box_id_t *srcntgt_box_ids, const box_id_t *split_box_ids,
    const int32_t *box_morton_bin_counts, const particle_id_t *box_srcntgt_starts,
    const box_level_t *box_levels, const particle_id_t *user_srcntgt_ids,
    const int32_t *box_has_children, const int32_t *box_force_split,
    particle_id_t *new_user_srcntgt_ids, box_id_t *new_srcntgt_box_ids)
{
    const int w = mbc_width(d, have_extent), nb = 1 << d, off = have_extent ? 1 : 0;
    /* independent work items (every particle writes its own destination) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        box_id_t ibox = srcntgt_box_ids[i];
        int do_split = (box_has_children[ibox] && box_levels[ibox] + 1 == level);
        if (level_restrict) do_split = do_split || box_force_split[ibox];
        if (!do_split) {
            new_user_srcntgt_ids[i] = user_srcntgt_ids[i];
            new_srcntgt_box_ids[i] = ibox;
            continue;
        }
        int mnr = morton_nrs[i];
        const int32_t *box_bins = box_morton_bin_counts + (int64_t)ibox * w;
        const int32_t *my_bins = morton_bin_counts + i * w;
        particle_id_t my_count = (mnr == -1) ? my_bins[0] : my_bins[off + mnr];
        particle_id_t tgt = box_srcntgt_starts[ibox] + my_count - 1;
        if (have_extent) tgt += (mnr >= 0) ? box_bins[0] : 0;
        for (int m = 0; m < nb; ++m) tgt += (mnr > m) ? box_bins[off + m] : 0;
        new_user_srcntgt_ids[tgt] = user_srcntgt_ids[i];
        box_id_t new_box_id = split_box_ids[ibox] - nb + mnr;
        if (have_extent && mnr == -1) new_box_id = ibox;
        new_srcntgt_box_ids[tgt] = new_box_id;
    }
}

/* adjacency predicate -- traversal.py:279-318; LEVEL_TO_RAD :234-235 */
static inline coord_t level_to_rad(coord_t root_extent, int level)
{ return (root_extent * 1 / (coord_t)ocl_shl_i(1, level + 1)); }

static inline int is_adjacent_or_overlapping_with_neighborhood(
    int d, coord_t root_extent, const coord_t *target_center, int target_level,
    coord_t target_box_neighborhood_size, const coord_t *source_center, int source_level)
{
    coord_t target_rad = level_to_rad(root_extent, target_level);
    coord_t source_rad = level_to_rad(root_extent, source_level);
    coord_t rad_sum = ((2 * (target_box_neighborhood_size - 1) + 1) * target_rad + source_rad);
    coord_t slack = rad_sum + COORD_FMIN(target_rad, source_rad);
    coord_t l_inf_dist = 0;
    for (int a = 0; a < d; ++a)
        l_inf_dist = COORD_FMAX(l_inf_dist, COORD_FABS(target_center[a] - source_center[a]));
    return l_inf_dist <= slack;
}

/* level_restrict kernel -- tree_build_kernels.py:825-970; tree_build.py:1181-1189
 * runs over boxes [slice_start, slice_start+slice_count) of one upper level.
 * The walk uses the per-Morton child arrays and per-axis centre arrays. */
void orc_level_restrict(
    int d, int level, coord_t root_extent, int64_t slice_start, int64_t slice_count,
    const int32_t *box_has_children, int32_t *box_force_split,
    int32_t *have_upper_level_split_box,
    const box_id_t *const *box_child_ids, const coord_t *const *box_centers)
{
    const int nb = 1 << d;
    enum { MAXLEV = 128 };
    /* work items only write their own box_force_split entry and read those of the level below
       (set by the previous launch of the sweep): independent, like the OpenCL kernel */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t bi = slice_start; bi < slice_start + slice_count; ++bi) {
        box_id_t box_id = (box_id_t)bi;
        if (box_has_children[box_id]) continue;
        box_id_t stack_box[MAXLEV]; int stack_mnr[MAXLEV];
        int stack_size = 0; box_id_t walk_parent = 0; int walk_mnr = 0; int cont = 1;
        while (cont) {
            box_id_t child = box_child_ids[walk_mnr][walk_parent];
            if (child) {
                int child_level = stack_size + 1;
                int adj;
                if (child == box_id) adj = 0;
                else {
                    coord_t bc[MAXDIM], cc[MAXDIM];
                    for (int a = 0; a < d; ++a) { bc[a] = box_centers[a][box_id]; cc[a] = box_centers[a][child]; }
                    adj = is_adjacent_or_overlapping_with_neighborhood(
                        d, root_extent, cc, child_level, (coord_t)1, bc, level);
                }
                if (adj) {
                    if (box_has_children[child]) {
                        if (child_level <= 1 + level) {
                            stack_box[stack_size] = walk_parent; stack_mnr[stack_size] = walk_mnr;
                            ++stack_size; walk_parent = child; walk_mnr = 0;
                            continue;
                        }
                    } else {
                        if (child_level == 2 + level ||
                            (child_level == 1 + level && box_force_split[child])) {
                            box_force_split[box_id] = 1;
#pragma omp atomic
                            *have_upper_level_split_box |= 1;
                            cont = 0;
                        }
                    }
                }
            }
            /* walk_advance, traversal.py:115-143 */
            for (;;) {
                ++walk_mnr;
                if (walk_mnr < nb) break;
                cont = (stack_size > 0);
                if (cont) { --stack_size; walk_parent = stack_box[stack_size]; walk_mnr = stack_mnr[stack_size]; }
                else break;
            }
        }
    }
}

/* extract_nonchild_srcntgt_count -- tree_build_kernels.py:979-1007 */
void orc_extract_nonchild_srcntgt_count(
    int d, int64_t nboxes, const int32_t *box_morton_bin_counts,
    const particle_id_t *box_srcntgt_counts_cumul, box_id_t highest_possibly_split_box_nr,
    particle_id_t *box_srcntgt_counts_nonchild)
{
    const int w = mbc_width(d, 1);
    for (int64_t i = 0; i < nboxes; ++i) {
        if (i >= highest_possibly_split_box_nr) box_srcntgt_counts_nonchild[i] = 0;
        else if (box_srcntgt_counts_cumul[i] == 0) box_srcntgt_counts_nonchild[i] = 0;
        else box_srcntgt_counts_nonchild[i] = box_morton_bin_counts[i * w];
    }
}

/* find_prune_indices scan -- tree_build_kernels.py:1697-1718 */
void orc_find_prune_indices(int64_t nboxes, const particle_id_t *box_srcntgt_counts_cumul,
    box_id_t *src_box_id, box_id_t *dst_box_id, box_id_t *nboxes_post_prune)
{
    box_id_t item = 0;
    for (int64_t i = 0; i < nboxes; ++i) {
        item += (box_srcntgt_counts_cumul[i] != 0);
        if (box_srcntgt_counts_cumul[i]) { dst_box_id[i] = item - 1; src_box_id[item - 1] = (box_id_t)i; }
        if (i + 1 == nboxes) *nboxes_post_prune = item;
    }
}

/* find_level_box_counts scan -- tree_build_kernels.py:1724-1742 */
void orc_find_level_box_counts(int64_t nboxes, const box_level_t *box_levels, box_id_t *level_box_counts)
{
    box_id_t item = 0;
    for (int64_t i = 0; i < nboxes; ++i) {
        int seg_start = (i == 0) || (box_levels[i] != box_levels[i - 1]);
        item = seg_start ? 1 : item + 1;
        if (i + 1 == nboxes || box_levels[i] != box_levels[i + 1])
            level_box_counts[box_levels[i]] = item;
    }
}

/* source_counter scan (exclusive output) -- tree_build_kernels.py:1770-1782 */
void orc_source_counter(int64_t n, const particle_id_t *user_srcntgt_ids, particle_id_t nsources,
    particle_id_t *source_numbers)
{
    particle_id_t prev = 0;
    for (int64_t i = 0; i < n; ++i) {
        source_numbers[i] = prev;
        prev += (user_srcntgt_ids[i] < nsources) ? 1 : 0;
    }
}

/* find_source_and_target_indices -- tree_build_kernels.py:1013-1164 */
void orc_source_and_target_index_finder(
    int have_extent, int64_t n,
    const particle_id_t *user_srcntgt_ids, particle_id_t nsources,
    const box_id_t *srcntgt_box_ids, const box_id_t *box_parent_ids,
    const particle_id_t *box_srcntgt_starts, const particle_id_t *box_srcntgt_counts_cumul,
    const particle_id_t *source_numbers, const particle_id_t *box_srcntgt_counts_nonchild,
    particle_id_t *user_source_ids, particle_id_t *srcntgt_target_ids,
    particle_id_t *sorted_target_ids,
    particle_id_t *box_source_starts, particle_id_t *box_source_counts_cumul,
    particle_id_t *box_target_starts, particle_id_t *box_target_counts_cumul,
    particle_id_t *box_source_counts_nonchild, particle_id_t *box_target_counts_nonchild)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        particle_id_t sorted_id = (particle_id_t)i;
        particle_id_t source_nr = source_numbers[i];
        particle_id_t target_nr = sorted_id - source_nr;
        box_id_t box_id = srcntgt_box_ids[i];
        particle_id_t box_start = box_srcntgt_starts[box_id];
        particle_id_t box_count = box_srcntgt_counts_cumul[box_id];
        particle_id_t uid = user_srcntgt_ids[i];
        int is_source = uid < nsources;
        {
            particle_id_t wstart = box_start; box_id_t wbox = box_id;
            while (sorted_id == wstart) {
                box_source_starts[wbox] = source_nr;
                box_target_starts[wbox] = target_nr;
                box_id_t nbx = box_parent_ids[wbox];
                if (nbx == wbox) break;
                wbox = nbx; wstart = box_srcntgt_starts[wbox];
            }
        }
        if (have_extent) {
            particle_id_t nonchild = box_srcntgt_counts_nonchild[box_id];
            if (sorted_id + 1 == box_start + nonchild) {
                particle_id_t s0 = source_numbers[box_start];
                particle_id_t t0 = box_start - s0;
                box_source_counts_nonchild[box_id] = source_nr + (particle_id_t)is_source - s0;
                box_target_counts_nonchild[box_id] = target_nr + 1 - (particle_id_t)is_source - t0;
            }
        }
        {
            particle_id_t wstart = box_start, wcount = box_count; box_id_t wbox = box_id;
            while (sorted_id + 1 == wstart + wcount) {
                particle_id_t s0 = source_numbers[wstart];
                particle_id_t t0 = wstart - s0;
                box_source_counts_cumul[wbox] = source_nr + (particle_id_t)is_source - s0;
                box_target_counts_cumul[wbox] = target_nr + 1 - (particle_id_t)is_source - t0;
                box_id_t nbx = box_parent_ids[wbox];
                if (nbx == wbox) break;
                wbox = nbx;
                wstart = box_srcntgt_starts[wbox]; wcount = box_srcntgt_counts_cumul[wbox];
            }
        }
        if (is_source) user_source_ids[source_nr] = uid;
        else {
            srcntgt_target_ids[target_nr] = uid;
            sorted_target_ids[uid - nsources] = target_nr;
        }
    }
}

/* box_info kernel -- tree_build_kernels.py:1192-1305; flag bits tree.py:133-142 */
#define BOX_IS_SOURCE_BOX 1
#define BOX_IS_TARGET_BOX 2
#define BOX_HAS_SOURCE_CHILD_BOXES 4
#define BOX_HAS_TARGET_CHILD_BOXES 8

void orc_box_info(
    int have_extent, int sources_are_targets, int64_t nboxes,
    const particle_id_t *box_srcntgt_counts_cumul,
    const particle_id_t *box_source_counts_cumul, const particle_id_t *box_target_counts_cumul,
    const int32_t *box_has_children,
    particle_id_t *box_source_counts_nonchild, particle_id_t *box_target_counts_nonchild,
    box_flags_t *box_flags)
{
    for (int64_t b = 0; b < nboxes; ++b) {
        particle_id_t particle_count = box_srcntgt_counts_cumul[b];
        particle_id_t nc_src = have_extent ? box_source_counts_nonchild[b] : 0;
        particle_id_t nc_tgt = have_extent ? box_target_counts_nonchild[b] : 0;
        particle_id_t nc = nc_src + nc_tgt;
        box_flags_t fl = 0;
        if (box_has_children[b]) {
            /* :1256 -- BOX_HAS_SOURCE_OR_TARGET_CHILD_BOXES (= both child bits) is set
               unconditionally for every non-leaf box */
            fl |= BOX_HAS_SOURCE_CHILD_BOXES | BOX_HAS_TARGET_CHILD_BOXES;
            if (sources_are_targets) {
                if (particle_count - nc)
                    fl |= BOX_HAS_SOURCE_CHILD_BOXES | BOX_HAS_TARGET_CHILD_BOXES;
            } else {
                if (box_source_counts_cumul[b] - nc_src) fl |= BOX_HAS_SOURCE_CHILD_BOXES;
                if (box_target_counts_cumul[b] - nc_tgt) fl |= BOX_HAS_TARGET_CHILD_BOXES;
            }
            if (nc_src) fl |= BOX_IS_SOURCE_BOX;
            if (nc_tgt) fl |= BOX_IS_TARGET_BOX;
        } else {
            if (sources_are_targets) {
                if (particle_count) fl |= BOX_IS_SOURCE_BOX | BOX_IS_TARGET_BOX;
                box_source_counts_nonchild[b] = particle_count;
            } else {
                particle_id_t ms = box_source_counts_cumul[b];
                particle_id_t mt = particle_count - ms;
                if (ms) fl |= BOX_IS_SOURCE_BOX;
                if (mt) fl |= BOX_IS_TARGET_BOX;
                box_source_counts_nonchild[b] = ms;
                box_target_counts_nonchild[b] = mt;
            }
        }
        box_flags[b] = fl;
    }
}

/* find_box_extents -- tree_build_kernels.py:1311-1399; tree_build.py:1751-1802
 * one launch covers boxes [start, stop) of one level. */
void orc_box_extents(
    int d, int have_extent, int64_t start, int64_t stop, int64_t aligned_nboxes,
    const box_id_t *box_child_ids /* [2^d, aligned] */, const coord_t *box_centers /* [d, aligned] */,
    const particle_id_t *box_particle_starts, const particle_id_t *box_particle_counts_nonchild,
    const coord_t *const *particles, const coord_t *particle_radii, int enable_radii,
    coord_t *bbox_min /* [d, aligned] */, coord_t *bbox_max)
{
    const int nb = 1 << d;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ibox = start; ibox < stop; ++ibox) {
        coord_t mn[MAXDIM], mx[MAXDIM];
        for (int a = 0; a < d; ++a) mn[a] = mx[a] = box_centers[a * aligned_nboxes + ibox];
        particle_id_t p0 = box_particle_starts[ibox];
        particle_id_t p1 = p0 + box_particle_counts_nonchild[ibox];
        for (particle_id_t ip = p0; ip < p1; ++ip) {
            coord_t rad = 0;
            if (have_extent && enable_radii) rad = particle_radii[ip];
            for (int a = 0; a < d; ++a) {
                coord_t c = particles[a][ip];
                coord_t lo = c - rad, hi = c + rad;
                mn[a] = (lo < mn[a]) ? lo : mn[a];
                mx[a] = (mx[a] < hi) ? hi : mx[a];
            }
        }
        for (int m = 0; m < nb; ++m) {
            box_id_t child = box_child_ids[m * aligned_nboxes + ibox];
            if (child == 0) continue;
            for (int a = 0; a < d; ++a) {
                coord_t cmn = bbox_min[a * aligned_nboxes + child];
                coord_t cmx = bbox_max[a * aligned_nboxes + child];
                mn[a] = (cmn < mn[a]) ? cmn : mn[a];
                mx[a] = (mx[a] < cmx) ? cmx : mx[a];
            }
        }
        for (int a = 0; a < d; ++a) {
            bbox_min[a * aligned_nboxes + ibox] = mn[a];
            bbox_max[a * aligned_nboxes + ibox] = mx[a];
        }
    }
}
