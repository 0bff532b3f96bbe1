/*
 * oracle/oracle_trav.c -- CPU restatement of the reference FMM-traversal kernels.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_tree.c).  Restated from the kernel sources
 * in /root/reference/boxtree/traversal.py; each function cites the lines it
 * follows.  PARITY PINNED against the reference itself: tests/refexec executes
 * the reference's FMMTraversalBuilder (host code and kernel templates unmodified)
 * on the CPU and this restatement reproduces every output array, dtype and byte
 * (tests/test_refexec.py; 172 sweep cases, random cases, BASELINE configs up to 1e7 points).  The list-of-lists builder mirrors the documented
 * behaviour of pyopencl.algorithm.ListOfListsBuilder (pyopencl is an unpinned
 * third-party dependency of the reference, `pyopencl>=2022.1`, not vendored):
 * one work-item per row, a count pass, an exclusive scan to `starts`, a write
 * pass; row content is in APPEND order of the row's work-item.
 *
 * This file is #included by oracle_tree.c's translation unit settings: it is
 * compiled together with it (same coord_t).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#if defined(COORD_F32)
typedef float coord_t;
#define COORD_SQRT sqrtf
#define COORD_FABS fabsf
#define COORD_FMAX fmaxf
#define COORD_FMIN fminf
#define COORD_EPS FLT_EPSILON
#elif defined(COORD_F64)
typedef double coord_t;
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
typedef uint8_t box_level_t;
typedef uint8_t box_flags_t;

#define MAXDIM 3
#define BOX_IS_SOURCE_BOX 1
#define BOX_IS_TARGET_BOX 2
#define BOX_HAS_SOURCE_CHILD_BOXES 4
#define BOX_HAS_TARGET_CHILD_BOXES 8

static inline int ocl_shl_i(int v, int s) { return (int)((unsigned)v << (s & 31)); }

/* LEVEL_TO_RAD -- traversal.py:234-235 */
static inline coord_t level_to_rad(coord_t root_extent, int level)
{ return (root_extent * 1 / (coord_t)ocl_shl_i(1, level + 1)); }

/* traversal.py:279-305 */
static inline int adj_nbhd(int d, coord_t root_extent, const coord_t *tc, int tl,
                           coord_t nbhd, const coord_t *sc, int sl)
{
    coord_t target_rad = level_to_rad(root_extent, tl);
    coord_t source_rad = level_to_rad(root_extent, sl);
    coord_t rad_sum = ((2 * (nbhd - 1) + 1) * target_rad + source_rad);
    coord_t slack = rad_sum + COORD_FMIN(target_rad, source_rad);
    coord_t l_inf_dist = 0;
    for (int a = 0; a < d; ++a)
        l_inf_dist = COORD_FMAX(l_inf_dist, COORD_FABS(tc[a] - sc[a]));
    return l_inf_dist <= slack;
}
/* traversal.py:307-318 */
static inline int adj(int d, coord_t root_extent, const coord_t *tc, int tl, const coord_t *sc, int sl)
{ return adj_nbhd(d, root_extent, tc, tl, (coord_t)1, sc, sl); }

/* ---- shared tree view --------------------------------------------------- */
typedef struct {
    int d;
    int64_t aligned_nboxes;
    coord_t root_extent;
    const coord_t *box_centers;       /* [d, aligned] */
    const box_level_t *box_levels;    /* [nboxes] */
    const box_id_t *box_child_ids;    /* [2^d, aligned] */
    const box_flags_t *box_flags;     /* [nboxes] */
    const box_id_t *box_parent_ids;
    int well_sep_is_n_away;
} tree_view_t;

static inline void load_center(const tree_view_t *t, box_id_t b, coord_t *c)
{ for (int a = 0; a < t->d; ++a) c[a] = t->box_centers[t->aligned_nboxes * a + b]; }

/* ---- emitter: count pass / write pass ----------------------------------- */
typedef struct {
    int nlists;
    int write;                 /* 0: count, 1: write */
    int64_t row;
    int32_t *counts[3];        /* per list: [nrows] (count pass) */
    const int32_t *starts[3];  /* per list: [nrows+1] (write pass) */
    int32_t *lists[3];         /* per list: output (write pass) */
    int32_t cursor[3];
} emitter_t;

static inline void emit(emitter_t *e, int ilist, box_id_t v)
{
    if (!e->write) { if (e->counts[ilist]) e->counts[ilist][e->row]++; }
    else if (e->lists[ilist]) {
        e->lists[ilist][e->starts[ilist][e->row] + e->cursor[ilist]] = v;
        e->cursor[ilist]++;
    }
}

/* walk state -- traversal.py:98-160 */
#define WALK_MAXLEV 128
typedef struct {
    box_id_t box_stack[WALK_MAXLEV]; int mnr_stack[WALK_MAXLEV];
    int stack_size; box_id_t parent; int mnr; int cont;
} walk_t;
static inline void walk_init(walk_t *w, box_id_t start) { w->stack_size = 0; w->parent = start; w->mnr = 0; w->cont = 1; }
static inline box_id_t walk_box(const tree_view_t *t, const walk_t *w)
{ return t->box_child_ids[w->mnr * t->aligned_nboxes + w->parent]; }
static inline void walk_advance(walk_t *w, int nb)
{
    for (;;) {
        ++w->mnr;
        if (w->mnr < nb) break;
        w->cont = (w->stack_size > 0);
        if (w->cont) { --w->stack_size; w->parent = w->box_stack[w->stack_size]; w->mnr = w->mnr_stack[w->stack_size]; }
        else break;
    }
}
static inline void walk_push(walk_t *w, box_id_t nb)
{ w->box_stack[w->stack_size] = w->parent; w->mnr_stack[w->stack_size] = w->mnr; ++w->stack_size; w->parent = nb; w->mnr = 0; }

/* ------------------------------------------------------------------------
 * b1: sources_parents_and_targets -- traversal.py:326-355, 1853-1866
 * lists: 0 source_parent_boxes, 1 source_boxes, 2 target_or_target_parent_boxes,
 *        3 target_boxes (only if !sources_are_targets).
 * counts_out[4]; outputs may be NULL on the count call.
 * ---------------------------------------------------------------------- */
void orc_sources_parents_and_targets(
    int64_t nboxes, const box_flags_t *box_flags, int sources_are_targets,
    const int8_t *source_boxes_mask, const int8_t *source_parent_boxes_mask,
    box_id_t *source_parent_boxes, box_id_t *source_boxes,
    box_id_t *target_or_target_parent_boxes, box_id_t *target_boxes, int64_t *counts_out)
{
    int64_t nsp = 0, ns = 0, ntp = 0, nt = 0;
    for (int64_t b = 0; b < nboxes; ++b) {
        box_flags_t fl = box_flags[b];
        if ((fl & BOX_IS_SOURCE_BOX) && (!source_boxes_mask || source_boxes_mask[b])) {
            if (source_boxes) source_boxes[ns] = (box_id_t)b;
            ++ns;
        }
        if ((fl & BOX_HAS_SOURCE_CHILD_BOXES) && (!source_parent_boxes_mask || source_parent_boxes_mask[b])) {
            if (source_parent_boxes) source_parent_boxes[nsp] = (box_id_t)b;
            ++nsp;
        }
        if (!sources_are_targets && (fl & BOX_IS_TARGET_BOX)) {
            if (target_boxes) target_boxes[nt] = (box_id_t)b;
            ++nt;
        }
        if (fl & (BOX_HAS_TARGET_CHILD_BOXES | BOX_IS_TARGET_BOX)) {
            if (target_or_target_parent_boxes) target_or_target_parent_boxes[ntp] = (box_id_t)b;
            ++ntp;
        }
    }
    counts_out[0] = nsp; counts_out[1] = ns; counts_out[2] = ntp; counts_out[3] = nt;
}

/* b2: extract_level_start_box_nrs -- traversal.py:361-392 (host fix-up is in the driver) */
void orc_extract_level_start_box_nrs(
    int64_t nlist, const box_id_t *level_start_box_nrs, const box_level_t *box_levels,
    const box_id_t *box_list, box_id_t *list_level_start_box_nrs)
{
    for (int64_t i = 0; i < nlist; ++i) {
        box_id_t my_box = box_list[i];
        int my_level = box_levels[my_box];
        int leading;
        if (i == 0) leading = 1;
        else {
            box_id_t prev = box_list[i - 1];
            box_id_t my_level_start = level_start_box_nrs[my_level];
            leading = (prev < my_level_start && my_level_start <= my_box);
        }
        if (leading) list_level_start_box_nrs[my_level] = (box_id_t)i;
    }
}

/* ---- generate() functions ------------------------------------------------ */

/* b3: same_level_non_well_sep_boxes -- traversal.py:398-464 */
static void gen_colleagues(const tree_view_t *t, emitter_t *e, box_id_t box_id)
{
    const int nb = 1 << t->d;
    coord_t center[MAXDIM]; load_center(t, box_id, center);
    if (box_id == 0) return;
    int level = t->box_levels[box_id];
    walk_t w; walk_init(&w, 0);
    while (w.cont) {
        box_id_t wb = walk_box(t, &w);
        if (wb) {
            coord_t wc[MAXDIM]; load_center(t, wb, wc);
            int a_or_o = adj_nbhd(t->d, t->root_extent, center, level,
                                  (coord_t)t->well_sep_is_n_away, wc, t->box_levels[wb]);
            if (a_or_o) {
                if (w.stack_size + 1 == level && wb != box_id) emit(e, 0, wb);
                else { walk_push(&w, wb); continue; }
            }
        }
        walk_advance(&w, nb);
    }
}

/* N3: peer lists -- area_query.py:393-475 (PEER_LIST_FINDER_TEMPLATE) */
static void gen_peers(const tree_view_t *t, emitter_t *e, box_id_t box_id)
{
    const int nb = 1 << t->d;
    coord_t center[MAXDIM]; load_center(t, box_id, center);
    if (box_id == 0) { emit(e, 0, box_id); return; }           /* peer of root = self */
    int level = t->box_levels[box_id];
    walk_t w; walk_init(&w, 0);
    while (w.cont) {
        box_id_t wb = walk_box(t, &w);
        if (wb) {
            coord_t wc[MAXDIM]; load_center(t, wb, wc);
            /* walk box lives on level stack_size + 1 */
            int a_or_o = adj(t->d, t->root_extent, center, level, wc, w.stack_size + 1);
            if (a_or_o) {
                if (w.stack_size + 1 == level) emit(e, 0, wb);
                else if (!(t->box_flags[wb] & (BOX_HAS_SOURCE_CHILD_BOXES | BOX_HAS_TARGET_CHILD_BOXES)))
                    emit(e, 0, wb);
                else {
                    int must_be_peer = 1;
                    for (int m = 0; must_be_peer && m < nb; ++m) {
                        box_id_t c = t->box_child_ids[m * t->aligned_nboxes + wb];
                        if (c) {
                            coord_t cc[MAXDIM]; load_center(t, c, cc);
                            must_be_peer &= !adj(t->d, t->root_extent, center, level, cc, w.stack_size + 2);
                        }
                    }
                    if (must_be_peer) emit(e, 0, wb);
                    else { walk_push(&w, wb); continue; }
                }
            }
        }
        walk_advance(&w, nb);
    }
}

/* b4: neighbor_source_boxes (list 1) -- traversal.py:470-550 */
typedef struct { const box_id_t *target_boxes; } list1_args_t;
static void gen_list1(const tree_view_t *t, emitter_t *e, const list1_args_t *x, box_id_t target_box_number)
{
    const int nb = 1 << t->d;
    box_id_t box_id = x->target_boxes[target_box_number];
    coord_t center[MAXDIM]; load_center(t, box_id, center);
    int level = t->box_levels[box_id];
    if (t->box_flags[0] & BOX_IS_SOURCE_BOX) emit(e, 0, 0);
    walk_t w; walk_init(&w, 0);
    while (w.cont) {
        box_id_t wb = walk_box(t, &w);
        if (wb) {
            coord_t wc[MAXDIM]; load_center(t, wb, wc);
            int a_or_o = adj(t->d, t->root_extent, center, level, wc, t->box_levels[wb]);
            if (a_or_o) {
                box_flags_t fl = t->box_flags[wb];
                if (fl & BOX_IS_SOURCE_BOX) emit(e, 0, wb);
                if (fl & BOX_HAS_SOURCE_CHILD_BOXES) { walk_push(&w, wb); continue; }
            }
        }
        walk_advance(&w, nb);
    }
}

/* b5: from_sep_siblings (list 2) -- traversal.py:556-601 */
typedef struct {
    const box_id_t *tp_boxes; const box_id_t *coll_starts; const box_id_t *coll_lists;
} list2_args_t;
static void gen_list2(const tree_view_t *t, emitter_t *e, const list2_args_t *x, box_id_t itp)
{
    const int nb = 1 << t->d;
    box_id_t box_id = x->tp_boxes[itp];
    coord_t center[MAXDIM]; load_center(t, box_id, center);
    int level = t->box_levels[box_id];
    box_id_t parent = t->box_parent_ids[box_id];
    if (parent == box_id) return;
    for (box_id_t i = x->coll_starts[parent]; i < x->coll_starts[parent + 1]; ++i) {
        box_id_t parent_nf = x->coll_lists[i];
        for (int m = 0; m < nb; ++m) {
            box_id_t sib = t->box_child_ids[m * t->aligned_nboxes + parent_nf];
            if (sib == 0) continue;
            coord_t sc[MAXDIM]; load_center(t, sib, sc);
            int sep = !adj_nbhd(t->d, t->root_extent, center, level,
                                (coord_t)t->well_sep_is_n_away, sc, t->box_levels[sib]);
            if (sep) emit(e, 0, sib);
        }
    }
}

/* b6: from_sep_smaller (list 3 per level / list 3 close) -- traversal.py:607-875
 * crit: 0 static_linf, 1 precise_linf, 2 static_l2.  list 0 = from_sep_smaller,
 * list 1 = from_sep_close_smaller */
typedef struct {
    coord_t stick_out_factor; const box_id_t *target_boxes;
    const box_id_t *coll_starts; const box_id_t *coll_lists;
    int targets_have_extent; int sources_have_extent; int crit;
    const coord_t *box_target_bounding_box_min, *box_target_bounding_box_max;
    const particle_id_t *box_source_counts_cumul;
    particle_id_t min_nsources_cumul; int source_level;
} list3_args_t;

static void gen_list3(const tree_view_t *t, emitter_t *e, const list3_args_t *x, box_id_t target_box_number)
{
    const int nb = 1 << t->d, d = t->d;
    box_id_t tgt_box_id = x->target_boxes[target_box_number];
    coord_t tgt_center[MAXDIM]; load_center(t, tgt_box_id, tgt_center);
    int tgt_level = t->box_levels[tgt_box_id];
    coord_t tgt_stickout_l_inf_rad = 0, tgt_ext_center[MAXDIM], tgt_radii_vec[MAXDIM];
    if (x->targets_have_extent) {
        if (x->crit == 0 || x->crit == 2)
            tgt_stickout_l_inf_rad = (1 + x->stick_out_factor) * level_to_rad(t->root_extent, tgt_level);
        else { /* load_true_box_extent, traversal.py:177-198 */
            for (int a = 0; a < d; ++a) {
                coord_t mn = x->box_target_bounding_box_min[a * t->aligned_nboxes + tgt_box_id];
                coord_t mx = x->box_target_bounding_box_max[a * t->aligned_nboxes + tgt_box_id];
                tgt_ext_center[a] = ((coord_t)0.5) * (mn + mx);
                tgt_radii_vec[a] = ((coord_t)0.5) * (mx - mn);
            }
        }
    }
    const int close_lists_exist = x->sources_have_extent || x->targets_have_extent;
    for (box_id_t i = x->coll_starts[tgt_box_id]; i < x->coll_starts[tgt_box_id + 1]; ++i) {
        box_id_t same_lev_nws_box = x->coll_lists[i];
        if (same_lev_nws_box == tgt_box_id) continue;
        walk_t w; walk_init(&w, same_lev_nws_box);
        while (w.cont) {
            box_id_t wb = walk_box(t, &w);
            box_flags_t cfl = t->box_flags[wb];
            if (wb && (cfl & (BOX_IS_SOURCE_BOX | BOX_HAS_SOURCE_CHILD_BOXES))) {
                coord_t wc[MAXDIM]; load_center(t, wb, wc);
                int walk_level = t->box_levels[wb];
                int in_list_1 = adj(d, t->root_extent, tgt_center, tgt_level, wc, walk_level);
                if (in_list_1) {
                    if (cfl & BOX_HAS_SOURCE_CHILD_BOXES) {
                        if (walk_level <= x->source_level || x->source_level == -1) {
                            walk_push(&w, wb); continue;
                        }
                    }
                } else {
                    int meets_sep_crit;
                    if (!x->targets_have_extent) meets_sep_crit = 1;
                    else if (x->crit == 0) {
                        coord_t source_rad = level_to_rad(t->root_extent, walk_level);
                        coord_t l_inf_dist = 0;
                        for (int a = 0; a < d; ++a)
                            l_inf_dist = COORD_FMAX(l_inf_dist,
                                COORD_FABS(tgt_center[a] - wc[a]) - tgt_stickout_l_inf_rad - source_rad);
                        meets_sep_crit = l_inf_dist >= (2 - 8 * COORD_EPS) * source_rad;
                    } else if (x->crit == 1) {
                        coord_t source_rad = level_to_rad(t->root_extent, walk_level);
                        coord_t l_inf_dist = 0;
                        for (int a = 0; a < d; ++a)
                            l_inf_dist = COORD_FMAX(l_inf_dist,
                                COORD_FABS(tgt_ext_center[a] - wc[a]) - tgt_radii_vec[a] - source_rad);
                        meets_sep_crit = l_inf_dist >= (2 - 8 * COORD_EPS) * source_rad;
                    } else {
                        coord_t source_l_inf_rad = level_to_rad(t->root_extent, walk_level);
                        coord_t l2sq = 0;
                        for (int a = 0; a < d; ++a)
                            l2sq = l2sq + (tgt_center[a] - wc[a]) * (tgt_center[a] - wc[a]);
                        coord_t rhs = COORD_SQRT(l2sq)
                            - COORD_SQRT((coord_t)d) * tgt_stickout_l_inf_rad - source_l_inf_rad;
                        meets_sep_crit = ((2 - 8 * COORD_EPS) * source_l_inf_rad <= rhs);
                    }
                    int force_close = close_lists_exist &&
                        (x->box_source_counts_cumul[wb] < x->min_nsources_cumul);
                    if (meets_sep_crit && !force_close) {
                        if (x->source_level == walk_level) emit(e, 0, wb);
                    } else if (close_lists_exist) {
                        if ((cfl & BOX_IS_SOURCE_BOX) && x->source_level == -1) emit(e, 1, wb);
                        if (cfl & BOX_HAS_SOURCE_CHILD_BOXES) { walk_push(&w, wb); continue; }
                    }
                }
            }
            walk_advance(&w, nb);
        }
    }
}

/* b7: from_sep_bigger (list 4 / list 4 close) -- traversal.py:931-1146
 * list 0 = from_sep_bigger, list 1 = from_sep_close_bigger */
typedef struct {
    coord_t stick_out_factor; const box_id_t *tp_boxes;
    const box_id_t *coll_starts; const box_id_t *coll_lists; int with_extent;
} list4_args_t;

static int meets_sep_bigger_criterion(int d, coord_t root_extent, const coord_t *tc, int tl,
                                      const coord_t *sc, int sl, coord_t stick_out_factor)
{ /* traversal.py:933-972 */
    coord_t target_rad = level_to_rad(root_extent, tl);
    coord_t source_rad = level_to_rad(root_extent, sl);
    coord_t max_allowed = (3 * (1 + stick_out_factor) * target_rad + source_rad);
    coord_t l_inf_dist = 0;
    for (int a = 0; a < d; ++a)
        l_inf_dist = COORD_FMAX(l_inf_dist, COORD_FABS(tc[a] - sc[a]));
    return l_inf_dist >= max_allowed * (1 - 8 * COORD_EPS);
}

static void gen_list4(const tree_view_t *t, emitter_t *e, const list4_args_t *x, box_id_t itp)
{
    const int d = t->d;
    box_id_t tgt_ibox = x->tp_boxes[itp];
    coord_t tgt_center[MAXDIM]; load_center(t, tgt_ibox, tgt_center);
    int tgt_box_level = t->box_levels[tgt_ibox];
    if (tgt_box_level == 0) return;
    box_id_t tgt_parent = t->box_parent_ids[tgt_ibox];
    const int tgt_parent_level = tgt_box_level - 1;
    coord_t parent_center[MAXDIM]; load_center(t, tgt_parent, parent_center);
    box_flags_t tgt_box_flags = t->box_flags[tgt_ibox];
    int walk_level; box_id_t cur;
    if (t->well_sep_is_n_away == 1) { walk_level = tgt_box_level - 1; cur = tgt_parent; }
    else { walk_level = tgt_box_level; cur = tgt_ibox; }
    for (; walk_level != 0; --walk_level, cur = t->box_parent_ids[cur]) {
        for (box_id_t i = x->coll_starts[cur]; i < x->coll_starts[cur + 1]; ++i) {
            box_id_t sb = x->coll_lists[i];
            if (!(t->box_flags[sb] & BOX_IS_SOURCE_BOX)) continue;
            coord_t sc[MAXDIM]; load_center(t, sb, sc);
            int in_list_1 = adj(d, t->root_extent, tgt_center, tgt_box_level, sc, walk_level);
            if (in_list_1) continue;
            if (x->with_extent) {
                int tgt_meets = meets_sep_bigger_criterion(d, t->root_extent, tgt_center,
                    tgt_box_level, sc, walk_level, x->stick_out_factor);
                if (!tgt_meets) {
                    if (tgt_box_flags & BOX_IS_TARGET_BOX) emit(e, 1, sb);
                    continue;
                }
            }
            int in_parent_list_1 = adj(d, t->root_extent, parent_center, tgt_parent_level, sc, walk_level);
            int would_be_in_parent_list_4 = !in_parent_list_1;
            if (t->well_sep_is_n_away > 1)
                would_be_in_parent_list_4 = would_be_in_parent_list_4 && (walk_level < tgt_box_level);
            if (would_be_in_parent_list_4) {
                if (x->with_extent) {
                    int parent_meets = meets_sep_bigger_criterion(d, t->root_extent, parent_center,
                        tgt_parent_level, sc, walk_level, x->stick_out_factor);
                    if (!parent_meets) emit(e, 0, sb);
                }
            } else emit(e, 0, sb);
        }
    }
}

/* ---- list-of-lists drivers ----------------------------------------------- *
 * kind: 0 colleagues, 1 list1, 2 list2, 3 list3, 4 list4.
 * Pass write=0 with counts[k] (zero-initialised, [nrows]) to count, then
 * write=1 with starts[k] ([nrows+1]) and lists[k].  NULL entries = omitted list.
 */
typedef struct {
    tree_view_t tree;
    list1_args_t l1; list2_args_t l2; list3_args_t l3; list4_args_t l4;
} trav_args_t;

void orc_build_lists(int kind, const trav_args_t *A, int64_t nrows, int write,
                     int32_t *counts0, int32_t *counts1,
                     const int32_t *starts0, const int32_t *starts1,
                     int32_t *lists0, int32_t *lists1)
{
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < nrows; ++r) {
        emitter_t e; memset(&e, 0, sizeof e);
        e.nlists = 2; e.write = write; e.row = r;
        e.counts[0] = counts0; e.counts[1] = counts1;
        e.starts[0] = starts0; e.starts[1] = starts1;
        e.lists[0] = lists0; e.lists[1] = lists1;
        switch (kind) {
        case 0: gen_colleagues(&A->tree, &e, (box_id_t)r); break;
        case 1: gen_list1(&A->tree, &e, &A->l1, (box_id_t)r); break;
        case 2: gen_list2(&A->tree, &e, &A->l2, (box_id_t)r); break;
        case 3: gen_list3(&A->tree, &e, &A->l3, (box_id_t)r); break;
        case 4: gen_list4(&A->tree, &e, &A->l4, (box_id_t)r); break;
        case 5: gen_peers(&A->tree, &e, (box_id_t)r); break;
        }
    }
}

/* N3: area query -- area_query.py:168-392 (GUIDING_BOX_FINDER_MACRO, AREA_QUERY_WALKER_BODY)
 * and check_l_infty_ball_overlap (traversal.py:200-214).  One row per ball: the leaves that
 * overlap the l^inf ball.  write = 0: counts[nballs]; write = 1: lists at starts. */
static int ball_overlaps(const tree_view_t *t, box_id_t b, const coord_t *bc, coord_t br)
{
    coord_t c[MAXDIM]; load_center(t, b, c);
    coord_t size_sum = level_to_rad(t->root_extent, t->box_levels[b]) + br;
    coord_t max_dist = 0;
    for (int a = 0; a < t->d; ++a) max_dist = COORD_FMAX(max_dist, COORD_FABS(bc[a] - c[a]));
    return max_dist <= size_sum;
}

void orc_area_query(const tree_view_t *t, const int32_t *peer_starts, const box_id_t *peer_lists,
                    int64_t nballs, const coord_t *const *ball_centers, const coord_t *ball_radii,
                    const coord_t *bbox_min, int write, int32_t *counts, const int32_t *starts,
                    int32_t *lists)
{
    const int d = t->d, nb = 1 << d;
    const int haschild = BOX_HAS_SOURCE_CHILD_BOXES | BOX_HAS_TARGET_CHILD_BOXES;
    for (int64_t i = 0; i < nballs; ++i) {
        coord_t bc[MAXDIM]; for (int a = 0; a < d; ++a) bc[a] = ball_centers[a][i];
        const coord_t br = ball_radii[i];
        /* find_guiding_box, :168-264 */
        box_id_t box = 0;
        coord_t bbox_max[MAXDIM], qc[MAXDIM];
        for (int a = 0; a < d; ++a) {
            bbox_max[a] = bbox_min[a] + (coord_t)(t->root_extent / (1 + 1e-4));
            qc[a] = COORD_FMIN(bbox_max[a], COORD_FMAX(bbox_min[a], bc[a]));
        }
        coord_t qr = 0;
        for (int mnr = 0; mnr < nb; ++mnr) {
            for (int a = 0; a < d; ++a) {
                coord_t off = ((1 << (d - 1 - a)) & mnr) ? +br : -br;
                coord_t corner = COORD_FMIN(bbox_max[a], COORD_FMAX(bbox_min[a], bc[a] + off));
                qr = COORD_FMAX(qr, COORD_FABS(corner - qc[a]));
            }
        }
        if (level_to_rad(t->root_extent, 0) / 2 >= qr) {
            for (unsigned box_level = 0;; ++box_level) {
                if (!(t->box_flags[box] & haschild)
                    || (level_to_rad(t->root_extent, box_level) / 2 < qr
                        && qr <= level_to_rad(t->root_extent, box_level)))
                    break;
                int morton = 0;
                for (int a = 0; a < d; ++a) {
                    coord_t off_scaled = (qc[a] - bbox_min[a]) / t->root_extent;
                    unsigned bits = (unsigned)(off_scaled * (1U << ((1 + box_level) & 31)));
                    morton |= (bits & 1U) << (d - 1 - a);
                }
                box_id_t next = t->box_child_ids[morton * t->aligned_nboxes + box];
                if (next) box = next; else break;
            }
        }
        /* walk the peers, :266-365 */
        int32_t n = 0;
        int32_t *out = write ? lists + starts[i] : NULL;
        for (int32_t pi = peer_starts[box]; pi < peer_starts[box + 1]; ++pi) {
            box_id_t peer = peer_lists[pi];
            if (!(t->box_flags[peer] & haschild)) {
                if (ball_overlaps(t, peer, bc, br)) { if (write) out[n] = peer; ++n; }
            } else {
                walk_t w; walk_init(&w, peer);
                while (w.cont) {
                    box_id_t wb = walk_box(t, &w);
                    if (wb) {
                        if (!(t->box_flags[wb] & haschild)) {
                            if (ball_overlaps(t, wb, bc, br)) { if (write) out[n] = wb; ++n; }
                        } else { walk_push(&w, wb); continue; }
                    }
                    walk_advance(&w, nb);
                }
            }
        }
        if (!write) counts[i] = n;
    }
}

/* list merger -- traversal.py:1153-1214 (count kernel then write kernel) */
void orc_merge_lists_count(int64_t noutput, const box_id_t *output_to_input_box, int nlists,
                           const box_id_t *const *starts, box_id_t *new_counts /* [noutput+1] */)
{
    for (int64_t i = 0; i < noutput; ++i) {
        box_id_t ibox = output_to_input_box[i];
        box_id_t tot = 0;
        for (int l = 0; l < nlists; ++l) tot += starts[l][ibox + 1] - starts[l][ibox];
        if (i == 0) new_counts[0] = 0;
        new_counts[i + 1] = tot;
    }
}
void orc_merge_lists_write(int64_t noutput, const box_id_t *output_to_input_box, int nlists,
                           const box_id_t *const *starts, const box_id_t *const *lists,
                           const box_id_t *new_starts, box_id_t *new_lists)
{
    for (int64_t i = 0; i < noutput; ++i) {
        box_id_t ibox = output_to_input_box[i];
        box_id_t cur = new_starts[i];
        for (int l = 0; l < nlists; ++l) {
            box_id_t s = starts[l][ibox], c = starts[l][ibox + 1] - s;
            for (box_id_t j = 0; j < c; ++j) new_lists[cur++] = lists[l][s + j];
        }
    }
}
