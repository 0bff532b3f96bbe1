"""``FMMCostModel`` (``boxtree/cost.py:87-713, 715-1262``): modelled cost of every FMM stage per
box from a traversal -- the weights the reference hands to ``partition_work``
(``distributed/__init__.py:212-230``).  Per-box work is CSR segmented sums in torch (float64) on
the array context's stream; the translation cost model keeps the reference's symbolic structure
(``var("c_m2l") * e2e_cost(...)``) with a minimal expression type instead of pymbolic.
Consumer-side utility of the traversal.
"""
from __future__ import annotations

import numpy as np
import torch

from .array_context import TorchArrayContext


# {{{ a minimal stand-in for the pymbolic expressions the reference's model uses

class _Expr:
    def __add__(self, other):
        return _Bin("+", self, other)

    def __radd__(self, other):
        return _Bin("+", other, self)

    def __mul__(self, other):
        return _Bin("*", self, other)

    def __rmul__(self, other):
        return _Bin("*", other, self)

    def __pow__(self, other):
        return _Bin("**", self, other)


class var(_Expr):  # noqa: N801  (pymbolic's name)
    def __init__(self, name):
        self.name = name


class _Bin(_Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b


def evaluate(expr, context):
    if isinstance(expr, var):
        return context[expr.name]
    if isinstance(expr, _Bin):
        a, b = evaluate(expr.a, context), evaluate(expr.b, context)
        return a + b if expr.op == "+" else a * b if expr.op == "*" else a ** b
    return expr

# }}}


class FMMTranslationCostModel:
    """``cost.py:87-147``: modelled costs of single translations / evaluations."""

    def __init__(self, ncoeffs_fmm_by_level, uses_point_and_shoot):
        self.ncoeffs_fmm_by_level = ncoeffs_fmm_by_level
        self.uses_point_and_shoot = uses_point_and_shoot

    @staticmethod
    def direct():
        return var("c_p2p")

    def p2l(self, level):
        return var("c_p2l") * self.ncoeffs_fmm_by_level[level]

    def l2p(self, level):
        return var("c_l2p") * self.ncoeffs_fmm_by_level[level]

    def p2m(self, level):
        return var("c_p2m") * self.ncoeffs_fmm_by_level[level]

    def m2p(self, level):
        return var("c_m2p") * self.ncoeffs_fmm_by_level[level]

    def m2m(self, src_level, tgt_level):
        return var("c_m2m") * self.e2e_cost(self.ncoeffs_fmm_by_level[src_level],
                                            self.ncoeffs_fmm_by_level[tgt_level])

    def l2l(self, src_level, tgt_level):
        return var("c_l2l") * self.e2e_cost(self.ncoeffs_fmm_by_level[src_level],
                                            self.ncoeffs_fmm_by_level[tgt_level])

    def m2l(self, src_level, tgt_level):
        return var("c_m2l") * self.e2e_cost(self.ncoeffs_fmm_by_level[src_level],
                                            self.ncoeffs_fmm_by_level[tgt_level])

    def e2e_cost(self, nsource_coeffs, ntarget_coeffs):
        if self.uses_point_and_shoot:
            return (nsource_coeffs ** (3 / 2) + nsource_coeffs ** (1 / 2) * ntarget_coeffs
                    + ntarget_coeffs ** (3 / 2))
        return nsource_coeffs * ntarget_coeffs


def make_pde_aware_translation_cost_model(dim, nlevels):
    """``cost.py:152-166``."""
    p_fmm = [var(f"p_fmm_lev{i}") for i in range(nlevels)]
    return FMMTranslationCostModel(ncoeffs_fmm_by_level=[(p + 1) ** (dim - 1) for p in p_fmm],
                                   uses_point_and_shoot=dim == 3)


def make_taylor_translation_cost_model(dim, nlevels):
    """``cost.py:169-181``."""
    p_fmm = [var(f"p_fmm_lev{i}") for i in range(nlevels)]
    return FMMTranslationCostModel(ncoeffs_fmm_by_level=[(p + 1) ** dim for p in p_fmm],
                                   uses_point_and_shoot=False)


def _row_sums(starts, lists, values):
    """``out[i] = sum(values[lists[starts[i]:starts[i+1]]])`` (``bt_csr_row_sums``, float64)."""
    from . import _cabi
    from ._cabi import check, dptr
    nrows = int(starts.shape[0]) - 1
    out = torch.zeros(max(nrows, 1), dtype=torch.float64, device=values.device)
    values = values.to(torch.float64).contiguous()
    check(_cabi.load().bt_csr_row_sums(1, nrows, dptr(starts.contiguous()), dptr(lists.contiguous()),
                                       dptr(values), None, dptr(out), 0, 1.0,
                                       torch.cuda.current_stream(values.device).cuda_stream),
          "bt_csr_row_sums")
    return out[:nrows]


class FMMCostModel:
    """``cost.py:186-713`` (interface) with the per-box processes of ``:715-1262``."""

    _FMM_STAGE_TO_CALIBRATION_PARAMETER = {
        "form_multipoles": "c_p2m", "coarsen_multipoles": "c_m2m", "eval_direct": "c_p2p",
        "multipole_to_local": "c_m2l", "eval_multipoles": "c_m2p", "form_locals": "c_p2l",
        "refine_locals": "c_l2l", "eval_locals": "c_l2p"}

    def __init__(self, translation_cost_model_factory=make_pde_aware_translation_cost_model):
        self.translation_cost_model_factory = translation_cost_model_factory

    # {{{ per-box processes

    @staticmethod
    def _f64(actx, a):
        if isinstance(a, np.ndarray):
            a = actx.from_numpy(np.ascontiguousarray(a, np.float64))
        return a.to(torch.float64)

    def process_form_multipoles(self, actx, traversal, p2m_cost):
        tree = traversal.tree
        sb = traversal.source_boxes.long()
        return (tree.box_source_counts_nonchild.to(torch.float64)[sb]
                * self._f64(actx, p2m_cost)[tree.box_levels.long()[sb]])

    def process_coarsen_multipoles(self, actx, traversal, m2m_cost):
        tree = traversal.tree
        m2m_cost = self._f64(actx, m2m_cost).cpu().numpy()
        lssp = traversal.level_start_source_parent_box_nrs.cpu().tolist()
        nchildren = (tree.box_child_ids[:, :int(tree.nboxes)] != 0).sum(0)
        spb = traversal.source_parent_boxes.long()
        result = 0.0
        for source_level in range(int(tree.nlevels) - 1, 2, -1):
            target_level = source_level - 1
            boxes = spb[lssp[target_level]:lssp[target_level + 1]]
            result += float(m2m_cost[target_level]) * int(nchildren[boxes].sum().item())
        return result

    def get_ndirect_sources_per_target_box(self, actx, traversal):
        nsrc = traversal.tree.box_source_counts_nonchild.to(torch.float64)
        out = _row_sums(traversal.neighbor_source_boxes_starts,
                        traversal.neighbor_source_boxes_lists, nsrc)
        if traversal.from_sep_close_smaller_starts is not None:
            out = out + _row_sums(traversal.from_sep_close_smaller_starts,
                                  traversal.from_sep_close_smaller_lists, nsrc)
        if traversal.from_sep_close_bigger_starts is not None:
            out = out + _row_sums(traversal.from_sep_close_bigger_starts,
                                  traversal.from_sep_close_bigger_lists, nsrc)
        return out

    def process_direct(self, actx, traversal, ndirect_sources_by_itgt_box, p2p_cost,
                       box_target_counts_nonchild=None):
        if box_target_counts_nonchild is None:
            box_target_counts_nonchild = traversal.tree.box_target_counts_nonchild
        ntgt = box_target_counts_nonchild.to(torch.float64)[traversal.target_boxes.long()]
        return ntgt * ndirect_sources_by_itgt_box * float(p2p_cost)

    def process_list2(self, actx, traversal, m2l_cost):
        st = traversal.from_sep_siblings_starts.long()
        lev = traversal.tree.box_levels.long()[traversal.target_or_target_parent_boxes.long()]
        return self._f64(actx, m2l_cost)[lev] * (st[1:] - st[:-1]).to(torch.float64)

    def process_list3(self, actx, traversal, m2p_cost, box_target_counts_nonchild=None):
        tree = traversal.tree
        if box_target_counts_nonchild is None:
            box_target_counts_nonchild = tree.box_target_counts_nonchild
        ntgt = box_target_counts_nonchild.to(torch.float64)
        m2p_cost = self._f64(actx, m2p_cost)
        nm2p = torch.zeros(int(tree.nboxes), dtype=torch.float64, device=ntgt.device)
        for ilevel, ssn in enumerate(traversal.from_sep_smaller_by_level):
            tb = traversal.target_boxes_sep_smaller_by_source_level[ilevel].long()
            st = ssn.starts.long()
            nm2p.index_add_(0, tb, ntgt[tb] * (st[1:] - st[:-1]).to(torch.float64) * m2p_cost[ilevel])
        return nm2p

    def process_list4(self, actx, traversal, p2l_cost):
        tree = traversal.tree
        per_src = (tree.box_source_counts_nonchild.to(torch.float64)
                   * self._f64(actx, p2l_cost)[tree.box_levels.long()])
        return _row_sums(traversal.from_sep_bigger_starts, traversal.from_sep_bigger_lists, per_src)

    def process_eval_locals(self, actx, traversal, l2p_cost, box_target_counts_nonchild=None):
        tree = traversal.tree
        if box_target_counts_nonchild is None:
            box_target_counts_nonchild = tree.box_target_counts_nonchild
        tb = traversal.target_boxes.long()
        return (box_target_counts_nonchild.to(torch.float64)[tb]
                * self._f64(actx, l2p_cost)[tree.box_levels.long()[tb]])

    def process_refine_locals(self, actx, traversal, l2l_cost):
        l2l_cost = self._f64(actx, l2l_cost).cpu().numpy()
        lstp = traversal.level_start_target_or_target_parent_box_nrs.cpu().tolist()
        result = 0.0
        for target_lev in range(1, int(traversal.tree.nlevels)):
            result += (lstp[target_lev + 1] - lstp[target_lev]) * float(l2l_cost[target_lev - 1])
        return result

    # }}}

    @staticmethod
    def zero_cost_per_box(actx, nboxes):
        return actx.zeros(nboxes, np.float64)

    @staticmethod
    def aggregate_over_boxes(actx, per_box_result):
        if isinstance(per_box_result, float):
            return per_box_result
        return float(per_box_result.sum().item())

    def fmm_cost_factors_for_kernels_from_model(self, actx, nlevels, xlat_cost, context):
        """``cost.py:387-436``: evaluate the symbolic model for every level."""
        def ev(e):
            return evaluate(e, context)
        return {
            "p2m_cost": np.array([ev(xlat_cost.p2m(i)) for i in range(nlevels)], np.float64),
            "m2m_cost": np.array([ev(xlat_cost.m2m(i + 1, i)) for i in range(nlevels - 1)], np.float64),
            "c_p2p": ev(xlat_cost.direct()),
            "m2l_cost": np.array([ev(xlat_cost.m2l(i, i)) for i in range(nlevels)], np.float64),
            "m2p_cost": np.array([ev(xlat_cost.m2p(i)) for i in range(nlevels)], np.float64),
            "p2l_cost": np.array([ev(xlat_cost.p2l(i)) for i in range(nlevels)], np.float64),
            "l2l_cost": np.array([ev(xlat_cost.l2l(i, i + 1)) for i in range(nlevels - 1)], np.float64),
            "l2p_cost": np.array([ev(xlat_cost.l2p(i)) for i in range(nlevels)], np.float64),
        }

    def _factors(self, actx, traversal, level_to_order, calibration_params):
        tree = traversal.tree
        for ilevel in range(int(tree.nlevels)):
            calibration_params[f"p_fmm_lev{ilevel}"] = level_to_order[ilevel]
        xlat_cost = self.translation_cost_model_factory(int(tree.dimensions), int(tree.nlevels))
        return self.fmm_cost_factors_for_kernels_from_model(actx, int(tree.nlevels), xlat_cost,
                                                            calibration_params)

    def cost_per_box(self, actx, traversal, level_to_order, calibration_params,
                     ndirect_sources_per_target_box=None, box_target_counts_nonchild=None):
        """``cost.py:445-525``: float64 ``[nboxes]``, the modelled cost of all stages per box."""
        assert isinstance(actx, TorchArrayContext)
        with torch.cuda.stream(actx.stream):
            if ndirect_sources_per_target_box is None:
                ndirect_sources_per_target_box = self.get_ndirect_sources_per_target_box(actx, traversal)
            tree = traversal.tree
            cost = self._factors(actx, traversal, level_to_order, calibration_params)
            if box_target_counts_nonchild is None:
                box_target_counts_nonchild = tree.box_target_counts_nonchild
            result = self.zero_cost_per_box(actx, int(tree.nboxes))
            tb = traversal.target_boxes.long()
            tp = traversal.target_or_target_parent_boxes.long()
            result[traversal.source_boxes.long()] += self.process_form_multipoles(
                actx, traversal, cost["p2m_cost"])
            result[tb] += self.process_direct(
                actx, traversal, ndirect_sources_per_target_box, cost["c_p2p"],
                box_target_counts_nonchild=box_target_counts_nonchild)
            result[tp] += self.process_list2(actx, traversal, cost["m2l_cost"])
            result += self.process_list3(actx, traversal, cost["m2p_cost"],
                                         box_target_counts_nonchild=box_target_counts_nonchild)
            result[tp] += self.process_list4(actx, traversal, cost["p2l_cost"])
            result[tb] += self.process_eval_locals(
                actx, traversal, cost["l2p_cost"], box_target_counts_nonchild=box_target_counts_nonchild)
        return result

    def cost_per_stage(self, actx, traversal, level_to_order, calibration_params,
                       ndirect_sources_per_target_box=None, box_target_counts_nonchild=None):
        """``cost.py:527-625``: dict stage name -> modelled cost."""
        assert isinstance(actx, TorchArrayContext)
        with torch.cuda.stream(actx.stream):
            if ndirect_sources_per_target_box is None:
                ndirect_sources_per_target_box = self.get_ndirect_sources_per_target_box(actx, traversal)
            cost = self._factors(actx, traversal, level_to_order, calibration_params)
            if box_target_counts_nonchild is None:
                box_target_counts_nonchild = traversal.tree.box_target_counts_nonchild
            agg = self.aggregate_over_boxes
            return {
                "form_multipoles": agg(actx, self.process_form_multipoles(actx, traversal, cost["p2m_cost"])),
                "coarsen_multipoles": self.process_coarsen_multipoles(actx, traversal, cost["m2m_cost"]),
                "eval_direct": agg(actx, self.process_direct(
                    actx, traversal, ndirect_sources_per_target_box, cost["c_p2p"],
                    box_target_counts_nonchild=box_target_counts_nonchild)),
                "multipole_to_local": agg(actx, self.process_list2(actx, traversal, cost["m2l_cost"])),
                "eval_multipoles": agg(actx, self.process_list3(
                    actx, traversal, cost["m2p_cost"], box_target_counts_nonchild=box_target_counts_nonchild)),
                "form_locals": agg(actx, self.process_list4(actx, traversal, cost["p2l_cost"])),
                "refine_locals": self.process_refine_locals(actx, traversal, cost["l2l_cost"]),
                "eval_locals": agg(actx, self.process_eval_locals(
                    actx, traversal, cost["l2p_cost"], box_target_counts_nonchild=box_target_counts_nonchild)),
            }

    @staticmethod
    def get_unit_calibration_params():
        return {"c_l2l": 1.0, "c_l2p": 1.0, "c_m2l": 1.0, "c_m2m": 1.0, "c_m2p": 1.0,
                "c_p2l": 1.0, "c_p2m": 1.0, "c_p2p": 1.0}

    def estimate_calibration_params(self, model_results, timing_results,
                                    time_field_name="wall_elapsed",
                                    additional_stage_to_param_names=()):
        """``cost.py:650-713``: least-squares fit of one factor per stage."""
        nresults = len(model_results)
        assert len(timing_results) == nresults
        stage_to_param_names = dict(self._FMM_STAGE_TO_CALIBRATION_PARAMETER)
        stage_to_param_names.update(additional_stage_to_param_names)
        params = set(stage_to_param_names.values())
        uncalibrated_times = {p: np.zeros(nresults) for p in params}
        actual_times = {p: np.zeros(nresults) for p in params}
        for icase, model_result in enumerate(model_results):
            for stage_name, param_name in stage_to_param_names.items():
                if stage_name in model_result:
                    uncalibrated_times[param_name][icase] = model_result[stage_name]
        for icase, timing_result in enumerate(timing_results):
            for stage_name, time in timing_result.items():
                actual_times[stage_to_param_names[stage_name]][icase] = time[time_field_name]
        result = {}
        for param in params:
            uncalibrated, actual = uncalibrated_times[param], actual_times[param]
            if np.allclose(uncalibrated, 0):
                result[param] = 0.0
                continue
            result[param] = actual.dot(uncalibrated) / uncalibrated.dot(uncalibrated)
        return result
