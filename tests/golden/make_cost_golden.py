"""Generate ``tests/golden/cost_model.json``: per-box and per-stage costs computed by the
REFERENCE's own ``_PythonFMMCostModel`` (``/root/reference/boxtree/cost.py:1264-1443``, pure
Python), executed in place on the oracle's traversal of seeded inputs.  pymbolic, mako,
arraycontext and pyopencl (imported at the top of cost.py) are not installed: minimal stand-ins
with the handful of names cost.py touches are registered first; nothing is copied.

    python tests/golden/make_cost_golden.py
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE_ROOT = "/root/reference"


class _Expr:
    def __add__(self, o): return _Bin("+", self, o)
    def __radd__(self, o): return _Bin("+", o, self)
    def __mul__(self, o): return _Bin("*", self, o)
    def __rmul__(self, o): return _Bin("*", o, self)
    def __pow__(self, o): return _Bin("**", self, o)


class _Var(_Expr):
    def __init__(self, name): self.name = name


class _Bin(_Expr):
    def __init__(self, op, a, b): self.op, self.a, self.b = op, a, b


def _evaluate(e, context):
    if isinstance(e, _Var):
        return context[e.name]
    if isinstance(e, _Bin):
        a, b = _evaluate(e.a, context), _evaluate(e.b, context)
        return a + b if e.op == "+" else a * b if e.op == "*" else a ** b
    return e


def load_reference_cost():
    pkg = types.ModuleType("boxtree")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree")]
    sys.modules["boxtree"] = pkg

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    stub("mako")
    stub("mako.template", Template=lambda *a, **k: None)
    stub("arraycontext", ArrayContext=object, PyOpenCLArrayContext=object)
    stub("pymbolic", var=_Var, evaluate=_evaluate)
    stub("pyopencl")
    stub("pyopencl.elementwise", ElementwiseKernel=object)
    stub("pyopencl.tools", dtype_to_ctype=str)
    stub("pytools", keyed_memoize_method=lambda *a, **k: (lambda f: f))
    mod = importlib.import_module("boxtree.cost")
    assert mod.__file__.startswith(REFERENCE_ROOT)
    return mod


def cases():
    from tests.parity_util import config3_inputs, normal_particles
    out = {}
    out["3d-points"] = (normal_particles(6000, 3, np.float64), dict(max_particles_in_box=30), {})
    out["2d-points-2away"] = (normal_particles(6000, 2, np.float64), dict(max_particles_in_box=30),
                              dict(well_sep_is_n_away=2))
    s, t, r = config3_inputs(4000, 4000)
    out["3d-config3"] = (s, dict(max_particles_in_box=30, targets=t, target_radii=r,
                                 stick_out_factor=0.25, extent_norm="linf",
                                 kind="adaptive-level-restricted"), {})
    return out


def level_to_order(nlevels):
    return np.array([3 + (i % 4) for i in range(nlevels)], dtype=np.int64)


CALIBRATION = {"c_l2l": 1.5, "c_l2p": 0.75, "c_m2l": 2.0, "c_m2m": 1.25, "c_m2p": 0.5,
               "c_p2l": 3.0, "c_p2m": 1.0, "c_p2p": 0.125}


def main():
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    cost = load_reference_cost()
    golden = {}
    for name, (src, tkw, vkw) in cases().items():
        tree = build_tree(src, **tkw)
        trav = build_traversal(tree, **vkw)
        for factory in ("make_pde_aware_translation_cost_model", "make_taylor_translation_cost_model"):
            model = cost._PythonFMMCostModel(getattr(cost, factory))
            per_box = model.cost_per_box(None, trav, level_to_order(tree.nlevels), dict(CALIBRATION))
            per_stage = model.cost_per_stage(None, trav, level_to_order(tree.nlevels), dict(CALIBRATION))
            golden[f"{name}/{factory}"] = {
                "nboxes": int(tree.nboxes),
                "per_box_sum": float(np.sum(per_box)),
                "per_box_sha256": hashlib.sha256(np.ascontiguousarray(per_box).tobytes()).hexdigest(),
                "per_box_head": [float(x) for x in per_box[:64]],
                "per_box_every_97th": [float(x) for x in per_box[::97]],
                "per_stage": {k: float(v) for k, v in per_stage.items()},
            }
            print(name, factory, tree.nboxes, float(np.sum(per_box)))
    with open(os.path.join(HERE, "cost_model.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
