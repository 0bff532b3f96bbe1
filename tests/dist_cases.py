"""Inputs shared by the distributed-setup parity tests and the golden generator."""
import numpy as np

from tests.parity_util import config3_inputs, normal_particles

CASES = {
    "points": lambda: (normal_particles(20000, 3, np.float64), dict(max_particles_in_box=30), {}),
    "points2d-f32-2away": lambda: (normal_particles(20000, 2, np.float32),
                                   dict(max_particles_in_box=30), dict(well_sep_is_n_away=2)),
    "config3": lambda: (lambda s, t, r: (s, dict(
        max_particles_in_box=30, targets=t, target_radii=r, stick_out_factor=0.25,
        extent_norm="linf", kind="adaptive-level-restricted"), {}))(*config3_inputs(20000, 20000)),
}
NRANKS = (1, 3, 4)


def box_cost(tree):
    """The cost vector the tests partition by: 1 + own sources + own targets per box."""
    nb = tree.nboxes
    return (1.0 + np.asarray(tree.box_source_counts_nonchild)[:nb]
            + np.asarray(tree.box_target_counts_nonchild)[:nb]).astype(np.float64)
