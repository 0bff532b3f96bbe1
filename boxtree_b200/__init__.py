"""boxtree_b200 -- B200-native (sm_100a) tree build + FMM traversal generation.

Drop-in for the ``TreeBuilder.__call__`` -> ``Tree`` ->
``FMMTraversalBuilder.__call__`` -> ``FMMTraversalInfo`` path of
`inducer/boxtree <https://github.com/inducer/boxtree>`__ (same names, argument
meaning and error behaviour as ``boxtree/__init__.py:26-52``), implemented as
hand-written CUDA kernels behind a C ABI (``include/boxtree_b200.h``).

Particle orderings and CSR storage follow the reference's conventions
(``boxtree/__init__.py:114-166``): ``user_source_ids`` maps tree order to user
order for sources, ``sorted_target_ids`` maps user order to tree order for
targets; ``*_starts``/``*_lists`` pairs are CSR with ``starts`` of length
``nrows + 1``.
"""
from .array_context import TorchArrayContext, make_obj_array
from .tree import Tree, TreeOfBoxes, box_flags_enum
from .tree_build import MaxLevelsExceeded, TreeBuilder
from .traversal import BuiltList, FMMTraversalBuilder, FMMTraversalInfo
from .particle_filter import (FilteredTargetListsInTreeOrder, FilteredTargetListsInUserOrder,
                              ParticleListFilter)
from .point_sources import TreeWithLinkedPointSources, link_point_sources
from .area_query import (AreaQueryBuilder, AreaQueryResult, LeavesToBallsLookup,
                         LeavesToBallsLookupBuilder, PeerListFinder, PeerListLookup,
                         SpaceInvaderQueryBuilder)
from .translation_classes import (RotationClassesBuilder, RotationClassesInfo,
                                  TranslationClassesBuilder, TranslationClassesInfo)
from .cost import (FMMCostModel, FMMTranslationCostModel, make_pde_aware_translation_cost_model,
                   make_taylor_translation_cost_model)

__all__ = [
    "TorchArrayContext", "make_obj_array",
    "Tree", "TreeOfBoxes", "box_flags_enum",
    "TreeBuilder", "MaxLevelsExceeded",
    "FMMTraversalBuilder", "FMMTraversalInfo", "BuiltList",
    "ParticleListFilter", "FilteredTargetListsInUserOrder", "FilteredTargetListsInTreeOrder",
    "TreeWithLinkedPointSources", "link_point_sources",
    "PeerListFinder", "PeerListLookup", "AreaQueryBuilder", "AreaQueryResult",
    "LeavesToBallsLookupBuilder", "LeavesToBallsLookup", "SpaceInvaderQueryBuilder",
    "TranslationClassesBuilder", "TranslationClassesInfo", "RotationClassesBuilder",
    "RotationClassesInfo",
    "FMMCostModel", "FMMTranslationCostModel", "make_pde_aware_translation_cost_model",
    "make_taylor_translation_cost_model",
]
