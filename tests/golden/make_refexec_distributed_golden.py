"""``refexec_distributed_digests.json``: the REFERENCE's own distributed setup
(``partition_work``, ``get_box_masks``, ``generate_local_tree``, ``generate_local_travs`` of
``/root/reference/boxtree/distributed``, unmodified, one thread per rank over an in-process
communicator; see ``tests/refexec``) on the cases of ``tests/dist_cases.py``: per rank, a short
sha256 of every output.  ``tests/test_gpu_distributed.py`` holds the CUDA path to them,
``tests/test_refexec.py`` the oracle.

    python tests/golden/make_refexec_distributed_golden.py
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from refexec.run import reference_distributed_setup            # noqa: E402
from tests.dist_cases import CASES, NRANKS, box_cost            # noqa: E402
from tests.parity_util import distributed_rank_digests           # noqa: E402


def main():
    out = {}
    for name in sorted(CASES):
        src, tkw, vkw = CASES[name]()
        for nranks in NRANKS:
            tree, _, ranks = reference_distributed_setup(src, tkw, vkw, nranks, box_cost)
            out[f"{name}:{nranks}"] = [
                distributed_rank_digests(r["responsible_boxes_list"], r["masks"], r["local_tree"],
                                         r["src_idx"], r["tgt_idx"], r["local_trav"], tree.nboxes)
                for r in ranks]
            print(name, nranks, "nboxes", tree.nboxes, flush=True)
    with open(os.path.join(HERE, "refexec_distributed_digests.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
