"""CPU tests of the boundary: the C-ABI library loads and exports every symbol declared in
include/boxtree_b200.h; the product path fails loudly without a GPU; bench's reference arm."""
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "boxtree_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "bt_bounding_box" in syms and "bt_trav_list3" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol():
    from boxtree_b200 import _cabi
    lib = _cabi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    # and the binding table covers the header
    bound = set(_cabi.SIGNATURES) | {"bt_launch_count", "bt_prof_enable", "bt_prof_reset",
                                     "bt_prof_report", "bt_set_walk_mode", "bt_get_walk_mode"}
    assert set(declared_symbols()) == bound


def test_no_compute_entry_point_is_a_stub():
    from boxtree_b200 import _cabi
    assert _cabi.load().bt_max_key_level(3) == 19
    assert _cabi.load().bt_max_key_level(2) == 28
    assert _cabi.load().bt_max_tree_level(3) == 31
    assert _cabi.launch_count() == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "boxtree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the CPU oracle", "").replace(
                    "CPU oracle's", ""), f"{f} mentions the oracle"


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from boxtree_b200 import TorchArrayContext
    with pytest.raises(RuntimeError):
        TorchArrayContext()


def test_bench_reference_arm_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "config3", "--n", "40000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port"
