"""Run the reference's OWN test files (``/root/reference/test/test_*.py``) on the CPU through
``tests/refexec`` and print a summary.

    python tests/refexec/run_reference_tests.py [pytest args / test ids ...]

Why: a test of the reference passing here means the stand-ins in ``fakecl.py`` execute the
reference's kernels the way its authors expect -- evidence that outputs produced through refexec
(``tests/golden/refexec_*``) are the reference's outputs.  Nothing is written under
``/root/reference`` (no bytecode, no pytest cache; rootdir is a scratch directory).

Not run, and why: ``test_tree_of_boxes.py`` (imports meshmode), ``test_distributed.py`` (mpi4py;
the reference skips it without MPI), the pyfmmlib-based FMM tests (the reference skips them
without pyfmmlib), ``test_tools.py::test_device_record*`` (exercise arraycontext's own container
plumbing, which the stand-in does not model).
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
REFERENCE_TESTS = "/root/reference/test"

DEFAULT_FILES = ["test_tree.py", "test_traversal.py", "test_cost_model.py", "test_tools.py",
                 "test_fmm.py"]
# node ids as pytest forms them for files outside its rootdir
DESELECT = ["::test_device_record", "::test_device_record_array_context"]


def run(args, quiet=True):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([TESTS, ROOT, REFERENCE_TESTS,
                                           os.environ.get("PYTHONPATH", "")]))
    scratch = tempfile.mkdtemp(prefix="refexec-root-")
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p",
           "refexec.pytest_plugin", "--rootdir", scratch, "-q" if quiet else "-v",
           "-W", "ignore::DeprecationWarning"]
    for d in DESELECT:
        cmd += ["--deselect", d]
    cmd += args
    return subprocess.run(cmd, cwd=scratch, env=env, capture_output=True, text=True)


def main():
    args = sys.argv[1:]
    if args:
        proc = run([a if a.startswith("-") or os.path.isabs(a) else os.path.join(REFERENCE_TESTS, a)
                    for a in args], quiet=False)
        print(proc.stdout[-6000:])
        sys.exit(proc.returncode)
    worst = 0
    for name in DEFAULT_FILES:
        proc = run([os.path.join(REFERENCE_TESTS, name)])
        summary = [ln for ln in proc.stdout.splitlines() if " passed" in ln or " failed" in ln
                   or " error" in ln or " skipped" in ln]
        print(f"{name}: {summary[-1] if summary else proc.stdout[-400:]}", flush=True)
        for ln in proc.stdout.splitlines():
            if ln.startswith("FAILED") or ln.startswith("ERROR"):
                print("   ", ln)
        worst = max(worst, proc.returncode)
    sys.exit(worst)


if __name__ == "__main__":
    main()
