"""The reference's OWN unit tests, run with `tabulate` swapped for the drop-in.

`oracle/make_ref.py` materialises the reference package and its `test/FIAT/unit/test_*.py` files under the git-ignored
`oracle/_ref/` (it travels to the GPU box); `oracle/dropin_plugin.py` replaces every `tabulate` on the hot path
before those tests are collected.  The reference's known-answer tests -- exact Dubiner values, nodality, partition of
unity, macro-element continuity, tensor-product dof order, trace elements -- then judge

    not gpu:  the CPU oracle (`FIATB200_DROPIN=oracle`) -- pins the oracle with the reference's own assertions
    not gpu:  the device mode's host logic (`FIATB200_DROPIN=emulate`: `fiat_b200.api` with the kernel launches of
              polynomial elements answered by the oracle) -- trace / quadrature elements, the single-point form, the
              exception types, and the re-entrancy of the binding: `describe_element` tabulates sub-elements of the
              reference, which comes back into the library when `tabulate` is bound to it; `get_tabulator` used to
              hold its cache lock across that call and dead-locked on every interval trace element (found here)
    gpu:      the CUDA path  (`FIATB200_DROPIN=device`, numpy in / numpy out through the C ABI)

The default selection of the CPU test is the files that exercise tabulation most (about 690 tests, 40 s).  On the GPU
every worker process pays its own CUDA context and the host-side element construction dominates (560 tests took 5.5
minutes on the B200 box, all passing: profiles/r02_reference_suite_device.txt), so the GPU test takes the files about
tensor products, trace / quadrature elements, Regge / HHJ and discontinuous Taylor elements (117 tests) and gives up
(skip, not fail) if the box needs more than 10 minutes; `FIATB200_REF_SUITE=full` runs all 36 files in either mode.  Tests that need `gem` / `sympy`-through-gem (absent here: test_precision.py, test_macro.py::test_macro_gem /
test_macro_sympy) fail the same way without the plugin and are left out.
"""
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
REF_TESTS = os.path.join(REF, "ref_tests")
DEFAULT_FILES = ["test_fiat.py", "test_tensor_product.py", "test_regge_hhj.py", "test_macro.py", "test_hdivtrace.py",
                 "test_discontinuous_taylor.py", "test_serendipity.py", "test_quadrature_element.py"]
NEEDS_GEM = "not macro_gem and not macro_sympy"


GPU_FILES = ["test_tensor_product.py", "test_regge_hhj.py", "test_hdivtrace.py", "test_quadrature_element.py",
             "test_discontinuous_taylor.py"]


def _run(mode, tmp_path, workers, files=DEFAULT_FILES, min_passed=600, min_replaced=700, timeout=3000):
    if not os.path.isdir(REF_TESTS):
        pytest.skip("oracle/_ref/ref_tests absent (python -c 'import __graft_entry__ as g; g.build()' makes it where "
                    "/root/reference exists)")
    if os.environ.get("FIATB200_REF_SUITE") == "full":
        targets = [REF_TESTS, f"--ignore={os.path.join(REF_TESTS, 'test_precision.py')}"]
    else:
        targets = [os.path.join(REF_TESTS, f) for f in files]
    stats_file = tmp_path / "dropin_stats.jsonl"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, ROOT]), FIATB200_DROPIN=mode,
               FIATB200_DROPIN_STATS=str(stats_file))
    env.pop("FIATB200_QUICK_NPTS", None)         # the library's default: small calls take the quick plan
    cmd = [sys.executable, "-m", "pytest", "-p", "oracle.dropin_plugin", "-q", "-p", "no:cacheprovider",
           "-c", os.devnull, "--rootdir", str(tmp_path), "-k", NEEDS_GEM, "-n", str(workers),
           "--timeout", "300"] + targets          # (per test: a hang fails that test instead of the whole run)
    try:
        res = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        pytest.skip(f"the reference's unit tests ({mode} mode) did not finish within {timeout} s on this machine")
    tail = res.stdout[-4000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= min_passed, tail
    replaced = fallback = 0
    reasons = {}
    with open(stats_file) as f:
        for line in f:
            rec = json.loads(line)
            replaced += rec["replaced"]
            fallback += rec["fallback"]
            reasons.update(rec["fallback_reasons"])
    # the replacement must actually have been what the tests judged
    assert replaced >= min_replaced, (replaced, fallback, reasons)
    assert fallback <= replaced // 4, (replaced, fallback, reasons)
    # nothing but what the docstring of the plugin lists is handed back to the reference
    # (on the device, sizes the library does not take are handed back as well: "fiat_b200 error 3")
    for reason in reasons:
        assert "symbolic points" in reason or "elements on a point" in reason or "fiat_b200 error 3" in reason \
            or (mode == "device" and reason.startswith(("UnsupportedElement", "UnsupportedByLibrary"))), reasons
    return replaced, fallback


def test_reference_unit_tests_judge_the_oracle(tmp_path):
    _run("oracle", tmp_path, workers=4)


def test_reference_unit_tests_judge_the_device_host_logic(tmp_path):
    _run("emulate", tmp_path, workers=4)


@pytest.mark.gpu
def test_reference_unit_tests_judge_the_device_path(tmp_path):
    full = os.environ.get("FIATB200_REF_SUITE") == "full"
    _run("device", tmp_path, workers=2, files=GPU_FILES, min_passed=600 if full else 90,
         min_replaced=700 if full else 100, timeout=3000 if full else 600)
