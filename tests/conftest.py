import glob
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The GPU tests are about the streaming kernels: small calls must not be diverted to the quick plan (api.QUICK_NPTS).
# tests/test_gpu_parity.py::test_quick_plan_for_small_calls and the reference's unit tests over the drop-in
# (tests/test_reference_suite.py, run in a subprocess with the default) cover the quick plan itself.
os.environ.setdefault("FIATB200_QUICK_NPTS", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_case_names():
    return sorted(os.path.basename(p)[len("case_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


def _as_tuples(key):
    return tuple(_as_tuples(k) for k in key) if isinstance(key, list) else key


def load_case(name):
    from fiat_b200 import description
    case = description.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    ent = case["entity"]
    if ent == "none":
        case["entity"] = None
    else:
        dim, eid = ent
        case["entity"] = (_as_tuples(dim), eid)
    case["ref"] = {tuple(k): v for k, v in zip(case["keys"], case["values"])}
    case["error_keys"] = [tuple(k) for k in case.get("error_keys", [])]    # slots holding exception objects (trace elements)
    return case


def is_polynomial_case(case):
    """False for the elements that are not polynomial tabulations (HDivTrace, QuadratureElement): pinned by the golden
    files directly, no oracle / plan."""
    return case["desc"]["kind"] not in ("trace", "quadrature")


def load_desc(name):
    from fiat_b200 import description
    return description.load(os.path.join(GOLDEN, f"desc_{name}.npz"))


def tolerance(desc, alpha):
    """north_star: 1e-12 * max|ref| per derivative component; 1e-10 for order >= 2 at degree >= 8."""
    from oracle.tolerance import tolerance as tol
    return tol(desc, alpha)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
