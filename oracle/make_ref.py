"""Materialise the live reference under oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).

TEST INFRASTRUCTURE, not product: only tests/, __graft_entry__ and bench.py's CPU legs may import what this
produces.  The reference is pure Python, so "building" it means placing its own files where an interpreter
without /root/reference can import them:

    oracle/_ref/FIAT/            <- /root/reference/FIAT/*.py            (unmodified)
    oracle/_ref/gem/utils.py     <- /root/reference/gem/utils.py         (FIAT/reference_element.py:29 needs safe_repr;
                                    the package __init__ is left empty so that the rest of gem is not pulled in)
    oracle/_ref/recursivenodes/  <- tests/golden/gen/recursivenodes/     (our stand-in for the un-vendored PyPI
                                    package; construction-time only)
    oracle/_ref/ref_tests/       <- /root/reference/test/FIAT/unit/test_*.py  (the reference's own unit tests, run by
                                    tests/test_reference_suite.py with the tabulation path swapped out,
                                    oracle/dropin_plugin.py)

Run by `__graft_entry__.build()` in the build container whenever /root/reference is present; on the GPU box the
already materialised copy is used.  Nothing under oracle/_ref/ is committed.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FIAT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def make(force=False):
    """-> path of oracle/_ref, or None when neither the reference nor an earlier copy is available."""
    src = os.path.join(REF, "FIAT")
    stamp = os.path.join(OUT, "FIAT", "__init__.py")
    if not os.path.isdir(src):
        return OUT if os.path.exists(stamp) else None
    if os.path.exists(stamp) and not force:
        newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src) if f.endswith(".py"))
        if os.path.getmtime(stamp) >= newest:
            return OUT
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "gem"))
    shutil.copytree(src, os.path.join(OUT, "FIAT"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy(os.path.join(REF, "gem", "utils.py"), os.path.join(OUT, "gem", "utils.py"))
    open(os.path.join(OUT, "gem", "__init__.py"), "w").close()
    shutil.copytree(os.path.join(HERE, "..", "tests", "golden", "gen", "recursivenodes"),
                    os.path.join(OUT, "recursivenodes"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    tests_src = os.path.join(REF, "test", "FIAT", "unit")
    if os.path.isdir(tests_src):
        os.makedirs(os.path.join(OUT, "ref_tests"))
        for name in sorted(os.listdir(tests_src)):
            if name.startswith("test_") and name.endswith(".py"):
                shutil.copy(os.path.join(tests_src, name), os.path.join(OUT, "ref_tests", name))
    os.utime(stamp)
    return OUT


def import_reference():
    """Import the live reference from oracle/_ref (never from /root/reference).  -> the FIAT module, or None."""
    stamp = os.path.join(OUT, "FIAT", "__init__.py")
    if not os.path.exists(stamp):
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import FIAT
    return FIAT


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
