"""Stand-in for the third-party `recursivenodes` package (T. Isaac, PyPI, unpinned in the
reference's pyproject.toml).  The package is not installed in the build container and cannot
be fetched, so this module restates the handful of published formulas that FIAT needs at
*element construction time* (node families, the recursive simplex node rule, Gauss-Jacobi
rules).  None of this is on the tabulation hot path; it only makes the reference importable
so that golden fixtures can be generated (tests/golden/gen/make_golden.py).
"""
