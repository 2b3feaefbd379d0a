"""1-D node families on [0, 1] and Isaac's recursive simplex interpolation-node rule.

Used by FIAT/reference_element.py:30,89-98 (`make_lattice`).
"""
import numpy
from scipy.special import roots_jacobi


class _Family:
    """family[n] -> ascending array of n+1 nodes on [0, 1]."""

    def __init__(self, fn):
        self._fn = fn
        self._cache = {}

    def __getitem__(self, n):
        try:
            return self._cache[n]
        except KeyError:
            x = numpy.asarray(self._fn(n), dtype=float)
            x = 0.5 * (x + (1.0 - x[::-1]))  # enforce symmetry about 1/2
            return self._cache.setdefault(n, x)


def _equi(n):
    return numpy.array([0.5]) if n == 0 else numpy.linspace(0.0, 1.0, n + 1)


def _equi_interior(n):
    return (numpy.arange(n + 1) + 0.5) / (n + 1)


def _lgl(n):
    if n == 0:
        return numpy.array([0.5])
    if n == 1:
        return numpy.array([0.0, 1.0])
    xi, _ = roots_jacobi(n - 1, 1.0, 1.0)
    return 0.5 * (numpy.concatenate(([-1.0], xi, [1.0])) + 1.0)


def _gl(n):
    xi, _ = roots_jacobi(n + 1, 0.0, 0.0)
    return 0.5 * (xi + 1.0)


def _lgc(n):
    if n == 0:
        return numpy.array([0.5])
    return 0.5 * (1.0 - numpy.cos(numpy.pi * numpy.arange(n + 1) / n))


def _gc(n):
    k = numpy.arange(n + 1)
    return 0.5 * (1.0 - numpy.cos(numpy.pi * (2 * k + 1) / (2 * n + 2)))


_FAMILIES = {
    "equi": _Family(_equi),
    "equi_interior": _Family(_equi_interior),
    "lgl": _Family(_lgl),
    "gl": _Family(_gl),
    "lgc": _Family(_lgc),
    "gc": _Family(_gc),
}


def _decode_family(family):
    if isinstance(family, _Family):
        return family
    if family is None:
        family = "lgl"
    return _FAMILIES[family]


def _recursive(d, n, alpha, family):
    """Barycentric coordinates of the node with multi-index `alpha` (len d+1, sum n)."""
    alpha = tuple(int(a) for a in alpha)
    if family is _FAMILIES["equi"] and n > 0:
        # the recursion reproduces alpha/n for equispaced nodes; return it without round-off
        return numpy.array(alpha, dtype=float) / n
    xn = family[n]
    b = numpy.zeros(d + 1)
    if d == 1:
        b[0] = xn[alpha[0]]
        b[1] = xn[alpha[1]]
        return b
    weight = 0.0
    for i in range(d + 1):
        alpha_noti = alpha[:i] + alpha[i + 1:]
        n_noti = n - alpha[i]
        w = xn[n_noti]
        br = _recursive(d - 1, n_noti, alpha_noti, family)
        b[:i] += w * br[:i]
        b[i + 1:] += w * br[i:]
        weight += w
    return b / weight
