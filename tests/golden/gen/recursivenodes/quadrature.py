"""Gauss-Jacobi type rules on [-1, 1] and the Stroud conical rule on the (-1, 1) simplex.

Used by FIAT/quadrature.py:13,102,123,179.
"""
import numpy
from scipy.special import roots_jacobi, eval_legendre


def gaussjacobi(m, a=0.0, b=0.0):
    x, w = roots_jacobi(m, a, b)
    return x, w


def lobattogaussjacobi(m, a=0.0, b=0.0):
    assert a == 0 and b == 0
    if m == 2:
        return numpy.array([-1.0, 1.0]), numpy.array([1.0, 1.0])
    xi, _ = roots_jacobi(m - 2, 1.0, 1.0)
    x = numpy.concatenate(([-1.0], xi, [1.0]))
    w = 2.0 / (m * (m - 1) * eval_legendre(m - 1, x) ** 2)
    return x, w


def simplexgausslegendre(d, m):
    """m**d point conical product rule on the simplex with vertices at -1/+1.

    Collapsed axis i carries a Gauss-Jacobi(m, i, 0) rule; simplex coordinate i is
    x_i = 2 u_i prod_{j>i} (1 - u_j) - 1 with u = (1 + eta)/2, so the Jacobian
    prod_j (1 - u_j)^j is absorbed by the Jacobi weights (divided by 2^i each).
    """
    rules = [roots_jacobi(m, float(i), 0.0) for i in range(d)]
    grids = numpy.meshgrid(*[r[0] for r in rules], indexing="ij")
    wgrids = numpy.meshgrid(*[r[1] for r in rules], indexing="ij")
    u = [0.5 * (1.0 + g.ravel()) for g in grids]
    w = numpy.ones_like(u[0])
    for i, wg in enumerate(wgrids):
        w = w * wg.ravel() / 2.0 ** i
    x = numpy.zeros((len(w), d))
    rem = numpy.ones_like(w)
    for i in reversed(range(d)):
        x[:, i] = 2.0 * u[i] * rem - 1.0
        rem = rem * (1.0 - u[i])
    return x, w
