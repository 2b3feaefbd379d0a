"""The north-star tolerance (BASELINE.json): max|delta| <= 1e-12 * max|ref| per derivative component, 1e-10 for
derivative order >= 2 at degree >= 8.  Test infrastructure, like everything under oracle/."""


def degree_of(desc):
    kind = desc["kind"]
    if kind in ("trace", "quadrature"):
        return 0
    if kind == "simplex":
        return int(desc["degree"])
    if kind == "flattened":
        return degree_of(desc["element"])
    if kind == "composite":
        return max(degree_of(p["element"]) for p in desc["parts"])
    return max(degree_of(desc["A"]), degree_of(desc["B"]))


def tolerance(desc, alpha):
    return 1e-10 if (sum(alpha) >= 2 and degree_of(desc) >= 8) else 1e-12
