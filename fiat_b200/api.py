"""Public API: a drop-in for `FiniteElement.tabulate(order, points, entity=None)`
(FIAT/finite_element.py:98-109,181-197; TensorProductElement / FlattenedDimensions:
FIAT/tensor_product.py:231-336,396-407) that runs on the B200.

    from fiat_b200 import tabulate
    tab = tabulate(element, order, points)          # element built by FIAT itself
    tab[(1, 0, 0)]                                   # torch.float64 cuda tensor (ndofs, *value_shape, npts)

The result is the same dict as the reference's: keys are derivative multi-indices in the order
mis(sd,0), ..., mis(sd,order); values have shape (ndofs, *value_shape, npoints), float64.  The
values are views of ONE device allocation `(nalpha, ndofs, *value_shape, npoints)`.

`element` may also be an element description (`fiat_b200.extract.describe_element`, possibly
loaded from a file), which is how the GPU box runs without FIAT installed.  Plans (device tables)
are cached per element and order.  There is no CPU fallback: unsupported elements raise
`NotImplementedError`, a missing CUDA library or device raises.
"""
import collections
import os
import ctypes
import functools
import threading
import weakref

import numpy
import torch

from . import _lib, plan as planmod
from .extract import describe_element, UnsupportedElement

__all__ = ["tabulate", "tabulate_into", "tabulate_host", "locate_subcells", "Tabulator", "get_tabulator", "TraceError"]

FORCE_THREAD_PER_POINT = 1
FORCE_DMMA = 2
FORCE_GENERAL = 4          # do not use the product-form (lattice) kernel
NO_VALUE_TABLE = 8         # do not use the value-table kernel (derivative-folded coefficients)
NO_ALPHA_SPLIT = 16        # do not split the tabulation into one derived order-0 element per alpha
NO_MERGED_SPLIT = 32       # keep the derived elements of a split as separate launches
NO_MACRO_MERGED = 64       # do not use the split-cell tile kernel (derived element of plan.macro_merged)
KERNEL_NAMES = {0: "none", 1: "cellwise", 2: "mma", 3: "small", 4: "vals", 5: "lattice", 6: "tensor", 7: "mma_cells"}


def _resolve_simplex_entity(desc, entity):
    sd = int(desc["sd"])
    if entity is None:
        entity = (sd, 0)
    dim, ent = int(entity[0]), int(entity[1])
    keys = numpy.asarray(desc["ent_keys"]).reshape(-1, 2)
    hit = numpy.where((keys[:, 0] == dim) & (keys[:, 1] == ent))[0]
    if len(hit) == 0:
        if dim == sd and ent == 0:
            return sd, None
        raise KeyError(f"no entity {(dim, ent)} on this reference cell")
    j = int(hit[0])
    C = numpy.asarray(desc["ent_C"][j][:dim], dtype=float)
    off = numpy.asarray(desc["ent_off"][j], dtype=float)
    if dim == sd and numpy.array_equal(C, numpy.eye(sd)) and not off.any():
        return sd, None
    return dim, (C, off)


class TraceError(Exception):
    """Tabulating a trace element on the interior of a cell, or its gradient (FIAT/hdiv_trace.py:25-31)."""

    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


def _barycentric(points, vertices):
    """FIAT/hdiv_trace.py:338-352."""
    T = (numpy.asarray(vertices[:-1]) - vertices[-1]).T
    bary = numpy.einsum("ij,kj->ki", numpy.linalg.inv(T), points - vertices[-1])
    return numpy.concatenate([bary, (1 - bary.sum(axis=1))[:, None]], axis=1)


def _extract_facets(coordinates, tolerance=1e-10):
    """FIAT/hdiv_trace.py:306-335: every point must lie on exactly one facet; facet i excludes vertex i, except on the
    interval, where point i is facet i."""
    facet_to_pts = collections.defaultdict(list)
    for ipt, c in enumerate(coordinates):
        on_facet = [i for i, lam in enumerate(c) if abs(lam) < tolerance]
        if len(on_facet) != 1:
            return {}, False
        facet_to_pts[on_facet[0]].append(ipt)
    if coordinates.shape[1] == 2:
        facet_to_pts[0], facet_to_pts[1] = facet_to_pts[1], facet_to_pts[0]
    return facet_to_pts, True


class _Plan:
    """Owns one device plan handle."""

    def __init__(self, handle, keep):
        self.handle = handle
        self.keep = keep
        self.force_flags = 0        # library flags every launch of this plan carries (set by the self-check)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().fiatb200_plan_destroy(self.handle)
        except Exception:
            pass


def _quick_description(desc):
    """The same element with every simplex description marked `dense_only`: plan.compile_simplex then builds only the
    recurrence tables and dense per-cell matrices (about a millisecond), none of the row clustering, block packing,
    value tables or derived elements (up to seconds) that pay off on large point sets."""
    kind = desc["kind"]
    if kind == "simplex":
        return dict(desc, dense_only=True)
    if kind == "tensor":
        return dict(desc, A=_quick_description(desc["A"]), B=_quick_description(desc["B"]))
    if kind == "flattened":
        return dict(desc, element=_quick_description(desc["element"]))
    if kind == "composite":
        return dict(desc, parts=[dict(part, element=_quick_description(part["element"])) for part in desc["parts"]])
    return desc


# Calls with at most this many points go to the quick plan (thread-per-point kernels, no plan-time optimisation) as long
# as no larger call has built the optimised plan for that (order, entity): the typical use of the reference --
# an element tabulated once at a quadrature rule -- then costs milliseconds instead of the seconds of plan time the
# streaming kernels need.  FIATB200_QUICK_NPTS=0 switches it off (the GPU tests do: they are about those kernels).
QUICK_NPTS = int(os.environ.get("FIATB200_QUICK_NPTS", "4096"))


class Tabulator:
    """Device tabulation of one element (description) on one CUDA device."""

    def __init__(self, desc, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fiat_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.desc = desc
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.kind = desc["kind"]
        self._plans = {}
        self._lock = threading.Lock()
        self._quick = None

    def _quick_for(self, order, entity, npts, flags):
        """The quick-plan tabulator for a small call, or None (see QUICK_NPTS)."""
        if flags or npts > QUICK_NPTS or self.kind in ("trace", "quadrature") or self.desc.get("dense_only"):
            return None
        if ("resolved", order, None if entity is None else str(entity), 0) in self._plans:
            return None                         # the optimised plan exists already
        with self._lock:
            if self._quick is None:
                self._quick = Tabulator(_quick_description(self.desc), self.device)
        return self._quick

    # -- shapes -------------------------------------------------------------------------------
    def cell_dimension(self):
        return planmod._cell_dim(self.desc)

    def value_shape(self):
        return planmod.value_shape_of(self.desc)

    def alphas(self, order):
        return planmod.alpha_list(self.cell_dimension(), order)

    # -- plans --------------------------------------------------------------------------------
    def _simplex_plan(self, desc, order):
        key = ("simplex", id(desc), order)
        with self._lock:
            if key not in self._plans:
                prog = planmod.compile_simplex(desc, order)
                struct, keep = _lib.simplex_struct(prog)
                handle = ctypes.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.fiatb200_simplex_plan_create(ctypes.byref(struct), ctypes.byref(handle)))
                self._plans[key] = (_Plan(handle, keep), prog)
            return self._plans[key]

    def _lattice_plan(self, desc, order):
        """Product-form plan for equispaced Lagrange elements, or None.  Selected only after it has
        reproduced the general kernel on the device."""
        key = ("lattice", id(desc), order)
        with self._lock:
            if key in self._plans:
                return self._plans[key]
        plan = None
        rowmap = planmod.lattice_rowmap(desc) if order <= 2 else None
        if rowmap is not None:
            keep = [numpy.ascontiguousarray(rowmap, dtype=numpy.int32)]
            handle = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_lattice_plan_create(
                    int(desc["sd"]), int(desc["degree"]), order, keep[0].ctypes.data_as(_lib.p_i32), len(rowmap),
                    ctypes.byref(handle)))
            cand = _Plan(handle, keep)
            # the product-form kernel only depends on (sd, degree, order): once it has reproduced the general kernel
            # on this device the verdict is kept for the process, and later elements skip the general plan
            vkey = (str(self.device), int(desc["sd"]), int(desc["degree"]), order)
            verdict = _LATTICE_VERDICT.get(vkey)
            if verdict is None:
                verdict = _LATTICE_VERDICT[vkey] = self._lattice_agrees(desc, order, cand)
            if verdict:
                plan = cand
        with self._lock:
            self._plans[key] = plan
        return plan

    def _lattice_agrees(self, desc, order, cand):
        sd = int(desc["sd"])
        general, prog = self._simplex_plan(desc, order)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(20261018)
        u, _ = torch.sort(torch.rand((96, sd), generator=gen, dtype=torch.float64), dim=1)
        pts = torch.diff(torch.cat([torch.zeros((96, 1), dtype=torch.float64), u], dim=1), dim=1)
        pts = pts.to(self.device).contiguous()
        ent = _lib.entity_struct(sd, None)
        a = torch.empty((prog.na, prog.nrows, 96), dtype=torch.float64, device=self.device)
        b = torch.empty_like(a)
        self._launch(general, ent, pts, a, 96, FORCE_GENERAL)
        self._launch(cand, ent, pts, b, 96, 0)
        scale = a.abs().amax(dim=(1, 2)).clamp_min(1e-300)
        err = (a - b).abs().amax(dim=(1, 2)) / scale
        return bool((err <= 1e-13).all().item())

    def _self_check_flags(self, desc, order):
        """On-device self-check of every path that replaces the derivative jets by host-folded derivative matrices
        (value-table kernel, per-alpha / stacked derived elements, split-cell tile kernel): on first use the default
        path and the thread-per-point jet kernel (FIAT/expansions.py:66-137 carried through the recurrence, no
        derived matrix) tabulate 96 points of the cell; if any table differs by more than SELF_CHECK_TOL of its
        largest entry the derived paths are switched off for this element and order.  -> flags to OR in."""
        key = ("selfcheck", id(desc), order)
        with self._lock:
            if key in self._plans:
                return self._plans[key]
        verdict = 0
        if desc["kind"] == "simplex" and desc["expansion"] == "dubiner" and int(desc["degree"]) >= 1 and order >= 0 \
                and not desc.get("raw_members"):
            sd = int(desc["sd"])
            main, prog = self._simplex_plan(desc, order)
            derived = [self.lib.fiatb200_plan_kernel(main.handle, 0) == 4]
            derived.append(self._alpha_split_plans(desc, order, 0) is not None)
            derived.append(self._macro_merged_plan(desc, order, 0) is not None)
            if any(derived):
                gen = torch.Generator(device="cpu")
                gen.manual_seed(20261019)
                lam = torch.rand((96, sd + 1), generator=gen, dtype=torch.float64) + 0.02
                lam = lam / lam.sum(dim=1, keepdim=True)
                verts = torch.as_tensor(numpy.asarray(desc["vertices"], dtype=numpy.float64)[:sd + 1])
                pts = (lam @ verts).to(self.device).contiguous()
                ent = _lib.entity_struct(sd, None)
                ref = torch.empty((prog.na, prog.nrows, 96), dtype=torch.float64, device=self.device)
                self._launch(main, ent, pts, ref, 96, FORCE_THREAD_PER_POINT)
                got = torch.empty_like(ref)
                self._run(self._default_launches(desc, order, ent), None, pts, got, 96, 96, 0)
                scale = ref.abs().amax(dim=(1, 2)).clamp_min(1e-300)
                err = (got - ref).abs().amax(dim=(1, 2)) / scale
                if not bool((err <= SELF_CHECK_TOL).all().item()):
                    verdict = NO_VALUE_TABLE | NO_ALPHA_SPLIT | NO_MACRO_MERGED
                    main.force_flags = NO_VALUE_TABLE
        with self._lock:
            self._plans[key] = verdict
        return verdict

    def _default_launches(self, desc, order, ent):
        """Launch list of the default (derived) path of one plain simplex element, as _resolve builds it."""
        p, _ = self._simplex_plan(desc, order)
        split = self._alpha_split_plans(desc, order, 0)
        macro = self._macro_merged_plan(desc, order, 0)
        if macro is not None:
            p = macro
        if split is None:
            return [(p, ent, None, 0)]
        if split[-1][0] == "merged":
            return [(split[-1][1], ent, None, 0)]
        out = []
        for j, sub, _ in split:
            if sub is not None:
                out.append((sub, ent, None, j))
        return out

    def _alpha_split_plans(self, desc, order, flags):
        """Per-alpha derived order-0 plans (plan.alpha_split) when the element would otherwise run on the DMMA
        tile kernel (or, for derivative orders it does not cover, the thread-per-point kernel) and the split
        stores fewer coefficient blocks; None otherwise.
        -> [(alpha index, plan or None for an identically zero table)]"""
        if flags & (FORCE_THREAD_PER_POINT | FORCE_DMMA | NO_ALPHA_SPLIT) or order < 1:
            return None
        key = ("split", id(desc), order, flags & 11)      # the result depends on the kernel the flags select
        with self._lock:
            if key in self._plans:
                return self._plans[key]
        main, prog = self._simplex_plan(desc, order)
        out = None
        if self.lib.fiatb200_plan_kernel(main.handle, flags & 11) in (1, 2):      # thread-per-point or DMMA tile
            derived = planmod.alpha_split(desc, order, prog)
            if derived is not None:
                out = []
                for j, (alpha, d) in enumerate(derived):
                    out.append((j, None if d is None else self._simplex_plan(d, 0)[0], d))
                # all derived elements stacked into one (rows = the derivative tables one after the other):
                # one launch, one value recurrence; usable when the part's rows are written in place
                merged = planmod.merged_split(desc, order, derived)
                if merged is not None:
                    out.append(("merged", self._simplex_plan(merged, 0)[0], merged))
        with self._lock:
            self._plans[key] = out
        return out

    def _macro_merged_plan(self, desc, order, flags):
        """Derived order-0 plan of a split-cell element for the split-cell tile kernel (plan.macro_merged), when
        the element would otherwise run thread-per-point; None otherwise."""
        if flags & (FORCE_THREAD_PER_POINT | FORCE_DMMA | NO_MACRO_MERGED) or int(desc.get("ncells", 1)) < 2:
            return None
        key = ("macro", id(desc), order, flags & 11)
        with self._lock:
            if key in self._plans:
                return self._plans[key][0]
        main, prog = self._simplex_plan(desc, order)
        out, derived = None, None
        # (value-table elements stay where they are: HCT degree 4 runs at 482 Gval/s there, 232 on the tile kernel)
        if self.lib.fiatb200_plan_kernel(main.handle, flags & 11) == 1:
            derived = planmod.macro_merged(desc, order, prog)
            if derived is not None:
                cand = self._simplex_plan(derived, 0)[0]
                if self.lib.fiatb200_plan_kernel(cand.handle, 0) == 7:
                    out = cand
        with self._lock:
            self._plans[key] = (out, derived)
        return out

    def kernel_names(self, order, entity=None, flags=0):
        """Names of the kernels the launches of `tabulate(order, ...)` run on (diagnostics)."""
        launches = self._resolve(order, entity, flags)[0]
        return [KERNEL_NAMES.get(self.lib.fiatb200_plan_kernel(p.handle, (flags & 11) | p.force_flags), "?") if p is not None else "zero"
                for p, _, _, _ in launches]

    def kernel_path(self, order, flags=0):
        """Which device path `tabulate(order, ...)` takes: 'lattice', 'simplex' or 'tensor'."""
        if self.kind == "composite" or (self.kind == "flattened" and self.desc["element"]["kind"] == "composite"):
            return "composite"
        if self.kind != "simplex":
            return "tensor"
        if flags & (FORCE_GENERAL | FORCE_THREAD_PER_POINT | FORCE_DMMA):
            return "simplex"
        return "lattice" if self._lattice_plan(self.desc, order) is not None else "simplex"

    def _tensor_plan(self, desc, order, entity):
        ekey = None if entity is None else (tuple(entity[0]) if isinstance(entity[0], (list, tuple)) else entity[0], entity[1])
        key = ("tensor", id(desc), order, ekey)
        with self._lock:
            cached = self._plans.get(key)
        if cached is not None:
            return cached
        leaves = planmod.flatten_tensor(desc, entity)
        if len(leaves) > 4:
            raise UnsupportedElement("tensor-product elements with more than 4 factors")
        if sum(1 for lf in leaves if len(lf.desc["value_shape"])) > 1:
            raise NotImplementedError("tabulate does not support two vector-valued inputs")
        arr = (_lib.TensorLeafStruct * len(leaves))()
        keep, nrows, npdim = [], 1, 0
        for i, lf in enumerate(leaves):
            p, prog = self._simplex_plan(lf.desc, order)
            dim, tr = _resolve_simplex_entity(lf.desc, lf.entity)
            arr[i].plan = p.handle
            arr[i].entity = _lib.entity_struct(lf.sd, tr)
            if tr is None:
                arr[i].entity.dim = dim
            arr[i].point_offset = lf.point_offset
            keep.append(p)
            nrows *= prog.nrows
            npdim = max(npdim, lf.point_offset + lf.point_dim)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fiatb200_tensor_plan_create(arr, len(leaves), order, ctypes.byref(handle)))
        out = (_Plan(handle, keep), nrows, npdim)
        with self._lock:
            self._plans[key] = out
        return out

    def _zero_rows(self, parts, ndofs, nc_out):
        """Device list of the table rows that no part writes (they must read as zero)."""
        covered = numpy.zeros((ndofs, nc_out), dtype=bool)
        for part in parts:
            nd = planmod.num_dofs_of(part.desc)
            covered[part.dof_base:part.dof_base + nd, part.comp_out] = True
        rows = numpy.flatnonzero(~covered.reshape(-1)).astype(numpy.int32)
        if len(rows) == 0:
            return None
        return torch.as_tensor(rows, device=self.device)

    def _resolve(self, order, entity, flags=0):
        """-> (launches, total_rows, point dimension, result shape prefix, rows to zero).

        launches = [(plan, entity struct or None, row map struct or None)]; a plain element has one
        launch without a row map, wrapper elements one per part."""
        if order < 0:
            raise ValueError("order must be non-negative")
        ckey = ("resolved", order, None if entity is None else str(entity), flags & 127)
        hit = self._plans.get(ckey)
        if hit is not None:
            return hit
        resolved = self._resolve_uncached(order, entity, flags)
        with self._lock:
            self._plans[ckey] = resolved
        return resolved

    def _resolve_uncached(self, order, entity, flags):
        parts = planmod.resolve_parts(self.desc, entity)
        vs = planmod.value_shape_of(self.desc)
        nc_out = int(numpy.prod(vs)) if vs else 1
        ndofs = planmod.num_dofs_of(self.desc)
        total_rows = ndofs * nc_out
        single = len(parts) == 1 and parts[0].dof_base == 0 and parts[0].comp_out == list(range(nc_out)) \
            and all(sg == 1.0 for sg in parts[0].sign)
        launches, pdim, zero_alpha_rows = [], None, []
        for part in parts:
            d = part.desc
            split = None
            if d["kind"] == "simplex":
                p, prog = self._simplex_plan(d, order)
                fast = None
                if not (flags & (FORCE_GENERAL | FORCE_THREAD_PER_POINT | FORCE_DMMA)):
                    fast = self._lattice_plan(d, order)
                    p = fast if fast is not None else p
                if fast is None:
                    if not flags & (FORCE_THREAD_PER_POINT | FORCE_DMMA | NO_VALUE_TABLE | NO_ALPHA_SPLIT | NO_MACRO_MERGED):
                        flags |= self._self_check_flags(d, order)
                    split = self._alpha_split_plans(d, order, flags)
                    macro = self._macro_merged_plan(d, order, flags) if single else None
                    if macro is not None:
                        p = macro
                dim, tr = _resolve_simplex_entity(d, part.entity)
                ent = _lib.entity_struct(prog.sd, tr)
            else:
                p, _, dim = self._tensor_plan(d, order, part.entity)
                ent = None
            if pdim is not None and dim != pdim:
                raise ValueError("parts of a wrapper element disagree on the point dimension")
            pdim = dim
            rmap = None if single else _lib.row_map_struct(len(part.comp_out), nc_out, part.dof_base, total_rows,
                                                            part.comp_out, part.sign)
            if split is None:
                launches.append((p, ent, rmap, 0))
            elif single and split[-1][0] == "merged" and not (flags & NO_MERGED_SPLIT):
                launches.append((split[-1][1], ent, None, 0))
            else:
                split = [item for item in split if item[0] != "merged"]
                # one derived order-0 element per alpha, written into that alpha's table
                nd = planmod.num_dofs_of(d)
                for j, sub, _ in split:
                    if sub is not None:
                        launches.append((sub, ent, rmap, j))
                    else:
                        rows = (numpy.arange(part.dof_base, part.dof_base + nd)[:, None] * nc_out
                                + numpy.asarray(part.comp_out)[None, :]).reshape(-1)
                        zero_alpha_rows.append((j, torch.as_tensor(rows.astype(numpy.int32), device=self.device)))
        zero = None
        if zero_alpha_rows:
            launches.extend((None, None, rows, j) for j, rows in zero_alpha_rows)
        if not single:
            key = ("zero", entity if entity is None else str(entity))
            with self._lock:
                if key not in self._plans:
                    self._plans[key] = self._zero_rows(parts, ndofs, nc_out)
                zero = self._plans[key]
        return launches, total_rows, pdim, (ndofs,) + vs, zero

    # -- calls --------------------------------------------------------------------------------
    def _single_point(self, points, pdim):
        """One point given as a bare coordinate tuple, shape (sd,): the reference's tables then have no point axis
        (CiarletElement only; test/FIAT/unit/test_fiat.py test_single_point_tabulation, test_regge_hhj.py)."""
        if self.kind != "simplex" or pdim == 0:
            return False
        shape = tuple(points.shape) if isinstance(points, (torch.Tensor, numpy.ndarray)) else numpy.shape(points)
        return shape == (pdim,)

    def _points(self, points, pdim):
        if isinstance(points, torch.Tensor):
            pts = points.to(device=self.device, dtype=torch.float64)
        else:
            arr = numpy.asarray(points)
            if arr.dtype == object:
                raise NotImplementedError("symbolic points are not tabulated on the device")
            pts = torch.as_tensor(numpy.ascontiguousarray(arr, dtype=numpy.float64), device=self.device)
        pts = pts.reshape(-1, pdim) if pts.numel() or pts.ndim < 2 else pts.reshape(pts.shape[0], pdim)
        return pts.contiguous()

    # -- elements that are not polynomial tabulations: HDivTrace, QuadratureElement ---------------------------------
    def _tabulate_quadrature(self, order, points, entity):
        """QuadratureElement.tabulate (FIAT/quadrature_element.py:43-61): the identity at its own points."""
        d = self.desc
        if entity is not None and tuple(entity) != (int(d["dim"]), 0):
            raise ValueError('QuadratureElement does not "tabulate" on subentities.')
        if order:
            raise ValueError("Derivatives are not defined on a QuadratureElement.")
        own = numpy.asarray(d["points"], dtype=float)
        pts = numpy.asarray(points.cpu() if isinstance(points, torch.Tensor) else points, dtype=float).reshape(len(points), -1)
        if len(pts) != len(own) or abs(pts - own).max() > 1e-12:
            raise AssertionError("Mismatch of quadrature points!")
        return {(0,) * int(d["sd"]): torch.eye(len(own), dtype=torch.float64, device=self.device)}

    def _tabulate_trace(self, order, points, entity):
        """HDivTrace.tabulate (FIAT/hdiv_trace.py:133-233): values only on facets -- the facet's discontinuous element
        tabulated on the facet and written into that facet's rows; derivative slots hold TraceError objects, like
        the reference.  entity=None identifies the facet of every point geometrically (simplices; tolerance 1e-10,
        :22,316-335) on the host, then tabulates facet by facet on the device."""
        d = self.desc
        sd, ndofs = int(d["sd"]), int(d["ndofs"])
        evalkey = (0,) * sd
        if order < 0:
            raise ValueError("order must be non-negative")
        npts = len(points)
        phivals = {}
        for alpha in self.alphas(order):
            phivals[alpha] = torch.zeros((ndofs, npts), dtype=torch.float64, device=self.device) if alpha == evalkey \
                else TraceError("Gradients on trace elements are not well-defined.")
        facet_sd = sd - 1
        facets = d["facets"]

        def fkey(f):
            return tuple(f["dim"]) if isinstance(f["dim"], list) else f["dim"]

        def facet_table(f, order_, pts_):
            if f["element"]["kind"] == "point":       # constant on a point facet
                vals = torch.as_tensor(numpy.asarray(f["element"]["values"], dtype=float), device=self.device)
                return vals[:, None].expand(-1, len(pts_))
            return get_tabulator(f["element"], self.device).tabulate(order_, pts_)[(0,) * facet_sd]

        if entity is None or tuple(entity) == (sd, 0):
            if not d["simplex"]:
                raise NotImplementedError("Tabulating this element on a non-simplex cell without providing an entity "
                                          "is not currently supported.")
            pts = numpy.asarray(points.cpu() if isinstance(points, torch.Tensor) else points, dtype=float).reshape(npts, sd)
            verts = numpy.asarray(d["vertices"], dtype=float)
            bary = _barycentric(pts, verts)
            facet_to_pts, success = _extract_facets(bary)
            if not success:
                for key in phivals:
                    if entity is None:
                        phivals[key] = torch.full((ndofs, npts), float("nan"), dtype=torch.float64, device=self.device)
                    else:
                        phivals[key] = TraceError("The HDivTrace element can only be tabulated on facets.")
            (f,) = [f for f in facets if fkey(f) == facet_sd]
            ref_verts = numpy.concatenate([numpy.zeros((1, facet_sd)), numpy.eye(facet_sd)]) if facet_sd else numpy.zeros((1, 0))
            for facet, ipts in facet_to_pts.items():
                coords = numpy.delete(bary[ipts], facet, axis=1)           # barycentric coordinates on the facet
                new_points = coords @ ref_verts
                vals = facet_table(f, order, new_points)
                idx = torch.as_tensor(ipts, device=self.device, dtype=torch.long)
                phivals[evalkey][f["nf"] * facet:f["nf"] * (facet + 1), idx] = vals
            return phivals
        edim = tuple(entity[0]) if isinstance(entity[0], (list, tuple)) else entity[0]
        if not any(fkey(f) == edim for f in facets):
            return {key: TraceError("The HDivTrace element can only be tabulated on facets.") for key in phivals}
        offset = 0
        for f in facets:
            for i in range(f["count"]):
                if (fkey(f), i) == (edim, int(entity[1])):
                    phivals[evalkey][offset:offset + f["nf"]] = facet_table(f, 0, points)
                offset += f["nf"]
        return phivals

    def tabulate(self, order, points, entity=None, flags=0):
        if self.kind == "quadrature":
            return self._tabulate_quadrature(order, points, entity)
        if self.kind == "trace":
            return self._tabulate_trace(order, points, entity)
        quick = self._quick_for(order, entity, len(points) if hasattr(points, "__len__") else QUICK_NPTS + 1, flags)
        if quick is not None:
            return quick.tabulate(order, points, entity, FORCE_THREAD_PER_POINT)
        launches, nrows, pdim, prefix, zero = self._resolve(order, entity, flags)
        pts = self._points(points, pdim)
        npts = pts.shape[0]
        alphas = self.alphas(order)
        out = torch.empty((len(alphas), nrows, npts), dtype=torch.float64, device=self.device)
        self._run(launches, zero, pts, out, npts, npts, flags)
        if self._single_point(points, pdim):
            return {a: out[j].reshape(prefix) for j, a in enumerate(alphas)}
        return {a: out[j].reshape(prefix + (npts,)) for j, a in enumerate(alphas)}

    def tabulate_into(self, out, order, points, entity=None, flags=0):
        """Streaming form: write into a caller-owned (nalpha, nrows, >=npts) float64 cuda tensor
        (row stride = out.stride(1)); returns the number of points written."""
        launches, nrows, pdim, _, zero = self._resolve(order, entity, flags)
        pts = self._points(points, pdim)
        npts = pts.shape[0]
        na = len(self.alphas(order))
        if out.dtype != torch.float64 or out.device != self.device or out.ndim != 3 or out.shape[0] != na \
                or out.shape[1] != nrows or out.shape[2] < npts or out.stride(2) != 1 \
                or out.stride(0) != nrows * out.stride(1):
            raise ValueError("out must be a float64 cuda tensor (nalpha, nrows, >=npts) with unit point stride")
        self._run(launches, zero, pts, out, npts, out.stride(1), flags)
        return npts

    def _run(self, launches, zero, pts, out, npts, row_stride, flags):
        if npts == 0:
            return
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ld = pts.stride(0) if pts.shape[1] else 0
        cflags = flags & 11
        with torch.cuda.device(self.device):
            if zero is not None:
                _lib.check(self.lib.fiatb200_zero_rows(out.data_ptr(), row_stride, npts, out.shape[1], out.shape[0],
                                                        zero.data_ptr(), zero.numel(), stream))
            for p, ent, rmap, aoff in launches:
                # aoff: derivative table this launch writes (launches of per-alpha derived elements)
                optr = out.data_ptr() + 8 * aoff * out.shape[1] * row_stride
                if p is None:       # identically zero derivative table of a per-alpha split
                    _lib.check(self.lib.fiatb200_zero_rows(optr, row_stride, npts, out.shape[1], 1,
                                                            rmap.data_ptr(), rmap.numel(), stream))
                    continue
                _lib.check(self.lib.fiatb200_tabulate_mapped(
                    p.handle, ctypes.byref(ent) if ent is not None else None, pts.data_ptr(), npts, ld,
                    optr, row_stride, ctypes.byref(rmap) if rmap is not None else None, cflags | p.force_flags, stream))

    def _launch(self, p, ent, pts, out, npts, flags, row_stride=None):
        self._run([(p, ent, None, 0)], None, pts, out, npts, npts if row_stride is None else row_stride, flags)

    def tabulate_host(self, order, points, entity=None, chunk_pts=1 << 16, flags=0, out=None):
        """End-to-end with host (numpy) buffers: returns a dict of numpy arrays, like the reference."""
        if self.kind in ("trace", "quadrature"):
            # not polynomial tabulations: facet by facet / the identity (slots that are not defined hold TraceError)
            res = self.tabulate(order, points, entity)
            return {a: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for a, v in res.items()}
        quick = self._quick_for(order, entity, len(points) if hasattr(points, "__len__") else QUICK_NPTS + 1, flags)
        if quick is not None:
            return quick.tabulate_host(order, points, entity, chunk_pts, FORCE_THREAD_PER_POINT, out)
        launches, nrows, pdim, prefix, zero = self._resolve(order, entity, flags)
        pts = numpy.ascontiguousarray(numpy.asarray(points, dtype=numpy.float64)).reshape(-1, pdim)
        npts = pts.shape[0]
        alphas = self.alphas(order)
        if out is None:
            # page-locked result (recycled by torch's caching host allocator): the device-to-host copies then run at
            # PCIe speed, 2.8 x the rate into pageable memory (bench.py: e2e vs e2e.pageable_value)
            out = torch.empty((len(alphas), nrows, npts), dtype=torch.float64, pin_memory=True).numpy()
        elif not (isinstance(out, numpy.ndarray) and out.dtype == numpy.float64 and out.flags.c_contiguous
                  and out.flags.writeable and out.shape == (len(alphas), nrows, npts)):
            # the library writes (nalpha * nrows) rows of npts doubles through the raw pointer
            raise ValueError(f"out must be a writeable C-contiguous float64 ndarray of shape {(len(alphas), nrows, npts)}")
        if npts and len(launches) == 1 and launches[0][2] is None:
            p, ent, _, _ = launches[0]
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_tabulate_host(
                    p.handle, ctypes.byref(ent) if ent is not None else None, pts.ctypes.data, npts, pdim,
                    out.ctypes.data, chunk_pts, (flags & 11) | p.force_flags))
        elif npts:
            # wrapper elements / per-alpha splits: the launch list goes through the same staged pipeline
            arr = (_lib.LaunchStruct * len(launches))()
            for i, (p, ent, rmap, aoff) in enumerate(launches):
                arr[i].alpha_offset = aoff
                if p is None:
                    arr[i].plan, arr[i].zero_rows_dev, arr[i].nzero_rows = None, rmap.data_ptr(), rmap.numel()
                else:
                    arr[i].plan = p.handle
                    arr[i].entity = ctypes.pointer(ent) if ent is not None else None
                    arr[i].map = ctypes.pointer(rmap) if rmap is not None else None
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_tabulate_host_list(
                    arr, len(launches), len(alphas), nrows, zero.data_ptr() if zero is not None else None,
                    zero.numel() if zero is not None else 0, pts.ctypes.data, npts, pdim, out.ctypes.data, chunk_pts,
                    (flags & 11) | functools.reduce(lambda a, b: a | b, (p.force_flags for p, _, _, _ in launches
                                                                         if p is not None), 0)))
        if self._single_point(points, pdim):
            return {a: out[j].reshape(prefix) for j, a in enumerate(alphas)}
        return {a: out[j].reshape(prefix + (npts,)) for j, a in enumerate(alphas)}

    def tabulate_factors(self, order, points, entity=None):
        """Factored result for tensor-product elements: the per-factor tables instead of their outer
        product (what FInAT's sum-factorisation consumes, finat/tensor_product.py:98-144), which avoids
        writing prod(n_l) values per point.  Returns a list with one entry per leaf factor, in dof-major
        order: (alpha_offset, sd_l, {alpha_l: tensor (n_l, *value_shape_l, npts)}); the full table is
        out[alpha][(i0 * n1 + i1) * n2 + i2] = prod_l tab_l[alpha[slice_l]][i_l]."""
        d = self.desc
        if d["kind"] not in ("tensor", "flattened") or (d["kind"] == "flattened" and d["element"]["kind"] != "tensor"):
            raise UnsupportedElement("tabulate_factors() is defined for tensor-product elements")
        leaves = planmod.flatten_tensor(d, entity)
        pdim = max(lf.point_offset + lf.point_dim for lf in leaves)
        pts = self._points(points, pdim)
        out, aoff = [], 0
        for lf in leaves:
            sub = get_tabulator(lf.desc, self.device)
            sl = pts[:, lf.point_offset:lf.point_offset + lf.point_dim]
            out.append((aoff, lf.sd, sub.tabulate(order, sl, lf.entity)))
            aoff += lf.sd
        return out

    def evaluate(self, coefficients, order, points, entity=None):
        """Fused consumer: derivatives of the finite-element functions u_f = sum_i coefficients[f, i] phi_i
        at the points, without ever writing the (ndofs, npts) tables (point evaluation /
        interpolation).  Returns {alpha: tensor (nfunc, *value_shape, npts)}.

        Tabulation is linear in the coefficient tensor (FIAT/polynomial_set.py:71): the functions' tables are the
        tables of the weights coefficients . C, which the library forms on the device in the same call
        (`fiatb200_evaluate_simplex`), so a new coefficient vector needs no re-planning.  Scalar tensor-product
        elements nest the sum over the factors (`fiatb200_evaluate_tensor`); wrapper elements (enriched, mixed,
        H(div)/H(curl) on tensor products) add up their parts' evaluations with the parts' dof slices, component
        placement and signs (FIAT/enriched.py:88-113, mixed.py:61-92, hdivcurl.py)."""
        out, alphas, shape = self._evaluate_device(coefficients, order, points, entity)
        return {a: out[j].reshape(shape) for j, a in enumerate(alphas)}

    def _coefficients(self, coefficients, ndofs):
        if isinstance(coefficients, torch.Tensor):
            u = coefficients.to(device=self.device, dtype=torch.float64)
            u = u.reshape(1, -1) if u.ndim == 1 else u
        else:
            u = torch.as_tensor(numpy.ascontiguousarray(numpy.atleast_2d(numpy.asarray(coefficients, dtype=numpy.float64))),
                                device=self.device)
        if u.ndim != 2 or u.shape[1] != ndofs:
            raise ValueError(f"expected {ndofs} coefficients per function, got {tuple(u.shape)}")
        return u.contiguous()

    def _eval_plan(self, desc, order):
        """Plan of the stacked derived element (plan.stacked_derived) -> (plan, nstack, ndofs, ncomp), or None."""
        key = ("eval", id(desc), order)
        with self._lock:
            if key in self._plans:
                return self._plans[key]
        out = None
        if desc["kind"] == "simplex" and desc["expansion"] == "dubiner":
            _, prog = self._simplex_plan(desc, order)
            stacked = planmod.stacked_derived(desc, order, prog)
            if stacked is not None:
                p = self._simplex_plan(stacked, 0)[0]
                out = (p, prog.na, prog.ndofs, prog.nrows // max(prog.ndofs, 1), stacked)
        with self._lock:
            self._plans[key] = out
        return out

    def _evaluate_device(self, coefficients, order, points, entity):
        """-> (tensor (nalpha, nfunc * ncomp, npts), alphas, (nfunc, *value_shape, npts))"""
        if order < 0:
            raise ValueError("order must be non-negative")
        alphas = self.alphas(order)
        vs = planmod.value_shape_of(self.desc)
        nc_out = int(numpy.prod(vs)) if vs else 1
        ndofs = planmod.num_dofs_of(self.desc)
        u = self._coefficients(coefficients, ndofs)
        nfunc = u.shape[0]
        parts = planmod.resolve_parts(self.desc, entity)
        single = len(parts) == 1 and parts[0].dof_base == 0 and parts[0].comp_out == list(range(nc_out)) \
            and all(sg == 1.0 for sg in parts[0].sign)
        total = None
        for part in parts:
            nd = planmod.num_dofs_of(part.desc)
            up = u if single else u[:, part.dof_base:part.dof_base + nd].contiguous()
            if part.desc["kind"] == "simplex":
                got = self._evaluate_simplex_part(part.desc, up, order, points, part.entity)
            else:
                got = self._evaluate_tensor_part(part.desc, up, order, points, part.entity)
            if single:
                total = got
                break
            # got: (nalpha, nfunc * nc_in, npts) -> components comp_out of the wrapper's value, with signs
            npts = got.shape[-1]
            if total is None:
                total = torch.zeros((len(alphas), nfunc, nc_out, npts), dtype=torch.float64, device=self.device)
            g = got.reshape(len(alphas), nfunc, len(part.comp_out), npts)
            for k, (c, sg) in enumerate(zip(part.comp_out, part.sign)):
                total[:, :, c, :] += sg * g[:, :, k, :]
        npts = total.shape[-1]
        return total.reshape(len(alphas), nfunc * nc_out, npts), alphas, (nfunc,) + vs + (npts,)

    def _evaluate_simplex_part(self, desc, u, order, points, entity):
        sd = int(desc["sd"])
        dim, tr = _resolve_simplex_entity(desc, entity)
        ent = _lib.entity_struct(sd, tr)
        pts = self._points(points, dim)
        npts = pts.shape[0]
        nfunc = u.shape[0]
        ep = self._eval_plan(desc, order)
        if ep is None:
            # sets without derivative matrices (1-D Legendre / Lagrange lines, degree 0, very high degrees): tabulate
            # the derived element whose coefficient tensor is u . coeffs (planned per coefficient set)
            coeffs = numpy.asarray(desc["coeffs"], dtype=numpy.float64)
            derived = dict(desc)
            derived["coeffs"] = numpy.einsum("fd,dck->fck", u.cpu().numpy(), coeffs)
            derived.pop("nodes", None)
            tab = Tabulator(derived, self.device).tabulate(order, points, entity)
            return torch.stack([v.reshape(nfunc * coeffs.shape[1], npts) for v in tab.values()])
        p, nstack, ndofs, ncomp, _ = ep
        out = torch.empty((nstack, nfunc * ncomp, npts), dtype=torch.float64, device=self.device)
        if npts:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_evaluate_simplex(
                    p.handle, nstack, ndofs, u.data_ptr(), nfunc, ctypes.byref(ent), pts.data_ptr(), npts,
                    pts.stride(0) if pts.shape[1] else 0, out.data_ptr(), npts, stream))
        return out

    def _evaluate_tensor_part(self, desc, u, order, points, entity):
        if planmod.value_shape_of(desc):
            # one vector-valued factor (FIAT/tensor_product.py:293-335): contract the table (no nested-sum kernel yet)
            tab = get_tabulator(desc, self.device).tabulate(order, points, entity)
            return torch.stack([torch.einsum("fd,dcp->fcp", u, v.reshape(v.shape[0], -1, v.shape[-1])).reshape(
                u.shape[0] * (v.numel() // (v.shape[0] * v.shape[-1])), v.shape[-1]) for v in tab.values()])
        p, nrows, pdim = self._tensor_plan(desc, order, entity)
        pts = self._points(points, pdim)
        npts = pts.shape[0]
        out = torch.empty((len(self.alphas(order)), u.shape[0], npts), dtype=torch.float64, device=self.device)
        if npts:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_evaluate_tensor(p.handle, u.data_ptr(), u.shape[0], pts.data_ptr(), npts,
                                                              pts.stride(0) if pts.shape[1] else 0, out.data_ptr(), npts,
                                                              stream))
        return out

    def evaluate_host(self, coefficients, order, points, entity=None, out=None, chunk_pts=1 << 17):
        """evaluate() with host (numpy) buffers: points in, {alpha: ndarray (nfunc, *value_shape, npts)} out, chunked and
        double-buffered inside the library (`fiatb200_evaluate_host`).  Per point 8 * dim bytes go to the device and
        8 * nalpha * nfunc * ncomp bytes come back -- not the tables.  Plain Ciarlet elements and scalar tensor-product
        elements take the pipelined path; wrapper elements go through evaluate() and one copy."""
        alphas = self.alphas(order)
        vs = planmod.value_shape_of(self.desc)
        nc = int(numpy.prod(vs)) if vs else 1
        ndofs = planmod.num_dofs_of(self.desc)
        u = numpy.ascontiguousarray(numpy.atleast_2d(numpy.asarray(coefficients, dtype=numpy.float64)))
        if u.shape[1] != ndofs:
            raise ValueError(f"expected {ndofs} coefficients per function, got {u.shape}")
        nfunc = u.shape[0]
        parts = planmod.resolve_parts(self.desc, entity)
        plain = len(parts) == 1 and parts[0].dof_base == 0 and parts[0].comp_out == list(range(nc)) \
            and all(sg == 1.0 for sg in parts[0].sign)
        handle = ent = None
        if plain and parts[0].desc["kind"] == "simplex":
            ep = self._eval_plan(parts[0].desc, order)
            if ep is not None:
                handle, nstack = ep[0].handle, ep[1]
                pdim, tr = _resolve_simplex_entity(parts[0].desc, parts[0].entity)
                ent = _lib.entity_struct(int(parts[0].desc["sd"]), tr)
        elif plain and not vs:
            p, nrows, pdim = self._tensor_plan(parts[0].desc, order, parts[0].entity)
            handle, nstack = p.handle, len(alphas)
        if handle is None:
            dev, _, shape = self._evaluate_device(u, order, points, entity)
            res = dev.cpu().numpy()
            if out is not None:
                out[...] = res.reshape(out.shape)
                res = out
            return {a: res[j].reshape(shape) for j, a in enumerate(alphas)}
        pts = numpy.ascontiguousarray(numpy.asarray(points, dtype=numpy.float64)).reshape(-1, pdim)
        npts = pts.shape[0]
        if out is None:
            out = torch.empty((len(alphas), nfunc * nc, npts), dtype=torch.float64, pin_memory=True).numpy()
        elif not (isinstance(out, numpy.ndarray) and out.dtype == numpy.float64 and out.flags.c_contiguous
                  and out.flags.writeable and out.size == len(alphas) * nfunc * nc * npts):
            raise ValueError(f"out must be a writeable C-contiguous float64 ndarray with {len(alphas) * nfunc * nc * npts} entries")
        if npts:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_evaluate_host(
                    handle, nstack, ndofs, u.ctypes.data, nfunc, ctypes.byref(ent) if ent is not None else None,
                    pts.ctypes.data, npts, pdim, out.ctypes.data, chunk_pts))
        flat = out.reshape(len(alphas), nfunc * nc, npts)
        return {a: flat[j].reshape((nfunc,) + vs + (npts,)) for j, a in enumerate(alphas)}

    def mapped(self, mapping, J=None, Jinv=None, Jdet=None):
        """Tabulator of the pulled-back element: its tables are `pullback(tabulate(...), mapping, J, Jinv, Jdet)` of
        FIAT/macro.py:601-645 (Piola maps of a constant Jacobian), folded into the coefficient tensor on the host
        (plan.pullback_description) -- same kernels, no extra work per point."""
        return Tabulator(planmod.pullback_description(self.desc, mapping, J, Jinv, Jdet), self.device)

    def locate_subcells(self, points, unique, entity=None):
        """Bitmask (uint32 as int64 tensor) of the subcells each point is binned to."""
        if self.kind != "simplex":
            raise UnsupportedElement("subcell location is defined for simplicial complexes")
        p, prog = self._simplex_plan(self.desc, 0)
        dim, tr = _resolve_simplex_entity(self.desc, entity)
        ent = _lib.entity_struct(prog.sd, tr)
        pts = self._points(points, dim)
        mask = torch.zeros(pts.shape[0], dtype=torch.int32, device=self.device)
        if pts.shape[0]:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fiatb200_locate_subcells(p.handle, ctypes.byref(ent), pts.data_ptr(), pts.shape[0],
                                                             pts.stride(0) if dim else 0, int(bool(unique)),
                                                             mask.data_ptr(), stream))
        return mask


_LATTICE_VERDICT = {}      # (device, sd, degree, order) -> the product-form kernel reproduced the general one
SELF_CHECK_TOL = 2e-13     # derived paths must reproduce the jet kernel this well (relative to each table's max), else jets
#                            (the thread-per-point jet kernel stays within 2e-15 of the reference up to degree 12)
_cache_lock = threading.RLock()
_by_element = weakref.WeakKeyDictionary()
_by_desc = collections.OrderedDict()        # description dicts are not weak-referenceable: bounded LRU instead
MAX_CACHED_DESCRIPTIONS = 64


def get_tabulator(element, device=None):
    """Cached Tabulator of a FIAT element object or of an element description dict."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if isinstance(element, dict):
        key = (id(element), str(dev))
        with _cache_lock:
            hit = _by_desc.get(key)
            if hit is None or hit[0] is not element:
                hit = (element, Tabulator(element, dev))
                _by_desc[key] = hit
                while len(_by_desc) > MAX_CACHED_DESCRIPTIONS:
                    _by_desc.popitem(last=False)
            else:
                _by_desc.move_to_end(key)
            return hit[1]
    with _cache_lock:
        try:
            per_dev = _by_element.setdefault(element, {})
        except TypeError:          # not weak-referenceable
            per_dev = element.__dict__.setdefault("_fiat_b200_tabulators", {})
        tab = per_dev.get(str(dev))
    if tab is None:
        # described OUTSIDE the lock: describe_element may tabulate sub-elements of the reference (the point facets
        # of an interval trace, extract._describe_trace), and where the reference's `tabulate` is bound to this
        # library -- the maintainer's integration, oracle/dropin_plugin.py -- that call comes back in here
        new = Tabulator(describe_element(element), dev)
        with _cache_lock:
            tab = per_dev.setdefault(str(dev), new)
    return tab


def tabulate(element, order, points, entity=None, device=None):
    """Device drop-in for `element.tabulate(order, points, entity)`."""
    return get_tabulator(element, device).tabulate(order, points, entity)


def tabulate_into(out, element, order, points, entity=None, device=None):
    return get_tabulator(element, device).tabulate_into(out, order, points, entity)


def tabulate_host(element, order, points, entity=None, device=None, chunk_pts=1 << 16):
    return get_tabulator(element, device).tabulate_host(order, points, entity, chunk_pts=chunk_pts)


def locate_subcells(element, points, unique, device=None):
    return get_tabulator(element, device).locate_subcells(points, unique)
