"""ctypes binding of the C ABI declared in include/fiat_b200.h.

The library is built in-tree (fiat_b200/csrc/libfiat_b200.so).  There is no fallback: if the
library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfiat_b200.so")

c_i32, c_i64, c_u32, c_dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_double
p_i32, p_dbl, p_void = ctypes.POINTER(c_i32), ctypes.POINTER(c_dbl), ctypes.c_void_p


class SimplexProgramStruct(ctypes.Structure):
    _fields_ = [
        ("sd", c_i32), ("degree", c_i32), ("order", c_i32), ("na", c_i32), ("expansion", c_i32),
        ("ncells", c_i32), ("nslots", c_i32), ("nrows", c_i32), ("ncomp", c_i32), ("unique", c_i32),
        ("nsteps", c_i32), ("nlevels", c_i32), ("nfix", c_i32), ("nfixgrp", c_i32), ("line_n", c_i32),
        ("start_slot", c_i32),
        ("step_idx", p_i32), ("step_abc", p_dbl), ("nat_abc", p_dbl), ("level_ptr", p_i32),
        ("fix_idx", p_i32), ("fix_w", p_dbl), ("fix_grp", p_i32),
        ("geom", p_dbl), ("bary", p_dbl), ("ccell", p_dbl), ("ccell_morton", p_dbl),
        ("low1", p_i32), ("mul1", p_dbl), ("low2", p_i32), ("mul2", p_dbl),
        ("line_tab", p_dbl), ("line_tab_len", c_i64),
        ("nrb", c_i32), ("kpad", c_i32), ("nblk", c_i32),
        ("blk_ptr", p_i32), ("blk_kb", p_i32), ("blk_frag", p_dbl), ("rb_order", p_i32), ("row_perm", p_i32),
        ("cderiv", p_dbl), ("cderiv_len", c_i64), ("ncp", c_i32), ("blk_cells", c_i32),
        ("cstream", p_dbl), ("cstream_len", c_i64), ("cstep_ptr", p_i32), ("cnsteps", c_i32), ("crb", c_i32),
    ]


class EntityMapStruct(ctypes.Structure):
    _fields_ = [("dim", c_i32), ("identity", c_i32), ("C", c_dbl * 9), ("offset", c_dbl * 3)]


class TensorLeafStruct(ctypes.Structure):
    _fields_ = [("plan", p_void), ("entity", EntityMapStruct), ("point_offset", c_i32)]


class RowMapStruct(ctypes.Structure):
    _fields_ = [("nc_in", c_i32), ("nc_out", c_i32), ("dof_base", c_i32), ("total_rows", c_i32),
                ("comp_out", c_i32 * 9), ("sign", c_dbl * 9)]


def row_map_struct(nc_in, nc_out, dof_base, total_rows, comp_out, sign):
    m = RowMapStruct()
    m.nc_in, m.nc_out, m.dof_base, m.total_rows = nc_in, nc_out, dof_base, total_rows
    co = list(comp_out) + [0] * (9 - len(comp_out))
    sg = list(sign) + [1.0] * (9 - len(sign))
    m.comp_out = (c_i32 * 9)(*[int(v) for v in co])
    m.sign = (c_dbl * 9)(*[float(v) for v in sg])
    return m


class LaunchStruct(ctypes.Structure):
    _fields_ = [("plan", p_void), ("entity", ctypes.POINTER(EntityMapStruct)), ("map", ctypes.POINTER(RowMapStruct)),
                ("alpha_offset", c_i32), ("zero_rows_dev", p_void), ("nzero_rows", c_i32)]


class LibraryError(RuntimeError):
    pass


class UnsupportedByLibrary(LibraryError, NotImplementedError):
    """FIATB200_ERR_UNSUPPORTED: a size or element the device path does not take (expansion degree above the device
    tables, more than 32 subcells, ...); callers that have the reference at hand may fall back on it."""


_lib = None

EXPORTS = [
    "fiatb200_version", "fiatb200_last_error", "fiatb200_simplex_plan_create", "fiatb200_tensor_plan_create",
    "fiatb200_lattice_plan_create", "fiatb200_plan_destroy", "fiatb200_plan_shape", "fiatb200_plan_kernel", "fiatb200_tabulate", "fiatb200_tabulate_mapped", "fiatb200_zero_rows", "fiatb200_locate_subcells",
    "fiatb200_tabulate_host", "fiatb200_tabulate_host_list", "fiatb200_evaluate_tensor", "fiatb200_launch_count",
    "fiatb200_cluster_rows", "fiatb200_colour_members", "fiatb200_evaluate_simplex", "fiatb200_evaluate_host",
]


def load():
    """Load libfiat_b200.so (once).  Raises if it has not been built -- no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(f"{LIB_PATH} is missing: build it with fiat_b200/csrc/build.sh "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.fiatb200_version.restype = ctypes.c_int
    lib.fiatb200_last_error.restype = ctypes.c_char_p
    lib.fiatb200_launch_count.restype = c_i64
    lib.fiatb200_simplex_plan_create.argtypes = [ctypes.POINTER(SimplexProgramStruct), ctypes.POINTER(p_void)]
    lib.fiatb200_tensor_plan_create.argtypes = [ctypes.POINTER(TensorLeafStruct), c_i32, c_i32, ctypes.POINTER(p_void)]
    lib.fiatb200_lattice_plan_create.argtypes = [c_i32, c_i32, c_i32, p_i32, c_i32, ctypes.POINTER(p_void)]
    lib.fiatb200_plan_destroy.argtypes = [p_void]
    lib.fiatb200_plan_kernel.argtypes = [p_void, c_u32]
    lib.fiatb200_plan_shape.argtypes = [p_void, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]
    lib.fiatb200_tabulate.argtypes = [p_void, ctypes.POINTER(EntityMapStruct), p_void, c_i64, c_i64, p_void, c_i64,
                                      c_u32, p_void]
    lib.fiatb200_tabulate_mapped.argtypes = [p_void, ctypes.POINTER(EntityMapStruct), p_void, c_i64, c_i64, p_void, c_i64,
                                             ctypes.POINTER(RowMapStruct), c_u32, p_void]
    lib.fiatb200_zero_rows.argtypes = [p_void, c_i64, c_i64, c_i64, c_i32, p_void, c_i32, p_void]
    lib.fiatb200_locate_subcells.argtypes = [p_void, ctypes.POINTER(EntityMapStruct), p_void, c_i64, c_i64, c_i32,
                                             p_void, p_void]
    lib.fiatb200_tabulate_host.argtypes = [p_void, ctypes.POINTER(EntityMapStruct), p_void, c_i64, c_i64, p_void,
                                           c_i64, c_u32]
    lib.fiatb200_evaluate_tensor.argtypes = [p_void, p_void, c_i32, p_void, c_i64, c_i64, p_void, c_i64, p_void]
    lib.fiatb200_tabulate_host_list.argtypes = [ctypes.POINTER(LaunchStruct), c_i32, c_i32, c_i64, p_void, c_i32, p_void,
                                                c_i64, c_i64, p_void, c_i64, c_u32]
    lib.fiatb200_evaluate_simplex.argtypes = [p_void, c_i32, c_i32, p_void, c_i32, ctypes.POINTER(EntityMapStruct), p_void,
                                              c_i64, c_i64, p_void, c_i64, p_void]
    lib.fiatb200_evaluate_host.argtypes = [p_void, c_i32, c_i32, p_void, c_i32, ctypes.POINTER(EntityMapStruct), p_void,
                                           c_i64, c_i64, p_void, c_i64]
    p_u8 = ctypes.POINTER(ctypes.c_uint8)
    lib.fiatb200_cluster_rows.argtypes = [p_u8, c_i32, c_i32, c_i32, c_i32, p_i32, c_i64, ctypes.c_uint64, p_i32]
    lib.fiatb200_colour_members.argtypes = [p_u8, c_i32, c_i32, c_i32, p_i32, c_i64, ctypes.c_uint64, p_i32, p_i32]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().fiatb200_last_error()
        text = f"fiat_b200 error {rc}: {msg.decode() if msg else '?'}"
        raise UnsupportedByLibrary(text) if rc == 3 else LibraryError(text)


def _ptr(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def simplex_struct(prog):
    """ctypes view of a plan.SimplexProgram; returns (struct, keepalive list)."""
    keep = []

    def f64(a):
        a = numpy.ascontiguousarray(a, dtype=numpy.float64)
        keep.append(a)
        return _ptr(a, c_dbl)

    def i32(a):
        a = numpy.ascontiguousarray(a, dtype=numpy.int32)
        keep.append(a)
        return _ptr(a, c_i32)

    s = SimplexProgramStruct()
    s.sd, s.degree, s.order, s.na, s.expansion = prog.sd, prog.degree, prog.order, prog.na, prog.expansion
    s.ncells, s.nslots, s.nrows, s.unique = prog.ncells, prog.nslots, prog.nrows, prog.unique
    s.ncomp = prog.nrows // max(prog.ndofs, 1)
    s.nsteps, s.nlevels, s.nfix, s.line_n = len(prog.step_idx), len(prog.level_ptr) - 1, len(prog.fix_idx), prog.line_n
    s.nfixgrp = len(prog.fix_grp)
    s.start_slot = prog.start_slot
    s.step_idx, s.step_abc, s.level_ptr = i32(prog.step_idx), f64(prog.step_abc), i32(prog.level_ptr)
    s.nat_abc, s.ccell_morton = f64(prog.nat_abc), f64(prog.ccell_morton)
    s.fix_idx, s.fix_w, s.fix_grp = i32(prog.fix_idx), f64(prog.fix_w), i32(prog.fix_grp)
    s.geom, s.bary, s.ccell = f64(prog.geom), f64(prog.bary), f64(prog.ccell)
    s.low1, s.mul1, s.low2, s.mul2 = i32(prog.low1), f64(prog.mul1), i32(prog.low2), f64(prog.mul2)
    s.line_tab, s.line_tab_len = f64(prog.line_tab), int(numpy.size(prog.line_tab))
    s.blk_cells = int(prog.blk_cells)
    s.nrb = len(prog.blk_ptr) // prog.blk_cells - 1 if prog.blk_cells > 1 else len(prog.blk_ptr) - 1
    s.kpad, s.nblk = prog.kpad, len(prog.blk_kb) // 4
    s.blk_ptr, s.blk_kb, s.blk_frag, s.rb_order = i32(prog.blk_ptr), i32(prog.blk_kb), f64(prog.blk_frag), i32(prog.rb_order)
    s.row_perm = i32(prog.row_perm if len(prog.row_perm) else numpy.arange(prog.nrows))
    s.cderiv, s.cderiv_len, s.ncp = f64(prog.cderiv), int(numpy.size(prog.cderiv)), int(prog.ncp)
    s.cstream, s.cstream_len = f64(prog.cstream), int(numpy.size(prog.cstream))
    s.cstep_ptr, s.cnsteps, s.crb = i32(prog.cstep_ptr), len(prog.cstep_ptr) - 1, int(prog.crb)
    return s, keep


def entity_struct(sd, transform):
    """transform: None (identity) or (C (dim x sd), offset (sd,))."""
    e = EntityMapStruct()
    if transform is None:
        e.dim, e.identity = sd, 1
        return e
    C, off = transform
    C = numpy.asarray(C, dtype=float).reshape(-1, sd)
    e.dim, e.identity = C.shape[0], 0
    flat = numpy.zeros(9)
    flat[:C.size] = C.reshape(-1)
    e.C = (c_dbl * 9)(*flat)
    o = numpy.zeros(3)
    o[:sd] = off
    e.offset = (c_dbl * 3)(*o)
    return e
