"""Host-side plan compiler checked on CPU: the compiled recurrence program, fix-ups, folded
coefficients and block packing reproduce the reference outputs stored in the golden fixtures."""
import numpy
import pytest

from conftest import golden_case_names, is_polynomial_case, load_case, tolerance
from fiat_b200 import plan as planmod
from oracle import fiat_oracle
import program_emulator as emu


def _simplex_dubiner_cases():
    out = []
    for name in golden_case_names():
        case = load_case(name)
        d = case["desc"]
        if d["kind"] == "simplex" and d["expansion"] == "dubiner":
            out.append(name)
    return out


@pytest.mark.parametrize("name", _simplex_dubiner_cases())
def test_program_reproduces_reference(name):
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    prog = planmod.compile_simplex(desc, order)
    pts = numpy.asarray(case["points"], dtype=float)
    tr = fiat_oracle.resolve_entity(desc, case["entity"])
    if tr is not None:
        pts = pts.reshape(len(pts), tr[0].shape[0]) @ tr[0] + tr[1]
    near = fiat_oracle.locate_cells(desc, pts, unique=bool(prog.unique))
    out = emu.run_simplex(prog, pts, near)
    for j, alpha in enumerate(emu.keys(prog)):
        ref = case["ref"][alpha].reshape(prog.nrows, -1)
        scale = max(abs(ref).max(), 1e-300)
        assert abs(out[j] - ref).max() <= tolerance(desc, alpha) * scale, alpha


@pytest.mark.parametrize("name", _simplex_dubiner_cases())
def test_value_table_reproduces_reference(name):
    """Derivative-folded coefficient tables (value-table kernel): out_alpha = C_alpha . member values."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    prog = planmod.compile_simplex(desc, order)
    if prog.ncp == 0:
        pytest.skip("outside the value-table kernel's range")
    pts = numpy.asarray(case["points"], dtype=float)
    tr = fiat_oracle.resolve_entity(desc, case["entity"])
    if tr is not None:
        pts = pts.reshape(len(pts), tr[0].shape[0]) @ tr[0] + tr[1]
    near = fiat_oracle.locate_cells(desc, pts, unique=bool(prog.unique))
    out = emu.run_value_table(prog, pts, near)
    worst = 0.0
    for j, alpha in enumerate(emu.keys(prog)):
        ref = case["ref"][alpha].reshape(prog.nrows, -1)
        scale = max(abs(ref).max(), 1e-300)
        worst = max(worst, abs(out[j] - ref).max() / scale)
        assert abs(out[j] - ref).max() <= tolerance(desc, alpha) * scale, alpha
    print(name, "worst relative error", worst)


@pytest.mark.parametrize("name", ["n2curl4_tet_o1", "p8_tet_o2", "p5_tet_o3", "hermite3_tet_o2", "p6_tri_o4",
                                  "p4_tet_face2_o2", "p10_tri_o2", "argyris_tri_o2"])
def test_alpha_split_reproduces_reference(name):
    """Per-alpha derived order-0 elements (plan.alpha_split) and their stacked form (plan.merged_split):
    their order-0 tabulation is the reference's derivative table."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    split = planmod.alpha_split(desc, order)
    assert split is not None and [a for a, _ in split] == planmod.alpha_list(int(desc["sd"]), order)
    pts = numpy.asarray(case["points"], dtype=float)
    tr = fiat_oracle.resolve_entity(desc, case["entity"])
    if tr is not None:
        pts = pts.reshape(len(pts), tr[0].shape[0]) @ tr[0] + tr[1]
    near = numpy.ones((1, len(pts)), dtype=bool)
    for alpha, derived in split:
        ref = case["ref"][alpha].reshape(-1, len(pts))
        if derived is None:
            assert not ref.any()
            continue
        prog = planmod.compile_simplex(derived, 0)
        out = emu.run_simplex(prog, pts, near)[0]
        assert abs(out - ref).max() <= tolerance(desc, alpha) * max(abs(ref).max(), 1e-300), alpha
    merged = planmod.merged_split(desc, order, split)
    assert merged is not None
    prog = planmod.compile_simplex(merged, 0)
    out = emu.run_simplex(prog, pts, near)[0]
    nrows = out.shape[0] // len(split)
    for j, (alpha, _) in enumerate(split):
        ref = case["ref"][alpha].reshape(-1, len(pts))
        assert abs(out[j * nrows:(j + 1) * nrows] - ref).max() <= tolerance(desc, alpha) * max(abs(ref).max(), 1e-300), alpha
    assert prog.kpad % 4 == 0 and emu.blocks_to_dense(prog).shape == (prog.nrows, prog.nslots)


@pytest.mark.parametrize("name", ["gn_tet_o2", "walkington_tet_o2", "hct4_tri_o2", "hct_o2", "ps12_o2",
                                  "alfeld_sorokina_tet_adv_o2", "ch_wf_tet_adv_o1", "p2_wf_tet_adv_o2", "p1_ps_tet_adv_o1",
                                  "hct6_tri_o2", "jm_tri_o2"])
def test_macro_merged_reproduces_reference(name):
    """Derived order-0 element of a split-cell element (plan.macro_merged): per-subcell stacked matrices on the
    subcell's un-normalised members; its order-0 tabulation (with the reference's binning and multiplicities) is the
    stack of the reference's derivative tables.  Also checks the per-subcell block packing the tile kernel streams."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    merged = planmod.macro_merged(desc, order)
    assert merged is not None and merged["unique"] == int(bool(desc["c0"]) and order == 0)
    prog = planmod.compile_simplex(merged, 0)
    assert prog.blk_cells == prog.ncells == int(desc["ncells"]) and prog.unique == merged["unique"]
    pts = numpy.asarray(case["points"], dtype=float)
    near = fiat_oracle.locate_cells(desc, pts, unique=bool(prog.unique))
    out = emu.run_simplex(prog, pts, near)[0]
    alphas = planmod.alpha_list(int(desc["sd"]), order)
    nrows = out.shape[0] // len(alphas)
    for j, alpha in enumerate(alphas):
        ref = case["ref"][alpha].reshape(-1, len(pts))
        assert abs(out[j * nrows:(j + 1) * nrows] - ref).max() <= tolerance(desc, alpha) * max(abs(ref).max(), 1e-300), alpha
    # block stream of every subcell: one (nrb + 1) pointer row per subcell, >= 1 block per row block
    nrb = len(prog.blk_ptr) // prog.ncells - 1
    assert len(prog.blk_ptr) == prog.ncells * (nrb + 1) and nrb == -(-prog.nrows // 8)
    assert len(prog.blk_kb) == 4 * prog.blk_ptr[-1] and len(prog.blk_frag) == 32 * prog.blk_ptr[-1]
    assert prog.blk_kb.min() >= 0 and prog.blk_kb.max() < prog.kpad
    for c in range(prog.ncells):
        ptr = prog.blk_ptr[c * (nrb + 1):(c + 1) * (nrb + 1)]
        assert (numpy.diff(ptr) >= 1).all()
    # the fixed-k stream of the register-operand kernel holds the same matrices, step by step
    assert prog.crb == planmod.CELLS_STEP_RB and len(prog.cstream) == prog.cstep_ptr[-1] and (numpy.diff(prog.cstep_ptr) % 2 == 0).all()
    assert (len(prog.cstep_ptr) - 1) * prog.crb >= nrb
    for c in range(prog.ncells):
        assert numpy.array_equal(emu.stream_to_dense(prog, c), emu.blocks_to_dense(prog, c))
    # the packed matrices (insignificant entries dropped, plan.significant_entries) still reproduce the reference
    packed = emu.run_simplex(prog, pts, near, packed=True)[0]
    for j, alpha in enumerate(alphas):
        ref = case["ref"][alpha].reshape(-1, len(pts))
        assert abs(packed[j * nrows:(j + 1) * nrows] - ref).max() <= 2e-13 * max(abs(ref).max(), 1e-300), alpha


def test_mis_order_matches_reference_keys():
    for name in golden_case_names():
        case = load_case(name)
        if case["desc"]["kind"] != "simplex":
            continue
        sd = int(case["desc"]["sd"])
        assert [tuple(k) for k in case["keys"]] == planmod.alpha_list(sd, case["order"])


@pytest.mark.parametrize("name", ["p8_tet_o2", "n2curl4_tet_o1", "p3_tri_o1", "p12_tri_o2", "p10_spectral_tet_o2"])
def test_block_packing_reproduces_reference(name):
    """The 8x4 gather packing (fix-ups folded in, insignificant entries dropped by plan.significant_entries) still
    reproduces the reference: jets path and, for mid-size elements, the stacked derived element -- the P12 triangle
    is the case where dropping by coefficient size alone lost 1e-12 of the first-derivative tables."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    pts = numpy.asarray(case["points"], dtype=float)
    near = numpy.ones((1, len(pts)), dtype=bool)
    prog = planmod.compile_simplex(desc, order)
    idx = prog.blk_kb.reshape(-1, 4)
    assert prog.kpad % 4 == 0 and len(prog.rb_order) == len(prog.blk_ptr) - 1
    assert len(idx) == prog.blk_ptr[-1] and idx.min() >= 0 and idx.max() < prog.kpad
    # the gather of a block is bank-conflict free when its four slots differ mod 4: true for nearly all blocks
    if len(idx) >= 64:
        assert (numpy.sort(idx % 4, axis=1) == numpy.arange(4)).all(axis=1).mean() >= 0.9
    out = emu.run_simplex(prog, pts, near, packed=True)
    for j, alpha in enumerate(emu.keys(prog)):
        ref = case["ref"][alpha].reshape(prog.nrows, -1)
        assert abs(out[j] - ref).max() <= 2e-13 * max(abs(ref).max(), 1e-300), alpha
    split = planmod.alpha_split(desc, order, prog)
    merged = None if split is None else planmod.merged_split(desc, order, split)
    if merged is not None:
        pm = planmod.compile_simplex(merged, 0)
        got = emu.run_simplex(pm, pts, near, packed=True)[0]
        nrows = got.shape[0] // len(split)
        for j, (alpha, _) in enumerate(split):
            ref = case["ref"][alpha].reshape(-1, len(pts))
            assert abs(got[j * nrows:(j + 1) * nrows] - ref).max() <= 2e-13 * max(abs(ref).max(), 1e-300), alpha


def test_tensor_flattening_hex():
    from conftest import load_desc
    leaves = planmod.flatten_tensor(load_desc("gll_q10_hex"))
    assert [(lf.point_offset, lf.point_dim, lf.sd) for lf in leaves] == [(0, 1, 1), (1, 1, 1), (2, 1, 1)]
    assert all(lf.entity == (1, 0) for lf in leaves)


def _simplex_leaves(desc):
    kind = desc["kind"]
    if kind == "simplex":
        yield desc
    elif kind == "flattened":
        yield from _simplex_leaves(desc["element"])
    elif kind == "tensor":
        yield from _simplex_leaves(desc["A"])
        yield from _simplex_leaves(desc["B"])
    elif kind == "composite":
        for part in desc["parts"]:
            yield from _simplex_leaves(part["element"])


@pytest.mark.parametrize("name", golden_case_names())
def test_every_golden_description_compiles(name):
    """Plan compilation (incl. degenerate degree-0 line sets, wrapper elements) needs no GPU."""
    case = load_case(name)
    if not is_polynomial_case(case):
        pytest.skip("not a polynomial tabulation")
    for leaf in _simplex_leaves(case["desc"]):
        for order in range(0, case["order"] + 1):
            prog = planmod.compile_simplex(leaf, order)
            assert prog.nrows == prog.ndofs * max(1, int(numpy.prod(prog.value_shape)) if prog.value_shape else 1)
            assert numpy.isfinite(prog.ccell).all() and numpy.isfinite(prog.line_tab).all()
    parts = planmod.resolve_parts(case["desc"], case["entity"])
    ndofs = planmod.num_dofs_of(case["desc"])
    ref0 = next(iter(case["ref"].values()))
    assert ref0.shape[0] == ndofs and ref0.shape[1:-1] == planmod.value_shape_of(case["desc"])
    assert all(0 <= p.dof_base < ndofs for p in parts)


@pytest.mark.parametrize("name", ["p8_tet_o2", "n2curl4_tet_o1", "hct_o2", "ps12_o0", "gn_tet_o2", "p2_tri_facet1_o1",
                                  "p2_wf_tet_adv_o2", "regge2_tet_o1"])
def test_stacked_derived_reproduces_reference(name):
    """The operand of the fused point evaluation (plan.stacked_derived): an order-0 element whose rows are all the
    derivative tables of the element, with dense per-subcell matrices; its order-0 tabulation with the reference's
    binning is the stack of the reference's tables, and u . (that) are the derivatives of u = sum_i u_i phi_i."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    stacked = planmod.stacked_derived(desc, order)
    assert stacked is not None and stacked["dense_only"]
    prog = planmod.compile_simplex(stacked, 0)
    assert len(prog.blk_kb) == 0 and prog.ncp == 0          # no packing, no value table: only the dense matrices
    pts = numpy.asarray(case["points"], dtype=float)
    tr = fiat_oracle.resolve_entity(desc, case["entity"])
    if tr is not None:
        pts = pts.reshape(len(pts), tr[0].shape[0]) @ tr[0] + tr[1]
    near = fiat_oracle.locate_cells(desc, pts, unique=bool(prog.unique))
    out = emu.run_simplex(prog, pts, near)[0]
    alphas = planmod.alpha_list(int(desc["sd"]), order)
    nrows = out.shape[0] // len(alphas)
    for j, alpha in enumerate(alphas):
        ref = case["ref"][alpha].reshape(-1, len(pts))
        assert abs(out[j * nrows:(j + 1) * nrows] - ref).max() <= tolerance(desc, alpha) * max(abs(ref).max(), 1e-300), alpha


@pytest.mark.parametrize("name", ["p8_tet_o2", "rtcf1_quad_o1", "nested_tpe_o1", "gll_q3_hex_face4_o2", "gn_tet_o2"])
def test_quick_description_builds_only_dense_tables(name):
    """api._quick_description (small calls): every simplex description in the tree is marked dense_only, and
    compile_simplex then builds no block packing, value table or fixed-k stream -- and the same dense matrices."""
    from fiat_b200 import api
    case = load_case(name)
    quick = api._quick_description(case["desc"])

    def leaves(d):
        kind = d["kind"]
        if kind == "simplex":
            yield d
        elif kind == "tensor":
            yield from leaves(d["A"])
            yield from leaves(d["B"])
        elif kind == "flattened":
            yield from leaves(d["element"])
        elif kind == "composite":
            for part in d["parts"]:
                yield from leaves(part["element"])

    full_leaves, quick_leaves = list(leaves(case["desc"])), list(leaves(quick))
    assert len(full_leaves) == len(quick_leaves) >= 1
    assert not any(d.get("dense_only") for d in full_leaves)          # the caller's description is left alone
    for d_full, d_quick in zip(full_leaves, quick_leaves):
        assert d_quick["dense_only"] is True
        prog = planmod.compile_simplex(d_quick, case["order"])
        assert len(prog.blk_kb) == 0 and len(prog.cstream) == 0 and prog.ncp == 0
        ref = planmod.compile_simplex(d_full, case["order"])
        if ref.nslots == prog.nslots and len(ref.blk_kb) == 0:
            assert numpy.array_equal(ref.ccell_morton, prog.ccell_morton)
        assert prog.ccell_morton.shape == ref.ccell_morton.shape


@pytest.mark.parametrize("name", ["gn_tet_adv_o2", "alfeld_sorokina_tet_adv_o2", "p2_wf_tet_adv_o2", "hct5_tri_o2"])
def test_cells_reg_tile_algorithm_reproduces_reference(name):
    """Executable restatement of the register-operand split-cell kernel (program_emulator.run_cells_reg: column sort
    with octet padding, B fragments per octet, prefix block stream, scatter through the column permutation, points in
    several subcells) on adversarial point sets, against the reference's derivative tables."""
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    merged = planmod.macro_merged(desc, order)
    prog = planmod.compile_simplex(merged, 0)
    assert prog.crb > 0
    pts = numpy.asarray(case["points"], dtype=float)[:150]
    near = fiat_oracle.locate_cells(desc, pts, unique=bool(prog.unique))
    assert (near.sum(axis=0) > 1).any() or name == "hct5_tri_o2"          # interior facets / vertices are in the set
    nwarps = 4 if prog.ncells <= 4 else 8                                # tile = 16 * nwarps - 8 * ncells points:
    out = emu.run_cells_reg(prog, pts, near, nwarps=nwarps)              # small tiles, several of them, a ragged last one
    assert not numpy.isnan(out).any()
    alphas = planmod.alpha_list(int(desc["sd"]), order)
    nrows = out.shape[0] // len(alphas)
    for j, alpha in enumerate(alphas):
        ref = case["ref"][alpha].reshape(-1, case["ref"][alpha].shape[-1])[:, :150]
        assert abs(out[j * nrows:(j + 1) * nrows] - ref).max() <= 2e-13 * max(abs(ref).max(), 1e-300), alpha
