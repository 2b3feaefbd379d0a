"""Test helper: interpret a compiled SimplexProgram with numpy, step by step, the way the CUDA
kernel does (un-normalised recurrence, fix-ups, folded per-cell coefficient matrices).  It lets the
CPU test-suite check the host-side plan compiler without a GPU.  Not a product path."""
import numpy

from fiat_b200.plan import alpha_list


def _jets(prog, cell, x, fixups=True):
    """Run the Dubiner program of one cell at default-simplex coordinates x (sd, npts)."""
    sd, na = prog.sd, prog.na
    npts = x.shape[1]
    T = numpy.zeros((prog.nslots, na, npts))
    T[prog.start_slot, 0] = prog.geom[cell, 12]
    X = [x[i] for i in range(sd)] + [-numpy.ones(npts), -numpy.ones(npts)]
    pairs = [(d1, d2) for d1 in range(sd) for d2 in range(d1, sd)]
    g = prog.geom[cell]
    for s, (nxt, cur, prv, codim) in enumerate(prog.step_idx):
        a, b, c = prog.step_abc[s]
        dfa, dfb = g[14 + 3 * codim:14 + 3 * codim + sd], g[23 + 3 * codim:23 + 3 * codim + sd]
        fb = 0.5 * (X[codim + 1] + X[codim + 2])
        fa = X[codim] + (fb + 1.0)
        F = a * fa - b * fb
        dF = a * dfa - b * dfb
        G = -c * (fb * fb)
        dG = [fb * (-2 * c * dfb[d]) for d in range(sd)]
        ddG = [(-2 * c * dfb[d1]) * dfb[d2] for (d1, d2) in pairs]
        for j in range(na):
            v = F * T[cur, j]
            for d in range(sd):
                if prog.low1[j, d] >= 0:
                    v = v + prog.mul1[j, d] * dF[d] * T[cur, prog.low1[j, d]]
            if prv >= 0:
                v = v + G * T[prv, j]
                for d in range(sd):
                    if prog.low1[j, d] >= 0:
                        v = v + prog.mul1[j, d] * dG[d] * T[prv, prog.low1[j, d]]
                for k in range(len(pairs)):
                    if prog.low2[j, k] >= 0:
                        v = v + prog.mul2[j, k] * ddG[k] * T[prv, prog.low2[j, k]]
            T[nxt, j] = v
    if fixups:
        for (t, s), w in zip(prog.fix_idx, prog.fix_w):
            T[t] -= w * T[s]
    return T


def run_simplex(prog, pts, near, packed=False):
    """out[alpha_index, row, point] for already-transformed points and a membership matrix.
    packed=True contracts with the matrices rebuilt from the 8x4 block packing (what the tile kernels multiply:
    fix-ups folded in, insignificant entries dropped) instead of the dense per-cell matrices."""
    assert prog.expansion == 0
    sd = prog.sd
    npts = len(pts)
    out = numpy.zeros((prog.na, prog.nrows, npts))
    mult = near.sum(axis=0)
    for c in range(prog.ncells):
        ip = numpy.where(near[c])[0]
        if len(ip) == 0:
            continue
        A = prog.geom[c, :sd * sd].reshape(sd, sd)
        b = prog.geom[c, 9:9 + sd]
        x = (pts[ip] @ A.T + b).T
        T = _jets(prog, c, x, fixups=not packed)
        vals = numpy.einsum("rk,kap->arp", blocks_to_dense(prog, c) if packed else prog.ccell[c], T)
        out[:, :, ip] += vals / (1.0 if prog.unique else mult[None, None, ip])
    return out


def run_value_table(prog, pts, near):
    """The value-table kernel: member values only (no jets, no fix-ups), derivative-folded coefficients."""
    import math
    sd, n = prog.sd, prog.degree
    npts = len(pts)
    out = numpy.zeros((prog.na, prog.nrows, npts))
    mult = near.sum(axis=0)
    for c in range(prog.ncells):
        ip = numpy.where(near[c])[0]
        if len(ip) == 0:
            continue
        A = prog.geom[c, :sd * sd].reshape(sd, sd)
        x = (pts[ip] @ A.T + prog.geom[c, 9:9 + sd]).T
        # raw recurrence values in slot order, then Morton order
        T = numpy.zeros((prog.nslots, len(ip)))
        T[prog.start_slot] = prog.geom[c, 12]
        X = [x[i] for i in range(sd)] + [-numpy.ones(len(ip))] * 2
        for s, (nxt, cur, prv, codim) in enumerate(prog.step_idx):
            a, b, cc = prog.step_abc[s]
            fb = 0.5 * (X[codim + 1] + X[codim + 2])
            fa = X[codim] + (fb + 1.0)
            T[nxt] = (a * fa - b * fb) * T[cur] - (cc * (fb * fb) * T[prv] if prv >= 0 else 0.0)
        T = T[prog.slot_of]
        off = 0
        for j, alpha in enumerate(alpha_list(sd, prog.order)):
            k = sum(alpha)
            nm = math.comb(n - k + sd, sd) if k <= n else 0
            blk = prog.cderiv[off:off + prog.nrows * nm * prog.ncp].reshape(prog.nrows, nm, prog.ncp)[:, :, c]
            off += prog.nrows * nm * prog.ncp
            out[j][:, ip] += (blk @ T[:nm]) / (1.0 if prog.unique else mult[None, ip])
    return out


def blocks_to_dense(prog, cell=0):
    """Rebuild the dense folded coefficient matrix (table rows x member slots) of one subcell from the 8x4 gather
    packing: block q holds the coefficients of 8 packed rows on the four member slots blk_kb[4 q .. 4 q + 3]."""
    ncells = max(prog.blk_cells, 1)
    nrb = len(prog.blk_ptr) // ncells - 1
    ptr = prog.blk_ptr[cell * (nrb + 1):(cell + 1) * (nrb + 1)]
    idx = numpy.asarray(prog.blk_kb).reshape(-1, 4)
    C = numpy.zeros((nrb * 8, prog.kpad))
    for rb in range(nrb):
        for q in range(ptr[rb], ptr[rb + 1]):
            frag = prog.blk_frag[q * 32:(q + 1) * 32].reshape(8, 4)
            for t in range(4):
                C[rb * 8:rb * 8 + 8, idx[q, t]] += frag[:, t]
    out = numpy.zeros((prog.nrows, prog.nslots))
    out[prog.row_perm] = C[:prog.nrows, :prog.nslots]          # packed row i is table row row_perm[i]
    return out


def stream_to_dense(prog, cell=0):
    """The same matrix rebuilt from the fixed-k block stream of the register-operand split-cell kernel
    (plan.pack_fixed_stream): steps of crb row blocks, int32 records n | first block << 16 (the k-blocks 0 .. n - 1 are stored), k-block j = slots 4j..4j+3."""
    ncells, RB = prog.ncells, prog.crb
    nstep = len(prog.cstep_ptr) - 1
    hdr = -(-(ncells * RB) // 4) * 2
    KB = prog.kpad // 4
    C = numpy.zeros((nstep * RB * 8, prog.kpad))
    for s in range(nstep):
        step = numpy.asarray(prog.cstream[prog.cstep_ptr[s]:prog.cstep_ptr[s + 1]])
        meta = step[:hdr].view(numpy.int32)
        frags = step[hdr:].reshape(-1, 8, 4)
        for r in range(RB):
            m = int(meta[cell * RB + r])
            q, n = m >> 16, m & 0xFFFF
            assert n <= KB
            for kb in range(n):
                C[(s * RB + r) * 8:(s * RB + r) * 8 + 8, 4 * kb:4 * kb + 4] = frags[q + kb]
    out = numpy.zeros((prog.nrows, prog.nslots))
    out[prog.row_perm] = C[:prog.nrows, :prog.nslots]
    return out


def run_cells_reg(prog, pts, near, nwarps=16):
    """The tile algorithm of csrc/cells_reg.cuh, statement by statement, on the host: a tile of 16 * nwarps -
    8 * ncells points; columns sorted by subcell, each subcell's range padded to octets; warp w owns octets 2w, 2w + 1
    and reads their B fragments (k-block j = slots 4j..4j+3) once; steps of crb row blocks from the prefix stream
    (int32 records n | first block << 16, blocks 0 .. n - 1); results through the column permutation into rows
    row_perm[...]; points in several subcells (or none) are finished thread-per-point from the dense matrices.
    -> out[row, point] of the order-0 derived element (prog.na == 1)."""
    assert prog.na == 1 and prog.crb > 0
    ncells, RB, K = prog.ncells, prog.crb, prog.nslots
    KB = prog.kpad // 4
    nstep = len(prog.cstep_ptr) - 1
    hdr = -(-(ncells * RB) // 4) * 2
    PTS = 16 * nwarps
    PT = PTS - 8 * ncells
    npts = len(pts)
    out = numpy.full((prog.nrows, npts), numpy.nan)
    mult = near.sum(axis=0)
    sd = prog.sd
    for base in range(0, npts, PT):
        tile = numpy.arange(base, min(base + PT, npts))
        cell_of = numpy.where(mult[tile] == 1, near[:, tile].argmax(axis=0), -1)
        # phase 0: columns sorted by subcell, ranges padded to octets
        perm, octcell = -numpy.ones(PTS, dtype=int), []
        off = 0
        for c in range(ncells):
            mine = numpy.flatnonzero(cell_of == c)
            perm[off:off + len(mine)] = mine
            noct = -(-len(mine) // 8)
            octcell += [c] * noct
            off += 8 * noct
        # phase 1: member values of every column in its own subcell's coordinates (padding columns stay zero)
        T = numpy.zeros((4 * KB, PTS))
        for c in range(ncells):
            cols = [j for j in range(off) if perm[j] >= 0 and cell_of[perm[j]] == c]
            if cols:
                A = prog.geom[c, :sd * sd].reshape(sd, sd)
                x = (pts[tile[perm[cols]]] @ A.T + prog.geom[c, 9:9 + sd]).T
                T[:K, cols] = _jets(prog, c, x, fixups=False)[:, 0, :]
        # phase 2: warps x steps
        for w in range(nwarps):
            for o in (2 * w, 2 * w + 1):
                if o >= len(octcell):
                    continue
                c = octcell[o]
                B = T[:, 8 * o:8 * o + 8].reshape(KB, 4, 8)              # B fragments, one per k-block
                for s in range(nstep):
                    step = numpy.asarray(prog.cstream[prog.cstep_ptr[s]:prog.cstep_ptr[s + 1]])
                    meta, frags = step[:hdr].view(numpy.int32), step[hdr:].reshape(-1, 8, 4)
                    for r in range(RB):
                        m = int(meta[c * RB + r])
                        n, first = m & 0xFFFF, m >> 16
                        acc = numpy.zeros((8, 8))
                        for kb in range(n - 1, -1, -1):                      # the fall-through run: n - 1, ..., 0
                            acc += frags[first + kb] @ B[kb]
                        rb = s * RB + r
                        for g in range(8):
                            row = prog.row_perm[rb * 8 + g] if rb * 8 + g < prog.nrows else -1
                            for col in range(8):
                                p = perm[8 * o + col]
                                if row >= 0 and p >= 0:
                                    out[row, tile[p]] = acc[g, col]
        # phase 3: points in several subcells (tables averaged) or in none (zero column)
        for i, p in enumerate(tile):
            if mult[p] == 1:
                continue
            out[:, p] = 0.0
            for c in numpy.flatnonzero(near[:, p]):
                A = prog.geom[c, :sd * sd].reshape(sd, sd)
                x = (pts[p:p + 1] @ A.T + prog.geom[c, 9:9 + sd]).T
                out[:, p] += prog.ccell[c] @ _jets(prog, c, x)[:, 0, 0] / (1.0 if prog.unique else mult[p])
    return out


def keys(prog):
    return alpha_list(prog.sd, prog.order)
