"""CPU ORACLE (test infrastructure, NOT product code).

Plain numpy/scipy restatement of the beam-FEM hot path of pyLatticeDSO
(pyLatticeSim).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product (``pylatticedso_b200``) never does: it fails loudly when
its CUDA library is missing.

Pinning status
--------------
* Element formula, frame rule, shear reduced integration, kappa, the gmsh 1-D
  subdivision rule, penalisation radius, boundary-DOF ordering and the
  condensation are PINNED by the reference's 30 stored dolfinx/PETSc Schur
  matrices (``data/outputs/schur_complement/Schur_complement_{BCC,Hybrid1,
  Hybrid4}.npz``; copies under ``tests/golden/``): see
  ``tests/test_oracle_golden.py`` (all 30 agree to <= 2e-12 relative).
* CSR structure / DOF numbering, displacements, reactions and radius
  gradients are NOT pinned by any test or fixture of the reference
  ("parity unpinned" for those rows).  They are
  anchored on reference-in-the-loop runs (the reference's own ``solve_DDM``
  and ``LatticeOpti.gradient`` fed with this oracle's Schur matrices), frozen
  as fixtures by ``tests/golden/make_golden.py``.

Each function cites the reference file:line (relative to the pyLatticeDSO
checkout) whose behaviour it restates.  The arithmetic of the reference lives
in dolfinx==0.9.0 / ufl==2024.2.0 / basix==0.9.0 / PETSc / gmsh>=4.14
(``pyproject.toml:18-29``), none of which is vendored; what is restated here
is the published P1xP1 Timoshenko formulation those libraries evaluate for the
UFL form in ``simulation_base.py``.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

KAPPA = 0.9          # material_definition.py:45
MESH_FRACTION = 0.05  # lattice_generation.py:50-64  (h = 0.05 * cell_size_x)
NDOF = 6             # point.py:68  [ux,uy,uz,rx,ry,rz]


# --------------------------------------------------------------------------
# A1: element stiffness
# --------------------------------------------------------------------------
def local_frame(t):
    """Local frame (a1, a2) of unit tangents ``t`` [E,3].

    beam_model.py:197-216: e1 = ey if |t_y| < |t_x| else ex;
    e2 = ez if |t_z| < |t.e1| else e1; a1 = t x e2 / |.|; a2 = t x a1 / |.|.
    Comparisons are strict on absolute values.
    """
    t = np.asarray(t, dtype=np.float64)
    ex = np.array([1.0, 0.0, 0.0])
    ey = np.array([0.0, 1.0, 0.0])
    ez = np.array([0.0, 0.0, 1.0])
    c1 = np.abs(t[:, 1]) < np.abs(t[:, 0])
    e1 = np.where(c1[:, None], ey, ex)
    te1 = np.einsum("ij,ij->i", t, e1)
    c2 = np.abs(t[:, 2]) < np.abs(te1)
    e2 = np.where(c2[:, None], ez, e1)
    a1 = np.cross(t, e2)
    a1 /= np.linalg.norm(a1, axis=1)[:, None]
    a2 = np.cross(t, a1)
    a2 /= np.linalg.norm(a2, axis=1)[:, None]
    return a1, a2


def section_stiffness(r, E, nu, kappa=KAPPA):
    """(ES, GS, GJ, EI) from radius. material_definition.py:142-156, :131."""
    r = np.asarray(r, dtype=np.float64)
    G = E / (2.0 * (1.0 + nu))
    S = math.pi * r ** 2
    I = (math.pi * r ** 4) / 4.0
    J = 2.0 * I
    return E * S, G * kappa * S, G * J, E * I


def section_stiffness_drad(r, E, nu, kappa=KAPPA):
    """d(ES, GS, GJ, EI)/dr. material_definition.py:207-223 (normal beams)."""
    r = np.asarray(r, dtype=np.float64)
    G = E / (2.0 * (1.0 + nu))
    dS = 2.0 * math.pi * r
    dI = math.pi * r ** 3
    return E * dS, G * kappa * dS, G * 2.0 * dI, E * dI


def strain_vectors(x1, x2):
    """Six 12-vectors b_i [E,6,12] and lengths L [E].

    simulation_base.py:141-156 (generalised strains), :190-197 and :220-225
    (degree-1 quadrature for the two shear terms => theta at the midpoint).
    Element DOF layout [w1(3), th1(3), w2(3), th2(3)] in global axes.
    """
    x1 = np.asarray(x1, dtype=np.float64).reshape(-1, 3)
    x2 = np.asarray(x2, dtype=np.float64).reshape(-1, 3)
    d = x2 - x1
    L = np.linalg.norm(d, axis=1)
    t = d / L[:, None]
    a1, a2 = local_frame(t)
    n = x1.shape[0]
    B = np.zeros((n, 6, 12))
    iL = (1.0 / L)[:, None]
    # axial
    B[:, 0, 0:3] = -t * iL
    B[:, 0, 6:9] = t * iL
    # shear 1: dw.a1/L - thbar.a2
    B[:, 1, 0:3] = -a1 * iL
    B[:, 1, 6:9] = a1 * iL
    B[:, 1, 3:6] = -0.5 * a2
    B[:, 1, 9:12] = -0.5 * a2
    # shear 2: dw.a2/L + thbar.a1
    B[:, 2, 0:3] = -a2 * iL
    B[:, 2, 6:9] = a2 * iL
    B[:, 2, 3:6] = 0.5 * a1
    B[:, 2, 9:12] = 0.5 * a1
    # torsion
    B[:, 3, 3:6] = -t * iL
    B[:, 3, 9:12] = t * iL
    # bending
    B[:, 4, 3:6] = -a1 * iL
    B[:, 4, 9:12] = a1 * iL
    B[:, 5, 3:6] = -a2 * iL
    B[:, 5, 9:12] = a2 * iL
    return B, L


def element_stiffness(x1, x2, r, E, nu, kappa=KAPPA, drad=False):
    """12x12 stiffness (or d/dr of it) of 2-node P1/P1 Timoshenko elements.

    K_e = L * sum_i D_i b_i b_i^T, D = (ES, GS, GS, GJ, EI, EI)
    (material_definition.py:111-113 pairs the stresses with the strains of
    simulation_base.py:151-156).
    """
    B, L = strain_vectors(x1, x2)
    r = np.broadcast_to(np.asarray(r, dtype=np.float64), L.shape)
    ES, GS, GJ, EI = (section_stiffness_drad if drad else section_stiffness)(r, E, nu, kappa)
    D = np.stack([ES, GS, GS, GJ, EI, EI], axis=1) * L[:, None]
    return np.einsum("ei,eia,eib->eab", D, B, B)


# --------------------------------------------------------------------------
# A2: element set (gmsh 1-D subdivision)
# --------------------------------------------------------------------------
def gmsh_segments(L, h):
    """Number of 2-node elements gmsh puts on a straight line of length L with
    target size h at both end points: max(1, int(L/h + 0.99)).
    lattice_generation.py:50-64,119,161 (rule probed against the 30 goldens)."""
    return max(1, int(L / h + 0.99))


def flatten_lattice(lattice, cell_index=None, elements_per_strut="gmsh"):
    """Loop-based flattening of a reference ``Lattice``/``LatticeSim`` object
    graph into arrays, in the canonical order used everywhere in this repo:
    lattice nodes by ``node.index``, beams by ``beam.index``, strut-interior
    nodes appended beam-major.

    lattice_generation.py:105-175 (which nodes/beams are meshed; ``radius<=0``
    beams skipped at :158), beam_model.py:145-166 (element radius = radius of
    its beam; already x1.5 on ``beam_mod`` beams, beam.py:405-411).

    Returns dict(xyz[N,3], en[E,2], rad[E], beam_of_elem[E], mod[E] (bool),
    point_index[Np] (node.index of the first Np nodes), points (list)).
    """
    if cell_index is None:
        cells = list(lattice.cells)
    else:
        cells = [c for c in lattice.cells if c.index == cell_index]
    pts = {}
    beams = {}
    for c in cells:
        for p in c.points_cell:
            pts[p.index] = p
        for b in c.beams_cell:
            if b.radius > 0:
                beams[b.index] = b
    order = sorted(pts)
    points = [pts[i] for i in order]
    loc = {idx: k for k, idx in enumerate(order)}
    xyz = [[p.x, p.y, p.z] for p in points]
    h = MESH_FRACTION * lattice.cell_size_x
    en, rad, bo, mod = [], [], [], []
    for bi in sorted(beams):
        b = beams[bi]
        p1, p2 = b.point1, b.point2
        a = np.array([p1.x, p1.y, p1.z])
        c_ = np.array([p2.x, p2.y, p2.z])
        L = float(np.linalg.norm(c_ - a))
        nseg = gmsh_segments(L, h) if elements_per_strut == "gmsh" else int(elements_per_strut)
        prev = loc[p1.index]
        for k in range(1, nseg):
            xyz.append(list(a + (c_ - a) * (k / nseg)))
            cur = len(xyz) - 1
            en.append((prev, cur))
            prev = cur
        en.append((prev, loc[p2.index]))
        rad.extend([b.radius] * nseg)
        bo.extend([bi] * nseg)
        mod.extend([bool(b.beam_mod)] * nseg)
    return dict(xyz=np.array(xyz, dtype=np.float64), en=np.array(en, dtype=np.int32).reshape(-1, 2),
                rad=np.array(rad, dtype=np.float64), beam_of_elem=np.array(bo, dtype=np.int64),
                mod=np.array(mod, dtype=bool), point_index=np.array(order, dtype=np.int64),
                points=points)


# --------------------------------------------------------------------------
# A3: assembly
# --------------------------------------------------------------------------
def assemble_csr(xyz, en, rad, E, nu, kappa=KAPPA, drad=False):
    """Global K as scipy CSR (sorted indices, duplicates summed; explicit zeros
    kept, i.e. the pattern is the full 6x6 block graph like dolfinx's
    ``create_matrix`` sparsity).  simulation_base.py:480-481,
    schur_complement.py:69-71."""
    xyz = np.asarray(xyz, dtype=np.float64)
    en = np.asarray(en)
    Ke = element_stiffness(xyz[en[:, 0]], xyz[en[:, 1]], rad, E, nu, kappa, drad=drad)
    dofs = (en[:, :, None] * NDOF + np.arange(NDOF)[None, None, :]).reshape(-1, 12)
    rows = np.repeat(dofs, 12, axis=1).ravel()
    cols = np.tile(dofs, (1, 12)).ravel()
    n = NDOF * xyz.shape[0]
    K = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


# --------------------------------------------------------------------------
# A4/A5/A10: BC elimination, solve, reactions
# --------------------------------------------------------------------------
def apply_dirichlet(K, fixed, g, f):
    """Return (K_bc, b): rows & cols of constrained DOFs zeroed with unit
    diagonal, b = f - K[:,c] g on free rows, b[c] = g[c].
    simulation_base.py:480 (assemble_matrix(bcs)), :490 (apply_lifting),
    :492 (set_bc), :495-499 (point loads added AFTER set_bc, also on
    constrained rows - see ``reference_rhs``)."""
    fixed = np.asarray(fixed, dtype=bool)
    g = np.where(fixed, np.asarray(g, dtype=np.float64), 0.0)
    b = np.asarray(f, dtype=np.float64) - K @ g
    b[fixed] = g[fixed]
    keep = sp.diags((~fixed).astype(np.float64))
    Kbc = (keep @ K @ keep + sp.diags(fixed.astype(np.float64))).tocsr()
    return Kbc, b


def solve_static(K, fixed, g, f):
    """u with K u = f on free DOFs, u[c] = g (direct sparse LU standing in for
    PETSc preonly+lu, simulation_base.py:502-511). Returns (u, reactions) with
    reactions = K u (unconstrained K, simulation_base.py:582-645)."""
    fixed = np.asarray(fixed, dtype=bool)
    free = np.flatnonzero(~fixed)
    gz = np.where(fixed, np.asarray(g, dtype=np.float64), 0.0)
    rhs = (np.asarray(f, dtype=np.float64) - K @ gz)[free]
    Kff = K[free][:, free].tocsc()
    u = gz.copy()
    u[free] = spla.splu(Kff).solve(rhs)
    return u, K @ u


# --------------------------------------------------------------------------
# A6: the reference's PCG
# --------------------------------------------------------------------------
def reference_pcg(A, b, Minv=None, maxiter=100, tol=1e-5, mintol=1e-5,
                  restart_every=1000, alpha_max=0.1):
    """Restatement of conjugate_gradient_solver.py:59-122 (x0=0, alpha clamp
    :78-79, restart :89-90, dual stop test :97/:102, info=2 flag :107-109).
    ``A`` and ``Minv`` are callables or support ``@``. Returns (x, info, iters)."""
    mv = (lambda v: A(v)) if callable(A) else (lambda v: A @ v)
    if Minv is None:
        pc = lambda v: v
    elif callable(Minv):
        pc = Minv
    else:
        pc = lambda v: Minv @ v
    n = b.shape[0]
    x = np.zeros(n)
    r = b - mv(x)
    z = pc(r)
    p = z.copy()
    rz_old = float(np.dot(r, z))
    norm_b = float(np.linalg.norm(b))
    info = 1
    it = 0
    for k in range(maxiter):
        it = k + 1
        Ap = mv(p)
        alpha = rz_old / float(np.dot(p, Ap))
        alpha = min(alpha, alpha_max)
        x += alpha * p
        r -= alpha * Ap
        if k % restart_every == 0 and k > 0:
            p = z.copy()
        rn = float(np.linalg.norm(r))
        dn = float(np.linalg.norm(p))
        sn = float(np.linalg.norm(x))
        if rn <= tol * norm_b:
            info = 0
            break
        if dn < mintol * (sn + 1e-12):
            info = 0
            break
        if alpha < 1e-6:
            info = 2
        z = pc(r)
        rz_new = float(np.dot(r, z))
        beta = rz_new / rz_old
        p = z + beta * p
        rz_old = rz_new
    return x, info, it


# --------------------------------------------------------------------------
# A7: per-cell Schur complement
# --------------------------------------------------------------------------
def schur_complement(K, bnd_dofs):
    """S = K_BB - K_BI K_II^-1 K_IB with B in the given order, I = the rest in
    ascending order. schur_complement.py:75-147."""
    K = K.toarray() if sp.issparse(K) else np.asarray(K)
    n = K.shape[0]
    bnd_dofs = np.asarray(bnd_dofs, dtype=np.int64)
    mask = np.ones(n, dtype=bool)
    mask[bnd_dofs] = False
    I = np.flatnonzero(mask)
    KBB = K[np.ix_(bnd_dofs, bnd_dofs)]
    if I.size == 0:
        return KBB.copy()
    KBI = K[np.ix_(bnd_dofs, I)]
    KII = K[np.ix_(I, I)]
    return KBB - KBI @ np.linalg.solve(KII, KBI.T)


def cell_schur_from_lattice(lattice, cell_index, E, nu, elements_per_strut="gmsh"):
    """utils_schur.py:22-53: boundary DOF order = 6 DOFs of each node of
    ``cell.node_in_order_simulation`` (cell.py:611-680), the cell's own beams
    only (``cell.beams_cell``)."""
    cell = next(c for c in lattice.cells if c.index == cell_index)
    cell.define_node_order_to_simulate()
    flat = flatten_lattice(lattice, cell_index, elements_per_strut)
    loc = {int(idx): k for k, idx in enumerate(flat["point_index"])}
    bnd = []
    for p in cell.node_in_order_simulation:
        k = loc[p.index]
        bnd.extend(range(NDOF * k, NDOF * k + NDOF))
    K = assemble_csr(flat["xyz"], flat["en"], flat["rad"], E, nu)
    return schur_complement(K, bnd)


# --------------------------------------------------------------------------
# A11: compliance gradient
# --------------------------------------------------------------------------
def compliance_gradient(xyz, en, rad, u, group, n_groups, E, nu, kappa=KAPPA, chain=None):
    """g[p] = - sum_{e in group p} chain_e * u_e^T (dK_e/dr)(r_e) u_e.

    Element form of lattice_opti.py:701-841 / lattice_sim.py:1020-1054 with the
    analytic section derivatives of material_definition.py:163-231;
    ``chain`` = d r_e / d r_param (1.5 on penalised segments, beam.py:422-436).
    Elements with group < 0 are ignored.
    """
    xyz = np.asarray(xyz, dtype=np.float64)
    dK = element_stiffness(xyz[en[:, 0]], xyz[en[:, 1]], rad, E, nu, kappa, drad=True)
    dofs = (en[:, :, None] * NDOF + np.arange(NDOF)[None, None, :]).reshape(-1, 12)
    ue = np.asarray(u, dtype=np.float64)[dofs]
    q = np.einsum("ea,eab,eb->e", ue, dK, ue)
    if chain is not None:
        q = q * chain
    g = np.zeros(n_groups)
    ok = np.asarray(group) >= 0
    np.add.at(g, np.asarray(group)[ok], -q[ok])
    return g


def element_action_closed_form(x_own, x_other, r, u_own, u_other, E, nu, kappa=KAPPA):
    """Force of ONE element on its end node ``own`` as the closed form the CUDA matrix-free operator evaluates
    (pylatticedso_b200/csrc/matfree.cuh) -- checker for that algebra, vectorised over a batch:

        dw = w_i - w_j,  st = th_i + th_j,  dth = th_i - th_j,  t = (x_j - x_i)/L
        f_w  = aI dw + aT t (t.dw) - c (t x st)
        f_th = bI (st - t (t.st)) + dI dth + dT t (t.dth) + c (t x dw)

    with aI = GS/L, aT = (ES-GS)/L, c = GS/2, bI = GS L/4, dI = EI/L, dT = (GJ-EI)/L -- equal to the rows of
    ``element_stiffness`` (beam_model.py:197-216 through the explicit frame) for either end of the element.
    Returns [n, 6]."""
    x_own, x_other = np.atleast_2d(x_own).astype(np.float64), np.atleast_2d(x_other).astype(np.float64)
    u_own, u_other = np.atleast_2d(u_own).astype(np.float64), np.atleast_2d(u_other).astype(np.float64)
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    d = x_other - x_own
    L = np.linalg.norm(d, axis=1)
    t = d / L[:, None]
    G = E / (2.0 * (1.0 + nu))
    S = np.pi * r ** 2
    I = np.pi * r ** 4 / 4.0
    ES, GS, EI, GJ = E * S, G * kappa * S, E * I, G * 2.0 * I
    aI, aT, c, bI, dI, dT = GS / L, (ES - GS) / L, 0.5 * GS, 0.25 * GS * L, EI / L, (GJ - EI) / L
    dw = u_own[:, :3] - u_other[:, :3]
    st = u_own[:, 3:] + u_other[:, 3:]
    dth = u_own[:, 3:] - u_other[:, 3:]
    dot = lambda a, b: np.einsum("ij,ij->i", a, b)[:, None]
    col = lambda v: v[:, None]
    f_w = col(aI) * dw + col(aT) * t * dot(t, dw) - col(c) * np.cross(t, st)
    f_t = col(bI) * (st - t * dot(t, st)) + col(dI) * dth + col(dT) * t * dot(t, dth) + col(c) * np.cross(t, dw)
    return np.concatenate([f_w, f_t], axis=1)


# ---------------------------------------------------------------------------------------------------------
# Chain condensation (checker for lat_schur_batch's strut pre-pass, pylatticedso_b200/csrc/lattice_schur.cu)
# ---------------------------------------------------------------------------------------------------------
def find_chains(n_nodes, en, keep, xyz, tol=1e-9):
    """Split a cell mesh into straight strut chains.  A CHAIN NODE is a node that is not in ``keep`` (the
    boundary nodes), has exactly two incident elements and these are collinear.  Returns a list of
    (A, B, [(element, flipped), ...]) walking from joint A to joint B; elements that touch no chain node
    come back as chains of length one.  Exact restatement target: eliminating the chain nodes first is plain
    static condensation, so the Schur complement on ``keep`` is unchanged."""
    en = np.asarray(en)
    deg = np.bincount(en.ravel(), minlength=n_nodes)
    inc = [[] for _ in range(n_nodes)]
    for e, (a, b) in enumerate(en):
        inc[a].append(e)
        inc[b].append(e)
    is_keep = np.zeros(n_nodes, dtype=bool)
    is_keep[np.asarray(keep)] = True
    d = xyz[en[:, 1]] - xyz[en[:, 0]]
    d /= np.linalg.norm(d, axis=1)[:, None]
    chain_node = np.zeros(n_nodes, dtype=bool)
    for n in range(n_nodes):
        if not is_keep[n] and deg[n] == 2:
            e0, e1 = inc[n]
            chain_node[n] = np.linalg.norm(np.cross(d[e0], d[e1])) < tol
    used = np.zeros(len(en), dtype=bool)
    chains = []
    for e_start in range(len(en)):
        if used[e_start]:
            continue
        a, b = en[e_start]
        if chain_node[a] and chain_node[b]:
            continue                      # interior of a chain: reached from one of its ends
        # orient so that the walk starts at a joint
        start, nxt, flipped = (a, b, False) if not chain_node[a] else (b, a, True)
        seq = [(e_start, flipped)]
        used[e_start] = True
        while chain_node[nxt]:
            e_next = [e for e in inc[nxt] if not used[e]][0]
            used[e_next] = True
            a2, b2 = en[e_next]
            fl = a2 != nxt
            seq.append((e_next, fl))
            nxt = b2 if not fl else a2
        chains.append((int(start), int(nxt), seq))
    assert used.all(), "closed loop of chain nodes"
    return chains


def condensed_strut(xyz, en, rad, chain, E, nu, kappa=KAPPA):
    """12x12 stiffness of a straight chain of elements condensed onto its two end joints, computed the way
    the CUDA pre-pass does: in the frame of the strut the problem splits into an axial spring series
    (sum L/ES), a torsion spring series (sum L/GJ) and ONE planar Timoshenko beam (the two bending planes are
    identical for a circular section), whose 4x4 matrix on (W_A, Phi_A, W_B, Phi_B) is condensed with 2x2 pivots:
        M_WW(re,ce) = sA GS/L,  M_WPhi(re,ce) = (re==0 ? -GS/2 : GS/2),  M_PhiPhi(re,ce) = GS L/4 + sA EI/L.
    Back in 3-D:  ww = M_WW (I - tt) + k_ax tt,  w-theta = M_WPhi [t]x,  theta-w = -M_PhiW [t]x,
                  theta-theta = M_PhiPhi (I - tt) + k_tor tt."""
    A, B, seq = chain
    G = E / (2.0 * (1.0 + nu))
    t = None
    flex_ax = flex_tor = 0.0
    M = None                         # planar beam condensed so far on (W_A, Phi_A, W_k, Phi_k)
    for e, flipped in seq:
        a, b = (en[e][1], en[e][0]) if flipped else (en[e][0], en[e][1])
        dvec = xyz[b] - xyz[a]
        L = np.linalg.norm(dvec)
        if t is None:
            t = dvec / L
        r = rad[e]
        S = np.pi * r * r
        I = np.pi * r ** 4 / 4.0
        ES, GS, EI, GJ = E * S, G * kappa * S, E * I, G * 2.0 * I
        flex_ax += L / ES
        flex_tor += L / GJ
        Me = np.array([[GS / L, -GS / 2, -GS / L, -GS / 2],
                       [-GS / 2, GS * L / 4 + EI / L, GS / 2, GS * L / 4 - EI / L],
                       [-GS / L, GS / 2, GS / L, GS / 2],
                       [-GS / 2, GS * L / 4 - EI / L, GS / 2, GS * L / 4 + EI / L]])
        if M is None:
            M = Me
        else:
            # 6x6 on (A, k, next); eliminate the middle pair k
            T = np.zeros((6, 6))
            T[:4, :4] += M
            T[2:, 2:] += Me
            keep = [0, 1, 4, 5]
            P = T[2:4, 2:4]
            C = T[np.ix_(keep, [2, 3])]
            M = T[np.ix_(keep, keep)] - C @ np.linalg.solve(P, C.T)
    k_ax, k_tor = 1.0 / flex_ax, 1.0 / flex_tor
    tt = np.outer(t, t)
    Pp = np.eye(3) - tt
    Sk = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    K = np.zeros((12, 12))
    for re in range(2):
        for ce in range(2):
            sA = 1.0 if re == ce else -1.0
            K[6 * re:6 * re + 3, 6 * ce:6 * ce + 3] = M[2 * re, 2 * ce] * Pp + sA * k_ax * tt
            K[6 * re:6 * re + 3, 6 * ce + 3:6 * ce + 6] = M[2 * re, 2 * ce + 1] * Sk
            K[6 * re + 3:6 * re + 6, 6 * ce:6 * ce + 3] = -M[2 * re + 1, 2 * ce] * Sk
            K[6 * re + 3:6 * re + 6, 6 * ce + 3:6 * ce + 6] = M[2 * re + 1, 2 * ce + 1] * Pp + sA * k_tor * tt
    return K, (A, B)


def schur_via_chain_condensation(xyz, en, rad, bnd_nodes, E, nu, kappa=KAPPA):
    """Schur complement on the boundary DOFs computed from the JOINT-ONLY cell (every straight strut replaced by
    its condensed 12x12 super-element).  Equal to ``schur_complement(assemble_csr(...), bnd_dofs)``."""
    xyz = np.asarray(xyz, dtype=np.float64)
    en = np.asarray(en)
    n_nodes = xyz.shape[0]
    chains = find_chains(n_nodes, en, bnd_nodes, xyz)
    joints = sorted({c[0] for c in chains} | {c[1] for c in chains} | set(int(b) for b in bnd_nodes))
    jpos = {n: k for k, n in enumerate(joints)}
    Kj = np.zeros((6 * len(joints), 6 * len(joints)))
    for ch in chains:
        Ks, (A, B) = condensed_strut(xyz, en, rad, ch, E, nu, kappa)
        idx = np.r_[6 * jpos[A] + np.arange(6), 6 * jpos[B] + np.arange(6)]
        Kj[np.ix_(idx, idx)] += Ks
    bd = np.concatenate([6 * jpos[int(b)] + np.arange(6) for b in bnd_nodes])
    it = np.setdiff1d(np.arange(Kj.shape[0]), bd)
    if it.size == 0:
        return Kj[np.ix_(bd, bd)], chains
    return Kj[np.ix_(bd, bd)] - Kj[np.ix_(bd, it)] @ np.linalg.solve(Kj[np.ix_(it, it)], Kj[np.ix_(it, bd)]), chains


def assemble_joint_only(xyz, en, rad, n_points, E, nu, kappa=KAPPA):
    """Joint-only stiffness of a subdivided lattice mesh (mesh.py numbering: the first ``n_points`` nodes are the
    lattice points, elements beam-major from point1 to point2): every strut replaced by its condensed 12x12
    super-element.  With no load and no constraint on strut-interior nodes, solving this system gives exactly the
    joint displacements and joint reactions of the full system (static condensation).  Checker for
    lat_assemble_bsr_struts.  Returns (K_joints csr, chains)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    en = np.asarray(en)
    starts = np.flatnonzero(en[:, 0] < n_points)
    ends = np.r_[starts[1:], len(en)]
    rows, cols, vals = [], [], []
    chains = []
    for s0, s1 in zip(starts, ends):
        ch = (int(en[s0, 0]), int(en[s1 - 1, 1]), [(e, False) for e in range(s0, s1)])
        assert ch[1] < n_points
        chains.append(ch)
        Ks, (A, B) = condensed_strut(xyz, en, rad, ch, E, nu, kappa)
        idx = np.r_[6 * A + np.arange(6), 6 * B + np.arange(6)]
        rows.append(np.repeat(idx, 12)); cols.append(np.tile(idx, 12)); vals.append(Ks.ravel())
    K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(6 * n_points, 6 * n_points)).tocsr()
    return K, chains

