"""Device pipeline of the full-lattice beam FEM and the drop-in replacement of
``pyLatticeSim.utils_simulation.solve_FEM_FenicsX`` (utils_simulation.py:21-56).

    mesh (host SoA) --H2D--> element generation + BSR assembly --> Dirichlet
    elimination --> Jacobi / 6x6 block-Jacobi PCG --> reactions --D2H--> Points

All numerics run in ``liblattice_b200.so`` (CUDA, sm_100a); this module only
moves arrays and mirrors the reference's write-back onto ``Point`` objects
(full_scale_lattice_simulation.py:77-120).
"""
from __future__ import annotations

import numpy as np

from . import lib as L
from .mesh import NDOF, BeamMesh, bc_arrays_from_lattice, flatten_lattice

KAPPA = 0.9  # material_definition.py:45

# pyLatticeDesign/materials/*.json (E, nu) -- data consumed as-is
MATERIALS = {"VeroClear": (1013.0, 0.3)}


def material_constants(lattice):
    """(E, nu) of ``lattice.material_name`` (beam_model.py:190-193 -> materials.py:9-52)."""
    name = getattr(lattice, "material_name", "VeroClear")
    if name in MATERIALS:
        return MATERIALS[name]
    try:  # fall back to the reference's own loader when it is importable
        from pyLatticeDesign.materials import MatProperties
        m = MatProperties(name)
        return float(m.young_modulus), float(m.poisson_ratio)
    except Exception as e:  # pragma: no cover
        raise KeyError(f"unknown material {name!r}") from e


def stretch_dominated(mesh: BeamMesh) -> bool:
    """Maxwell's rigidity count for pin-jointed frames: a 3-D lattice whose joints connect >= 12 struts on average in the
    bulk (Octet: 12, BCC: 8) carries load by strut tension / compression; its stiffness matrix is then governed by
    long-wavelength modes -- the ones the rigid-body-mode coarse space removes (two-level table)."""
    n_struts = int(np.count_nonzero(mesh.en0 < mesh.n_points))      # every strut starts at a lattice point (mesh.py numbering)
    return 2.0 * n_struts / max(1, mesh.n_points) >= 10.5           # surface joints pull the mean below the bulk value


class BeamFEM:
    """Assembled beam-FEM operator resident on one GPU."""

    def __init__(self, mesh: BeamMesh, young: float, nu: float, kappa: float = KAPPA, ctx: L.Context | None = None,
                 pinned: bool = False):
        import torch
        self.torch = torch
        # an operator owns a resident sparsity pattern inside its context: without an explicit ctx it gets a private one
        self.ctx = ctx or L.Context()
        self.mesh = mesh
        self.young, self.nu, self.kappa = float(young), float(nu), float(kappa)
        dev = self.ctx.device
        self.h2d_bytes = 0

        def up(a, dtype):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
            if pinned:
                t = t.pin_memory()
            self.h2d_bytes += t.numel() * t.element_size()
            return t.to(dev, non_blocking=pinned)

        self.x = up(mesh.x, np.float64)
        self.y = up(mesh.y, np.float64)
        self.z = up(mesh.z, np.float64)
        self.en0 = up(mesh.en0, np.int32)
        self.en1 = up(mesh.en1, np.int32)
        self.rad = up(mesh.rad, np.float64)
        self.n_nodes = mesh.n_nodes
        self.n_elems = mesh.n_elems
        self.n_dof = mesh.n_dof
        self.rowptr = self.colidx = self.vals = self.vals_bc = None

    # -- pattern + assembly ---------------------------------------------------
    def build_pattern(self):
        self.rowptr, self.colidx = self.ctx.bsr_pattern(self.en0, self.en1, self.n_nodes)
        self.nnzb = int(self.colidx.numel())
        return self.rowptr, self.colidx

    def assemble(self, mode=L.ASM_ROWS, out=None):
        if self.rowptr is None:
            self.build_pattern()
        self.vals = self.ctx.assemble_bsr(self.x, self.y, self.z, self.en0, self.en1, self.rad, self.n_nodes,
                                          self.nnzb, self.young, self.nu, self.kappa, mode=mode, out=out)
        return self.vals

    def set_radii(self, rad):
        """Update element radii (device copy) -- geometry/pattern unchanged."""
        t = self.torch.from_numpy(np.ascontiguousarray(rad, dtype=np.float64))
        self.rad.copy_(t)
        self.vals = None

    # -- two-level preconditioner ---------------------------------------------------
    def two_level(self, fixed, n_aggregates=None):
        """Rigid-body-mode coarse space of this operator under the constraints ``fixed`` (coarse.TwoLevel; resident in
        the context until the next call).  A matrix-free user pays one temporary assembly for the Galerkin product."""
        from . import coarse
        torch = self.torch
        if self.rowptr is None:
            self.build_pattern()
        dev = self.ctx.device
        fixed_d = torch.as_tensor(np.ascontiguousarray(fixed, dtype=np.uint8)).to(dev) if not torch.is_tensor(fixed) else fixed
        vals = self.vals
        if vals is None:
            vals = self.ctx.assemble_bsr(self.x, self.y, self.z, self.en0, self.en1, self.rad, self.n_nodes, self.nnzb,
                                         self.young, self.nu, self.kappa)
        return coarse.TwoLevel(self.ctx, self.x, self.y, self.z, fixed_d, self.rowptr, self.colidx, vals, n_aggregates)

    def stretch_dominated(self):
        return stretch_dominated(self.mesh)

    TWO_LEVEL_AUTO_NODES = 20000

    def _two_level_scope(self, two_level, fixed):
        import contextlib
        if isinstance(two_level, str):
            if two_level != "auto":
                raise ValueError("two_level: None / True / number of aggregates / coarse.TwoLevel / 'auto'")
            two_level = self.n_nodes >= self.TWO_LEVEL_AUTO_NODES and self.stretch_dominated()
        if two_level is None or two_level is False:
            return contextlib.nullcontext()
        from . import coarse
        if isinstance(two_level, coarse.TwoLevel):
            return two_level
        return self.two_level(fixed, None if two_level is True else int(two_level))

    # -- solve ------------------------------------------------------------------
    def solve(self, fixed, g, f, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, keep_unconstrained=True,
              want_reactions=True, two_level=None, **pcg_kw):
        """Static solve K u = f with u[c] = g.  Returns (u, reactions, info) as device tensors/dict.

        BC algebra of simulation_base.py:480-499; the sparse LU of :502-511 is replaced
        by PCG to ``tol`` relative residual.  ``two_level``: True / a number of aggregates / a coarse.TwoLevel adds the
        rigid-body-mode coarse correction to the preconditioner (csrc/coarse.cuh)."""
        torch = self.torch
        if self.vals is None:
            self.assemble()
        dev = self.ctx.device
        fixed_d = torch.as_tensor(np.ascontiguousarray(fixed, dtype=np.uint8)).to(dev) if not torch.is_tensor(fixed) else fixed
        g_d = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64)).to(dev) if not torch.is_tensor(g) else g
        f_d = torch.as_tensor(np.ascontiguousarray(f, dtype=np.float64)).to(dev) if not torch.is_tensor(f) else f
        keep = keep_unconstrained or want_reactions
        self.vals_bc, b = self.ctx.apply_dirichlet(self.rowptr, self.colidx, self.vals, fixed_d, g_d, f_d,
                                                   inplace=not keep)
        with self._two_level_scope(two_level, fixed_d):
            u, info = self.ctx.pcg(self.rowptr, self.colidx, self.vals_bc, b, tol=tol, maxiter=maxiter,
                                   precond=precond, **pcg_kw)
        self.ctx.set_dirichlet_values(fixed_d, g_d, u)
        R = None
        if want_reactions:
            R = self.ctx.spmv(self.rowptr, self.colidx, self.vals, u)   # R = K_unconstrained u
        if not keep:
            self.vals = None
        return u, R, info

    def solve_matrix_free(self, fixed, g, f, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, want_reactions=True,
                          two_level=None, **pcg_kw):
        """Same system and BC algebra as :meth:`solve`, but K is never assembled: every product regenerates the
        element action from the geometry (csrc/matfree.cuh).  No 288 B/block matrix in HBM, ~10x fewer bytes per
        PCG iteration.  Returns (u, reactions, info) like :meth:`solve`."""
        torch = self.torch
        if self.rowptr is None:
            self.build_pattern()      # the incidence lists of the pattern drive the operator
        dev = self.ctx.device
        fixed_d = torch.as_tensor(np.ascontiguousarray(fixed, dtype=np.uint8)).to(dev) if not torch.is_tensor(fixed) else fixed
        g_d = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64)).to(dev) if not torch.is_tensor(g) else g
        f_d = torch.as_tensor(np.ascontiguousarray(f, dtype=np.float64)).to(dev) if not torch.is_tensor(f) else f
        self.ctx.matfree_setup(self.x, self.y, self.z, self.en0, self.en1, self.rad, self.n_nodes, self.young,
                               self.nu, self.kappa, fixed=fixed_d)
        b = self.ctx.matfree_rhs(g_d, f_d)
        with self._two_level_scope(two_level, fixed_d):
            u, info = self.ctx.pcg_matfree(b, tol=tol, maxiter=maxiter, precond=precond, **pcg_kw)
        self.ctx.set_dirichlet_values(fixed_d, g_d, u)
        R = self.ctx.matfree_apply(u, eliminated=False) if want_reactions else None   # R = K_unconstrained u
        return u, R, info

    def strut_topology(self):
        """Host description of the struts of a subdivided mesh (mesh.py numbering: lattice points first, elements
        beam-major from point1 to point2): element ranges and the two end joints of every strut."""
        m = self.mesh
        starts = np.flatnonzero(m.en0 < m.n_points)
        ends = np.r_[starts[1:], m.n_elems]
        sa, sb = m.en0[starts].astype(np.int32), m.en1[ends - 1].astype(np.int32)
        if (sb >= m.n_points).any() or (np.diff(np.r_[starts, m.n_elems]) < 1).any():
            raise ValueError("mesh is not beam-major between lattice points")
        return np.r_[starts, m.n_elems].astype(np.int32), sa, sb

    def solve_condensed(self, fixed, g, f, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, full_field=False,
                        two_level=None, **pcg_kw):
        """Joint-only solve: every strut (all its elements) is condensed exactly onto its two lattice points
        (``lat_assemble_bsr_struts``), the BSR system over the ``n_points`` joints is solved by the same PCG.  Needs
        loads and constraints on lattice points only -- what the reference applies -- and returns
        (u_joints [6 n_points], reactions on the joints, info): identical to the joint entries of :meth:`solve`
        with 5x (2 elements per strut) to ~70x (the reference's 18) fewer DOFs and far fewer iterations.
        ``full_field=True`` additionally back-substitutes the strut-interior nodes (``lat_strut_recover``) and returns
        u and R over ALL nodes of the mesh (R is zero on interior nodes), e.g. for the element-form gradient.
        ``two_level`` (True / number of aggregates): coarse space over the joints (coarse.TwoLevel)."""
        torch = self.torch
        m, ctx, dev = self.mesh, self.ctx, self.ctx.device
        nj = 6 * m.n_points
        fixed, g, f = np.asarray(fixed), np.asarray(g, dtype=np.float64), np.asarray(f, dtype=np.float64)
        if fixed[nj:].any() or np.any(f[nj:] != 0.0):
            raise ValueError("solve_condensed: loads / constraints on strut-interior nodes; use solve()")
        ptr, sa, sb = self.strut_topology()
        t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
        sa_d, sb_d = t(sa, np.int32), t(sb, np.int32)
        rowptr, colidx = ctx.bsr_pattern(sa_d, sb_d, m.n_points)       # pattern of the JOINT mesh (now resident)
        self.rowptr = self.colidx = self.vals = None                   # the full-mesh pattern is no longer resident
        nnzb = int(colidx.numel())
        xyz = torch.stack([self.x, self.y, self.z], dim=1).contiguous()
        ne = m.n_elems
        vals = ctx.assemble_bsr_struts(xyz, self.en0, self.en1, self.rad, t(ptr, np.int32),
                                       t(np.arange(ne), np.int32), t(np.zeros(ne), np.int32), m.n_points, nnzb,
                                       self.young, self.nu, self.kappa)
        fd, gd, fv = t(fixed[:nj], np.uint8), t(g[:nj], np.float64), t(f[:nj], np.float64)
        vbc, b = ctx.apply_dirichlet(rowptr, colidx, vals, fd, gd, fv, inplace=False)
        import contextlib
        scope = contextlib.nullcontext()
        if two_level == "auto":
            two_level = m.n_points >= self.TWO_LEVEL_AUTO_NODES and self.stretch_dominated()
        if two_level:
            from . import coarse
            npts = m.n_points
            scope = coarse.TwoLevel(ctx, self.x[:npts].contiguous(), self.y[:npts].contiguous(), self.z[:npts].contiguous(),
                                    fd, rowptr, colidx, vbc, None if two_level is True else int(two_level))
        with scope:
            u, info = ctx.pcg(rowptr, colidx, vbc, b, tol=tol, maxiter=maxiter, precond=precond, **pcg_kw)
        ctx.set_dirichlet_values(fd, gd, u)
        R = ctx.spmv(rowptr, colidx, vals, u)
        info = dict(info, n_dof_condensed=nj, n_dof_full=m.n_dof)
        if full_field:
            u_full = torch.zeros(m.n_dof, dtype=torch.float64, device=dev)
            u_full[:nj] = u
            if m.n_nodes > m.n_points:
                ctx.strut_recover(xyz, self.en0, self.en1, self.rad, t(ptr, np.int32), t(np.arange(ne), np.int32),
                                  t(np.zeros(ne), np.int32), sa_d, sb_d, int(np.diff(ptr).max()), self.young, self.nu,
                                  self.kappa, u, u_full)
            R_full = torch.zeros(m.n_dof, dtype=torch.float64, device=dev)
            R_full[:nj] = R
            return u_full, R_full, info
        return u, R, info

    def adjoint_gradient(self, u, dJdu, fixed, group, n_groups, chain=None, tol=1e-10, maxiter=200000,
                         precond=L.PC_BLOCK6):
        """dJ/d(param) for an objective J(u) with dJ/du = ``dJdu`` (zero on constrained DOFs):
        solve K lambda = dJ/du with the SAME constrained operator, then g[p] = -sum_e lambda_e^T dK_e/dr u_e.
        (LatticeOpti's displacement objectives: adjoint S lambda = dJ/du and lambda^T dS u,
        lattice_opti.py:843-902,1487-1648.)  Requires a previous ``solve`` (uses its eliminated matrix)."""
        torch = self.torch
        if self.vals_bc is None:
            raise RuntimeError("adjoint_gradient needs the constrained operator of a previous solve()")
        dev = self.ctx.device
        fd = torch.as_tensor(np.ascontiguousarray(fixed, dtype=np.uint8)).to(dev) if not torch.is_tensor(fixed) else fixed
        q = torch.as_tensor(np.ascontiguousarray(dJdu, dtype=np.float64)).to(dev) if not torch.is_tensor(dJdu) else dJdu.clone()
        q = torch.where(fd.bool(), torch.zeros_like(q), q)
        lam, info = self.ctx.pcg(self.rowptr, self.colidx, self.vals_bc, q, tol=tol, maxiter=maxiter, precond=precond)
        g = self.compliance_gradient(u, group, n_groups, chain=chain, lam=lam)
        return g, lam, info

    def compliance_gradient(self, u, group, n_groups, chain=None, lam=None):
        torch = self.torch
        dev = self.ctx.device
        grp = torch.as_tensor(np.ascontiguousarray(group, dtype=np.int32)).to(dev) if not torch.is_tensor(group) else group
        ch = None
        if chain is not None:
            ch = torch.as_tensor(np.ascontiguousarray(chain, dtype=np.float64)).to(dev) if not torch.is_tensor(chain) else chain
        return self.ctx.compliance_grad(self.x, self.y, self.z, self.en0, self.en1, self.rad, grp, n_groups, u,
                                        self.young, self.nu, self.kappa, chain=ch, lam=lam)


class FEMResult:
    """What ``solve_FEM_FenicsX`` returns as its second value, reduced to what callers read."""

    def __init__(self, fem: BeamFEM, u, reactions, info, fixed):
        self.fem = fem
        self.u = u
        self.reactions = reactions
        self.info = info
        self.fixed = fixed


def solve_FEM_B200(lattice, elements_per_strut="gmsh", tol=1e-10, maxiter=500000, precond=L.PC_BLOCK6,
                   dedup_point_loads=False, ctx=None, matrix_free=False, condense_struts=False, two_level="auto"):
    """Drop-in for ``solve_FEM_FenicsX(lattice) -> (xsol, simulationModel)``
    (utils_simulation.py:21-56).

    ``matrix_free=True`` solves the same system without assembling K (csrc/matfree.cuh): same result to the
    solver tolerance, 1.4-2.8x faster iterations and no 288 B/block matrix in HBM.
    ``condense_struts=True`` solves the exact joint-only system (:meth:`BeamFEM.solve_condensed`) and
    back-substitutes the strut-interior nodes, so ``model.u`` / ``model.R`` are the same full fields as on the other
    paths (the write-back below only touches lattice points anyway).
    ``two_level`` (True / number of aggregates / "auto"): block-Jacobi + rigid-body-mode coarse space (coarse.TwoLevel) --
    4-6x fewer iterations on stretch-dominated lattices (Octet), no gain on BCC.  "auto" (default) switches it on for
    lattices of at least 20 000 nodes whose mean joint valence says stretch-dominated (``BeamFEM.stretch_dominated``).

    Leaves ``Point.displacement_vector`` on every lattice node and
    ``Point.reaction_force_vector`` on nodes with a fixed DOF
    (full_scale_lattice_simulation.py:77-120), then returns
    ``lattice.get_global_displacement()[0]`` (utils_simulation.py:55-56).
    """
    E, nu = material_constants(lattice)
    mesh = flatten_lattice(lattice, None, elements_per_strut)
    fixed, g, f = bc_arrays_from_lattice(lattice, mesh, dedup_point_loads=dedup_point_loads)
    fem = BeamFEM(mesh, E, nu, ctx=ctx or L.default_context())     # one shared workspace across drop-in calls
    if condense_struts:      # joint-only solve + back-substitution: model.u / model.R cover all nodes like the other paths
        u, R, info = fem.solve_condensed(fixed, g, f, tol=tol, maxiter=maxiter, precond=precond, full_field=True,
                                         two_level=two_level)
    else:
        solve = fem.solve_matrix_free if matrix_free else fem.solve
        u, R, info = solve(fixed, g, f, tol=tol, maxiter=maxiter, precond=precond, two_level=two_level)
    u_h = u.cpu().numpy().reshape(-1, NDOF)
    R_h = R.cpu().numpy().reshape(-1, NDOF)
    for k, p in enumerate(mesh.meta["points"]):
        p.displacement_vector[:] = [float(v) for v in u_h[k]]
    # Reactions: same loop as full_scale_lattice_simulation.py:111-120 -- one set_reaction_force per
    # (cell, node) pair, and Point.set_reaction_force ACCUMULATES (point.py:372-385), so a clamped node
    # shared by k cells ends up with k times K.u, exactly like the reference.
    loc = {int(i): k for k, i in enumerate(mesh.point_index)}
    for cell in lattice.cells:
        for node in cell.points_cell:
            if 1 in node.fixed_DOF:
                node.set_reaction_force([float(v) for v in R_h[loc[node.index]]])
    xsol, _ = lattice.get_global_displacement()
    return xsol, FEMResult(fem, u, R, info, fixed)


def cell_sensitivities_to_parameters(lattice, q, optimization_type=None):
    """Map q[c, j] = u_c^T (dS_c/dr_j) u_c to the parameter vector of ``LatticeOpti.calculate_gradient``
    (lattice_opti.py:752-839), RAW sign (the reference flips it in ``gradient``, :719):

    * ``unit_cell``: ``grad[cell.index * n_geom + j] = q[c, j]`` (:758-761)
    * ``constant``:  hybrid -> ``grad[j] = sum_c q[c, j]``; else ``grad[0] = sum q`` (:763-784)
    * ``linear``:    ``r_cell = a . centre + d`` shared by all geometries of a cell: ``grad[i] += dC_c * centre[dir_i]``,
      ``grad[-1] += dC_c`` with ``dC_c = sum_j q[c, j]``, cells whose unclamped radius lies outside
      ``(min_radius, max_radius)`` skipped (:785-839)."""
    params = getattr(lattice, "optimization_parameters", None) or {}
    opt_type = optimization_type or params.get("type", "unit_cell")
    cells = list(lattice.cells)
    n_geom = q.shape[1]
    if opt_type == "unit_cell":
        grad = np.zeros(len(cells) * n_geom)
        for k, c in enumerate(cells):
            grad[c.index * n_geom: c.index * n_geom + n_geom] += q[k]
        return grad
    if opt_type == "constant":
        if bool(params.get("hybrid", False)):
            return q.sum(axis=0)
        n_par = int(getattr(lattice, "number_parameters", 1))
        grad = np.zeros(max(n_par, 1))
        grad[0] = q.sum()
        return grad
    if opt_type == "linear":
        dirs = params.get("direction", ["x", "y", "z"])
        if any(d not in ("x", "y", "z") for d in dirs):
            raise ValueError(f"Invalid direction in {dirs}; valid are 'x', 'y', 'z'.")
        n_par = len(dirs) + 1
        if hasattr(lattice, "number_parameters") and lattice.number_parameters != n_par:
            raise ValueError(f"Mismatch in number of linear parameters: got {lattice.number_parameters}, expected {n_par}.")
        theta = [float(v) for v in lattice.actual_optimization_parameters]
        den = getattr(lattice, "denormalize_optimization_parameters", None)
        coef = {"x": 0.0, "y": 0.0, "z": 0.0}
        for i, dkey in enumerate(dirs):
            coef[dkey] = den([theta[i]])[0] if den else theta[i]
        d0 = den([theta[-1]])[0] if den else theta[-1]
        tol = 1e-12
        grad = np.zeros(n_par)
        for k, c in enumerate(cells):
            cx, cy, cz = c.center_point
            r_un = coef["x"] * cx + coef["y"] * cy + coef["z"] * cz + d0
            if not (lattice.min_radius + tol < r_un < lattice.max_radius - tol):
                continue
            dC = float(q[k].sum())
            for i, dkey in enumerate(dirs):
                grad[i] += dC * (cx if dkey == "x" else cy if dkey == "y" else cz)
            grad[len(dirs)] += dC
        return grad
    raise NotImplementedError(f"Gradient for optimization type '{opt_type}' not implemented yet.")


def compliance_gradient_lattice(lattice, model: FEMResult, optimization_type="unit_cell"):
    """dC/d(param) in the parameter order of ``LatticeOpti.calculate_gradient`` (lattice_opti.py:752-839:
    ``unit_cell`` / ``constant`` / ``linear``), sign of :719 included.

    The reference evaluates  -sum_c u_c^T (dS_c/dr_j) u_c  with one Schur matrix per cell built from
    ``cell.beams_cell`` -- a strut shared by k cells therefore contributes to the parameter of EACH of its
    cells (cell.py:914-915 changes it from every owner).  The element form used here reproduces that by
    giving every (cell, strut) incidence its own pass of ``lat_compliance_grad``; the per-(cell, geometry)
    sums are then mapped to the parameters by :func:`cell_sensitivities_to_parameters`.
    """
    import torch
    fem = model.fem
    mesh = fem.mesh
    n_geom = len(getattr(lattice, "geom_types", [0])) if hasattr(lattice, "geom_types") else 1
    cells = list(lattice.cells)
    pos = {c.index: k for k, c in enumerate(cells)}
    n_slots = len(cells) * n_geom
    # (beam.index -> list of (cell, geometry) slots), one entry per owning cell
    owners = {}
    for c in cells:
        for b in c.beams_cell:
            j = int(getattr(b, "type_beam", 0))
            owners.setdefault(b.index, []).append(pos[c.index] * n_geom + j)
    depth = max(len(v) for v in owners.values())
    acc = torch.zeros(n_slots, dtype=torch.float64, device=fem.ctx.device)
    chain = torch.from_numpy(np.ascontiguousarray(mesh.chain, dtype=np.float64)).to(fem.ctx.device)
    for layer in range(depth):
        table = {bi: (v[layer] if layer < len(v) else -1) for bi, v in owners.items()}
        grp = np.array([table.get(int(bi), -1) for bi in mesh.beam_of_elem], dtype=np.int32)
        acc += fem.ctx.compliance_grad(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad,
                                       torch.from_numpy(grp).to(fem.ctx.device), n_slots, model.u, fem.young, fem.nu,
                                       fem.kappa, chain=chain)
    # lat_compliance_grad carries the sign of :719 (g = -sum ...); the mapping works on the raw (+) term
    q = -acc.cpu().numpy().reshape(len(cells), n_geom)
    return -cell_sensitivities_to_parameters(lattice, q, optimization_type)
