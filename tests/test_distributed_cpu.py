"""CPU (gloo, world_size 2): the host-side slab partition / halo logic of the multi-GPU path.
The device exchange (ncclSend/ncclRecv) is replaced by torch.distributed gloo send/recv and the
local operators are assembled with the CPU oracle, so the index logic is what is tested."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import E_MOD, NU, ROOT


def _exchange(part, vec_nodes6):
    """gloo version of lat_halo_exchange: vec_nodes6 [n_local, 6]."""
    reqs, recv_bufs = [], []
    ro = 0
    for q, sl, rc in zip(part.peers, part.send_lists, part.recv_counts):
        if len(sl):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec_nodes6[sl])), dst=q))
        if rc:
            buf = torch.empty((rc, 6), dtype=torch.float64)
            reqs.append(dist.irecv(buf, src=q))
            recv_bufs.append((ro, rc, buf))
        ro += rc
    for r in reqs:
        r.wait()
    for ro, rc, buf in recv_bufs:
        vec_nodes6[part.n_owned + ro: part.n_owned + ro + rc] = buf.numpy()


def _worker(rank, world, port, geom, ncell, m_, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pylatticedso_b200 import mesh as M
        from pylatticedso_b200 import distributed as D
        from oracle import lattice_oracle as orc
        lat = M.synthetic_lattice(geom, ncell, [0.04])
        mesh = M.mesh_from_synthetic(lat, m_)
        part = D.partition_slab(mesh, rank, world)
        lm = D.local_mesh(mesh, part)
        # every node is owned exactly once
        cnt = torch.zeros(mesh.n_nodes, dtype=torch.int64)
        cnt[torch.from_numpy(part.owned)] = 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())
        # complete rows: every element touching an owned node is local
        touched = np.isin(mesh.en0, part.owned) | np.isin(mesh.en1, part.owned)
        assert np.array_equal(np.flatnonzero(touched), part.local_elems)
        # ghost section ordered by owner then global id, contiguous per peer
        assert np.all(np.diff(part.ghost_owner) >= 0)
        for qq in part.peers:
            g = part.ghosts[part.ghost_owner == qq]
            assert np.all(np.diff(g) > 0)
        # distributed SpMV == global SpMV on the owned rows
        Kg = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
        Kl = orc.assemble_csr(lm.xyz, np.stack([lm.en0, lm.en1], 1), lm.rad, E_MOD, NU)
        rng = np.random.default_rng(5)
        xg = rng.standard_normal(mesh.n_dof)
        xl = np.zeros((part.n_local, 6))
        xl[: part.n_owned] = xg.reshape(-1, 6)[part.owned]          # ghosts unknown until exchanged
        _exchange(part, xl)
        assert np.array_equal(xl[part.n_owned:], xg.reshape(-1, 6)[part.ghosts])
        yl = (Kl @ xl.ravel()).reshape(-1, 6)[: part.n_owned]
        yg = (Kg @ xg).reshape(-1, 6)[part.owned]
        assert np.abs(yl - yg).max() < 1e-12 * np.abs(yg).max()
        # dot products over owned DOFs all-reduce to the global dot
        d = torch.tensor([float((xl[: part.n_owned] ** 2).sum())], dtype=torch.float64)
        dist.all_reduce(d)
        assert abs(float(d) - float(xg @ xg)) < 1e-10 * float(xg @ xg)
        q.put((rank, "ok", part.n_owned, len(part.ghosts)))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc(), 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("geom,ncell,m_", [("BCC", (4, 2, 2), 2), ("Octet", (4, 2, 3), 1), ("BCC", (5, 2, 2), 1)])
def test_slab_partition_two_ranks_gloo(geom, ncell, m_):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + hash((geom, ncell, m_))) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, geom, ncell, m_, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[1]
    assert sum(r[2] for r in res) > 0 and all(r[3] > 0 for r in res)


def test_slab_bounds_and_owner():
    from pylatticedso_b200 import distributed as D
    b = D.slab_bounds(0.0, 100.0, 8)
    assert list(np.diff(b)) == [12, 13, 12, 13, 12, 13, 12, 13] or sum(np.diff(b)) == 100
    o = D.node_owner_by_x(np.array([0.0, 11.99, 12.0, 99.9, 100.0]), b)
    assert list(o) == [0, 0, 1, 7, 7]


def test_slab_bounds_from_nodes_never_collapse():
    """cell size 0.1 and more ranks than the unit-cell rule could serve collapsed the cuts; the
    data-derived rule balances node counts, and an impossible request fails identically on every rank."""
    from pylatticedso_b200 import distributed as D
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice("BCC", (10, 2, 2), [0.005], cell_size=(0.1, 0.1, 0.1))
    mesh = M.mesh_from_synthetic(lat, 2)
    b = D.slab_bounds_from_nodes(mesh.x, 4)
    assert np.all(np.diff(b) > 0)
    own = np.bincount(D.node_owner_by_x(mesh.x, b), minlength=4)
    assert own.min() > 0 and own.max() < 1.5 * own.mean()
    parts = [D.partition_slab(mesh, r, 4) for r in range(4)]
    assert sum(p.n_owned for p in parts) == mesh.n_nodes
    for p in parts:                       # send list of p -> q has the length q expects from p
        for q, sl in zip(p.peers, p.send_lists):
            assert parts[q].recv_counts[parts[q].peers.index(p.rank)] == len(sl)
    with pytest.raises(ValueError):
        D.partition_slab(mesh, 0, 64)     # more ranks than x-planes: every rank raises before any collective
    with pytest.raises(ValueError):
        D.slab_bounds(0.0, 1.0, 4, cell=1.0)


@pytest.mark.parametrize("geom,ncell,m_,world,grad", [("BCC", (6, 2, 3), 2, 3, None), ("Octet", (5, 2, 2), 1, 2, ("linear", [True, False, True], [0.1, 0, 0.05])),
                                                      ("BCC", (4, 2, 2), 1, 4, None)])
def test_per_slab_generation_equals_partition_of_the_global_mesh(geom, ncell, m_, world, grad):
    """generate_slab builds only the rank's cell layers (+1 overlap layer); the local mesh, the halo lists and the
    boundary conditions must be IDENTICAL to what partition_slab extracts from the global mesh with cuts on the same
    cell planes -- including the first-creator radius rule on graded lattices."""
    from pylatticedso_b200 import distributed as D
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice(geom, ncell, [0.04], grad_radius=grad)
    mesh = M.mesh_from_synthetic(lat, m_)
    fixed, g, f = M.compression_bc(mesh)
    layers = D.slab_layers(ncell[0], world)
    bounds = np.array([float(a) for a, _ in layers] + [float(ncell[0])])
    bounds[1:-1] -= 1e-9
    owner = D.node_owner_by_x(mesh.x, bounds)
    n_owned_total = 0
    for r in range(world):
        ref_part = D.partition_slab(mesh, r, world, owner=owner)
        ref_lm = D.local_mesh(mesh, ref_part)
        lm, part = D.generate_slab(geom, ncell, [0.04], m_, r, world, grad_radius=grad)
        for k in ("x", "y", "z", "en0", "en1", "rad"):
            assert np.array_equal(getattr(lm, k), getattr(ref_lm, k)), (r, k)
        assert part.peers == ref_part.peers and part.recv_counts == ref_part.recv_counts
        assert all(np.array_equal(a, b) for a, b in zip(part.send_lists, ref_part.send_lists))
        assert (part.n_owned, part.n_local) == (ref_part.n_owned, ref_part.n_local)
        dofs = D.local_dofs(ref_part)
        fl, gl, fl2 = D.compression_bc_local(lm)
        assert np.array_equal(fl, fixed[dofs]) and np.array_equal(gl, g[dofs]) and np.array_equal(fl2, f[dofs])
        n_owned_total += part.n_owned
    assert n_owned_total == mesh.n_nodes


def test_joint_only_partition_covers_every_strut_and_keeps_the_geometry():
    """Host logic of DistributedJointFEM: the joint mesh is partitioned like any mesh; every strut is local to the
    owner(s) of its two joints with ALL its elements, in a numbering [joints | strut-interior nodes]."""
    from pylatticedso_b200 import distributed as D, mesh as M
    lat = M.synthetic_lattice("Octet", (5, 2, 2), [0.04])
    mesh = M.mesh_from_synthetic(lat, 3)
    ptr, sa, sb = D.strut_topology(mesh)
    assert ptr[-1] == mesh.n_elems and (np.diff(ptr) == 3).all()
    jm = D.joint_mesh(mesh, sa, sb)
    world = 3
    owned_total, seen = 0, np.zeros(mesh.n_elems, dtype=int)
    for r in range(world):
        part = D.partition_slab(jm, r, world)
        f = D.local_strut_mesh(mesh, ptr, part)
        owned_total += part.n_owned
        seen[f["elems_global"]] += 1
        nj = part.n_local
        assert f["xyz"].shape[0] == nj + f["interior_global"].shape[0]
        assert (f["len1"][f["chain_ptr"][1:] - 1] < nj).all() and (f["len0"][f["chain_ptr"][:-1]] < nj).all()   # strut ends are joints
        glob = mesh.xyz[mesh.en1[f["elems_global"]]] - mesh.xyz[mesh.en0[f["elems_global"]]]
        np.testing.assert_array_equal(f["xyz"][f["len1"]] - f["xyz"][f["len0"]], glob)
        np.testing.assert_array_equal(f["rad"], mesh.rad[f["elems_global"]])
        lm = D.local_mesh(jm, part)                                  # strut ends in local joint numbering, same order
        np.testing.assert_array_equal(lm.en0, f["len0"][f["chain_ptr"][:-1]])
        np.testing.assert_array_equal(lm.en1, f["len1"][f["chain_ptr"][1:] - 1])
    assert owned_total == mesh.n_points
    assert seen.min() == 1 and seen.max() == 2                       # struts across a cut are condensed on both sides


def test_generate_slab_single_rank_fast_path_equals_the_general_path():
    """world = 1: the identity partition is built directly; same mesh, same partition as the general code."""
    from pylatticedso_b200 import distributed as D
    for geom, n, m_ in (("Octet", (5, 4, 3), 1), ("BCC", (4, 3, 3), 3)):
        a, pa = D.generate_slab(geom, n, [0.04], m_, 0, 1)
        b, pb = D.generate_slab(geom, n, [0.04], m_, 0, 1, _general=True)
        for k in ("x", "y", "z", "en0", "en1", "rad", "beam_of_elem", "chain", "point_index", "cell_of_elem"):
            va, vb = getattr(a, k), getattr(b, k)
            assert np.array_equal(va, vb) and va.dtype == vb.dtype, k
        assert a.n_points == b.n_points and np.array_equal(a.meta["is_point"], b.meta["is_point"])
        assert np.array_equal(a.meta["global_nodes"], b.meta["global_nodes"])
        assert pa.n_owned == pb.n_owned and pa.n_local == pb.n_local and pa.peers == pb.peers == []
        assert np.array_equal(pa.local_elems, pb.local_elems) and np.array_equal(pa.local_nodes, pb.local_nodes)
        assert pa.send_lists == pb.send_lists == [] and pa.recv_counts == pb.recv_counts == []
