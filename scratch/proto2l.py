import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, '.')
from pylatticedso_b200 import mesh as M
from oracle import lattice_oracle as O
E_MOD, NU = 1.0e5 if False else 210000.0, 0.3

def build(geom, nc, m_el):
    lat = M.synthetic_lattice(geom, nc, [0.05])
    m = M.mesh_from_synthetic(lat, m_el)
    en = np.stack([m.en0, m.en1], 1)
    K = O.assemble_csr(m.xyz, en, m.rad, E_MOD, NU)
    fixed, g, f = M.compression_bc(m)
    return lat, m, K, fixed, g, f

def rbm_Z(xyz, agg, n_agg, fixed):
    n = xyz.shape[0]
    cen = np.zeros((n_agg, 3)); cnt = np.bincount(agg, minlength=n_agg)
    for k in range(3): cen[:, k] = np.bincount(agg, xyz[:, k], minlength=n_agg) / np.maximum(cnt, 1)
    d = xyz - cen[agg]
    rows, cols, vals = [], [], []
    for i3 in range(3):       # translations
        rows.append(6*np.arange(n)+i3); cols.append(6*agg+i3); vals.append(np.ones(n))
        rows.append(6*np.arange(n)+3+i3); cols.append(6*agg+3+i3); vals.append(np.ones(n))
    # u = omega x d : u_x = wy dz - wz dy ; u_y = wz dx - wx dz ; u_z = wx dy - wy dx
    for (ui, wk, comp, sgn) in [(0,1,2,1),(0,2,1,-1),(1,2,0,1),(1,0,2,-1),(2,0,1,1),(2,1,0,-1)]:
        rows.append(6*np.arange(n)+ui); cols.append(6*agg+3+wk); vals.append(sgn*d[:, comp])
    Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6*n, 6*n_agg))
    Z = sp.diags((~fixed.astype(bool)).astype(float)) @ Z
    return Z

def run(geom, nc, m_el, box):
    lat, m, K, fixed, g, f = build(geom, nc, m_el)
    Kbc, rhs = O.apply_dirichlet(K, fixed, g, f)
    Kbc = Kbc.tocsr()
    n = m.n_nodes
    # block jacobi
    B = sp.bsr_matrix(Kbc, blocksize=(6, 6))
    D = np.zeros((n, 6, 6))
    B.sort_indices()
    for i in range(n):
        for j in range(B.indptr[i], B.indptr[i+1]):
            if B.indices[j] == i: D[i] = B.data[j]
    Dinv = np.linalg.inv(D)
    def bj(r): return np.einsum('nij,nj->ni', Dinv, r.reshape(n, 6)).ravel()
    xyz = m.xyz
    lo = xyz.min(0); hi = xyz.max(0)
    nb = np.maximum(1, np.round((hi-lo)/box).astype(int))
    ijk = np.minimum(((xyz-lo)/(hi-lo+1e-12)*nb).astype(int), nb-1)
    agg = (ijk[:,0]*nb[1]+ijk[:,1])*nb[2]+ijk[:,2]
    n_agg = int(nb.prod())
    Z = rbm_Z(xyz, agg, n_agg, fixed)
    Ec = (Z.T @ Kbc @ Z).toarray()
    w, V = np.linalg.eigh(Ec)
    keep = w > 1e-10*w.max()
    Einv = (V[:, keep] / w[keep]) @ V[:, keep].T
    def two(r): return bj(r) + Z @ (Einv @ (Z.T @ r))
    res = {}
    for name, Mfun in (('bj', bj), ('2l', two)):
        it = [0]
        def cb(x): it[0] += 1
        x, info = spl.cg(Kbc, rhs, rtol=1e-8, maxiter=20000, M=spl.LinearOperator(Kbc.shape, matvec=Mfun), callback=cb)
        res[name] = it[0]
    print(geom, nc, 'm', m_el, 'dof', 6*n, 'box', box, 'n_agg', n_agg, 'n_c', 6*n_agg, 'rank', keep.sum(), res, flush=True)

if __name__ == '__main__':
    run('BCC', (8,8,8), 1, 2.0)
    run('BCC', (8,8,8), 1, 1.0)
    run('BCC', (12,12,12), 1, 2.0)
    run('BCC', (12,12,12), 2, 2.0)
    run('BCC', (12,12,12), 2, 3.0)
    run('Octet', (10,10,10), 1, 2.0)
