"""Analytic known-answer tests pinning the oracle's restatement of the un-vendored third-party
arithmetic (tiny-cuda-nn / nerfacc 0.3.3) -- SURVEY.md §4.  These are the only anchors available:
the reference ships no tests or golden vectors for those pieces (parity unpinned, see oracle/__init__)."""
import numpy as np
import torch

from oracle import hashgrid, sh, mlp, march, render, neus, occgrid

NERF_CFG = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.447269237440378)
NEUS_CFG = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=1.3195079107728942)


def test_level_table_matches_survey():
    lt = hashgrid.level_table(NERF_CFG)
    assert list(lt['res'][:6]) == [16, 24, 34, 49, 71, 102]
    assert list(lt['size'][:5]) == [4096, 13824, 39304, 117656, 357912] and all(lt['size'][5:] == 1 << 19)
    assert int(lt['offset'][-1]) == 6299960 and lt['n_params'] == 12599920
    assert list(lt['dense']) == [True] * 5 + [False] * 11
    lt = hashgrid.level_table(NEUS_CFG)
    assert list(lt['res'][:5]) == [32, 43, 56, 74, 98]
    assert int(lt['offset'][-1]) == 6984576 and list(lt['dense']) == [True] * 4 + [False] * 12


def test_hash_index_python_int_restatement():
    rng = np.random.default_rng(0)
    xyz = rng.integers(0, 4097, size=(256, 3))
    T = 1 << 19
    want = [((int(x) * 1) ^ (int(y) * 2654435761) ^ (int(z) * 805459861)) % (1 << 32) % T for x, y, z in xyz]
    t = torch.from_numpy(xyz)
    got = hashgrid.corner_index(t[:, 0], t[:, 1], t[:, 2], res=4096, size=T, dense=False)
    assert got.tolist() == want
    got = hashgrid.corner_index(t[:, 0] % 16, t[:, 1] % 16, t[:, 2] % 16, res=16, size=4096, dense=True)
    assert got.tolist() == [(int(x) % 16 + (int(y) % 16) * 16 + (int(z) % 16) * 256) % 4096 for x, y, z in xyz]


def test_hashgrid_reproduces_affine_function_on_dense_levels():
    cfg = dict(NERF_CFG, n_levels=3)
    lt = hashgrid.level_table(cfg)
    table = torch.zeros(int(lt['offset'][-1]), 2, dtype=torch.float64)
    a, b = torch.tensor([0.3, -0.7, 0.2], dtype=torch.float64), torch.tensor([-0.1, 0.5, 0.9], dtype=torch.float64)
    for l in range(3):
        r, off, scale = int(lt['res'][l]), int(lt['offset'][l]), float(lt['scale'][l])
        g = torch.arange(r, dtype=torch.float64)
        X, Y, Z = torch.meshgrid(g, g, g, indexing='ij')
        idx = (X + Y * r + Z * r * r).long().flatten() + off
        # vertex v sits at x = (v - 0.5)/scale
        P = torch.stack([X, Y, Z], -1).reshape(-1, 3)
        xs = (P - 0.5) / scale
        table[idx, 0] = xs @ a + 0.25
        table[idx, 1] = xs @ b - 0.5
    # keep pos = x*scale+0.5 below res-1 so the +1 corner exists (at x ~ 1 it wraps into the next row: inherent to tcnn)
    x = torch.rand(500, 3, generator=torch.Generator().manual_seed(1), dtype=torch.float64) * 0.9 + 0.01
    out = hashgrid.hashgrid_fwd(x.float(), table, lt)
    xf = x.float().double()
    for l in range(3):
        np.testing.assert_allclose(out[:, 2 * l].numpy(), (xf @ a + 0.25).numpy(), atol=1e-5)
        np.testing.assert_allclose(out[:, 2 * l + 1].numpy(), (xf @ b - 0.5).numpy(), atol=1e-5)


def test_hashgrid_autograd_gradcheck():
    cfg = dict(n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4, per_level_scale=1.7)
    lt = hashgrid.level_table(cfg)
    g = torch.Generator().manual_seed(2)
    table = torch.randn(int(lt['offset'][-1]), 2, generator=g, dtype=torch.float64, requires_grad=True)
    x = (torch.rand(6, 3, generator=g) * 0.8 + 0.1)
    xd = x.double().requires_grad_(True)
    # gradcheck w.r.t. the table (linear) and x (piecewise-trilinear; points are away from cell faces w.h.p.)
    f = lambda t: hashgrid.hashgrid_fwd(x, t, lt)
    assert torch.autograd.gradcheck(f, (table,), eps=1e-6, atol=1e-6)
    assert torch.autograd.gradcheck(lambda xx: hashgrid.hashgrid_fwd(xx, table.detach(), lt), (xd,), eps=1e-6, atol=1e-4)
    # second order: d/dtable of (dy/dx . v) exists and matches finite differences
    def gx(t):
        xx = xd.detach().clone().requires_grad_(True)
        y = hashgrid.hashgrid_fwd(xx, t, lt)
        g, = torch.autograd.grad(y.sum(), xx, create_graph=True)
        return g
    assert torch.autograd.gradcheck(gx, (table,), eps=1e-6, atol=1e-5)


def test_sh4_axes_and_orthonormality():
    def S(d):
        return sh.sh4((torch.tensor([d], dtype=torch.float64) + 1) / 2)[0]
    c0, c1 = 0.28209479177387814, 0.48860251190291987
    z = S([0., 0., 1.])
    assert abs(z[0] - c0) < 1e-12 and abs(z[2] - c1) < 1e-12 and abs(z[1]) < 1e-12 and abs(z[3]) < 1e-12
    assert abs(z[6] - (0.94617469575755997 - 0.31539156525251999)) < 1e-12
    assert abs(S([1., 0., 0.])[3] + c1) < 1e-12 and abs(S([0., 1., 0.])[1] + c1) < 1e-12
    # orthonormality under quadrature: Gauss-Legendre in cos(theta) x uniform in phi
    mu, w = np.polynomial.legendre.leggauss(16)
    phi = (np.arange(32) + 0.5) / 32 * 2 * np.pi
    MU, PHI = np.meshgrid(mu, phi, indexing='ij')
    W = np.repeat(w[:, None], 32, 1) * (2 * np.pi / 32)
    st = np.sqrt(1 - MU ** 2)
    d = torch.from_numpy(np.stack([st * np.cos(PHI), st * np.sin(PHI), MU], -1).reshape(-1, 3))
    Y = sh.sh4((d + 1) / 2).numpy()
    G = (Y * W.reshape(-1, 1)).T @ Y
    np.testing.assert_allclose(G, np.eye(16), atol=1e-10)


def test_ffmlp_layout_matches_reference_doc():
    # models/network_utils.py:156: (in_pad + out_pad) * W + (n_hidden - 1) * W^2
    for n_in, n_out, nh in [(32, 16, 1), (32, 3, 2), (35, 13, 1), (16, 3, 3)]:
        shapes, n = mlp.ffmlp_layout(n_in, n_out, 64, nh)
        assert n == (mlp.pad16(n_in) + mlp.pad16(n_out)) * 64 + (nh - 1) * 64 * 64
    # padded inputs are ones: a 35-wide input behaves like 48 with 13 trailing ones
    p = mlp.ffmlp_init(35, 13, 64, 1)
    x = torch.randn(5, 35)
    y = mlp.ffmlp_fwd(x, p, 35, 13, emulate_fp16=False, compute_dtype=torch.float64)
    W1, W2 = p[:64 * 48].view(64, 48).double(), p[64 * 48:].view(16, 64).double()
    xp = torch.cat([x.double(), torch.ones(5, 13, dtype=torch.float64)], -1)
    np.testing.assert_allclose(y.numpy(), (torch.relu(xp @ W1.t()) @ W2.t())[:, :13].numpy(), atol=1e-12)


def _box_scene(R=16):
    binary = np.zeros((R, R, R), bool)
    binary[R // 4: 3 * R // 4, R // 4: 3 * R // 4, R // 4: 3 * R // 4] = True
    return binary


def test_ray_aabb_and_march_kats():
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    o = np.array([[-3, 0, 0], [-3, 0, 0], [0, 0, 0], [-3, 5, 0]], np.float32)
    d = np.array([[1, 0, 0], [-1, 0, 0], [0, 0, 1], [1, 0, 0]], np.float32)
    tmin, tmax = march.ray_aabb_intersect(o, d, aabb)
    np.testing.assert_allclose(tmin, [2, 1e10, 0, 1e10])
    np.testing.assert_allclose(tmax, [4, 1e10, 1, 1e10])
    step = np.float32(0.01)
    # all-ones grid: contiguous samples covering [t_min, t_max); empty grid / miss: nothing
    ones = np.ones((4, 4, 4), bool)
    ri, ts, te, pk = march.march_lattice(o, d, aabb, ones, step, tmin, tmax)
    assert pk[:, 1].tolist() == [200, 0, 100, 0] or abs(pk[0, 1] - 200) <= 1 and abs(pk[2, 1] - 100) <= 1
    assert pk[1, 1] == 0 and pk[3, 1] == 0
    r0 = ri == 0
    np.testing.assert_allclose(ts[r0][1:], te[r0][:-1])  # contiguous
    np.testing.assert_allclose(te[r0] - ts[r0], step, rtol=1e-3)
    ri, *_ = march.march_lattice(o, d, aabb, np.zeros((4, 4, 4), bool), step, tmin, tmax)
    assert ri.size == 0
    # box of half the extent: ray 0 crosses [-0.5, 0.5] => ~100 samples with midpoints inside
    ri, ts, te, pk = march.march_lattice(o, d, aabb, _box_scene(), step, tmin, tmax)
    tm = 0.5 * (ts + te)[ri == 0]
    assert abs(pk[0, 1] - 100) <= 1 and (np.abs(-3 + tm) <= 0.5 + 1e-6).all()


def test_lattice_equals_dda_reference():
    rng = np.random.default_rng(3)
    R = 16
    binary = rng.random((R, R, R)) < 0.3
    aabb = np.array([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], np.float32)
    n = 24
    o = (rng.normal(size=(n, 3)) * 0.2 + np.array([0, 0, -4.0])).astype(np.float32)
    d = rng.normal(size=(n, 3)) * 0.25 + np.array([0, 0, 1.0])
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    step = np.float32(1.732 * 3 / 256)
    tmin, tmax = march.ray_aabb_intersect(o, d, aabb)
    ri, ts, te, pk = march.march_lattice(o, d, aabb, binary, step, tmin, tmax)
    mismatched = 0
    for r in range(n):
        ref = march.march_dda_reference(o[r], d[r], aabb, binary, float(step), float(tmin[r]), float(tmax[r]))
        mine = list(zip(ts[ri == r], te[ri == r]))
        if len(ref) != len(mine):
            mismatched += 1  # a midpoint within rounding of a cell face may flip; must be rare
            continue
        if ref:
            np.testing.assert_allclose(np.array(ref), np.array(mine), atol=2e-4)
    assert mismatched <= 2


def test_march_sequential_cone():
    # all-occupied contracted grid: steps grow geometrically once t*cone > step
    o = np.zeros((1, 3), np.float32)
    d = np.array([[0, 0, 1]], np.float32)
    roi = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    cone = 10 ** (3 / 64) - 1
    ri, ts, te, pk = march.march_sequential(o, d, roi, np.ones((2, 2, 2), bool), 0.01, cone, np.array([0.1], np.float32),
                                            np.array([1e3], np.float32), march.UN_BOUNDED_SPHERE)
    dt = te - ts
    assert (np.abs(ts[1:] - te[:-1]) == 0).all()
    assert np.allclose(dt[ts * cone < 0.01], 0.01, rtol=1e-5)
    big = ts * cone > 0.0101
    assert np.allclose(dt[big] / ts[big], cone, rtol=1e-4)
    assert 80 <= len(ts) <= 110


def test_render_closed_forms_and_gradients():
    sigma, L, n = 3.0, 1.2, 60
    t = torch.linspace(0, L, n + 1, dtype=torch.float64)
    ts, te = t[:-1, None], t[1:, None]
    ri = torch.zeros(n, dtype=torch.long)
    sig = torch.full((n, 1), sigma, dtype=torch.float64, requires_grad=True)
    w = render.render_weight_from_density(ts, te, sig, ri, 1)
    op = render.accumulate_along_rays(w, ri, None, 1)
    np.testing.assert_allclose(op.item(), 1 - np.exp(-sigma * L), rtol=1e-12)
    # alpha path agrees with density path
    alpha = 1 - torch.exp(-sig.detach() * (te - ts))
    w2 = render.render_weight_from_alpha(alpha, ri, 1)
    np.testing.assert_allclose(w2.numpy(), w.detach().numpy(), rtol=1e-10)
    # analytic gradient: d w_j / d sigma_i
    g = torch.randn(n, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    gs, = torch.autograd.grad((w * g).sum(), sig)
    T = render.transmittance_from_density(sig.detach().view(-1), (te - ts).view(-1), ri, 1)
    wd, gd, dl = w.detach().view(-1), g.view(-1), (te - ts).view(-1)
    suffix = torch.flip(torch.cumsum(torch.flip(gd * wd, [0]), 0), [0]) - gd * wd
    np.testing.assert_allclose(gs.view(-1).numpy(), (dl * (gd * (T - wd) - suffix)).numpy(), rtol=1e-9, atol=1e-12)
    # two rays, visibility: early termination drops a suffix only
    a = torch.tensor([0.5] * 20 + [0.8] * 10, dtype=torch.float64)
    ri2 = torch.tensor([0] * 20 + [1] * 10)
    keep, Tr = render.render_visibility(a, ri2, 2, 1e-4)
    assert keep[:14].all() and not keep[14:20].any()      # 0.5^13 = 1.2e-4 >= 1e-4 > 0.5^14
    assert keep[20:26].all() and not keep[26:].any()      # 0.2^5 = 3.2e-4 kept, 0.2^6 = 6.4e-5 dropped
    assert abs(Tr[20].item() - 1.0) < 1e-15


def test_neus_alpha_planar_closed_form():
    # planar SDF f(x) = x.n - c, ray hitting the plane frontally: cos = -1
    s, dist = 20.0, 0.01
    sdf = torch.tensor([0.03, 0.0, -0.02], dtype=torch.float64)
    nrm = torch.tensor([[0., 0., 1.]] * 3, dtype=torch.float64)
    dirs = torch.tensor([[0., 0., -1.]] * 3, dtype=torch.float64)
    for ratio in (0.0, 1.0):
        a = neus.get_alpha(sdf, nrm, dirs, torch.full((3, 1), dist, dtype=torch.float64), torch.tensor(s, dtype=torch.float64), ratio)
        prev, nxt = torch.sigmoid((sdf + dist / 2) * s), torch.sigmoid((sdf - dist / 2) * s)
        np.testing.assert_allclose(a.numpy(), ((prev - nxt + 1e-5) / (prev + 1e-5)).numpy(), rtol=1e-12)
    # back-facing ray at ratio 1: iter_cos = 0 => alpha = 1e-5/(c+1e-5)
    a = neus.get_alpha(sdf, nrm, -dirs, torch.full((3, 1), dist, dtype=torch.float64), torch.tensor(s, dtype=torch.float64), 1.0)
    c = torch.sigmoid(sdf * s)
    np.testing.assert_allclose(a.numpy(), (1e-5 / (c + 1e-5)).numpy(), rtol=1e-9)


def test_occgrid_update_rule_and_bit_packing():
    R = 8
    occs = torch.zeros(R ** 3)
    cells = torch.arange(R ** 3)
    jit = torch.full((R ** 3, 3), 0.5)
    fn = lambda x: (x.norm(dim=-1) < 0.8).float()[:, None] * 0.5
    occs1, b1 = occgrid.update(occs, cells, jit, fn, 1.5, 0, R, occ_thre=0.01)
    centers = ((occgrid.cell_coords(cells, R).float() + 0.5) / R * 3 - 1.5)
    assert torch.equal(b1.view(-1), centers.norm(dim=-1) < 0.8)
    occs2, _ = occgrid.update(occs1, cells, jit, lambda x: torch.zeros(len(x), 1), 1.5, 0, R)
    np.testing.assert_allclose(occs2.numpy(), occs1.numpy() * 0.95, rtol=1e-6)
    bits = occgrid.pack_bits(b1.numpy())
    flat = b1.view(-1).numpy()
    for idx in (0, 1, 37, 255, 300, 511):
        assert bool((bits[idx >> 5] >> (idx & 31)) & 1) == bool(flat[idx])


def test_neus_field_manual_formulas_match_autograd():
    """oracle/neus_field.py (the hand-derived first + second order backward the fused NeuS kernels implement) == torch autograd of
    VolumeSDF's analytic-normal construction (models/geometry.py:158-180), in float64."""
    from oracle import neus_field
    cfg = dict(n_levels=5, n_features_per_level=2, log2_hashmap_size=9, base_resolution=4, per_level_scale=1.6)
    lt = hashgrid.level_table(cfg)
    g = torch.Generator().manual_seed(0)
    n, n_in, n_out, r = 40, 3 + 2 * 5, 13, 1.5
    table = (torch.randn(lt['n_params'], generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
    W1 = (torch.randn(64, n_in, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    b1 = (torch.randn(64, generator=g, dtype=torch.float64) * 0.01).requires_grad_(True)
    W2 = (torch.randn(n_out, 64, generator=g, dtype=torch.float64) * 0.2).requires_grad_(True)
    b2 = (torch.randn(n_out, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    pts = ((torch.rand(n, 3, generator=g, dtype=torch.float64) * 2 - 1) * 1.2).requires_grad_(True)
    # autograd path, as the reference builds it
    x01 = (pts + r) / (2 * r)
    enc = torch.cat([x01 * 2 - 1, hashgrid.hashgrid_fwd(x01, table.view(-1, 2), lt)], -1)
    out = torch.nn.functional.softplus(enc @ W1.t() + b1, beta=100) @ W2.t() + b2
    sdf = out[:, 0]
    grad, = torch.autograd.grad(sdf, pts, torch.ones_like(sdf), create_graph=True)
    g_out = torch.randn(n, n_out, generator=g, dtype=torch.float64)
    g_grad = torch.randn(n, 3, generator=g, dtype=torch.float64)
    ((out * g_out).sum() + (grad * g_grad).sum()).backward()
    # manual path
    sdf_m, grad_m, out_m, cache = neus_field.forward(pts.detach(), table.detach(), lt, W1.detach(), b1.detach(), W2.detach(), b2.detach(), r)
    np.testing.assert_allclose(out_m.numpy(), out.detach().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(grad_m.numpy(), grad.detach().numpy(), rtol=1e-8, atol=1e-10)
    gm = neus_field.backward(cache, table.detach(), lt, W1.detach(), b1.detach(), W2.detach(), b2.detach(), r, g_out, g_grad)
    for name, ref in (('W1', W1.grad), ('b1', b1.grad), ('W2', W2.grad), ('b2', b2.grad), ('table', table.grad)):
        np.testing.assert_allclose(gm[name].numpy(), ref.numpy(), rtol=1e-7, atol=1e-9, err_msg=name)


def test_adamw_oracle_matches_torch():
    """the reference's optimizer is torch.optim.AdamW itself (systems/utils.py:314-325): pin the restatement against it"""
    from oracle import optim
    torch.manual_seed(0)
    p0 = torch.randn(4099) * 0.1
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pt], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p, m, v = p0.numpy().copy(), np.zeros(4099, np.float32), np.zeros(4099, np.float32)
    for step in range(1, 8):
        g = torch.randn(4099) * (10.0 ** -(step % 4))
        g[::7] = 0.0
        pt.grad = g.clone()
        opt.step()
        optim.adamw_step(p, g.numpy().copy(), m, v, step)
        np.testing.assert_allclose(p, pt.detach().numpy(), rtol=2e-6, atol=2e-8)
    st = opt.state[pt]
    np.testing.assert_allclose(m, st['exp_avg'].numpy(), rtol=2e-6, atol=1e-7 * float(np.abs(m).max()))  # lerp cancellation near 0
    np.testing.assert_allclose(v, st['exp_avg_sq'].numpy(), rtol=2e-6, atol=1e-7 * float(v.max()))


def _unit_grid(n, r=1.0):
    g = np.linspace(-r, r, n, dtype=np.float32)
    return np.meshgrid(g, g, g, indexing='ij')


def test_marching_cubes_oracle_properties_and_product_case_table():
    """oracle/mcubes.py (PyMCubes stand-in, parity unpinned) pinned by what a mesh consumer relies on, and the product's generated case
    table (instant-nsr-pl_b200/mc_table.py -> csrc/mc_table.inc) checked against the oracle's independent per-cell construction."""
    import importlib.util
    import os
    from oracle import mcubes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('nsr_mc_table', os.path.join(root, 'instant-nsr-pl_b200', 'mc_table.py'))
    mt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mt)
    table = mt.build_table()
    assert len(table) == 256 and table[0] == [] and table[255] == [] and max(len(t) for t in table) == 5
    for case in range(256):
        assert mcubes.case_triangles(case) == table[case], case
    # the committed header is the generator's output
    inc = open(os.path.join(root, 'instant-nsr-pl_b200', 'csrc', 'mc_table.inc')).read()
    rows = [ln for ln in inc.splitlines() if ln.startswith('    {')]
    assert len(rows) == 256
    for case, ln in enumerate(rows):
        flat = [int(v) for v in ln.strip(' {},').split(',')]
        want = [e for t in table[case] for e in t]
        assert flat[:len(want)] == want and all(v == -1 for v in flat[len(want):]) and len(flat) == 16

    # sphere: closed, outward, vertices on the iso-crossings, area / volume converge from below
    X, Y, Z = _unit_grid(33)
    sdf = np.sqrt(X * X + Y * Y + Z * Z) - np.float32(0.6)
    v, f = mcubes.marching_cubes(sdf, 0.0, lo=(-1, -1, -1), hi=(1, 1, 1), negate=True)
    assert mcubes.directed_edge_defects(f) == 0
    assert len(v) - 3 * len(f) // 2 + len(f) == 2                      # Euler characteristic of a sphere (E = 3F/2 on a closed mesh)
    vol, area = mcubes.signed_volume(v, f), mcubes.area(v, f)
    assert 0.98 * (4 / 3 * np.pi * 0.6 ** 3) < vol < 4 / 3 * np.pi * 0.6 ** 3
    assert 0.99 * (4 * np.pi * 0.36) < area < 4 * np.pi * 0.36
    assert np.abs(np.linalg.norm(v, axis=1) - 0.6).max() < 1e-3       # linear interpolation of a distance field
    # each vertex lies on a grid edge (two lattice coordinates) of the box
    lat = (v + 1) / 2 * 32
    assert ((np.abs(lat - np.round(lat)) < 1e-4).sum(axis=1) >= 2).all()
    # flipping the sign flips the orientation (the reference negates the level for exactly this reason, geometry.py:62)
    v2, f2 = mcubes.marching_cubes(sdf, 0.0, lo=(-1, -1, -1), hi=(1, 1, 1), negate=False)
    assert mcubes.signed_volume(v2, f2) < 0 and len(f2) == len(f)
    # torus: genus 1
    q = np.sqrt(X * X + Y * Y) - np.float32(0.55)
    tor = np.sqrt(q * q + Z * Z) - np.float32(0.22)
    v, f = mcubes.marching_cubes(tor, 0.0, negate=True)
    assert mcubes.directed_edge_defects(f) == 0 and len(v) - 3 * len(f) // 2 + len(f) == 0
    # white noise (every ambiguous configuration occurs) with an outside border: still closed, and non-cubic grids work
    rng = np.random.default_rng(0)
    fld = rng.standard_normal((14, 15, 16)).astype(np.float32)
    fld[0] = fld[-1] = fld[:, 0] = fld[:, -1] = -10
    fld[:, :, 0] = fld[:, :, -1] = -10
    v, f = mcubes.marching_cubes(fld, 0.1)
    assert len(f) > 5000 and mcubes.directed_edge_defects(f) == 0 and mcubes.signed_volume(v, f) > 0
    # nothing to extract
    v, f = mcubes.marching_cubes(np.ones((4, 4, 4), np.float32), 2.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)


def test_hashgrid_one_gather_form_matches_the_per_corner_form():
    """the timed CPU baseline evaluates the hash grid with one indexing op over all corners (bench.py): same values and gradients"""
    cfg = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.5)
    lt = hashgrid.level_table(cfg)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(257, 3, generator=g)
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 2e-6)):
        t1 = (hashgrid.init_table(lt, dtype=dtype) * 1e3).requires_grad_(True)
        t2 = t1.detach().clone().requires_grad_(True)
        go = torch.randn(257, 32, generator=g).to(dtype)
        a = hashgrid.hashgrid_fwd(x, t1, lt, compute_dtype=dtype)
        b = hashgrid.hashgrid_fwd(x, t2, lt, compute_dtype=dtype, one_gather=True)
        (a * go).sum().backward()
        (b * go).sum().backward()
        np.testing.assert_allclose(b.detach().numpy(), a.detach().numpy(), rtol=0, atol=tol)
        np.testing.assert_allclose(t2.grad.numpy(), t1.grad.numpy(), rtol=0, atol=tol * 10)


def test_dense_level_index_wrap_equals_the_modulo_for_every_in_range_cell():
    """csrc/common.cuh nsr_corner_indices (round 2): on dense levels the table index is wrapped with `i = min(i, 2 size - 1); i -= (i >= size) *
    size` instead of `i % size` (the integer division was 11 % of the per-ray forward kernel's instructions).  The two agree iff i < 2 size for
    every corner of every cell a position in [0, 1]^3 can fall into -- checked here exhaustively on the corner extremes for every dense level
    of the shipped grid configs (tcnn index: x + y res + z res^2, cell = floor(scale x + 0.5) in [0, res - 1], corner offsets 0 / 1)."""
    from nsr_b200 import configs, ops
    seen = 0
    for cfg in (configs.nerf_blender()['geometry']['xyz_encoding_config'], configs.neus_blender()['geometry']['xyz_encoding_config'],
                configs.neus_dtu()['geometry_bg']['xyz_encoding_config'] if 'geometry_bg' in configs.neus_dtu() else None):
        if cfg is None:
            continue
        g = ops.GridSpec(cfg)
        for l in range(g.n_levels):
            if not g.dense[l]:
                continue
            r, size = int(g.res[l]), int(g.size[l])
            assert r >= 2 and r ** 3 <= size
            # cell coordinates 0 .. r - 1 (scale x + 0.5 <= scale + 0.5 < r), corners add 0 / 1 per axis
            cells = np.arange(r, dtype=np.int64)
            top = (cells[:, None, None] + 1) + (cells[None, :, None] + 1) * r + (cells[None, None, :] + 1) * r * r   # the largest corner index per cell
            assert int(top.max()) == r + r * r + r ** 3 and int(top.max()) < 2 * size
            i = top.reshape(-1)
            wrapped = np.minimum(i, 2 * size - 1)
            wrapped = wrapped - (wrapped >= size) * size
            assert np.array_equal(wrapped, i % size)
            seen += 1
    assert seen >= 5   # nerf-blender: levels 0..4 are dense
