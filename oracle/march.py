"""Occupancy-grid ray marching (nerfacc 0.3.3 ``ray_aabb_intersect`` / ``ray_marching``) -- CPU oracle.

[3P, parity unpinned] restated from SURVEY.md Appendix A.1/A.2; reference call sites
models/nerf.py:82-93, models/neus.py:153-169,209-220.

All arithmetic is float32 numpy, operation by operation, in exactly the order the CUDA kernels
use (they are compiled with -fmad=false and spell out every fma), so sample SETS are compared
bit-exactly.  Two marchers:

* ``march_lattice``  -- cone_angle == 0 (AABB mode, every reference fg config): sample k of a ray
  spans [fma(k,dt,t_min), fma(k+1,dt,t_min)); it is emitted iff its midpoint lies in an occupied
  cell.  nerfacc's voxel-skipping DDA only ever jumps in multiples of dt inside an empty voxel, so
  this brute-force lattice is the same set (``march_dda_reference`` restates the DDA loop to
  demonstrate that, up to the rounding of its chained ``t += dt``).
* ``march_sequential`` -- cone_angle > 0 (contracted background, models/neus.py:141-169): blind
  stepping with dt = clamp(t*cone_angle, step, 1e10), chained t0 <- t1.
"""
import numpy as np

F32 = np.float32
AABB = 0
UN_BOUNDED_SPHERE = 2


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def ray_aabb_intersect(rays_o, rays_d, aabb):
    """-> (t_min[N], t_max[N]) float32; miss => (1e10, 1e10); hit => (max(near,0), far)."""
    o = np.asarray(rays_o, F32)
    d = np.asarray(rays_d, F32)
    aabb = np.asarray(aabb, F32)
    with np.errstate(divide='ignore', invalid='ignore'):
        t1 = (aabb[None, 0:3] - o) / d
        t2 = (aabb[None, 3:6] - o) / d
    lo = np.fmin(t1, t2)
    hi = np.fmax(t1, t2)
    near = np.fmax(np.fmax(lo[:, 0], lo[:, 1]), lo[:, 2])
    far = np.fmin(np.fmin(hi[:, 0], hi[:, 1]), hi[:, 2])
    near0 = np.fmax(near, F32(0))
    with np.errstate(invalid='ignore'):
        hit = far > near0
    t_min = np.where(hit, near0, F32(1e10)).astype(F32)
    t_max = np.where(hit, far, F32(1e10)).astype(F32)
    return t_min, t_max


def ray_interval(rays_o, rays_d, scene_aabb, near_plane, far_plane, step, jitter):
    """t_min/t_max per ray as ``ray_marching`` prepares them."""
    n = rays_o.shape[0]
    if scene_aabb is not None:
        t_min, t_max = ray_aabb_intersect(rays_o, rays_d, scene_aabb)
    else:
        t_min = np.zeros(n, F32)
        t_max = np.full(n, 1e10, F32)
    if near_plane is not None:
        t_min = np.fmax(t_min, np.broadcast_to(np.asarray(near_plane, F32), (n,))).astype(F32)
    if far_plane is not None:
        t_max = np.fmin(t_max, F32(far_plane)).astype(F32)
    if jitter is not None:  # stratified: one U[0,1) draw per ray
        t_min = (t_min + np.asarray(jitter, F32) * F32(step)).astype(F32)
    return t_min, t_max


def occupied(p, roi, binary, contraction):
    """p [..,3] float32 world points -> bool.  binary: bool [R,R,R] (flat ix*R*R + iy*R + iz)."""
    R = binary.shape[0]
    lo, hi = roi[0:3].astype(F32), roi[3:6].astype(F32)
    # (p - lo) * fp32(1 / (hi - lo)): one fp32 reciprocal of the extent, then a multiply (what the kernels do; nerfacc divides --
    # the two differ only for points within one ulp of a cell face)
    inv = (F32(1) / (hi - lo)).astype(F32)
    unit = ((p - lo).astype(F32) * inv).astype(F32)
    if contraction == AABB:
        inside = np.all((p >= lo) & (p <= hi), axis=-1)
    else:
        u = (unit * F32(2) - F32(1)).astype(F32)
        n = np.sqrt(((u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1]).astype(F32) + u[..., 2] * u[..., 2]).astype(F32))
        with np.errstate(divide='ignore', invalid='ignore'):
            s = (F32(2) - F32(1) / n).astype(F32)
            uc = (s[..., None] * (u / n[..., None]).astype(F32)).astype(F32)
        u = np.where((n > 1)[..., None], uc, u)
        unit = (u * F32(0.25) + F32(0.5)).astype(F32)
        inside = np.ones(p.shape[:-1], bool)
    with np.errstate(invalid='ignore'):
        cell = np.clip((unit * F32(R)).astype(F32), -1e9, 1e9)
    cell = np.clip(np.nan_to_num(cell).astype(np.int64), 0, R - 1)
    flat = cell[..., 0] * R * R + cell[..., 1] * R + cell[..., 2]
    return inside & binary.reshape(-1)[flat]


def pack(counts):
    start = np.zeros(len(counts), np.int64)
    start[1:] = np.cumsum(counts)[:-1]
    return np.stack([start, counts], axis=-1)


def march_lattice(rays_o, rays_d, roi, binary, step, t_min, t_max, contraction=AABB, chunk=2048):
    """-> ray_indices int32 [M], t_starts f32 [M], t_ends f32 [M], packed_info int64 [N,2]."""
    o = np.asarray(rays_o, F32)
    d = np.asarray(rays_d, F32)
    step = F32(step)
    n = o.shape[0]
    ri, ts, te, counts = [], [], [], np.zeros(n, np.int64)
    for b in range(0, n, chunk):
        sl = slice(b, min(n, b + chunk))
        tmin, tmax = t_min[sl], t_max[sl]
        span = np.where(tmax > tmin, (tmax.astype(np.float64) - tmin) / step, 0)
        kmax = int(min(np.nanmax(span) if span.size else 0, 1 << 20)) + 3
        k = np.arange(kmax, dtype=F32)[None, :]
        t0 = fma(k, step, tmin[:, None])
        t1 = fma(k + F32(1), step, tmin[:, None])
        tm = ((t0 + t1).astype(F32) * F32(0.5)).astype(F32)
        valid = tm < tmax[:, None]
        p = fma(tm[..., None], d[sl][:, None, :], o[sl][:, None, :])
        occ = occupied(p, np.asarray(roi, F32), binary, contraction) & valid
        rr, kk = np.nonzero(occ)
        ri.append((rr + b).astype(np.int32))
        ts.append(t0[rr, kk])
        te.append(t1[rr, kk])
        counts[sl] = occ.sum(axis=1)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return cat(ri, np.int32), cat(ts, F32), cat(te, F32), pack(counts)


def march_sequential(rays_o, rays_d, roi, binary, step, cone_angle, t_min, t_max, contraction,
                     max_steps=1 << 16):
    """Blind stepping with growing dt (contracted background pass)."""
    o = np.asarray(rays_o, F32)
    d = np.asarray(rays_d, F32)
    step, cone = F32(step), F32(cone_angle)
    n = o.shape[0]

    def dt(t):
        return np.fmin(np.fmax((t * cone).astype(F32), step), F32(1e10)).astype(F32)

    t0 = t_min.astype(F32).copy()
    t1 = (t0 + dt(t0)).astype(F32)
    tm = ((t0 + t1).astype(F32) * F32(0.5)).astype(F32)
    rec = [[] for _ in range(n)]
    for _ in range(max_steps):
        act = tm < t_max
        if not act.any():
            break
        p = fma(tm[:, None], d, o)
        occ = occupied(p, np.asarray(roi, F32), binary, contraction) & act
        for r in np.nonzero(occ)[0]:
            rec[r].append((t0[r], t1[r]))
        t0 = np.where(act, t1, t0).astype(F32)
        t1 = np.where(act, (t0 + dt(t0)).astype(F32), t1).astype(F32)
        tm = np.where(act, ((t0 + t1).astype(F32) * F32(0.5)).astype(F32), tm).astype(F32)
    counts = np.array([len(r) for r in rec], np.int64)
    ri = np.repeat(np.arange(n, dtype=np.int32), counts)
    flat = [x for r in rec for x in r]
    ts = np.array([a for a, _ in flat], F32)
    te = np.array([b for _, b in flat], F32)
    return ri, ts, te, pack(counts)


def march_dda_reference(o, d, roi, binary, step, t_min, t_max):
    """One ray, pure python: the nerfacc AABB-mode loop with voxel skipping,
    cone_angle = 0.  Only used by a KAT to show lattice == DDA."""
    o = np.asarray(o, np.float64)
    d = np.asarray(d, np.float64)
    roi = np.asarray(roi, np.float64)
    R = binary.shape[0]
    lo, hi = roi[:3], roi[3:]
    out = []
    t0, t1 = float(t_min), float(t_min) + step
    tm = 0.5 * (t0 + t1)
    while tm < t_max:
        p = o + tm * d
        inside = np.all((p >= lo) & (p <= hi))
        unit = (p - lo) / (hi - lo)
        cell = np.clip((unit * R).astype(np.int64), 0, R - 1)
        if inside and binary[cell[0], cell[1], cell[2]]:
            out.append((t0, t1))
            t0 = t1
            t1 = t0 + step
            tm = 0.5 * (t0 + t1)
        else:
            g = unit * R
            with np.errstate(divide='ignore', invalid='ignore'):
                tx = (np.floor(g + 0.5 + 0.5 * np.sign(d)) - g) / d / R * (hi - lo)
            t_exit = max(np.nanmin(tx), 0.0)
            target = tm + t_exit
            tm += step
            while tm < target:
                tm += step
            t0, t1 = tm - 0.5 * step, tm + 0.5 * step
    return out
