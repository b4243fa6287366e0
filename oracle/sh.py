"""Spherical harmonics degree 4 (tiny-cuda-nn ``SphericalHarmonics``) -- CPU oracle.

[3P, parity unpinned] restated from SURVEY.md Appendix A.5; reference call site
models/network_utils.py:90 <- configs/nerf-blender.yaml:59-61; input convention (dir+1)/2 from
models/texture.py:24.  Pinned by KATs: axis values + orthonormality under quadrature.
"""
import torch


def sh4(v01):
    """v01 [N,3] in [0,1] (tcnn convention); returns the 16 real SH basis values of d = 2*v01-1."""
    d = v01 * 2.0 - 1.0
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return torch.stack(out, dim=-1)
