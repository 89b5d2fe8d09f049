"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy restatement of ``cv2.VariationalRefinement.calc(I0, I1, flow)`` as the reference calls it at
``tobac_flow/flow.py:359,513-519`` (``vr_model = cv2.VariationalRefinement.create()``, defaults alpha=20, delta=5,
gamma=10, omega=1.6, 5 fixed-point x 5 SOR iterations, zeta=0.1, epsilon=1e-3).

The arithmetic lives in OpenCV (opencv/modules/video/src/variational_refinement.cpp; not vendored in
/root/reference).  Restated here in plain (non red-black-buffer) form: Brox-style data terms (brightness + gradient
constancy) on derivatives of the averaged image, a robust first-order smoothness term, red-black SOR.
Pinned against cv2 itself in ``tests/test_oracle_varref.py``.
"""
import numpy as np

F32 = np.float32

try:  # pragma: no cover
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None

DEFAULTS = dict(alpha=20.0, delta=5.0, gamma=10.0, omega=1.6, fixed_point_iterations=5, sor_iterations=5,
                zeta=0.1, epsilon=0.001)


def _dx(a):
    """Sobel(a, dx=1, dy=0, ksize=1, BORDER_REPLICATE): a[x+1] - a[x-1]."""
    p = np.pad(a, ((0, 0), (1, 1)), mode="edge")
    return (p[:, 2:] - p[:, :-2]).astype(F32)


def _dy(a):
    p = np.pad(a, ((1, 1), (0, 0)), mode="edge")
    return (p[2:, :] - p[:-2, :]).astype(F32)


def warp_linear(img, u, v):
    """remap(I1 as CV_32F, x + u, y + v, INTER_LINEAR, BORDER_REPLICATE) — via the oracle's remap restatement."""
    from . import remap_np
    h, w = img.shape
    px = (np.arange(w, dtype=F32)[None, :] + u).astype(F32)
    py = (np.arange(h, dtype=F32)[:, None] + v).astype(F32)
    return remap_np.remap_linear_replicate(img.astype(F32), px, py)


def variational_refinement(I0, I1, flow, alpha=20.0, delta=5.0, gamma=10.0, omega=1.6, fixed_point_iterations=5,
                           sor_iterations=5, zeta=0.1, epsilon=0.001, trace=None):
    I0 = np.asarray(I0)
    I1 = np.asarray(I1)
    W_u = np.ascontiguousarray(flow[..., 0], dtype=F32)
    W_v = np.ascontiguousarray(flow[..., 1], dtype=F32)
    h, w = W_u.shape
    I0f = I0.astype(F32)
    warped = warp_linear(I1, W_u, W_v)
    avg = (F32(0.5) * I0f + F32(0.5) * warped).astype(F32)
    Ix, Iy = _dx(avg), _dy(avg)
    Iz = (warped - I0f).astype(F32)
    Ixx, Ixy, Iyy = _dx(Ix), _dy(Ix), _dy(Iy)
    Ixz, Iyz = _dx(Iz), _dy(Iz)
    if trace is not None:
        trace.update(warped=warped, Ix=Ix, Iy=Iy, Iz=Iz, Ixx=Ixx, Ixy=Ixy, Iyy=Iyy, Ixz=Ixz, Iyz=Iyz)

    zeta2 = F32(zeta * zeta)
    eps2 = F32(epsilon * epsilon)
    delta2 = F32(delta / 2)
    gamma2 = F32(gamma / 2)
    alpha2 = F32(alpha / 2)
    omega = F32(omega)

    du = np.zeros((h, w), F32)
    dv = np.zeros((h, w), F32)
    tu, tv = W_u.copy(), W_v.copy()
    yy, xx = np.mgrid[0:h, 0:w]
    red = ((yy + xx) % 2) == 0

    def fwd_x(a):
        out = np.zeros_like(a)
        out[:, :-1] = a[:, 1:] - a[:, :-1]
        return out

    def fwd_y(a):
        out = np.zeros_like(a)
        out[:-1, :] = a[1:, :] - a[:-1, :]
        return out

    for _ in range(fixed_point_iterations):
        # ---- data term ------------------------------------------------------------------------------------------
        dn = Ix * Ix + Iy * Iy + zeta2
        Ik1z = Iz + Ix * du + Iy * dv
        wgt = (delta2 / np.sqrt(Ik1z * Ik1z / dn + eps2)) / dn
        A11 = wgt * (Ix * Ix) + zeta2
        A12 = wgt * (Ix * Iy)
        A22 = wgt * (Iy * Iy) + zeta2
        b1 = -wgt * (Iz * Ix)
        b2 = -wgt * (Iz * Iy)
        dn1 = Ixx * Ixx + Ixy * Ixy + zeta2
        dn2 = Iyy * Iyy + Ixy * Ixy + zeta2
        Ik1zx = Ixz + Ixx * du + Ixy * dv
        Ik1zy = Iyz + Ixy * du + Iyy * dv
        wgt = gamma2 / np.sqrt(Ik1zx * Ik1zx / dn1 + Ik1zy * Ik1zy / dn2 + eps2)
        A11 = A11 + wgt * (Ixx * Ixx / dn1 + Ixy * Ixy / dn2)
        A12 = A12 + wgt * (Ixx * Ixy / dn1 + Ixy * Iyy / dn2)
        A22 = A22 + wgt * (Ixy * Ixy / dn1 + Iyy * Iyy / dn2)
        b1 = b1 - wgt * (Ixx * Ixz / dn1 + Ixy * Iyz / dn2)
        b2 = b2 - wgt * (Ixy * Ixz / dn1 + Iyy * Iyz / dn2)
        # ---- smoothness term: weights from the current flow, right-hand side from the input flow ---------------
        ux, vx, uy, vy = fwd_x(tu), fwd_x(tv), fwd_y(tu), fwd_y(tv)
        sw = (alpha2 / np.sqrt(ux * ux + vx * vx + uy * uy + vy * vy + eps2)).astype(F32)
        # horizontal edges (x, x+1)
        wx = sw.copy()
        wx[:, -1] = 0
        gux, gvx = fwd_x(W_u), fwd_x(W_v)
        A11 = A11 + wx
        A22 = A22 + wx
        b1 = b1 + wx * gux
        b2 = b2 + wx * gvx
        A11[:, 1:] += wx[:, :-1]
        A22[:, 1:] += wx[:, :-1]
        b1[:, 1:] -= (wx * gux)[:, :-1]
        b2[:, 1:] -= (wx * gvx)[:, :-1]
        # vertical edges (y, y+1)
        wy = sw.copy()
        wy[-1, :] = 0
        guy, gvy = fwd_y(W_u), fwd_y(W_v)
        A11 = A11 + wy
        A22 = A22 + wy
        b1 = b1 + wy * guy
        b2 = b2 + wy * gvy
        A11[1:, :] += wy[:-1, :]
        A22[1:, :] += wy[:-1, :]
        b1[1:, :] -= (wy * guy)[:-1, :]
        b2[1:, :] -= (wy * gvy)[:-1, :]
        A11, A12, A22, b1, b2 = (a.astype(F32) for a in (A11, A12, A22, b1, b2))
        # ---- red-black SOR ----------------------------------------------------------------------------------------
        for _s in range(sor_iterations):
            for colour in (red, ~red):
                pu = np.pad(du, 1)
                pv = np.pad(dv, 1)
                pwx = np.pad(wx, 1)
                pwy = np.pad(wy, 1)
                sig_u = (pwx[1:-1, :-2] * pu[1:-1, :-2] + pwx[1:-1, 1:-1] * pu[1:-1, 2:] +
                         pwy[:-2, 1:-1] * pu[:-2, 1:-1] + pwy[1:-1, 1:-1] * pu[2:, 1:-1])
                new_u = du + omega * ((sig_u + b1 - dv * A12) / A11 - du)
                du = np.where(colour, new_u, du).astype(F32)
                sig_v = (pwx[1:-1, :-2] * pv[1:-1, :-2] + pwx[1:-1, 1:-1] * pv[1:-1, 2:] +
                         pwy[:-2, 1:-1] * pv[:-2, 1:-1] + pwy[1:-1, 1:-1] * pv[2:, 1:-1])
                new_v = dv + omega * ((sig_v + b2 - du * A12) / A22 - dv)
                dv = np.where(colour, new_v, dv).astype(F32)
        tu = (W_u + du).astype(F32)
        tv = (W_v + dv).astype(F32)
    return np.stack([tu, tv], -1)
