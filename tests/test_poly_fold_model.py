"""CPU model of the polynomial / folded form of the source coordinates (csrc/edf_fx.cuh: edf_poly_build, edf_poly_fold).

Along a column (fixed z, x) of the output and inside one control interval of y, the displacement is a cubic in the
row's fractional control position u; the kernels fold the row index itself (y + off = (j + u) r, r = (I - 1)/(P - 1)),
the crop offset and the affine map into that cubic.  This test restates the construction in NumPy and checks it
against the straightforward evaluation (B-spline of the mirror-extended control grid through
scipy.ndimage.map_coordinates, then deform.c:771-781): agreement to ~1e-11 voxel, i.e. far inside the 2^-21
re-evaluation zone of the kernels."""
import numpy as np
import scipy.ndimage


def bspline3(u):
    v = 1.0 - u
    return np.array([v ** 3 / 6.0, (u * u * (u - 2.0) * 3.0 + 4.0) / 6.0, (v * v * (v - 2.0) * 3.0 + 4.0) / 6.0, u ** 3 / 6.0])


def mirror(i, n):
    if n <= 1:
        return 0
    p = 2 * n - 2
    i = abs(i) % p
    return p - i if i >= n else i


def column_polynomial(D, idim, off, A, z, x, j):
    """Coefficients c[h][k] of the source coordinate along axis h as a cubic in u, for output column (z, x) and the
    control interval floor(cp_y) == j -- the arithmetic of edf_poly_build + edf_poly_fold (without the 1.5 * 2^29)."""
    P = D.shape[1:]
    cpz = (P[0] - 1) * (z + off[0]) / (idim[0] - 1)
    cpx = (P[2] - 1) * (x + off[2]) / (idim[2] - 1)
    jz, jx = int(np.floor(cpz)), int(np.floor(cpx))
    wz, wx = bspline3(cpz - jz), bspline3(cpx - jx)
    r = (idim[1] - 1) / (P[1] - 1)
    out = np.zeros((3, 4))
    for h in range(3):
        E = np.zeros(4)
        for jj in range(4):
            for i in range(4):
                for k in range(4):
                    E[jj] += D[h, mirror(jz - 1 + i, P[0]), mirror(j - 1 + jj, P[1]), mirror(jx - 1 + k, P[2])] * wz[i] * wx[k]
        a = np.array([(E[0] + 4 * E[1] + E[2]) / 6.0, (E[2] - E[0]) / 2.0, (E[0] - 2 * E[1] + E[2]) / 2.0,
                      ((E[3] - E[0]) + 3 * (E[1] - E[2])) / 6.0])
        yo = j * r - off[1]                                   # output row index at u = 0
        out[h] = a
        out[h, 0] += A[h, 0] * z + A[h, 1] * yo + A[h, 2] * x + A[h, 3] + off[h]
        out[h, 1] += A[h, 1] * r
    return out


def test_folded_cubic_reproduces_the_source_coordinates():
    rng = np.random.default_rng(3)
    idim = (40, 90, 56)
    for case in range(4):
        P = [(5, 5, 5), (3, 4, 6), (4, 7, 3), (2, 2, 2)][case]
        D = rng.standard_normal((3,) + P) * 6.0
        off = [(0, 0, 0), (3, 11, 5), (0, 20, 0), (7, 0, 9)][case]
        th = np.deg2rad(10.0 * case)
        A = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, np.cos(th), -np.sin(th), 2.0 * case], [0.0, np.sin(th), np.cos(th), -1.5 * case]])
        if case == 0:
            A = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)             # no affine map
        odim = tuple(n - o for n, o in zip(idim, off))
        worst = 0.0
        for _ in range(40):
            z, x = int(rng.integers(0, odim[0])), int(rng.integers(0, odim[2]))
            ys = np.arange(odim[1])
            cp = [np.full(ys.shape, (P[0] - 1) * (z + off[0]) / (idim[0] - 1)),
                  (P[1] - 1) * (ys + off[1]) / (idim[1] - 1),
                  np.full(ys.shape, (P[2] - 1) * (x + off[2]) / (idim[2] - 1))]
            disp = [scipy.ndimage.map_coordinates(D[h], cp, order=3, mode="mirror", prefilter=False) for h in range(3)]
            ref = [A[h, 0] * z + A[h, 1] * ys + A[h, 2] * x + A[h, 3] + off[h] + disp[h] for h in range(3)]
            jy = np.floor(cp[1]).astype(int)
            for j in np.unique(jy):
                rows = np.nonzero(jy == j)[0]
                c = column_polynomial(D, idim, off, A, z, x, int(j))
                u = cp[1][rows] - j
                for h in range(3):
                    val = ((c[h, 3] * u + c[h, 2]) * u + c[h, 1]) * u + c[h, 0]
                    worst = max(worst, float(np.abs(val - ref[h][rows]).max()))
        assert worst < 1e-10, (case, worst)
