"""Oracle (test infrastructure): robust losses used on the hot path, fp64 NumPy.

Follows
  * func_step / func_piece / redescending_loss   /root/reference/src/build.py:382-395
    (called as misc.redescending_loss with (a,b,c) = (3,10,20),
     /root/reference/src/all_optimizations.py:25-27,497)
  * SciPy's Cauchy loss as used by the SBA solves /root/reference/src/calib/calib.py:335,381
    (scipy.optimize.least_squares(loss='cauchy', f_scale=C): rho(z) = ln(1+z),
     cost = 0.5 * sum C^2 rho((f/C)^2)) - SciPy is an un-vendored dependency; the formula
     is SciPy's documented one and is pinned by the notebook trace values in tests.
"""
import numpy as np

REDESC_A, REDESC_B, REDESC_C = 3.0, 10.0, 20.0


def func_step(start, x):
    return 1.0 / (1.0 + np.exp(-(x - start)))


def redescending_loss(err, a=REDESC_A, b=REDESC_B, c=REDESC_C):
    """Literal evaluation of build.py:388-395 (the logistic gates are never exactly 0/1)."""
    e = np.abs(err)
    sa, sb, sc = func_step(a, e), func_step(b, e), func_step(c, e)
    cost = (1 - sa) / 2 * e ** 2
    cost = cost + (sa - sb) * (a * e - a ** 2 / 2)
    cost = cost + (sb - sc) * (a * b - a ** 2 / 2 + (a * (c - b) / 2) * (1 - ((c - e) / (c - b)) ** 2))
    cost = cost + sc * (a * b - a ** 2 / 2 + (a * (c - b) / 2))
    return cost


def redescending_dloss(err, a=REDESC_A, b=REDESC_B, c=REDESC_C):
    """(rho(err), d rho / d err, d^2 rho / d err^2) of the literal formula.

    rho depends on |err|; d/d err = sign(err) * d/d e.  At err == 0 the one-sided
    derivative in e is non-zero (the blend has a tiny cusp there); we return
    sign(0) * (.) = 0 for the first derivative at exactly 0.
    """
    err = np.asarray(err, dtype=np.float64)
    e = np.abs(err)
    sa, sb, sc = func_step(a, e), func_step(b, e), func_step(c, e)
    dsa, dsb, dsc = sa * (1 - sa), sb * (1 - sb), sc * (1 - sc)
    ddsa, ddsb, ddsc = dsa * (1 - 2 * sa), dsb * (1 - 2 * sb), dsc * (1 - 2 * sc)
    k3 = a * (c - b) / 2
    p1, dp1, ddp1 = e * e / 2, e, 1.0
    p2, dp2 = a * e - a * a / 2, a
    u = (c - e) / (c - b)
    p3, dp3, ddp3 = a * b - a * a / 2 + k3 * (1 - u * u), a * u, -a / (c - b)
    p4 = a * b - a * a / 2 + k3
    rho = (1 - sa) * p1 + (sa - sb) * p2 + (sb - sc) * p3 + sc * p4
    d = (-dsa * p1 + (1 - sa) * dp1 + (dsa - dsb) * p2 + (sa - sb) * dp2
         + (dsb - dsc) * p3 + (sb - sc) * dp3 + dsc * p4)
    dd = (-ddsa * p1 - 2 * dsa * dp1 + (1 - sa) * ddp1
          + (ddsa - ddsb) * p2 + 2 * (dsa - dsb) * dp2
          + (ddsb - ddsc) * p3 + 2 * (dsb - dsc) * dp3 + (sb - sc) * ddp3
          + ddsc * p4)
    return rho, np.sign(err) * d, dd


def redescending_irls_weight(err, a=REDESC_A, b=REDESC_B, c=REDESC_C):
    """Gauss-Newton curvature weight psi(e) = max(rho'(e)/e, 1 - sigma_a(e)), e = |err|.

    rho'/e is the IRLS (iteratively re-weighted least squares) majoriser weight of the
    robust term; it is floored by the frozen-gate curvature of the quadratic piece
    (1 - sigma_a) e^2/2.  The floor takes over for e < ~0.45 where the literal blend has
    a tiny cusp (rho'(0+) = -0.0616 < 0) that makes rho'/e negative and ill-conditioned;
    psi >= 0 everywhere and psi -> 0 beyond c.  This is a solver design choice of the B200
    LM loop (the reference hands the exact objective to IPOPT with an L-BFGS Hessian,
    all_optimizations.py:515); cost and gradient are exact.
    """
    err = np.asarray(err, dtype=np.float64)
    e = np.abs(err)
    _, d, _ = redescending_dloss(e, a, b, c)
    floor = 1.0 - func_step(a, e)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(e > 0, d / np.where(e > 0, e, 1.0), -np.inf)
    return np.maximum(w, floor)


def cauchy_cost(f, f_scale=1.0):
    """SciPy least_squares cost with loss='cauchy': 0.5 * sum C^2 ln(1 + (f/C)^2)."""
    f = np.asarray(f, dtype=np.float64)
    z = (f / f_scale) ** 2
    return 0.5 * np.sum(f_scale ** 2 * np.log1p(z))


def cauchy_rho(z):
    """(rho, rho', rho'') of SciPy's cauchy loss as functions of z = (f/C)^2."""
    z = np.asarray(z, dtype=np.float64)
    t = 1 + z
    return np.log1p(z), 1 / t, -1 / t ** 2
