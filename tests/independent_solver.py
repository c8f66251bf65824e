"""TEST INFRASTRUCTURE ONLY -- an independent discrete-ordinate flux solve in numpy.

Used to pin the C oracle (oracle/disort_oracle.c) on multi-layer beam and thermal problems: it
shares no code and no algorithmic shortcut with either the oracle (reduced nn x nn ASYMTX
eigenproblem + LINPACK band solve) or the CUDA kernels (symmetrised Jacobi + adding sweep).
It takes the full N x N stream matrix, a general `numpy.linalg.eig`, dense particular-solution
solves and one dense global boundary-value matrix.  Mathematically the discrete-ordinate
solution is unique given (double-Gauss quadrature, delta-M, boundary conditions), so agreement
to ~1e-11 is expected; DISORT's absorption cut-off is NOT applied here.

Conventions follow DISORT: layers top-down, mu > 0 upward.  `planck_levels` are the
band-integrated Planck values at the nlyr+1 levels (the PLKAVG series is pinned separately).
"""
import numpy as np


def double_gauss(nstr):
    n = nstr // 2
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def solve_fluxes(dtau, ssa, pmom, nstr, *, fbeam=0.0, umu0=1.0, albedo=0.0, fisot=0.0,
                 planck_levels=None, top_emission=0.0, bottom_planck=0.0):
    """dtau[L], ssa[L], pmom[L, >=nstr+1] (moment 0 = 1).  Returns (flup[L+1], fdn_total[L+1])."""
    dtau, ssa, pmom = (np.asarray(a, dtype=np.float64) for a in (dtau, ssa, pmom))
    L, N, n = len(dtau), nstr, nstr // 2
    mu_h, w_h = double_gauss(N)
    mu = np.concatenate([-mu_h, mu_h])          # first n downward, last n upward
    wt = np.concatenate([w_h, w_h])
    # delta-M
    f = pmom[:, N] if pmom.shape[1] > N else np.zeros(L)
    om = ssa * (1 - f) / (1 - f * ssa)
    dt = (1 - f * ssa) * dtau
    chi = (pmom[:, :N] - f[:, None]) / (1 - f[:, None])
    tau = np.concatenate([[0.0], np.cumsum(dt)])
    tau_un = np.concatenate([[0.0], np.cumsum(dtau)])
    ls = np.arange(N)
    P = np.polynomial.legendre.legvander(mu, N - 1)               # P[i, l]
    P0 = np.polynomial.legendre.legvander(np.array([-umu0]), N - 1)[0]
    beam = fbeam > 0
    lam, G, Zb, Z0, Z1 = [], [], [], [], []
    for lc in range(L):
        gl = (2 * ls + 1) * om[lc] * chi[lc]
        D = 0.5 * (P * gl) @ P.T * wt[None, :]
        A = (np.eye(N) - D) / mu[:, None]
        ev, vec = np.linalg.eig(A)
        assert np.max(np.abs(ev.imag)) < 1e-9 * np.max(np.abs(ev.real))
        lam.append(ev.real)
        G.append(vec.real)
        if beam:
            X = fbeam / (4 * np.pi) * (P * gl) @ P0
            Zb.append(np.linalg.solve(np.eye(N) + np.diag(mu) / umu0 - D, X))
        else:
            Zb.append(np.zeros(N))
        if planck_levels is not None:
            b1 = (planck_levels[lc + 1] - planck_levels[lc]) / dt[lc] if dt[lc] > 0 else 0.0
            b0 = planck_levels[lc] - b1 * tau[lc]
            z1 = np.linalg.solve(np.eye(N) - D, (1 - om[lc]) * b1 * np.ones(N))
            z0 = np.linalg.solve(np.eye(N) - D, (1 - om[lc]) * b0 * np.ones(N) + mu * z1)
            Z0.append(z0)
            Z1.append(z1)
        else:
            Z0.append(np.zeros(N))
            Z1.append(np.zeros(N))

    def hom(lc, t):
        """matrix H[i, j] = G_ij * scaled exponential at optical depth t inside layer lc"""
        ref = np.where(lam[lc] > 0, tau[lc + 1], tau[lc])
        return G[lc] * np.exp(lam[lc] * (t - ref))[None, :]

    def part(lc, t):
        return Zb[lc] * np.exp(-t / umu0) * (1.0 if beam else 0.0) + Z0[lc] + Z1[lc] * t

    M = np.zeros((N * L, N * L))
    rhs = np.zeros(N * L)
    row = 0
    dn, up = slice(0, n), slice(n, N)
    # top: downward streams
    M[row:row + n, 0:N] = hom(0, 0.0)[dn]
    rhs[row:row + n] = fisot + top_emission - part(0, 0.0)[dn]
    row += n
    for lc in range(L - 1):
        t = tau[lc + 1]
        M[row:row + N, lc * N:(lc + 1) * N] = hom(lc, t)
        M[row:row + N, (lc + 1) * N:(lc + 2) * N] = -hom(lc + 1, t)
        rhs[row:row + N] = part(lc + 1, t) - part(lc, t)
        row += N
    # bottom: Lambertian surface
    t = tau[L]
    H, p = hom(L - 1, t), part(L - 1, t)
    refl = 2.0 * albedo * (wt[dn] * np.abs(mu[dn]))                # (albedo/pi) * 2 pi sum w mu u-
    M[row:row + n, (L - 1) * N:] = H[up] - np.outer(np.ones(n), refl @ H[dn])
    rhs[row:row + n] = (-p[up] + refl @ p[dn] + albedo * umu0 * fbeam / np.pi * np.exp(-t / umu0) * beam
                        + (1 - albedo) * bottom_planck)
    C = np.linalg.solve(M, rhs)
    flup, fdn = np.zeros(L + 1), np.zeros(L + 1)
    for lev in range(L + 1):
        lc = min(lev, L - 1)
        u = hom(lc, tau[lev]) @ C[lc * N:(lc + 1) * N] + part(lc, tau[lev])
        flup[lev] = 2 * np.pi * np.sum(wt[up] * mu[up] * u[up])
        fdn[lev] = 2 * np.pi * np.sum(wt[dn] * np.abs(mu[dn]) * u[dn]) + beam * umu0 * fbeam * np.exp(-tau[lev] / umu0)
    return flup, fdn
