"""Closed-form identities of the reference's hard-coded matrices, checked on the oracle's derived
versions: Q_base / basis (rbp_planner.hpp L327-L347) and Aeq_base (L353-L405).  Nothing here reads the reference."""
from math import comb

import numpy as np

import oracle


def bernstein(i, n, t):
    return comb(n, i) * t ** i * (1 - t) ** (n - i)


def test_basis_rows_are_the_monomial_coefficients_of_the_quintic_bernstein_polynomials():
    _, B = oracle.Q_base_and_basis()
    t = np.linspace(0, 1, 13)
    for i in range(6):
        assert np.allclose(np.polyval(B[i], t), bernstein(i, 5, t), atol=1e-13)     # highest power first
    assert np.allclose(B.sum(0), [0, 0, 0, 0, 0, 1])                                  # partition of unity


def test_q_base_is_the_jerk_gram_matrix_of_a_quintic_bezier_curve():
    Q, B = oracle.Q_base_and_basis()
    assert np.array_equal(Q, Q.T) and np.array_equal(Q, np.rint(Q)) and np.all(Q % 120 == 0)
    # c' Q c = integral over [0, 1] of (third derivative)^2, by Gauss-Legendre on random control points
    rng = np.random.default_rng(0)
    xs, ws = np.polynomial.legendre.leggauss(8)
    xs, ws = 0.5 * (xs + 1), 0.5 * ws
    for _ in range(5):
        c = rng.standard_normal(6)
        mono = c @ B                                            # highest power first
        jerk = np.polyval(np.polyder(mono, 3), xs)
        assert np.isclose(c @ Q @ c, (ws * jerk ** 2).sum(), rtol=1e-12)
    # rank 3: quadratics have zero jerk
    assert np.linalg.matrix_rank(Q) == 3
    for p in ([0, 0, 1], [0, 1, 0], [1, 0, 0]):                 # control points of 1, t, t^2 in the Bernstein basis
        ctrl = np.linalg.solve(B.T, np.concatenate([np.zeros(3), p]))
        assert np.abs(Q @ ctrl).max() < 1e-9


def test_aeq_base_rows_are_endpoint_states_and_c2_continuity():
    T = np.array([0.0, 1.0, 2.5, 3.0, 5.0])
    M = len(T) - 1
    A = oracle.Aeq_base(T)
    assert A.shape == (3 * M + 3, 6 * M)
    _, B = oracle.Q_base_and_basis()
    rng = np.random.default_rng(1)
    ctrl = rng.standard_normal((M, 6))
    dt = np.diff(T)

    def state(m, tau):                                          # (pos, vel, acc) of segment m at local parameter tau in [0, 1]
        mono = ctrl[m] @ B
        return np.array([np.polyval(np.polyder(mono, d), tau) / dt[m] ** d for d in range(3)])

    y = A @ ctrl.reshape(-1)
    assert np.allclose(y[0:3], state(0, 0.0)) and np.allclose(y[3:6], state(M - 1, 1.0))
    for m in range(1, M):
        assert np.allclose(y[6 + 3 * (m - 1): 6 + 3 * m], state(m - 1, 1.0) - state(m, 0.0), atol=1e-12)
