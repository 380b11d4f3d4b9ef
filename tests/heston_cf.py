"""Closed-form Heston European call by characteristic-function quadrature
(test-only analytic anchor).  The reference's MC has no
drift and no discounting (src/HSimulation.tpp:75-80), so MC prices are compared
with r = 0; r is a parameter only to validate this formula against the
reference's SWIFT known-answer prices (src/UnitTest.cpp:155-188,221-261, r=0.02)."""
import numpy as np
from scipy.integrate import quad


def heston_call(S, K, T, v0, theta, rho, kappa, sigma, r=0.0):
    x = np.log(S)

    def cf(u, j):
        # Albrecher et al. "little Heston trap" form
        uj = 0.5 if j == 1 else -0.5
        bj = kappa - rho * sigma if j == 1 else kappa
        a = kappa * theta
        d = np.sqrt((rho * sigma * 1j * u - bj) ** 2 - sigma ** 2 * (2 * uj * 1j * u - u * u))
        g = (bj - rho * sigma * 1j * u - d) / (bj - rho * sigma * 1j * u + d)
        e = np.exp(-d * T)
        C = r * 1j * u * T + a / sigma ** 2 * ((bj - rho * sigma * 1j * u - d) * T
                                               - 2 * np.log((1 - g * e) / (1 - g)))
        D = (bj - rho * sigma * 1j * u - d) / sigma ** 2 * (1 - e) / (1 - g * e)
        return np.exp(C + D * v0 + 1j * u * x)

    def P(j):
        f = lambda u: (np.exp(-1j * u * np.log(K)) * cf(u, j) / (1j * u)).real
        val, _ = quad(f, 1e-12, 400.0, limit=2000, epsabs=1e-12, epsrel=1e-12)
        return 0.5 + val / np.pi

    return S * P(1) - K * np.exp(-r * T) * P(2)
