"""CPU restatement of the matrix-Fisher normalising constant (training-side consumer)
-- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows reference losses/matrix_fisher_loss.py:9-192: `bessel0_exp_scaled` (:31-48, Horner :15-28),
`torch_trapezoid_integral` (:51-73), the forward integrand (:76-99), the backward integrand (:102-131) and
`LogMFNormConstant.forward/backward` (:147-192). PINNED: `oracle/make_golden.py` asserts bit equality with the imported
reference on the fixtures (tests/golden/mf_norm.npz)."""
import torch

_A = [1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.360768e-1, 0.45813e-2][::-1]
_B = [0.39894228, 0.1328592e-1, 0.225319e-2, -0.157565e-2, 0.916281e-2, -0.2057706e-1, 0.2635537e-1, -0.1647633e-1,
      0.392377e-2][::-1]
NUM_TRAPS = 512


def _horner(coeffs, x):
    z = torch.full_like(x, coeffs[0])
    for c in coeffs[1:]:
        z = z * x + c
    return z


def bessel0_exp_scaled(x):
    ax = x.abs()
    small = _horner(_A, (ax / 3.75) ** 2) / torch.exp(ax)
    large = _horner(_B, 3.75 / ax) / torch.sqrt(ax)
    return torch.where(ax <= 3.75, small, large)


def _trapezoid(func, s):
    i = torch.arange(NUM_TRAPS, dtype=s.dtype)
    u = (i * (2 / (NUM_TRAPS - 1)) + (-1)).view(1, NUM_TRAPS)
    w = torch.ones(1, NUM_TRAPS, dtype=s.dtype)
    w[0, 0] = 0.5
    w[0, -1] = 0.5
    return torch.sum(func(u, s) * w, dim=1) * 2 / (NUM_TRAPS - 1)


def _integrand(u, s_i, s_j, s_k):
    return (bessel0_exp_scaled((s_i - s_j) * 0.5 * (1 - u)) * bessel0_exp_scaled((s_i + s_j) * 0.5 * (1 + u))
            * torch.exp((s_j + s_k) * (u - 1)))


def _forward_integrand(u, s):
    return _integrand(u, s[:, [1]], s[:, [2]], s[:, [0]])


def _backward_integrand(u, s):
    s_i = torch.max(s[:, 1:], dim=1, keepdim=True).values
    s_j = torch.min(s[:, 1:], dim=1, keepdim=True).values
    return _integrand(u, s_i, s_j, s[:, [0]]) * u


def log_mf_norm_constant(S):
    """S (n,3) proper singular values -> (log c(S) (n,), d log c / d S (n,3))."""
    S = S.float()
    c_bar = 0.5 * _trapezoid(_forward_integrand, S)
    log_c = torch.log(c_bar) + torch.sum(S, dim=1)
    d = torch.empty(S.shape[0], 3, dtype=S.dtype)
    for k in range(3):
        d[:, k] = 0.5 * _trapezoid(_backward_integrand, torch.cat((S[:, k:], S[:, :k]), dim=1))
    return log_c, d / c_bar.view(-1, 1)
