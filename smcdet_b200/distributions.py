"""The reference's custom distributions (smcdet/distributions.py) as small torch classes.

On the hot path their arithmetic lives inside the CUDA kernels (the truncated-normal random walk
in ``smcdet_mh_mutate``, the truncated-Pareto flux prior in ``smcdet_prior_logprob`` /
``smcdet_prior_sample``); these classes keep the reference's public surface for code that builds
on it and are device-agnostic torch programs.
"""

import math

import torch
from torch.distributions import Distribution


class DiscreteUniform(Distribution):
    """Uniform pmf on the integers low..high (reference distributions.py:5-19)."""

    def __init__(self, low, high):
        self.low, self.high = low, high
        super().__init__(validate_args=False)

    def sample(self, sample_shape=torch.Size()):
        return torch.randint(self.low, self.high + 1, sample_shape)

    def log_prob(self, value):
        inside = (value >= self.low) & (value <= self.high)
        out = torch.full_like(value, float("-inf"), dtype=torch.get_default_dtype())
        return out.masked_fill(inside, -math.log(self.high - self.low + 1))


class TruncatedDiagonalMVN(Distribution):
    """Independent normals truncated to the box [lb, ub] (reference distributions.py:22-58)."""

    _SQRT2 = math.sqrt(2.0)

    def __init__(self, mu, sigma, lb, ub):
        super().__init__(validate_args=False)
        self.mu, self.sigma, self.lb, self.ub = mu, sigma, lb, ub
        self.dim = mu.size()
        self._cdf_lb = self._phi(lb)
        # log of the normal mass inside the box; nan -> 0, -inf -> most negative float
        self.log_prob_in_box = (self._phi(ub) - self._cdf_lb).log().nan_to_num()

    def _phi(self, x):
        return 0.5 * (1 + torch.erf((x - self.mu) / self.sigma / self._SQRT2))

    def sample(self, shape=None):
        shape = tuple(self.dim) if shape is None else shape
        eps = 1e-6
        p = torch.rand(shape, device=self.mu.device).clamp(eps, 1.0 - eps)
        q = (self._cdf_lb + p * self.log_prob_in_box.exp()).clamp(eps, 1.0 - eps)
        x = self.mu + self.sigma * self._SQRT2 * torch.erfinv(2 * q - 1)
        return torch.maximum(torch.minimum(x, torch.as_tensor(self.ub, device=x.device)),
                             torch.as_tensor(self.lb, device=x.device))

    def log_prob(self, value):
        assert (value >= self.lb).all() and (value <= self.ub).all()
        z = (value - self.mu) / self.sigma
        return -0.5 * z * z - torch.as_tensor(self.sigma).log() - 0.5 * math.log(2 * math.pi) - self.log_prob_in_box

    def cdf(self, value):
        num = (self._phi(value) - self._cdf_lb + 1e-9).log().sum(-1)
        return (num - self.log_prob_in_box).exp()


class TruncatedPareto(Distribution):
    """Bounded Pareto on [lower, upper] (reference distributions.py:61-89)."""

    def __init__(self, alpha, lower, upper):
        super().__init__(validate_args=False)
        self.alpha = torch.tensor(alpha)
        self.lower = torch.tensor(lower)
        self.upper = torch.tensor(upper)
        # log(alpha) + alpha log L + alpha log U - log(U^alpha - L^alpha)
        self.logpdf_norm_const = (self.alpha.log() + alpha * self.lower.log() + self.alpha * self.upper.log()
                                  - (self.upper**self.alpha - self.lower**self.alpha).log())

    def sample(self, shape=[]):
        u = torch.rand(shape)
        ua, la = self.upper**self.alpha, self.lower**self.alpha
        return ((ua - u * ua + u * la) / (la * ua)) ** (-1 / self.alpha)

    def log_prob(self, value):
        assert (value >= self.lower).all() and (value <= self.upper).all()
        return self.logpdf_norm_const - (self.alpha + 1) * value.log()
