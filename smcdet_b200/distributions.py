"""Box-truncated normal, bounded Pareto and integer-uniform laws with the reference's class names.

On the hot path this arithmetic lives inside the CUDA kernels: the truncated-normal random walk and its
proposal-density ratio in ``smcdet_mh_mutate`` / ``smcdet_mala_mutate`` (``smcdet_math.cuh``:
``truncnormal_make*``, ``truncnormal_draw``), the bounded-Pareto flux prior in ``smcdet_prior_logprob`` /
``smcdet_prior_sample``.  The classes below keep the public surface of ``smcdet/distributions.py`` for code
that builds on it; they are small device-agnostic torch programs written around three helpers.
"""

import math

import torch
from torch.distributions import Distribution

_INV_SQRT2 = 1.0 / math.sqrt(2.0)
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
_EDGE = 1e-6  # uniforms and cdf values are kept this far from 0 and 1 before inversion


def _gauss_cdf(x, centre, width):
    """Phi((x - centre) / width) written with erf, as torch.distributions.Normal.cdf does."""
    return 0.5 * (1.0 + torch.erf((x - centre) / width * _INV_SQRT2))


def _gauss_logpdf(x, centre, width):
    z = (x - centre) / width
    return -0.5 * z * z - torch.as_tensor(width).log() - _HALF_LOG_2PI


def _pin(t, lo, hi):
    """Clip to [lo, hi] where the bounds may be python numbers or tensors."""
    like = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
    return torch.minimum(torch.maximum(like, torch.as_tensor(lo, device=like.device)), torch.as_tensor(hi, device=like.device))


class DiscreteUniform(Distribution):
    """P(k) = 1 / (high - low + 1) on the integers low..high; -inf log-mass elsewhere."""

    def __init__(self, low, high):
        self.low, self.high = low, high
        self._log_mass = -math.log(high - low + 1)
        super().__init__(validate_args=False)

    def sample(self, sample_shape=torch.Size()):
        return torch.randint(self.low, self.high + 1, sample_shape)

    def log_prob(self, value):
        inside = (value >= self.low) & (value <= self.high)
        return torch.where(inside, self._log_mass, float("-inf"))


class TruncatedDiagonalMVN(Distribution):
    """Independent N(mu, sigma^2) coordinates restricted to the box [lb, ub].

    ``log_prob_in_box`` is the log of the Gaussian mass inside the box per coordinate, with nan mapped to 0 and
    -inf to the most negative float (torch.nan_to_num defaults), which the reference relies on when a proposal
    mean sits far outside the box."""

    def __init__(self, mu, sigma, lb, ub):
        super().__init__(validate_args=False)
        self.mu, self.sigma, self.lb, self.ub = mu, sigma, lb, ub
        self.dim = mu.size()
        self._below = _gauss_cdf(lb, mu, sigma)                       # mass to the left of the box
        inside = _gauss_cdf(ub, mu, sigma) - self._below
        self.log_prob_in_box = torch.nan_to_num(torch.log(inside))

    def sample(self, shape=None):
        """Inverse-cdf draw: u -> Phi^-1(below + u * mass), clipped back into the box."""
        want = tuple(self.dim) if shape is None else shape
        u = torch.rand(want, device=self.mu.device).clamp_(_EDGE, 1.0 - _EDGE)
        level = (self._below + u * torch.exp(self.log_prob_in_box)).clamp_(_EDGE, 1.0 - _EDGE)
        draw = self.mu + self.sigma * math.sqrt(2.0) * torch.erfinv(2.0 * level - 1.0)
        return _pin(draw, self.lb, self.ub)

    def log_prob(self, value):
        assert (value >= self.lb).all() and (value <= self.ub).all()
        return _gauss_logpdf(value, self.mu, self.sigma) - self.log_prob_in_box

    def cdf(self, value):
        partial = torch.log(_gauss_cdf(value, self.mu, self.sigma) - self._below + 1e-9).sum(-1)
        return torch.exp(partial - self.log_prob_in_box)


class TruncatedPareto(Distribution):
    """Bounded Pareto: density proportional to f^-(alpha+1) on [lower, upper]."""

    def __init__(self, alpha, lower, upper):
        super().__init__(validate_args=False)
        self.alpha, self.lower, self.upper = torch.tensor(alpha), torch.tensor(lower), torch.tensor(upper)
        a, lo, up = self.alpha, self.lower, self.upper
        # log(alpha L^alpha U^alpha / (U^alpha - L^alpha))
        self.logpdf_norm_const = a.log() + alpha * lo.log() + a * up.log() - (up**a - lo**a).log()

    def sample(self, shape=[]):
        u = torch.rand(shape)
        top, bottom = self.upper**self.alpha, self.lower**self.alpha
        return ((top - u * top + u * bottom) / (bottom * top)) ** (-1 / self.alpha)

    def log_prob(self, value):
        assert (value >= self.lower).all() and (value <= self.upper).all()
        return self.logpdf_norm_const - (self.alpha + 1) * value.log()
