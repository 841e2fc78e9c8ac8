"""Draw tape for running the upstream reference with injected random draws.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (in the build
container, where ``/root/reference`` exists) to replay pre-generated uniforms /
categorical indices through the reference's own call sites, so that the
reference, the C oracle and the CUDA path all consume identical draws.

The reference never passes ``generator=``; its RNG call sites are
(file:line in /root/reference):
  * ``torch.rand``                smcdet/distributions.py:44, :77 ; smcdet/sampler.py:143
  * ``torch.multinomial``         via torch.distributions.Categorical.sample, reached from
                                  ``Multinomial.sample`` at smcdet/kernel.py:44
  * ``Tensor.multinomial``        smcdet/sampler.py:129
  * ``Uniform(...).sample()``     smcdet/kernel.py:115, smcdet/prior.py:59  (-> ``torch.rand``)
  * ``torch.rand_like``           smcdet/kernel.py:261 (SingleComponentMALA)
"""

import contextlib

import torch


class DrawTape:
    """FIFO of tensors handed out to patched torch RNG entry points."""

    def __init__(self):
        self.rand_queue = []
        self.multinomial_queue = []
        self.log = []

    def push_rand(self, t):
        self.rand_queue.append(t)

    def push_multinomial(self, t):
        self.multinomial_queue.append(t)

    def _rand(self, *size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        t = self.rand_queue.pop(0)
        if tuple(t.shape) != tuple(size):
            raise RuntimeError(f"tape rand shape {tuple(t.shape)} != requested {tuple(size)}")
        self.log.append(("rand", tuple(size)))
        dtype = kw.get("dtype", None)
        return t.clone() if dtype is None else t.to(dtype)

    def _rand_like(self, other, **kw):
        t = self.rand_queue.pop(0)
        if tuple(t.shape) != tuple(other.shape):
            raise RuntimeError(f"tape rand_like shape {tuple(t.shape)} != requested {tuple(other.shape)}")
        self.log.append(("rand_like", tuple(other.shape)))
        return t.to(other.dtype)

    def _multinomial(self, probs, num_samples, replacement=False, **kw):
        t = self.multinomial_queue.pop(0)
        want = tuple(probs.shape[:-1]) + (num_samples,)
        if tuple(t.shape) != want:
            raise RuntimeError(f"tape multinomial shape {tuple(t.shape)} != requested {want}")
        self.log.append(("multinomial", want))
        return t.clone()

    @contextlib.contextmanager
    def active(self):
        orig_rand = torch.rand
        orig_rand_like = torch.rand_like
        orig_mn = torch.multinomial
        orig_tmn = torch.Tensor.multinomial
        tape = self

        def tensor_multinomial(self_t, num_samples, replacement=False, **kw):
            return tape._multinomial(self_t, num_samples, replacement, **kw)

        torch.rand = self._rand
        torch.rand_like = self._rand_like
        torch.multinomial = self._multinomial
        torch.Tensor.multinomial = tensor_multinomial
        try:
            yield self
        finally:
            torch.rand = orig_rand
            torch.rand_like = orig_rand_like
            torch.multinomial = orig_mn
            torch.Tensor.multinomial = orig_tmn
