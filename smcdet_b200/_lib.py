"""Loader and thin call helpers for ``libsmcdet_b200.so`` (C ABI in ``include/smcdet_b200.h``).

The library is the only compute path of this package: there is no CPU fallback.  Every tensor
handed to it must be a contiguous CUDA tensor; calls are asynchronous on torch's current
stream of the tensor's device.
"""

import ctypes as C
import os
import subprocess

import torch

from . import _abi as A

_PKG = os.path.dirname(os.path.abspath(__file__))
_SRC_DIR = os.path.join(_PKG, "csrc")
LIB_PATH = os.environ.get("SMCDET_B200_LIB") or os.path.join(_PKG, "libsmcdet_b200.so")
_SOURCES = [os.path.join(_SRC_DIR, "smcdet_kernels.cu"), os.path.join(_SRC_DIR, "smcdet_math.cuh"),
            os.path.join(os.path.dirname(_PKG), "include", "smcdet_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in _SOURCES)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [_SOURCES[0], "-o", LIB_PATH]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


# kernels launched by one call of each entry point (smcdet_mh_mutate: zero + mh + divide)
KERNELS_PER_CALL = {
    "smcdet_loglik": 1, "smcdet_loglik_segments": 1, "smcdet_psf": 1, "smcdet_psf_radial": 1, "smcdet_render": 1, "smcdet_prior_logprob": 1, "smcdet_prior_sample": 1,
    "smcdet_temper_update": 1, "smcdet_resample": 1, "smcdet_gather": 1, "smcdet_mh_mutate": 3, "smcdet_mh_mutate_resampled": 3, "smcdet_mala_mutate": 3, "smcdet_prune": 1,
    "smcdet_match_catalogs": 1, "smcdet_agg_join": 1, "smcdet_agg_unjoin": 1, "smcdet_agg_mutate": 3,
}


class _CountingLib(object):
    """Proxy over the bound CDLL that counts kernel launches per entry point (bench.py reports them)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self.launches = 0
        self.calls = {}
        for name, n in KERNELS_PER_CALL.items():
            setattr(self, name, self._wrap(name, getattr(cdll, name), n))

    def _wrap(self, name, fn, n):
        def call(*args):
            self.launches += n
            self.calls[name] = self.calls.get(name, 0) + 1
            return fn(*args)
        return call

    def adjust(self, n):
        """Correct the launch count of the last call (e.g. smcdet_mh_mutate with acc_as_count launches 1 kernel, not 3)."""
        self.launches += n

    def __getattr__(self, name):
        return getattr(self._cdll, name)


_lib = None


def lib():
    """The bound shared library.  Fails loudly if it has not been built or exports are missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(smcdet_b200 has no CPU or eager-PyTorch fallback)")
        cdll = C.CDLL(LIB_PATH)
        A.bind(cdll)
        if cdll.smcdet_version() != A.ABI_VERSION:
            raise RuntimeError("libsmcdet_b200.so ABI version mismatch; rebuild it")
        _lib = _CountingLib(cdll)
    return _lib


class SmcdetError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise SmcdetError(f"libsmcdet_b200 call failed (code {rc}): {lib().smcdet_last_error_string().decode()}")


def device():
    """Device new tensors are created on: torch's default device if it is a CUDA device (the
    reference's convention, experiments/basic/run_smc.py:22-24), else the current CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("smcdet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    d = torch.get_default_device()
    if d.type == "cuda":
        return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())
    return torch.device("cuda", torch.cuda.current_device())


def f32(t, dev=None):
    """A contiguous float32 CUDA tensor with the values of ``t``."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if dev is None:
        dev = t.device if t.is_cuda else device()
    return t.to(device=dev, dtype=torch.float32).contiguous()


def ptr(t, dtype=torch.float32):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("smcdet_b200 kernels take CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_for(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def fresh_seed():
    """A 62-bit Philox seed drawn from torch's global CPU generator, so torch.manual_seed()
    makes a whole run reproducible."""
    return int(torch.randint(0, 2**62, (1,), device="cpu").item())
