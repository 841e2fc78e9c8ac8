"""Simulation-based calibration of the sampler on the GPU (tests/sbclib.py): per functional of the catalog, the mean rank of
the truth among the posterior draws and the coverage of the central 50 % / 90 % intervals, with their binomial standard errors.
usage: python scripts/gpu_sbc.py [images] [particles] [stars] [sweeps] [basic]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sbclib import sbc

if __name__ == "__main__":
    basic = "basic" in sys.argv[1:]      # the Gaussian-PSF / Poisson model of BASELINE config 1
    a = [int(x) for x in sys.argv[1:] if x != "basic"]
    n = a[0] if a else 400
    res, iters = sbc(*a, basic=basic)
    print(f"{n} images, {iters} SMC iterations for the slowest; standard errors: mean rank {0.2887 / n ** 0.5:.3f}, "
          f"50 % coverage {0.5 / n ** 0.5:.3f}, 90 % coverage {0.3 / n ** 0.5:.3f}")
    for k, v in res.items():
        print(f"{k:26s} mean rank {v['mean_rank']:.3f}   50 %: {v['cover50']:.3f}   90 %: {v['cover90']:.3f}")
