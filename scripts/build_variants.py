"""Build variants of the CUDA library with different -D switches into build/variants/ (development aid for
kernel tuning: one gpurun call then times them all with scripts/variant_sweep.py).
usage: python scripts/build_variants.py name1="-DSMC_X=1 -DSMC_Y=2" name2="..." """
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smcdet_b200 import _lib

out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)


def one(spec):
    name, flags = spec.split("=", 1)
    out = os.path.join(out_dir, name + ".so")
    cmd = ["/usr/local/cuda/bin/nvcc"] + _lib.NVCC_FLAGS + flags.split() + [_lib._SOURCES[0], "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, (r.stdout + r.stderr)[-2000:]


with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "6"))) as ex:
    for name, rc, log in ex.map(one, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED\n" + log)
