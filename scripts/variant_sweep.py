"""Time mh_kernel (148 tiles x 10 000 particles x D = 10, 100 sweeps) for every library in build/variants/
(GPU side of scripts/build_variants.py).  usage: python scripts/variant_sweep.py [names...]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
want = sys.argv[1:]
for lib in libs:
    name = os.path.basename(lib)[:-3]
    if want and name not in want:
        continue
    env = dict(os.environ, SMCDET_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_time_mh.py")], capture_output=True, text=True, env=env, cwd=ROOT)
    print(f"{name:24s}", (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
