#!/usr/bin/env python
"""Benchmark of the smcdet per-tile SMC hot path on B200 (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]          own arm (CUDA library)
  python bench.py --impl reference [...]                        reference arm (the reference's path on the host CPU)
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU

Workloads (--workload):
  m71synthetic      BASELINE.json configs[1] (default): synthetic SDSS-r-band-like field drawn from the M71 model, cut
                    into 8x8 tiles; per tile a likelihood-tempered SMC sampler with N = 10 000 catalogs of D = 10 stars,
                    100 single-site MH sweeps per SMC iteration, ESS threshold 0.5 N, multinomial resampling, run to
                    temperature 1 (notebooks/smc.ipynb cells 3-7 of the reference)
  m71semisynthetic  configs[2]: three times the source density, D = 16
  basic             configs[0]: Gaussian PSF + Poisson likelihood, D = 8, pad 2 (experiments/basic/run_smc.py:44-105)
  allstrata         north_star's target sentence: every tile x every count stratum 0..D as count-stratified SMC
                    (manuscript.tex:312-356), the (tile, count) strata spread over the GPUs by expected cost

Scaling (--scaling):
  weak    every rank owns --tiles-per-gpu tiles of its own field (tiles are independent: no data-path collective)
  strong  ONE field of --field-tiles tiles (BASELINE.json configs[3]) sharded round-robin over the ranks
  both    (default) the line's `value` is the weak-scaling number; the object `strong` carries the fixed-field number
          and a checksum of the field's per-tile summaries, which must not depend on the number of GPUs
End-to-end region: pinned host tiles -> H2D -> samplers -> the reference's `Aggregate` finish (aggregate.py:583-589)
-> pruned catalogs and summaries D2H.  For ONE field (strong scaling, and any 1-GPU run) the weighted catalogs of all
tiles are first gathered with NCCL onto rank 0, which runs the finish; in weak scaling with several GPUs every rank owns
its own field and finishes it itself (no catalog collective).  The last warm-up step goes through this path as well.

A "step" is one complete run to temperature 1 over the rank's tiles.  `value` = particle-likelihood evaluations per
second over the whole job, counted as the reference evaluates them: per SMC iteration and live tile N*(num_iters + 2)
(kernel.py:64-70, :89-96; sampler.py:100-102), plus N for initialize.
"""

import argparse
import contextlib
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# canonical parameters of the reference (notebooks/smc.ipynb raw lines 53-63)
M71 = dict(background=104.1486587524414, adu_per_nmgy=241.02658081054688,
           psf_params=[1.107237458229065, 2.0800251960754395, 2.3254318237304688, 5.240590572357178,
                       0.7346734404563904, 0.5114791393280029],
           psf_radius=8, noise_additive=1.0000007072408224e-10, noise_multiplicative=1.936462640762329)
PRIOR = dict(counts_rate=0.030264640226960182, flux_alpha=0.21411753249015655, flux_lower=0.06291294097900389,
             flux_upper=1804.6791992187502)
DETECTION = 0.25165176391601557
TILE = 8
# experiments/basic/run_smc.py:44-105 (Gaussian PSF, Poisson likelihood)
BASIC_STDEV, BASIC_BG = 0.93, 200.0
_PSF_MAX = 1.0 / (2 * np.pi * BASIC_STDEV**2)
BASIC_SCALE = float(5 * np.sqrt(BASIC_BG) / _PSF_MAX)          # 384.265: the detection threshold
BASIC_ALPHA = float(-np.log(1 - 0.99) / (np.log(50 * np.sqrt(BASIC_BG) / _PSF_MAX) - np.log(BASIC_SCALE)))
METRIC, UNIT = "particle-likelihood evals/sec", "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="m71synthetic", choices=["m71synthetic", "m71semisynthetic", "basic", "allstrata"])
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"])
    ap.add_argument("--tiles-per-gpu", type=int, default=None, help="weak scaling: tiles per rank (default 800; 96 for allstrata)")
    ap.add_argument("--field-tiles", type=int, default=None, help="strong scaling: tiles of the one field (default 800; 96 for allstrata)")
    ap.add_argument("--particles", type=int, default=10000)
    ap.add_argument("--stars", type=int, default=None, help="stars per catalog (10; 16 m71semisynthetic; 8 basic)")
    ap.add_argument("--mh-iters", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.stars is None:
        a.stars = {"m71semisynthetic": 16, "basic": 8}.get(a.workload, 10)
    small = a.workload == "allstrata"
    if a.tiles_per_gpu is None:
        a.tiles_per_gpu = 96 if small else 800
    if a.field_tiles is None:
        a.field_tiles = 96 if small else 800
    return a


def is_m71(a):
    return a.workload != "basic"


def pad_of(a):
    return 2 if a.workload == "basic" else 4


def workload_config(a):
    names = {
        "m71synthetic": "m71synthetic: M71 PSF + Normal likelihood, 8x8 tiles, psf_radius 8, pad 4 "
                        "(BASELINE.json configs[1]; notebooks/smc.ipynb of the reference)",
        "m71semisynthetic": "m71semisynthetic-shaped: as m71synthetic with three times the source density and larger catalogs "
                            "(BASELINE.json configs[2]; SURVEY.md 8d config 3)",
        "basic": "basic: Gaussian PSF (stdev 0.93) + Poisson likelihood, 8x8 tiles, psf_radius 8, pad 2, background 200 "
                 "(BASELINE.json configs[0]; experiments/basic/run_smc.py:44-105 of the reference)",
        "allstrata": "allstrata: count-stratified SMC, every tile x every count 0..D as its own tempered sampler "
                     "(north_star target; manuscript.tex:312-356), M71 model as m71synthetic",
    }
    sharding = ("(tile, count) strata assigned to ranks by expected cost (longest-processing-time), per-stratum evidences "
                "all-gathered (NCCL)") if a.workload == "allstrata" else (
        "tiles sharded round-robin, no data-path collective.  End to end, strong scaling (one field): NCCL gather "
        "(dist.gather) of every tile's weighted catalogs (counts, locs, fluxes, weights, log Z) onto rank 0 and the "
        "Aggregate finish there (aggregate.py:583-589); weak scaling with more than one GPU (every rank its own field): "
        "each rank runs that finish on its own tiles and reads its own catalogs back, no catalog collective")
    return {"workload": names[a.workload],
            "tiles_per_gpu": a.tiles_per_gpu, "field_tiles_strong": a.field_tiles, "particles_per_tile": a.particles,
            "stars_per_catalog": a.stars, "mh_iters": a.mh_iters, "ess_threshold_prop": 0.5, "resample": "multinomial",
            "loglik_for_tempering": "from the incrementally updated rate image (default; 7e-7 relative drift measured)",
            "step": "one full run to temperature 1 over all tiles of the rank",
            "finished_tiles": "frozen (each tile runs as in the reference's per-tile loop, experiments/m71/run_smc.py:113-124)",
            "parallelism": f"{a.gpus} GPU(s): {sharding}",
            "l2": "particle state per rank (>= 1 GB at the default size) exceeds the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (C port of the reference's arithmetic), all host threads; and the unmodified reference
# ----------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_smc_iteration(O, om, op, mh, tiles, state, rng, N, iters):
    """One SMC iteration of the reference (sampler.py:244-247) on the oracle: resample, MH, temper,
    update_weights.  Returns the number of particle-likelihood evaluations."""
    T = tiles.shape[0]
    u = rng.random((T, N))
    idx = O.resample(O.RESAMPLE_MULTINOMIAL, state["weights"], u)
    state["counts"], state["locs"], state["fluxes"] = O.gather(idx, state["counts"], state["locs"], state["fluxes"])
    D = state["fluxes"].shape[-1]
    comp = rng.integers(0, D, (iters, T, N), dtype=np.int32)
    r = O.mh_run(om, op, mh, tiles, state["counts"], state["locs"], state["fluxes"], state["tau"], comp,
                 rng.random((iters, T, N, 2), dtype=np.float32), rng.random((iters, T, N), dtype=np.float32),
                 rng.random((iters, T, N), dtype=np.float32), traces=False)
    state["locs"], state["fluxes"] = r["locs"], r["fluxes"]
    ll = O.loglik(om, tiles, state["locs"], state["fluxes"])  # the recompute of sampler.py:100-102
    tau_new, _, _ = O.temper(ll, state["tau"], 0.5 * N)
    _, state["weights"], _, state["logz"] = O.update_weights(ll, tau_new, state["tau"], state["logz"])
    state["tau"] = tau_new
    return T * N * (iters + 2)


def cpu_sample_setup(a, n_tiles, seed=0):
    from oracle import api as O

    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm always takes every core it may run on
    O.set_num_threads(host_cores())
    pad = pad_of(a)
    rng = np.random.default_rng(seed)
    N, D = a.particles, a.stars
    if is_m71(a):
        om = O.m71_model(M71["psf_radius"], M71["psf_params"], M71["background"], M71["adu_per_nmgy"],
                         M71["noise_additive"], M71["noise_multiplicative"])
        op = O.m71_prior(D, D, PRIOR["counts_rate"], TILE, TILE, PRIOR["flux_alpha"], PRIOR["flux_lower"], PRIOR["flux_upper"], pad=pad)
        mh = O.make_mh(a.mh_iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"], (-pad, -pad), (TILE + pad, TILE + pad))
        tp = O.m71_prior(4, 4, PRIOR["counts_rate"], TILE, TILE, PRIOR["flux_alpha"], DETECTION, PRIOR["flux_upper"], pad=pad)
    else:
        om = O.gauss_model(8, BASIC_STDEV, BASIC_BG)
        op = O.pareto_prior(D, D, TILE, TILE, 0.9 * BASIC_SCALE, BASIC_ALPHA, pad=pad)
        mh = O.make_mh(a.mh_iters, 0.1, 100.0, 0.9 * BASIC_SCALE, 1e6, (-pad, -pad), (TILE + pad, TILE + pad))
        tp = O.pareto_prior(4, 4, TILE, TILE, BASIC_SCALE, BASIC_ALPHA, pad=pad)
    # observed tiles: drawn from the model with a handful of stars
    _, tl, tf = O.prior_sample(tp, rng.random((n_tiles, 1, 4, 2), dtype=np.float32), rng.random((n_tiles, 1, 4), dtype=np.float32), 1)
    rate = O.render(om, tl, tf, TILE, TILE)[..., 0]
    if is_m71(a):
        tiles = (rate + rng.standard_normal(rate.shape) * np.sqrt(M71["noise_additive"] + M71["noise_multiplicative"] * rate)).astype(np.float32)
    else:
        tiles = rng.poisson(rate).astype(np.float32)
    counts, locs, fluxes = O.prior_sample(op, rng.random((n_tiles, N, D, 2), dtype=np.float32),
                                          rng.random((n_tiles, N, D), dtype=np.float32), N)
    state = dict(counts=counts, locs=locs, fluxes=fluxes, weights=np.full((n_tiles, N), 1.0 / N, np.float32),
                 tau=np.full(n_tiles, 0.05, np.float32), logz=np.zeros(n_tiles, np.float32))
    return O, om, op, mh, tiles, state, rng


def cpu_calibrate(a):
    """evals/s of the CPU arm on a small probe, to size the bounded sample."""
    import copy

    small = copy.copy(a)
    small.particles, small.mh_iters = 1000, 10
    O, om, op, mh, tiles, state, rng = cpu_sample_setup(small, 1)
    t0 = time.perf_counter()
    n = cpu_smc_iteration(O, om, op, mh, tiles, state, rng, small.particles, small.mh_iters)
    return n / (time.perf_counter() - t0), O.num_threads()


def cpu_measure(a, target_seconds, steps=1, warmup=0):
    rate, threads = cpu_calibrate(a)
    per_tile = a.particles * (a.mh_iters + 2)
    n_tiles = int(max(1, min(64, round(rate * target_seconds / per_tile))))
    O, om, op, mh, tiles, state, rng = cpu_sample_setup(a, n_tiles)
    times, evals = [], 0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        n = cpu_smc_iteration(O, om, op, mh, tiles, state, rng, a.particles, a.mh_iters)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
            evals += n
    total = sum(times)
    sample = (f"one SMC iteration (multinomial resample, {a.mh_iters} MH sweeps, likelihood recompute, Brent tempering, "
              f"weight update) on {n_tiles} tile(s) x {a.particles} particles x {a.stars} stars, {steps} repetition(s)")
    return dict(value=evals / total, unit=UNIT, cores=threads, kind="port", sample=sample), total / max(1, steps) * 1e3, n_tiles


def reference_measure(a, target_seconds=20.0):
    """The UNMODIFIED reference (oracle/_ref/smcdet, copied from /root/reference by oracle/make_ref.sh at build time;
    pure Python / PyTorch) timed on the host cores: SMC iterations of its own SMCsampler (resample, mutate, temper,
    update_weights: sampler.py:244-247) on one tile at reduced N and MH sweeps.  None if the copy is absent."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "smcdet", "sampler.py")):
        return None
    import torch

    sys.path.insert(0, ref_dir)
    try:
        from smcdet.images import ImageModel, M71ImageModel
        from smcdet.kernel import SingleComponentMH
        from smcdet.prior import M71Prior, ParetoStarPrior
        from smcdet.sampler import SMCsampler
    finally:
        sys.path.remove(ref_dir)
    cores = host_cores()
    torch.set_num_threads(cores)
    N, iters, D, pad = 1000, 10, a.stars, pad_of(a)
    torch.manual_seed(0)
    if is_m71(a):
        im = M71ImageModel(image_height=TILE, image_width=TILE, **M71)
        pr = M71Prior(min_objects=D, max_objects=D, image_height=TILE, image_width=TILE, pad=pad, **PRIOR)
        tp = M71Prior(min_objects=4, max_objects=4, image_height=TILE, image_width=TILE, pad=pad,
                      **dict(PRIOR, flux_lower=DETECTION))
        mh = SingleComponentMH(iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        thr = DETECTION
    else:
        im = ImageModel(image_height=TILE, image_width=TILE, psf_radius=8, psf_stdev=BASIC_STDEV, background=BASIC_BG)
        pr = ParetoStarPrior(min_objects=D, max_objects=D, image_height=TILE, image_width=TILE, flux_scale=0.9 * BASIC_SCALE,
                             flux_alpha=BASIC_ALPHA, pad=pad)
        tp = ParetoStarPrior(min_objects=4, max_objects=4, image_height=TILE, image_width=TILE, flux_scale=BASIC_SCALE,
                             flux_alpha=BASIC_ALPHA, pad=pad)
        mh = SingleComponentMH(iters, 0.1, 100.0, 0.9 * BASIC_SCALE, 1e6)
        thr = BASIC_SCALE
    _, tl, tf = tp.sample(num_tiles_per_side=1, stratify_by_count=True, num_catalogs_per_count=1)
    image = im.sample(tl, tf)[0, 0, :, :, 0].contiguous()
    with contextlib.redirect_stdout(sys.stderr):
        s = SMCsampler(image, TILE, pr, im, mh, N, 0.5, "multinomial", thr, 100, print_every=10**6)
        s.initialize()
        s.temper()
        s.update_weights()
        done, evals, t_total = 0, 0, 0.0
        while t_total < target_seconds and done < 12:
            t0 = time.perf_counter()
            s.resample(); s.mutate(); s.temper(); s.update_weights()
            dt = time.perf_counter() - t0
            if done > 0:  # the first iteration warms torch's thread pool and allocator
                t_total += dt
                evals += N * (iters + 2)
            done += 1
    if evals == 0:
        return None
    return dict(value=evals / t_total, unit=UNIT, cores=cores, kind="reference",
                sample=f"the unmodified reference's SMCsampler (oracle/_ref/smcdet, torch {torch.__version__} on the CPU, "
                       f"{cores} threads): {done - 1} SMC iterations (resample, mutate, temper, update_weights) on 1 tile x "
                       f"{N} particles x {D} stars, {iters} MH sweeps")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms, n_tiles = cpu_measure(a, target_seconds=6.0, steps=a.steps, warmup=a.warmup)
    real = None
    try:
        real = reference_measure(a, target_seconds=10.0)
    except Exception as exc:  # noqa: BLE001  (the copy of the reference is optional test infrastructure)
        real = {"kind": "reference", "unavailable": repr(exc)[:200]}
    base_all = dict(base, also=[real] if real else [])
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": base_all,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tiles_per_sec_full_smc_est": base["value"] / (a.particles * (a.mh_iters + 2) * 12.0),
            "note": "value = CPU port (oracle/smcdet_oracle.c, OpenMP, all host cores whatever OMP_NUM_THREADS says) of the "
                    "reference's arithmetic -- the faster of the two CPU baselines, so ratios against it are conservative; "
                    "cpu_baseline.also = the unmodified torch reference timed on the same cores at reduced N"}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons of one GPU every 200 ms during the timed region, through NVML in
    this process (a polling `nvidia-smi` child was measured to slow the timed region by several percent)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_flag = index, [], threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            visible = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
            entry = visible[self.index] if self.index < len(visible) else str(self.index)
            if entry.isdigit():
                h = nv.nvmlDeviceGetHandleByIndex(int(entry))
            else:  # CUDA_VISIBLE_DEVICES given as UUIDs
                h = nv.nvmlDeviceGetHandleByUUID(entry.encode() if isinstance(entry, str) else entry)
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop_flag.is_set():
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  [k for k, b in bits.items() if reasons & b]))
                self._stop_flag.wait(0.2)
        except Exception as exc:  # noqa: BLE001
            self.error = repr(exc)
            self._poll_nvidia_smi()

    def _poll_nvidia_smi(self):
        """Fallback when NVML cannot be used in-process: a polling nvidia-smi child (the recipe's clocks line)."""
        import subprocess

        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                     "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            for line in proc.stdout:
                c = [x.strip() for x in line.split(",")]
                try:
                    self.rows.append((float(c[0]), float(c[1]), float(c[2]),
                                      [n for n, v in zip(names, c[3:7]) if v.lower().startswith("active")]))
                except (ValueError, IndexError):
                    pass
                if self._stop_flag.is_set():
                    proc.terminate()
                    break
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        self._stop_flag.set()
        self.join(timeout=2)
        sm = [r[0] for r in self.rows]
        power = [r[2] for r in self.rows]
        reasons = sorted({x for r in self.rows for x in r[3]})
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons,
                "source": "NVML (pynvml) every 200 ms during the timed regions", **({"error": self.error} if hasattr(self, "error") else {})}


def make_objects(a):
    """(image model, inference prior, true prior, detection threshold, MH kernel factory) of the workload."""
    from smcdet_b200.images import ImageModel, M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior, ParetoStarPrior

    D, pad = a.stars, pad_of(a)
    if is_m71(a):
        model = M71ImageModel(TILE, TILE, **M71)
        lo = 0 if a.workload == "allstrata" else D
        prior = M71Prior(lo, D, PRIOR["counts_rate"], TILE, TILE, flux_alpha=PRIOR["flux_alpha"],
                         flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=pad)
        true_prior = M71Prior(0, 24, PRIOR["counts_rate"], TILE, TILE, flux_alpha=PRIOR["flux_alpha"],
                              flux_lower=DETECTION, flux_upper=PRIOR["flux_upper"], pad=pad)
        return model, prior, true_prior, DETECTION, lambda: SingleComponentMH(a.mh_iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
    model = ImageModel(TILE, TILE, psf_radius=8, psf_stdev=BASIC_STDEV, background=BASIC_BG)
    prior = ParetoStarPrior(D, D, TILE, TILE, flux_scale=0.9 * BASIC_SCALE, flux_alpha=BASIC_ALPHA, pad=pad)
    true_prior = ParetoStarPrior(0, 8, TILE, TILE, flux_scale=BASIC_SCALE, flux_alpha=BASIC_ALPHA, pad=pad)
    return model, prior, true_prior, BASIC_SCALE, lambda: SingleComponentMH(a.mh_iters, 0.1, 100.0, 0.9 * BASIC_SCALE, 1e6)


def make_field(a, num_tiles, seed, dev):
    """Synthetic field of `num_tiles` tiles drawn from the workload's image model with its true prior
    (notebooks/smc.ipynb cell 3; experiments/basic/generate_images.py:26-76): [T, 8, 8] on `dev`."""
    import torch

    model, _, _, _, _ = make_objects(a)
    pad = pad_of(a)
    T, D = num_tiles, 24
    g = torch.Generator(device=dev).manual_seed(seed)
    low, high = -pad, TILE + pad
    if is_m71(a):
        density = 3.0 if a.workload == "m71semisynthetic" else 1.0
        mean = density * PRIOR["counts_rate"] * (TILE + 2 * pad) ** 2
        counts = torch.poisson(torch.full((T,), mean, device=dev), generator=g).clamp(max=D)
        al = PRIOR["flux_alpha"]
        ua, la = PRIOR["flux_upper"] ** al, DETECTION ** al
        u = torch.rand(T, 1, 1, D, device=dev, generator=g)
        fluxes = ((ua - u * ua + u * la) / (la * ua)) ** (-1.0 / al)
    else:
        counts = torch.randint(0, 9, (T,), device=dev, generator=g).float()   # DiscreteUniform{0..8} (prior.py:157-162)
        u = torch.rand(T, 1, 1, D, device=dev, generator=g)
        fluxes = BASIC_SCALE * (1.0 - u) ** (-1.0 / BASIC_ALPHA)               # Pareto(scale, alpha)
    locs = low + torch.rand(T, 1, 1, D, 2, device=dev, generator=g) * (high - low)
    mask = torch.arange(D, device=dev).view(1, 1, 1, D) < counts.view(T, 1, 1, 1)
    rate = model._rate(locs * mask.unsqueeze(-1), fluxes * mask)  # [T,1,8,8,1]
    if is_m71(a):
        noise = torch.randn(rate.shape, device=dev, generator=g)
        img = rate + noise * (model.noise_additive + model.noise_multiplicative * rate).sqrt()
    else:
        img = torch.poisson(rate, generator=g)
    return img[:, 0, :, :, 0].contiguous()  # [T,8,8]


def run_own(a):
    import torch
    import torch.distributed as dist

    from smcdet_b200 import _lib as L
    from smcdet_b200.cssmc import CountStratifiedSMC
    from smcdet_b200.shard import ShardedSMC, gather_tiles, shard_tile_ids

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION/WARN/INFO);
        # stdout must carry the JSON line only, so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"

    N, D, iters = a.particles, a.stars, a.mh_iters
    model, prior, _, threshold, new_mh = make_objects(a)
    strata = a.workload == "allstrata"
    ns = D + 1 if strata else 1
    lib = L.lib()
    quiet = contextlib.redirect_stdout(sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ------------------------------------------------------------------------------------------
    # one job = one field of `num_tiles` tiles sharded over the ranks
    # ------------------------------------------------------------------------------------------
    class Job(object):
        def __init__(self, num_tiles, field_seed, gather):
            # gather: the end-to-end region collects every tile's catalogs on rank 0 (ONE field sharded over the ranks);
            # otherwise every rank finishes and reads back its own tiles (ranks own separate parts of the job)
            self.gather = gather or world == 1
            self.T = num_tiles
            self.field = make_field(a, num_tiles, field_seed, dev)          # the whole field, identical on every rank
            self.field_host = self.field.cpu().pin_memory()
            self.local_ids = shard_tile_ids(num_tiles, world, rank)
            self.T_local = len(self.local_ids)

        def run(self, seed, tiles=None, e2e=False):
            """One step.  Returns (evals, live segment-iterations, MH event log, outputs for the host, summaries)."""
            torch.manual_seed(seed)
            mh = new_mh()
            mh.event_log = []
            field = self.field if tiles is None else tiles
            if strata:
                with quiet:
                    cs = CountStratifiedSMC(field.view(self.T, 1, TILE, TILE), TILE, prior, model, mh, N, 0.5, "multinomial",
                                            threshold, 200, verbose=False, keep_samplers=True, rank=rank, world=world,
                                            seed=seed)
                    cs.run()
                smp = cs.samplers["all"]
                mhk = smp.MutationKernel
                live = sum(cs.live_strata)
                S = cs.local_strata.numel()
                evals = live * N * (iters + 2) + S * N
                summ = torch.stack([cs.log_evidence.view(-1), cs.posterior_mean_count().view(-1),
                                    cs.posterior_count_probs.view(self.T, ns).max(-1).values,
                                    cs.log_normalizing_constant.view(self.T, ns)[:, -1]], -1)   # [T, 4], identical on all ranks
                outs = [summ, cs.posterior_count_probs.view(self.T, ns)] if rank == 0 else []
                return evals, live, mhk.event_log, outs, summ, int(smp.iter)
            sh = ShardedSMC(field, TILE, prior, model, mh, N, 0.5, "multinomial", threshold, 200, seed=seed, device=dev)
            sh.run()
            s = sh.sampler
            live = sum(s.live_tiles)
            evals = live * N * (iters + 2) + self.T_local * N
            if not e2e:
                summ = sh.local_results()["summaries"]
                return evals, live, mh.event_log, [], summ, int(s.iter)
            # the reference's finish: all tiles' weighted catalogs -> Aggregate on rank 0 (NCCL gather for world > 1),
            # or -- weak scaling, where ranks own separate fields -- every rank's own tiles -> its own Aggregate
            with quiet:
                agg = sh.sink(local=not self.gather)
            outs = []
            if agg is not None:
                Tn = self.T if self.gather else self.T_local
                outs = [agg.summaries, agg.pruned_counts.view(Tn, N).to(torch.int16), agg.pruned_locs.view(Tn, N, D, 2),
                        agg.pruned_fluxes.view(Tn, N, D)]
            return evals, live, mh.event_log, outs, None, int(s.iter)

        def summaries_global(self, summ):
            """[T, k] per-tile summaries in global tile order on every rank (strata: already global)."""
            if strata or world == 1:
                return summ
            return gather_tiles(summ.contiguous(), self.T)

    def timed(job, steps, seed0, e2e):
        """`steps` steps of `job` between two events; e2e: tiles start in pinned host memory and the step's results
        land in pinned host buffers (the read-back of step k overlaps step k + 1 on a copy stream; everything is complete
        before the closing event).  Returns a dict of measurements (times are the max over ranks)."""
        copy_stream = torch.cuda.Stream(device=dev)
        host_sets, pending = {}, []
        if e2e and (rank == 0 or not job.gather):  # pinned landing buffers, allocated before the clock starts (two sets)
            Tn = job.T if (job.gather or strata) else job.T_local
            shapes = ([((Tn, 4), torch.float32), ((Tn, ns), torch.float32)] if strata else
                      [((Tn, 6), torch.float32), ((Tn, N), torch.int16), ((Tn, N, D, 2), torch.float32), ((Tn, N, D), torch.float32)])
            for slot in range(min(2, steps)):
                host_sets[slot] = [torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.launches
        evals = live_total = h2d = d2h = 0
        logs, smc_iters, summ = [], [], None
        barrier()
        e0.record(torch.cuda.current_stream(dev))
        for k in range(steps):
            tiles = None
            if e2e:
                tiles = job.field_host.to(dev, non_blocking=True)
                h2d += job.field_host.numel() * 4
            ev, live, log, outs, summ_k, it = job.run(seed0 + k, tiles, e2e)
            evals += ev
            live_total += live
            logs.append(log)
            smc_iters.append(it)
            summ = summ if summ is not None else summ_k   # the FIRST timed step's summaries (same seed whatever --steps)
            if e2e and outs:
                slot = k % 2
                if slot not in host_sets:
                    host_sets[slot] = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ready)
                    for hbuf, o in zip(host_sets[slot], outs):
                        hbuf.copy_(o, non_blocking=True)
                        o.record_stream(copy_stream)
                        d2h += o.numel() * o.element_size()
                pending = [outs]
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        e1.record(torch.cuda.current_stream(dev))
        barrier()
        ms = reduce_ranks(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
        del pending
        return dict(ms=ms, evals=reduce_ranks(float(evals), dist.ReduceOp.SUM if world > 1 else None), live=live_total,
                    logs=logs, smc_iters=smc_iters, launches=lib.launches - launches0, h2d=h2d // steps,
                    d2h=int(reduce_ranks(float(d2h), dist.ReduceOp.SUM if world > 1 else None)) // steps, summ=summ)

    do_weak = a.scaling in ("weak", "both")
    do_strong = a.scaling in ("strong", "both")
    same = do_weak and do_strong and a.tiles_per_gpu * world == a.field_tiles
    weak_job = Job(a.tiles_per_gpu * world, 1234, gather=same) if do_weak else None
    strong_job = weak_job if same else (Job(a.field_tiles, 1234, gather=True) if do_strong else None)
    main_job = weak_job if do_weak else strong_job

    # ---- warm-up
    # (the last warm-up step of each job goes through the end-to-end path as well: NCCL sets up its peer-to-peer
    # channels on the first gather, and the gather buffers enter torch's caching allocator)
    for w in range(a.warmup):
        last = w == a.warmup - 1
        main_job.run(100 + w, main_job.field_host.to(dev) if last else None, e2e=last)
        if strong_job is not None and strong_job is not main_job and (w == 0 or last):
            strong_job.run(100, strong_job.field_host.to(dev) if last else None, e2e=last)
    torch.cuda.synchronize(dev)
    barrier()

    clock = ClockSampler(local_rank)
    clock.start()
    time.sleep(0.3)
    main = timed(main_job, a.steps, 1000, e2e=False)         # device-resident
    main_e2e = timed(main_job, a.steps, 1000, e2e=True)      # pinned host tiles -> ... -> pinned host results
    strong = None
    if do_strong:
        if strong_job is main_job:
            s_dev, s_e2e = main, main_e2e
        else:
            s_dev = timed(strong_job, a.steps, 1000, e2e=False)
            s_e2e = timed(strong_job, a.steps, 1000, e2e=True)
        full = strong_job.summaries_global(s_dev["summ"]).float().contiguous()
        digest = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()[:16]
        strong = {"field_tiles": strong_job.T, "value": s_dev["evals"] / (s_dev["ms"] * 1e-3), "unit": UNIT,
                  "ms_per_step": s_dev["ms"] / a.steps, "tiles_per_sec": strong_job.T * a.steps / (s_dev["ms"] * 1e-3),
                  "e2e": {"value": s_e2e["evals"] / (s_e2e["ms"] * 1e-3), "unit": UNIT, "ms_per_step": s_e2e["ms"] / a.steps,
                          "h2d_bytes_per_step": s_e2e["h2d"], "d2h_bytes_per_step": s_e2e["d2h"]},
                  "smc_iters_per_step": s_dev["smc_iters"],
                  "checksum": {"sha256_16": digest, "sum_logz": float(full[:, 0].double().sum()),
                               "of": "per-tile summaries of the first timed step in global tile order "
                                     "(log Z, ESS, temperature, acceptance, mean detected count, mean detected flux; "
                                     "allstrata: log evidence, mean count, max count probability, log Z of the top count); "
                                     "seeded by the global tile id, so identical for every number of GPUs"},
                  "limit": "per-rank work shrinks with the GPU count while the slowest tile's serial chain of SMC "
                           "iterations stays: the tail of launches with few live tiles underfills a GPU"}
    clocks = clock.stop()

    value = main["evals"] / (main["ms"] * 1e-3)
    units_per_step = main_job.T * ns
    tiles_per_sec = main_job.T * a.steps / (main["ms"] * 1e-3)

    # ---- roofline of the dominant kernel (mh_kernel), timed live with CUDA events per launch.
    # Units one launch processes, per live particle: 1 full render (entry state) of D stars and num_iters sweeps that
    # each evaluate 2 stars (the one removed and the one proposed) on the P pixels, plus P pixel terms per evaluation.
    # Per-unit figures are SURVEY.md 8(d)'s: M71 4 MUFU / 12 FP32 instr per (star, pixel) and 2 / 7 per Normal pixel term;
    # Gaussian-PSF model 1 / 6 per (star, pixel) and 1 / 3 per Poisson pixel term.
    c_psf, i_psf, c_pix, i_pix = (4, 12, 2, 7) if is_m71(a) else (1, 6, 1, 3)
    mh_ms = sum(ev0.elapsed_time(ev1) for log in main["logs"] for (ev0, ev1, *_r) in log)
    n_launch = sum(len(log) for log in main["logs"])
    P = TILE * TILE
    star_pixels = (1 * D + 2 * iters) * P   # entry render + two-star sweeps (no final refresh by default)
    pixel_terms = (iters + 1) * P
    mufu_per_particle = c_psf * star_pixels + c_pix * pixel_terms
    fp32_per_particle = i_psf * star_pixels + i_pix * pixel_terms
    # what the kernel issues per (sweep, star): separable Gaussians (2 terms x (8 + 8) ex2; 1 term for the Gaussian
    # model), and for the M71 wing 1 lg2 + 0.75 ex2 per star-pixel (a quarter of the ex2 run on the FMA pipe); four
    # pixels share one rcp and one lg2 (Normal) or one lg2 per pixel (Poisson); ~45 in the proposal step
    per_star = (2 * 16 + 1.75 * P) if is_m71(a) else 16
    exec_mufu_per_particle = (1 * D + 2 * iters) * per_star + pixel_terms * (0.5 if is_m71(a) else 1.0) + 45 * iters
    live_particles = main["live"] * N
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    sfu_peak = 148 * 16 * sm_max * 1e6 / 1e12          # TOP/s (MUFU results per second)
    fp32_peak = 148 * 128 * sm_max * 1e6 / 1e12        # T instr/s (FMA = 1 instr)
    achieved = live_particles * mufu_per_particle / (mh_ms * 1e-3) / 1e12
    # read the catalog and its count through the 8-byte resampling index, write catalog + count, loglik out
    bytes_per_launch_particle = 2 * (12 * D + 4) + 8 + 4
    hbm_gbs = live_particles * bytes_per_launch_particle / (mh_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per particle-launch from the committed ncu --set full capture of this kernel
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_mh_kernel.json")))
        # (captured at D = 10: 248 algorithmic bytes per particle; other catalog sizes are scaled by their own figure)
        traffic = (prof["dram_bytes_per_particle"] * bytes_per_launch_particle / prof["algorithmic_bytes_per_particle"]
                   * live_particles / max(1, n_launch))
    except Exception:  # noqa: BLE001
        pass
    kname = "mh_kernel<M71,8,8,TPP=1,GATHER>" if is_m71(a) else "mh_kernel<GAUSS,8,8,TPP=1,GATHER>"
    roofline = {"kernel": f"{kname} (smcdet_mh_mutate_resampled: the MH sweeps with the resampling step's gather fused in)",
                "bound": "sfu",
                "achieved": achieved, "peak": sfu_peak, "unit": "TOP/s (MUFU)", "frac": achieved / sfu_peak,
                "peak_source": f"derived: 148 SMs x 16 MUFU lanes x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); "
                               "the path is SFU/FP32-bound, not HBM- or tensor-bound (SURVEY.md 8d)",
                "definition": "algorithmic MUFU of the units a launch processes (SURVEY 8d per (star,pixel) PSF evaluation "
                              "and per pixel term; 1 full render + num_iters two-star sweeps per particle) / CUDA-event "
                              "time of the launches.  It can exceed 1 because the kernel evaluates the Gaussian PSF terms "
                              "separably and part of the wing's exponentials on the FMA pipe: executed_frac is the MUFU "
                              "actually issued / peak (ncu sm__inst_executed_pipe_xu agrees, profiles/)",
                "executed_frac": live_particles * exec_mufu_per_particle / (mh_ms * 1e-3) / 1e12 / sfu_peak,
                "fp32_achieved_tinstr": live_particles * fp32_per_particle / (mh_ms * 1e-3) / 1e12,
                "fp32_peak_tinstr": fp32_peak,
                "launches": n_launch, "avg_launch_ms": mh_ms / max(1, n_launch),
                "share_of_step": mh_ms / main["ms"], "traffic": traffic,
                "traffic_source": "dram__bytes_read+write per particle of the committed ncu --set full capture "
                                  "(profiles/r02_ncu_mh_kernel.json: M71 model, D = 10; scaled by 24 D + 8 bytes for other "
                                  "catalog sizes) x the live particles of this run",
                "bound_note": None if is_m71(a) else "the Gaussian-PSF kernel issues ~110 MUFU against ~1500 other "
                              "instructions per sweep: it is bound by instruction issue / latency (ncu: XU 40 %, issue "
                              "62 %, profiles/r02_ncu_mh_gauss.txt), so frac is not its pipe utilisation; "
                              "fp32_achieved_tinstr / fp32_peak_tinstr is the figure that applies (DESIGN.md 3.7)",
                "algorithmic_bytes_per_launch": live_particles * bytes_per_launch_particle / max(1, n_launch),
                "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None}}

    # ---- the standalone likelihood kernel on the same field (dense: D stars x P pixels per evaluation)
    Tl = min(main_job.T_local, 800)
    prior_d = make_objects(argparse.Namespace(**dict(vars(a), workload="m71synthetic" if strata else a.workload)))[1]
    counts0, locs0, fluxes0 = prior_d._sample_grid(Tl, 1, None, True, N, seed=7)
    tiles_ll = main_job.field[:Tl].view(Tl, 1, TILE, TILE)
    for _ in range(3):
        model.loglikelihood(tiles_ll, locs0, fluxes0)
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for x0, x1 in evs:
        x0.record(torch.cuda.current_stream(dev))
        model.loglikelihood(tiles_ll, locs0, fluxes0)
        x1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    ll_ms = sorted(x0.elapsed_time(x1) for x0, x1 in evs)[len(evs) // 2]
    ll_rate = Tl * N / (ll_ms * 1e-3)
    ll_mufu = c_psf * D * P + c_pix * P
    roofline_loglik = {"kernel": f"loglik_kernel<{'M71' if is_m71(a) else 'GAUSS'},8,8,TPP=1> (smcdet_loglik)", "bound": "sfu",
                       "evals_per_s": ll_rate, "launch_ms": ll_ms,
                       "achieved": ll_rate * ll_mufu / 1e12, "peak": sfu_peak, "unit": "TOP/s (MUFU)",
                       "frac": ll_rate * ll_mufu / 1e12 / sfu_peak,
                       "executed_frac": ll_rate * (D * per_star + P * (0.5 if is_m71(a) else 1.0)) / 1e12 / sfu_peak,
                       "hbm_gbs": ll_rate * (12 * D + 8) / 1e9}
    del counts0, locs0, fluxes0

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base, _, _ = cpu_measure(a, target_seconds=12.0, steps=1, warmup=0)
        try:
            real = reference_measure(a, target_seconds=12.0)
        except Exception as exc:  # noqa: BLE001
            real = {"kind": "reference", "unavailable": repr(exc)[:200]}
        cpu_base["also"] = [real] if real else []

    if rank == 0:
        e2e_value = main_e2e["evals"] / (main_e2e["ms"] * 1e-3)
        per_eval = (iters + 2)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": main["ms"] / a.steps, "higher_is_better": True,
                "scaling": "weak" if do_weak else "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(a), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": main_e2e["h2d"],
                        "d2h_bytes_per_step": main_e2e["d2h"], "ms_per_step": main_e2e["ms"] / a.steps,
                        "tiles_per_sec": main_job.T * a.steps / (main_e2e["ms"] * 1e-3),
                        "path": ("pinned host tiles -> H2D -> sampler -> " + ("gather of the weighted catalogs onto rank 0 -> "
                                 if main_job.gather else "(per rank) ") +
                                 "Aggregate finish (resample + prune) -> pruned catalogs + summaries D2H into pinned buffers")
                                if not strata else
                                "pinned host tiles -> H2D -> count-stratified samplers on every rank -> all-gather of the "
                                "per-stratum evidences -> count posterior per tile -> D2H"},
                "gpu_launches": main["launches"] + main_e2e["launches"], "kernel_calls": dict(lib.calls),
                "launches_per_step": main["launches"] / a.steps,
                "roofline": roofline, "roofline_loglik": roofline_loglik, "cpu_baseline": cpu_base,
                "tiles_per_sec": tiles_per_sec, "smc_iters_per_step": main["smc_iters"],
                "mean_smc_iters_per_segment": main["live"] * (world if not strata else 1) / (units_per_step * a.steps)
                if not strata else None,
                # the same rate in other units: `value` counts N*(iters+2) evaluations per live tile-iteration as the
                # reference performs them; the kernel does iters+1 pixel sums and two-star incremental updates instead
                "proposals_per_s": value * iters / per_eval,
                "dense_equiv": {"definition": "evaluations/s a kernel that re-rendered all D stars per evaluation (as the "
                                              "reference does; BASELINE.md section 3 SFU peak: 1.73e9 for M71 D = 10) would "
                                              "need for the same step time", "value": value,
                                "dense_sfu_peak_evals_per_s": sfu_peak * 1e12 / ll_mufu,
                                "note": "value exceeds the dense SFU peak because a sweep re-renders 2 stars, not D"},
                "strong": strong}
        if strata:
            line["strata_per_sec"] = units_per_step * a.steps / (main["ms"] * 1e-3)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)
