#!/usr/bin/env python
"""Benchmark of the smcdet per-tile SMC hot path on B200 (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]          own arm (CUDA library)
  python bench.py --impl reference [...]                        reference arm (CPU port of the reference)
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU

Workload (BASELINE.json configs[1], "m71synthetic"): a synthetic SDSS-r-band-like field drawn from the
M71 model itself, cut into 8x8 tiles; per tile a likelihood-tempered SMC sampler with N = 10 000
catalogs of D = 10 stars, 100 single-site MH sweeps per SMC iteration, ESS threshold 0.5 N,
multinomial resampling, run to temperature 1 (notebooks/smc.ipynb cells 3-7 of the reference).
Every rank owns `--tiles-per-gpu` tiles (weak scaling; tiles are independent, no data-path collective).

A "step" is one complete SMCsampler.run() over the rank's tiles.  `value` = particle-likelihood
evaluations per second over the whole job, counted as the reference evaluates them: per SMC iteration
and live tile N*(num_iters + 2) (kernel.py:64-70, :89-96; sampler.py:100-102), plus N for initialize.
"""

import argparse
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
TILE, PAD = 8, 4
METRIC, UNIT = "particle-likelihood evals/sec", "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--tiles-per-gpu", type=int, default=800)
    ap.add_argument("--particles", type=int, default=10000)
    ap.add_argument("--stars", type=int, default=None, help="stars per catalog (default 10; 16 for m71semisynthetic)")
    ap.add_argument("--workload", default="m71synthetic", choices=["m71synthetic", "m71semisynthetic"],
                    help="m71semisynthetic = BASELINE.json configs[2]: three times the source density, D = 16 (SURVEY.md 8d)")
    ap.add_argument("--mh-iters", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.stars is None:
        a.stars = 16 if a.workload == "m71semisynthetic" else 10
    return a


def workload_config(a):
    dense = getattr(a, "workload", "m71synthetic") == "m71semisynthetic"
    name = ("m71semisynthetic-shaped: as m71synthetic with three times the source density and larger catalogs "
            "(BASELINE.json configs[2]; SURVEY.md 8d config 3)") if dense else (
        "m71synthetic: M71 PSF + Normal likelihood, 8x8 tiles, psf_radius 8, pad 4 "
        "(BASELINE.json configs[1]; notebooks/smc.ipynb of the reference)")
    return {"workload": name,
            "tiles_per_gpu": a.tiles_per_gpu, "particles_per_tile": a.particles, "stars_per_catalog": a.stars,
            "mh_iters": a.mh_iters, "ess_threshold_prop": 0.5, "resample": "multinomial",
            "loglik_for_tempering": "from the incrementally updated rate image (default; 7e-7 relative drift measured)",
            "step": "one full SMCsampler.run() to temperature 1 over all tiles of the rank",
            "finished_tiles": "frozen (each tile runs as in the reference's per-tile loop, experiments/m71/run_smc.py:113-124)",
            "parallelism": f"tiles sharded over {a.gpus} GPU(s), no data-path collective; final all_gather of catalogs",
            "l2": "particle state per rank (>= 1 GB at the default size) exceeds the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (C port of the reference's arithmetic), all host threads
# ----------------------------------------------------------------------------------------------
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

    om = O.m71_model(M71["psf_radius"], M71["psf_params"], M71["background"], M71["adu_per_nmgy"],
                     M71["noise_additive"], M71["noise_multiplicative"])
    op = O.m71_prior(a.stars, a.stars, PRIOR["counts_rate"], TILE, TILE, PRIOR["flux_alpha"], PRIOR["flux_lower"],
                     PRIOR["flux_upper"], pad=PAD)
    mh = O.make_mh(a.mh_iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"], (-PAD, -PAD), (TILE + PAD, TILE + PAD))
    rng = np.random.default_rng(seed)
    N, D = a.particles, a.stars
    # observed tiles: drawn from the model with a handful of stars (true prior of the notebook)
    tp = O.m71_prior(4, 4, PRIOR["counts_rate"], TILE, TILE, PRIOR["flux_alpha"], DETECTION, PRIOR["flux_upper"], pad=PAD)
    _, tl, tf = O.prior_sample(tp, rng.random((n_tiles, 1, 4, 2), dtype=np.float32), rng.random((n_tiles, 1, 4), dtype=np.float32), 1)
    rate = O.render(om, tl, tf, TILE, TILE)[..., 0]
    tiles = (rate + rng.standard_normal(rate.shape) * np.sqrt(M71["noise_additive"] + M71["noise_multiplicative"] * rate)).astype(np.float32)
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


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms, n_tiles = cpu_measure(a, target_seconds=6.0, steps=a.steps, warmup=a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tiles_per_sec_full_smc_est": base["value"] / (a.particles * (a.mh_iters + 2) * 12.0),
            "note": "CPU port (oracle/smcdet_oracle.c, OpenMP) of the reference's arithmetic: the reference itself is "
                    "pure Python/PyTorch and /root/reference does not exist on the GPU box; the unmodified torch path "
                    "measured 1.2-1.45e4 evals/s on 8 threads in the build container (BASELINE.md section 2)"}
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
                "source": "NVML (pynvml) every 200 ms during the timed region", **({"error": self.error} if hasattr(self, "error") else {})}


def make_field(a, rank, dev):
    """Synthetic field: tiles drawn from the M71 model with the reference's true prior
    (notebooks/smc.ipynb cell 3); tile g of the job is seeded by its global id."""
    import torch

    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.prior import M71Prior

    model = M71ImageModel(TILE, TILE, **M71)
    true_prior = M71Prior(0, 24, PRIOR["counts_rate"], TILE, TILE, flux_alpha=PRIOR["flux_alpha"],
                          flux_lower=DETECTION, flux_upper=PRIOR["flux_upper"], pad=PAD)
    T = a.tiles_per_gpu
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    density = 3.0 if getattr(a, "workload", "m71synthetic") == "m71semisynthetic" else 1.0
    counts = torch.poisson(torch.full((T,), density * float(true_prior._count_rate()), device=dev), generator=g).clamp(max=24)
    D = 24
    low, high = -PAD, TILE + PAD
    locs = low + torch.rand(T, 1, 1, D, 2, device=dev, generator=g) * (high - low)
    al = PRIOR["flux_alpha"]
    ua, la = PRIOR["flux_upper"] ** al, DETECTION ** al
    u = torch.rand(T, 1, 1, D, device=dev, generator=g)
    fluxes = ((ua - u * ua + u * la) / (la * ua)) ** (-1.0 / al)
    mask = torch.arange(D, device=dev).view(1, 1, 1, D) < counts.view(T, 1, 1, 1)
    rate = model._rate(locs * mask.unsqueeze(-1), fluxes * mask)  # [T,1,8,8,1]
    noise = torch.randn(rate.shape, device=dev, generator=g)
    img = rate + noise * (model.noise_additive + model.noise_multiplicative * rate).sqrt()
    return img[..., 0].contiguous()  # [T,1,8,8]


def run_own(a):
    import torch
    import torch.distributed as dist

    from smcdet_b200 import _lib as L
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior
    from smcdet_b200.sampler import SMCsampler
    from smcdet_b200.shard import gather_tiles

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

    T, N, D, iters = a.tiles_per_gpu, a.particles, a.stars, a.mh_iters
    model = M71ImageModel(TILE, TILE, **M71)
    prior = M71Prior(D, D, PRIOR["counts_rate"], TILE, TILE, flux_alpha=PRIOR["flux_alpha"],
                     flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=PAD)
    tiles_dev = make_field(a, rank, dev)
    tiles_host = tiles_dev.cpu().pin_memory()
    tile_ids = (torch.arange(T, device=dev, dtype=torch.int64) * world + rank).view(T, 1)
    lib = L.lib()

    def one_run(tiles, seed, mh):
        torch.manual_seed(seed)
        s = SMCsampler(tiles, TILE, prior, model, mh, N, 0.5, "multinomial", DETECTION, 200, tile_ids=tile_ids,
                       freeze_finished=True, verbose=False)
        s.run()
        return s

    def count_evals(s, mh):
        # per launch: live tiles x N x (iters + 2); + T*N for initialize
        live = sum(int(T if act is None else act.sum().item()) for (_, _, act, *_r) in mh.event_log)
        return live * N * (iters + 2) + T * N, live

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- warm-up
    for w in range(a.warmup):
        mh = SingleComponentMH(iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        mh.event_log = []
        one_run(tiles_dev, 100 + w, mh)
    barrier()

    # ---- device-resident timing: K steps
    sampler_clock = ClockSampler(local_rank)
    sampler_clock.start()
    time.sleep(0.3)
    launches0 = lib.launches
    evals = live_total = 0
    mh_logs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(torch.cuda.current_stream(dev))
    iters_smc = []
    for k in range(a.steps):
        mh = SingleComponentMH(iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        mh.event_log = []
        s = one_run(tiles_dev, 1000 + k, mh)
        mh_logs.append(mh.event_log)
        iters_smc.append(s.iter)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.launches - launches0
    clocks = sampler_clock.stop()
    for log in mh_logs:
        live = sum(int(T if act is None else act.sum().item()) for (_, _, act, *_r) in log)
        live_total += live
        evals += live * N * (iters + 2) + T * N
    evals_all = sum_over_ranks(float(evals))
    value = evals_all / (elapsed_ms * 1e-3)
    tiles_per_sec = a.gpus * T * a.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (mh_kernel), timed live with CUDA events per launch.
    # Units one launch processes, per live particle: 1 full render (entry state) of D stars and
    # num_iters sweeps that each evaluate 2 stars (the one removed and the one proposed) on the P pixels, plus
    # P pixel terms per evaluation.  Per-unit figures are SURVEY.md 8(d)'s: 4 MUFU / 12 FP32 instr per
    # (star, pixel) PSF evaluation of the M71 model, 2 MUFU / 7 FP32 instr per Normal pixel term.
    mh_ms = sum(ev0.elapsed_time(ev1) for log in mh_logs for (ev0, ev1, *_r) in log)
    n_launch = sum(len(log) for log in mh_logs)
    P = TILE * TILE
    star_pixels = (1 * D + 2 * iters) * P   # entry render + two-star sweeps (no final refresh by default)
    pixel_terms = (iters + 1) * P
    mufu_per_particle = 4 * star_pixels + 2 * pixel_terms
    fp32_per_particle = 12 * star_pixels + 7 * pixel_terms
    # what the kernel issues: separable Gaussians (2*(8+8) ex2 per star) + 2 MUFU per star-pixel for the wing +
    # 1 for the star weight; four pixels share one rcp and one lg2 (0.5 per pixel); ~45 in the proposal step
    exec_mufu_per_particle = (1 * D + 2 * iters) * (2 * 16 + 2 * P + 1) + pixel_terms // 2 + 45 * iters
    live_particles = live_total * N
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    sfu_peak = 148 * 16 * sm_max * 1e6 / 1e12          # TOP/s (MUFU results per second)
    fp32_peak = 148 * 128 * sm_max * 1e6 / 1e12        # T instr/s (FMA = 1 instr)
    achieved = live_particles * mufu_per_particle / (mh_ms * 1e-3) / 1e12
    bytes_per_launch_particle = 2 * (12 * D) + 4 + 4   # read + write catalog, count, loglik out
    hbm_gbs = live_particles * bytes_per_launch_particle / (mh_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per particle-launch from the committed ncu --set full capture of this kernel
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_mh_kernel.json")))
        traffic = prof["dram_bytes_per_particle"] * live_particles / max(1, n_launch)
    except Exception:
        pass
    roofline = {"kernel": "mh_kernel<M71,8,8,TPP=1> (smcdet_mh_mutate)", "bound": "sfu",
                "achieved": achieved, "peak": sfu_peak, "unit": "TOP/s (MUFU)", "frac": achieved / sfu_peak,
                "peak_source": f"derived: 148 SMs x 16 MUFU lanes x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); "
                               "the path is SFU/FP32-bound, not HBM- or tensor-bound (SURVEY.md 8d)",
                "definition": "algorithmic MUFU of the units a launch processes (SURVEY 8d: 4 per (star,pixel) PSF "
                              "evaluation, 2 per pixel term; 1 full render + num_iters two-star sweeps per particle) / "
                              "CUDA-event time of the launches.  It can exceed 1 because the kernel evaluates the two "
                              "Gaussian PSF terms separably (about 2.5 MUFU per star-pixel issued): executed_frac is "
                              "the MUFU actually issued / peak (ncu sm__inst_executed_pipe_xu agrees, profiles/)",
                "executed_frac": live_particles * exec_mufu_per_particle / (mh_ms * 1e-3) / 1e12 / sfu_peak,
                "fp32_achieved_tinstr": live_particles * fp32_per_particle / (mh_ms * 1e-3) / 1e12,
                "fp32_peak_tinstr": fp32_peak,
                "launches": n_launch, "avg_launch_ms": mh_ms / max(1, n_launch),
                "share_of_step": mh_ms / elapsed_ms, "traffic": traffic,
                "algorithmic_bytes_per_launch": live_particles * bytes_per_launch_particle / max(1, n_launch),
                "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None}}

    # ---- the standalone likelihood kernel on the same field (dense: D stars x P pixels per evaluation)
    counts0, locs0, fluxes0 = prior._sample_grid(T, 1, None, True, N, seed=7)
    for _ in range(3):
        model.loglikelihood(tiles_dev, locs0, fluxes0)
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for x0, x1 in evs:
        x0.record(torch.cuda.current_stream(dev))
        model.loglikelihood(tiles_dev, locs0, fluxes0)
        x1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    ll_ms = sorted(x0.elapsed_time(x1) for x0, x1 in evs)[len(evs) // 2]
    ll_rate = T * N / (ll_ms * 1e-3)
    roofline_loglik = {"kernel": "loglik_kernel<M71,8,8,TPP=1> (smcdet_loglik)", "bound": "sfu",
                       "evals_per_s": ll_rate, "launch_ms": ll_ms,
                       "achieved": ll_rate * (4 * D * P + 2 * P) / 1e12, "peak": sfu_peak, "unit": "TOP/s (MUFU)",
                       "frac": ll_rate * (4 * D * P + 2 * P) / 1e12 / sfu_peak,
                       "executed_frac": ll_rate * (D * (2 * 16 + 2 * P + 1) + P // 2) / 1e12 / sfu_peak,
                       "hbm_gbs": ll_rate * (12 * D + 8) / 1e9}
    del counts0, locs0, fluxes0

    # ---- end to end through the public API: pinned host tiles -> H2D -> run -> results D2H into pinned buffers.
    #      Per-tile summaries are all-gathered to every rank; each rank reads back its own posterior catalogs.
    #      The read-back of step k runs on a copy stream and overlaps the sampling of step k+1 (two pinned
    #      buffer sets); everything is complete before the closing event.
    copy_stream = torch.cuda.Stream(device=dev)
    shapes = [((T * world, 4), torch.float32), ((T, N), torch.int16), ((T, N, D, 2), torch.float32), ((T, N, D), torch.float32)]
    host_out = [[torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes] for _ in range(2)]  # allocated once
    pending = []
    barrier()
    e0.record(torch.cuda.current_stream(dev))
    h2d = d2h = 0
    e2e_evals = 0
    e2e_logs = []
    for k in range(a.steps):
        mh = SingleComponentMH(iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        mh.event_log = []
        tiles = tiles_host.to(dev, non_blocking=True)
        h2d += tiles_host.numel() * 4
        s = one_run(tiles, 1000 + k, mh)
        summ = torch.stack([s.log_normalizing_constant, s.ess, s.posterior_mean_count(s.pruned_counts.float()),
                            s.posterior_mean_total_flux(s.pruned_fluxes)], -1).view(T, 4)
        if world > 1:
            summ = gather_tiles(summ.contiguous(), T * world)
        outs = [summ, s.pruned_counts.view(T, N).to(torch.int16), s.pruned_locs.view(T, N, D, 2),
                s.pruned_fluxes.view(T, N, D)]
        slot = k % 2
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            for hbuf, o in zip(host_out[slot], outs):
                hbuf.copy_(o, non_blocking=True)
                o.record_stream(copy_stream)
                d2h += o.numel() * o.element_size()
        pending.append(outs)
        if len(pending) > 1:
            pending.pop(0)
        e2e_logs.append(mh.event_log)
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    for log in e2e_logs:
        live = sum(int(T if act is None else act.sum().item()) for (_, _, act, *_r) in log)
        e2e_evals += live * N * (iters + 2) + T * N
    e2e_value = sum_over_ranks(float(e2e_evals)) / (e2e_ms * 1e-3)

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base, _, _ = cpu_measure(a, target_seconds=12.0, steps=1, warmup=0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(a), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps,
                        "d2h_bytes_per_step": d2h // a.steps, "ms_per_step": e2e_ms / a.steps,
                        "tiles_per_sec": a.gpus * T * a.steps / (e2e_ms * 1e-3)},
                "gpu_launches": launches, "kernel_calls": dict(lib.calls), "roofline": roofline,
                "roofline_loglik": roofline_loglik,
                "cpu_baseline": cpu_base, "tiles_per_sec": tiles_per_sec, "smc_iters_per_step": iters_smc,
                "mean_smc_iters_per_tile": live_total / (T * a.steps)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)
