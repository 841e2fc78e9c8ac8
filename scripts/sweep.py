"""Likelihood microbenchmark sweep (BASELINE.json configs[4]): particles 1k-1M x stars 1-16 x tile side 8-32,
both image models, plus the fused MH kernel per tile size.  Writes gpurun_out/sweep.json and sweep.md."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from smcdet_b200.images import ImageModel, M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior, ParetoStarPrior
from smcdet_b200.sampler import SMCsampler
from bench import M71, PRIOR, DETECTION

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
SFU_PEAK = 148 * 16 * 1.965e9

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]

def objects(model, tile, D):
    if model == "m71":
        im = M71ImageModel(tile, tile, **M71)
        pr = M71Prior(D, D, PRIOR["counts_rate"], tile, tile, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"],
                      flux_upper=PRIOR["flux_upper"], pad=4)
        mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        c_psf, c_pix = 4, 2
    else:
        im = ImageModel(tile, tile, background=200.0, psf_radius=8, psf_stdev=0.93)
        pr = ParetoStarPrior(D, D, tile, tile, flux_scale=345.84, flux_alpha=2.0, pad=2)
        mh = SingleComponentMH(100, 0.1, 100.0, 345.84, 1e6)
        c_psf, c_pix = 1, 1
    return im, pr, mh, c_psf, c_pix

rows = []
for model in ("m71", "gauss"):
    for tile in (8, 16, 32):
        for D in (1, 2, 4, 8, 16):
            for N in (1000, 10000, 100000, 1000000):
                T = max(1, min(4096, (1 << 20) // N))
                if T * N * D * 12 > 8e9:
                    continue
                im, pr, mh, c_psf, c_pix = objects(model, tile, D)
                counts, locs, fluxes = pr._sample_grid(T, 1, None, True, N, seed=1)
                tiles = (200.0 + 50.0 * torch.rand(T, 1, tile, tile, device=dev)).round()
                ms = timeit(lambda: im.loglikelihood(tiles, locs, fluxes))
                evals = T * N / (ms * 1e-3)
                mufu = c_psf * D * tile * tile + c_pix * tile * tile
                rows.append(dict(kernel="loglik", model=model, tile=tile, D=D, N=N, T=T, ms=round(ms, 4), evals_per_s=evals,
                                 sfu_frac_algorithmic=evals * mufu / SFU_PEAK))
                print(json.dumps(rows[-1]), flush=True)
    for tile, D, T, N in ((8, 10, 296, 10000), (8, 8, 296, 10000), (16, 10, 74, 10000), (32, 10, 20, 10000), (8, 16, 148, 10000)):
        im, pr, mh, c_psf, c_pix = objects(model, tile, D)
        counts, locs, fluxes = pr._sample_grid(T, 1, None, True, N, seed=1)
        tiles = (200.0 + 50.0 * torch.rand(T, 1, tile, tile, device=dev)).round()
        s = SMCsampler(tiles, tile, pr, im, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
        s.temperature = torch.full((T, 1), 0.3, device=dev)
        ms = timeit(lambda: mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1), n=3, warm=1)
        evals = T * N * 102 / (ms * 1e-3)
        mufu = c_psf * D * tile * tile + c_pix * tile * tile
        rows.append(dict(kernel="mh(100 sweeps)", model=model, tile=tile, D=D, N=N, T=T, ms=round(ms, 3), evals_per_s=evals,
                         sfu_frac_algorithmic=evals * mufu / SFU_PEAK))
        print(json.dumps(rows[-1]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=0)
with open(os.path.join(ROOT, "gpurun_out", "sweep.md"), "w") as f:
    f.write("| kernel | model | tile | D | N | T | ms | evals/s | algorithmic MUFU / SFU peak |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['kernel']} | {r['model']} | {r['tile']} | {r['D']} | {r['N']} | {r['T']} | {r['ms']} | {r['evals_per_s']:.3e} | {r['sfu_frac_algorithmic']:.2f} |\n")
