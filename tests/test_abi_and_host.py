"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes mirror matches the header, the package fails loudly without CUDA, and the
tile-sharding host logic works across two gloo ranks."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "smcdet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smcdet_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from smcdet_b200 import _abi, _lib

    path = _lib.build()
    cdll = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/smcdet_b200.h but not exported"
    assert sorted(_abi.PROTOTYPES) == names, "smcdet_b200/_abi.py and the header disagree"
    _abi.bind(cdll)
    assert cdll.smcdet_version() == _abi.ABI_VERSION
    # sm_100a SASS is in the binary
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_the_header():
    from smcdet_b200 import _abi

    assert ctypes.sizeof(_abi.ModelParams) == 15 * 4
    assert ctypes.sizeof(_abi.PriorParams) == 4 * 4 + 4 + 4 * 4 + 6 * 4
    # 13 four-byte members (52 bytes), padding to the pointer's alignment, the tile_of_segment pointer
    assert ctypes.sizeof(_abi.MHParams) == 56 + ctypes.sizeof(ctypes.c_void_p)
    assert _abi.MHParams.tile_of_segment.offset == 56 and _abi.MHParams.acc_as_count.offset == 44
    assert ctypes.sizeof(_abi.LoopState) == 4 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_abi.DrawTape) == 4 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_abi.MHTrace) == 5 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_abi.ResampledSource) == 8 * ctypes.sizeof(ctypes.c_void_p)
    assert _abi.ResampledSource.counts_out.offset == 4 * ctypes.sizeof(ctypes.c_void_p)
    assert _abi.ResampledSource.rates_out.offset == 7 * ctypes.sizeof(ctypes.c_void_p)


def test_invalid_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from smcdet_b200 import _abi, _lib

    cdll = _abi.bind(ctypes.CDLL(_lib.build()))
    m = _abi.ModelParams()
    m.model_kind = 7
    assert cdll.smcdet_loglik(ctypes.byref(m), None, None, None, None, 1, 1, 1, 8, 8, None) == _abi.E_INVALID
    assert b"smcdet_loglik" in cdll.smcdet_last_error_string()
    m.model_kind = _abi.MODEL_M71_NORMAL
    assert cdll.smcdet_loglik(ctypes.byref(m), None, None, None, None, 1, 1, 1, 8, 8, None) == _abi.E_INVALID
    assert cdll.smcdet_resample(5, None, None, 0, None, None, None, None, 1, 1, None) == _abi.E_INVALID
    assert cdll.smcdet_temper_update(None, None, None, 1.0, 1, None, None, None, None, None, None, None, 1, 1, None) == _abi.E_INVALID
    # carried expected-count images (ABI v8): reading and writing the same buffer is refused before anything is launched
    buf = [ctypes.create_string_buffer(64) for _ in range(12)]
    ad = [ctypes.addressof(b) for b in buf]
    pr, mh = _abi.PriorParams(), _abi.MHParams()
    mh.num_iters, mh.refresh_loglik = 1, 1
    src = _abi.ResampledSource(ad[0], ad[1], ad[2], ad[3], ad[4], None, ad[5], ad[5])
    rc = cdll.smcdet_mh_mutate_resampled(ctypes.byref(m), ctypes.byref(pr), ctypes.byref(mh), ad[6], ctypes.byref(src), ad[7], ad[8],
                                         ad[9], ad[10], ad[11], None, None, 0, 0, None, None, None, 1, 1, 1, 8, 8, None)
    assert rc == _abi.E_INVALID and b"rates" in cdll.smcdet_last_error_string()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from smcdet_b200 import _lib
    from smcdet_b200.images import M71ImageModel

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.device()
    model = M71ImageModel(8, 8, background=100.0, psf_radius=8, adu_per_nmgy=240.0,
                          psf_params=[1.1, 2.0, 2.3, 5.2, 0.73, 0.51], noise_additive=0.0, noise_multiplicative=1.9)
    assert abs(float(model.psf_normalizing_constant) - 12.75) < 0.5  # host-side constant still computed
    with pytest.raises((RuntimeError, TypeError)):
        model.loglikelihood(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 4, 2, 2), torch.zeros(1, 1, 4, 2))


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "smcdet_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "smcdet_b200", fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} mentions the oracle"
            assert "hostsim" not in src, f"{fn} mentions hostsim"


def test_round_robin_sharding():
    from smcdet_b200.shard import shard_sizes, shard_tile_ids

    for T, G in [(800, 8), (7, 4), (3, 8), (1024, 1)]:
        ids = [shard_tile_ids(T, G, r) for r in range(G)]
        assert sorted(torch.cat(ids).tolist()) == list(range(T))
        assert [len(i) for i in ids] == shard_sizes(T, G)
        assert max(len(i) for i in ids) - min(len(i) for i in ids) <= 1
    # sharding by merge blocks: every tile once, a rank's tiles block by block (row-major inside a block)
    from smcdet_b200.shard import block_tile_ids

    for G in (1, 2, 3, 8):
        parts = [block_tile_ids((40, 20), 4, G, r) for r in range(G)]
        assert sorted(torch.cat([p[0] for p in parts]).tolist()) == list(range(800))
        assert sorted(torch.cat([p[1] for p in parts]).tolist()) == list(range(50))
        ids, blocks = parts[0]
        first = ids[:16].view(4, 4)
        assert first[0].tolist() == [0, 1, 2, 3] and first[1].tolist() == [20, 21, 22, 23] and blocks[0] == 0


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from smcdet_b200.shard import gather_tiles, gather_tiles_to_root, shard_tile_ids
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
T = 7
ids = shard_tile_ids(T, world, rank)
local = torch.stack([torch.full((3, 2), float(t)) + torch.arange(6).view(3, 2) / 10 for t in ids.tolist()]) if len(ids) else torch.zeros(0, 3, 2)
full = gather_tiles(local, T)
want = torch.stack([torch.full((3, 2), float(t)) + torch.arange(6).view(3, 2) / 10 for t in range(T)])
assert torch.equal(full, want), (rank, full)
cnt = gather_tiles(ids.clone(), T)
assert cnt.tolist() == list(range(T))
# the gather onto the rank that runs the Aggregate sink: the full field on the root only
root = gather_tiles_to_root(local, T, dst=0)
assert (root is None) == (rank != 0)
if rank == 0:
    assert torch.equal(root, want)
# a field with fewer tiles than ranks: the last rank's shard is empty and still takes part in the collectives
one = shard_tile_ids(1, world, rank)
assert len(one) == (1 if rank == 0 else 0)
loc1 = torch.full((len(one), 2), 5.0)
assert torch.equal(gather_tiles(loc1, 1), torch.full((1, 2), 5.0))
r1 = gather_tiles_to_root(loc1, 1, dst=0)
assert (r1 is None) == (rank != 0) and (rank != 0 or torch.equal(r1, torch.full((1, 2), 5.0)))
# the (tile, count) strata of count-stratified SMC as a second sharding axis: every stratum on exactly one rank,
# loads balanced by expected cost, the assignment identical on every rank
from smcdet_b200.cssmc import CountStratifiedSMC as CS
parts = CS.assign_strata(13, list(range(0, 9)), world)
assert sorted(sum(parts, [])) == list(range(13 * 9)) and parts == CS.assign_strata(13, list(range(0, 9)), world)
loads = [sum(CS.stratum_cost(i % 9) for i in p) for p in parts]
assert max(loads) - min(loads) <= CS.stratum_cost(8)
mine = torch.tensor([float(len(parts[rank]))])
both = [torch.zeros(1) for _ in range(world)]
dist.all_gather(both, mine)
assert sum(int(b.item()) for b in both) == 13 * 9
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_strata_assignment_balances_expected_cost():
    from smcdet_b200.cssmc import CountStratifiedSMC as CS

    counts = list(range(0, 11))
    for world in (1, 2, 8):
        parts = CS.assign_strata(100, counts, world)
        assert sorted(sum(parts, [])) == list(range(100 * 11))
        loads = [sum(CS.stratum_cost(counts[i % 11]) for i in p) for p in parts]
        assert max(loads) - min(loads) <= CS.stratum_cost(10), loads
    # round-robin over the flat stratum index would leave rank loads as unequal as the counts themselves
    rr = [sum(CS.stratum_cost(counts[i % 11]) for i in range(r, 1100, 11)) for r in range(11)]
    assert max(rr) > 3 * min(rr)


def test_gather_tiles_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_distribution_classes_follow_their_definitions():
    """The thin torch classes of smcdet_b200.distributions (device-agnostic, so checkable on the CPU):
    truncated normal draws stay in the box and its density integrates to one, bounded Pareto matches its
    closed form, the integer uniform gives -inf outside its support."""
    from smcdet_b200.distributions import DiscreteUniform, TruncatedDiagonalMVN, TruncatedPareto

    torch.manual_seed(0)
    mu = torch.tensor([[-3.95, 5.0], [11.9, 0.0]])
    d = TruncatedDiagonalMVN(mu, torch.tensor(0.1), torch.tensor([-4.0, -4.0]), torch.tensor([12.0, 12.0]))
    x = d.sample()
    assert x.shape == mu.shape and (x >= -4).all() and (x <= 12).all()
    grid = torch.linspace(-4.0, -3.0, 20001).unsqueeze(-1)
    one = TruncatedDiagonalMVN(torch.tensor([-3.95]), torch.tensor(0.1), torch.tensor([-4.0]), torch.tensor([12.0]))
    mass = torch.trapezoid(one.log_prob(grid).exp().squeeze(-1), grid.squeeze(-1))
    assert abs(float(mass) - 1.0) < 1e-3
    with pytest.raises(AssertionError):
        d.log_prob(torch.full_like(mu, 13.0))
    p = TruncatedPareto(0.2, 0.06, 1800.0)
    f = p.sample([2000])
    assert (f >= 0.06 * (1 - 1e-6)).all() and (f <= 1800.0 * (1 + 1e-6)).all()
    a, lo, up = 0.2, 0.06, 1800.0
    ref = np.log(a * lo**a / (1 - (lo / up) ** a)) - (a + 1) * np.log(7.0)
    assert abs(float(p.log_prob(torch.tensor(7.0))) - ref) < 1e-5
    u = DiscreteUniform(2, 5)
    lp = u.log_prob(torch.tensor([1.0, 2.0, 5.0, 6.0]))
    assert torch.isneginf(lp[0]) and torch.isneginf(lp[3]) and abs(float(lp[1]) + np.log(4)) < 1e-6
    s = u.sample([100])
    assert int(s.min()) >= 2 and int(s.max()) <= 5


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference on the host cores) on a tiny setting: exactly one
    JSON line on stdout with the contract's keys; and the own arm refuses to run without a GPU instead of
    falling back."""
    import json
    import subprocess
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--particles", "300", "--mh-iters", "3"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and "workload" in d["config"]
    if not torch.cuda.is_available():
        own = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                             capture_output=True, text=True, timeout=300)
        assert own.returncode != 0 and "no CPU fallback" in (own.stderr + own.stdout)


def test_committed_bench_line_carries_the_contract():
    """The committed 1-GPU bench line of the current round (profiles/r02_bench_n1.json, written by `python bench.py` on a
    B200) has every key of the measurement contract, with consistent arithmetic: value = evals / time, the roofline
    fraction = achieved / peak, traffic measured and below the algorithmic bytes (no wasted re-reads), the dominant
    kernel's share of the step below 1, both CPU baselines with their core counts, clean clocks."""
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.loads(open(os.path.join(root, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "m71synthetic" in d["config"]["workload"]
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["share_of_step"] < 1
    assert 0 < r["executed_frac"] < 1 and r["traffic"] is not None and 0 < r["traffic"] <= r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["also"][0]["kind"] == "reference"
    s = d["strong"]
    assert len(s["checksum"]["sha256_16"]) == 16 and s["e2e"]["d2h_bytes_per_step"] > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/smcdet"), reason="the reference tree only exists in the build container")
def test_public_surface_matches_the_reference():
    """Drop-in check against the unmodified reference (build container only): every class and function the reference's
    library modules define exists here under the same name, constructors and methods take the reference's positional
    arguments in the reference's order (extensions are keyword-only or trail them), and no public method is missing."""
    import importlib
    import inspect

    sys.path.insert(0, "/root/reference")
    try:
        problems = []
        for mod in ("sampler", "prior", "images", "kernel", "aggregate", "metrics", "distributions"):
            ref, own = importlib.import_module("smcdet." + mod), importlib.import_module("smcdet_b200." + mod)
            for name, rv in vars(ref).items():
                if not (inspect.isclass(rv) or inspect.isfunction(rv)) or getattr(rv, "__module__", "") != ref.__name__:
                    continue
                if not hasattr(own, name):
                    problems.append(f"missing {mod}.{name}")
                    continue
                ov = getattr(own, name)

                def positional(fn):
                    return [p.name for p in inspect.signature(fn).parameters.values()
                            if p.name != "self" and p.kind == p.POSITIONAL_OR_KEYWORD]

                if inspect.isfunction(rv):
                    if positional(ov)[:len(positional(rv))] != positional(rv):
                        problems.append(f"signature {mod}.{name}")
                    continue
                for m, member in vars(rv).items():
                    if m.startswith("_") and m != "__init__":
                        continue
                    if not hasattr(ov, m):
                        problems.append(f"missing {mod}.{name}.{m}")
                    elif inspect.isfunction(member) and inspect.isfunction(getattr(ov, m)):
                        a, b = positional(member), positional(getattr(ov, m))
                        if b[:len(a)] != a:
                            problems.append(f"signature {mod}.{name}.{m}: {a} vs {b}")
        assert not problems, problems
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "smcdet" or k.startswith("smcdet.")]:
            del sys.modules[k]
