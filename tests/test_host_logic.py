"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares,
host-side packing / schedule / sharding logic, and loud failure without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import gdss_oracle as O
from oracle.cases import CKPTS, SHORT_CASES, load_golden_ckpt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_CUDA = torch.cuda.is_available()


def test_library_exports_every_symbol_of_the_header():
    from gdss_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "gdss_b200.h")).read()
    names = set(re.findall(r"\b(gdss_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/gdss_b200.h but not exported"
    assert set(L._gdss_symbols) == names, set(L._gdss_symbols) ^ names
    assert L.gdss_abi_version() == 1
    assert L.gdss_device_count() >= 0
    assert b"workspace" in L.gdss_error_string(-3)


def test_tf32_variant_library_exports_the_same_abi():
    """libgdss_b200_tf32.so (GDSS_PRECISION=tf32) is the same C ABI built with single-pass TF32 contractions."""
    import ctypes
    from gdss_b200 import _lib
    assert _lib.PRECISION == "fp32"                         # the default library is the fp32-parity build
    alt = ctypes.CDLL(os.path.join(ROOT, "gdss_b200", "csrc", "libgdss_b200_tf32.so"))
    hdr = open(os.path.join(ROOT, "include", "gdss_b200.h")).read()
    for n in sorted(set(re.findall(r"\b(gdss_[a-z0-9_]+)\s*\(", hdr))):
        assert hasattr(alt, n), n
    assert alt.gdss_abi_version() == 1


def test_struct_sizes_match_the_header():
    from gdss_b200 import _lib
    assert C.sizeof(_lib.NetDesc) == 48
    assert C.sizeof(_lib.SdeDesc) == 24
    assert C.sizeof(_lib.SamplerDesc) == 24 * 2 + 8 + 24 + 16
    assert _lib.STEP_COEF_FLOATS == 24 and max(_lib.COL.values()) < 24


@pytest.mark.parametrize("name", list(CKPTS))
def test_describe_and_pack_every_checkpoint(name):
    from gdss_b200 import _lib
    from gdss_b200.engine import describe_network, pack_network
    ck = load_golden_ckpt(name)
    N = ck["params_adj"]["max_node_num"]
    for sd, params in ((ck["x_state_dict"], ck["params_x"]), (ck["adj_state_dict"], ck["params_adj"])):
        d = describe_network(sd, N)
        kind = {"ScoreNetworkX": _lib.NET_X, "ScoreNetworkX_GMH": _lib.NET_X_GMH, "ScoreNetworkA": _lib.NET_A}[params["model_type"]]
        assert d.model_type == kind
        assert d.max_feat_num == params["max_feat_num"] and d.nhid == params["nhid"]
        assert d.depth == params.get("depth", params.get("num_layers"))
        if kind != _lib.NET_X:
            for k in ("num_linears", "c_init", "c_hid", "c_final", "adim"):
                assert getattr(d, k) == params[k], k
        desc, blob = pack_network(sd, N)
        n_param = sum(v.numel() for v in sd.values())
        assert n_param <= blob.numel() <= 1.35 * n_param + 64 * len(sd)      # rows / strides are padded to 4
        # every parameter value survives packing (multiset of values, transposition aside)
        src = torch.cat([v.reshape(-1) for v in sd.values()]).double()
        assert abs(float(blob.double().sum()) - float(src.sum())) <= 1e-6 * float(src.abs().sum())
        assert abs(float((blob.double() ** 2).sum()) - float((src ** 2).sum())) <= 1e-9 * float((src ** 2).sum())
        # wrong tensor count / wrong shape are rejected
        bad = dict(list(sd.items())[:-1])
        with pytest.raises(Exception):
            pack_network(bad, N, desc)


def test_linear_weights_are_stored_transposed():
    from gdss_b200.engine import pack_network
    ck = load_golden_ckpt("qm9")
    sd = ck["x_state_dict"]             # ScoreNetworkX: gcn weights first, then final.linears
    _, blob = pack_network(sd, 9)
    w0 = sd["layers.0.weight"]          # [in, out] stays as is at offset 0
    assert torch.equal(blob[: w0.numel()].view_as(w0), w0)
    wl = sd["final.linears.2.weight"]   # [out, in] -> [in, out]
    tail = blob[-(wl.numel() + 4 + 0):]
    flat = wl.t().contiguous().view(-1)
    s = blob.unfold(0, flat.numel(), 1)
    assert any(torch.equal(row, flat) for row in s[-64:])


def test_models_load_reference_checkpoints_unchanged():
    from gdss_b200.models import load_model
    for name in CKPTS:
        ck = load_golden_ckpt(name)
        for params, sd in ((ck["params_x"], ck["x_state_dict"]), (ck["params_adj"], ck["adj_state_dict"])):
            m = load_model(params)
            assert list(m.state_dict().keys()) == list(sd.keys())
            m.load_state_dict(sd, strict=True)
    with pytest.raises(ValueError, match="Unknown"):
        load_model({"model_type": "Nope"})


def _oracle_coeffs(case, i):
    sx = O.SDESpec(*case["sde_x"], case["scales"])
    sa = O.SDESpec(*case["sde_adj"], case["scales"])
    t = torch.linspace(1, 1e-4, case["scales"])[i]
    return sx, sa, t


@pytest.mark.parametrize("case", SHORT_CASES, ids=[c["name"] for c in SHORT_CASES])
def test_schedule_table_matches_oracle_scalars(case):
    """Apply the folded affine coefficients to random scalars and compare with the oracle's own
    step algebra evaluated on the same numbers."""
    from gdss_b200._lib import COL
    from gdss_b200.schedule import build_table
    from tests.gpu_common import sde_from
    sdx, sda = sde_from(case["sde_x"], case["scales"]), sde_from(case["sde_adj"], case["scales"])
    tab = build_table(sdx, sda, case["predictor"], case["corrector"], case["snr"], case["scale_eps"], 1e-4)
    assert tab.shape == (case["scales"], 24) and tab.dtype == np.float32
    rs = np.random.RandomState(0)
    for i in (0, case["scales"] // 2, case["scales"] - 1):
        sx, sa, t = _oracle_coeffs(case, i)
        for d, sd in enumerate((sx, sa)):
            y, raw, z = [torch.tensor(float(v)) for v in rs.standard_normal(3)]
            row = tab[i]
            cs = -1.0 / sd.marginal_std(t) if sd.kind != "VE" else torch.ones(())
            assert np.isclose(row[COL["cs"] + d], float(cs), rtol=1e-6)
            s = cs * raw
            if case["predictor"] == "S4":
                hdt = torch.ones(()) * (-0.5 / case["scales"])
                m1, s1 = sd.transition(t, hdt)
                m2, s2 = sd.transition(t + hdt, hdt)
                want = m1 * y + s1 * z
                want = want + (-sd.diffusion(t) ** 2 * s) * (-1.0 / case["scales"])
                want = m2 * want
                got = row[COL["mu1"] + d] * y + row[COL["sig1"] + d] * z
                got = got + row[COL["drift"] + d] * raw
                got = row[COL["mu2"] + d] * got
                assert np.isclose(row[COL["sig2"] + d], float(s2), rtol=1e-6, equal_nan=True)
            elif case["predictor"] == "Reverse":
                f, G = sd.discretize(y, t)
                want = y - (f - G ** 2 * s) + G * z
                got = row[COL["pa"] + d] * y + row[COL["pb"] + d] * raw + row[COL["pc"] + d] * z
            else:
                dt = -1.0 / sd.N
                g = sd.diffusion(t)
                want = y + (sd.drift_coeff(t) * y - g ** 2 * s) * dt + g * np.sqrt(-dt) * z
                got = row[COL["pa"] + d] * y + row[COL["pb"] + d] * raw + row[COL["pc"] + d] * z
            assert abs(float(got) - float(want)) <= 2e-6 * max(1.0, abs(float(want))), (i, d, float(got), float(want))
            if case["corrector"] == "Langevin":
                assert np.isclose(row[COL["alpha"] + d], float(sd.langevin_alpha(t, False)), rtol=0, atol=0)


def test_unknown_sde_and_predictor_behave_like_the_reference():
    from gdss_b200.sde import load_sde, VPSDE
    from gdss_b200.schedule import build_table
    with pytest.raises(NotImplementedError, match="not yet supported"):
        load_sde(dict(type="XX", beta_min=.1, beta_max=1., num_scales=10))
    sde = VPSDE(.1, 1., 10)
    a = build_table(sde, sde, "Bogus", "Bogus", .1, 1., 1e-3)      # silently Euler / None (solver.py:163-164)
    b = build_table(sde, sde, "Euler", "None", .1, 1., 1e-3)
    assert np.array_equal(a, b)


def test_shard_bounds_cover_the_batch():
    from gdss_b200.dist import shard_bounds
    for total in (1, 7, 8, 10000, 10001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("name,fusable", [("community_small", 1), ("ego_small", 1), ("qm9", 1), ("zinc250k", 1),
                                          ("zinc250k_v2", 1), ("enzymes", 0), ("grid", 0)])
def test_ctx_query_reports_the_plan_of_every_checkpoint(name, fusable):
    """Planning is host logic: a context can be created (not run) without a device, and gdss_ctx_query tells which networks
    fit the per-graph kernel and how much shared memory its staged-weights plan needs."""
    from gdss_b200._lib import lib
    from gdss_b200.engine import pack_network
    L = lib()
    ck = load_golden_ckpt(name)
    N = int(ck["params_adj"]["max_node_num"])
    dx, bx = pack_network(ck["x_state_dict"], N)
    da, ba = pack_network(ck["adj_state_dict"], N)
    ctx = L.gdss_ctx_create(N, int(dx.max_feat_num), C.byref(dx), bx.data_ptr(), C.byref(da), ba.data_ptr())
    assert ctx
    try:
        q = lambda k: int(L.gdss_ctx_query(ctx, k.encode()))
        assert q("fused_fusable") == fusable
        smem = q("fused_smem_bytes")
        assert (smem > 0) == bool(fusable)
        if name == "zinc250k_v2":
            assert smem <= 227 * 1024                      # one 512-thread CTA of the largest graphs fits an SM
            half = (227 * 1024) // 2 - 1024                # two CTAs per SM: n <= 23 with staged weights, n <= 32 without
            assert q("bucket_smem_bytes:23:1") <= half < q("bucket_smem_bytes:24:1")
            assert q("bucket_smem_bytes:32:0") <= half < q("bucket_smem_bytes:32:1")
            assert q("bucket_smem_bytes:99:0") < 0
        if name == "community_small":
            assert smem > 227 * 1024                       # 96 KB of Q|K|V weights per layer: runs the weight-less plan
        assert q("bogus") < 0
        if not HAS_CUDA:
            assert q("smem_optin") == 48 * 1024            # no device: the default limit, nothing is launched
    finally:
        L.gdss_ctx_destroy(ctx)


@pytest.mark.skipif(HAS_CUDA, reason="CPU-only behaviour")
def test_compute_calls_fail_loudly_without_cuda():
    from gdss_b200 import _lib
    from gdss_b200.engine import Engine
    from gdss_b200.solver import get_pc_sampler
    from gdss_b200.sde import VPSDE
    from gdss_b200.models import load_model
    from gdss_b200 import graph_utils as G
    ck = load_golden_ckpt("qm9")
    with pytest.raises(_lib.GdssError, match="no CPU fallback"):
        Engine(ck["x_state_dict"], ck["adj_state_dict"], 9, "cuda")
    fn = get_pc_sampler(VPSDE(.1, 1., 10), VPSDE(.1, 1., 10), (2, 9, 4), (2, 9, 9), device="cpu")
    with pytest.raises(Exception):
        fn(ck["x_state_dict"], ck["adj_state_dict"], torch.ones(2, 9))
    m = load_model(ck["params_adj"])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 9, 4), torch.zeros(1, 9, 9), torch.ones(1, 9))
    with pytest.raises(RuntimeError, match="CUDA"):
        G.mask_x(torch.zeros(1, 9, 4), torch.ones(1, 9))
    from gdss_b200.optim import FusedAdamEMA
    with pytest.raises(_lib.GdssError, match="no CPU fallback"):
        FusedAdamEMA(torch.zeros(8))
    # the C entry points themselves refuse, they do not silently compute on the host
    L = _lib.lib()
    assert L.gdss_quantize(C.c_void_p(8), C.c_void_p(8), 0.5, 4, None) == -2
    assert L.gdss_adam_ema_step(C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), None, 4, 1e-3, .9, .999, 1e-8, 0., 1.,
                                1., 1, .999, 0, None, C.c_void_p(8), None, None) == -2


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["GDSS_ROOT"])
from gdss_b200.dist import ShardInfo, shard_bounds
dist.init_process_group("gloo")
s = ShardInfo(create_comm=False)
ident = s.exchange_id(lambda: bytes(range(128)))
assert ident == bytes(range(128)), ident[:8]
lo, hi = s.bounds(11)
# the 4-float Langevin exchange: every rank contributes its shard's norm sums
local = torch.tensor([float(hi - lo), 1.0, 2.0 * s.rank, 3.0])
dist.all_reduce(local)
assert local.tolist() == [11.0, 2.0, 2.0, 6.0], local
open(os.path.join(os.environ["GDSS_OUT"], f"rank{s.rank}.txt"), "w").write(f"ok {lo} {hi}")
"""


def test_two_rank_gloo_sharding_and_id_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GDSS_ROOT=ROOT, GDSS_OUT=str(tmp_path), MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (tmp_path / "rank0.txt").read_text() == "ok 0 6" and (tmp_path / "rank1.txt").read_text() == "ok 6 11"


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU oracle port timed on the host cores) prints one JSON line with the keys the driver
    reads; bounded sample (1 timed step of 1000 on 32 graphs)."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "graphs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["metric"].startswith("sampled graphs/sec")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "ZINC250k v2" in line["config"]["workload"] and line["dtype"] == "f32"


def test_clock_sampler_parses_single_and_multi_gpu_rows():
    """bench.ClockSampler.stop(): medians per GPU, the slowest clock / highest power / any throttle reason over the GPUs of
    the job (one poller on rank 0 for all of them), from canned nvidia-smi rows."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    def make(primary, n, rows):
        cs = bench.ClockSampler.__new__(bench.ClockSampler)
        cs.rows, cs.proc, cs.primary, cs.n_gpus = [[c.strip() for c in r.split(",")] for r in rows], FakeProc(), primary, n
        return cs.stop()

    na = "Not Active"
    one = make(0, 1, [f"0, 1965, 1965, 350.5, {na}, {na}, {na}, {na}", f"0, 1950, 1965, 371.0, {na}, {na}, {na}, Active", "garbage"])
    assert one["sm_mhz"] == 1957.5 and one["sm_max_mhz"] == 1965.0 and one["power_w_max"] == 371.0 and one["samples"] == 2
    assert one["reasons"] == ["sw_power_cap"] and "sm_mhz_min_over_gpus" not in one
    rows = []
    for g, (mhz, pw, therm) in enumerate([(1965, 400.0, na), (1800, 410.5, "Active"), (1965, 390.0, na)]):
        rows += [f"{g}, {mhz}, 1965, {pw}, {na}, {na}, {therm}, {na}"] * 2
    many = make(0, 3, rows)
    assert many["sm_mhz"] == 1965.0 and many["reasons"] == [] and many["gpus_sampled"] == 3
    assert many["sm_mhz_min_over_gpus"] == 1800.0 and many["power_w_max_over_gpus"] == 410.5
    assert many["reasons_any_gpu"] == ["sw_thermal_slowdown"]


def test_dead_chain_analysis_of_the_shipped_checkpoints():
    """gdss_analyze_weights (host code, what gdss_ctx_create switches off): the weight-decayed dead sub-networks of the shipped
    checkpoints -- e.g. gdss_zinc250k_v2's ScoreNetworkA: every gnn_q / gnn_k of layer 0 and the whole last AttentionLayer are in
    the fp32 denormal range -- are found, liveness of x_out is propagated backwards, random-init weights lose nothing."""
    import ctypes as C
    from gdss_b200 import _lib
    from gdss_b200.engine import pack_network, model_state
    L = _lib.lib()

    def report(sd, N):
        desc, blob = pack_network(model_state(sd), N)
        out = (C.c_int32 * 64)()
        n = L.gdss_analyze_weights(C.byref(desc), N, blob.data_ptr(), out, 16)
        assert n >= 0
        return n, [tuple(out[4 * l + k] for k in range(4)) for l in range(desc.depth)]

    ck = load_golden_ckpt("zinc250k_v2")
    n, rep = report(ck["adj_state_dict"], 38)
    assert n > 0
    assert rep[0][0] == 0 and rep[0][1] == 0b11                     # layer 0: both Q/K chains dead, both V live
    assert [bin(r[0]).count("1") for r in rep[1:5]] == [6, 6, 6, 6]  # layers 1-4: 6 of 8 maps live
    assert rep[5] == (0, 0, 0, 1)                                   # last layer: constant per-edge MLP, nothing else
    assert rep[4][1] == 0 and rep[4][2] == 0                        # its input x is never read: layer 4 needs no V / node MLP
    assert rep[3][2] == 1
    n_x, rep_x = report(ck["x_state_dict"], 38)
    assert rep_x[0][0] == 0 and rep_x[0][2] == 1 and rep_x[1][2] == 1    # X_GMH: every x_l feeds the final concat
    # a dead block must be dead in the reference's sense too: removing it changes nothing (oracle, CPU)
    sd = {k: v.clone() for k, v in ck["adj_state_dict"].items()}
    for k in sd:
        if k.startswith("layers.5."):
            assert float(sd[k].abs().max()) < 1e-30 or ".mlp.linears." in k or ".multi_channel." in k or "gnn_v" in k, k
    # random-init weights: nothing is dead
    from gdss_b200.models import load_model
    torch.manual_seed(0)
    n_r, rep_r = report(load_model(ck["params_adj"]).state_dict(), 38)
    assert n_r == 0 and all(r[0] == r[1] and r[3] == 0 for r in rep_r[:-1]) and rep_r[-1][1] == 0b11111111 or n_r == 0


def test_state_dict_is_packed_by_key_not_by_position():
    """A re-ordered (sorted) state_dict packs to the same blob; a missing / foreign key fails loudly."""
    from gdss_b200 import _lib
    from gdss_b200.engine import pack_network
    ck = load_golden_ckpt("qm9")
    sd = ck["adj_state_dict"]
    _, a = pack_network(sd, 9)
    _, b = pack_network({k: sd[k] for k in sorted(sd)}, 9)
    assert torch.equal(a, b)
    bad = dict(sd)
    bad.pop("final.linears.2.bias")
    with pytest.raises(_lib.GdssError):
        pack_network(bad, 9)


def test_probability_flow_and_n_steps_in_the_schedule():
    """sde.py:98-99: the ODE halves the reverse drift and drops the predictor noise; Euler + probability_flow fails in the
    reference at call time (solver.py:52) and so does the mirror."""
    from gdss_b200.schedule import build_table, COL
    from gdss_b200.sde import VPSDE, VESDE
    from gdss_b200.solver import get_pc_sampler
    sx, sa = VPSDE(.1, 1., 50), VESDE(.2, 1., 50)
    t0 = build_table(sx, sa, "Reverse", "Langevin", .2, .8, 1e-4, False)
    t1 = build_table(sx, sa, "Reverse", "Langevin", .2, .8, 1e-4, True)
    assert np.allclose(t1[:, COL["pb"]:COL["pb"] + 2], 0.5 * t0[:, COL["pb"]:COL["pb"] + 2], rtol=1e-6)
    assert np.all(t1[:, COL["pc"]:COL["pc"] + 2] == 0) and np.all(t0[:, COL["pc"]:COL["pc"] + 2] > 0)
    assert np.array_equal(t1[:, COL["pa"]:COL["pa"] + 2], t0[:, COL["pa"]:COL["pa"] + 2])
    fn = get_pc_sampler(sx, sa, (2, 9, 4), (2, 9, 9), predictor="Euler", probability_flow=True, device="cpu")
    with pytest.raises(TypeError, match="not subscriptable"):
        fn({}, {}, torch.ones(2, 9))
    with pytest.raises(ValueError):
        get_pc_sampler(sx, sa, (2, 9, 4), (2, 9, 9), predictor="Reverse", corrector="Langevin", n_steps=0)


def test_adam_step_constants_follow_torch_and_the_reference_ema():
    """gdss_adam_step_consts (host code; what a graph-replayed training step copies to the device before every replay): lr, the
    two Adam bias corrections as torch.optim.Adam computes them from the python step count, and the EMA factor of
    utils/ema.py:26-28 (num_updates += 1; decay = min(decay, (1 + num_updates) / (10 + num_updates)))."""
    import ctypes as C
    import math
    from gdss_b200 import _lib
    L = _lib.lib()
    out = (C.c_float * 4)()
    for step, nu in ((1, 0), (2, 1), (10, 9), (1000, 999), (200000, 5)):
        rc = L.gdss_adam_step_consts(0.005, 0.9, 0.999, step, 0.999, nu, C.cast(out, C.c_void_p))
        assert rc == 0
        n1 = nu + 1
        decay = min(0.999, (1 + n1) / (10 + n1))
        b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))       # the ABI takes the betas as C floats
        want = [0.005, 1 - b1 ** step, math.sqrt(1 - b2 ** step), 1 - decay]
        for a, b in zip(out, want):
            assert abs(a - b) <= 2e-7 * max(1.0, abs(b)), (step, nu, list(out), want)
    rc = L.gdss_adam_step_consts(0.005, 0.9, 0.999, 3, 0.5, -1, C.cast(out, C.c_void_p))      # -1: the decay as it is
    assert rc == 0 and abs(out[3] - 0.5) < 1e-7
    assert L.gdss_adam_step_consts(0.005, 0.9, 0.999, 0, 0.999, 0, C.cast(out, C.c_void_p)) < 0       # step is 1-based


def test_bench_secondary_ceiling_follows_baseline_md():
    """the MUFU ceiling of C3 is ~2.4 k graphs/s per GPU (306 k tanh + 646 k ELU elements per joint
    evaluation, two evaluations per step, 148 SMs x 16 results per clock at 1965 MHz); the reported fraction is the achieved
    dense-equivalent element rate against it and scales with the number of GPUs."""
    import bench
    c = bench.mufu_ceiling(bench.WORKLOADS["zinc250k_v2"], 1214.07, 1965.0)
    assert abs(c["ceiling_graphs_per_s"] - 2443.9) < 1 and abs(c["frac"] - 1214.07 / c["ceiling_graphs_per_s"]) < 1e-12
    c8 = bench.mufu_ceiling(bench.WORKLOADS["zinc250k_v2"], 8000.0, 1965.0, n_gpus=8)
    assert abs(c8["ceiling_graphs_per_s"] - 8 * c["ceiling_graphs_per_s"]) < 1e-6
    for wl in bench.WORKLOADS.values():
        assert len(wl["transcendentals"]) == 2


def test_load_sampling_fn_forwards_the_reference_configs(monkeypatch):
    """utils/loader.py:121-147: the sampler closure is built from (config_train, config.sampler, config.sample, device) -- here
    the checkpoint's own `model_config` (plain dict or attribute objects, as the reference's EasyDict gives them) and the
    sampler / sample blocks of config/sample_zinc250k.yaml and config/sample_qm9.yaml: 10 000-graph shapes for the molecule sets,
    the data batch size otherwise, S4_solver for predictor 'S4', `cuda:<first id>` for a device list."""
    from types import SimpleNamespace
    from gdss_b200 import solver
    from gdss_b200.ckpt import load_ckpt
    seen = []
    monkeypatch.setattr(solver, "_make", lambda *a, **k: seen.append((a, k)) or "closure")

    def ns(d):
        return SimpleNamespace(**{k: ns(v) if isinstance(v, dict) else v for k, v in d.items()})

    sample = dict(use_ema=False, noise_removal=True, probability_flow=False, eps=1e-4, seed=42)
    zinc = load_ckpt("zinc250k_v2")["model_config"]
    module = dict(predictor="Reverse", corrector="Langevin", snr=0.2, scale_eps=0.9, n_steps=1)
    for cfg, mod, smp in ((zinc, module, sample), (ns(zinc), ns(module), ns(sample))):
        assert solver.load_sampling_fn(cfg, mod, smp, [3, 4]) == "closure"
        a, k = seen.pop()
        sde_x, sde_adj, shape_x, shape_adj, pred, corr, snr, seps, n_steps, pf, cont, denoise, eps, dev, s4 = a
        assert (sde_x.kind, sde_x.beta_0, sde_x.beta_1, sde_x.N) == ("VP", 0.1, 1.0, 1000)
        assert (sde_adj.kind, sde_adj.sigma_min, sde_adj.sigma_max, sde_adj.N) == ("VE", 0.2, 1.0, 1000)
        assert shape_x == (10000, 38, 9) and shape_adj == (10000, 38, 38)
        assert (pred, corr, snr, seps, n_steps, pf, cont, denoise, eps, dev, s4) == \
            ("Reverse", "Langevin", 0.2, 0.9, 1, False, True, True, 1e-4, "cuda:3", False)
    qm9 = load_ckpt("qm9")["model_config"]
    assert solver.load_sampling_fn(qm9, dict(predictor="S4", corrector="None", snr=0.2, scale_eps=0.7, n_steps=1), sample, "cuda:0")
    a, _ = seen.pop()
    assert a[2] == (10000, 9, 4) and a[-1] is True and a[-2] == "cuda:0"
    com = load_ckpt("community_small")["model_config"]
    assert solver.load_sampling_fn(com, dict(predictor="Euler", corrector="Langevin", snr=0.05, scale_eps=0.7, n_steps=1), sample, "cuda")
    a, _ = seen.pop()
    assert a[2] == (com["data"]["batch_size"], 20, 10) and a[3] == (com["data"]["batch_size"], 20, 20) and a[-1] is False


def test_sampling_weights_apply_the_ema_shadow_like_the_reference_sampler():
    """sampler.py:47-52 / utils/ema.py:37-47: with `sample.use_ema` the i-th parameter is overwritten by the i-th shadow
    parameter (ENZYMES, grid ship EMA and ask for it; QM9 has none); a shadow list that does not fit the state_dict is refused."""
    from gdss_b200.ckpt import ema_weights, load_ckpt, sampling_weights
    ck = load_ckpt("enzymes")
    sdx, px, sda, pa = sampling_weights(ck)                       # the checkpoint's own config asks for EMA
    raw_x, _, raw_a, _ = sampling_weights(ck, use_ema=False)
    assert list(sdx) == list(raw_x) and list(sda) == list(raw_a) and px["model_type"] == "ScoreNetworkX"
    for sd, raw, ema in ((sdx, raw_x, ck["ema_x"]), (sda, raw_a, ck["ema_adj"])):
        for (k, v), s in zip(sd.items(), ema["shadow_params"]):
            assert torch.equal(v, s.float()) and v.shape == raw[k].shape
        assert max(float((sd[k] - raw[k]).abs().max()) for k in sd) > 0          # the shadow differs from the raw weights
    q = load_ckpt("qm9")
    assert all(torch.equal(a, b) for a, b in zip(sampling_weights(q)[0].values(), q["x_state_dict"].values()))
    with pytest.raises(ValueError):
        ema_weights(raw_x, {"shadow_params": ck["ema_x"]["shadow_params"][:-1]})
    bad = list(ck["ema_x"]["shadow_params"])
    bad[0] = bad[0].reshape(-1)
    with pytest.raises(ValueError):
        ema_weights(raw_x, {"shadow_params": bad})
