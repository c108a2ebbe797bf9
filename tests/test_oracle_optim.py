"""CPU tests of the optimizer-side sweep: the oracle restatement against the golden produced by the
unmodified reference classes, and the host logic of the fused optimizer (descriptor packing, hyper-parameters, state
layout, loud failure without a GPU)."""
import ctypes
import os

import pytest
import torch

from dualdiffusion_b200 import _lib as L, ops
from dualdiffusion_b200.training import optim as fo
from oracle import optim_oracle as oo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim_small.pt")


def _load():
    return torch.load(GOLD, weights_only=False)


def run_oracle(g, upto=3):
    """Replays the golden's steps through the oracle; yields (step record, oracle state) per step."""
    names = g["names"]
    params = {n: g["init"][n].clone() for n in names}
    m = {n: torch.zeros_like(params[n]) for n in names}
    v = {n: torch.zeros_like(params[n]) for n in names}
    emas = [{n: g["init"][n].clone().to(torch.float64 if cfg.get("use_float64") else torch.float32) for n in names}
            for cfg in g["emas"].values()]
    betas = [cfg["beta"] for cfg in g["emas"].values()]
    fbs = [cfg.get("feedback_beta") for cfg in g["emas"].values()]
    hy = g["hyper"]
    for it, rec in enumerate(g["steps"][:upto]):
        norm = oo.train_update(params, rec["grads"], m, v, it + 1, lr=hy["lr"], betas=hy["betas"], eps=hy["eps"],
                               weight_decay=rec["weight_decay"], max_norm=hy["max_norm"], emas=emas, ema_betas=betas,
                               feedback_betas=fbs, fan_in=g["fan_in"])
        yield rec, dict(norm=norm, params=params, exp_avg=m, exp_avg_sq=v, emas=emas)


def test_oracle_matches_reference_golden():
    g = _load()
    for rec, st in run_oracle(g):
        assert abs(float(st["norm"]) - float(rec["grad_norm"])) <= 1e-6 * float(rec["grad_norm"])
        for n in g["names"]:
            torch.testing.assert_close(st["params"][n], rec["params"][n], rtol=2e-6, atol=1e-7)
            torch.testing.assert_close(st["exp_avg"][n], rec["exp_avg"][n], rtol=2e-6, atol=1e-9)
            torch.testing.assert_close(st["exp_avg_sq"][n], rec["exp_avg_sq"][n], rtol=2e-6, atol=1e-12)
            for k, name in enumerate(g["emas"]):
                ref = rec["emas"][name][n]
                assert st["emas"][k][n].dtype == ref.dtype
                torch.testing.assert_close(st["emas"][k][n], ref, rtol=2e-6, atol=1e-7)


def test_golden_covers_clipped_and_unclipped_steps_and_weight_norm():
    g = _load()
    norms = [float(s["grad_norm"]) for s in g["steps"]]
    assert norms[0] > g["hyper"]["max_norm"] and norms[2] < g["hyper"]["max_norm"]
    last = g["steps"][-1]["params"]
    for n, f in g["fan_in"].items():
        w = last[n]
        if f > 0:       # re-normalised rows have RMS 1 (up to the 1e-4 epsilon)
            rms = w.reshape(w.shape[0], -1).pow(2).mean(dim=1).sqrt()
            assert torch.allclose(rms, torch.ones_like(rms), atol=2e-4)
    assert g["fan_in"]["free.weight"] == 0 and g["fan_in"]["gain"] == 0


def test_pack_optim_descs_rows_and_prefix_sums():
    p1, p2, p3 = torch.zeros(16, 6, 3, 3), torch.zeros(()), torch.zeros(5000)
    ents = [dict(p=p, g=p, m=p, v=p, emas=[p], fan_in=f) for p, f in ((p1, 54), (p2, 0), (p3, 0))]
    arr, rows = ops.pack_optim_descs(ents)
    assert rows == 16 + 1 + 2
    assert [(d.rows, d.row_len, d.normalize, d.row_begin, d.numel) for d in arr] == [
        (16, 54, 1, 0, 864), (1, ops.OPTIM_FLAT_ROW, 0, 16, 1), (2, ops.OPTIM_FLAT_ROW, 0, 17, 5000)]
    assert arr[0].ema[0] == p1.data_ptr() and not arr[0].ema[1]
    with pytest.raises(ValueError):
        ops.pack_optim_descs([dict(p=p3, g=p3, m=p3, v=p3, fan_in=7)])
    with pytest.raises(ValueError):
        ops.pack_optim_descs([dict(p=p3, g=p3, m=p3, v=p3, emas=[p3] * 5)])
    garr, chunks = ops.pack_gnorm_descs([p1, p2, torch.zeros(3 * L.GNORM_CHUNK + 1)])
    assert chunks == 1 + 1 + 4 and [d.chunk_begin for d in garr] == [0, 1, 2]


def test_struct_sizes_match_header_layout():
    # natural alignment of the C structs in include/dualdiffusion_b200.h (pointers 8, long long 8, double 8, int 4)
    assert ctypes.sizeof(L.GnormDesc) == 24
    assert ctypes.sizeof(L.OptimDesc) == 4 * 8 + 4 * 8 + 8 + 4 * 4
    assert ctypes.sizeof(L.OptimHyper) == 7 * 8 + 4 * 8 + 4 * 8 + 4 * 4 + 8
    assert L.OptimDesc.numel.offset == 64 and L.OptimHyper.n_ema.offset == 136


def test_make_hyper_bias_corrections_and_feedback_encoding():
    h = fo.make_hyper(1e-2, (0.9, 0.99), 1e-8, 0.0, 3.0, [0.9999, 0.99], [0.9999, None], [0, 1])
    assert h.bias_correction1 == pytest.approx(1 - 0.9 ** 3) and h.bias_correction2 == pytest.approx(1 - 0.99 ** 3)
    assert h.n_ema == 2 and h.feedback_beta[0] == 0.9999 and h.feedback_beta[1] == -1.0 and h.feedback_beta[3] == -1.0
    assert list(h.ema_is_f64) == [0, 1, 0, 0]


def test_fused_adamw_is_a_torch_optimizer_with_adamw_state_layout_and_no_cpu_path():
    w = torch.nn.Parameter(torch.randn(4, 3))
    opt = fo.FusedAdamW([w], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0, fused=True)
    assert isinstance(opt, torch.optim.Optimizer)
    assert set(opt.param_groups[0]) >= {"lr", "betas", "eps", "weight_decay", "params"}
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.5)      # trainer.py:42 drives the lr through the group
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-3)
    w.grad = torch.randn(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.clip_grad_norm_(10.0)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.randn(4, 3))], lr=1e-2)
    ref.param_groups[0]["params"][0].grad = torch.randn(4, 3)
    ref.step()
    sd = ref.state_dict()
    opt.load_state_dict({"state": sd["state"], "param_groups": [dict(opt.state_dict()["param_groups"][0])]})
    assert set(opt.state[w]) == {"step", "exp_avg", "exp_avg_sq"}
    with pytest.raises(NotImplementedError):
        fo.FusedAdamW([w], amsgrad=True)
    del sched


def test_weight_norm_fan_in_follows_normalize_weights_membership():
    from dualdiffusion_b200.modules.mp_tools import MPConv

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = MPConv(12, 16, (3, 3), groups=2)
            self.b = MPConv(9, 5, (), disable_weight_norm=True)
            self.gain = torch.nn.Parameter(torch.zeros(()))
    net = Net()
    f = fo.weight_norm_fan_in(net)
    assert f == {id(net.a.weight): 54}
    opt = fo.FusedAdamW(net.parameters())
    opt.attach_module(net)
    with pytest.raises(ValueError):
        opt.attach_emas([[net.a.weight]], [0.99])
    opt.attach_emas([[p.detach().clone() for p in net.parameters()]], [0.99], [0.999])
    opt.set_ema_betas([0.5])
    assert opt._ema_feedback == [0.999]


def test_descriptor_table_covers_every_unet_parameter_once(monkeypatch):
    """Host dry run of FusedAdamW.step() on the reduced UNet (kernel launch replaced by a recorder): every parameter
    element is covered by exactly one row, weight-normalised tensors are exactly the MPConv weights normalize_weights()
    touches, and the EMA pointers follow the parameter order."""
    from dualdiffusion_b200.modules.mp_tools import MPConv
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    from oracle import unet_oracle as uo
    spec = uo.small_spec()
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    params = list(net.parameters())
    for p in params:
        p.grad = torch.zeros_like(p)
    emas = [[p.detach().clone() for p in params]]
    calls = []
    monkeypatch.setattr(fo, "_require_f32_cuda", lambda t, what: None)
    monkeypatch.setattr(ops, "descs_to_device", lambda arr, dev: arr)
    monkeypatch.setattr(ops, "optim_step_batched", lambda descs, n, rows, hyper, coef: calls.append((descs, n, rows, hyper)))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    opt = fo.FusedAdamW(params, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.0)
    opt.attach_module(net)
    opt.attach_emas(emas, [0.999], [0.9999])
    opt.step()
    monkeypatch.undo()
    assert len(calls) == 1
    descs, n, rows, hyper = calls[0]
    assert n == len(params) and hyper.n_ema == 1 and hyper.bias_correction1 == pytest.approx(0.1)
    normed = {c.weight.data_ptr() for c in net.modules() if isinstance(c, MPConv) and not c.disable_weight_norm}
    begin = 0
    for d, p, e in zip(descs, params, emas[0]):
        assert d.p == p.data_ptr() and d.g == p.grad.data_ptr() and d.ema[0] == e.data_ptr() and not d.ema[1]
        assert d.m == opt.state[p]["exp_avg"].data_ptr() and d.v == opt.state[p]["exp_avg_sq"].data_ptr()
        assert d.numel == p.numel() and d.row_begin == begin
        assert bool(d.normalize) == (p.data_ptr() in normed)
        if d.normalize:
            assert d.rows == p.shape[0] and d.rows * d.row_len == p.numel()
        else:
            assert (d.rows - 1) * d.row_len < p.numel() <= d.rows * d.row_len
        begin += d.rows
    assert begin == rows
    assert all(float(opt.state[p]["step"]) == 1.0 for p in params)


@pytest.fixture(scope="module")
def host_harness(tmp_path_factory):
    """tests/csrc/optim_host_check.cpp built with g++: the kernels' per-element functions (csrc/optim_math.cuh) and
    descriptor walk on host memory."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path_factory.mktemp("optim_host") / "liboptim_host_check.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(root, "include"),
                    "-I", os.path.join(root, "dualdiffusion_b200", "csrc"), "-x", "c++",
                    os.path.join(root, "tests", "csrc", "optim_host_check.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.optim_host_grad_norm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    lib.optim_host_grad_norm.restype = None
    lib.optim_host_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(L.OptimHyper),
                                    ctypes.c_float]
    lib.optim_host_step.restype = ctypes.c_int
    return lib


def test_kernel_arithmetic_and_descriptor_walk_replay_reference_golden_on_host(host_harness):
    """The exact per-element code the CUDA kernels compile (optim_math.cuh), driven by ops.pack_optim_descs /
    pack_gnorm_descs / make_hyper on host tensors, replays the reference golden: clip, AdamW with and without decay,
    fp32 + fp64 EMA copies, feedback, weight re-normalisation, flat tensors with a short tail row."""
    g = _load()
    names = g["names"]
    p = {n: g["init"][n].clone().contiguous() for n in names}
    m = {n: torch.zeros_like(p[n]) for n in names}
    v = {n: torch.zeros_like(p[n]) for n in names}
    cfgs = list(g["emas"].values())
    emas = [{n: g["init"][n].clone().to(torch.float64 if c.get("use_float64") else torch.float32) for n in names}
            for c in cfgs]
    hy = g["hyper"]
    for it, rec in enumerate(g["steps"]):
        grads = [rec["grads"][n].contiguous() for n in names]
        garr, _ = ops.pack_gnorm_descs(grads)
        out2 = (ctypes.c_float * 2)()
        host_harness.optim_host_grad_norm(garr, len(grads), hy["max_norm"], out2)
        assert abs(out2[0] - float(rec["grad_norm"])) <= 2e-6 * float(rec["grad_norm"])
        assert out2[1] == pytest.approx(min(1.0, hy["max_norm"] / (float(rec["grad_norm"]) + 1e-6)), rel=1e-5)
        arr, rows = ops.pack_optim_descs([dict(p=p[n], g=gr, m=m[n], v=v[n], emas=[e[n] for e in emas],
                                               fan_in=g["fan_in"][n]) for n, gr in zip(names, grads)])
        hyper = fo.make_hyper(hy["lr"], hy["betas"], hy["eps"], rec["weight_decay"], float(it + 1),
                              [c["beta"] for c in cfgs], [c.get("feedback_beta") for c in cfgs],
                              [int(bool(c.get("use_float64"))) for c in cfgs])
        assert host_harness.optim_host_step(arr, len(names), rows, ctypes.byref(hyper), out2[1]) == 0
        for n in names:
            torch.testing.assert_close(p[n], rec["params"][n], rtol=1e-5, atol=1e-7, msg=f"step {it} param {n}")
            torch.testing.assert_close(m[n], rec["exp_avg"][n], rtol=1e-5, atol=2e-8, msg=f"step {it} exp_avg {n}")
            torch.testing.assert_close(v[n], rec["exp_avg_sq"][n], rtol=1e-5, atol=1e-12)
            for k, name in enumerate(g["emas"]):
                torch.testing.assert_close(emas[k][n], rec["emas"][name][n], rtol=1e-5, atol=1e-7,
                                           msg=f"step {it} ema {name} {n}")


def test_parameter_without_gradient_gets_ema_and_renormalisation_only(host_harness):
    """g == NULL in the descriptor: torch's AdamW skips the parameter, EMA_Manager.update() and normalize_weights()
    still cover it (ema.py:292 takes every module parameter)."""
    gen = torch.Generator().manual_seed(3)
    p = torch.randn(6, 20, generator=gen)
    e = torch.randn(6, 20, generator=gen)
    exp_e = torch.lerp(e, p, 1 - 0.99)
    exp_p = oo.normalize_rows(torch.lerp(p, exp_e, 1 - 0.999))
    arr, rows = ops.pack_optim_descs([dict(p=p, g=None, m=None, v=None, emas=[e], fan_in=20)])
    assert not arr[0].g and not arr[0].m and not arr[0].v
    hyper = fo.make_hyper(1e-2, (0.9, 0.99), 1e-8, 0.1, 1.0, [0.99], [0.999], [0])
    assert host_harness.optim_host_step(arr, 1, rows, ctypes.byref(hyper), 1.0) == 0
    torch.testing.assert_close(e, exp_e, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(p, exp_p, rtol=1e-6, atol=1e-7)
