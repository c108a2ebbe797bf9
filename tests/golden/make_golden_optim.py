"""Regenerates tests/golden/optim_small.pt from the UNMODIFIED reference (CPU, build container only): the optimizer-side
sweep the reference trainer runs after backward() -- clip_grad_norm_ (training/trainer.py:1044), torch.optim.AdamW.step
(trainer.py:461-473,1062), EMA_Manager.update (training/ema.py:284-313, the real class driven by a stand-in trainer
object) and DualDiffusionModule.normalize_weights (modules/module.py:186-191 -> MPConv.normalize_weights) -- applied for
three consecutive steps to a small module built from the reference's own MPConv.

    python tests/golden/make_golden_optim.py      (needs /root/reference; never runs on the GPU box)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

HYPER = dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0, max_norm=10.0)     # config/models/default/unet_train.json
EMAS = {"a": dict(beta=0.9999, feedback_beta=0.9999), "b": dict(beta=0.99), "c": dict(beta=0.999, use_float64=True)}


def main():
    ref_shim.install()
    from modules.module import DualDiffusionModule, DualDiffusionModuleConfig
    from modules.mp_tools import MPConv
    from training.ema import EMA_Manager
    from dataclasses import dataclass

    @dataclass
    class TinyConfig(DualDiffusionModuleConfig):
        pass

    class Tiny(DualDiffusionModule):
        module_name = "tiny"

        def __init__(self, config):
            super().__init__()
            self.config = config
            self.conv_a = MPConv(12, 16, kernel=(3, 3), groups=2)          # rows of 54
            self.conv_b = MPConv(16, 10, kernel=(1, 1))                     # rows of 16
            self.linear = MPConv(40, 7, kernel=())                          # rows of 40 (odd offsets: scalar path)
            self.free = MPConv(9, 5, kernel=(), disable_weight_norm=True)   # never re-normalised
            self.gain = torch.nn.Parameter(torch.tensor(0.3))
            self.vec = torch.nn.Parameter(torch.randn(5000))                # crosses a 4096-element chunk

        def forward(self, x):
            raise NotImplementedError

    torch.manual_seed(1234)
    net = Tiny(TinyConfig())
    net.normalize_weights()
    names = [n for n, _ in net.named_parameters()]
    trainer = types.SimpleNamespace(
        accelerator=types.SimpleNamespace(device=torch.device("cpu"), is_main_process=False),
        persistent_state=types.SimpleNamespace(total_samples_processed=0), total_batch_size=32, global_step=1,
        config=types.SimpleNamespace(model_path="/nonexistent"))
    ema = EMA_Manager("tiny", net, {k: dict(v) for k, v in EMAS.items()}, trainer)
    opt = torch.optim.AdamW(net.parameters(), lr=HYPER["lr"], betas=HYPER["betas"], eps=HYPER["eps"],
                            weight_decay=HYPER["weight_decay"])
    init = {n: p.detach().clone() for n, p in net.named_parameters()}
    steps = []
    wds = [0.0, 0.05, 0.0]
    scales = [1.0, 40.0, 0.01]          # step 2's gradient is large enough to be clipped
    for it in range(3):
        for g in opt.param_groups:
            g["weight_decay"] = wds[it]
        grads = {}
        for n, p in net.named_parameters():
            p.grad = torch.randn(p.shape) * scales[it]
            grads[n] = p.grad.clone()
        norm = torch.nn.utils.clip_grad_norm_(list(net.parameters()), HYPER["max_norm"])
        opt.step()
        opt.zero_grad()
        ema.update()
        net.normalize_weights()
        steps.append(dict(
            grads=grads, weight_decay=wds[it], grad_norm=norm.clone(),
            params={n: p.detach().clone() for n, p in net.named_parameters()},
            exp_avg={n: opt.state[p]["exp_avg"].clone() for n, p in net.named_parameters()},
            exp_avg_sq={n: opt.state[p]["exp_avg_sq"].clone() for n, p in net.named_parameters()},
            emas={k: {n: p.detach().clone() for n, p in ema.ema_modules[k].named_parameters()} for k in EMAS}))
    fan_in = {n: 0 for n in names}
    for mn, m in net.named_modules():
        if isinstance(m, MPConv) and not m.disable_weight_norm:
            fan_in[mn + ".weight"] = m.weight[0].numel()
    torch.save(dict(names=names, init=init, hyper=HYPER, emas=EMAS, fan_in=fan_in, steps=steps),
               os.path.join(OUT, "optim_small.pt"))
    print("wrote optim_small.pt:", {n: tuple(v.shape) for n, v in init.items()}, "grad norms",
          [float(s["grad_norm"]) for s in steps])


if __name__ == "__main__":
    main()
