"""Runs in a fresh interpreter (import order matters): the UNMODIFIED reference is imported first through
oracle/ref_shim.py, then the drop-in classes are loaded the way the reference itself loads plugins -- a model
directory whose model_index.json names `dualdiffusion_b200.*` packages (src/pipelines/dual_diffusion_pipeline.py:217-228)
read by the reference's own `DualDiffusionPipeline.from_pretrained` (:232-300) / `DualDiffusionModule.from_pretrained`
(src/modules/module.py:59-84).

    python tests/boundary_driver.py cpu <tmpdir>     # load path only (no GPU)
    python tests/boundary_driver.py gpu <tmpdir>     # + the reference's own diffusion_decode and train-batch math on cuda:0

Prints one JSON object on the last line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode, tmp = sys.argv[1], sys.argv[2]

import torch  # noqa: E402

from oracle import ref_shim, unet_oracle as uo  # noqa: E402

ref_shim.install()
# the reference's own modules, imported BEFORE the drop-in package so that the drop-ins subclass its bases
from modules.module import DualDiffusionModule  # noqa: E402
from modules.unets.unet_edm2_b4 import UNet as RefUNet, UNetConfig as RefUNetConfig  # noqa: E402
from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat as RefFormat, MS_MDCT_DualFormatConfig as RefFormatConfig  # noqa: E402
from pipelines.dual_diffusion_pipeline import DualDiffusionPipeline as RefPipeline, SampleParams as RefSampleParams  # noqa: E402

out = {}
spec = uo.small_spec()
sd = uo.synth_state_dict(spec, seed=0)
cfg = RefUNetConfig(**{k: getattr(spec, k) for k in RefUNetConfig.__dataclass_fields__ if hasattr(spec, k)})

# 1. a model directory written by the REFERENCE (its own save_pretrained: config json + safetensors)
ref_unet = RefUNet(cfg)
ref_unet.load_state_dict(sd, strict=True)
ref_pipe = RefPipeline({"unet": ref_unet, "format": RefFormat(RefFormatConfig())})
model_dir = os.path.join(tmp, "model")
ref_pipe.save_pretrained(model_dir)
index = json.load(open(os.path.join(model_dir, "model_index.json")))
out["reference_index"] = json.loads(json.dumps(index["modules"]))

# 2. the one registration step of SURVEY 8(b): point `package` at the new implementation
index["modules"]["unet"] = {"package": "dualdiffusion_b200.modules.unets.unet_edm2_b4", "class": "UNet"}
index["modules"]["format"] = {"package": "dualdiffusion_b200.modules.formats.ms_mdct_dual", "class": "MS_MDCT_DualFormat"}
json.dump(index, open(os.path.join(model_dir, "model_index.json"), "w"))

classes = RefPipeline.get_model_module_classes(model_dir)
out["classes"] = {k: f"{v.__module__}.{v.__name__}" for k, v in classes.items()}
out["subclass_of_reference_base"] = all(issubclass(c, DualDiffusionModule) for c in classes.values())

# 3. the reference's own loader: from_pretrained -> module.from_pretrained -> strict load_state_dict
pipe = RefPipeline.from_pretrained(model_dir, torch_dtype=torch.float32, device="cpu")
out["isinstance"] = isinstance(pipe.unet, DualDiffusionModule) and isinstance(pipe.format, DualDiffusionModule)
out["unet_class"] = type(pipe.unet).__module__
loaded = pipe.unet.state_dict()
out["state_dict_equal"] = sorted(loaded.keys()) == sorted(sd.keys()) and all(torch.equal(loaded[k], sd[k]) for k in sd)
_norm = lambda v: list(v) if isinstance(v, (list, tuple)) else v          # JSON turns tuples into lists
out["config_equal"] = all(_norm(getattr(pipe.unet.config, k)) == _norm(getattr(cfg, k)) for k in RefUNetConfig.__dataclass_fields__)
out["last_global_step"] = pipe.model_metadata["last_global_step"]

# every other drop-in class resolves its config class the same way (module.py:72)
import importlib  # noqa: E402
import inspect  # noqa: E402
from dataclasses import is_dataclass  # noqa: E402
resolved = {}
for pkg, cls in (("dualdiffusion_b200.modules.daes.dae_edm2_d3", "DAE_D3"),
                 ("dualdiffusion_b200.modules.formats.spectrogram", "SpectrogramFormat"),
                 ("dualdiffusion_b200.modules.unets.unet_edm2_ddec_mclt_b1", "DDec_MCLT_UNet_B1"),
                 ("dualdiffusion_b200.modules.unets.unet_edm2_q4_ddec", "UNet"),
                 ("dualdiffusion_b200.modules.unets.unet_edm2_b4_2", "UNet"),
                 ("dualdiffusion_b200.modules.daes.dae_edm2_q4", "DAE"),
                 ("dualdiffusion_b200.modules.formats.ms_mdct_dual_2", "MS_MDCT_DualFormat")):
    c = getattr(importlib.import_module(pkg), cls)
    cc = c.config_class or inspect.signature(c.__init__).parameters["config"].annotation
    resolved[f"{pkg}.{cls}"] = bool(is_dataclass(cc)) and issubclass(c, DualDiffusionModule)
out["other_dropins_resolve"] = resolved

if mode == "gpu":
    dev = torch.device("cuda:0")
    pipe = pipe.to(device=dev)
    ref_pipe = ref_pipe.to(device=dev)
    g = torch.Generator().manual_seed(3)
    clap = torch.randn(1, spec.in_channels_emb, generator=g)
    # 4. the reference trainer's loss math (unet_trainer.py:259-280) + backward through the drop-in (train mode)
    x = torch.randn(2, spec.in_channels, 32, 48, generator=g).to(dev)
    noise = torch.randn(2, spec.in_channels, 32, 48, generator=g).to(dev)
    sigma = torch.tensor([2.0, 0.4], device=dev)
    emb_in = torch.randn(2, spec.in_channels_emb, generator=g).to(dev)
    mask = torch.tensor([True, False], device=dev)

    def train_loss(unet):
        unet.train().requires_grad_(True)
        sigma_data = unet.config.sigma_data
        emb = unet.get_embeddings(emb_in, mask)
        denoised = unet(x + noise * sigma.view(-1, 1, 1, 1), sigma, pipe.format, emb, None)
        w = (sigma ** 2 + sigma_data ** 2) / (sigma * sigma_data) ** 2
        batch_weighted_loss = torch.nn.functional.mse_loss(denoised, x, reduction="none").mean(dim=(1, 2, 3)) * w
        error_logvar = unet.get_sigma_loss_logvar(sigma)
        return (batch_weighted_loss / error_logvar.exp() + error_logvar).mean()

    la, lb = train_loss(pipe.unet), train_loss(ref_pipe.unet)
    la.backward()
    lb.backward()
    out["train_loss_rel_err"] = abs(float(la) - float(lb)) / abs(float(lb))
    ga = dict(pipe.unet.named_parameters())
    gb = dict(ref_pipe.unet.named_parameters())
    worst = 0.0
    per = {}
    for name in ("dec.block0_layer0.conv_res1.weight", "enc.block1_layer0.conv_res0.weight", "emb_noise.weight",
                 "emb_label.weight", "logvar_linear.weight"):
        a, b = ga[name].grad.float(), gb[name].grad.float()
        per[name] = float((a - b).norm() / (b.norm() + 1e-30))
        out.setdefault("dbg", {})[name] = (float(a.norm()), float(b.norm()), float((a * b).sum() / (a.norm() * b.norm() + 1e-30)))
        worst = max(worst, per[name])
    out["train_grad_rel_err"] = worst
    # third opinion: autograd through the CPU oracle on the same inputs
    sdg = {k: (v.clone().requires_grad_(True) if "fourier" not in k else v) for k, v in sd.items()}
    xc, nc, sc = x.cpu(), noise.cpu(), sigma.cpu()
    u_ = uo.mp_conv(torch.ones(1), sdg["emb_label_unconditional.weight"], training=True)
    c_ = uo.mp_conv(uo.normalize(emb_in.cpu()), sdg["emb_label.weight"], training=True)
    emb_o = uo.mp_sum(u_, c_, mask.cpu().unsqueeze(1).float())
    d_o = uo.unet_forward(sdg, spec, xc + nc * sc.view(-1, 1, 1, 1), sc, emb_o, training=True)
    w_o = (sc ** 2 + spec.sigma_data ** 2) / (sc * spec.sigma_data) ** 2
    bwl = torch.nn.functional.mse_loss(d_o, xc, reduction="none").mean(dim=(1, 2, 3)) * w_o
    lv = uo.sigma_loss_logvar(sdg, sc)
    if lv is not None:
        lo = (bwl / lv.exp() + lv).mean()
        lo.backward()
        name = "dec.block0_layer0.conv_res1.weight"
        go = sdg[name].grad
        cos = lambda p_, q_: float((p_ * q_).sum() / (p_.norm() * q_.norm() + 1e-30))
        out["oracle_vs"] = dict(loss_oracle=float(lo), loss_dropin=float(la), loss_ref_gpu=float(lb),
                                cos_dropin=cos(go, ga[name].grad.float().cpu()), cos_ref_gpu=cos(go, gb[name].grad.float().cpu()))
    out["train_grad_rel_err_per_param"] = per
    for net in (pipe.unet, ref_pipe.unet):
        net.zero_grad(set_to_none=True)
        net.requires_grad_(False).train(False)

    # 5. the reference's OWN sampler loop (pipeline.py:589-752) over the drop-in UNet vs over its own UNet, same seed
    params = RefSampleParams(seed=11, num_steps=4, batch_size=1, cfg_scale=1.5, use_heun=True)
    shape = (1, spec.in_channels, 32, 48)
    with torch.no_grad():
        got = pipe.diffusion_decode(params, audio_embedding=clap.to(dev), sample_shape=shape, quiet=True)
        want = ref_pipe.diffusion_decode(params, audio_embedding=clap.to(dev), sample_shape=shape, quiet=True)
    out["decode_rel_err"] = float((got.float() - want.float()).norm() / want.float().norm())
print(json.dumps(out))
