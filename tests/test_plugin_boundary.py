"""The drop-in boundary, exercised through the REFERENCE's own loader and callers:
a model directory's model_index.json names `dualdiffusion_b200.*` packages and the reference's unmodified
`DualDiffusionPipeline.from_pretrained` (src/pipelines/dual_diffusion_pipeline.py:217-300) /
`DualDiffusionModule.from_pretrained` (src/modules/module.py:59-84) load them; on the GPU the reference's own
`diffusion_decode` (pipeline.py:589-752) and the trainer's loss math (unet_trainer.py:259-280) + backward run over the
drop-in UNet and are compared with the same calls over the reference's own UNet on the same device.

The driver runs in a fresh interpreter because the drop-in base class is chosen at import time (reference loaded first).
The reference tree is /root/reference here and the unmodified copy staged under the git-ignored baseline/_ref/ on the
GPU box (__graft_entry__.build)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _run(mode, tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "boundary_driver.py"), mode, str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _check_load(out):
    assert out["reference_index"]["unet"]["package"] == "modules.unets.unet_edm2_b4"        # written by the reference
    assert out["classes"] == {"unet": "dualdiffusion_b200.modules.unets.unet_edm2_b4.UNet",
                              "format": "dualdiffusion_b200.modules.formats.ms_mdct_dual.MS_MDCT_DualFormat"}
    assert out["subclass_of_reference_base"] and out["isinstance"]
    assert out["unet_class"] == "dualdiffusion_b200.modules.unets.unet_edm2_b4"
    assert out["state_dict_equal"], "strict load of a reference-saved safetensors changed the weights"
    assert out["config_equal"]
    assert out["last_global_step"] == {"unet": 0, "format": 0}
    assert all(out["other_dropins_resolve"].values()), out["other_dropins_resolve"]


@needs_reference
def test_reference_from_pretrained_loads_the_dropins(tmp_path):
    _check_load(_run("cpu", tmp_path))


@pytest.mark.gpu
@needs_reference
def test_reference_sampler_and_train_math_over_the_dropin(tmp_path):
    out = _run("gpu", tmp_path)
    _check_load(out)
    # bf16 tensor-core body vs the reference's fp32 eager path on the same GPU: the whole-network tolerance (3e-2)
    assert out["decode_rel_err"] < 3e-2, out
    assert out["train_loss_rel_err"] < 2e-2, out
    assert out["train_grad_rel_err"] < 6e-2, out
