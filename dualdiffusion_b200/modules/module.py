"""Module base of the drop-in classes.

Mirrors the part of the reference's `DualDiffusionModule` contract (src/modules/module.py:42-191) that the hot path
needs: loading (config dataclass <- `<name>.json`, `<name>.safetensors` weights with strict key matching), dtype /
device / memory_format tracking and `normalize_weights()`.  The control-plane methods (save_pretrained, load_ema,
blend_weights) are not restated: under the reference they come from its own base class.
When the reference itself is loaded in the process (its
`modules.module` is in sys.modules because its pipeline / trainer imported us through model_index.json),
the classes here *are* subclasses of the reference bases, so `isinstance(x, DualDiffusionModule)` checks in
src/pipelines/dual_diffusion_pipeline.py:131-133 pass.  Stand-alone (GPU box, no reference) the local
mirror below is used.
"""
from __future__ import annotations

import inspect
import json
import os
import sys
from abc import ABC
from dataclasses import dataclass, fields
from typing import Optional, Type, Union

import torch


def _reference_loaded() -> bool:
    return "modules.module" in sys.modules and hasattr(sys.modules["modules.module"], "DualDiffusionModule")


if _reference_loaded():   # running under the reference's own pipeline / trainer
    from modules.module import DualDiffusionModule, DualDiffusionModuleConfig  # type: ignore
else:

    @dataclass
    class DualDiffusionModuleConfig(ABC):
        last_global_step: int = 0

    class DualDiffusionModule(torch.nn.Module, ABC):
        config_class: Optional[Type[DualDiffusionModuleConfig]] = None
        module_name: Optional[str] = None
        has_trainable_parameters: bool = True
        supports_half_precision: bool = True
        supports_channels_last: Union[bool, str] = True
        supports_compile: bool = False

        def __init__(self) -> None:
            super().__init__()
            self.dtype = torch.get_default_dtype()
            self.device = torch.device("cpu")
            self.memory_format = torch.contiguous_format
            self.module_path = None

        # ---- persistence: module.py:59-102 ----
        @classmethod
        @torch.no_grad()
        def from_pretrained(cls, module_path: str, subfolder: Optional[str] = None,
                            torch_dtype: Optional[torch.dtype] = None, device: Optional[torch.device] = None,
                            load_config_only: bool = False) -> "DualDiffusionModule":
            if subfolder is not None:
                module_path = os.path.join(module_path, subfolder)
            config_class = cls.config_class or inspect.signature(cls.__init__).parameters["config"].annotation
            name = os.path.basename(module_path)
            with open(os.path.join(module_path, f"{name}.json")) as fh:
                raw = json.load(fh)
            known = {f.name for f in fields(config_class)}
            module = cls(config_class(**{k: v for k, v in raw.items() if k in known}))
            module.requires_grad_(False).train(False)
            if not load_config_only and cls.has_trainable_parameters:
                from safetensors.torch import load_file
                module.load_state_dict(load_file(os.path.join(module_path, f"{name}.safetensors")))
            module.module_path = module_path
            return module.to(dtype=torch_dtype, device=device)

        # ---- dtype / device tracking: module.py:104-143 ----
        def to(self, device=None, dtype=None, memory_format=None, **kwargs) -> "DualDiffusionModule":
            if device is not None:
                device = torch.device(device)
            if dtype in (torch.float16, torch.bfloat16) and not type(self).supports_half_precision:
                dtype = torch.float32
            if memory_format == torch.channels_last:
                if type(self).supports_channels_last is False:
                    memory_format = None
                elif type(self).supports_channels_last == "3d":
                    memory_format = torch.channels_last_3d
            super().to(device=device, dtype=dtype, memory_format=memory_format, **kwargs)
            self.dtype = dtype or self.dtype
            self.device = device or self.device
            self.memory_format = memory_format or self.memory_format
            return self

        def float(self):
            return self.to(dtype=torch.float32)

        def half(self):
            return self.to(dtype=torch.bfloat16)

        def cpu(self, **kwargs):
            return self.to(device="cpu", **kwargs)

        def cuda(self, device: Optional[int] = None):
            return self.to(device="cuda" if device is None else f"cuda:{device}")

        def compile(self, **kwargs) -> None:
            # module.py:145-149 wraps forward in torch.compile; the B200 path is hand-written kernels +
            # CUDA graphs, so this is a deliberate no-op (supports_compile = False).
            return None

        @torch.no_grad()
        def normalize_weights(self) -> None:
            if not type(self).has_trainable_parameters:
                return
            for module in self.modules():
                if hasattr(module, "normalize_weights") and module is not self:
                    module.normalize_weights()
