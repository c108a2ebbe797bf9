"""Format base class mirror (reference src/modules/formats/format.py:29-42)."""
from __future__ import annotations

import sys
from abc import ABC
from dataclasses import dataclass

from ..module import DualDiffusionModule, DualDiffusionModuleConfig

if "modules.formats.format" in sys.modules and hasattr(sys.modules["modules.formats.format"], "DualDiffusionFormat"):
    from modules.formats.format import DualDiffusionFormat, DualDiffusionFormatConfig  # type: ignore
else:

    @dataclass
    class DualDiffusionFormatConfig(DualDiffusionModuleConfig):
        sample_rate: int = 32000
        num_raw_channels: int = 2
        default_raw_length: int = 1408768

    class DualDiffusionFormat(DualDiffusionModule, ABC):
        module_name: str = "format"
        has_trainable_parameters: bool = False
        supports_half_precision: bool = False
        supports_compile: bool = False
