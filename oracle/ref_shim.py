"""TEST INFRASTRUCTURE ONLY — not part of the product path.

Imports the *unmodified* reference (parlance-zz/dualdiffusion, mounted read-only at
/root/reference) on CPU so its own PyTorch code can (a) validate the oracle
restatement in this directory and (b) generate the golden vectors under
tests/golden/.  /root/reference does not exist on the GPU box, so nothing in
`-m gpu` tests, smoke() or bench.py may import this module.

Seven I/O-only third-party packages the reference imports at module top level are
absent from this image; none is touched by the hot path, so they
are replaced by empty stub modules before import.
"""
import json
import os
import re
import sys
import types

def _find_reference() -> str:
    """/root/reference in the build container; on the GPU box the unmodified src/ tree staged by
    __graft_entry__.build() under the git-ignored baseline/_ref/."""
    env = os.environ.get("DUALDIFFUSION_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/src/modules"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REFERENCE_ROOT = _find_reference()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "modules"))


def _strip_json5(text: str) -> str:
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r",(\s*[}\]])", r"\1", text)
    return text


def install() -> None:
    """Put the reference's src/ on sys.path with stubs for absent I/O deps."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("mutagen", "mutagen.flac", "pyloudnorm", "librosa",
                 "accelerate", "accelerate.logging", "accelerate.utils"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "pyjson5" not in sys.modules:
        try:
            __import__("pyjson5")
        except Exception:
            stub = types.ModuleType("pyjson5")
            stub.load = lambda f: json.loads(_strip_json5(f.read()))
            stub.loads = lambda s: json.loads(_strip_json5(s))
            stub.dump = lambda obj, f, **kw: json.dump(obj, f, **kw)
            stub.dumps = lambda obj, **kw: json.dumps(obj, **kw)
            sys.modules["pyjson5"] = stub
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
