"""Loader for the UNMODIFIED reference tree that __graft_entry__.build() stages under the git-ignored baseline/_ref/src
(the reference has no setup.py / pyproject: it is a source tree run with PYTHONPATH=src, so it is staged, not
pip-installed).  Used by bench.py's reference arm (its CPU path on the host cores) and its `gpu_eager` comparator (the
reference's own eager PyTorch path on the same B200) -- never by the product package.  Seven I/O-only third-party
packages the reference imports at module top level are absent from this image; none is touched
by the hot path, so they are replaced by empty stub modules before import."""
import json
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.abspath(__file__))


def reference_src() -> str:
    for cand in (os.environ.get("DUALDIFFUSION_REFERENCE_SRC"), os.path.join(ROOT, "_ref", "src"), "/root/reference/src"):
        if cand and os.path.isdir(os.path.join(cand, "modules")):
            return cand
    return ""


def available() -> bool:
    return bool(reference_src())


def _strip_json5(text: str) -> str:
    text = re.sub(r"//[^\n]*", "", text)
    return re.sub(r",(\s*[}\]])", r"\1", text)


def install() -> str:
    src = reference_src()
    if not src:
        raise RuntimeError("reference tree not staged (run __graft_entry__.build() where /root/reference exists)")
    for name in ("mutagen", "mutagen.flac", "pyloudnorm", "librosa", "accelerate", "accelerate.logging", "accelerate.utils"):
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
    if src not in sys.path:
        sys.path.insert(0, src)
    return src
