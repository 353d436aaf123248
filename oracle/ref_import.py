"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (read-only /root/reference) in the build container.

Used by oracle/make_golden.py (fixture generation) and by `-m "not gpu"` tests that pin oracle/ff_oracle.py
against the reference when /root/reference is present.  Nothing in the product (freefine_b200/), in `-m gpu`
tests, smoke() or bench.py may import this module: /root/reference does not exist on the GPU box.

Stub recipe: SURVEY.md Appendix D.1.  diffusers 0.18.0 / matplotlib / rembg / pytorch_lightning / pytorch3d are
absent from the image, the reference's hot-path modules only need their *names* at import time.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF_ROOT = os.environ.get("FREEFINE_REF_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "utils", "attention.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Permissive(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return type(k, (), {})


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    _installed = True
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mp = _mod("matplotlib")
            mp.pyplot = _mod("matplotlib.pyplot")
    try:
        import diffusers  # noqa: F401
    except Exception:
        d = _mod("diffusers", StableDiffusionPipeline=type("StableDiffusionPipeline", (), {}),
                 DDIMScheduler=type("DDIMScheduler", (), {}))
        d.utils = _mod("diffusers.utils")

        def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
            return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)

        d.utils.torch_utils = _mod("diffusers.utils.torch_utils", randn_tensor=randn_tensor)
    try:
        import rembg  # noqa: F401
    except Exception:
        _mod("rembg", remove=lambda *a, **k: None)
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        def seed_everything(seed, *a, **k):
            torch.manual_seed(seed)
            return seed
        pl = _mod("pytorch_lightning", seed_everything=seed_everything)
        pl.utilities = _mod("pytorch_lightning.utilities", rank_zero_warn=lambda *a, **k: None)
    try:
        import pytorch3d  # noqa: F401
    except Exception:
        for n in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.points", "pytorch3d.transforms",
                  "pytorch3d.structures", "pytorch3d.renderer.points.rasterizer", "pytorch3d.renderer.points.compositor",
                  "pytorch3d.ops"):
            sys.modules[n] = _Permissive(n)


def load():
    """Returns a namespace with the reference modules: attention, model, vis_utils, geo_utils (None if it fails)."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    attention = importlib.import_module("src.utils.attention")
    model = importlib.import_module("src.demo.model")
    vis_utils = importlib.import_module("src.utils.vis_utils")
    try:
        geo_utils = importlib.import_module("src.utils.geo_utils")
    except Exception:  # pragma: no cover
        geo_utils = None
    return types.SimpleNamespace(attention=attention, model=model, vis_utils=vis_utils, geo_utils=geo_utils)


def make_reference_pipeline(ref, parts, device="cpu", flavour="edit"):
    """FreeFinePipeline.__new__ + stand-in parts, wired like freefine_batch_infer_2d.py:149-155."""
    P = ref.model.FreeFinePipeline
    if not isinstance(getattr(P, "device", None), property):
        P.device = property(lambda self: self._ff_device)
    pipe = P.__new__(P)
    pipe._ff_device = torch.device(device)
    pipe.unet, pipe.vae = parts.unet, parts.vae
    pipe.tokenizer, pipe.text_encoder, pipe.scheduler = parts.tokenizer, parts.text_encoder, parts.scheduler
    controller = ref.attention.Attention_Modulator(start_layer=10)
    pipe.controller = controller
    {"edit": ref.attention.register_attention_control, "bggen": ref.attention.register_attention_control_4bggen,
     "compose": ref.attention.register_attention_control_compose}[flavour](pipe, controller)
    pipe.modify_unet_forward()
    return pipe, controller
