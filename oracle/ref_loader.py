"""Import the reference's own modules (build container only; test infrastructure).

`/root/reference` exists only in the build container (the GPU box gets `baseline/_ref`, see `_find_root`). Two third-party imports of the reference
are absent from this image and are stubbed exactly as SURVEY.md §8(c) describes:
`torchsde` (used only by Brownian-tree classes, components/utils.py:54-102) and
`einops_exts.rearrange_many` (attention_utils.py:5).
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """`/root/reference` in the build container; on the GPU box the verbatim, git-ignored copy that baseline/install_ref.py
    made of the modules on the hot path (`baseline/_ref`)."""
    for cand in (os.environ.get("ADB_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "models", "components")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models", "components"))


def import_reference():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True       # the tree is read-only
    if "torchsde" not in sys.modules:
        m = types.ModuleType("torchsde")
        m.BrownianTree = object
        sys.modules["torchsde"] = m
    if "einops_exts" not in sys.modules:
        import einops
        m = types.ModuleType("einops_exts")
        m.rearrange_many = lambda ts, p, **kw: tuple(einops.rearrange(t, p, **kw) for t in ts)
        sys.modules["einops_exts"] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    mods = types.SimpleNamespace()
    mods.wavenet = importlib.import_module("src.models.backbones.wavenet")
    mods.diffusion = importlib.import_module("src.models.components.diffusion")
    mods.sampler_edm = importlib.import_module("src.models.components.sampler_edm")
    mods.scheduler = importlib.import_module("src.models.components.scheduler")
    mods.distribution = importlib.import_module("src.models.components.distribution")
    mods.stochastic_sampler_edm = importlib.import_module("src.models.components.stochastic_sampler_edm")
    return mods


class WaveNetAdapter:
    """net(x[B,1,L], t, **kw) -> wavenet(x.squeeze(1), t): the adapter SURVEY.md §8(c) requires
    because WaveNetNoise.forward takes [B,L] and no kwargs (wavenet.py:170) while denoise_fn calls
    net(c_in*x, c_noise, cond_drop_prob=0., **kw) on [B,1,L] (diffusion.py:50)."""

    def __init__(self, net):
        self.net = net

    def __call__(self, x, t, **kw):
        return self.net(x.squeeze(1), t)
