"""TEST INFRASTRUCTURE — import the UNMODIFIED reference forward (``oracle/_ref``, staged by ``oracle/make_ref.py``).

Used by ``tests/``, ``__graft_entry__`` and ``bench.py``'s baseline legs (``--impl reference``, ``cpu_baseline``,
``gpu_eager_baseline``) only; the product (``miles_credit_b200``) never imports this module.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def available(variant: str = "crossformer") -> bool:
    name = {"fuxi": "fuxi.py", "wxformer": os.path.join("wxformer", "crossformer.py")}.get(variant, "crossformer.py")
    return os.path.isfile(os.path.join(STAGED, "credit", "models", name))


def _install_stub():
    """``credit/models/crossformer.py:10`` imports ``credit.postblock.gen1`` (-> xarray, absent here); the class is only
    instantiated under ``post_conf.activate`` (crossformer.py:580-586), which the hot-path configs switch off."""
    if "credit.postblock.gen1" in sys.modules:
        return
    from torch import nn

    stub = types.ModuleType("credit.postblock.gen1")

    class PostBlock(nn.Module):  # never instantiated (post_conf.activate=False)
        def __init__(self, *a, **k):
            super().__init__()
            raise RuntimeError("credit.postblock.gen1 is stubbed: post_conf.activate must be False")

    stub.PostBlock = PostBlock
    sys.modules["credit.postblock.gen1"] = stub


def load_reference(root: str = None):
    """Returns the reference's own ``credit.models.load_model`` (credit/models/__init__.py:301-387)."""
    root = root or STAGED
    if not os.path.isdir(os.path.join(root, "credit")):
        raise FileNotFoundError(f"{root}: the reference is not staged (run oracle/make_ref.py in the build container)")
    if root not in sys.path:
        sys.path.insert(0, root)
    _install_stub()
    from credit.models import load_model  # type: ignore

    return load_model


def reference_model(kwargs: dict, state_dict: dict, variant: str = "crossformer"):
    """The reference ``nn.Module`` for our constructor kwargs with ``state_dict`` loaded ``strict=True``, in ``eval()``."""
    import copy

    if variant == "fuxi":
        # credit/models/fuxi.py:4-5 imports timm's SwinTransformerV2Stage; timm is in neither the reference tree nor this
        # image: the restatement of oracle/swin_v2.py stands in (pinned against HuggingFace's Swinv2Stage, tests/test_swin_v2_vs_hf.py)
        from oracle import swin_v2

        swin_v2.install_timm_stub()
    load_model = load_reference()
    conf = copy.deepcopy({k: v for k, v in kwargs.items() if k != "variant"})
    conf["type"] = variant
    model = load_model({"model": conf})
    if variant == "fuxi":
        model = move(model, "cpu")  # Fuxi.__init__ moves itself to cuda when one is visible (fuxi.py:433-435); callers place it
    model.load_state_dict(state_dict, strict=True)
    return model.eval()


def move(model, device):
    """``model.to(device)`` that keeps the old-style spectral-norm alias intact.

    ``torch.nn.utils.spectral_norm`` leaves ``module.weight`` behind as a PLAIN attribute aliasing ``weight_orig``'s storage
    (so a checkpoint loaded in place shows through it); the forward pre-hook overwrites it on every module call.  timm's
    Swin-V2 attention never calls its ``qkv`` module (``F.linear(x, self.qkv.weight, ...)``), so for that layer the alias IS
    the weight in use (oracle/swin_v2.py).  ``nn.Module.to`` re-creates parameter storage but not plain attributes: after a
    device move the alias would point at the old device and the construction-time values.  In the reference's own flow the
    model is built on its final device and checkpoints load in place (credit/models/base_model.py:72-85), so the alias always
    holds there; this helper restores exactly that state after moving."""
    model = model.to(device)
    for m in model.modules():
        if "weight_orig" in getattr(m, "_parameters", {}) and "weight" in m.__dict__:
            m.__dict__["weight"] = m._parameters["weight_orig"].data
    return model


def seed_policy():
    """The operating conditions of the reference's rollout apps (credit/seed.py:7-25): exact fp32, deterministic."""
    import torch

    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
