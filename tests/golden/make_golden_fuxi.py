"""Generate the FuXi golden vector from the UNMODIFIED reference module (run in the build container only).

    python tests/golden/make_golden_fuxi.py

``credit/models/fuxi.py`` needs ``timm`` (``:4-5``), which is absent from the reference tree and from this image: the
Swin-V2 stand-in of ``oracle/swin_v2.py`` is registered under timm's two import paths first, so everything FuXi owns in the
reference tree (padding, cube embedding, DownBlock / UpBlock, head, un-patchify, resize, spectral-norm hooks on every
Conv2d / Linear / ConvTranspose2d) runs as the reference wrote it, while the third-party stage is the restatement on both
sides (parity of the stage: unpinned).  The model goes through the reference's own registry path
``credit.models.load_model(conf)``; weights are the reference's default init under a fixed seed, warmed by train-mode
forwards so the spectral-norm ``u/v`` vectors are meaningful; the fixture stores input, output and the
state dict (small model: 1.3 MB).
"""
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

stub = types.ModuleType("credit.postblock.gen1")


class PostBlock(nn.Module):  # never instantiated (post_conf.activate=False)
    def __init__(self, *a, **k):
        super().__init__()


stub.PostBlock = PostBlock
sys.modules["credit.postblock.gen1"] = stub

from oracle import swin_v2  # noqa: E402

assert swin_v2.install_timm_stub(), "a real timm is importable here: regenerate against it and mark the stage pinned"

from credit.models import load_model  # noqa: E402

from oracle import fuxi_oracle  # noqa: E402

CASES = {
    # 45 x 88 image, earth pad (9, 10) / (12, 12) -> 64 x 112, patch (2, 4, 4) -> 16 x 28 patches, token grid 8 x 14,
    # window 3 -> zero pad to 9 x 15 (top 0 / bottom 1, left 0 / right 1); 4 blocks (two of them shifted)
    "unit_fuxi": dict(
        image_height=45, image_width=88, patch_height=4, patch_width=4, frames=2, frame_patch_size=2, levels=3, channels=2,
        surface_channels=2, input_only_channels=1, output_only_channels=2, dim=32, num_groups=4, num_heads=2, depth=4,
        window_size=3, use_spectral_norm=True, interp=True,
        padding_conf=dict(activate=True, mode="earth", pad_lat=[9, 10], pad_lon=[12, 12]),
        post_conf={"activate": False}),
    # no padding, window larger than one side of the token grid (clamped, shift dropped on that axis)
    "unit_fuxi_nopad": dict(
        image_height=32, image_width=96, patch_height=4, patch_width=4, frames=2, frame_patch_size=2, levels=2, channels=2,
        surface_channels=1, input_only_channels=0, output_only_channels=0, dim=32, num_groups=8, num_heads=4, depth=2,
        window_size=5, use_spectral_norm=True, interp=True, padding_conf=dict(activate=False),
        post_conf={"activate": False}),
}


def main():
    for name, kw in CASES.items():
        torch.manual_seed(1000)
        model = load_model({"model": dict(kw, type="fuxi")}).cpu().float()
        spec = fuxi_oracle.FuxiSpec.from_kwargs(**kw)
        x = torch.randn(2, spec.in_chans, spec.frames, spec.image_height, spec.image_width)
        model.train()
        with torch.no_grad():
            for _ in range(5):  # power iterations of the spectral-norm hooks
                model(x)
        model.eval()
        with torch.no_grad():
            y = model(x)
            y2 = model(x)
        assert torch.equal(y, y2)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            yo = fuxi_oracle.forward(x, sd, spec)
        err = float((yo - y).abs().max() / y.abs().max())
        print(f"{name}: out {tuple(y.shape)} abs-max {float(y.abs().max()):.3f}  oracle vs reference rel-max {err:.3e}  "
              f"{len(sd)} state-dict keys")
        torch.save({"kwargs": kw, "x": x, "y": y, "state_dict": sd, "oracle_rel_max": err,
                    "swin_stage": "restatement (oracle/swin_v2.py) on both sides: timm absent"},
                   os.path.join(HERE, f"{name}.pt"))


if __name__ == "__main__":
    main()
