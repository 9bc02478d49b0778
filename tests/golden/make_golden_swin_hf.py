"""Golden vectors that pin ``oracle/swin_v2.py`` (the restatement of timm's ``SwinTransformerV2Stage`` FuXi instantiates,
credit/models/fuxi.py:250-260) against an INDEPENDENT public implementation of the same algorithm: HuggingFace
``transformers.models.swinv2.modeling_swinv2.Swinv2Stage`` (transformers is in the build image; timm is not).  Both are
ports of the Swin-V2 reference code (Liu et al.): scaled-cosine window attention with a clamped learned logit scale,
log-spaced continuous position bias (2 -> 512 -> heads MLP, 16 sigmoid), q / v bias without k bias, shifted windows with
-100 masks, res-post-norm, MLP ratio 4 with exact GELU.  HF keeps q, k, v as three Linear layers and names the post-norms
``layernorm_before`` / ``layernorm_after``; ``hf_to_timm`` below maps its parameters onto timm's tree (the names
``oracle/swin_v2.py`` and ``FuxiB200`` use).

    python tests/golden/make_golden_swin_hf.py        # writes tests/golden/swin_v2_hf.pt (run in the build container)

Every parameter (LayerNorm weights, logit scales, biases included) is randomised so that no term of the block is hidden
behind an identity initialisation."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [  # name, dim, heads, resolution, window, depth, batch
    ("w7_14x21", 32, 4, (14, 21), 7, 4, 2),
    ("w4_8x12", 48, 3, (8, 12), 4, 3, 1),
    ("w7_14x14_dh32", 64, 2, (14, 14), 7, 2, 1),
]
# Not comparable: a grid side <= the window.  HF clamps with ONE scalar (window = min(side), no shift at all), timm clamps
# per dimension (``_calc_window_shift``: the other side keeps its shift) - the restatement follows timm; FuXi pads its token
# grid to window multiples (fuxi.py:67-79), and the 0.25 deg / 1 deg grids have several windows per side.


def hf_stage(dim, heads, res, window, depth, seed):
    from transformers.models.swinv2.configuration_swinv2 import Swinv2Config
    from transformers.models.swinv2.modeling_swinv2 import Swinv2Stage

    cfg = Swinv2Config(window_size=window, hidden_act="gelu", qkv_bias=True, hidden_dropout_prob=0.0,
                       attention_probs_dropout_prob=0.0, drop_path_rate=0.0, layer_norm_eps=1e-5, mlp_ratio=4.0)
    torch.manual_seed(seed)
    stage = Swinv2Stage(cfg, dim=dim, input_resolution=res, depth=depth, num_heads=heads, drop_path=[0.0] * depth,
                        downsample=None, pretrained_window_size=0).eval()
    with torch.no_grad():
        for name, prm in stage.named_parameters():
            if name.endswith("logit_scale"):
                prm.copy_(torch.log(torch.rand_like(prm) * 30 + 1))      # some heads above the ln(100) clamp
            elif "layernorm" in name and name.endswith("weight"):
                prm.copy_(1.0 + 0.3 * torch.randn_like(prm))
            elif name.endswith("bias"):
                prm.copy_(0.2 * torch.randn_like(prm))
            else:
                prm.copy_(torch.randn_like(prm) * (0.5 / prm.shape[-1] ** 0.5 if prm.dim() > 1 else 1.0))
    return stage


def hf_to_timm(sd, depth):
    """HF ``Swinv2Stage.state_dict()`` -> timm ``SwinTransformerV2Stage`` parameter names."""
    out = {}
    for i in range(depth):
        h, t = f"blocks.{i}.", f"blocks.{i}."
        a = h + "attention.self."
        out[t + "attn.qkv.weight"] = torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]])
        out[t + "attn.q_bias"] = sd[a + "query.bias"]
        out[t + "attn.v_bias"] = sd[a + "value.bias"]
        out[t + "attn.logit_scale"] = sd[a + "logit_scale"]
        out[t + "attn.cpb_mlp.0.weight"] = sd[a + "continuous_position_bias_mlp.0.weight"]
        out[t + "attn.cpb_mlp.0.bias"] = sd[a + "continuous_position_bias_mlp.0.bias"]
        out[t + "attn.cpb_mlp.2.weight"] = sd[a + "continuous_position_bias_mlp.2.weight"]
        out[t + "attn.proj.weight"] = sd[h + "attention.output.dense.weight"]
        out[t + "attn.proj.bias"] = sd[h + "attention.output.dense.bias"]
        out[t + "norm1.weight"], out[t + "norm1.bias"] = sd[h + "layernorm_before.weight"], sd[h + "layernorm_before.bias"]
        out[t + "mlp.fc1.weight"], out[t + "mlp.fc1.bias"] = sd[h + "intermediate.dense.weight"], sd[h + "intermediate.dense.bias"]
        out[t + "mlp.fc2.weight"], out[t + "mlp.fc2.bias"] = sd[h + "output.dense.weight"], sd[h + "output.dense.bias"]
        out[t + "norm2.weight"], out[t + "norm2.bias"] = sd[h + "layernorm_after.weight"], sd[h + "layernorm_after.bias"]
    return {k: v.detach().clone() for k, v in out.items()}


def main():
    import transformers

    cases = {}
    for k, (name, dim, heads, res, window, depth, batch) in enumerate(CASES):
        stage = hf_stage(dim, heads, res, window, depth, seed=100 + k)
        torch.manual_seed(200 + k)
        x = torch.randn(batch, res[0], res[1], dim)
        with torch.no_grad():
            y = stage(x.reshape(batch, -1, dim), res)[0].reshape(batch, res[0], res[1], dim)
        cases[name] = dict(dim=dim, heads=heads, resolution=res, window=window, depth=depth, x=x, y=y,
                           state_dict=hf_to_timm(stage.state_dict(), depth))
    out = os.path.join(ROOT, "tests", "golden", "swin_v2_hf.pt")
    torch.save({"transformers": transformers.__version__, "cases": cases}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
