"""Configuration and grid geometry of the WXFormer/CrossFormer forecast step.

This is host logic only (no tensors): it turns the reference's ``model:`` YAML block
(constructor kwargs of ``credit.models.crossformer.CrossFormer``, reference
``credit/models/crossformer.py:372-401``) into the per-stage grid sizes, the channel
plan of every cross-embed branch and the list of state-dict keys with their shapes
(reference key layout: SURVEY.md §8b, probed from ``CrossFormer.state_dict()``).

Everything downstream (oracle, weight preparation, CUDA launch plans, tests) reads the
geometry from here so that there is exactly one statement of "what shape is what".
"""

from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple


def _tup(val, length=4):
    """Mirror of the reference's ``cast_tuple`` (crossformer.py:19-20) for YAML lists."""
    if isinstance(val, (list, tuple)):
        return tuple(val)
    return (val,) * length


@dataclass(frozen=True)
class Padding:
    """``padding_conf`` of the reference (boundary_padding.py:6-18)."""

    activate: bool = False
    mode: str = "earth"
    pad_lat: Tuple[int, int] = (0, 0)
    pad_lon: Tuple[int, int] = (0, 0)

    @staticmethod
    def from_conf(conf) -> "Padding":
        if not conf or not conf.get("activate", False):
            return Padding(False, "earth", (0, 0), (0, 0))
        mode = conf.get("mode", "earth")
        if mode not in ("earth", "mirror"):
            raise ValueError(f"padding mode must be 'earth' or 'mirror', got {mode!r}")
        pad_lat = conf.get("pad_lat", (40, 40))
        pad_lon = conf.get("pad_lon", (40, 40))
        # credit/parser.py:436-440 turns an int into [p, p]
        if isinstance(pad_lat, int):
            pad_lat = (pad_lat, pad_lat)
        if isinstance(pad_lon, int):
            pad_lon = (pad_lon, pad_lon)
        return Padding(True, mode, (int(pad_lat[0]), int(pad_lat[1])), (int(pad_lon[0]), int(pad_lon[1])))


@dataclass(frozen=True)
class Branch:
    """One Conv2d of a CrossEmbedLayer (crossformer.py:139-148)."""

    kernel: int
    stride: int
    pad: int
    c_out: int
    c_off: int  # channel offset of this branch inside the concatenated stage output


@dataclass(frozen=True)
class Stage:
    index: int
    c_in: int
    dim: int
    depth: int
    heads: int
    h_in: int
    w_in: int
    h: int
    w: int
    local_window: int
    global_window: int
    branches: Tuple[Branch, ...]


@dataclass(frozen=True)
class UpStage:
    """One decoder UpBlock (crossformer.py:70-122): ConvT k2s2 then 2x(conv3x3+GN+SiLU)+skip."""

    name: str
    c_in: int
    c_out: int
    h_in: int
    w_in: int
    groups: int


@dataclass(frozen=True)
class Geometry:
    image_height: int
    image_width: int
    frames: int
    output_frames: int
    channels: int
    levels: int
    surface_channels: int
    input_only_channels: int
    output_only_channels: int
    base_input_channels: int
    input_channels: int
    base_output_channels: int
    output_channels: int
    dim: Tuple[int, ...]
    depth: Tuple[int, ...]
    dim_head: int
    use_spectral_norm: bool
    interp: bool
    padding: Padding
    h_pad: int
    w_pad: int
    stages: Tuple[Stage, ...]
    ups: Tuple[UpStage, ...]
    h_dec: int  # decoder output grid (before unpad)
    w_dec: int
    h_crop: int  # after unpad
    w_crop: int
    h_out: int  # after optional bilinear resize
    w_out: int
    variant: str = "crossformer"  # "crossformer": ConvTranspose decoder; "wxformer": PixelShuffle decoder

    @property
    def in_shape(self):
        return (self.base_input_channels, self.frames, self.image_height, self.image_width)

    @property
    def out_shape(self):
        return (self.base_output_channels, self.output_frames, self.h_out, self.w_out)


def cross_embed_channel_split(dim_out: int, kernel_sizes: Sequence[int]) -> List[int]:
    """Channel count per kernel, sorted-kernel order (crossformer.py:131-136)."""
    n = len(kernel_sizes)
    scales = [int(dim_out / (2**i)) for i in range(1, n)]
    return [*scales, dim_out - sum(scales)]


def build_geometry(
    image_height: int = 640,
    patch_height: int = 1,
    image_width: int = 1280,
    patch_width: int = 1,
    frames: int = 2,
    output_frames: int = 1,
    channels: int = 4,
    surface_channels: int = 7,
    input_only_channels: int = 3,
    output_only_channels: int = 0,
    levels: int = 15,
    dim=(64, 128, 256, 512),
    depth=(2, 2, 8, 2),
    dim_head: int = 32,
    global_window_size=(5, 5, 2, 1),
    local_window_size=10,
    cross_embed_kernel_sizes=((4, 8, 16, 32), (2, 4), (2, 4), (2, 4)),
    cross_embed_strides=(4, 2, 2, 2),
    attn_dropout: float = 0.0,
    ff_dropout: float = 0.0,
    use_spectral_norm: bool = True,
    attention_type=None,
    interp: bool = True,
    upsample_v_conv: bool = False,
    padding_conf=None,
    post_conf=None,
    variant: str = "crossformer",
    upsample_with_ps: bool = True,
    **kwargs,
) -> Geometry:
    """Same keyword surface and defaults as ``CrossFormer.__init__`` (crossformer.py:372-401).

    ``variant="wxformer"`` selects credit/models/wxformer/crossformer.py (registry keys ``wxformer`` /
    ``wxformer_base``): same encoder, cross-embed branches wrapped in an explicit ZeroPad2d (:199-236), PixelShuffle
    decoder (``UpBlockPS`` :137-162, ``up_block4`` :813-830); ``upsample_with_ps`` is accepted and ignored like there.

    Unknown keys (e.g. ``frame_patch_size``) are swallowed like the reference's ``**kwargs``.
    Features outside the forecast hot path raise ``NotImplementedError`` instead of silently
    computing something else.
    """
    if patch_height != 1 or patch_width != 1:
        raise NotImplementedError("cube embedding (patch_height/patch_width > 1) is outside the hot path")
    if variant not in ("crossformer", "wxformer"):
        raise ValueError(f"unknown variant {variant!r}")
    if upsample_v_conv and variant == "crossformer":
        # (the `wxformer` registry class has no such option: the key falls into its **kwargs and is ignored,
        # wxformer/crossformer.py:636-653 - e.g. config/gen_2/smoke/smoke_gen2_multistep_casper.yml sets it)
        raise NotImplementedError("upsample_v_conv=True decoder is not built (no BASELINE config uses it)")
    if attention_type is not None:
        raise NotImplementedError("UpBlock attention_type is not built (None in every BASELINE config)")
    if post_conf is not None and post_conf.get("activate", False):
        raise NotImplementedError("in-model PostBlock is out of scope; use post_conf.activate=False")
    if kwargs.get("diffusion"):
        raise NotImplementedError("diffusion conditioning is out of scope")

    dim = _tup(dim)
    depth = _tup(depth)
    gws = _tup(global_window_size)
    lws = _tup(local_window_size)
    kernels = tuple(tuple(k) for k in cross_embed_kernel_sizes)
    if len(kernels) and not isinstance(cross_embed_kernel_sizes[0], (list, tuple)):
        kernels = (tuple(cross_embed_kernel_sizes),) * 4
    strides = _tup(cross_embed_strides)
    for name, v in (("dim", dim), ("depth", depth), ("global_window_size", gws), ("local_window_size", lws),
                    ("cross_embed_kernel_sizes", kernels), ("cross_embed_strides", strides)):
        if len(v) != 4:
            raise AssertionError(f"{name} must have 4 entries, got {len(v)}")

    padding = Padding.from_conf(padding_conf)
    base_in = channels * levels + surface_channels + input_only_channels
    base_out = channels * levels + surface_channels + output_only_channels
    c_in0 = base_in * frames
    c_out = base_out * output_frames

    h_pad = image_height + padding.pad_lat[0] + padding.pad_lat[1]
    w_pad = image_width + padding.pad_lon[0] + padding.pad_lon[1]

    stages = []
    h, w, c = h_pad, w_pad, c_in0
    for i in range(4):
        ks = sorted(kernels[i])
        split = cross_embed_channel_split(dim[i], ks)
        branches, off, sizes = [], 0, set()
        for k, co in zip(ks, split):
            p = (k - strides[i]) // 2
            if variant == "wxformer":  # ZeroPad2d(left = (k-s)//2, right = (k-s) - left) then an unpadded conv
                ho = (h + (k - strides[i]) - k) // strides[i] + 1
                wo = (w + (k - strides[i]) - k) // strides[i] + 1
            else:
                ho = (h + 2 * p - k) // strides[i] + 1
                wo = (w + 2 * p - k) // strides[i] + 1
            sizes.add((ho, wo))
            branches.append(Branch(k, strides[i], p, co, off))
            off += co
        if len(sizes) != 1:
            raise ValueError(f"stage {i}: cross-embed branches disagree on output size {sorted(sizes)}")
        ho, wo = sizes.pop()
        if dim[i] % dim_head:
            raise ValueError(f"dim[{i}]={dim[i]} is not a multiple of dim_head={dim_head}")
        for wsz, kind in ((lws[i], "local"), (gws[i], "global")):
            if ho % wsz or wo % wsz:
                # the reference fails inside einops.rearrange at the first forward
                raise ValueError(f"stage {i}: grid {ho}x{wo} is not divisible by {kind} window {wsz}")
        stages.append(Stage(i, c, dim[i], depth[i], dim[i] // dim_head, h, w, ho, wo, lws[i], gws[i], tuple(branches)))
        h, w, c = ho, wo, dim[i]

    last = dim[-1]
    ups = []
    uh, uw = stages[3].h, stages[3].w
    plan = (("up_block1", last, last // 2, 2), ("up_block2", 2 * (last // 2), last // 4, 1),
            ("up_block3", 2 * (last // 4), last // 8, 0))
    for name, ci, co, skip in plan:
        ups.append(UpStage(name, ci, co, uh, uw, dim[0]))
        uh, uw = 2 * uh, 2 * uw
        if (uh, uw) != (stages[skip].h, stages[skip].w):
            raise ValueError(f"{name}: decoder grid {uh}x{uw} does not match encoder stage {skip} "
                             f"grid {stages[skip].h}x{stages[skip].w}")
        if co != dim[skip]:
            raise ValueError(f"{name}: skip concat needs dim[{skip}]={dim[skip]} == {co}")
        if co % dim[0]:
            raise ValueError(f"{name}: GroupNorm needs {co} channels divisible by {dim[0]} groups")
    h_dec, w_dec = 2 * uh, 2 * uw  # up_block4: ConvTranspose2d k4 s2 p1 doubles the grid
    h_crop = h_dec - padding.pad_lat[0] - padding.pad_lat[1]
    w_crop = w_dec - padding.pad_lon[0] - padding.pad_lon[1]
    if interp:
        h_out, w_out = image_height, image_width
    else:
        h_out, w_out = h_crop, w_crop

    return Geometry(
        image_height, image_width, frames, output_frames, channels, levels, surface_channels,
        input_only_channels, output_only_channels, base_in, c_in0, base_out, c_out, dim, depth, dim_head,
        bool(use_spectral_norm), bool(interp), padding, h_pad, w_pad, tuple(stages), tuple(ups),
        h_dec, w_dec, h_crop, w_crop, h_out, w_out, variant,
    )


# --------------------------------------------------------------------------------------
# state-dict layout


def _sn(spec, prefix, w_shape, sn, sn_dim=0, bias=True, bias_len=None):
    """Keys of one Conv/Linear module, with or without the old-style spectral-norm hook.

    With the hook (torch.nn.utils.spectral_norm, applied at crossformer.py:23-26) a module
    owns ``bias, weight_orig, weight_u, weight_v``; ConvTranspose2d uses dim=1.
    """
    if bias:
        spec[prefix + ".bias"] = ((bias_len if bias_len is not None else w_shape[0],), "bias")
    if sn:
        height = w_shape[sn_dim]
        width = 1
        for i, s in enumerate(w_shape):
            if i != sn_dim:
                width *= s
        spec[prefix + ".weight_orig"] = (tuple(w_shape), f"weight:{sn_dim}")
        spec[prefix + ".weight_u"] = ((height,), "u")
        spec[prefix + ".weight_v"] = ((width,), "v")
    else:
        spec[prefix + ".weight"] = (tuple(w_shape), f"weight:{sn_dim}")


def state_spec(geo: Geometry) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Every persistent tensor of the reference module: key -> (shape, role)."""
    sn = geo.use_spectral_norm
    wx = geo.variant == "wxformer"
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()
    for st in geo.stages:
        s = st.index
        for i, br in enumerate(st.branches):
            # wxformer: nn.Sequential(ZeroPad2d, Conv2d) -> the conv is child "1" (tests/test_legacy_checkpoint_compat.py:24-26)
            _sn(spec, f"layers.{s}.0.convs.{i}" + (".1" if wx else ""), (br.c_out, st.c_in, br.kernel, br.kernel), sn)
        d = st.dim
        dq = d // 4
        for l in range(st.depth):
            for a in (0, 2):  # short, long attention
                p = f"layers.{s}.1.layers.{l}.{a}"
                spec[p + ".norm.g"] = ((1, d, 1, 1), "gain")
                spec[p + ".norm.b"] = ((1, d, 1, 1), "shift")
                _sn(spec, p + ".to_qkv", (3 * d, d, 1, 1), sn, bias=False)
                _sn(spec, p + ".to_out", (d, d, 1, 1), sn)
                _sn(spec, p + ".dpb.layers.0", (dq, 2), sn)
                for li, nxt in ((1, 3), (4, 6), (7, 9)):
                    spec[p + f".dpb.layers.{li}.weight"] = ((dq,), "gain")
                    spec[p + f".dpb.layers.{li}.bias"] = ((dq,), "shift")
                    _sn(spec, p + f".dpb.layers.{nxt}", ((dq if nxt != 9 else 1), dq), sn)
            for f in (1, 3):  # feed-forward blocks
                p = f"layers.{s}.1.layers.{l}.{f}.layers"
                spec[p + ".0.g"] = ((1, d, 1, 1), "gain")
                spec[p + ".0.b"] = ((1, d, 1, 1), "shift")
                _sn(spec, p + ".1", (4 * d, d, 1, 1), sn)
                _sn(spec, p + ".4", (d, 4 * d, 1, 1), sn)
    # cube_embedding is constructed but unused when patch=1 (crossformer.py:533-538, 601-602);
    # Conv3d is not spectral-normed (crossformer.py:25)
    spec["cube_embedding.proj.weight"] = ((geo.dim[0], geo.input_channels, geo.frames, 1, 1), "weight:0")
    spec["cube_embedding.proj.bias"] = ((geo.dim[0],), "bias")
    spec["cube_embedding.norm.weight"] = ((geo.dim[0],), "gain")
    spec["cube_embedding.norm.bias"] = ((geo.dim[0],), "shift")
    for up in geo.ups:
        if wx:  # UpBlockPS (wxformer/crossformer.py:137-162): conv3x3 -> 4*C, PixelShuffle, sharpen conv, residual stack
            _sn(spec, f"{up.name}.conv", (4 * up.c_out, up.c_in, 3, 3), sn)
            _sn(spec, f"{up.name}.sharp", (up.c_out, up.c_out, 3, 3), sn)
        else:
            _sn(spec, f"{up.name}.conv", (up.c_in, up.c_out, 2, 2), sn, sn_dim=1, bias_len=up.c_out)
        for ci, gi in ((0, 1), (3, 4)):
            _sn(spec, f"{up.name}.b.{ci}", (up.c_out, up.c_out, 3, 3), sn)
            spec[f"{up.name}.b.{gi}.weight"] = ((up.c_out,), "gain")
            spec[f"{up.name}.b.{gi}.bias"] = ((up.c_out,), "shift")
    c4 = 2 * (geo.dim[-1] // 8)
    if wx:  # up_block4 = Sequential(conv3x3 -> 4*C_out, PixelShuffle(2), conv3x3) (wxformer/crossformer.py:813-830)
        _sn(spec, "up_block4.0", (4 * geo.output_channels, c4, 3, 3), sn)
        _sn(spec, "up_block4.2", (geo.output_channels, geo.output_channels, 3, 3), sn)
    else:
        _sn(spec, "up_block4", (c4, geo.output_channels, 4, 4), sn, sn_dim=1, bias_len=geo.output_channels)
    return spec


def flops_per_forward(geo: Geometry) -> Dict[str, float]:
    """Algorithmic FLOPs (2*MAC) by op class, the figure SURVEY.md §8(d) defines."""
    out = {"cross_embed": 0.0, "qkv": 0.0, "qk": 0.0, "pv": 0.0, "out_proj": 0.0, "ff": 0.0,
           "dec_conv3x3": 0.0, "dec_up": 0.0}
    for st in geo.stages:
        n = st.h * st.w
        for br in st.branches:
            out["cross_embed"] += 2.0 * n * br.c_out * st.c_in * br.kernel * br.kernel
        d = st.dim
        for wsz in (st.local_window, st.global_window):
            L = wsz * wsz
            per = st.depth
            out["qkv"] += per * 2.0 * n * d * 3 * d
            out["out_proj"] += per * 2.0 * n * d * d
            out["qk"] += per * 2.0 * n * L * d
            out["pv"] += per * 2.0 * n * L * d
        out["ff"] += 2 * st.depth * 2.0 * n * d * 4 * d * 2
    wx = geo.variant == "wxformer"
    for up in geo.ups:
        n_in = up.h_in * up.w_in
        if wx:
            out["dec_up"] += 2.0 * n_in * up.c_in * 4 * up.c_out * 9 + 2.0 * (4 * n_in) * up.c_out * up.c_out * 9
        else:
            out["dec_up"] += 2.0 * n_in * up.c_in * up.c_out * 4
        out["dec_conv3x3"] += 2 * 2.0 * (4 * n_in) * up.c_out * up.c_out * 9
    c4 = 2 * (geo.dim[-1] // 8)
    n0 = geo.stages[0].h * geo.stages[0].w
    if wx:
        out["dec_up"] += 2.0 * n0 * c4 * 4 * geo.output_channels * 9 + 2.0 * (4 * n0) * geo.output_channels ** 2 * 9
    else:
        out["dec_up"] += 2.0 * n0 * c4 * geo.output_channels * 16
    out["total"] = sum(out.values())
    return out


# --------------------------------------------------------------------------------------
# named workloads (BASELINE.json configs -> constructor kwargs)


def check_kernel_limits(geo: "Geometry") -> List[str]:
    """Limits of the sm_100a kernels, checked at construction instead of at the first forward.  Raises for geometries no
    kernel covers; returns warnings for geometries that run on a slower path."""
    notes = []
    if geo.dim_head != 32:
        raise NotImplementedError(f"dim_head={geo.dim_head}: the window-attention kernels are built for dim_head = 32 "
                                  "(every CREDIT WXFormer/CrossFormer config)")
    for st in geo.stages:
        for wsz, kind in ((st.local_window, "local"), (st.global_window, "global")):
            if wsz * wsz > 128:
                raise NotImplementedError(f"stage {st.index}: {kind} window {wsz}x{wsz} = {wsz * wsz} tokens; the attention "
                                          "kernels hold at most 128 tokens per window")
        if st.dim % 8:
            raise NotImplementedError(f"dim[{st.index}]={st.dim} must be a multiple of 8 (16-byte operand rows)")
    st0 = geo.stages[0]
    if st0.c_in > 64:
        notes.append(f"{st0.c_in} input channels > 64: the stage-0 cross-embed runs on the exact-fp32 CUDA-core kernel "
                     "(much slower than the Toeplitz tensor-core kernel) and the domain decomposition is unavailable")
    if geo.output_channels % 4:
        notes.append(f"{geo.output_channels} output channels (not a multiple of 4): the decoder head runs on the exact-fp32 "
                     "CUDA-core kernel")
    return notes


def workload(name: str) -> dict:
    """Constructor kwargs of the named BASELINE.json configs."""
    wx6h = dict(
        frames=1, levels=13, channels=4, surface_channels=4, input_only_channels=4, output_only_channels=8,
        patch_width=1, patch_height=1, frame_patch_size=1, dim=[128, 256, 512, 1024], depth=[2, 2, 8, 2],
        global_window_size=[10, 5, 2, 1], local_window_size=10,
        cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4], [2, 4], [2, 4]], cross_embed_strides=[2, 2, 2, 2],
        attn_dropout=0.0, ff_dropout=0.0, interp=True, use_spectral_norm=True, post_conf={"activate": False},
    )
    if name == "wxformer_6h_025deg":  # config/gen_2/examples/wxformer_era5_025deg_6hr.yml:169-208
        return dict(wx6h, image_height=721, image_width=1440,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[40, 40], pad_lon=[80, 80]))
    if name == "wxformer_6h_1deg":  # same architecture at 181x360
        return dict(wx6h, image_height=181, image_width=360,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[69, 70], pad_lon=[60, 60]))
    if name == "smoke_1deg":  # credit_smoke_test_v2.yml:119-160 as shipped
        return dict(
            frames=1, image_height=181, image_width=360, levels=18, channels=4, surface_channels=4,
            input_only_channels=4, output_only_channels=8, patch_width=1, patch_height=1, frame_patch_size=1,
            dim=[64, 128, 256, 512], depth=[2, 2, 4, 2], global_window_size=[8, 4, 2, 1], local_window_size=3,
            cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4], [2, 4], [2, 4]], cross_embed_strides=[2, 2, 2, 2],
            attn_dropout=0.0, ff_dropout=0.0, interp=True, use_spectral_norm=True,
            padding_conf=dict(activate=True, mode="earth", pad_lat=[30, 30], pad_lon=[12, 12]),
            post_conf={"activate": False},
        )
    if name == "smoke_tiny":  # credit_smoke_test_v2.yml shrunk to 64x128 (BASELINE.json configs[0])
        return dict(workload("smoke_1deg"), image_height=64, image_width=128,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[16, 16], pad_lon=[8, 8]))
    if name == "unit":  # fixture-sized model used by the golden vectors (tests/golden)
        return dict(
            frames=1, image_height=45, image_width=96, levels=3, channels=2, surface_channels=2,
            input_only_channels=2, output_only_channels=1, patch_width=1, patch_height=1,
            dim=[32, 64, 128, 256], depth=[1, 1, 2, 1], global_window_size=[8, 4, 2, 1], local_window_size=3,
            cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4], [2, 4], [2, 4]], cross_embed_strides=[2, 2, 2, 2],
            attn_dropout=0.0, ff_dropout=0.0, interp=True, use_spectral_norm=True,
            padding_conf=dict(activate=True, mode="earth", pad_lat=[25, 27], pad_lon=[24, 24]),
            post_conf={"activate": False},
        )
    raise KeyError(name)
