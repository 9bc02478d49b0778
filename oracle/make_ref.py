"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference forward under ``oracle/_ref/``.

    python oracle/make_ref.py            # build container only (needs /root/reference)

The reference (NCAR/miles-credit @ 5e14b83c) is pure Python, so "compiling" it means staging the import closure of its
forward path where the GPU box can import it: the files are copied byte for byte from ``/root/reference`` (never edited,
never committed: ``oracle/_ref/`` is git-ignored and travels to the box with the snapshot like a built ``.so``).  The
closure is discovered by importing the path here — ``credit.models.load_model`` for the ``crossformer`` / ``wxformer``
registry keys, ``credit.boundary_padding`` — and listing the ``credit.*`` modules that import pulled in; a sha256
manifest (``oracle/_ref/MANIFEST.json``) records every staged file against its source.

The single obstacle to importing the path is ``credit/models/crossformer.py:10`` -> ``credit.postblock.gen1`` -> xarray
(absent from this image): ``oracle/ref_loader.py`` places a stub module in ``sys.modules`` first; the stubbed class is never
instantiated when ``post_conf.activate`` is False.  Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s
baseline legs use what is staged here; the product never imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CREDIT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "credit")):
        raise SystemExit(f"{REF} is not present: oracle/_ref can only be staged in the build container")
    sys.path.insert(0, HERE)
    import ref_loader  # noqa: E402  (installs the postblock stub, puts REF on sys.path)

    load_model = ref_loader.load_reference(root=REF)
    import torch

    tiny = dict(
        frames=1, image_height=45, image_width=96, levels=3, channels=2, surface_channels=2, input_only_channels=2,
        output_only_channels=1, patch_width=1, patch_height=1, dim=[32, 64, 128, 256], depth=[1, 1, 2, 1],
        global_window_size=[8, 4, 2, 1], local_window_size=3, cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4], [2, 4], [2, 4]],
        cross_embed_strides=[2, 2, 2, 2], attn_dropout=0.0, ff_dropout=0.0, interp=True, use_spectral_norm=True,
        padding_conf=dict(activate=True, mode="earth", pad_lat=[25, 27], pad_lon=[24, 24]), post_conf={"activate": False})
    for key in ("crossformer", "wxformer"):
        m = load_model({"model": dict(tiny, type=key)})
        with torch.no_grad():
            m.train()
            m(torch.randn(1, 10, 1, 45, 96))
    # FuXi (credit/models/fuxi.py) with the Swin-V2 stand-in registered as timm (oracle/swin_v2.py)
    import swin_v2  # noqa: E402

    swin_v2.install_timm_stub()
    fx = dict(image_height=32, image_width=96, patch_height=4, patch_width=4, frames=2, frame_patch_size=2, levels=2, channels=2,
              surface_channels=1, input_only_channels=0, output_only_channels=0, dim=32, num_groups=8, num_heads=4, depth=2,
              window_size=5, use_spectral_norm=True, interp=True, padding_conf=dict(activate=False), post_conf={"activate": False})
    m = load_model({"model": dict(fx, type="fuxi")}).cpu()
    with torch.no_grad():
        m(torch.randn(1, 5, 2, 32, 96))
    import credit.boundary_padding  # noqa: F401,E402
    import credit.seed  # noqa: F401,E402

    files = {}
    for name, mod in sorted(sys.modules.items()):
        f = getattr(mod, "__file__", None)
        if (name == "credit" or name.startswith("credit.")) and f and os.path.abspath(f).startswith(REF + os.sep):
            files[os.path.relpath(os.path.abspath(f), REF)] = os.path.abspath(f)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {"source": REF, "commit": None, "files": {}}
    try:
        manifest["commit"] = json.load(open(os.path.join(REF, ".SUBMODULES.json")))["commit"]
    except Exception:  # noqa: BLE001
        pass
    for rel, src in files.items():
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = _sha(src)
        assert _sha(dst) == manifest["files"][rel]
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(files)} unmodified reference files under {DST}")
        for rel in files:
            print("  ", rel)
    return manifest


if __name__ == "__main__":
    stage()
