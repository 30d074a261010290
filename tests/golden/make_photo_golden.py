#!/usr/bin/env python
"""Generate tests/golden/photo_golden.npz by calling the REFERENCE's own FlameTracker.compute_photometric_energy
(vhap/model/tracker.py:391-478, with get_background_color :288-304 and the render_rgba wrapper :306-338, imported unmodified) on a
bare instance whose `render` is a RECORDING FAKE: it stores the arguments the tracker hands to NVDiffRenderer.render_rgba and returns
a prepared render_out dict.  What gets pinned is everything around the (unavailable) nvdiffrast renderer: background selection per
mode, the v-flip of the UVs, the texture / light arguments, the region look-ups of the stage's align_*_except lists, the
enable_disturbance flag, the [B,H,W,C] -> [B,C,H,W] permutes and the L1 normalisation sum|gt - pred| / sum(alpha > 0 over 3 channels)
with its gradient.

    PYTHONPATH=/root/reference python tests/golden/make_photo_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_stub("nvdiffrast").torch = _stub("nvdiffrast.torch", RasterizeCudaContext=lambda *a, **k: object(), RasterizeGLContext=lambda *a, **k: object())
_stub("pytorch3d"); _stub("pytorch3d.io", load_obj=None); _stub("pytorch3d.structures"); _stub("pytorch3d.structures.meshes", Meshes=None)
_stub("matplotlib", cm=None); _stub("matplotlib.pyplot")
import vhap.model.tracker as RT                         # noqa: E402
from vhap.config import base as RB                      # noqa: E402
from vhap_b200.flame_model import FlameModelData        # noqa: E402


class RecordingRenderer:
    backend = "nvdiffrast"

    def __init__(self, out):
        self.out, self.calls = out, []

    def render_rgba(self, rast_dict, verts, faces, verts_uv, faces_uv, tex, lights, background_color, align_texture_except_fid,
                    align_boundary_except_vid, enable_disturbance):
        self.calls.append(dict(verts_uv=verts_uv.clone(), faces_uv=faces_uv, tex=tex, lights=lights, background_color=background_color,
                               tex_exc=align_texture_except_fid, bnd_exc=align_boundary_except_vid, enable_disturbance=enable_disturbance))
        return dict(self.out)


def main():
    g = torch.Generator().manual_seed(31)
    m = FlameModelData.synthetic()
    B, H, W, T = 2, 12, 20, 8
    gt_rgb = torch.rand(B, 3, H, W, generator=g)
    alpha = (torch.rand(B, H, W, 1, generator=g) > 0.4).float() * torch.rand(B, H, W, 1, generator=g)      # AA'd alpha: 0 or in (0,1]
    rgba = torch.cat([torch.rand(B, H, W, 3, generator=g), alpha], -1).requires_grad_(True)
    render_out = {"rgba": rgba, "albedo": torch.rand(B, H, W, 3, generator=g), "diffuse_detach_normal": torch.rand(B, H, W, 3, generator=g)}
    out = {"gt_rgb": gt_rgb.numpy(), "rgba": rgba.detach().numpy(), "verts_uv": m.verts_uv.astype(np.float32)}
    stages = {"rgb_init_all": RB.StageRgbInitAllConfig(), "rgb_global_tracking": RB.StageRgbGlobalTrackingConfig()}
    pipeline = stages                                     # `self.cfg.pipeline[stage]` (tracker.py:417,420): a mapping by stage name
    for mode in ("target", "white", "black"):
        for stage in ("rgb_init_all", "rgb_global_tracking", None):
            trk = object.__new__(RT.FlameTracker)
            trk.device = "cpu"
            trk.cfg = types.SimpleNamespace(render=types.SimpleNamespace(backend="nvdiffrast", background_train=mode, background_eval=mode), pipeline=pipeline)
            trk.lights = torch.rand(9, 3, generator=g)
            trk.flame = types.SimpleNamespace(
                textures_idx=torch.as_tensor(m.faces_uv.astype(np.int64)), verts_uvs=torch.as_tensor(m.verts_uv.astype(np.float32)),
                mask=types.SimpleNamespace(get_fid_by_region=lambda r: torch.as_tensor(m.get_fid_by_region(list(r))),
                                           get_vid_by_region=lambda r: torch.as_tensor(m.get_vid_by_region(list(r)))))
            trk.render = RecordingRenderer(render_out)
            albedos = torch.rand(1, 3, T, T, generator=g)
            rgba.grad = None
            loss, res = trk.compute_photometric_energy({"rgb": gt_rgb}, torch.zeros(B, 5, 3), None, albedos, {}, stage=stage)
            loss.backward()
            c = trk.render.calls[0]
            key = f"{mode}/{stage}"
            out[key + "/loss"] = np.float64(loss.item())
            out[key + "/g_rgba"] = rgba.grad.numpy().copy()
            out[key + "/verts_uv_arg"] = c["verts_uv"].numpy()
            out[key + "/enable_disturbance"] = np.bool_(c["enable_disturbance"])
            bg = c["background_color"]
            out[key + "/bg_is_tensor"] = np.bool_(torch.is_tensor(bg))
            out[key + "/bg"] = bg.numpy() if torch.is_tensor(bg) else np.asarray(bg, np.float32)
            out[key + "/tex_exc"] = c["tex_exc"].numpy() if c["tex_exc"] is not None else np.zeros(0, np.int64) - 1
            out[key + "/bnd_exc"] = c["bnd_exc"].numpy() if c["bnd_exc"] is not None else np.zeros(0, np.int64) - 1
            out[key + "/tex_is_albedos"] = np.bool_(c["tex"] is albedos)
            out[key + "/lights_shape"] = np.asarray(c["lights"].shape)
            out[key + "/out_shapes"] = np.asarray(res["rgba"].shape)
            print(key, float(loss), c["enable_disturbance"], None if c["tex_exc"] is None else len(c["tex_exc"]), None if c["bnd_exc"] is None else len(c["bnd_exc"]))
    path = Path(__file__).with_name("photo_golden.npz")
    np.savez_compressed(path, **out)
    print(path)


if __name__ == "__main__":
    main()
