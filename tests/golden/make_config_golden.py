#!/usr/bin/env python
"""Generate tests/golden/config_golden.json from the REFERENCE's own config dataclasses (vhap/config/base.py, imported
unmodified with PYTHONPATH=/root/reference): loss weights, learning rates, render / model defaults and the stage table
(optimisable parameter groups, photometric flag, jawline switch, align_*_except lists, step counts).  The reference is not
available on the GPU box, so the values are committed and tests/test_config_golden.py holds vhap_b200/config.py to them.

    PYTHONPATH=/root/reference python tests/golden/make_config_golden.py
"""
import dataclasses
import json
import sys
from pathlib import Path

sys.path.insert(0, "/root/reference")
from vhap.config import base as B      # noqa: E402

STAGE_CLASSES = {
    "lmk_init_rigid": B.StageLmkInitRigidConfig, "lmk_init_all": B.StageLmkInitAllConfig,
    "lmk_sequential_tracking": B.StageLmkSequentialTrackingConfig, "lmk_global_tracking": B.StageLmkGlobalTrackingConfig,
    "rgb_init_texture": B.StageRgbInitTextureConfig, "rgb_init_all": B.StageRgbInitAllConfig, "rgb_init_offset": B.StageRgbInitOffsetConfig,
    "rgb_sequential_tracking": B.StageRgbSequentialTrackingConfig, "rgb_global_tracking": B.StageRgbGlobalTrackingConfig,
}


def main():
    out = {"w": dataclasses.asdict(B.LossWeightConfig()), "lr": dataclasses.asdict(B.LearningRateConfig()),
           "render": dataclasses.asdict(B.RenderConfig()), "model": dataclasses.asdict(B.ModelConfig()), "stages": {}}
    for name, cls in STAGE_CLASSES.items():
        d = dataclasses.asdict(cls())
        d["photometric"] = isinstance(cls(), B.PhotometricStageConfig)
        out["stages"][name] = d
    path = Path(__file__).with_name("config_golden.json")
    path.write_text(json.dumps(out, indent=1, sort_keys=True, default=str) + "\n")
    print(path, {k: list(v) for k, v in out["stages"].items()})


if __name__ == "__main__":
    main()


def optimizer_groups():
    """{stage: {parameter name: lr}} from the reference's own get_train_parameters + configure_optimizer (tracker.py:1465-1513,159-211)
    run on a bare tracker (no __init__), lr_scale 0.1 like global tracking (tracker.py:1385)."""
    import types
    import torch

    def _stub(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m
    _stub("nvdiffrast").torch = _stub("nvdiffrast.torch", RasterizeCudaContext=lambda *a, **k: object(), RasterizeGLContext=lambda *a, **k: object())
    _stub("pytorch3d"); _stub("pytorch3d.io", load_obj=None); _stub("pytorch3d.structures"); _stub("pytorch3d.structures.meshes", Meshes=None)
    _stub("matplotlib", cm=None); _stub("matplotlib.pyplot")
    import vhap.model.tracker as RT
    names = {"focal_length": (1,), "shape": (300,), "tex_extra": (3, 4, 4), "static_offset": (1, 7, 3), "lights": (9, 3), "translation": (2, 3), "rotation": (2, 3),
             "eyes_pose": (2, 6), "neck_pose": (2, 3), "jaw_pose": (2, 3), "expr": (2, 100)}
    res = {}
    for stage, cls in STAGE_CLASSES.items():
        trk = object.__new__(RT.GlobalTracker)            # get_train_parameters lives on the subclass (tracker.py:1221,1465)
        trk.calibrated = False
        trk.cfg = types.SimpleNamespace(pipeline={stage: cls()}, model=B.ModelConfig(), lr=B.LearningRateConfig())
        tensors = {n: torch.zeros(*shp, requires_grad=True) for n, shp in names.items()}
        for n, t in tensors.items():
            setattr(trk, n, t)
        trk.tex_pca = torch.zeros(100, requires_grad=True)
        trk.dynamic_offset = None
        optim = trk.configure_optimizer(trk.get_train_parameters(stage), lr_scale=0.1)
        by_id = {id(t): n for n, t in tensors.items()}
        res[stage] = {by_id[id(p)]: g["lr"] for g in optim.param_groups for p in g["params"]}
    return res


def nersemble():
    """the NeRSemble overrides (vhap/config/nersemble.py:23-82): data flags the engine reads, loss weights, the two overridden stages"""
    from vhap.config import nersemble as N
    data = N.NersembleDataConfig
    out = {"w": dataclasses.asdict(N.NersembleLossWeightConfig()),
           "data": {f.name: f.default for f in dataclasses.fields(data) if f.name in ("calibrated", "scale_factor", "n_downsample_rgb", "target_extrinsic_type",
                                                                                       "background_color", "image_size_during_calibration")},
           "stages": {}}
    for name, cls in (("rgb_sequential_tracking", N.NersembleStageRgbSequentialTrackingConfig), ("rgb_global_tracking", N.NersembleStageRgbGlobalTrackingConfig)):
        d = dataclasses.asdict(cls())
        d["photometric"] = isinstance(cls(), B.PhotometricStageConfig)
        out["stages"][name] = d
    return out


def stage_schedules():
    """[(batch index, lr of the `base` group / cfg.lr.base)] per iteration, from the reference's own optimize_stage (tracker.py:1391-1416)
    run on a bare tracker with optimize_iter replaced by a recorder: a 3-sample "dataloader" for the two global stages (lr_scale 0.1,
    ExponentialLR) and a single sample for two per-sample stages."""
    import types
    import torch
    import vhap.model.tracker as RT
    res = {}
    for stage, cls, loader in (("rgb_global_tracking", B.StageRgbGlobalTrackingConfig, True), ("lmk_global_tracking", B.StageLmkGlobalTrackingConfig, True),
                               ("rgb_init_texture", B.StageRgbInitTextureConfig, False), ("rgb_sequential_tracking", B.StageRgbSequentialTrackingConfig, False)):
        trk = object.__new__(RT.GlobalTracker)
        trk.calibrated = False
        trk.logger = types.SimpleNamespace(info=lambda *a, **k: None)
        trk.cfg = types.SimpleNamespace(pipeline={stage: cls()}, model=B.ModelConfig(), lr=B.LearningRateConfig())
        for n, shp in {"focal_length": (1,), "shape": (300,), "tex_extra": (3, 4, 4), "static_offset": (1, 7, 3), "lights": (9, 3), "translation": (2, 3),
                       "rotation": (2, 3), "eyes_pose": (2, 6), "neck_pose": (2, 3), "jaw_pose": (2, 3), "expr": (2, 100)}.items():
            setattr(trk, n, torch.zeros(*shp, requires_grad=True))
        trk.tex_pca = torch.zeros(100, requires_grad=True)
        trk.dynamic_offset = None
        rec = []
        trk.optimize_iter = lambda sample, optimizer, st: rec.append((sample, optimizer.param_groups[-1]["lr"] / B.LearningRateConfig().base))
        trk.evaluate = lambda *a, **k: None
        if loader:
            trk.optimize_stage(stage=stage, dataloader=[0, 1, 2], lr_scale=0.1)
        else:
            trk.optimize_stage(stage=stage, sample=0, lr_scale=1.0)
        res[stage] = {"per_sample": not loader, "n_batches": 3 if loader else 1, "lr_scale": 0.1 if loader else 1.0, "schedule": rec}
    return res


if __name__ == "__main__":
    path = Path(__file__).with_name("config_golden.json")
    d = json.loads(path.read_text())
    d["nersemble"] = nersemble()
    d["optimizer_groups_lr_scale_0.1"] = optimizer_groups()
    d["stage_schedules"] = stage_schedules()
    path.write_text(json.dumps(d, indent=1, sort_keys=True, default=str) + "\n")
    print({k: v for k, v in d["optimizer_groups_lr_scale_0.1"].items() if k in ("rgb_global_tracking", "lmk_init_rigid")})
