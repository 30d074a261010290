"""Host-side configuration mirroring the reference's dataclass tree (vhap/config/base.py) for the fields the
photometric inner loop reads.  Field names and defaults follow base.py so a `BaseTrackingConfig` instance from the
reference can be passed in unchanged (only attribute access is used)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple


@dataclass
class LossWeightConfig:            # base.py:126-195
    landmark: Optional[float] = 10.0
    always_enable_jawline_landmarks: bool = True
    photo: Optional[float] = 30.0
    reg_shape: float = 3e-1
    reg_neck: float = 3e-1
    reg_jaw: float = 3e-1
    reg_eyes: float = 3e-2
    reg_expr: float = 3e-2
    reg_tex_res_clusters: Optional[float] = 1e1
    reg_tex_res_for: Tuple[str, ...] = ("sclerae", "teeth")
    reg_tex_tv: Optional[float] = 1e4
    reg_light: Optional[float] = None
    reg_diffuse: Optional[float] = 1e2
    reg_offset: Optional[float] = 3e2
    reg_offset_relax_coef: float = 1.0
    reg_offset_relax_for: Tuple[str, ...] = ("hair", "ears")
    reg_offset_lap: Optional[float] = 1e6
    reg_offset_lap_relax_coef: float = 0.1
    reg_offset_lap_relax_for: Tuple[str, ...] = ("hair", "ears")
    reg_offset_rigid: Optional[float] = 3e2
    reg_offset_rigid_for: Tuple[str, ...] = ("left_ear", "right_ear", "neck", "left_eye", "right_eye", "lips_tight")
    reg_offset_dynamic: Optional[float] = 3e5
    blur_iter: int = 0
    smooth_trans: float = 3e2
    smooth_rot: float = 3e1
    smooth_neck: float = 3e1
    smooth_jaw: float = 1e-1
    smooth_eyes: float = 0.0
    smooth_expr: float = 1e0


@dataclass
class LearningRateConfig:          # base.py:114-122
    base: float = 5e-3
    translation: float = 1e-3
    expr: float = 5e-2
    static_offset: float = 5e-4
    dynamic_offset: float = 5e-4
    camera: float = 5e-3
    light: float = 5e-3


@dataclass
class RenderConfig:                # base.py:94-110
    backend: str = "b200"
    background_train: str = "target"
    background_eval: str = "target"
    disturb_rate_fg: Optional[float] = 0.5
    disturb_rate_bg: Optional[float] = 0.5
    lighting_type: str = "SH"
    lighting_space: str = "world"


@dataclass
class StageConfig:                 # base.py:215-295
    name: str = "rgb_global_tracking"
    optimizable_params: Tuple[str, ...] = ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset")
    photometric: bool = True
    disable_jawline_landmarks: bool = True
    align_texture_except: Tuple[str, ...] = ()
    align_boundary_except: Tuple[str, ...] = ("bottomline",)


# step / epoch counts of the reference's pipeline (base.py:228-295): how many optimisation steps a stage runs per call
STAGE_STEPS = {"lmk_init_rigid": 500, "lmk_init_all": 500, "lmk_sequential_tracking": 50, "rgb_init_texture": 500, "rgb_init_all": 500,
               "rgb_init_offset": 500, "rgb_sequential_tracking": 50}
STAGE_EPOCHS = {"lmk_global_tracking": 30, "rgb_global_tracking": 30}

# (the reference also lists "dynamic_offset" for the two rgb tracking stages; dynamic offsets are off by default, base.py:69, and
#  rejected by this engine)
STAGES = {
    "lmk_init_rigid": StageConfig("lmk_init_rigid", ("cam", "pose"), False, False, (), ()),
    "lmk_init_all": StageConfig("lmk_init_all", ("cam", "pose", "shape", "joints", "expr"), False, False, (), ()),
    "lmk_sequential_tracking": StageConfig("lmk_sequential_tracking", ("pose", "joints", "expr"), False, False, (), ()),
    "lmk_global_tracking": StageConfig("lmk_global_tracking", ("cam", "pose", "shape", "joints", "expr"), False, False, (), ()),
    "rgb_init_texture": StageConfig("rgb_init_texture", ("cam", "shape", "texture", "lights"), True, False,
                                    ("hair", "boundary", "neck"), ("hair", "boundary")),
    "rgb_init_all": StageConfig("rgb_init_all", ("cam", "pose", "shape", "joints", "expr", "texture", "lights"), True, True,
                                ("hair", "boundary", "neck"), ("hair", "bottomline")),
    "rgb_init_offset": StageConfig("rgb_init_offset", ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset"),
                                   True, True, ("hair", "boundary", "neck"), ("bottomline",)),
    "rgb_sequential_tracking": StageConfig("rgb_sequential_tracking", ("pose", "joints", "expr", "texture"), True, True, (), ("bottomline",)),
    "rgb_global_tracking": StageConfig("rgb_global_tracking", ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset"),
                                       True, True, (), ("bottomline",)),
}


@dataclass
class EngineConfig:
    w: LossWeightConfig = field(default_factory=LossWeightConfig)
    lr: LearningRateConfig = field(default_factory=LearningRateConfig)
    render: RenderConfig = field(default_factory=RenderConfig)
    n_shape: int = 300
    n_expr: int = 100
    tex_resolution: int = 2048
    tex_clusters: Tuple[str, ...] = ("skin", "hair", "boundary", "lips_tight", "teeth", "sclerae", "irises")
    scale_factor: float = 1.0            # cfg.data.scale_factor (tracker.py:530)
    n_downsample_rgb: Optional[int] = None
    calibrated: bool = False


def opt_dict_for(stage: StageConfig) -> dict:
    """tracker.py:1465-1513 `get_train_parameters`: which parameter groups a stage optimises."""
    keys = ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset", "dynamic_offset")
    return {k: (k in stage.optimizable_params) for k in keys}


# engine parameter name -> name of its learning rate in LearningRateConfig (everything else uses `base`), tracker.py:159-211
_LR_OF = {"translation": "translation", "expr": "expr", "lights": "light", "focal_length": "camera", "static_offset": "static_offset"}


def adam_param_lrs(stage: StageConfig, lr: LearningRateConfig, lr_scale: float = 1.0, calibrated: bool = False) -> dict:
    """The Adam parameter groups of a stage: {engine parameter name: learning rate}, i.e. what the reference builds with
    get_train_parameters (tracker.py:1465-1513) + configure_optimizer (tracker.py:159-211); texture = `tex_extra`."""
    opt = opt_dict_for(stage)
    names = []
    if opt["cam"] and not calibrated: names.append("focal_length")
    if opt["shape"]: names.append("shape")
    if opt["texture"]: names.append("tex_extra")
    if opt["static_offset"]: names.append("static_offset")
    if opt["lights"]: names.append("lights")
    if opt["pose"]: names += ["translation", "rotation"]
    if opt["joints"]: names += ["eyes_pose", "neck_pose", "jaw_pose"]
    if opt["expr"]: names.append("expr")
    return {n: getattr(lr, _LR_OF.get(n, "base")) * lr_scale for n in names}


def stage_schedule(stage_name: str, n_batches: int, lr_scale: float = 1.0, per_sample: bool = False):
    """The iteration schedule of FlameTracker.optimize_stage (tracker.py:1391-1416) as a list of (batch index, learning-rate scale):
    * per_sample (a single staged sample, the init / sequential stages): `num_steps` iterations on batch 0 at constant lr_scale;
    * otherwise (a dataloader, the global stages): `num_epochs` passes over the batches in the order given (shuffling is the
      loader's business) with torch's ExponentialLR(gamma=0.9) stepped after every epoch -- lr_scale * 0.9 ** epoch.
    One fresh Adam state per call (tracker.py:1399), so the caller starts with Engine.set_stage(stage_name, lr_scale)."""
    if per_sample:
        return [(0, lr_scale)] * STAGE_STEPS[stage_name]
    out = []
    for epoch in range(STAGE_EPOCHS[stage_name]):
        out += [(b, lr_scale * 0.9 ** epoch) for b in range(n_batches)]
    return out



# ---------------------------------------------------------------------------------------------- NeRSemble (calibrated multi-view)
@dataclass
class NersembleLossWeightConfig(LossWeightConfig):      # vhap/config/nersemble.py:36-42
    landmark: Optional[float] = 3.0
    always_enable_jawline_landmarks: bool = False
    reg_expr: float = 1e-2
    reg_tex_tv: Optional[float] = 1e5
    smooth_expr: float = 0.0


# the two stages NeRSemble overrides (nersemble.py:44-61); every other stage is the base one.  (rgb_sequential_tracking additionally drops
# the texture group there, nersemble.py:46: ("pose", "joints", "expr", "dynamic_offset") -- dynamic offsets are off by default, base.py:69.)
NERSEMBLE_STAGES = dict(STAGES)
NERSEMBLE_STAGES["rgb_sequential_tracking"] = StageConfig("rgb_sequential_tracking", ("pose", "joints", "expr"), True, True, ("boundary",), ("boundary",))
NERSEMBLE_STAGES["rgb_global_tracking"] = StageConfig("rgb_global_tracking", ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset"),
                                                      True, True, ("boundary",), ("boundary",))


def nersemble_config(**overrides) -> EngineConfig:
    """EngineConfig of NersembleTrackingConfig (vhap/config/nersemble.py:23-82): calibrated cameras (per-frame 'intrinsic' /
    'extrinsic' in the sample, no focal-length parameter), the NeRSemble loss weights; stages from NERSEMBLE_STAGES."""
    return EngineConfig(w=NersembleLossWeightConfig(), calibrated=True, **overrides)
