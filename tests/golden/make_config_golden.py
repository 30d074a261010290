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
