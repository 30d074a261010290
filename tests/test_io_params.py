"""CPU test of the tracked_flame_params file format: keys, shapes and loader semantics of the reference
(vhap/model/tracker.py:79-129, 1152-1218, parameter shapes tracker.py:1279-1341)."""
import numpy as np

from vhap_b200 import io_params


def _params(n_t, V=11, T=4, n_shape=300, n_expr=100, seed=0):
    r = np.random.default_rng(seed)
    f = lambda *s: r.normal(size=s).astype(np.float32).reshape(-1)          # the engine hands out flat arrays
    return {"shape": f(n_shape), "expr": f(n_t, n_expr), "rotation": f(n_t, 3), "translation": f(n_t, 3), "neck_pose": f(n_t, 3),
            "jaw_pose": f(n_t, 3), "eyes_pose": f(n_t, 6), "lights": f(9, 3), "focal_length": f(1), "static_offset": f(1, V, 3),
            "tex_extra": r.normal(size=(3, T, T)).astype(np.float32)}


def test_keys_shapes_and_name(tmp_path):
    n_t, V, T = 5, 11, 4
    p = _params(n_t, V, T)
    rep = io_params.engine_params_to_report(p, timestep_ids=[f"{i:05d}" for i in range(n_t)], n_processed_frames=n_t, image_size=(512, 384))
    path = io_params.save_tracked_flame_params(tmp_path, rep, epoch=7)
    assert path.name == "tracked_flame_params_7.npz"                                     # tracker.py:1215-1218
    assert io_params.save_tracked_flame_params(tmp_path, rep).name == "tracked_flame_params.npz"
    z = io_params.load_tracked_flame_params(path)
    expect = {"rotation": (n_t, 3), "translation": (n_t, 3), "neck_pose": (n_t, 3), "jaw_pose": (n_t, 3), "eyes_pose": (n_t, 6),
              "shape": (300,), "expr": (n_t, 100), "timestep_id": (n_t,), "n_processed_frames": (), "focal_length": (1,),
              "tex_extra": (3, T, T), "lights": (9, 3), "static_offset": (1, V, 3), "image_size": (2,)}
    assert list(z.keys()) == list(expect.keys())                                         # the reference's key order (tracker.py:1158-1213)
    for k, shp in expect.items():
        assert z[k].shape == shp, (k, z[k].shape)
    for k in ("rotation", "shape", "expr", "tex_extra", "lights", "static_offset", "focal_length"):
        assert z[k].dtype == np.float32
    assert int(z["n_processed_frames"]) == n_t and list(z["image_size"]) == [512, 384]


def test_calibrated_and_optional_keys():
    p = _params(3)
    rep = io_params.engine_params_to_report(p, [0, 1, 2], 3, (64, 64), calibrated=True, tex_extra=False, use_static_offset=False)
    assert "focal_length" not in rep and "tex_extra" not in rep and "static_offset" not in rep


def test_round_trip_and_partial_load(tmp_path):
    n_file, n_eng = 3, 5
    src = _params(n_file, seed=1)
    rep = io_params.engine_params_to_report(src, list(range(n_file)), n_file, (32, 32))
    path = io_params.save_tracked_flame_params(tmp_path, rep)
    cur = _params(n_eng, seed=2)
    warned = []
    out = io_params.report_to_engine_params(io_params.load_tracked_flame_params(path), cur, n_eng, warn=warned.append)
    assert not warned
    for k in ("rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "expr"):   # first min(N_t, len) rows (tracker.py:92-94)
        a, b, c = out[k].reshape(n_eng, -1), src[k].reshape(n_file, -1), cur[k].reshape(n_eng, -1)
        assert np.array_equal(a[:n_file], b) and np.array_equal(a[n_file:], c[n_file:])
        assert out[k].shape == cur[k].shape
    for k in ("shape", "lights", "focal_length", "static_offset", "tex_extra"):
        assert np.array_equal(out[k].reshape(-1), src[k].reshape(-1))
    # optional keys missing from the file keep the engine's value and warn like the reference (tracker.py:113-123)
    rep2 = {k: v for k, v in rep.items() if k not in ("tex_extra", "static_offset")}
    out2 = io_params.report_to_engine_params(rep2, cur, n_eng, warn=warned.append)
    assert len(warned) == 2 and np.array_equal(out2["tex_extra"], cur["tex_extra"])
