"""CPU oracle for the VHAP photometric inner loop  --  TEST INFRASTRUCTURE ONLY.

This package is a float64-capable PyTorch / numpy restatement of the reference's hot path
(SURVEY.md section 8a).  It exists so that tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg have something to check the CUDA engine against.  Nothing under
`vhap_b200/` imports it and the product path never routes through it.

Pinning status
  * FLAME / LBS / landmarks / camera (oracle/lbs.py, oracle/camera.py): PINNED.  Checked against
    outputs of the reference's own `vhap/model/lbs.py` and `vhap/util/mesh.py`, imported unmodified
    in the authoring container; vectors in tests/golden/lbs_golden.npz, generator
    tests/golden/make_golden.py.
  * FlameHead.forward / FlameTracker.forward_flame composition (oracle/lbs.py flame_forward): PINNED against the reference's own
    methods run on bare instances with this repo's synthetic buffers; tests/golden/flame_golden.npz, make_flame_golden.py.
  * camera transforms, vertex normals, SH shading, detach_by_indices (oracle/camera.py, oracle/render.py
    compute_v_normals / sh_shading): PINNED against outputs and autograd gradients of the reference's own
    `vhap/util/render_nvdiffrast.py` methods (imported unmodified, a stub standing in for its absent nvdiffrast
    import); vectors in tests/golden/render_golden.npz, generator tests/golden/make_render_golden.py.
  * landmark energy and every regulariser (oracle/energy.py lmk_energy / regularization_energy): PINNED against values and
    gradients of the reference's own FlameTracker.compute_lmk_energy / compute_regularization_energy (tracker.py:347-389,
    480-690) run on a bare instance; tests/golden/energy_golden.npz, generator tests/golden/make_energy_golden.py.
  * photometric glue around the renderer (compute_energy: background modes, UV flip, align_*_except look-ups, disturbance flag,
    L1 normalisation + gradient): PINNED against the reference's FlameTracker.compute_photometric_energy run with a recording fake
    renderer; tests/golden/photo_golden.npz, generator tests/golden/make_photo_golden.py.
  * loss weights / learning rates / stage table (vhap_b200/config.py): PINNED against the reference's dataclasses
    (tests/golden/config_golden.json, generator tests/golden/make_config_golden.py).
  * render_rgba minus the nvdiffrast ops (oracle/render.py render_rgba / disturb): PINNED against the reference's own
    NVDiffRenderer.render_rgba run end to end with its dr.* calls served by this oracle's op restatements and its random draws
    injected (values + gradients); tests/golden/rgba_golden.npz, generator tests/golden/make_rgba_golden.py.
  * the whole energy end to end (oracle/energy.py compute_energy: all terms, total, gradients w.r.t. every parameter): PINNED, modulo
    the four nvdiffrast ops, against the reference's own FlameTracker.compute_energy run on CPU with all of its code unmodified
    (tracker, FlameHead, lbs, NVDiffRenderer) and only dr.rasterize / interpolate / texture / antialias served by this oracle;
    tests/golden/e2e_golden.npz, generator tests/golden/make_e2e_golden.py.
  * the nvdiffrast ops themselves -- rasterise / interpolate / texture / antialias (oracle/raster.py, oracle/render.py):
    PARITY UNPINNED.  The arithmetic lives in the third-party dependency `nvdiffrast`
    (ShenhanQian/nvdiffrast@backface-culling, pinned by branch name only at
    /root/reference/pyproject.toml:30) whose source is absent from /root/reference and which
    cannot be installed here (no network, no GPU).  The reference has no tests or golden vectors
    for this path.  These modules restate nvdiffrast's *published* algorithm
    (Laine et al. 2020, "Modular Primitives for High-Performance Differentiable Rendering") at the
    reference's call sites (vhap/util/render_nvdiffrast.py:254,384,389,399,465); the exact
    fixed-point snapping / fill rule / depth tie-break are this repo's own specification
    (DESIGN.md "Rasteriser specification") and triangle-id bit-exactness is claimed against
    this oracle only.
"""
