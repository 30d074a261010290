#!/usr/bin/env python
"""Build `vhap_b200/assets/flame_topology.npz` from the reference's shipped *data* assets.

Run in the authoring container only (needs /root/reference/asset/flame); the GPU box
has no /root/reference, so the output is committed.  No reference source is copied: the
fixture holds mesh topology / UVs / landmark embedding (OBJ + npy data) and region masks
recovered from `uv_masks.npz`.

Reference citations:
  * template mesh + UVs   : asset/flame/head_template_mesh.obj, loaded at vhap/model/flame.py:149
  * landmark embedding    : asset/flame/landmark_embedding_with_eyes.npy, vhap/model/flame.py:126-138
  * region masks          : asset/flame/uv_masks.npz (rendered by vhap/generate_flame_uvmask.py:25-76);
                            the licensed FLAME_masks.pkl (vhap/model/flame.py:757-772) is absent, so
                            per-vertex membership is *recovered* by probing the uv mask close to each
                            face corner (SURVEY.md §8c "Substitutes for missing assets").
  * lip rings for teeth   : the two ordered 15-vertex id lists at vhap/model/flame.py:802-814 (data).
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "vhap_b200" / "assets" / "flame_topology.npz"


def parse_obj(path):
    v, vt, f, ft = [], [], [], []
    for line in open(path):
        if line.startswith("v "):
            v.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("vt "):
            vt.append([float(x) for x in line.split()[1:3]])
        elif line.startswith("f "):
            a, b = [], []
            for tok in line.split()[1:4]:
                p = tok.split("/")
                a.append(int(p[0]) - 1)
                b.append(int(p[1]) - 1)
            f.append(a)
            ft.append(b)
    return (np.asarray(v, np.float32), np.asarray(vt, np.float32),
            np.asarray(f, np.int32), np.asarray(ft, np.int32))


def parse_id_list(src, name):
    m = re.search(r'"%s",\s*torch\.tensor\(\[\s*([0-9,\s]+)\]\)' % name, src)
    return np.asarray([int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()], np.int32)


def main():
    v, vt, f, ft = parse_obj(REF / "asset/flame/head_template_mesh.obj")
    assert v.shape == (5023, 3) and vt.shape == (5118, 2) and f.shape == (9976, 3)
    emb = np.load(REF / "asset/flame/landmark_embedding_with_eyes.npy", allow_pickle=True, encoding="latin1")[()]
    lmk_faces = np.asarray(emb["full_lmk_faces_idx"]).reshape(-1).astype(np.int32)
    lmk_bary = np.asarray(emb["full_lmk_bary_coords"]).reshape(-1, 3).astype(np.float32)

    src = open(REF / "vhap/model/flame.py").read()
    lip_up = parse_id_list(src, "lip_outside_ring_upper")
    lip_lo = parse_id_list(src, "lip_outside_ring_lower")
    assert lip_up.shape == (15,) and lip_lo.shape == (15,)

    # ---- recover per-vertex region membership from the uv masks -------------------------------
    masks = np.load(REF / "asset/flame/uv_masks.npz")
    names = sorted(masks.files)
    T = 2048
    V = v.shape[0]
    votes_yes = np.zeros((len(names), V), np.int32)
    votes_all = np.zeros(V, np.int32)
    w_main, w_oth = 0.8, 0.1
    for k in range(3):
        bary = np.full(3, w_oth, np.float32)
        bary[k] = w_main
        uv = (vt[ft] * bary[None, :, None]).sum(1)            # [F,2] probe point close to corner k
        col = np.clip((uv[:, 0] * T).astype(np.int64), 0, T - 1)
        row = np.clip(((1.0 - uv[:, 1]) * T).astype(np.int64), 0, T - 1)
        np.add.at(votes_all, f[:, k], 1)
        for i, n in enumerate(names):
            hit = masks[n][row, col]
            np.add.at(votes_yes[i], f[:, k], hit.astype(np.int32))
    member = votes_yes * 2 > votes_all[None, :]                # majority of incident corners
    bits = np.zeros(V, np.uint64)
    for i in range(len(names)):
        bits |= member[i].astype(np.uint64) << np.uint64(i)

    # uv-space mask used by reg_tex_res_clusters (tracker.py:536-539): sclerae U teeth, bit-packed
    res_mask = masks["sclerae"] | masks["teeth"]
    res_mask_packed = np.packbits(res_mask, axis=1)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(
        OUT, v_template=v, verts_uv=vt, faces=f, faces_uv=ft,
        lmk_faces_idx=lmk_faces, lmk_bary=lmk_bary,
        lip_ring_upper=lip_up, lip_ring_lower=lip_lo,
        region_names=np.asarray(names), region_bits=bits,
        uvmask_sclerae_teeth_packed=res_mask_packed,
    )
    print("wrote", OUT, OUT.stat().st_size, "bytes")
    for i, n in enumerate(names):
        print(f"  {n:40s} {int(member[i].sum()):5d} verts")


if __name__ == "__main__":
    sys.exit(main())
