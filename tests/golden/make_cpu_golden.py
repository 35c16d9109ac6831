"""Generate tests/golden/cpu_*.npz from the reference's OWN pure-torch code, imported in the
build container (needs /root/reference; the outputs are committed because the reference tree
does not exist on the GPU box).

What can run on CPU in the reference and is therefore used to pin the oracle:
  * testing/test_shencoder.py  SHEncoder_torch  -- closed-form SH, degree <= 5
  * SealNeRF/color_utils.py    rgb2hsv_torch / hsv2rgb_torch
  * SealNeRF/seal_utils.py     moller_trumbore, points_in_mesh, modify_hsv, modify_rgb, project_points,
                               SealMapper.map_mask / .map_color (texture branch), SealBBoxMapper.map_to_origin,
                               SealBrushMapper.map_to_origin, SealAnchorMapper.map_to_origin
  * activation.py              trunc_exp (forward + clamped backward)
seal_utils.py and test_shencoder.py cannot be *imported* here (pytorch3d/trimesh/json5/open3d,
CUDA-only shencoder), so the needed definitions are lifted from the reference source with
``ast`` at run time and exec'd unmodified -- nothing is copied into this repository.

Run:  python tests/golden/make_cpu_golden.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SEAL3D_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, glb):
    """exec the top-level defs/classes `names` of reference file `path` into namespace glb."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    exec(compile(mod, path, "exec"), glb)
    return glb


def synth_bbox_edit():
    """The synthetic axis-aligned bbox edit of SURVEY.md 8d: source box [-0.15,0.15]^3 moved by (0.3,0,0)."""
    lo, hi = np.array([-0.15, -0.15, -0.15]), np.array([0.15, 0.15, 0.15])
    T = np.eye(4)
    T[:3, 3] = [0.3, 0.0, 0.0]
    ang = 0.3  # a small rotation about z so the rotation path is exercised too
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    T[:3, :3] = R
    scale = np.array([1.2, 0.9, 1.0])
    center = (lo + hi) / 2
    # 8 corners of the source box -> scaled about centre -> transformed: the target mesh
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    v = (corners - center) * scale + center
    v = (T[:3, :3] @ v.T).T + T[:3, 3]
    faces = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                      [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
    tris = v[faces]
    map_data = {
        "map_bound": np.stack([v.min(0), v.max(0)]),
        "transform": np.linalg.inv(T),
        "rotation": np.linalg.inv(R),
        "scale": 1.0 / scale,
        "center": center,
        "empty_bound": np.stack([lo, hi]),
        "map_source": np.array([0.9, 0.9, 0.9]),
    }
    return map_data, tris


def main():
    sys.path.insert(0, REF)
    g = torch.Generator().manual_seed(1234)

    # ---- SH closed form ---------------------------------------------------------------
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    lift(os.path.join(REF, "testing", "test_shencoder.py"), {"SHEncoder_torch"}, ns)
    d = torch.rand(512, 3, generator=g) * 2 - 1
    d = d / (torch.norm(d, dim=-1, keepdim=True) + 1e-8)
    sh = {"dirs": d.numpy()}
    for deg in (1, 2, 3, 4, 5):
        sh["deg%d" % deg] = ns["SHEncoder_torch"](degree=deg)(d).numpy()
    np.savez_compressed(os.path.join(OUT, "cpu_sh.npz"), **sh)

    # ---- colour utils + seal proxy functions -------------------------------------------
    # SealNeRF/__init__ chain pulls heavy deps; load color_utils as a standalone module
    cu = types.ModuleType("color_utils")
    exec(compile(open(os.path.join(REF, "SealNeRF", "color_utils.py")).read(), "color_utils.py", "exec"), cu.__dict__)
    ns = {"torch": torch, "np": np, "Tuple": tuple, "Union": None,
          "rgb2hsv_torch": cu.rgb2hsv_torch, "hsv2rgb_torch": cu.hsv2rgb_torch}
    ns["Tuple"] = __import__("typing").Tuple
    ns["Union"] = __import__("typing").Union
    ns["Meshes"] = object
    lift(os.path.join(REF, "SealNeRF", "seal_utils.py"),
         {"moller_trumbore", "points_in_mesh", "modify_hsv", "modify_rgb", "convert_tensor", "project_points", "SealMapper",
          "SealBBoxMapper", "SealBrushMapper", "SealAnchorMapper"}, ns)

    rgb = torch.rand(2000, 3, generator=g)
    rgb[:50] = rgb[:50, :1]  # greys (delta == 0)
    rgb[50:60] = 0.0
    hsv = cu.rgb2hsv_torch(rgb.view(-1, 3, 1).clone()).view(-1, 3)
    back = cu.hsv2rgb_torch(hsv.view(-1, 3, 1).clone()).view(-1, 3)
    mod = torch.tensor([0.3, -0.1, 0.05])
    out_hsv = ns["modify_hsv"](rgb.clone(), mod)
    tgt = torch.tensor([0.8, 0.2, 0.1])
    out_rgb = ns["modify_rgb"](rgb.clone(), tgt, 0.05)
    np.savez_compressed(os.path.join(OUT, "cpu_color.npz"), rgb=rgb.numpy(), hsv=hsv.numpy(), back=back.numpy(),
                        mod=mod.numpy(), out_hsv=out_hsv.numpy(), target=tgt.numpy(), light=np.float32(0.05),
                        out_rgb=out_rgb.numpy())

    map_data, tris = synth_bbox_edit()
    mapper = object.__new__(ns["SealBBoxMapper"])
    ns["SealMapper"].__init__(mapper, {})
    mapper.map_data = dict(map_data)
    mapper.map_triangles = torch.from_numpy(tris)
    mapper.map_data_conversion(force=True)  # -> float32 cpu tensors, as the reference does
    pts = torch.rand(20000, 3, generator=g) * 1.2 - 0.4
    pts[:64] = 0.0            # zero padding rows of the marcher's over-allocated buffer
    pts[64:80, 1] = 0.0       # rows with one exactly-zero coordinate are excluded too
    dirs = torch.randn(20000, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    mask = mapper.map_mask(pts.clone())
    mp, md, mm = mapper.map_to_origin(pts.clone(), dirs.clone())
    assert torch.equal(mask, mm)
    in_mesh = ns["points_in_mesh"](pts, mapper.map_triangles)
    np.savez_compressed(
        os.path.join(OUT, "cpu_proxy.npz"), points=pts.numpy(), dirs=dirs.numpy(), tris=tris.astype(np.float32),
        in_mesh=in_mesh.numpy(), mask=mm.numpy(), mapped_points=mp.numpy(), mapped_dirs=md.numpy(),
        **{"md_" + k: np.asarray(v, dtype=np.float32) for k, v in map_data.items()})
    print("proxy: %d / %d inside" % (int(mm.sum()), mm.numel()))

    # ---- brush / anchor mappers and the texture colour map ------------------------------
    # the mappers' __init__ needs trimesh / skspatial / pytorch3d (absent): build map_data by hand, as for the bbox mapper
    def box_tris(lo, hi):
        c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
        return c[f]

    pts2 = torch.rand(20000, 3, generator=g) * 1.0 - 0.5
    pts2[:32] = 0.0
    # brush: two boxes, plane z = 0.05, pressure along +z, ring of border points
    b_lo = [np.array([-0.3, -0.25, -0.1]), np.array([0.1, 0.05, -0.1])]
    b_hi = [np.array([0.05, 0.2, 0.2]), np.array([0.4, 0.35, 0.2])]
    btris = np.concatenate([box_tris(l, h) for l, h in zip(b_lo, b_hi)])
    normal_expand = np.array([0.01, -0.02, 0.08])
    ang2 = np.linspace(0, 2 * np.pi, 37)[:-1]
    border = np.stack([0.22 * np.cos(ang2) - 0.1, 0.18 * np.sin(ang2), np.full_like(ang2, 0.05)], 1)
    for mode in ("linear", "dry"):
        brush = object.__new__(ns["SealBrushMapper"])
        ns["SealMapper"].__init__(brush, {})
        brush.map_data = {"map_bound": np.stack([np.stack([l, h]) for l, h in zip(b_lo, b_hi)]), "normal_expand": normal_expand,
                          "center": np.array([0.0, 0.0, 0.05]), "border_points": border, "attenuation_distance": 0.12,
                          "attenuation_mode": mode}
        brush.map_triangles = torch.from_numpy(btris)
        brush.map_test_dir = torch.from_numpy(normal_expand[None])
        brush.map_data_conversion(force=True)
        bp, _, bm = brush.map_to_origin(pts2.clone(), None)
        if mode == "linear":
            brush_lin = (bp.numpy(), bm.numpy())
        else:
            brush_dry = (bp.numpy(), bm.numpy())
    # anchor
    a_lo, a_hi = np.array([-0.35, -0.35, -0.3]), np.array([0.35, 0.35, 0.4])
    atris = box_tris(a_lo, a_hi)
    v_h = np.array([0.0, 0.02, -0.25])
    anchor = object.__new__(ns["SealAnchorMapper"])
    ns["SealMapper"].__init__(anchor, {})
    anchor.map_data = {"map_bound": np.stack([a_lo, a_hi]), "v_anchor": np.array([0.02, -0.01, 0.0]), "v_offset": np.array([0.05, 0.03, 0.0]),
                       "v_h": v_h, "len_h": float(np.linalg.norm(v_h)), "radius": 0.25, "scale": np.array([1.1, 0.9, 1.0])}
    anchor.map_triangles = torch.from_numpy(atris)
    anchor.map_data_conversion(force=True)
    ap, _, am = anchor.map_to_origin(pts2.clone(), None)
    far = pts2.clone() + 5.0          # nothing inside the map region: the early exit returns the (all-false) map mask
    fp_, _, fm = anchor.map_to_origin(far.clone(), None)
    assert not fm.any() and torch.equal(fp_, far)
    # texture colour map on 3000 colours
    img = torch.rand(24, 32, 3, generator=g)
    alpha = (torch.rand(24, 32, generator=g) > 0.3).to(torch.float32) * torch.rand(24, 32, generator=g)
    tex = object.__new__(ns["SealBrushMapper"])
    ns["SealMapper"].__init__(tex, {})
    v_o, v_w, v_hh = np.array([-0.3, -0.3, 0.05]), np.array([0.3, -0.28, 0.05]), np.array([-0.32, 0.3, 0.05])
    tex.map_data = {"image": img.numpy(), "image_mask": alpha.numpy(), "v_image_norm": np.array([0.0, 0.0, 1.0]), "v_image_o": v_o,
                    "v_image_w": v_w, "v_image_h": v_hh, "rgb_light_offset": 0.05}
    tex.map_data_conversion(force=True)
    cpts = torch.rand(3000, 3, generator=g) * 0.8 - 0.4
    ccol = torch.rand(3000, 3, generator=g)
    cout = tex.map_color(cpts.clone(), None, ccol.clone())
    np.savez_compressed(
        os.path.join(OUT, "cpu_mappers.npz"), points=pts2.numpy(),
        brush_tris=btris.astype(np.float32), brush_bounds=np.stack([np.stack([l, h]) for l, h in zip(b_lo, b_hi)]).astype(np.float32),
        brush_normal_expand=normal_expand.astype(np.float32), brush_center=np.array([0.0, 0.0, 0.05], dtype=np.float32),
        brush_border=border.astype(np.float32), brush_att=np.float32(0.12),
        brush_linear_points=brush_lin[0], brush_linear_mask=brush_lin[1], brush_dry_points=brush_dry[0], brush_dry_mask=brush_dry[1],
        anchor_tris=atris.astype(np.float32), anchor_bounds=np.stack([a_lo, a_hi]).astype(np.float32),
        anchor_v_anchor=np.array([0.02, -0.01, 0.0], dtype=np.float32), anchor_v_offset=np.array([0.05, 0.03, 0.0], dtype=np.float32),
        anchor_v_h=v_h.astype(np.float32), anchor_len_h=np.float32(np.linalg.norm(v_h)), anchor_radius=np.float32(0.25),
        anchor_scale=np.array([1.1, 0.9, 1.0], dtype=np.float32), anchor_points=ap.numpy(), anchor_mask=am.numpy(),
        tex_image=img.numpy(), tex_alpha=alpha.numpy(), tex_norm=np.array([0.0, 0.0, 1.0], dtype=np.float32), tex_o=v_o.astype(np.float32),
        tex_w=v_w.astype(np.float32), tex_h=v_hh.astype(np.float32), tex_light=np.float32(0.05), tex_points=cpts.numpy(), tex_colors=ccol.numpy(),
        tex_out=cout.numpy())
    print("brush: %d inside, anchor: %d valid" % (int(brush_lin[1].sum()), int(am.sum())))

    # ---- trunc_exp --------------------------------------------------------------------
    act = types.ModuleType("activation")
    exec(compile(open(os.path.join(REF, "activation.py")).read(), "activation.py", "exec"), act.__dict__)
    x = (torch.randn(256, generator=g) * 8).requires_grad_(True)
    y = act.trunc_exp(x)
    gy = torch.randn(256, generator=g)
    y.backward(gy)
    np.savez_compressed(os.path.join(OUT, "cpu_trunc_exp.npz"), x=x.detach().numpy(), y=y.detach().numpy(),
                        gy=gy.numpy(), gx=x.grad.numpy())
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("cpu_")))


if __name__ == "__main__":
    main()
