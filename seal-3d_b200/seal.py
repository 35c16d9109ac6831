"""Seal-3D proxy side: bbox / brush / anchor mappers with their colour edits, teacher/student renderers (mirror of
the hot-path parts of ``SealNeRF/seal_utils.py`` and ``SealNeRF/renderer.py``).

The mappers take ready ``map_data`` tensors (the dicts the reference constructors build, seal_utils.py:222-236,
:362-402, :491-512) plus the target-mesh triangles; building them from ``seal.json`` needs trimesh / pytorch3d /
skspatial and is out of scope.  Use ``synth.bbox_edit`` for an axis-aligned or rotated box edit.
"""
import numpy as np
import torch

from . import _lib
from .network import NeRFNetwork
from .tensorf import TensoRFNetwork


class _ColorEdits:
    """colour side of SealMapper.map_color (seal_utils.py:48-81), shared by the three mappers: hsv shift, rgb replacement,
    texture image -- applied in the reference's order"""

    def _init_color(self, md):
        self._stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._h_hsv = _lib.host_f32(md["hsv"]) if "hsv" in md else None
        self._h_rgb = _lib.host_f32(md["rgb"]) if "rgb" in md else None
        self._light = float(md["rgb_light_offset"]) if "rgb_light_offset" in md else 0.0
        self._image = None
        if "image" in md:
            self._image = torch.from_numpy(np.ascontiguousarray(md["image"], dtype=np.float32)).to(self.device)
            self._alpha = torch.from_numpy(np.ascontiguousarray(md["image_mask"], dtype=np.float32)).to(self.device)
            self._h_img = [_lib.host_f32(md[k]) for k in ("v_image_norm", "v_image_o", "v_image_w", "v_image_h")]

    def has_color_edit(self):
        return self._h_hsv is not None or self._h_rgb is not None or self._image is not None

    def map_color_(self, rgbs, mask, points=None):
        """in-place colour edit of the masked rows of rgbs [M,3] (float32, contiguous); `points` [M,3] = the (mapped) sample
        positions, needed by the texture edit only"""
        if not self.has_color_edit():
            return rgbs
        m8 = mask.to(torch.uint8).contiguous() if mask is not None else None
        if self._h_hsv is not None or self._h_rgb is not None:
            _lib.call("s3d_seal_map_color", rgbs, m8, rgbs.shape[0], self._h_hsv[1] if self._h_hsv else None,
                      self._h_rgb[1] if self._h_rgb else None, self._light, self._stats)
        if self._image is not None:
            if points is None:
                raise ValueError("the texture colour map needs the sample positions")
            _lib.call("s3d_seal_map_color_image", rgbs, points.contiguous().float(), m8, rgbs.shape[0], self._image, self._alpha,
                      self._image.shape[0], self._image.shape[1], self._h_img[0][1], self._h_img[1][1], self._h_img[2][1], self._h_img[3][1],
                      self._light, self._stats)
        return rgbs

    def map_color(self, points, dirs, colors):
        """seal_utils.py:48-81: returns the edited copy of `colors` [P,3]"""
        out = colors.detach().float().contiguous().clone()
        return self.map_color_(out, None, points)


class SealBBoxMapper(_ColorEdits):
    def __init__(self, map_data, triangles, test_dir=None, device="cuda"):
        self.device = torch.device(device)
        md = {k: np.asarray(v, dtype=np.float32) for k, v in map_data.items()}
        self.map_data = {k: torch.from_numpy(v).to(self.device) for k, v in md.items()}
        self._h = {k: _lib.host_f32(md[k]) for k in ("transform", "rotation", "scale", "center")}
        self._h_test = _lib.host_f32(test_dir) if test_dir is not None else None
        self._h_src = _lib.host_f32(md["empty_bound"]) if "map_source" in md else None
        self._h_ms = _lib.host_f32(md["map_source"]) if "map_source" in md else None
        self.bounds = torch.from_numpy(md["map_bound"].reshape(-1, 2, 3)).to(self.device).contiguous()
        self.map_triangles = torch.from_numpy(np.asarray(triangles, np.float32).reshape(-1, 3, 3)).to(self.device).contiguous()
        self._init_color(md)

    def map_to_origin(self, points, dirs=None):
        """seal_utils.py:237-279 -> (points', dirs', mask) ; clones with the masked rows replaced."""
        points = points.contiguous().float()
        P = points.shape[0]
        dirs_c = dirs.contiguous().float() if dirs is not None else None
        out_p = torch.empty_like(points)
        out_d = torch.empty_like(dirs_c) if dirs_c is not None else None
        mask = torch.empty(P, dtype=torch.uint8, device=points.device)
        _lib.call("s3d_seal_bbox_map_to_origin", points, dirs_c, P, self._h["transform"][1], self._h["rotation"][1],
                  self._h["scale"][1], self._h["center"][1], self.bounds, self.bounds.shape[0], self.map_triangles,
                  self.map_triangles.shape[0], self._h_test[1] if self._h_test else None,
                  self._h_src[1] if self._h_src else None, self._h_ms[1] if self._h_ms else None, out_p, out_d, mask)
        return out_p, out_d, mask.bool()

    def map_mask(self, points):
        return self.map_to_origin(points, None)[2]


class SealBrushMapper(_ColorEdits):
    """SealNeRF/seal_utils.py:282-453 with ready `map_data` (the constructor's mesh fitting needs trimesh / skspatial /
    pytorch3d, which are outside the hot path): keys map_bound [B,2,3], normal_expand, center, border_points [K,3],
    attenuation_distance, attenuation_mode ('linear' | 'dry'), optional force_fill_bound, hsv / rgb / rgb_light_offset /
    image, image_mask, v_image_norm / _o / _w / _h.  The mesh test uses the brush normal as its ray direction (:357)."""

    def __init__(self, map_data, triangles, device="cuda"):
        self.device = torch.device(device)
        self.mode = {"linear": 0, "dry": 1}.get(map_data.get("attenuation_mode", "linear"))
        if self.mode is None:
            raise NotImplementedError("attenuation_mode %r (seal_utils.py:433-438 raises too)" % map_data.get("attenuation_mode"))
        md = {k: np.asarray(v, dtype=np.float32) for k, v in map_data.items() if k != "attenuation_mode"}
        self.map_data = {k: torch.from_numpy(v).to(self.device) for k, v in md.items()}
        self.map_data.setdefault("force_fill_bound", self.map_data["map_bound"])
        self._h = {k: _lib.host_f32(md[k]) for k in ("normal_expand", "center")}
        self._att = float(md["attenuation_distance"])
        self.bounds = torch.from_numpy(md["map_bound"].reshape(-1, 2, 3)).to(self.device).contiguous()
        self.map_triangles = torch.from_numpy(np.asarray(triangles, np.float32).reshape(-1, 3, 3)).to(self.device).contiguous()
        self.border = torch.from_numpy(md["border_points"].reshape(-1, 3)).to(self.device).contiguous()
        self._init_color(md)

    def map_to_origin(self, points, dirs=None):
        """seal_utils.py:408-453 -> (points', dirs (untouched), mask)"""
        points = points.contiguous().float()
        P = points.shape[0]
        out_p = torch.empty_like(points)
        mask = torch.empty(P, dtype=torch.uint8, device=points.device)
        _lib.call("s3d_seal_brush_map_to_origin", points, P, self.bounds, self.bounds.shape[0], self.map_triangles, self.map_triangles.shape[0],
                  self._h["normal_expand"][1], self._h["normal_expand"][1], self._h["center"][1], self.border, self.border.shape[0], self._att,
                  self.mode, out_p, mask)
        return out_p, dirs, mask.bool()

    def map_mask(self, points):
        return self.map_to_origin(points, None)[2]


class SealAnchorMapper(_ColorEdits):
    """SealNeRF/seal_utils.py:456-570 with ready `map_data`: map_bound, v_anchor, v_offset, v_h, len_h, radius, scale
    (+ optional force_fill_bound and the colour keys)."""

    def __init__(self, map_data, triangles, test_dir=None, device="cuda"):
        self.device = torch.device(device)
        md = {k: np.asarray(v, dtype=np.float32) for k, v in map_data.items()}
        self.map_data = {k: torch.from_numpy(v).to(self.device) for k, v in md.items()}
        self.map_data.setdefault("force_fill_bound", self.map_data["map_bound"])
        self._h = {k: _lib.host_f32(md[k]) for k in ("v_anchor", "v_offset", "v_h", "scale")}
        self._h_test = _lib.host_f32(test_dir) if test_dir is not None else None
        self._len_h, self._radius = float(md["len_h"]), float(md["radius"])
        self.bounds = torch.from_numpy(md["map_bound"].reshape(-1, 2, 3)).to(self.device).contiguous()
        self.map_triangles = torch.from_numpy(np.asarray(triangles, np.float32).reshape(-1, 3, 3)).to(self.device).contiguous()
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._init_color(md)

    def map_to_origin(self, points, dirs=None):
        """seal_utils.py:514-570 -> (points', dirs (untouched), cone mask)"""
        points = points.contiguous().float()
        P = points.shape[0]
        out_p = torch.empty_like(points)
        mask = torch.empty(P, dtype=torch.uint8, device=points.device)
        _lib.call("s3d_seal_anchor_map_to_origin", points, P, self.bounds, self.bounds.shape[0], self.map_triangles, self.map_triangles.shape[0],
                  self._h_test[1] if self._h_test else None, self._h["v_anchor"][1], self._h["v_offset"][1], self._h["v_h"][1], self._len_h,
                  self._radius, self._h["scale"][1], self._flag, out_p, mask)
        return out_p, dirs, mask.bool()


class SealRendererMixin:
    """SealNeRF/renderer.py:8-74: mapper + occupancy force-fill of the edit region."""

    seal_mapper = None
    density_bitfield_origin = None
    density_bitfield_hacked = False

    def init_mapper(self, mapper):
        self.seal_mapper = mapper
        bounds = mapper.map_data["force_fill_bound"].detach().cpu().numpy().reshape(-1, 2, 3).copy()
        lo = np.maximum(bounds[:, 0, :], -self.bound)
        hi = np.minimum(bounds[:, 1, :], self.bound)
        H = self.grid_size
        self._fill_boxes = []
        for i in range(bounds.shape[0]):
            cmin = np.floor(((lo[i] + self.bound) / self.bound / 2) * H).astype(np.int32)
            cmax = np.floor(((hi[i] + self.bound) / self.bound / 2) * H).astype(np.int32)
            self._fill_boxes.append((_lib.host_i32(np.clip(cmin, 0, H)), _lib.host_i32(np.clip(cmax, 0, H))))

    @torch.no_grad()
    def hack_bitfield(self):
        if self.density_bitfield_origin is None:
            self.density_bitfield_origin = self.density_bitfield.clone()
        for lo, hi in self._fill_boxes:
            _lib.call("s3d_seal_force_fill_bitfield", self.density_bitfield, lo[1], hi[1], self.grid_size, 0)
        self.density_bitfield_hacked = True

    @torch.no_grad()
    def restore_bitfield(self):
        self.density_bitfield.copy_(self.density_bitfield_origin)
        self.density_bitfield_hacked = False

    def update_extra_state(self, decay=0.95, S=128, seed=None):
        super().update_extra_state(decay, S, seed)
        self.density_bitfield_origin = None
        if self.seal_mapper is not None:
            self.hack_bitfield()


class SealTeacherMixin:
    """SealNeRFTeacherRenderer (SealNeRF/renderer.py:77-418): samples are mapped to the original space before the
    field query and the colours of mapped samples are edited afterwards."""

    def _map_samples(self, xyzs, dirs):
        if self.seal_mapper is None:
            return xyzs, dirs, None
        return self.seal_mapper.map_to_origin(xyzs.view(-1, 3), dirs.view(-1, 3))

    def _map_colors(self, xyzs, dirs, rgbs, mask):
        if self.seal_mapper is None or not self.seal_mapper.has_color_edit():
            return rgbs
        rgbs = rgbs.float().contiguous()
        return self.seal_mapper.map_color_(rgbs, mask, xyzs.view(-1, 3) if xyzs is not None else None)


# SealNeRF/network.py:7-50 get_network(backbone, character): the four backbone x character classes
class TeacherNetwork(SealTeacherMixin, SealRendererMixin, NeRFNetwork):
    """NeRFNetwork_NGP_Teacher"""


class StudentNetwork(SealRendererMixin, NeRFNetwork):
    """NeRFNetwork_NGP_Student -- SealNeRFStudentRenderder (SealNeRF/renderer.py:421-424): plain renderer + the force-filled bitfield."""


class TensoRFTeacherNetwork(SealTeacherMixin, SealRendererMixin, TensoRFNetwork):
    """NeRFNetwork_TensoRF_Teacher (main_SealTensoRF.py:173-183)"""


class TensoRFStudentNetwork(SealRendererMixin, TensoRFNetwork):
    """NeRFNetwork_TensoRF_Student (main_SealTensoRF.py:187-197)"""
