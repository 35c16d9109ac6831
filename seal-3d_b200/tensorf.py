"""TensoRF (vector-matrix decomposition) backbone used by ``main_SealTensoRF.py`` -- mirror of ``tensoRF/network.py``
on top of the renderer mirror (BASELINE config 5).

Same constructor arguments, parameter names and state-dict SHAPES as the reference (``sigma_mat.i [1,R,H,W]``,
``sigma_vec.i [1,R,D,1]``, ``color_mat.i``, ``color_vec.i``, ``basis_mat.weight [27,144]``, ``color_net.l.weight``),
same methods (``get_sigma_feat``, ``get_color_feat``, ``forward``, ``density``, ``color``, ``density_loss``,
``upsample_model``, ``shrink_model``, ``get_params``).  What differs is the memory format: the factor images are kept in
torch's ``channels_last`` format, i.e. physically ``[H,W,R]`` / ``[D,R]``, which is what the VM kernels
(csrc/tensorf.cu) read -- one 128-bit load per tap per 4 channels instead of the reference's twelve
``F.grid_sample`` calls with R scattered scalar gathers per tap.  ``load_state_dict`` of a reference checkpoint copies into
that format; ``state_dict()`` returns the reference shapes.

Under fp16 autocast (the reference's ``--fp16``) ``basis_mat`` and the colour MLP (150-128-128-3) run on the tcgen05 kernels
of csrc/ffmlp_wide.cu (``_linear_tc``, ``_mlp_head_tc``: fp16 operands, fp32 accumulation, like the autocast ``F.linear`` GEMMs
of tensoRF/network.py:148, :172-178 they replace); in fp32 they are ``F.linear`` like in the reference.  The background model
(``bg_radius > 0``) is not built.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib
from . import raymarching
from .network import trunc_exp, get_encoder
from .renderer import NeRFRenderer


def _channel_last(t):
    """[1,R,H,W] tensor -> the same values in channels_last strides (physical [H,W,R]); size-1 dims make torch's own
    ``contiguous(memory_format=...)`` ambiguous, so the strides are set explicitly"""
    _, R, H, W = t.shape
    out = torch.empty_strided((1, R, H, W), (H * W * R, 1, W * R, R), dtype=t.dtype, device=t.device)
    out.copy_(t)
    return out


def _is_channel_last(t):
    _, R, H, W = t.shape
    want = (H * W * R, 1, W * R, R)
    return all(t.shape[i] == 1 or t.stride(i) == want[i] for i in range(4))


class _vm_lookup(Function):
    """six lookups + products (+ sums) in one launch; gradients flow to the six factor images only (x does not need one on
    the hot path, exactly like the reference's detached sample positions)"""

    @staticmethod
    def forward(ctx, x, aabb, reduce, *imgs):
        """reduce: True -> [M] fp32 (sigma), False -> [M, 3R] fp32, "half" -> [M, 3R] fp16 (colour features of an fp16 step)"""
        x = x.contiguous().float()
        _lib.check_cuda(x)
        for im in imgs:
            if not _is_channel_last(im) or im.dtype != torch.float32:
                raise _lib.S3DError("VM factor images must be float32 in channels_last layout (use TensoRFNetwork's parameters)")
        mats, vecs = imgs[:3], imgs[3:]
        R = mats[0].shape[1]
        dims = [[m.shape[2], m.shape[3], v.shape[2]] for m, v in zip(mats, vecs)]
        hd = _lib.host_i32(dims)
        M = x.shape[0]
        if reduce == "half":
            out = torch.empty(M, 3 * R, dtype=torch.float16, device=x.device)
            _lib.call("s3d_vm_forward_f16", x, M, aabb, *mats, *vecs, hd[1], R, out)
        else:
            out = torch.empty(M if reduce else (M, 3 * R), dtype=torch.float32, device=x.device)
            _lib.call("s3d_vm_forward", x, M, aabb, *mats, *vecs, hd[1], R, 1 if reduce else 0, out)
        ctx.save_for_backward(x, aabb, *imgs)
        ctx.cfg = (hd, R, reduce)
        return out

    @staticmethod
    def backward(ctx, g):
        x, aabb, *imgs = ctx.saved_tensors
        hd, R, reduce = ctx.cfg
        grads = [torch.zeros_like(im) for im in imgs]      # preserve_format: channels_last like the parameter
        if reduce == "half":
            _lib.call("s3d_vm_backward_f16", x, x.shape[0], aabb, *imgs, hd[1], R, g.contiguous().to(torch.float16), *grads)
        else:
            _lib.call("s3d_vm_backward", x, x.shape[0], aabb, *imgs, hd[1], R, 1 if reduce else 0, g.contiguous().float(), *grads)
        return (None, None, None, *grads)


class _linear_tc(Function):
    """bias-free nn.Linear on tensor cores (s3d_linear_*, csrc/ffmlp_wide.cu): fp16 operands, fp32 accumulation -- what the
    reference's autocast F.linear does through cuBLAS (tensoRF/network.py:155 basis_mat)"""

    @staticmethod
    def forward(ctx, x, weight):
        """-> y fp16 [B, out_pad]: out_pad = out rounded up to 16 (zero weight rows), so rows are 32-byte aligned vector stores;
        the caller uses the first `out` columns"""
        cout = weight.shape[0]
        out_pad = (cout + 15) // 16 * 16
        x16 = x.detach().to(torch.float16).contiguous()
        w16 = F.pad(weight.detach().to(torch.float16), (0, 0, 0, out_pad - cout)).contiguous()
        B, cin = x16.shape
        y = torch.empty(B, out_pad, dtype=torch.float16, device=x.device)
        _lib.call("s3d_linear_forward", x16, w16, B, cin, out_pad, y)
        ctx.save_for_backward(x16, w16)
        ctx.cfg = (x.dtype, cout)
        return y

    @staticmethod
    def backward(ctx, g):
        x16, w16 = ctx.saved_tensors
        in_dtype, cout = ctx.cfg
        B, cin = x16.shape
        g16 = g.to(torch.float16).contiguous()
        gx = torch.empty_like(x16) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w16)
        _lib.call("s3d_linear_backward", g16, x16, w16, B, cin, w16.shape[0], gx, gw)
        return (gx.to(in_dtype) if gx is not None else None), gw[:cout].float()


class _head_encode(Function):
    """h = cat([encoder(feat), encoder_dir(d)]) (both `frequency`, tensoRF/network.py:170-172) zero padded to a multiple of 16
    columns, fp16, in one launch (s3d_tensorf_head_encode); gradient flows to feat only (directions are data)"""

    @staticmethod
    def forward(ctx, feat16, d, deg, Fd):
        """feat16 [B, ld] fp16 with Fd valid columns (the padded rows of _linear_tc)"""
        feat16 = feat16.contiguous()
        B, ld = feat16.shape
        K = ((Fd + 3) * (1 + 2 * deg) + 15) // 16 * 16
        h = torch.empty(B, K, dtype=torch.float16, device=feat16.device)
        _lib.call("s3d_tensorf_head_encode", feat16, Fd, ld, d.contiguous().float(), B, deg, K, h)
        ctx.save_for_backward(h)
        ctx.cfg = (Fd, ld, deg, K)
        return h

    @staticmethod
    def backward(ctx, gh):
        (h,) = ctx.saved_tensors
        Fd, ld, deg, K = ctx.cfg
        B = h.shape[0]
        gf = torch.empty(B, ld, dtype=torch.float16, device=h.device)
        _lib.call("s3d_tensorf_head_encode_backward", gh.to(torch.float16).contiguous(), h, Fd, ld, B, deg, K, gf)
        return gf, None, None, None


class _mlp_head_tc(Function):
    """The colour MLP of tensoRF/network.py:53-67, 172-178 (bias-free Linear + ReLU chain, any hidden width the FFMLP kernels
    take) as ONE fused tensor-core launch per direction: the nn.Linear matrices are packed into the flat fp16 layout of
    ffmlp/ffmlp.py:121-122 (input columns zero padded to a multiple of 16, output rows to 16), s3d_ffmlp_forward / _backward do
    the rest; gradients are unpacked back onto the separate nn.Linear weights, so state dicts keep the reference's shapes."""

    @staticmethod
    def forward(ctx, h, *weights):
        B, cin = h.shape
        hidden, nl, cout = weights[0].shape[0], len(weights) - 1, weights[-1].shape[0]
        in_pad, out_pad = (cin + 15) // 16 * 16, 16
        w16 = [w.detach().to(torch.float16) for w in weights]
        flat = torch.cat([F.pad(w16[0], (0, in_pad - cin)).reshape(-1)] + [w.reshape(-1) for w in w16[1:-1]] +
                         [F.pad(w16[-1], (0, 0, 0, out_pad - cout)).reshape(-1)]).contiguous()
        x16 = h.detach()
        if x16.dtype != torch.float16 or in_pad != cin or not x16.is_contiguous():
            x16 = F.pad(x16.to(torch.float16), (0, in_pad - cin)).contiguous()
        out = torch.empty(B, out_pad, dtype=torch.float16, device=h.device)
        if not any(ctx.needs_input_grad):          # the frozen teacher: no activation buffers to write
            _lib.call("s3d_ffmlp_inference", x16, flat, B, in_pad, out_pad, hidden, nl, 0, 6, None, out)
            return out[:, :cout]
        fb = torch.empty(nl, B, hidden, dtype=torch.float16, device=h.device)
        _lib.call("s3d_ffmlp_forward", x16, flat, B, in_pad, out_pad, hidden, nl, 0, 6, fb, out)
        ctx.save_for_backward(x16, flat, fb)
        ctx.cfg = (cin, in_pad, cout, out_pad, hidden, nl, h.dtype)
        return out[:, :cout]

    @staticmethod
    def backward(ctx, g):
        x16, flat, fb = ctx.saved_tensors
        cin, in_pad, cout, out_pad, hidden, nl, in_dtype = ctx.cfg
        B = x16.shape[0]
        g16 = F.pad(g.to(torch.float16), (0, out_pad - cout)).contiguous()
        bb = torch.empty(nl, B, hidden, dtype=torch.float16, device=g.device)
        gx = torch.empty(B, in_pad, dtype=torch.float16, device=g.device)
        gw = torch.empty_like(flat)
        _lib.call("s3d_ffmlp_backward", g16, x16, flat, fb, B, in_pad, out_pad, hidden, nl, 0, 6, 1, bb, gx, gw)
        grads, off = [], 0
        grads.append(gw[off:off + hidden * in_pad].view(hidden, in_pad)[:, :cin].float()); off += hidden * in_pad
        for _ in range(nl - 1):
            grads.append(gw[off:off + hidden * hidden].view(hidden, hidden).float()); off += hidden * hidden
        grads.append(gw[off:off + out_pad * hidden].view(out_pad, hidden)[:cout].float())
        return ((gx if (cin == in_pad and in_dtype == torch.float16) else gx[:, :cin].to(in_dtype)), *grads)


class TensoRFNetwork(NeRFRenderer):
    mat_ids = [[0, 1], [0, 2], [1, 2]]
    vec_ids = [2, 1, 0]

    def __init__(self, resolution=[128] * 3, sigma_rank=[16] * 3, color_rank=[48] * 3, bg_resolution=[512, 512], bg_rank=8,
                 color_feat_dim=27, num_layers=3, hidden_dim=128, num_layers_bg=2, hidden_dim_bg=64, bound=1, **kwargs):
        super().__init__(bound, **kwargs)
        if self.bg_radius > 0:
            raise NotImplementedError("the TensoRF background model (tensoRF/network.py:68-97) is not built")
        if len(set(sigma_rank)) != 1 or len(set(color_rank)) != 1:
            raise NotImplementedError("the VM kernels take one rank per field (the reference default [16]*3 / [48]*3)")
        self.resolution = list(resolution)
        self.sigma_rank, self.color_rank, self.color_feat_dim = sigma_rank, color_rank, color_feat_dim
        self.sigma_mat, self.sigma_vec = self.init_one_svd(sigma_rank, self.resolution)
        self.color_mat, self.color_vec = self.init_one_svd(color_rank, self.resolution)
        self.basis_mat = nn.Linear(sum(color_rank), color_feat_dim, bias=False)
        self.num_layers, self.hidden_dim = num_layers, hidden_dim
        self.encoder, enc_dim = get_encoder("frequency", input_dim=color_feat_dim, multires=2)
        self.encoder_dir, enc_dim_dir = get_encoder("frequency", input_dim=3, multires=2)
        self.in_dim = enc_dim + enc_dim_dir
        self.color_net = nn.ModuleList([
            nn.Linear(self.in_dim if l == 0 else hidden_dim, 3 if l == num_layers - 1 else hidden_dim, bias=False)
            for l in range(num_layers)])
        self.bg_net = None

    def init_one_svd(self, n_component, resolution, scale=0.1):
        """tensoRF/network.py:101-112 (same shapes, same randn order), stored channels_last"""
        mat, vec = [], []
        for i in range(3):
            vec_id = self.vec_ids[i]
            m0, m1 = self.mat_ids[i]
            mat.append(nn.Parameter(_channel_last(scale * torch.randn((1, n_component[i], resolution[m1], resolution[m0])))))
            vec.append(nn.Parameter(_channel_last(scale * torch.randn((1, n_component[i], resolution[vec_id], 1)))))
        return nn.ParameterList(mat), nn.ParameterList(vec)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # a shrunk / upsampled reference checkpoint has other resolutions than the constructor's: re-allocate first.  Like
        # upsample_model / shrink_model this replaces nn.Parameter objects, so a trainer built before must be rebuilt
        # afterwards (the reference re-creates its optimizer at the same points, tensoRF/utils.py:126-128)
        for name in ("sigma_mat", "sigma_vec", "color_mat", "color_vec"):
            plist = getattr(self, name)
            for i in range(3):
                key = "%s%s.%d" % (prefix, name, i)
                if key in state_dict and tuple(state_dict[key].shape) != tuple(plist[i].shape):
                    plist[i] = nn.Parameter(_channel_last(torch.zeros(state_dict[key].shape, dtype=plist[i].dtype, device=plist[i].device)))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        if prefix + "sigma_mat.0" in state_dict:
            s = self.sigma_mat
            self.resolution = [s[0].shape[3], s[0].shape[2], s[1].shape[2]]

    def _lookup(self, x, mats, vecs, reduce, aabb=None):
        return _vm_lookup.apply(x, aabb, reduce, *mats, *vecs)

    def get_sigma_feat(self, x, aabb=None):
        """tensoRF/network.py:115-135: x [N,3] in [-1,1] -> [N]"""
        return self._lookup(x, self.sigma_mat, self.sigma_vec, True, aabb)

    def get_color_feat(self, x, aabb=None):
        """tensoRF/network.py:138-158: x [N,3] in [-1,1] -> [N, color_feat_dim]"""
        return self.basis_mat(self._lookup(x, self.color_mat, self.color_vec, False, aabb))

    def _color_mlp(self, x, d, aabb):
        if torch.is_autocast_enabled() and x.is_cuda and self.hidden_dim in (16, 32, 64, 128, 256) and self.num_layers >= 3:
            # fp16 step (the reference's --fp16): basis_mat and the colour MLP on the tcgen05 kernels instead of cuBLAS GEMMs
            feat = _linear_tc.apply(self._lookup(x, self.color_mat, self.color_vec, "half", aabb), self.basis_mat.weight)
            with torch.autocast("cuda", enabled=False):
                deg = getattr(self.encoder, "degree", None)
                if deg is not None and deg == getattr(self.encoder_dir, "degree", None):
                    h = _head_encode.apply(feat, d, deg, self.color_feat_dim)   # [N, 160] fp16: both encodings + padding in one launch
                    w0 = self.color_net[0].weight
                    w0 = F.pad(w0, (0, h.shape[1] - w0.shape[1]))       # the padded columns meet zero weights
                    out = _mlp_head_tc.apply(h, w0, *[lin.weight for lin in self.color_net[1:]])
                else:
                    h = torch.cat([self.encoder(feat[:, :self.color_feat_dim].float()), self.encoder_dir(d.float())], dim=-1)
                    out = _mlp_head_tc.apply(h, *[lin.weight for lin in self.color_net])
                return torch.sigmoid(out.float())
        h = torch.cat([self.encoder(self.get_color_feat(x, aabb)), self.encoder_dir(d)], dim=-1)
        for l in range(self.num_layers):
            h = self.color_net[l](h)
            if l != self.num_layers - 1:
                h = F.relu(h, inplace=True)
        return torch.sigmoid(h)

    def forward(self, x, d):
        """tensoRF/network.py:161-191; the aabb normalisation (:166) happens inside the lookup kernel"""
        x = x.reshape(-1, 3)
        sigma = trunc_exp(self.get_sigma_feat(x, self.aabb_train))
        return sigma, self._color_mlp(x, d.reshape(-1, 3), self.aabb_train)

    def density(self, x):
        return {"sigma": trunc_exp(self.get_sigma_feat(x.reshape(-1, 3), self.aabb_train))}

    def color(self, x, d, mask=None, **kwargs):
        """tensoRF/network.py:222-257 (masked inference)"""
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            x, d = x[mask], d[mask]
        h = self._color_mlp(x, d, self.aabb_train)
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
            return rgbs
        return h

    def density_loss(self):
        """tensoRF/network.py:261-265: L1 penalty on the sigma factors"""
        loss = 0
        for i in range(3):
            loss = loss + torch.mean(torch.abs(self.sigma_mat[i])) + torch.mean(torch.abs(self.sigma_vec[i]))
        return loss

    @torch.no_grad()
    def upsample_params(self, mat, vec, resolution):
        """tensoRF/network.py:268-275: bilinear, align_corners=True"""
        for i in range(3):
            vec_id = self.vec_ids[i]
            m0, m1 = self.mat_ids[i]
            for plist, (H2, W2) in ((mat, (resolution[m1], resolution[m0])), (vec, (resolution[vec_id], 1))):
                src = plist[i].data
                _, R, H, W = src.shape
                dst = _channel_last(torch.zeros(1, R, H2, W2, dtype=src.dtype, device=src.device))
                _lib.call("s3d_vm_resize", src, H, W, dst, H2, W2, R)
                plist[i] = nn.Parameter(dst)

    @torch.no_grad()
    def upsample_model(self, resolution):
        self.upsample_params(self.sigma_mat, self.sigma_vec, resolution)
        self.upsample_params(self.color_mat, self.color_vec, resolution)
        self.resolution = list(resolution)

    @torch.no_grad()
    def shrink_model(self):
        """tensoRF/network.py:285-322: crop the factors (and aabb_train) to the occupied part of the coarsest density grid"""
        half_grid_size = self.bound / self.grid_size
        thresh = min(self.density_thresh, self.mean_density)
        valid_grid = self.density_grid[self.cascade - 1] > thresh
        valid_pos = raymarching.morton3D_invert(torch.nonzero(valid_grid))
        valid_pos = (2 * valid_pos / (self.grid_size - 1) - 1) * (self.bound - half_grid_size)
        min_pos = valid_pos.amin(0) - half_grid_size
        max_pos = valid_pos.amax(0) + half_grid_size
        reso = torch.LongTensor(self.resolution).to(self.aabb_train.device)
        units = (self.aabb_train[3:] - self.aabb_train[:3]) / reso
        tl = torch.round((min_pos - self.aabb_train[:3]) / units).long().clamp(min=0)
        br = torch.minimum(torch.round((max_pos - self.aabb_train[:3]) / units).long(), reso)
        tl, br = tl.tolist(), br.tolist()
        for i in range(3):
            vec_id = self.vec_ids[i]
            m0, m1 = self.mat_ids[i]
            for v in (self.sigma_vec, self.color_vec):
                v[i] = nn.Parameter(_channel_last(v[i].data[..., tl[vec_id]:br[vec_id], :]))
            for m in (self.sigma_mat, self.color_mat):
                m[i] = nn.Parameter(_channel_last(m[i].data[..., tl[m1]:br[m1], tl[m0]:br[m0]]))
        self.aabb_train = torch.cat([min_pos, max_pos], dim=0)   # self.resolution stays, like the reference: upsample_model follows
        return tl, br

    def get_params(self, lr1, lr2=None):
        """tensoRF/network.py:326-340 (group order = gradient-arena order)"""
        lr2 = lr1 if lr2 is None else lr2
        return [{"params": self.sigma_mat, "lr": lr1}, {"params": self.sigma_vec, "lr": lr1},
                {"params": self.color_mat, "lr": lr1}, {"params": self.color_vec, "lr": lr1},
                {"params": self.basis_mat.parameters(), "lr": lr2}, {"params": self.color_net.parameters(), "lr": lr2}]
