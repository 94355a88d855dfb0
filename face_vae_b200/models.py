"""VAE bottleneck and the anchor model of the hot path.

``flatten_vae_nl`` keeps the reference's class name, forward signature and return convention (reference
models.py:525-570).  ``FaceVAE`` is the composition SURVEY.md section 8 defines from reference patterns (the reference
has no single image->image VAE class): EFE_conv5.down encoder (models.py:731,749) -> flatten_vae_nl bottleneck
(models.py:559-561) -> mid_conv (models.py:750/1096) -> ResBlock2D x n_res (models.py:1097) -> UpBlock2D stack
(models.py:1098, 462) -> 7x7 out_conv + sigmoid (models.py:1099,1110).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import functional as Fn
from . import ops
from .modules import (Conv2d, DownBlock2D, ResBlock2D, SameBlock2D, UpBlock2D, as_nchw, as_nhwc, chain_res_blocks, to_float_nchw)
from .ops import MODE_NONE, MODE_UP, OUT_NCHW_F32, OUT_NHWC_BF16


class flatten_vae_nl(nn.Module):
    """Parameter-free VAE bottleneck (reference models.py:525-570): mu / logstd are the first / last half of the
    channels, flattened; z = mu + exp(logstd) * randn.  The network predicts log(sigma), not log(sigma^2).

    ``forward(x, train_vae)`` -> ``(mu, logstd, x_hat)`` when train_vae else ``(None, None, x_hat)`` with x_hat == mu
    exactly.  The reference hard-codes ``view(b, 16, 4, 4)`` (models.py:564); here x_hat keeps x's spatial size.
    ``eps`` may be injected (tests, reproducibility); otherwise it is drawn with torch.randn as in the reference.
    """

    def forward(self, x, train_vae, eps: Optional[torch.Tensor] = None):
        x = to_float_nchw(x)
        b, c2, h, w = x.shape
        zc = c2 // 2
        flat = x.reshape(b, -1)
        d = zc * h * w
        mu, logstd = flat[:, :d], flat[:, d:]
        if not train_vae:
            return None, None, mu.reshape(b, zc, h, w)
        if eps is None:
            eps = torch.randn(b, d, device=x.device)
        z = Fn.Reparam.apply(mu, logstd, eps.reshape(b, d).contiguous().float())
        return mu, logstd, z.view(b, zc, h, w)


class flatten_vae(nn.Module):
    """reference models.py:484-522: LinearELR encoder -> (mu_fc * 0.1, logstd_fc * 0.01) -> z = mu + exp(logstd) * randn ->
    view back to the input shape (no decoder).  ``forward(x, train_vae)`` -> ``(mu, logstd, x_hat)`` when train_vae, else
    ``(None, None, x_hat)`` with x_hat == mu (the reference multiplies logstd and the noise by 0).  Same constructor,
    sub-module names and state_dict keys; the sampling runs in the fused re-parameterisation kernel; ``eps`` may be injected."""

    def __init__(self, down_seq=[16 * 4 * 4, 256], up_seq=[256], vae_seq=[256, 256], use_weight_norm=False, lin=None) -> None:
        super().__init__()
        from .modules import LinearELR
        lin = LinearELR if lin is None else lin
        self.encoder = nn.Sequential(*[lin(down_seq[i], down_seq[i + 1], norm="demod", act=nn.LeakyReLU(0.2)) for i in range(len(down_seq) - 1)])
        self.mu_fc = lin(vae_seq[0], vae_seq[1])
        self.logstd_fc = lin(vae_seq[0], vae_seq[1])

    def forward(self, x, train_vae, eps: Optional[torch.Tensor] = None):
        x = to_float_nchw(x) if x.dim() == 4 else x
        shape = x.shape
        x_fl = self.encoder(x.flatten(start_dim=1))
        mu = self.mu_fc(x_fl) * 0.1
        if not train_vae:
            return None, None, mu.view(shape)
        logstd = self.logstd_fc(x_fl) * 0.01
        if eps is None:
            eps = torch.randn(*logstd.size(), device=logstd.device)
        z = Fn.Reparam.apply(mu.contiguous(), logstd.contiguous(), eps.reshape(mu.shape).contiguous().float())
        return mu, logstd, z.view(shape)


class local_vae(nn.Module):
    """reference models.py:442-482: DownBlock2D encoder -> flatten -> ``map_fc1`` (LinearELR, demodulated, LeakyReLU) ->
    ``map_fc2`` -> ``view(b, up_seq[0], 4, 4)`` -> UpBlock2D decoder.  The sampling lines are commented out in the reference, so the
    module is a deterministic bottleneck: ``forward(x)`` -> ``(None, None, x_hat)``.  Same constructor, sub-module names and
    state_dict keys (``map_fc1`` is hard-wired to 128 * 4 * 4 inputs as in the reference, models.py:464).  The blocks run on the
    sm_100a kernels (pool fused into the norm pass, up-sampling folded into the convolution); the two fully connected layers are
    library GEMMs on the flattened NCHW-ordered features."""

    def __init__(self, down_seq=[128, 128], up_seq=[128, 128], vae_seq=[512, 256], use_weight_norm=False, lin=None) -> None:
        super().__init__()
        from .modules import LinearELR
        lin = LinearELR if lin is None else lin
        self.up_seq = up_seq
        self.encoder = nn.Sequential(*[DownBlock2D(down_seq[i], down_seq[i + 1], use_weight_norm) for i in range(len(down_seq) - 1)])
        self.decoder = nn.Sequential(*[UpBlock2D(up_seq[i], up_seq[i + 1], use_weight_norm) for i in range(len(up_seq) - 1)])
        self.map_fc1 = lin(128 * 4 * 4, vae_seq[0], norm="demod", act=nn.LeakyReLU(0.2))
        self.map_fc2 = lin(vae_seq[0], 128 * 4 * 4, norm="demod", act=nn.LeakyReLU(0.2))

    def forward(self, x):
        b = x.shape[0]
        t = as_nhwc(x)
        last = len(self.encoder) - 1
        for i, blk in enumerate(self.encoder):
            t = blk.forward_nhwc(t, out_nchw_f32=(i == last))        # the last block writes fp32 NCHW: flatten() is then a view
        if len(self.encoder) == 0:
            t = to_float_nchw(x)
        x_fl = self.map_fc1(t.flatten(start_dim=1))
        x_de = self.map_fc2(x_fl).view(b, self.up_seq[0], 4, 4)
        t = Fn.ToNHWC.apply(x_de.contiguous())
        for blk in self.decoder:
            t = blk.forward_nhwc(t)
        return None, None, as_nchw(t, self.decoder[-1].out_channels if len(self.decoder) else self.up_seq[0])


class flatten_vae6(nn.Module):
    """reference models.py:802-833: LinearELR encoder -> (mu_fc * 0.1, logstd_fc * 0.01) -> z = mu + exp(logstd) * randn (training)
    / z = mu (eval) -> LinearELR decoder -> view back to the input shape.  ``forward(x)`` -> ``(mu, logstd, x_hat)``.  The
    sampling runs in the fused re-parameterisation kernel; ``eps`` may be injected."""

    def __init__(self, down_seq=[16 * 4 * 4, 256], up_seq=[256, 16 * 4 * 4], vae_seq=[256, 256], training=True, lin=None) -> None:
        super().__init__()
        from .modules import LinearELR
        lin = LinearELR if lin is None else lin
        self.encoder = nn.Sequential(*[lin(down_seq[i], down_seq[i + 1], norm="demod", act=nn.LeakyReLU(0.2)) for i in range(len(down_seq) - 1)])
        self.decoder = nn.Sequential(*[lin(up_seq[i], up_seq[i + 1], norm="demod", act=nn.LeakyReLU(0.2)) for i in range(len(up_seq) - 1)])
        self.mu_fc = lin(vae_seq[0], vae_seq[1])
        self.logstd_fc = lin(vae_seq[0], vae_seq[1])
        self.training = training

    def forward(self, x, eps: Optional[torch.Tensor] = None):
        x = to_float_nchw(x) if x.dim() == 4 else x
        shape = x.shape
        x_en = self.encoder(x.flatten(start_dim=1))
        mu = self.mu_fc(x_en) * 0.1
        logstd = self.logstd_fc(x_en) * 0.01
        if self.training:
            if eps is None:
                eps = torch.randn(*logstd.size(), device=logstd.device)
            z = Fn.Reparam.apply(mu.contiguous(), logstd.contiguous(), eps.reshape(mu.shape).contiguous().float())
        else:
            z = mu
        x_hat = self.decoder(z).view(shape)
        return mu, logstd, x_hat


class EFE_conv5(nn.Module):
    """The 2-D stage of the reference's default expression-feature extractor (models.py:724-798, the trainer's EFE): input
    pre-scale ``F.interpolate(bilinear, scale_factor, recompute_scale_factor=True)`` (models.py:764) -> ``down`` (SameBlock2D +
    DownBlock2D stack, :749) -> ``vae`` (flatten_vae_nl, :775-778) -> ``mid_conv`` (:785) -> ``view(N, C, D, H, W)`` (:786-787),
    the hand-off to the 3-D keypoint decoder.  Same constructor arguments and the same sub-module names / state_dict keys
    for ``down``, ``mid_conv`` and ``vae``; the 3-D part (``up``, ``out_conv``, ``mix``, ``mix_out``: Conv3d / heat-maps) is
    outside the hot path, so ``forward`` is not provided -- ``forward_2d`` returns what the 3-D part consumes."""

    def __init__(self, use_weight_norm=False, down_seq=[3, 32, 64, 128, 256, 32], up_seq=[256, 256, 128, 64, 32, 32], D=16, K=15, n_res=3,
                 scale_factor=0.25, use_vae=True) -> None:
        super().__init__()
        self.down = nn.Sequential(*[SameBlock2D(down_seq[i], down_seq[i + 1], use_weight_norm) if i == 0 else
                                    DownBlock2D(down_seq[i], down_seq[i + 1], use_weight_norm) for i in range(len(down_seq) - 1)])
        self.mid_conv = Conv2d(down_seq[-1] // 2, up_seq[0] * D, 1, 1, 0)
        self.C, self.D = up_seq[0], D
        self.scale_factor = scale_factor
        self.vae = flatten_vae_nl() if use_vae else None

    def forward_2d(self, x, x_a=None, train_vae=None, eps: Optional[torch.Tensor] = None):
        """-> (x3d [N, C, D, h, w] bf16, x_c, x_a_c, (x_mu, x_logstd), (x_vae, x_hat)) -- models.py:764-787."""
        def encode(t):
            t = ops.bilinear_resize(t.contiguous().float(), self.scale_factor)
            last = len(self.down) - 1
            for i, blk in enumerate(self.down):
                if i == 0:
                    t = blk.forward_from_frames(t) if blk.pointwise_ok(t) else blk.forward_nhwc(as_nhwc(t))
                else:
                    t = blk.forward_nhwc(t, out_nchw_f32=(i == last))
            return t
        x = encode(x)
        x_z = x
        x_c, x_a_c = (x, encode(x_a)) if x_a is not None else (None, None)
        if self.vae is not None:
            x_vae = x
            x_mu, x_logstd, x_hat = self.vae(x_vae, train_vae, eps)
            x_z = x_hat
        else:
            x_mu = x_logstd = x_hat = x_vae = None
        t = Fn.ConvOnly.apply(Fn.ToNHWC.apply(x_z.contiguous()), self.mid_conv.weight, self.mid_conv.bias, 1, OUT_NHWC_BF16)[0]
        y = as_nchw(t, self.mid_conv.out_channels)                 # logical [N, C*D, h, w]
        n, _, h, w = y.shape
        return y.reshape(n, self.C, self.D, h, w), x_c, x_a_c, (x_mu, x_logstd), (x_vae, x_hat)

    def forward(self, x, x_a=None, kpc=None, train_vae=None):
        raise NotImplementedError("EFE_conv5: the 3-D keypoint decoder (UpBlock3D / Conv3d / heat-maps, models.py:788-798) is outside "
                                  "the hot path (SURVEY.md 2.1); use forward_2d for the 2-D stage")


class Generator(nn.Module):
    """The 2-D decoder of the reference's ``Generator`` (models.py:1085-1111): in_conv (CNA 3x3, LeakyReLU) -> mid_conv (1x1) ->
    occlusion gate -> ResBlock2D x n_res -> UpBlock2D stack -> out_conv (7x7) -> sigmoid, every block spectral-normalised
    (``use_weight_norm=True``).  Same constructor, sub-module names and state_dict keys.  The reference's ``forward(fs,
    deformation, occlusion)`` first warps the 3-D feature volume with ``F.grid_sample`` (keypoint / motion geometry, outside the
    hot path); ``forward_2d(fs2d, occlusion)`` takes the warped features already flattened to [N, C*D, H, W] (models.py:1102)."""

    def __init__(self, use_weight_norm=True, n_res=6, up_seq=[256, 128, 64], D=16, C=32):
        super().__init__()
        from .modules import ConvBlock2D
        self.in_conv = ConvBlock2D("CNA", C * D, up_seq[0], 3, 1, 1, use_weight_norm, nonlinearity_type="leakyrelu")
        self.mid_conv = Conv2d(up_seq[0], up_seq[0], 1, 1, 0)
        self.res = nn.Sequential(*[ResBlock2D(up_seq[0], use_weight_norm) for _ in range(n_res)])
        chain_res_blocks(self.res)
        self.up = nn.Sequential(*[UpBlock2D(up_seq[i], up_seq[i + 1], use_weight_norm) for i in range(len(up_seq) - 1)])
        self.out_conv = Conv2d(up_seq[-1], 3, 7, 1, 3)

    def forward_2d(self, fs2d, occlusion=None):
        t = self.in_conv.forward_nhwc(as_nhwc(fs2d))
        t = Fn.ConvOnly.apply(t, self.mid_conv.weight, self.mid_conv.bias, 1, OUT_NHWC_BF16)[0]
        if occlusion is not None:                                    # fs = fs * occlusion (models.py:1105): [N,1,H,W] gate
            t = t * occlusion.permute(0, 2, 3, 1).to(t.dtype)
        for blk in self.res:
            t = blk.forward_nhwc(t.contiguous())
        for blk in self.up:
            t = blk.forward_nhwc(t)
        logits = Fn.ConvOnly.apply(t, self.out_conv.weight, self.out_conv.bias, 7, OUT_NCHW_F32)[0]
        return torch.sigmoid(logits)

    def forward(self, fs, deformation, occlusion):
        raise NotImplementedError("Generator.forward warps a 3-D feature volume with F.grid_sample (models.py:1101-1102), which is outside "
                                  "the hot path; use forward_2d(fs2d, occlusion) on the warped, flattened features")


class Discriminator(nn.Module):
    """The reference's patch ``Discriminator`` (models.py:1114-1139): spectral-normalised CNA blocks with instance norm and
    LeakyReLU(0.2), 3x3 stride 2 / stride 1, and a final un-normalised 3x3 ``CN`` block.  Same constructor, ``layers`` ModuleList
    and state_dict keys.  ``forward(x, kp)`` needs the keypoint heat-map ``kp2gaussian_2d`` (reference utils.py, outside the hot
    path); ``forward_features(x_cat)`` takes the concatenated ``[frame | heat-map]`` tensor (models.py:1130-1131) and returns
    ``(output, features)`` as the reference does."""

    def __init__(self, use_weight_norm=True, down_seq=[64, 128, 256, 512], K=15):
        super().__init__()
        from .modules import ConvBlock2D
        layers = [ConvBlock2D("CNA", 3 + K, down_seq[0], 3, 2, 1, use_weight_norm, "instance", "leakyrelu")]
        layers.extend([ConvBlock2D("CNA", down_seq[i], down_seq[i + 1], 3, 2 if i < len(down_seq) - 2 else 1, 1, use_weight_norm, "instance",
                                   "leakyrelu") for i in range(len(down_seq) - 1)])
        layers.append(ConvBlock2D("CN", down_seq[-1], 1, 3, 1, 1, use_weight_norm, activation_type="none"))
        self.layers = nn.ModuleList(layers)

    def forward_features(self, x_cat):
        res = [x_cat]
        t = as_nhwc(x_cat)
        for layer in self.layers:
            t = layer.forward_nhwc(t)
            res.append(as_nchw(t, layer.out_channels))
        return res[-1], res[1:-1]

    def forward(self, x, kp):
        raise NotImplementedError("Discriminator.forward builds a keypoint heat-map with kp2gaussian_2d (reference utils.py:121-127), which is "
                                  "outside the hot path; use forward_features(torch.cat([x, heatmap], dim=1))")


class FaceVAE(nn.Module):
    """The anchor "face-vae" (SURVEY.md section 8).  Sub-module names give the oracle's state_dict keys."""

    def __init__(self, down_seq: Sequence[int] = (3, 32, 64, 128, 256, 32), up_seq: Sequence[int] = (256, 256, 128, 64, 32),
                 n_res: int = 2, use_weight_norm: bool = False):
        super().__init__()
        d, u = list(down_seq), list(up_seq)
        self.enc = nn.Sequential(*[SameBlock2D(d[i], d[i + 1], use_weight_norm) if i == 0 else
                                   DownBlock2D(d[i], d[i + 1], use_weight_norm) for i in range(len(d) - 1)])
        self.vae = flatten_vae_nl()
        self.zc = d[-1] // 2
        self.mid_conv = Conv2d(self.zc, u[0], 1, 1, 0)
        self.res = nn.Sequential(*[ResBlock2D(u[0], use_weight_norm) for _ in range(n_res)])
        chain_res_blocks(self.res)
        self.up = nn.Sequential(*[UpBlock2D(u[i], u[i + 1], use_weight_norm) for i in range(len(u) - 1)])
        self.out_conv = Conv2d(u[-1], 3, 7, 1, 3)
        self.out_conv.prep_kind = -1 if u[-1] == 32 else 0     # the tap-folded out_conv kernels prepare their own operands
        self.n_down = len(d) - 2

    def latent_dim(self, h: int, w: int) -> int:
        f = 2 ** self.n_down
        return self.zc * (h // f) * (w // f)

    # -- pieces ---------------------------------------------------------------------------------------------
    def encode_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """frames [N,3,H,W] fp32 -> h [N, 2*zc, H/f, W/f] fp32 NCHW (mu | logstd kept in fp32 for the KL term)."""
        last = len(self.enc) - 1
        t = x
        for i, blk in enumerate(self.enc):
            if i == 0:
                # raw RGB frames: per-pixel affine fast path (statistics from the input moments), else the generic block
                t = blk.forward_from_frames(x) if blk.pointwise_ok(x) else blk.forward_nhwc(as_nhwc(x))
            else:
                t = blk.forward_nhwc(t, out_nchw_f32=(i == last))
        return t

    def decode_nhwc(self, z_nchw: torch.Tensor) -> torch.Tensor:
        """z [N,zc,h,w] fp32 -> decoder features NHWC bf16 in front of out_conv."""
        t = Fn.ToNHWC.apply(z_nchw)
        # mid_conv's epilogue emits the batch sums of its output when a ResBlock2D (which normalises its input) follows
        emit = self.training and len(self.res) > 0
        t, sums = Fn.ConvOnly.apply(t, self.mid_conv.weight, self.mid_conv.bias, 1, OUT_NHWC_BF16, emit)
        ops.attach_stats(t, sums)
        for blk in self.res:
            t = blk.forward_nhwc(t)
        for blk in self.up:
            t = blk.forward_nhwc(t)        # up-sampling folded into the convolution: no 4x tensor between the blocks
        return t

    # -- public ---------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, train_vae: bool = True, eps: Optional[torch.Tensor] = None):
        """-> (mu, logstd, x_hat): mu/logstd [N,Dz] fp32 (None when not train_vae), x_hat [N,3,H,W] fp32 in (0,1)."""
        h = self.encode_nchw(x)
        mu, logstd, z = self.vae(h, train_vae, eps)
        d = self.decode_nhwc(z)
        logits = Fn.ConvOnly.apply(d, self.out_conv.weight, self.out_conv.bias, 7, OUT_NCHW_F32)[0]
        return mu, logstd, torch.sigmoid(logits)

    def forward_loss(self, x: torch.Tensor, eps: Optional[torch.Tensor] = None, l1: bool = False):
        """Fused training path -> dict(K, R, x_hat, mu, logstd): un-weighted KL and reconstruction terms with the
        re-parameterisation+KL and out_conv+sigmoid+loss(+gradient) stages each running as one fused kernel."""
        h = self.encode_nchw(x)
        b, c2, hh, ww = h.shape
        dz = self.zc * hh * ww
        if eps is None:
            eps = torch.randn(b, dz, device=x.device)
        flat = h.view(b, 2 * dz)
        z, kl = Fn.ReparamKL.apply(flat, eps.reshape(b, dz).contiguous().float())
        d = self.decode_nhwc(z.view(b, self.zc, hh, ww))
        x_hat, rec = Fn.ConvSigmoidRecon.apply(d, self.out_conv.weight, self.out_conv.bias, x, l1)
        return {"K": kl, "R": rec, "x_hat": x_hat, "mu": flat[:, :dz], "logstd": flat[:, dz:]}


def face_vae_256(**kw) -> FaceVAE:
    """BASELINE.json configs[1..2]: 256x256 anchor."""
    return FaceVAE(**kw)


def face_vae_512(**kw) -> FaceVAE:
    """BASELINE.json configs[3]: deeper encoder/decoder, larger latent."""
    return FaceVAE(down_seq=(3, 32, 64, 128, 256, 512, 64), up_seq=(512, 512, 256, 128, 64, 32), **kw)
