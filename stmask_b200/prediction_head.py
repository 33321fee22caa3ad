"""FCA / FCB prediction head over the FPN levels (reference layers/modules/prediction_head_FC.py:13-247,
shared over P3..P7 by STMask.py:86-100 with cfg.share_prediction_module).

Same constructor order, parameter names and shapes as the reference's `PredictionModule_FC` for the STMask configs
(`upfeature.0`, `{conf,bbox,mask,track}_extra.{0,2}`, `centerness_layer.k`, `bbox_layer.k`, `conf_layer.k[.conv_offset|
.conv_adaption|.conv]`, `track_layer.k`, `mask_layer.k`), so a released state_dict loads unchanged and a module built
under the same seed has the same weights.

B200 path (`forward_levels`): every convolution of the head runs ONCE for all five levels as a grouped launch of the
tcgen05 implicit-GEMM main loop (plain-conv mode of the deformable-conv kernel, NHWC bf16, bias + ReLU fused) — the
reference runs each of them once per level:

    upfeature                                   1 launch   (256 -> 256, ReLU)
    first conv of the four *_extra stacks       1 launch   (same input: fused into ONE 256 -> 1024 conv)
    second conv of the four *_extra stacks      4 launches
    per anchor kernel k (3x3, 3x5, 5x3):
        bbox_layer.k + centerness_layer.k       1 launch   (same input: fused 256 -> 4+1, fp32 output — box deltas steer the
                                                            FCB sampling positions and must not be rounded to bf16)
        FCB: conv_adaption with offsets derived inside the kernel from those box deltas + ReLU      1 launch
             conf_layer.k.conv (256 -> 41)      1 launch   (FCA only: conf_layer.k itself)
        track_layer.k, mask_layer.k             2 launches

21 launches per batch of frames for the whole head instead of ~100 per frame.  `make_priors` (pure Python, 15 345
iterations per frame in the reference, prediction_head_FC.py:224-247; its `prior_cache`, STMask.py:16, is never used)
is computed vectorised once per level size and cached.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .feature_align import FeatureAlign

HEAD_KERNELS = ((3, 3), (3, 5), (5, 3))          # cfg.head_layer_params (config.py:657-659)
_PRIOR_CACHE: Dict[Tuple, torch.Tensor] = {}


def make_priors(conv_h: int, conv_w: int, aspect_ratios: Sequence[Sequence[float]] = HEAD_KERNELS,
                scales: Sequence[float] = (1.0,), device="cpu") -> torch.Tensor:
    """[1, conv_h*conv_w*len(aspect_ratios)*len(scales), 4] priors (cx, cy, w, h), identical to the reference's loop
    (prediction_head_FC.py:224-247), vectorised and cached per (size, anchors, device)."""
    key = (conv_h, conv_w, tuple(tuple(a) for a in aspect_ratios), tuple(scales), str(device))
    hit = _PRIOR_CACHE.get(key)
    if hit is None:
        ys, xs = torch.meshgrid(torch.arange(conv_h, dtype=torch.float64), torch.arange(conv_w, dtype=torch.float64), indexing="ij")
        cx, cy = ((xs + 0.5) / conv_w).reshape(-1, 1), ((ys + 0.5) / conv_h).reshape(-1, 1)
        wh = [(s / scales[0] * arw / conv_w, s / scales[0] * arh / conv_h) for (arh, arw) in aspect_ratios for s in scales]
        w = torch.tensor([a for a, _ in wh], dtype=torch.float64).view(1, -1)
        h = torch.tensor([b for _, b in wh], dtype=torch.float64).view(1, -1)
        pri = torch.stack([cx.expand(-1, w.shape[1]), cy.expand(-1, w.shape[1]), w.expand(cx.shape[0], -1), h.expand(cx.shape[0], -1)], -1)
        hit = pri.reshape(1, -1, 4).float().to(device)
        _PRIOR_CACHE[key] = hit
    return hit


def _extra(n_layers: int, ch: int):
    if n_layers == 0:
        return nn.Identity()
    return nn.Sequential(*sum([[nn.Conv2d(ch, ch, kernel_size=3, padding=1), nn.ReLU(inplace=True)] for _ in range(n_layers)], []))


class PredictionHeadFC(nn.Module):
    def __init__(self, in_channels: int = 256, num_classes: int = 41, mask_dim: int = 32, embed_dim: int = 128,
                 fcb: Optional[str] = "ada", extra_layers: Tuple[int, int, int, int] = (2, 2, 2, 2),
                 head_kernels: Sequence[Tuple[int, int]] = HEAD_KERNELS, deform_groups: int = 1):
        super().__init__()
        c = in_channels
        self.out_channels, self.num_classes, self.mask_dim, self.embed_dim = c, num_classes, mask_dim, embed_dim
        self.head_kernels = tuple(tuple(k) for k in head_kernels)
        self.fcb = fcb
        # construction order == the reference's (prediction_head_FC.py:57-127): same seed -> same parameters
        self.upfeature = nn.Sequential(nn.Conv2d(c, c, 3, padding=1), nn.ReLU(inplace=True))     # cfg.extra_head_net
        self.bbox_layer, self.track_layer, self.mask_layer = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.centerness_layer, self.conf_layer = nn.ModuleList(), nn.ModuleList()
        for (kh, kw) in self.head_kernels:
            pad = ((kh - 1) // 2, (kw - 1) // 2)
            self.centerness_layer.append(nn.Conv2d(c, 1, (kh, kw), padding=pad))
            self.bbox_layer.append(nn.Conv2d(c, 4, (kh, kw), padding=pad))
            if fcb:
                self.conf_layer.append(FeatureAlign(c, num_classes, kernel_size=(kh, kw), deformable_groups=deform_groups,
                                                    use_pred_offset=(fcb == "ada")))
            else:
                self.conf_layer.append(nn.Conv2d(c, num_classes, (kh, kw), padding=pad))
            self.track_layer.append(nn.Conv2d(c, embed_dim, (kh, kw), padding=pad))
            self.mask_layer.append(nn.Conv2d(c, mask_dim, (kh, kw), padding=pad))
        self.track_extra = _extra(extra_layers[2], c)
        self.conf_extra = _extra(extra_layers[0], c)
        self.bbox_extra, self.mask_extra = _extra(extra_layers[0], c), _extra(extra_layers[1], c)
        self._pc: Dict[str, ops.PlainConv] = {}
        self._fused_key = None
        self._fused = None

    # ---------------------------------------------------------------- grouped plain convs
    def _conv(self, name: str, xs, weight, bias, padding, relu=False, out_f32=False):
        pc = self._pc.get(name)
        if pc is None:
            pc = self._pc[name] = ops.PlainConv()
        return pc(xs, weight, bias, 1, padding, relu=relu, out_f32=out_f32)

    def _fused_params(self):
        """First convs of the four *_extra stacks as ONE 256 -> 1024 conv; bbox_layer.k + centerness_layer.k as one 256 -> 5."""
        srcs = [self.conf_extra[0], self.bbox_extra[0], self.mask_extra[0], self.track_extra[0]]
        ps = [p for m in srcs for p in (m.weight, m.bias)] + [p for k in range(len(self.head_kernels))
                                                              for m in (self.bbox_layer[k], self.centerness_layer[k]) for p in (m.weight, m.bias)]
        key = tuple((id(p), p._version, p.data_ptr(), p.dtype) for p in ps)
        if key != self._fused_key:
            ex_w = torch.cat([m.weight.detach() for m in srcs], 0)
            ex_b = torch.cat([m.bias.detach() for m in srcs], 0)
            bc = [(torch.cat([self.bbox_layer[k].weight.detach(), self.centerness_layer[k].weight.detach()], 0),
                   torch.cat([self.bbox_layer[k].bias.detach(), self.centerness_layer[k].bias.detach()], 0))
                  for k in range(len(self.head_kernels))]
            self._fused = (ex_w, ex_b, bc)
            self._fused_key = key
        return self._fused

    @torch.no_grad()
    def forward_levels(self, xs: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """xs: the FPN levels [B, 256, H_l, W_l] (channels-last; bf16 -> tcgen05, fp32 -> CUDA cores).  Returns what
        `STMask.forward_single` collects over the levels (STMask.py:245-279): loc [B, P, 4], centerness [B, P, 1] (tanh),
        conf [B, P, num_classes] (logits), mask_coeff [B, P, mask_dim], track [B, P, embed_dim] (L2-normalised), priors
        [1, P, 4] and T2S_feat (list per level), P = 3 * sum(H_l * W_l)."""
        if len(self.conf_extra) != 4 or len(self.bbox_extra) != 4:
            raise NotImplementedError("forward_levels implements the STMask head layout: extra_layers = (2, 2, 2, 2)")
        xs = list(xs)
        c = self.out_channels
        x = self._conv("upfeature", xs, self.upfeature[0].weight, self.upfeature[0].bias, 1, relu=True)
        ex_w, ex_b, bc = self._fused_params()
        h1 = self._conv("extra1", x, ex_w, ex_b, 1, relu=True)                       # [B, 4*256, H, W]: conf | bbox | mask | track
        stacks = {}
        for i, (name, seq) in enumerate((("conf", self.conf_extra), ("bbox", self.bbox_extra), ("mask", self.mask_extra),
                                         ("track", self.track_extra))):
            stacks[name] = self._conv(name + "_extra2", [t[:, i * c:(i + 1) * c] for t in h1], seq[2].weight, seq[2].bias, 1, relu=True)
        n_lvl = len(xs)
        per_level: List[Dict[str, List[torch.Tensor]]] = [dict(loc=[], ctr=[], conf=[], mask=[], track=[]) for _ in range(n_lvl)]
        for k, (kh, kw) in enumerate(self.head_kernels):
            pad = ((kh - 1) // 2, (kw - 1) // 2)
            bw, bb = bc[k]
            bbox_ctr = self._conv(f"bbox_ctr{k}", stacks["bbox"], bw, bb, pad, out_f32=True)         # fp32 [B, 16, H, W]: 4 deltas | centerness
            deltas = [t[:, :4] for t in bbox_ctr]
            cl = self.conf_layer[k]
            if isinstance(cl, FeatureAlign):
                cal = cl.calibrate_levels(stacks["conf"], deltas)                                # offsets from the deltas inside the kernel, ReLU
                conf = self._conv(f"conf{k}", cal, cl.conv.weight, cl.conv.bias, pad)
            else:
                conf = self._conv(f"conf{k}", stacks["conf"], cl.weight, cl.bias, pad)
            track = self._conv(f"track{k}", stacks["track"], self.track_layer[k].weight, self.track_layer[k].bias, pad)
            mask = self._conv(f"mask{k}", stacks["mask"], self.mask_layer[k].weight, self.mask_layer[k].bias, pad)
            for l in range(n_lvl):
                nhwc = lambda t, n: t.permute(0, 2, 3, 1)[..., :n]
                per_level[l]["loc"].append(nhwc(bbox_ctr[l], 4).float())
                per_level[l]["ctr"].append(bbox_ctr[l].permute(0, 2, 3, 1)[..., 4:5].float())
                per_level[l]["conf"].append(nhwc(conf[l], self.num_classes).float())
                per_level[l]["track"].append(nhwc(track[l], self.embed_dim).float())
                per_level[l]["mask"].append(nhwc(mask[l], self.mask_dim).float())
        out = {"loc": [], "centerness": [], "conf": [], "mask_coeff": [], "track": [], "priors": []}
        for l, d in enumerate(per_level):
            b = xs[l].shape[0]
            out["loc"].append(torch.cat(d["loc"], -1).reshape(b, -1, 4))
            # the reference concatenates the per-kernel centerness maps along dim 1 of [B, H, W, 1] (prediction_head_FC.py:190):
            # kernel-major order inside a level, unlike loc / conf / mask / track (pixel-major, kernels innermost)
            out["centerness"].append(torch.tanh(torch.cat(d["ctr"], 1).reshape(b, -1, 1)))
            out["conf"].append(torch.cat(d["conf"], -1).reshape(b, -1, self.num_classes))
            out["mask_coeff"].append(torch.cat(d["mask"], -1).reshape(b, -1, self.mask_dim))
            out["track"].append(F.normalize(torch.cat(d["track"], -1).reshape(b, -1, self.embed_dim), dim=-1))
            out["priors"].append(make_priors(xs[l].shape[2], xs[l].shape[3], self.head_kernels, (1.0,), xs[l].device))
        res = {k: torch.cat(v, 1) for k, v in out.items()}
        res["T2S_feat"] = x
        return res
