"""Drop-in for the reference's networks/fusion_module.py (FusionModule, fusion_module.py:52-130): multi-frame feature
fusion.  Source-frame features are warped to the target frame with the VFI flows, every feature gets a Fourier
embedding of the (down-scaled) flow appended, the two warped sources are blended with the VFI merge mask, and a 1x1
ConvBlock per level brings [target | blended] back to the encoder's width.  state_dict keys: fusion_conv.{i}.conv.conv.*"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import warp_ops
from ..layers import ConvBlock1x1
from .IFRNet import warp


def fourier_embed(x, num_freqs):
    """[x, sin(2^k x), cos(2^k x) for k < num_freqs] along channels (fusion_module.py:8-48: log-sampled bands)"""
    out = [x]
    for k in range(num_freqs):
        f = 2.0 ** k
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, 1)


class FusionModule(nn.Module):
    def __init__(self, args, num_ch_enc, embed_multires=10):
        super().__init__()
        self.embed_multires = embed_multires
        self.embed_dim = 2 * (1 + 2 * embed_multires)  # flow has two channels
        self.num_ch_enc = num_ch_enc
        self.backbone = args.backbone
        self.convs = OrderedDict()
        for i in range(len(num_ch_enc) - 1, -1, -1):
            self.convs[("conv1x1", i)] = ConvBlock1x1(2 * (int(num_ch_enc[i]) + self.embed_dim), int(num_ch_enc[i]))
        self.fusion_conv = nn.ModuleList(list(self.convs.values()))

    def get_embedding_flow(self, x):
        """embedding of the flow at every feature resolution: each halving of the resolution halves the displacement
        (Lite-Mono's first level sits at 1/4, so it is halved twice)"""
        outs = []
        for i in range(len(self.num_ch_enc)):
            for _ in range(2 if (i == 0 and self.backbone == "LiteMono") else 1):
                x = warp_ops.resize_bilinear(x, scale_factor=0.5, mul=(0.5, 0.5))
            outs.append(fourier_embed(x, self.embed_multires))
        return outs

    def warp_features(self, features, flow):
        fh, fw = flow.shape[-2:]
        outs = []
        for feat in features:
            H, W = feat.shape[-2:]
            fl = warp_ops.resize_bilinear(flow, size=(H, W), mul=(W / fw, H / fh))
            outs.append(warp(feat, fl))
        return outs

    def merge_features(self, features_warped, merge_mask):
        feats_n1, feats_0, feats_p1 = features_warped
        outs = []
        for f_n1, f_0, f_p1 in zip(feats_n1, feats_0, feats_p1):
            m = warp_ops.resize_bilinear(merge_mask, size=f_0.shape[-2:])
            outs.append(torch.cat([f_0, m * f_n1 + (1 - m) * f_p1], 1))
        return outs

    def forward(self, features, flows, merge_mask):
        feats_n1, feats_0, feats_p1 = features
        flow_0_n1, flow_0_p1 = flows
        warped_n1 = self.warp_features(feats_n1, flow_0_n1)
        warped_p1 = self.warp_features(feats_p1, flow_0_p1)
        emb_0 = self.get_embedding_flow(torch.zeros_like(flow_0_n1))
        emb_n1 = self.get_embedding_flow(flow_0_n1)
        emb_p1 = self.get_embedding_flow(flow_0_p1)
        cat = lambda fs, es: [torch.cat([f, e], 1) for f, e in zip(fs, es)]
        merged = self.merge_features([cat(warped_n1, emb_n1), cat(feats_0, emb_0), cat(warped_p1, emb_p1)], merge_mask)
        return [self.convs[("conv1x1", i)](m) for i, m in enumerate(merged)]
