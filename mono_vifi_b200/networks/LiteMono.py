"""Drop-in for the reference's networks/LiteMono.py (LiteMono.py:296-505): Lite-Mono depth encoder (consecutive dilated
depth-wise convolutions + one cross-covariance-attention block per stage) and its three-level decoder.  Same class
names, constructor signatures, attributes (`num_ch_enc`) and state_dict keys (`downsample_layers.*`, `stem2.*`,
`stages.{i}.{j}.*`, `decoder.{i}.*`).  `timm` is not required: DropPath and trunc_normal_ are restated here."""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import litemono_ops
from ..bn_act import bn_act
from ..conv import Conv2d, conv2d
from ..layers import ConvBlock, Conv3x3, upsample


class DropPath(nn.Module):
    """stochastic depth per sample (timm.models.layers.DropPath)"""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class PositionalEncodingFourier(nn.Module):
    """DETR-style sine/cosine position code projected to `dim` channels by a 1x1 conv (LiteMono.py:13-47)"""

    def __init__(self, hidden_dim=32, dim=768, temperature=10000):
        super().__init__()
        self.token_projection = nn.Conv2d(hidden_dim * 2, dim, kernel_size=1)
        self.scale = 2 * math.pi
        self.temperature = temperature
        self.hidden_dim = hidden_dim
        self.dim = dim

    def forward(self, B, H, W):
        dev = self.token_projection.weight.device
        eps = 1e-6
        y = torch.arange(1, H + 1, dtype=torch.float32, device=dev).view(1, H, 1).expand(B, H, W)
        x = torch.arange(1, W + 1, dtype=torch.float32, device=dev).view(1, 1, W).expand(B, H, W)
        y = y / (y[:, -1:, :] + eps) * self.scale
        x = x / (x[:, :, -1:] + eps) * self.scale
        k = torch.arange(self.hidden_dim, dtype=torch.float32, device=dev)
        dim_t = self.temperature ** (2 * torch.div(k, 2, rounding_mode="floor") / self.hidden_dim)

        def code(p):
            p = p[:, :, :, None] / dim_t
            return torch.stack((p[:, :, :, 0::2].sin(), p[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos = torch.cat((code(y), code(x)), dim=3).permute(0, 3, 1, 2)
        return self.token_projection(pos)


class XCA(nn.Module):
    """cross-covariance attention: attention over channels (d_h x d_h per head) instead of tokens (LiteMono.py:50-90)"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 4, 1)  # [B, h, d, N]
        q = F.normalize(q, dim=-1)
        k = F.normalize(k, dim=-1)
        attn = self.attn_drop(((q @ k.transpose(-2, -1)) * self.temperature).softmax(dim=-1))
        x = (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            return litemono_ops.layer_norm_cl(x, self.weight, self.bias, self.eps)   # mvf_layernorm_cl_* on CUDA
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * ((x - u) / torch.sqrt(s + self.eps)) + self.bias[:, None, None]


class GELU(nn.GELU):
    """nn.GELU() (erf form); mvf_gelu_* on CUDA"""

    def forward(self, x):
        return litemono_ops.gelu(x)


class BNGELU(nn.Module):
    def __init__(self, nIn):
        super().__init__()
        self.bn = nn.BatchNorm2d(nIn, eps=1e-5)
        self.act = GELU()

    def forward(self, x):
        return self.act(bn_act(self.bn, x, relu=False))


class Conv(nn.Module):
    def __init__(self, nIn, nOut, kSize, stride, padding=0, dilation=(1, 1), groups=1, bn_act=False, bias=False):
        super().__init__()
        self.bn_act = bn_act
        self.conv = Conv2d(nIn, nOut, kernel_size=kSize, stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias)
        if bn_act:
            self.bn_gelu = BNGELU(nOut)

    def forward(self, x):
        y = self.conv(x)
        return self.bn_gelu(y) if self.bn_act else y


class CDilated(nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1, groups=1, bias=False):
        super().__init__()
        self.conv = Conv2d(nIn, nOut, kSize, stride=stride, padding=int((kSize - 1) / 2) * d, bias=bias, dilation=d, groups=groups)

    def forward(self, x):
        c = self.conv
        if litemono_ops.dwconv3x3_usable(x, c.weight, c.stride, c.padding, c.dilation, c.groups):
            return litemono_ops.dwconv3x3(x, c.weight, c.bias, c.dilation)   # mvf_dwconv3x3_*
        return c(x)


class _Mlp(nn.Module):
    """mixin: pwconv1 -> GELU -> pwconv2 -> layer scale, applied on channels-last tokens"""

    def _mlp(self, x):
        x = self.pwconv2(self.act(self.pwconv1(x)))
        return x if self.gamma is None else self.gamma * x

    def _mlp_cl(self, x):
        """the same on an NCHW-shaped channels-last tensor: the two nn.Linear layers are 1x1 convolutions on the tensor-core path"""
        h = self.act(conv2d(x, self.pwconv1.weight[:, :, None, None], self.pwconv1.bias))
        y = conv2d(h, self.pwconv2.weight[:, :, None, None], self.pwconv2.bias)
        return y if self.gamma is None else y * self.gamma.view(1, -1, 1, 1)


class DilatedConv(_Mlp):
    """depth-wise dilated 3x3 -> BN -> inverted-bottleneck MLP, residual (LiteMono.py:157-201; its `norm` is unused there too)"""

    def __init__(self, dim, k, dilation=1, stride=1, drop_path=0., layer_scale_init_value=1e-6, expan_ratio=6):
        super().__init__()
        self.ddwconv = CDilated(dim, dim, kSize=k, stride=stride, groups=dim, d=dilation)
        self.bn1 = nn.BatchNorm2d(dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, expan_ratio * dim)
        self.act = GELU()
        self.pwconv2 = nn.Linear(expan_ratio * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x):
        if x.is_cuda:
            return x + self.drop_path(self._mlp_cl(bn_act(self.bn1, self.ddwconv(x), relu=False)))
        y = self.bn1(self.ddwconv(x)).permute(0, 2, 3, 1)
        return x + self.drop_path(self._mlp(y).permute(0, 3, 1, 2))


class LGFI(_Mlp):
    """local-global feature interaction: (position code) + XCA, then the MLP, both residual (LiteMono.py:204-256)"""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6, expan_ratio=6, use_pos_emb=True, num_heads=6,
                 qkv_bias=True, attn_drop=0., drop=0.):
        super().__init__()
        self.dim = dim
        self.pos_embd = PositionalEncodingFourier(dim=dim) if use_pos_emb else None
        self.norm_xca = LayerNorm(dim, eps=1e-6)
        self.gamma_xca = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.xca = XCA(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, expan_ratio * dim)
        self.act = GELU()
        self.pwconv2 = nn.Linear(expan_ratio * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        t = x.reshape(B, C, H * W).permute(0, 2, 1)
        if self.pos_embd is not None:
            t = t + self.pos_embd(B, H, W).reshape(B, -1, t.shape[1]).permute(0, 2, 1)
        t = t + self.gamma_xca * self.xca(self.norm_xca(t))
        if x.is_cuda:
            y = self._mlp_cl(self.norm(t.reshape(B, H, W, C)).permute(0, 3, 1, 2))
        else:
            y = self._mlp(self.norm(t.reshape(B, H, W, C))).permute(0, 3, 1, 2)
        return x + self.drop_path(y)


class AvgPool(nn.Module):
    def __init__(self, ratio):
        super().__init__()
        self.pool = nn.ModuleList([nn.AvgPool2d(3, stride=2, padding=1) for _ in range(ratio)])

    def forward(self, x):
        for p in self.pool:
            x = p(x)
        return x


# widths, blocks per stage and dilation schedules (192x640/512 ; 320x1024) of the published variants (LiteMono.py:307-341)
_VARIANTS = {
    "lite-mono": ([48, 80, 128], [4, 4, 10], ([1, 2, 3], [1, 2, 3], [1, 2, 3, 1, 2, 3, 2, 4, 6]), ([1, 2, 5], [1, 2, 5], [1, 2, 5, 1, 2, 5, 2, 4, 10])),
    "lite-mono-small": ([48, 80, 128], [4, 4, 7], ([1, 2, 3], [1, 2, 3], [1, 2, 3, 2, 4, 6]), ([1, 2, 5], [1, 2, 5], [1, 2, 5, 2, 4, 10])),
    "lite-mono-tiny": ([32, 64, 128], [4, 4, 7], ([1, 2, 3], [1, 2, 3], [1, 2, 3, 2, 4, 6]), ([1, 2, 5], [1, 2, 5], [1, 2, 5, 2, 4, 10])),
    "lite-mono-8m": ([64, 128, 224], [4, 4, 10], ([1, 2, 3], [1, 2, 3], [1, 2, 3, 1, 2, 3, 2, 4, 6]), ([1, 2, 3], [1, 2, 3], [1, 2, 3, 1, 2, 3, 2, 4, 6])),
}


class DepthEncoder(nn.Module):
    def __init__(self, in_chans=3, model='lite-mono', height=192, width=640, global_block=[1, 1, 1],
                 global_block_type=['LGFI', 'LGFI', 'LGFI'], drop_path_rate=0.2, layer_scale_init_value=1e-6, expan_ratio=6,
                 heads=[8, 8, 8], use_pos_embd_xca=[True, False, False], **kwargs):
        super().__init__()
        dims, depth, dil_lo, dil_hi = _VARIANTS[model]
        self.num_ch_enc = np.array(dims)
        self.depth, self.dims = list(depth), list(dims)
        if height == 192 and width in (640, 512):
            self.dilation = [list(d) for d in dil_lo]
        elif height == 320 and width == 1024:
            self.dilation = [list(d) for d in dil_hi]
        for g in global_block_type:
            assert g in ['None', 'LGFI']
        d0 = dims[0]
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(Conv(in_chans, d0, kSize=3, stride=2, padding=1, bn_act=True),
                                                    Conv(d0, d0, kSize=3, stride=1, padding=1, bn_act=True),
                                                    Conv(d0, d0, kSize=3, stride=1, padding=1, bn_act=True)))
        self.stem2 = nn.Sequential(Conv(d0 + 3, d0, kSize=3, stride=2, padding=1, bn_act=False))
        self.input_downsample = nn.ModuleList([AvgPool(i) for i in range(1, 5)])
        for i in range(2):
            self.downsample_layers.append(nn.Sequential(Conv(dims[i] * 2 + 3, dims[i + 1], kSize=3, stride=2, padding=1, bn_act=False)))
        self.stages = nn.ModuleList()
        dp = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depth))]
        cur = 0
        for i in range(3):
            blocks = []
            for j in range(depth[i]):
                if j > depth[i] - global_block[i] - 1:
                    if global_block_type[i] != 'LGFI':
                        raise NotImplementedError
                    blocks.append(LGFI(dim=dims[i], drop_path=dp[cur + j], expan_ratio=expan_ratio, use_pos_emb=use_pos_embd_xca[i],
                                       num_heads=heads[i], layer_scale_init_value=layer_scale_init_value))
                else:
                    blocks.append(DilatedConv(dim=dims[i], k=3, dilation=self.dilation[i][j], drop_path=dp[cur + j],
                                              layer_scale_init_value=layer_scale_init_value, expan_ratio=expan_ratio))
            self.stages.append(nn.Sequential(*blocks))
            cur += depth[i]
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, (LayerNorm, nn.LayerNorm)):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def forward_features(self, x):
        x = (x - 0.45) / 0.225
        pooled = [p(x) for p in self.input_downsample]            # the image at 1/2 .. 1/16
        x = self.stem2(torch.cat((self.downsample_layers[0](x), pooled[0]), dim=1))
        features, carry = [], [x]
        for i in range(3):
            if i > 0:                                              # previous stage's input + output + the pooled image
                carry.append(pooled[i])
                x = self.downsample_layers[i](torch.cat(carry, dim=1))
                carry = [x]
            x = self.stages[i](x)
            carry.append(x)
            features.append(x)
        return features

    def forward(self, x):
        return self.forward_features(x)


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = 'bilinear'
        self.scales = scales
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = (self.num_ch_enc / 2).astype('int')
        self.convs = OrderedDict()
        for i in range(2, -1, -1):
            cin = self.num_ch_enc[-1] if i == 2 else self.num_ch_dec[i + 1]
            self.convs[("upconv", i, 0)] = ConvBlock(cin, self.num_ch_dec[i])
            cin = self.num_ch_dec[i] + (self.num_ch_enc[i - 1] if (self.use_skips and i > 0) else 0)
            self.convs[("upconv", i, 1)] = ConvBlock(cin, self.num_ch_dec[i])
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(self.num_ch_dec[s], self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, input_features):
        self.outputs = {}
        x = input_features[-1]
        for i in range(2, -1, -1):
            x = [upsample(self.convs[("upconv", i, 0)](x), mode='bilinear')]
            if self.use_skips and i > 0:
                x.append(input_features[i - 1])
            x = self.convs[("upconv", i, 1)](torch.cat(x, 1))
            if i in self.scales:
                self.outputs[("disp", i)] = self.sigmoid(upsample(self.convs[("dispconv", i)](x), mode='bilinear'))
        return self.outputs
