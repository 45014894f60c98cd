"""`KEYPOINT_HEAD` drop-in (reference: dll/models/keypoint_head.py:9-90) -- the coordinate-regression head the reference
package exports but `MultiPersonKeypointModel` never builds.  Same module tree / state_dict names; the
forward runs on the generic fp32 layer kernels (csrc/kpd_layers.cu), eval mode (Dropout = identity)."""
import torch
import torch.nn as nn

from ..configs import KeypointHeadConfig
from . import _layers as L


class _Slot(nn.Identity):
    """Placeholder for a parameter-free layer of the reference's Sequential (ReLU6, pooling, Flatten, Dropout, Sigmoid): it
    only keeps the numbering of the parameterised layers -- and with it the state_dict keys -- in place."""


def _conv_bn(cin, cout, k, pad):
    return [nn.Conv2d(cin, cout, kernel_size=k, padding=pad), nn.BatchNorm2d(cout), _Slot()]


class ResidualBlock(nn.Module):
    """keypoint_head.py:65-90: relu6(bn1(relu6(bn(conv3x3(x))))) + shortcut -> relu6; shortcut = conv1x1 + BN when widths differ."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Sequential(*_conv_bn(in_channels, out_channels, 3, 1))
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels))
        self._fold = L.FoldCache(self, lambda: (L.fold_conv_bn(self.conv1[0], self.conv1[1]), L.bn_as_depthwise(self.bn1),
                                                L.fold_conv_bn(self.downsample[0], self.downsample[1]) if self.downsample is not None else None))

    def run(self, xh):
        (w, b), (ws, bs), short = self._fold.get()
        y = L.conv2d(xh, w, b, 3, 1, 1, act="relu6")
        identity = xh if short is None else L.conv2d(xh, short[0], short[1], 1)
        c = y.shape[3]
        return L.conv2d(y, ws, bs, 1, groups=c, act="relu6", residual=identity, act2="relu6")       # bn1 -> relu6 -> + shortcut -> relu6

    def forward(self, x):
        return L.to_nchw_view(self.run(L.to_nhwc(x)))


def _flatten_as_nchw(weight, channels, cells):
    """Linear weight whose inputs were flattened from NCHW ([c][y][x]) re-ordered for the NHWC flattening ([y][x][c])."""
    return weight.detach().float().view(weight.shape[0], channels, cells).permute(0, 2, 1).reshape(weight.shape[0], -1).contiguous()


class KEYPOINT_HEAD(nn.Module):
    def __init__(self, config: KeypointHeadConfig):
        super().__init__()
        self.num_keypoints = config.num_keypoints
        c, r, v = config.in_channels, config.regression_channels, config.visibility_channels
        self._pool_r = (config.height // 4, config.width // 4)
        self.spatial_attention = nn.Sequential(nn.Conv2d(c, c // 2, kernel_size=1), _Slot(), nn.Conv2d(c // 2, 1, kernel_size=1), _Slot())
        self.regression_branch = nn.Sequential(
            ResidualBlock(c, config.fine_branch_channels), ResidualBlock(config.fine_branch_channels, r),
            *_conv_bn(r, r // 2, 3, 1), _Slot(), _Slot(),
            nn.Linear(r // 2 * self._pool_r[0] * self._pool_r[1], 256), nn.LayerNorm(256), _Slot(), _Slot(),
            nn.Linear(256, self.num_keypoints * 2))
        self.visibility_branch = nn.Sequential(
            *_conv_bn(c, v, 3, 1), _Slot(), _Slot(),
            nn.Linear(v * 16, 128), nn.LayerNorm(128), _Slot(), _Slot(),
            nn.Linear(128, self.num_keypoints * 3))
        self._fold = L.FoldCache(self, self._folded)

    def _folded(self):
        reg, vis, att = self.regression_branch, self.visibility_branch, self.spatial_attention
        cells_r = self._pool_r[0] * self._pool_r[1]
        f = lambda t: t.detach().float().contiguous()            # noqa: E731
        return {"att0": L.fold_conv_bn(att[0]), "att2": L.fold_conv_bn(att[2]),
                "reg_conv": L.fold_conv_bn(reg[2], reg[3]), "reg_fc1": (_flatten_as_nchw(reg[7].weight, reg[2].out_channels, cells_r), f(reg[7].bias)),
                "reg_fc2": (f(reg[11].weight), f(reg[11].bias)),
                "vis_conv": L.fold_conv_bn(vis[0], vis[1]), "vis_fc1": (_flatten_as_nchw(vis[5].weight, vis[0].out_channels, 16), f(vis[5].bias)),
                "vis_fc2": (f(vis[9].weight), f(vis[9].bias))}

    def forward(self, x):
        """x [B, in_channels, H, W] -> (keypoints [B, K, 2], visibility [B, K, 3]), both sigmoid outputs (keypoint_head.py:51-63)."""
        if self.training:
            raise RuntimeError("inference-only drop-in: call .eval()")
        W = self._fold.get()
        n = x.size(0)
        xh = L.to_nhwc(x)
        amap = L.conv2d(L.conv2d(xh, *W["att0"], 1, act="relu6"), *W["att2"], 1, act="sigmoid")
        xh = L.mul_map(xh, amap)
        reg, vis = self.regression_branch, self.visibility_branch
        y = reg[1].run(reg[0].run(xh))
        y = L.conv2d(y, *W["reg_conv"], 3, 1, 1, act="relu6")
        y = L.adaptive_avgpool(y, *self._pool_r).view(n, -1)
        y = L.layernorm(L.linear(y, *W["reg_fc1"]), reg[8], act="relu6")
        keypoints = L.linear(y, *W["reg_fc2"], act="sigmoid").view(n, self.num_keypoints, 2)
        z = L.conv2d(xh, *W["vis_conv"], 3, 1, 1, act="relu6")
        z = L.adaptive_avgpool(z, 4, 4).view(n, -1)
        z = L.layernorm(L.linear(z, *W["vis_fc1"]), vis[6], act="relu6")
        visibility = L.linear(z, *W["vis_fc2"], act="sigmoid").view(n, self.num_keypoints, 3)
        return keypoints, visibility
