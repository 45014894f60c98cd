"""Thin Python face of the generic fp32 layer kernels (csrc/kpd_layers.cu) used by the classes the reference exports but its
model never instantiates (`BACKBONE`, `KEYPOINT_HEAD`).  Activations travel as NHWC float32 tensors; BatchNorm
(eval mode) is folded into the preceding convolution here, on the device, and cached until a parameter changes.
PyTorch only allocates and reshapes; every arithmetic step of a forward is a libkpd_b200 kernel on the current stream."""
import torch

from .. import _native as N

ACT = {None: 0, "relu": 1, "hswish": 2, "hsigmoid": 3, "sigmoid": 4, "relu6": 5}


def fold_conv_bn(conv, bn=None):
    """(weight [kh][kw][cin/g][cout], bias [cout]) of conv (+ eval BatchNorm): w' = w * g / sqrt(var + eps), b' = (b - mean) * s + beta."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.size(0), device=w.device)
    if bn is not None:
        s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w.permute(2, 3, 1, 0).contiguous(), b.contiguous()


def bn_as_depthwise(bn):
    """A lone eval BatchNorm as a depthwise 1x1 conv: weight [1][1][1][C] = g / sqrt(var + eps), bias = beta - mean * s."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return s.view(1, 1, 1, -1).contiguous(), (bn.bias.detach().float() - bn.running_mean.detach().float() * s).contiguous()


class FoldCache:
    """Folded weights of one module tree, rebuilt when any parameter / buffer changes (version counter or storage)."""

    def __init__(self, module, builder):
        self._module, self._builder, self._sig, self._val = module, builder, None, None

    def get(self):
        ts = list(self._module.parameters()) + list(self._module.buffers())
        sig = tuple((t.data_ptr(), t._version) for t in ts)
        if sig != self._sig:
            with torch.no_grad():
                self._val = self._builder()
            self._sig = sig
        return self._val


def _dev(x):
    if not x.is_cuda:
        raise N.NativeLibraryError("move the module and its input to a CUDA device (no CPU fallback)")
    return x.device


def to_nhwc(x):
    return x.detach().float().permute(0, 2, 3, 1).contiguous()


def to_nchw_view(x):
    return x.permute(0, 3, 1, 2)


def conv2d(x, w, b, k, stride=1, pad=0, groups=1, act=None, in_scale=None, residual=None, act2=None):
    """x NHWC [n,h,w,cin]; w [k][k][cin/groups][cout]; -> NHWC [n,oh,ow,cout] = act2(act(conv + b) + residual)."""
    dev = _dev(x)
    n, h, ww, cin = x.shape
    cout = w.shape[-1]
    oh, ow = (h + 2 * pad - k) // stride + 1, (ww + 2 * pad - k) // stride + 1
    out = torch.empty((n, oh, ow, cout), device=dev, dtype=torch.float32)
    N.check(N.load().kpd_conv2d(N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(in_scale), N.ptr(residual), N.ptr(out), n, h, ww, cin, cout, k, stride,
                                pad, groups, ACT[act], ACT[act2], N.stream_ptr(dev)), "kpd_conv2d")
    return out


def adaptive_avgpool(x, oh, ow):
    dev = _dev(x)
    n, h, w, c = x.shape
    out = torch.empty((n, oh, ow, c), device=dev, dtype=torch.float32)
    N.check(N.load().kpd_adaptive_avgpool(N.ptr(x), N.ptr(out), n, h, w, c, oh, ow, N.stream_ptr(dev)), "kpd_adaptive_avgpool")
    return out


def linear(x, w, b=None, act=None):
    """x [m,k]; w [n,k] (torch layout) -> [m,n]."""
    dev = _dev(x)
    m, k = x.shape
    n = w.shape[0]
    out = torch.empty((m, n), device=dev, dtype=torch.float32)
    N.check(N.load().kpd_linear(N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(out), m, k, n, ACT[act], N.stream_ptr(dev)), "kpd_linear")
    return out


def layernorm(x, ln, act=None):
    dev = _dev(x)
    m, n = x.shape
    out = torch.empty_like(x)
    g = ln.weight.detach().float().contiguous() if ln.weight is not None else None
    b = ln.bias.detach().float().contiguous() if ln.bias is not None else None
    N.check(N.load().kpd_layernorm(N.ptr(x), N.ptr(g), N.ptr(b), float(ln.eps), N.ptr(out), m, n, ACT[act], N.stream_ptr(dev)), "kpd_layernorm")
    return out


def mul_map(x, amap):
    """x NHWC [n,h,w,c] * amap [n,h,w,1] broadcast over the channels."""
    dev = _dev(x)
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    N.check(N.load().kpd_mul_map(N.ptr(x), N.ptr(amap), N.ptr(out), n * h * w, c, N.stream_ptr(dev)), "kpd_mul_map")
    return out


def upsample_add_(dst, src):
    """dst += nearest-neighbour resize of src to dst's size (both NHWC, same n and c), in place."""
    dev = _dev(dst)
    n, h, w, c = dst.shape
    N.check(N.load().kpd_upsample_add(N.ptr(dst), N.ptr(src), n, h, w, src.shape[1], src.shape[2], c, N.stream_ptr(dev)), "kpd_upsample_add")
    return dst
