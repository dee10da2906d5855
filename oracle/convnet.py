"""Oracle forward pass of the reference's encoder-decoder CNNs (torch CPU float32).

Test infrastructure only.  Restates, as a directly executed graph walk:
  sleap/nn/architectures/unet.py:132-278        (UNet block stacks, from_config)
  sleap/nn/architectures/encoder_decoder.py:94-144, 275-399, 508-676
  sleap/nn/architectures/hourglass.py:17-305
  sleap/nn/heads.py:42-63                        (1x1 linear heads)
  sleap/nn/model.py:312-364                      (head taps by output stride)
Keras/TF layer semantics (SAME padding, Conv2DTranspose, UpSampling2D, BatchNorm eps=1e-3)
restated per SURVEY Appendix A.14.  Conv numerics: "parity unpinned" by the reference's
tests (no activation values asserted anywhere); structure pinned by parameter counts.

Weights: dict ``{keras_layer_name: {"kernel", "bias"[, "gamma","beta","mean","var"]}}`` in
Keras layouts: Conv2D kernel (kh, kw, Cin, Cout); Conv2DTranspose kernel (kh, kw, Cout, Cin).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))


def _same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, b, stride=1):
    """Keras Conv2D(padding='same').  x NCHW, w Keras (kh,kw,Cin,Cout)."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb = _same_pad(x.shape[2], kh, stride)
    pl, pr = _same_pad(x.shape[3], kw, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, _t(w).permute(3, 2, 0, 1).contiguous(), None if b is None else _t(b), stride=stride)


def conv2d_transpose_same(x, w, b, stride=2):
    """Keras Conv2DTranspose(k, strides=2, padding='same'): out = stride*H; forward-conv pad 0|1."""
    k = w.shape[0]
    y = F.conv_transpose2d(x, _t(w).permute(3, 2, 0, 1).contiguous(), None if b is None else _t(b), stride=stride)
    H, W = x.shape[2] * stride, x.shape[3] * stride
    total = max((x.shape[2] - 1) * stride + k - H, 0)
    p0 = total // 2
    return y[:, :, p0:p0 + H, p0:p0 + W]


def maxpool2_same(x, stride=2):
    pt, pb = _same_pad(x.shape[2], 2, stride)
    pl, pr = _same_pad(x.shape[3], 2, stride)
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(x, 2, stride)


def upsample2(x, method):
    if method == "nearest":
        return F.interpolate(x, scale_factor=2, mode="nearest")
    # tf.image.resize bilinear, half-pixel centres == torch align_corners=False
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def batchnorm(x, p):
    scale = _t(p["gamma"]) / torch.sqrt(_t(p["var"]) + BN_EPS)
    shift = _t(p["beta"]) - _t(p["mean"]) * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


class _Walker:
    def __init__(self, weights):
        self.w = weights

    def conv(self, x, name, stride=1, relu=True):
        p = self.w[name]
        y = conv2d_same(x, p["kernel"], p.get("bias"), stride)
        return F.relu(y) if relu else y

    def tconv(self, x, name):
        p = self.w[name]
        return F.relu(conv2d_transpose_same(x, p["kernel"], p.get("bias"), 2))

    def conv_relu_bn(self, x, prefix, k=3, stride=1):
        y = self.conv(x, prefix + "_conv", stride=stride, relu=True)
        return batchnorm(y, self.w[prefix + "_bn"])


def unet_blocks(cfg):
    """UNet.from_config (unet.py:250-278) -> block counts."""
    stem_blocks = 0
    if cfg.get("stem_stride"):
        stem_blocks = int(math.log2(cfg["stem_stride"]))
    down_blocks = int(math.log2(cfg["max_stride"])) - stem_blocks
    up_blocks = int(math.log2(cfg["max_stride"] / cfg["output_stride"]))
    return stem_blocks, down_blocks, up_blocks


def unet_forward(x, cfg, weights):
    """x NCHW float32 -> (list of per-stack outputs, list of per-stack intermediate (tensor, stride))."""
    wk = _Walker(weights)
    filters, rate = cfg["filters"], cfg.get("filters_rate", 2)
    convs = cfg.get("convs_per_block", 2)
    middle = cfg.get("middle_block", True)
    interp = cfg.get("up_interpolate", False)
    contraction = cfg.get("block_contraction", False)
    stacks = cfg.get("stacks", 1)
    stem_blocks, down_blocks, up_blocks = unet_blocks(cfg)

    stride = 1
    stem_output = []
    if stem_blocks > 0:
        for b in range(stem_blocks):
            if b > 0:
                x = maxpool2_same(x)
            for i in range(convs):
                x = wk.conv(x, f"stem{b}_conv{i}")
        x = maxpool2_same(x)
        stride = 2 ** stem_blocks
        stem_output = [(x, stride)]
    stem_stride = stride

    outs, mids = [], []
    for s in range(stacks):
        cur = stem_stride
        feats = []

        def note(t, st):
            if st not in [f[1] for f in feats]:
                feats.append((t, st))

        bi = 0
        for b in range(down_blocks):
            if b > 0:
                x = maxpool2_same(x)
                cur *= 2
            for i in range(convs):
                x = wk.conv(x, f"stack{s}_enc{bi}_conv{i}")
            note(x, cur)
            bi += 1
        x = maxpool2_same(x)   # PoolingBlock "_last_pool"
        cur *= 2
        note(x, cur)
        bi += 1
        if middle:
            if convs > 1:
                for i in range(convs - 1):
                    x = wk.conv(x, f"stack{s}_enc{bi}_middle_expand_conv{i}")
                note(x, cur)
                bi += 1
            x = wk.conv(x, f"stack{s}_enc{bi}_middle_contract_conv0")
            note(x, cur)
            bi += 1
        skips = stem_output + feats[:-1]

        inter = []
        for b in range(up_blocks):
            inter.append((x, cur))
            nxt = cur // 2
            prefix = f"stack{s}_dec{b}_s{cur}_to_s{nxt}"
            if interp:
                x = upsample2(x, "bilinear")
            else:
                x = wk.tconv(x, prefix + "_trans_conv")
            skip = None
            for (t, st) in skips:
                if st == nxt:
                    skip = t
                    break
            if skip is not None:
                x = torch.cat([skip, x], dim=1)
            for i in range(convs):
                x = wk.conv(x, prefix + f"_refine_conv{i}")
            cur = nxt
        outs.append(x)
        mids.append(inter)
    return outs, mids, cur


def hourglass_forward(x, cfg, weights):
    wk = _Walker(weights)
    stem_stride = cfg.get("stem_stride", 4)
    stem_blocks = int(math.log2(stem_stride))
    down_blocks = int(math.log2(cfg.get("max_stride", 64))) - stem_blocks
    up_blocks = int(math.log2(cfg.get("max_stride", 64) / cfg.get("output_stride", 4)))
    filters = cfg.get("filters", 256)
    inc = cfg.get("filter_increase", 128)
    stacks = cfg.get("stacks", 3)

    x = wk.conv_relu_bn(x, "stem0_conv7x7", stride=2 if stem_stride == 4 else 1)
    x = wk.conv_relu_bn(x, "stem0_conv3x3")
    x = maxpool2_same(x, stride=2 if stem_stride > 1 else 1)
    x = wk.conv_relu_bn(x, "stem0_conv3x3_out")
    stem_output = [(x, stem_stride)]

    outs, mids = [], []
    cur = stem_stride
    for s in range(stacks):
        cur = stem_stride
        feats = []
        for b in range(down_blocks):
            x = maxpool2_same(x)
            cur *= 2
            x = wk.conv_relu_bn(x, f"stack{s}_enc{b}_conv")
            if cur not in [f[1] for f in feats]:
                feats.append((x, cur))
        skips = stem_output + feats[:-1]
        inter = []
        for b in range(up_blocks):
            inter.append((x, cur))
            nxt = cur // 2
            prefix = f"stack{s}_dec{b}"
            skip = [t for (t, st) in skips if st == nxt][0]
            x = wk.conv_relu_bn(x, prefix + "_conv")
            x = upsample2(x, "nearest")
            x = x + wk.conv_relu_bn(skip, prefix + "_skip")
            cur = nxt
        outs.append(x)
        mids.append(inter)
    return outs, mids, cur


def model_forward(images_nhwc, spec, weights, n_threads=None):
    """Model.make_model (model.py:312-364): backbone + heads; returns list of NHWC float32 arrays,
    one per head, evaluated on the LAST stack (inference.py:2885-2888).

    spec: {"backbone": "unet"|"hourglass", "backbone_cfg": {...},
           "heads": [{"name", "channels", "output_stride"}, ...]}
    """
    if n_threads:
        torch.set_num_threads(n_threads)
    x = _t(images_nhwc).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        if spec["backbone"] == "unet":
            outs, mids, out_stride = unet_forward(x, spec["backbone_cfg"], weights)
        else:
            outs, mids, out_stride = hourglass_forward(x, spec["backbone_cfg"], weights)
        res = []
        for h in spec["heads"]:
            if h["output_stride"] == out_stride:
                feat = outs[-1]
            else:
                cands = [t for (t, st) in mids[-1] if st == h["output_stride"]]
                if not cands:
                    raise ValueError(f"Could not find a feature activation for output at stride {h['output_stride']}.")
                feat = cands[0]
            p = weights[h["name"]]
            if h.get("vector"):
                # ClassVectorsHead.make_head (sleap/nn/heads.py:431-460): GlobalMaxPool2D (or Flatten in NHWC order) ->
                # (Dense + ReLU) x num_fc_layers -> Dense + softmax
                f = feat.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32)
                v = f.max(axis=(1, 2)) if h.get("global_pool", True) else f.reshape(len(f), -1)
                for i in range(int(h.get("num_fc_layers", 1))):
                    d = weights[f"pre_classification{i}_fc"]
                    v = np.maximum(v @ np.asarray(d["kernel"], np.float32) + np.asarray(d["bias"], np.float32), np.float32(0))
                z = v @ np.asarray(p["kernel"], np.float32) + np.asarray(p["bias"], np.float32)
                z = z - z.max(axis=1, keepdims=True)
                e = np.exp(z)
                res.append((e / e.sum(axis=1, keepdims=True)).astype(np.float32))
                continue
            y = conv2d_same(feat, p["kernel"], p.get("bias"), 1)
            res.append(y.permute(0, 2, 3, 1).contiguous().numpy())
    return res
