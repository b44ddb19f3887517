"""Synthetic model zoo: the three model families BASELINE.json names, emitted as real
`.tflite` graphs with the op sequence the LCE converter produces (LceQuantize ->
LceBconv2d with folded BatchNorm multiplier/bias and fused ReLU -> ADD shortcut ...),
random weights (seeded), correct topology.

The topologies are NOT in /root/reference (they live in larq-zoo, absent offline);
they are restated from the public QuickNet / Bi-Real Net descriptions:
  QuickNet(-Large): stem [conv3x3 s2 -> BN/ReLU -> depthwise3x3 s2 -> conv1x1 -> BN],
    sections of residual blocks [sign -> binary conv3x3 (one-padding, ReLU fused, BN
    folded) -> + shortcut] with filters (64,128,256,512), blocks (4,4,4,4) / (6,8,12,6),
    transitions [ReLU? no: maxpool2x2 s1 -> blur depthwise3x3 s2 -> conv1x1 -> BN],
    head [ReLU -> global average pool -> dense 1000 -> softmax].
  Bi-RealNet-18: conv7x7 s2 + BN + maxpool3x3 s2; 16 binary conv3x3 (zero padding,
    per-channel scale folded into the multiplier) each with its own shortcut,
    downsampling shortcuts avgpool2x2 s2 -> conv1x1 -> BN; head as above.
"""
from __future__ import annotations

import numpy as np

# Pure Python on purpose (numpy + the flatbuffer / flexbuffer writers of tflite_writer.py): no
# native library is loaded, so bench.py's reference arm can synthesise the very same model bytes
# by loading these two files alone (importlib, not the package).
try:
    from .tflite_writer import TFLiteModel, bconv2d_options
except ImportError:      # loaded as a stand-alone file next to tflite_writer.py
    from tflite_writer import TFLiteModel, bconv2d_options


# Quantisation of an int8 image input (`input_type="int8"`): what the converter attaches to the
# input tensor with inference_input_type=tf.int8; N(0,1) images span about +-4.
INPUT_SCALE, INPUT_ZERO_POINT = 4.0 / 127.0, 0


def quantize_images(x, scale=INPUT_SCALE, zero_point=INPUT_ZERO_POINT):
    """float images -> the int8 tensor an int8-input model expects."""
    return np.clip(np.rint(np.asarray(x, np.float32) / np.float32(scale)) + zero_point, -128, 127).astype(np.int8)


class _Builder:
    def __init__(self, seed, batch, name):
        self.rng = np.random.default_rng(seed)
        self.m = TFLiteModel(name)
        self.batch = batch
        self.n = 0

    def _name(self, p):
        self.n += 1
        return f"{p}_{self.n}"

    def act(self, shape, dtype=np.float32):
        return self.m.add_tensor(self._name("t"), shape, dtype)

    def image_input(self, shape, input_type):
        """The graph's input: float32 images, or int8 images + the DEQUANTIZE the converter
        places behind an int8 input (then a step ships a quarter of the bytes to the device)."""
        if input_type in ("float32", np.float32, None):
            x = self.m.add_tensor("input", shape, np.float32)
            self.m.inputs = [x]
            return x
        if input_type not in ("int8", np.int8):
            raise ValueError(f"unsupported input_type {input_type!r}")
        q = self.m.add_tensor("input", shape, np.int8, scale=INPUT_SCALE, zero_point=INPUT_ZERO_POINT)
        self.m.inputs = [q]
        x = self.act(shape)
        self.m.add_op("DEQUANTIZE", [q], [x])
        return x

    def const(self, data, dtype=np.float32, name="w"):
        return self.m.add_tensor(self._name(name), None, dtype, data=data)

    def bn(self, c):
        """BatchNorm folded to y = x * scale + shift."""
        scale = self.rng.uniform(0.5, 1.5, c).astype(np.float32)
        shift = self.rng.uniform(-0.5, 0.5, c).astype(np.float32)
        return scale, shift

    # ---- float builtins ----
    def conv(self, x, xs, cout, k, stride, padding="SAME", activation="NONE", bn=True):
        b, h, w, cin = xs
        wt = self.rng.standard_normal((cout, k, k, cin)).astype(np.float32) * np.float32(
            1.0 / np.sqrt(k * k * cin))
        bias = np.zeros(cout, np.float32)
        if bn:
            s, t = self.bn(cout)
            wt = wt * s[:, None, None, None]
            bias = t
        oh = -(-h // stride) if padding == "SAME" else (h - k) // stride + 1
        ow = -(-w // stride) if padding == "SAME" else (w - k) // stride + 1
        y = self.act((b, oh, ow, cout))
        self.m.add_op("CONV_2D", [x, self.const(wt), self.const(bias, name="b")], [y],
                      padding=padding, stride=(stride, stride), activation=activation)
        return y, (b, oh, ow, cout)

    def depthwise(self, x, xs, k, stride, padding="SAME", blur=False, bn=False):
        b, h, w, c = xs
        if blur:   # blurpool: fixed binomial kernel
            a = np.array([1.0, 2.0, 1.0], np.float32)
            kern = np.outer(a, a) / 16.0
            wt = np.broadcast_to(kern[None, :, :, None], (1, k, k, c)).astype(np.float32).copy()
        else:
            wt = self.rng.standard_normal((1, k, k, c)).astype(np.float32) / np.float32(k)
        bias = np.zeros(c, np.float32)
        if bn:
            s, t = self.bn(c)
            wt = wt * s[None, None, None, :]
            bias = t
        oh = -(-h // stride) if padding == "SAME" else (h - k) // stride + 1
        ow = -(-w // stride) if padding == "SAME" else (w - k) // stride + 1
        y = self.act((b, oh, ow, c))
        self.m.add_op("DEPTHWISE_CONV_2D", [x, self.const(wt), self.const(bias, name="b")], [y],
                      padding=padding, stride=(stride, stride), depth_multiplier=1)
        return y, (b, oh, ow, c)

    def pool(self, kind, x, xs, k, stride, padding):
        b, h, w, c = xs
        oh = -(-h // stride) if padding == "SAME" else (h - k) // stride + 1
        ow = -(-w // stride) if padding == "SAME" else (w - k) // stride + 1
        y = self.act((b, oh, ow, c))
        self.m.add_op(kind, [x], [y], padding=padding, stride=(stride, stride), filter=(k, k))
        return y, (b, oh, ow, c)

    def add(self, a, b_, shape, activation="NONE"):
        y = self.act(shape)
        self.m.add_op("ADD", [a, b_], [y], activation=activation)
        return y

    def relu(self, x, shape):
        y = self.act(shape)
        self.m.add_op("RELU", [x], [y])
        return y

    # ---- LCE custom ops ----
    def quantize(self, x, xs):
        b, h, w, c = xs
        y = self.act((b, h, w, (c + 31) // 32), np.int32)
        self.m.add_op("LceQuantize", [x], [y], custom_options=b"")
        return y

    def bconv(self, xq, xs, cout, k=3, stride=1, pad_values=1, activation="RELU"):
        """LceBconv2d with float output; BN (and for Bi-RealNet the kernel scale)
        folded into post_activation_multiplier / bias like the converter does
        (LCE/mlir/transforms/optimize_patterns_common.td:39-113)."""
        b, h, w, cin = xs
        signs = self.rng.integers(0, 2, (cout, k, k, cin), dtype=np.int8) * 2 - 1
        bits = np.zeros((cout, k, k, ((cin + 31) // 32) * 32), np.uint32)
        bits[..., :cin] = signs < 0
        filt = (bits.reshape(cout, k, k, -1, 32) << np.arange(32, dtype=np.uint32)).sum(
            -1, dtype=np.uint64).astype(np.uint32).view(np.int32)
        s, t = self.bn(cout)
        mul = (s / np.float32(k * k * cin) * np.float32(4.0)).astype(np.float32)
        oh, ow = -(-h // stride), -(-w // stride)
        y = self.act((b, oh, ow, cout))
        act_code = {"NONE": 0, "RELU": 1}[activation]
        opts = bconv2d_options(cin, (stride, stride), (1, 1), 0, pad_values, act_code)
        self.m.add_op("LceBconv2d",
                      [xq, self.const(filt, np.int32, "bfilter"), self.const(mul, name="pmul"),
                       self.const(t, name="pbias"), -1], [y], custom_options=opts)
        return y, (b, oh, ow, cout)

    def head(self, x, xs, classes=1000, relu=True):
        b, h, w, c = xs
        if relu:
            x = self.relu(x, xs)
        axes = self.const(np.array([1, 2], np.int32), np.int32, "axes")
        pooled = self.act((b, c))
        self.m.add_op("MEAN", [x, axes], [pooled], keep_dims=False)
        wt = self.rng.standard_normal((classes, c)).astype(np.float32) / np.float32(np.sqrt(c))
        logits = self.act((b, classes))
        self.m.add_op("FULLY_CONNECTED", [pooled, self.const(wt),
                                          self.const(np.zeros(classes, np.float32), name="b")],
                      [logits])
        probs = self.act((b, classes))
        self.m.add_op("SOFTMAX", [logits], [probs], beta=1.0)
        return probs


def quicknet(batch=1, blocks=(4, 4, 4, 4), filters=(64, 128, 256, 512), seed=0, image=224,
             name="QuickNet", input_type="float32"):
    """Returns the serialized .tflite bytes."""
    g = _Builder(seed, batch, name)
    xs = (batch, image, image, 3)
    x = g.image_input(xs, input_type)
    # stem
    x, xs = g.conv(x, xs, filters[0] // 4, 3, 2, activation="RELU")
    x, xs = g.depthwise(x, xs, 3, 2, bn=True)
    x, xs = g.conv(x, xs, filters[0], 1, 1)
    for si, (nb, f) in enumerate(zip(blocks, filters)):
        if si > 0:   # transition block
            x, xs = g.pool("MAX_POOL_2D", x, xs, 2, 1, "VALID")
            x, xs = g.depthwise(x, xs, 3, 2, blur=True)
            x, xs = g.conv(x, xs, f, 1, 1)
        for _ in range(nb):
            q = g.quantize(x, xs)
            y, _ = g.bconv(q, xs, f, pad_values=1, activation="RELU")
            x = g.add(y, x, xs)
    probs = g.head(x, xs)
    g.m.outputs = [probs]
    return g.m.serialize()


def quicknet_large(batch=1, seed=0, image=224, input_type="float32"):
    return quicknet(batch, blocks=(6, 8, 12, 6), seed=seed, image=image, name="QuickNetLarge",
                    input_type=input_type)


def birealnet18(batch=1, seed=0, image=224, input_type="float32"):
    g = _Builder(seed, batch, "BiRealNet18")
    xs = (batch, image, image, 3)
    x = g.image_input(xs, input_type)
    x, xs = g.conv(x, xs, 64, 7, 2)
    x, xs = g.pool("MAX_POOL_2D", x, xs, 3, 2, "SAME")
    cin = 64
    for stage, f in enumerate((64, 128, 256, 512)):
        for blk in range(4):
            stride = 2 if (stage > 0 and blk == 0) else 1
            q = g.quantize(x, xs)
            y, ys = g.bconv(q, xs, f, stride=stride, pad_values=0, activation="NONE")
            if stride == 2 or cin != f:
                s, ss = g.pool("AVERAGE_POOL_2D", x, xs, 2, 2, "VALID")
                s, ss = g.conv(s, ss, f, 1, 1)
            else:
                s = x
            x, xs = g.add(y, s, ys), ys
            cin = f
    probs = g.head(x, xs, relu=False)
    g.m.outputs = [probs]
    return g.m.serialize()


MODELS = {"quicknet": quicknet, "quicknet_large": quicknet_large, "birealnet18": birealnet18}
