"""TEST INFRASTRUCTURE: an independent pure-Python `.tflite` reader and a CPU graph
executor used as the whole-model checker. LCE custom ops run on the CPU oracle
(tests/lce_testlib.py); the float builtins run as plain PyTorch fp32 on the CPU (the
"torch fp32 reference" for floating-point kernels). Nothing here is imported by the
product.
"""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.nn.functional as F

import lce_testlib as L

_TYPES = {0: np.float32, 2: np.int32, 3: np.uint8, 4: np.int64, 6: np.bool_, 9: np.int8}


class _T:
    def __init__(self, buf, pos):
        self.b, self.pos = buf, pos

    def _field(self, fid):
        if not self.pos:
            return 0
        vt = self.pos - struct.unpack_from("<i", self.b, self.pos)[0]
        vlen = struct.unpack_from("<H", self.b, vt)[0]
        slot = 4 + 2 * fid
        if slot + 2 > vlen:
            return 0
        off = struct.unpack_from("<H", self.b, vt + slot)[0]
        return self.pos + off if off else 0

    def scalar(self, fid, fmt, default=0):
        f = self._field(fid)
        return struct.unpack_from("<" + fmt, self.b, f)[0] if f else default

    def _ind(self, fid):
        f = self._field(fid)
        return f + struct.unpack_from("<I", self.b, f)[0] if f else 0

    def table(self, fid):
        return _T(self.b, self._ind(fid))

    def vector(self, fid, fmt):
        v = self._ind(fid)
        if not v:
            return []
        n = struct.unpack_from("<I", self.b, v)[0]
        return list(struct.unpack_from(f"<{n}{fmt}", self.b, v + 4))

    def bytes(self, fid):
        v = self._ind(fid)
        if not v:
            return b""
        n = struct.unpack_from("<I", self.b, v)[0]
        return bytes(self.b[v + 4:v + 4 + n])

    def string(self, fid):
        return self.bytes(fid).decode()

    def tables(self, fid):
        v = self._ind(fid)
        if not v:
            return []
        n = struct.unpack_from("<I", self.b, v)[0]
        out = []
        for i in range(n):
            slot = v + 4 + 4 * i
            out.append(_T(self.b, slot + struct.unpack_from("<I", self.b, slot)[0]))
        return out


def parse(model_bytes: bytes):
    b = memoryview(model_bytes)
    assert bytes(b[4:8]) == b"TFL3"
    model = _T(b, struct.unpack_from("<I", b, 0)[0])
    codes = []
    for c in model.tables(1):
        bc = c.scalar(3, "i") or c.scalar(0, "b")
        codes.append((bc, c.string(1)))
    buffers = [t.bytes(0) for t in model.tables(4)]
    sg = model.tables(2)[0]
    tensors = []
    for t in sg.tables(0):
        q = t.table(4)
        scales = q.vector(2, "f") if q.pos else []
        zps = q.vector(3, "q") if q.pos else []
        dtype = _TYPES[t.scalar(1, "b")]
        data = buffers[t.scalar(2, "I")]
        shape = tuple(t.vector(0, "i"))
        tensors.append({"shape": shape, "dtype": dtype, "name": t.string(3),
                        "data": np.frombuffer(data, dtype).reshape(shape).copy() if data else None,
                        "scale": scales[0] if scales else None,
                        "zero_point": zps[0] if zps else 0})
    ops = []
    for o in sg.tables(3):
        bc, custom = codes[o.scalar(0, "I")]
        ops.append({"code": bc, "custom": custom, "inputs": o.vector(1, "i"),
                    "outputs": o.vector(2, "i"), "options": o.table(4),
                    "custom_options": o.bytes(5)})
    return {"version": model.scalar(0, "I"), "tensors": tensors, "ops": ops,
            "inputs": sg.vector(1, "i"), "outputs": sg.vector(2, "i"),
            "description": model.string(3)}


def _flex_ints(blob):
    """Decode a flexbuffer map of ints (independent of the C++ reader)."""
    if not blob:
        return {}
    w = blob[-1]
    packed = blob[-2]
    slot = len(blob) - 2 - w
    p = slot - int.from_bytes(blob[slot:slot + w], "little")
    ew = 1 << (packed & 3)
    n = int.from_bytes(blob[p - ew:p], "little")
    kw = int.from_bytes(blob[p - 2 * ew:p - ew], "little")
    kslot = p - 3 * ew
    keys = kslot - int.from_bytes(blob[kslot:kslot + ew], "little")
    out = {}
    for i in range(n):
        ks = keys + i * kw
        ka = ks - int.from_bytes(blob[ks:ks + kw], "little")
        key = bytes(blob[ka:blob.index(0, ka)]).decode()
        out[key] = int.from_bytes(blob[p + i * ew:p + (i + 1) * ew], "little", signed=True)
    return out


def _same_pads(size, k, stride, dil=1):
    out = -(-size // stride)
    total = max(0, (out - 1) * stride + (k - 1) * dil + 1 - size)
    return total // 2, total - total // 2


def _act(x, a):
    if a == 1:
        return torch.relu(x)
    if a == 2:
        return torch.clamp(x, -1, 1)
    if a == 3:
        return torch.clamp(x, 0, 6)
    return x


def _conv(x, w, bias, o, depthwise):
    pad = o.scalar(0, "b")
    sw, sh = o.scalar(1, "i"), o.scalar(2, "i")
    if depthwise:
        act, dw, dh = o.scalar(4, "b"), o.scalar(5, "i", 1), o.scalar(6, "i", 1)
    else:
        act, dw, dh = o.scalar(3, "b"), o.scalar(4, "i", 1), o.scalar(5, "i", 1)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    if depthwise:
        c = w.shape[3]
        wt = torch.from_numpy(w).permute(3, 0, 1, 2)          # [C,1,kh,kw]
        groups = c
    else:
        wt = torch.from_numpy(w).permute(0, 3, 1, 2)          # [O,I,kh,kw]
        groups = 1
    kh, kw = wt.shape[2], wt.shape[3]
    if pad == 0:
        pt, pb = _same_pads(x.shape[1], kh, sh, dh)
        pl, pr = _same_pads(x.shape[2], kw, sw, dw)
        xt = F.pad(xt, (pl, pr, pt, pb))
    y = F.conv2d(xt, wt, None if bias is None else torch.from_numpy(bias), (sh, sw), 0, (dh, dw),
                 groups)
    return _act(y, act).permute(0, 2, 3, 1).contiguous().numpy()


def _pool(x, o, is_max):
    pad = o.scalar(0, "b")
    sw, sh, fw, fh, act = (o.scalar(1, "i"), o.scalar(2, "i"), o.scalar(3, "i"), o.scalar(4, "i"),
                           o.scalar(5, "b"))
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    if pad == 0:
        pt, pb = _same_pads(x.shape[1], fh, sh)
        pl, pr = _same_pads(x.shape[2], fw, sw)
    else:
        pt = pb = pl = pr = 0
    if is_max:
        xt = F.pad(xt, (pl, pr, pt, pb), value=float("-inf"))
        y = F.max_pool2d(xt, (fh, fw), (sh, sw))
    else:
        ones = F.pad(torch.ones_like(xt[:, :1]), (pl, pr, pt, pb))
        s = F.avg_pool2d(F.pad(xt, (pl, pr, pt, pb)), (fh, fw), (sh, sw), divisor_override=1)
        cnt = F.avg_pool2d(ones, (fh, fw), (sh, sw), divisor_override=1)
        y = s / cnt
    return _act(y, act).permute(0, 2, 3, 1).contiguous().numpy()


def run(model, inputs, threads=8, lce_impl="oracle", bconv_kind=1):
    """Execute the parsed model on the CPU. `inputs`: list of arrays (any batch)."""
    vals = {}
    for i, t in enumerate(model["tensors"]):
        if t["data"] is not None:
            vals[i] = t["data"]
    for idx, a in zip(model["inputs"], inputs):
        vals[idx] = np.ascontiguousarray(a)
    for op in model["ops"]:
        ins = [vals.get(i) if i >= 0 else None for i in op["inputs"]]
        o = op["options"]
        code, custom = op["code"], op["custom"]
        if code == 32 and custom == "LceQuantize":
            zp = model["tensors"][op["inputs"][0]]["zero_point"]
            out = L.quantize(ins[0], zp, impl=lce_impl)
        elif code == 32 and custom == "LceDequantize":
            ot = model["tensors"][op["outputs"][0]]
            kind = {np.float32: L.T_FLOAT, np.int8: L.T_INT8, np.bool_: L.T_BOOL}[ot["dtype"]]
            out = L.dequantize(ins[0], ot["shape"][-1], kind, ot["scale"] or 1.0, ot["zero_point"])
        elif code == 32 and custom == "LceBconv2d":
            a = _flex_ints(op["custom_options"])
            x, filt = ins[0], ins[1]
            ot = model["tensors"][op["outputs"][0]]
            out_type = {np.float32: L.OUT_FLOAT, np.int8: L.OUT_INT8, np.int32: L.OUT_BITPACKED}[
                ot["dtype"]]
            cin = a["channels_in"]
            groups = L.cdiv(cin, 32) // filt.shape[3]
            d = L.BconvDesc(x.shape[0], x.shape[1], x.shape[2], cin, filt.shape[1], filt.shape[2],
                            filt.shape[0], groups, a["stride_height"], a["stride_width"],
                            a["dilation_height_factor"], a["dilation_width_factor"], a["padding"],
                            a["pad_values"], a["fused_activation_function"], out_type,
                            float(ot["scale"] or 1.0), int(ot["zero_point"]))
            out = L.bconv2d(d, x, filt, ins[2], ins[3], ins[4], impl=lce_impl, kind=bconv_kind,
                            threads=threads)
        elif code == 32 and custom == "LceBMaxPool2d":
            a = _flex_ints(op["custom_options"])
            x = ins[0]
            d = L.BMaxPoolDesc(x.shape[0], x.shape[1], x.shape[2], x.shape[3], a["filter_height"],
                               a["filter_width"], a["stride_height"], a["stride_width"],
                               a["padding"])
            out = L.bmaxpool(d, x)
        elif code == 3:
            out = _conv(ins[0], ins[1], ins[2], o, False)
        elif code == 4:
            out = _conv(ins[0], ins[1], ins[2], o, True)
        elif code in (1, 17):
            out = _pool(ins[0], o, code == 17)
        elif code in (0, 18):
            a, b = torch.from_numpy(ins[0]), torch.from_numpy(np.asarray(ins[1]))
            out = _act(a + b if code == 0 else a * b, o.scalar(0, "b")).numpy()
        elif code == 19:
            out = np.maximum(ins[0], 0)
        elif code == 40:
            axes = tuple(int(v) for v in ins[1])
            out = torch.from_numpy(ins[0]).mean(dim=axes, keepdim=bool(o.scalar(0, "b"))).numpy()
        elif code == 9:
            x = torch.from_numpy(ins[0]).reshape(-1, ins[1].shape[1])
            y = x @ torch.from_numpy(ins[1]).T
            if ins[2] is not None:
                y = y + torch.from_numpy(ins[2])
            out = _act(y, o.scalar(0, "b")).numpy()
        elif code == 25:
            out = torch.softmax(torch.from_numpy(ins[0]) * o.scalar(0, "f", 1.0), dim=-1).numpy()
        elif code == 22:
            shape = o.vector(0, "i") or [int(v) for v in ins[1]]
            out = ins[0].reshape(shape)
        elif code in (34, 60):       # PAD / PADV2
            pads = np.asarray(ins[1]).reshape(-1, 2)
            fill = ins[2].reshape(()) if len(ins) > 2 and ins[2] is not None else 0
            out = np.pad(ins[0], [(int(a), int(b)) for a, b in pads], constant_values=fill)
        elif code == 6:              # DEQUANTIZE (reference/dequantize.h:32-49: the product in double)
            it = model["tensors"][op["inputs"][0]]
            out = (np.float64(np.float32(it["scale"])) *
                   (ins[0].astype(np.int64) - int(it["zero_point"])).astype(np.float64)).astype(np.float32)
        elif code == 2:              # CONCATENATION
            out = np.concatenate(ins, axis=o.scalar(0, "i"))
        else:
            raise NotImplementedError(f"op {code} {custom}")
        vals[op["outputs"][0]] = out
    return [vals[i] for i in model["outputs"]], vals
