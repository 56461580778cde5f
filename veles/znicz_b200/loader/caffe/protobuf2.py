"""Caffe protobuf messages without generated code.

The reference vendors 3.6 kLoC of ``protoc`` output
(/root/reference/loader/caffe/protobuf2.py) although only ``Datum`` is consumed
(/root/reference/loader/loader_lmdb.py:114-119). This module implements the protobuf
*wire format* directly (varint / 64-bit / length-delimited / 32-bit) with declarative
field tables, so ``Datum`` (LMDB records), ``BlobProto`` / ``BlobShape`` (mean files,
``.caffemodel`` weights) and the structural subset of ``NetParameter`` /
``LayerParameter`` (names, types, bottoms/tops, blobs, conv / pooling / LRN /
inner-product params) can be read and written. ``parse_net_text`` reads the prototxt
text format into nested dicts for model import.
"""
from __future__ import annotations

import re
import struct

import numpy

VARINT, FIXED64, BYTES, FIXED32 = 0, 1, 2, 5


# -- wire primitives ----------------------------------------------------------------------
def _read_varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _write_varint(out, value):
    if value < 0:
        value += 1 << 64
    while True:
        b = value & 0x7F
        value >>= 7
        if value:
            out.append(b | 0x80)
        else:
            out.append(b)
            return


def decode_message(buf):
    """→ list of (field_number, wire_type, value); BYTES values are memoryview slices."""
    buf = memoryview(buf)
    pos, n, out = 0, len(buf), []
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == VARINT:
            v, pos = _read_varint(buf, pos)
        elif wt == FIXED64:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == BYTES:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == FIXED32:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        if pos > n:
            raise ValueError("truncated message")
        out.append((field, wt, v))
    return out


def _signed(v, bits=64):
    return v - (1 << bits) if v >> (bits - 1) else v


# -- declarative messages -------------------------------------------------------------------
class Message(object):
    """FIELDS: number → (name, kind, repeated). kinds: int32 int64 uint32 bool float double
    bytes string enum or a Message subclass."""
    FIELDS = {}

    def __init__(self, **kwargs):
        for num, (name, kind, rep) in self.FIELDS.items():
            setattr(self, name, [] if rep else self._default(kind))
        for k, v in kwargs.items():
            if not any(f[0] == k for f in self.FIELDS.values()):
                raise AttributeError("%s has no field %s" % (type(self).__name__, k))
            setattr(self, k, v)

    @staticmethod
    def _default(kind):
        if isinstance(kind, type):
            return None
        return {"bytes": b"", "string": "", "bool": False, "float": 0.0,
                "double": 0.0}.get(kind, 0)

    # decoding -----------------------------------------------------------------------------
    @classmethod
    def FromString(cls, buf):
        m = cls()
        m.ParseFromString(buf)
        return m

    def ParseFromString(self, buf):
        for field, wt, v in decode_message(buf):
            spec = self.FIELDS.get(field)
            if spec is None:
                continue                     # unknown field: skipped (forward compatible)
            name, kind, rep = spec
            vals = self._decode_value(kind, wt, v)
            if rep:
                getattr(self, name).extend(vals)
            else:
                setattr(self, name, vals[-1])
        return self

    @staticmethod
    def _decode_value(kind, wt, v):
        if isinstance(kind, type):
            return [kind.FromString(v)]
        if kind in ("bytes", "string"):
            b = bytes(v)
            return [b.decode("utf-8", "replace") if kind == "string" else b]
        if kind == "float":
            if wt == BYTES:                  # packed
                return numpy.frombuffer(v, dtype="<f4").tolist()
            return [struct.unpack("<f", v)[0]]
        if kind == "double":
            if wt == BYTES:
                return numpy.frombuffer(v, dtype="<f8").tolist()
            return [struct.unpack("<d", v)[0]]
        # varint family
        if wt == BYTES:                      # packed varints
            out, pos, n = [], 0, len(v)
            while pos < n:
                x, pos = _read_varint(v, pos)
                out.append(x)
        else:
            out = [v]
        if kind == "bool":
            return [bool(x) for x in out]
        if kind in ("int32", "int64", "enum"):
            return [_signed(x) for x in out]
        return out

    # encoding -----------------------------------------------------------------------------
    def SerializeToString(self):
        out = bytearray()
        for num in sorted(self.FIELDS):
            name, kind, rep = self.FIELDS[num]
            val = getattr(self, name)
            if rep:
                if not len(val):
                    continue
                if kind in ("float", "double"):          # packed
                    raw = numpy.asarray(val, dtype="<f4" if kind == "float" else "<f8").tobytes()
                    _write_varint(out, (num << 3) | BYTES)
                    _write_varint(out, len(raw))
                    out += raw
                    continue
                for item in val:
                    self._encode_one(out, num, kind, item)
            else:
                if val is None or (val == self._default(kind) and not isinstance(kind, type)):
                    continue
                self._encode_one(out, num, kind, val)
        return bytes(out)

    @staticmethod
    def _encode_one(out, num, kind, val):
        if isinstance(kind, type):
            raw = val.SerializeToString()
            _write_varint(out, (num << 3) | BYTES)
            _write_varint(out, len(raw))
            out += raw
        elif kind in ("bytes", "string"):
            raw = val.encode("utf-8") if isinstance(val, str) else bytes(val)
            _write_varint(out, (num << 3) | BYTES)
            _write_varint(out, len(raw))
            out += raw
        elif kind == "float":
            _write_varint(out, (num << 3) | FIXED32)
            out += struct.pack("<f", val)
        elif kind == "double":
            _write_varint(out, (num << 3) | FIXED64)
            out += struct.pack("<d", val)
        else:
            _write_varint(out, (num << 3) | VARINT)
            _write_varint(out, int(val))

    def __repr__(self):
        parts = []
        for _, (name, kind, rep) in sorted(self.FIELDS.items()):
            v = getattr(self, name)
            if isinstance(v, (bytes, list)) and len(v) > 8:
                v = "<%d items>" % len(v)
            parts.append("%s=%r" % (name, v))
        return "%s(%s)" % (type(self).__name__, ", ".join(parts))


class BlobShape(Message):
    FIELDS = {1: ("dim", "int64", True)}


class BlobProto(Message):
    FIELDS = {1: ("num", "int32", False), 2: ("channels", "int32", False),
              3: ("height", "int32", False), 4: ("width", "int32", False),
              5: ("data", "float", True), 6: ("diff", "float", True),
              7: ("shape", BlobShape, False),
              8: ("double_data", "double", True), 9: ("double_diff", "double", True)}

    def to_array(self):
        if self.shape is not None and self.shape.dim:
            dims = tuple(self.shape.dim)
        else:
            dims = tuple(d for d in (self.num, self.channels, self.height, self.width))
            while len(dims) > 1 and dims[0] in (0, 1) and numpy.prod(dims[1:]) == len(self.data):
                dims = dims[1:]
        src = self.data if len(self.data) else self.double_data
        return numpy.asarray(src, dtype=numpy.float32).reshape(dims)

    @classmethod
    def from_array(cls, arr):
        arr = numpy.asarray(arr, dtype=numpy.float32)
        return cls(shape=BlobShape(dim=list(arr.shape)), data=arr.ravel().tolist())


class Datum(Message):
    """One LMDB record of a Caffe dataset: CHW uint8 ``data`` (or ``float_data``)."""
    FIELDS = {1: ("channels", "int32", False), 2: ("height", "int32", False),
              3: ("width", "int32", False), 4: ("data", "bytes", False),
              5: ("label", "int32", False), 6: ("float_data", "float", True),
              7: ("encoded", "bool", False)}

    def to_hwc(self, splitted_channels=True):
        """→ HWC array (CHW→HWC when the channels are stored as planes, the Caffe default:
        /root/reference/loader/loader_lmdb.py:95-100)."""
        if self.encoded:
            import cv2
            img = cv2.imdecode(numpy.frombuffer(self.data, numpy.uint8), cv2.IMREAD_UNCHANGED)
            return img if img.ndim == 3 else img[:, :, None]
        if len(self.data):
            flat = numpy.frombuffer(self.data, dtype=numpy.uint8)
        else:
            flat = numpy.asarray(self.float_data, dtype=numpy.float32)
        c, h, w = self.channels, self.height, self.width
        if splitted_channels:
            return numpy.ascontiguousarray(flat.reshape(c, h, w).transpose(1, 2, 0))
        return flat.reshape(h, w, c)

    @classmethod
    def from_hwc(cls, img, label=0):
        img = numpy.asarray(img)
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, c = img.shape
        chw = numpy.ascontiguousarray(img.transpose(2, 0, 1))
        if img.dtype == numpy.uint8:
            return cls(channels=c, height=h, width=w, data=chw.tobytes(), label=int(label))
        return cls(channels=c, height=h, width=w, label=int(label),
                   float_data=chw.astype(numpy.float32).ravel().tolist())


class ConvolutionParameter(Message):
    FIELDS = {1: ("num_output", "uint32", False), 2: ("bias_term", "bool", False),
              3: ("pad", "uint32", True), 4: ("kernel_size", "uint32", True),
              5: ("group", "uint32", False), 6: ("stride", "uint32", True),
              9: ("pad_h", "uint32", False), 10: ("pad_w", "uint32", False),
              11: ("kernel_h", "uint32", False), 12: ("kernel_w", "uint32", False),
              13: ("stride_h", "uint32", False), 14: ("stride_w", "uint32", False)}


class PoolingParameter(Message):
    FIELDS = {1: ("pool", "enum", False), 2: ("kernel_size", "uint32", False),
              3: ("stride", "uint32", False), 4: ("pad", "uint32", False),
              5: ("kernel_h", "uint32", False), 6: ("kernel_w", "uint32", False),
              7: ("stride_h", "uint32", False), 8: ("stride_w", "uint32", False),
              12: ("global_pooling", "bool", False)}
    MAX, AVE, STOCHASTIC = 0, 1, 2


class LRNParameter(Message):
    FIELDS = {1: ("local_size", "uint32", False), 2: ("alpha", "float", False),
              3: ("beta", "float", False), 4: ("norm_region", "enum", False),
              5: ("k", "float", False)}


class InnerProductParameter(Message):
    FIELDS = {1: ("num_output", "uint32", False), 2: ("bias_term", "bool", False)}


class DropoutParameter(Message):
    FIELDS = {1: ("dropout_ratio", "float", False)}


class LayerParameter(Message):
    FIELDS = {1: ("name", "string", False), 2: ("type", "string", False),
              3: ("bottom", "string", True), 4: ("top", "string", True),
              7: ("blobs", BlobProto, True),
              106: ("convolution_param", ConvolutionParameter, False),
              108: ("dropout_param", DropoutParameter, False),
              117: ("inner_product_param", InnerProductParameter, False),
              118: ("lrn_param", LRNParameter, False),
              121: ("pooling_param", PoolingParameter, False)}


class NetParameter(Message):
    FIELDS = {1: ("name", "string", False), 3: ("input", "string", True),
              4: ("input_dim", "int32", True), 100: ("layer", LayerParameter, True)}


class SolverParameter(Message):
    FIELDS = {24: ("net", "string", False), 5: ("base_lr", "float", False),
              7: ("max_iter", "int32", False), 8: ("lr_policy", "string", False),
              9: ("gamma", "float", False), 10: ("power", "float", False),
              11: ("momentum", "float", False), 12: ("weight_decay", "float", False),
              13: ("stepsize", "int32", False), 34: ("stepvalue", "int32", True)}


# -- prototxt (text format) -------------------------------------------------------------------
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|("(?:[^"\\]|\\.)*")|([{}:])|([^\s{}:#"]+))')


def parse_net_text(text):
    """prototxt → nested dict; repeated fields become lists."""
    tokens = []
    pos = 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None:
            break
        pos = m.end()
        if m.group(1):
            continue
        tokens.append(m.group(2) or m.group(3) or m.group(4))
    it = iter(tokens)

    def add(d, k, v):
        if k in d:
            if not isinstance(d[k], list):
                d[k] = [d[k]]
            d[k].append(v)
        else:
            d[k] = v

    def scalar(tok):
        if tok.startswith('"'):
            return tok[1:-1]
        if tok in ("true", "false"):
            return tok == "true"
        try:
            return int(tok)
        except ValueError:
            try:
                return float(tok)
            except ValueError:
                return tok

    def block():
        d = {}
        for tok in it:
            if tok == "}":
                return d
            nxt = next(it)
            if nxt == ":":
                nxt = next(it)
                if nxt == "{":
                    add(d, tok, block())
                else:
                    add(d, tok, scalar(nxt))
            elif nxt == "{":
                add(d, tok, block())
            else:
                raise ValueError("unexpected token %r after %r" % (nxt, tok))
        return d
    return block()
