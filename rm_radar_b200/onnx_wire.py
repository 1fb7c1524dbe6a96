"""Minimal ONNX protobuf wire-format reader (no `onnx` package in this image).

The reference resolves its "engine path" to a sibling ``.onnx`` and parses it with
nvonnxparser (``/root/reference/src/detect/detector.cpp:74-99,177-205``).  We read the same
files directly: the protobuf wire format is varint / length-delimited records and only a
handful of field numbers are needed.

Used by the engine builder (`rm_radar_b200.engine`) and by the oracle's torch interpreter.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def _varint(buf: memoryview, pos: int):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) for one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _packed_ints(val) -> list[int]:
    out = []
    pos = 0
    n = len(val)
    while pos < n:
        v, pos = _varint(val, pos)
        if v >= 1 << 63:
            v -= 1 << 64
        out.append(v)
    return out


def _sint(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


@dataclass
class Tensor:
    name: str = ""
    dims: tuple = ()
    data_type: int = 0
    array: np.ndarray | None = None


@dataclass
class Node:
    op: str = ""
    name: str = ""
    inputs: list = field(default_factory=list)
    outputs: list = field(default_factory=list)
    attrs: dict = field(default_factory=dict)


@dataclass
class ValueInfo:
    name: str = ""
    elem_type: int = 0
    shape: tuple = ()  # ints or str (dim_param)


@dataclass
class Graph:
    nodes: list
    initializers: dict
    inputs: list
    outputs: list
    metadata: dict
    producer: str = ""
    opset: int = 0


_DT = {1: np.float32, 7: np.int64, 6: np.int32, 10: np.float16, 11: np.float64, 9: np.bool_}


def _parse_tensor(buf) -> Tensor:
    t = Tensor()
    dims = []
    raw = None
    floats = []
    int64s = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dims.extend(_packed_ints(val) if wt == 2 else [_sint(val)])
        elif fno == 2:
            t.data_type = val
        elif fno == 4:  # float_data
            if wt == 2:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", bytes(val)))
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif fno == 7:  # int64_data
            int64s.extend(_packed_ints(val) if wt == 2 else [_sint(val)])
        elif fno == 8:
            t.name = bytes(val).decode()
        elif fno == 9:
            raw = bytes(val)
    t.dims = tuple(dims)
    dt = _DT[t.data_type]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt)
    elif floats:
        arr = np.asarray(floats, dtype=dt)
    elif int64s:
        arr = np.asarray(int64s, dtype=dt)
    else:
        arr = np.zeros(0, dtype=dt)
    t.array = arr.reshape(t.dims).copy()
    return t


def _parse_attr(buf):
    name = ""
    f = i = s = t = None
    floats = []
    ints = []
    atype = 0
    for fno, wt, val in _fields(buf):
        if fno == 1:
            name = bytes(val).decode()
        elif fno == 2:
            f = struct.unpack("<f", val)[0]
        elif fno == 3:
            i = _sint(val)
        elif fno == 4:
            s = bytes(val)
        elif fno == 5:
            t = _parse_tensor(val)
        elif fno == 7:
            if wt == 2:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", bytes(val)))
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif fno == 8:
            ints.extend(_packed_ints(val) if wt == 2 else [_sint(val)])
        elif fno == 20:
            atype = val
    if atype == 1:
        return name, f
    if atype == 2:
        return name, i
    if atype == 3:
        return name, s.decode()
    if atype == 4:
        return name, t
    if atype == 6:
        return name, floats
    if atype == 7:
        return name, ints
    # untyped fall-backs
    for v in (t, s, f, i):
        if v is not None:
            return name, v
    return name, ints or floats


def _parse_node(buf) -> Node:
    n = Node()
    for fno, wt, val in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(val).decode())
        elif fno == 2:
            n.outputs.append(bytes(val).decode())
        elif fno == 3:
            n.name = bytes(val).decode()
        elif fno == 4:
            n.op = bytes(val).decode()
        elif fno == 5:
            k, v = _parse_attr(val)
            n.attrs[k] = v
    return n


def _parse_value_info(buf) -> ValueInfo:
    vi = ValueInfo()
    for fno, wt, val in _fields(buf):
        if fno == 1:
            vi.name = bytes(val).decode()
        elif fno == 2:  # TypeProto
            for f2, _, v2 in _fields(val):
                if f2 == 1:  # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            vi.elem_type = v3
                        elif f3 == 2:  # shape
                            dims = []
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:
                                    d = None
                                    for f5, w5, v5 in _fields(v4):
                                        if f5 == 1:
                                            d = _sint(v5)
                                        elif f5 == 2:
                                            d = bytes(v5).decode()
                                    dims.append(d)
                            vi.shape = tuple(dims)
    return vi


def load(path: str) -> Graph:
    with open(path, "rb") as fh:
        data = memoryview(fh.read())
    graph_buf = None
    meta = {}
    producer = ""
    opset = 0
    for fno, wt, val in _fields(data):
        if fno == 7:
            graph_buf = val
        elif fno == 2:
            producer = bytes(val).decode()
        elif fno == 8:
            for f2, _, v2 in _fields(val):
                if f2 == 2:
                    opset = max(opset, v2)
        elif fno == 14:
            k = v = ""
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    k = bytes(v2).decode()
                elif f2 == 2:
                    v = bytes(v2).decode()
            meta[k] = v
    if graph_buf is None:
        raise ValueError(f"{path}: no GraphProto")
    nodes, inits, inputs, outputs = [], {}, [], []
    for fno, wt, val in _fields(graph_buf):
        if fno == 1:
            nodes.append(_parse_node(val))
        elif fno == 5:
            t = _parse_tensor(val)
            inits[t.name] = t.array
        elif fno == 11:
            inputs.append(_parse_value_info(val))
        elif fno == 12:
            outputs.append(_parse_value_info(val))
    inputs = [v for v in inputs if v.name not in inits]
    return Graph(nodes, inits, inputs, outputs, meta, producer, opset)
