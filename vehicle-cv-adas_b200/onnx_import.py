"""onnx_import.py -- ingest the reference's model files: `.onnx` -> packed sm_100a plan (`.b200w`).

The reference hands `.onnx` / `.trt` files to ONNXRuntime / TensorRT (coreEngine.py:54-55,164-166); its models come from
ultralytics / yolov5 exports (README.md:53-58) and from `TrafficLaneDetector/convertPytorchToONNX.py:60-87` (UFLD).  This module
is the B200 replacement of that ingestion step: it reads the ONNX protobuf directly (the `onnx` package is
not a dependency -- the wire format is parsed here), recovers the convolution / linear / LayerNorm parameters, recognises the
architecture (YOLOv8 / YOLOv5 / UFLDv2, scale, class count, input size) and drives the same `plan.build_*` builders that the
state_dict path uses.  Nothing here runs the network: the graph is only a parameter container plus a shape oracle.

How parameters are matched to layers:
  * by NAME when the exporter kept module names.  ultralytics / yolov5 fuse Conv+BN in PyTorch before exporting, so their
    files carry `model.N...conv.weight` / `.conv.bias` (already folded) -- taken as they are;
  * by ORDER for convolutions whose names were lost: `torch.onnx.export` folds eval-mode BatchNorm into the preceding Conv and
    the folded tensors get anonymous names (`onnx::Conv_123`).  These are consumed in graph order, which is the module
    execution order the plan builders follow, and every shape is checked;
  * un-fused exports (Conv followed by BatchNormalization with named parameters) fold through the normal `Weights.conv_bn`.
"""
from __future__ import annotations

import hashlib
import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import plan


# ---------------------------------------------------------------------------------------------------------------
# protobuf wire format (only what ONNX uses: varint, 64-bit, length-delimited, 32-bit)
# ---------------------------------------------------------------------------------------------------------------
def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) for one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _sint64(v: int) -> int:          # int64 fields are plain two's-complement varints
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_ints(wt: int, v) -> List[int]:
    if wt == 0:
        return [_sint64(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_sint64(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """TensorProto: dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9, double_data=10."""
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    external = False
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_ints(wt, v)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats += list(np.frombuffer(v, "<f4")) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno in (5, 7):
            ints += _packed_ints(wt, v)
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            floats += list(np.frombuffer(v, "<f8")) if wt == 2 else [struct.unpack("<d", v)[0]]
        elif fno == 13 or (fno == 14 and v == 1):
            external = True
    if external:
        raise Exception(f"initializer {name}: external tensor data is not supported (re-export with a single .onnx file)")
    if dtype not in _DTYPES:
        raise Exception(f"initializer {name}: unsupported ONNX data type {dtype}")
    np_t = _DTYPES[dtype]
    if raw is not None:
        a = np.frombuffer(raw, dtype=np.dtype(np_t).newbyteorder("<")).astype(np_t)
    elif floats:
        a = np.asarray(floats, dtype=np_t)
    elif dtype == 10 and ints:           # fp16 stored as uint16 bit patterns in int32_data
        a = np.asarray(ints, dtype=np.uint16).view(np.float16)
    else:
        a = np.asarray(ints, dtype=np_t)
    return name, a.reshape(dims) if dims else a.reshape(())


@dataclass
class OnnxNode:
    op_type: str
    name: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)


def _attribute(buf: memoryview):
    """AttributeProto: name=1, f=2, i=3, s=4, t=5, floats=7, ints=8."""
    name, val = "", None
    floats: List[float] = []
    ints: List[int] = []
    has_list = False
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _sint64(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            has_list = True
            floats += list(np.frombuffer(v, "<f4")) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            has_list = True
            ints += _packed_ints(wt, v)
    if has_list:
        val = ints if ints else floats
    return name, val


def _node(buf: memoryview) -> OnnxNode:
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5."""
    n = OnnxNode("", "", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op_type = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(v)
            n.attrs[k] = a
    return n


def _value_info(buf: memoryview) -> Tuple[str, List[Optional[int]]]:
    """ValueInfoProto: name=1, type=2 -> TypeProto.tensor_type=1 -> shape=2 -> dim=1 -> dim_value=1 / dim_param=2."""
    name, shape = "", []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 != 2:
                        continue
                    for f4, _, v4 in _fields(v3):
                        if f4 != 1:
                            continue
                        d = None
                        for f5, w5, v5 in _fields(v4):
                            if f5 == 1:
                                d = _sint64(v5)
                        shape.append(d)
    return name, shape


@dataclass
class OnnxModel:
    nodes: List[OnnxNode]
    initializers: Dict[str, np.ndarray]
    inputs: List[Tuple[str, List[Optional[int]]]]
    outputs: List[Tuple[str, List[Optional[int]]]]
    opset: int
    producer: str


def read_onnx(path: str) -> OnnxModel:
    """ModelProto: producer_name=2, graph=7, opset_import=8;  GraphProto: node=1, initializer=5, input=11, output=12."""
    if not os.path.isfile(path):
        raise Exception("The model path [%s] can't not found!" % path)
    with open(path, "rb") as f:
        data = memoryview(f.read())
    try:
        return _read_model(data, path)
    except (IndexError, ValueError, struct.error, UnicodeDecodeError) as e:
        raise Exception("The model path [%s] is not a readable ONNX file (%s: %s)" % (path, type(e).__name__, e)) from e


def _read_model(data: memoryview, path: str) -> OnnxModel:
    graph, opset, producer = None, 0, ""
    for fno, wt, v in _fields(data):
        if fno == 7:
            graph = v
        elif fno == 2:
            producer = bytes(v).decode()
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = bytes(v2).decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = max(opset, ver)
    if graph is None:
        raise Exception("The model path [%s] is not an ONNX ModelProto (no graph)" % path)
    m = OnnxModel([], {}, [], [], opset, producer)
    for fno, wt, v in _fields(graph):
        if fno == 1:
            m.nodes.append(_node(v))
        elif fno == 5:
            name, a = _tensor(v)
            m.initializers[name] = a
        elif fno == 11:
            m.inputs.append(_value_info(v))
        elif fno == 12:
            m.outputs.append(_value_info(v))
    for n in m.nodes:                   # Constant nodes are parameters too (some exporters emit weights this way); identical
        if n.op_type == "Constant" and "value" in n.attrs and n.outputs:      # initializers are stored once and aliased by Identity
            m.initializers.setdefault(n.outputs[0], np.asarray(n.attrs["value"]))
        elif n.op_type == "Identity" and n.inputs and n.inputs[0] in m.initializers and n.outputs:
            m.initializers.setdefault(n.outputs[0], m.initializers[n.inputs[0]])
    m.inputs = [(k, s) for k, s in m.inputs if k not in m.initializers]      # old exporters list initializers as inputs
    return m


# ---------------------------------------------------------------------------------------------------------------
# parameters -> plan.Weights
# ---------------------------------------------------------------------------------------------------------------
class OnnxWeights(plan.Weights):
    """`plan.Weights` whose parameters come from an ONNX graph (see the module docstring for the matching rules)."""

    def __init__(self, model: OnnxModel):
        named = {k: v for k, v in model.initializers.items() if v.dtype.kind == "f" and v.ndim >= 1}
        super().__init__({k: v.astype(np.float32) for k, v in named.items()})
        # convolutions in graph order: (weight name, weight, bias or None)
        self.convs: List[Tuple[str, np.ndarray, Optional[np.ndarray]]] = []
        for n in model.nodes:
            if n.op_type != "Conv" or len(n.inputs) < 2 or n.inputs[1] not in model.initializers:
                continue
            w = model.initializers[n.inputs[1]].astype(np.float32)
            b = model.initializers[n.inputs[2]].astype(np.float32) if len(n.inputs) > 2 and n.inputs[2] in model.initializers else None
            self.convs.append((n.inputs[1], w, b))
        # the bias a named convolution actually uses is the one its node references (exporters share identical initializers, so
        # `x.bias` may be stored once under another module's name)
        self._node_bias = {name: b for name, _, b in self.convs}
        self._anon = [c for c in self.convs if not _is_module_name(c[0])]
        self._anon_used = set()
        self.used_anonymous = 0

    def conv_bn(self, prefix: str, cout: int, cin: int, k: int, eps: float, conv_key="conv", bn_key="bn", res_branch=False):
        wkey = f"{prefix}.{conv_key}.weight" if conv_key else f"{prefix}.weight"
        bkey = wkey[:-len("weight")] + "bias"
        sd = self.state_dict
        if wkey in sd and f"{prefix}.{bn_key}.running_var" in sd:                # un-fused export: fold here
            return super().conv_bn(prefix, cout, cin, k, eps, conv_key, bn_key, res_branch)
        if wkey in sd:                                                            # fused before export, names kept
            w = sd[wkey]
            assert tuple(w.shape) == (cout, cin, k, k), f"{wkey}: expected {(cout, cin, k, k)}, file has {tuple(w.shape)}"
            b = self._node_bias.get(wkey)
            if b is None:
                b = sd[bkey] if bkey in sd else np.zeros(cout, np.float32)
            return w.astype(np.float32), b.astype(np.float32)
        # BN folded by the exporter: anonymous tensors, consumed in graph (= execution) order.  A residual block's 1x1 shortcut may
        # be traced before or after its two 3x3 convolutions (torchvision runs it after bn2), so the first unconsumed tensor of the
        # expected shape among the next three is taken.
        pending = [i for i in range(len(self._anon)) if i not in self._anon_used][:3]
        if not pending:
            raise Exception(f"ONNX file has no parameters left for {prefix} (architecture mismatch?)")
        for i in pending:
            name, w, b = self._anon[i]
            if tuple(w.shape) == (cout, cin, k, k):
                self._anon_used.add(i)
                self.used_anonymous += 1
                return w, (b if b is not None else np.zeros(cout, np.float32))
        name, w, _ = self._anon[pending[0]]
        raise Exception(f"{prefix}: expected a {(cout, cin, k, k)} convolution, the next unnamed Conv in the file ({name}) is {tuple(w.shape)}; "
                        "export with the module names kept (fuse Conv+BN in PyTorch before torch.onnx.export, as ultralytics does)")

    def conv_bias(self, prefix: str, cout: int, cin: int, k: int):
        return self.conv_bn(prefix, cout, cin, k, 0.0, conv_key="", bn_key="__no_bn__")


def _is_module_name(name: str) -> bool:
    """True for exporter-kept parameter names (`model.0.conv.weight`, `pool.weight`), False for `onnx::Conv_123` and friends."""
    return name.endswith(".weight") and "::" not in name and not name.split(".")[0].isdigit()


# ---------------------------------------------------------------------------------------------------------------
# architecture recognition
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    kind: str                 # "yolov8" | "yolov5" | "ufldv2"
    scale: str                # YOLO scale letter or ResNet depth ("18" / "34")
    nc: int = 80
    in_h: int = 640
    in_w: int = 640


_V8_WIDTH = {16: "n", 32: "s", 48: "m", 64: "l", 80: "x"}


def recognise(model: OnnxModel) -> ModelSpec:
    w = OnnxWeights(model)
    if not w.convs:
        raise Exception("no convolutions found in the ONNX graph")
    in_shape = model.inputs[0][1] if model.inputs else []
    in_h = int(in_shape[2]) if len(in_shape) == 4 and in_shape[2] else 0
    in_w = int(in_shape[3]) if len(in_shape) == 4 and in_shape[3] else 0
    first = w.convs[0][1]
    shapes = [tuple(c[1].shape) for c in w.convs]
    if first.shape[1:] == (3, 7, 7):                                          # torchvision ResNet stem -> UFLDv2
        n3 = sum(1 for s in shapes if s[2:] == (3, 3))
        depth = {16: "18", 32: "34"}.get(n3)
        if depth is None:
            raise Exception(f"UFLD backbone with {n3} 3x3 convolutions is not supported (ResNet-18/34 only)")
        return ModelSpec("ufldv2", depth, 0, in_h or 320, in_w or 1600)
    cout0, k0 = first.shape[0], first.shape[2]
    if cout0 not in _V8_WIDTH:
        raise Exception(f"unrecognised YOLO width: first convolution has {cout0} output channels")
    scale = _V8_WIDTH[cout0]
    named_nc = None                     # the class count is read from the head's own tensors when their names survived
    for name, cw, _ in w.convs:
        if re.fullmatch(r"model\.\d+\.cv3\.\d+\.2\.weight", name):      # YOLOv8 Detect.cv3[i][2]: Conv2d(c3, nc, 1)
            named_nc = cw.shape[0]
        elif re.fullmatch(r"model\.\d+\.m\.\d+\.weight", name) and cw.shape[2:] == (1, 1) and cw.shape[0] % 3 == 0:
            named_nc = cw.shape[0] // 3 - 5                                     # YOLOv5 Detect.m[i]: Conv2d(c, 3 * (nc + 5), 1)
    if k0 == 6:                                                               # yolov5 v6.x stem Conv(3, c, 6, 2, 2)
        if named_nc is not None:
            return ModelSpec("yolov5", scale, named_nc, in_h or 640, in_w or 640)
        no = [s[0] for s in shapes if s[2:] == (1, 1)][-1]                    # Detect.m[i]: 3 * (nc + 5)
        assert no % 3 == 0, f"YOLOv5 head with {no} outputs"
        return ModelSpec("yolov5", scale, no // 3 - 5, in_h or 640, in_w or 640)
    if k0 == 3:
        if named_nc is not None:
            return ModelSpec("yolov8", scale, named_nc, in_h or 640, in_w or 640)
        # Detect.cv3[i][2]: Conv2d(c3, nc, 1) -- the last 1x1 convolutions before the (optional) fixed DFL conv
        ones = [s for s in shapes if s[2:] == (1, 1) and s[0] != 1]
        return ModelSpec("yolov8", scale, ones[-1][0], in_h or 640, in_w or 640)
    raise Exception(f"unrecognised first convolution {first.shape}")


def build_plan(model: OnnxModel, spec: Optional[ModelSpec] = None) -> "plan.PlanBuilder":
    spec = spec or recognise(model)
    w = OnnxWeights(model)
    if spec.kind == "yolov8":
        return plan.build_yolov8(w, spec.scale, nc=spec.nc, in_h=spec.in_h, in_w=spec.in_w)
    if spec.kind == "yolov5":
        return plan.build_yolov5(w, spec.scale, nc=spec.nc, in_h=spec.in_h, in_w=spec.in_w)
    if spec.kind == "ufldv2":
        # the dataset follows from the input binding (ModelConfig: CULane 320x1600, TuSimple 320x800); the engine rejects any other
        cfg = dict(plan.UFLD_TUSIMPLE if (spec.in_h, spec.in_w) == (320, 800) else plan.UFLD_CULANE)
        cfg["in_h"], cfg["in_w"] = spec.in_h, spec.in_w
        return plan.build_ufldv2(w, spec.scale, cfg)
    raise Exception(f"unsupported model kind {spec.kind}")


def plan_from_onnx(onnx_path: str, out_path: Optional[str] = None) -> str:
    """Convert once and cache: returns the path of the `.b200w` plan for `onnx_path` (the counterpart of the reference's
    convertOnnxToTensorRT.py, which writes a `.trt` next to the `.onnx`)."""
    if not os.path.isfile(onnx_path):
        raise Exception("The model path [%s] can't not found!" % onnx_path)
    st = os.stat(onnx_path)
    if out_path is None:
        tag = hashlib.sha1(f"{os.path.abspath(onnx_path)}:{st.st_size}:{st.st_mtime_ns}:{plan.PLAN_VERSION}".encode()).hexdigest()[:16]
        cache = plan.cache_dir()
        out_path = os.path.join(cache, f"{os.path.splitext(os.path.basename(onnx_path))[0]}-{tag}.b200w")
    if os.path.isfile(out_path) and os.path.getmtime(out_path) >= st.st_mtime:
        return out_path
    pb = build_plan(read_onnx(onnx_path))
    tmp = out_path + f".tmp{os.getpid()}"
    pb.write(tmp)
    os.replace(tmp, out_path)
    return out_path
