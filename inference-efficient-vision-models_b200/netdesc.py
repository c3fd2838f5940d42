"""Host-side graph flattening: from the artefacts the reference's quantization stage produces to the
flat ``ievm_net_desc`` the C ABI consumes (include/ievm.h).

Three producers are accepted, exactly the three the reference hands to ``model(images)``:

* ``from_converted(gm)`` -- the live FX-converted ``GraphModule`` returned by
  ``QuantizationEngine.static_quantize`` (quantization/engines.py:95-121) or the inline PTQ of
  quantization/main.py:185-242.  The FX graph is walked node by node, so any stage widths and either
  qconfig flavour work; ``quantized.add_relu`` nodes are folded into the conv that feeds them.
* ``from_quantized_state_dict(sd)`` -- the ``state_dict()`` the reference saves for the INT8 variant
  (quantization/main.py:306-308); topology is the torchvision BasicBlock ResNet naming.
* ``from_half_module(m)`` -- the ``.half()``-cast torchvision ResNet (quantization/engines.py:84-93,
  quantization/main.py:256-262) with BasicBlock (student) or Bottleneck (ResNet-50 teacher,
  knowledge_distillation/utils.py:28-38) blocks; eval-mode BatchNorm is folded into the conv in fp32.

For calibration on the GPU a fourth walker, ``from_prepared(prepared)``, flattens the
observer-instrumented float ``GraphModule`` that ``prepare_fx`` returns (quantization/engines.py:109,
quantization/main.py:232) into an FP16 network whose residual adds stay separate ``OP_ADD_RELU`` layers, together
with the observer plan: which observer module sees which tensor, in call order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

OP_CONV, OP_MAXPOOL, OP_HEAD, OP_ADD_RELU = 0, 1, 2, 3
DTYPE_I8, DTYPE_F16 = 0, 1


@dataclass
class LayerSpec:
    op: int
    name: str
    in_tensor: int
    out_tensor: int = -1
    res_tensor: int = -1
    cin: int = 0
    cout: int = 0
    ksize: int = 1
    stride: int = 1
    pad: int = 0
    relu: bool = False
    weight: Optional[np.ndarray] = None     # int8 or float16, [cout, cin, k, k] ([classes, cin] for HEAD)
    bias: Optional[np.ndarray] = None       # float32 [cout]
    w_scale: Optional[np.ndarray] = None    # float32 [cout] (I8)
    in_scale: float = 1.0
    in_zp: int = 0
    out_scale: float = 1.0
    out_zp: int = 0
    res_scale: float = 1.0
    res_zp: int = 0
    add_scale: float = 1.0
    add_zp: int = 0


@dataclass
class NetSpec:
    dtype: int
    in_c: int = 3
    in_h: int = 224
    in_w: int = 224
    num_classes: int = 6
    in_scale: float = 1.0
    in_zp: int = 0
    layers: List[LayerSpec] = field(default_factory=list)
    tensor_names: Dict[int, str] = field(default_factory=dict)   # tensor id -> reference graph node name

    # ---- engine blob ("an export of our packed/padded weight blob for fast start") -------------
    _SCALARS = ("op", "name", "in_tensor", "out_tensor", "res_tensor", "cin", "cout", "ksize", "stride", "pad", "relu",
                "in_scale", "in_zp", "out_scale", "out_zp", "res_scale", "res_zp", "add_scale", "add_zp")

    def save(self, path) -> None:
        """Write the flattened network as one ``.npz`` (no pickle): a JSON header with every scalar field plus the
        weight / bias / scale arrays.  ``NetSpec.load`` + the engine constructor then start without torch.ao, FX or
        the reference's module classes.  float32 scalars are stored as their exact decimal repr (round-trip safe)."""
        import json
        head = {"format": "ievm-netspec-1", "dtype": self.dtype, "in_c": self.in_c, "in_h": self.in_h, "in_w": self.in_w,
                "num_classes": self.num_classes, "in_scale": repr(float(self.in_scale)), "in_zp": int(self.in_zp),
                "tensor_names": {str(k): v for k, v in self.tensor_names.items()}, "layers": []}
        arrays = {}
        for i, L in enumerate(self.layers):
            row = {}
            for f in self._SCALARS:
                v = getattr(L, f)
                row[f] = repr(float(v)) if isinstance(v, float) else (bool(v) if isinstance(v, (bool, np.bool_)) else
                                                                       (v if isinstance(v, str) else int(v)))
            head["layers"].append(row)
            for f in ("weight", "bias", "w_scale"):
                a = getattr(L, f)
                if a is not None:
                    arrays[f"L{i}.{f}"] = np.ascontiguousarray(a)
        arrays["header"] = np.frombuffer(json.dumps(head).encode(), dtype=np.uint8)
        with open(path, "wb") as fh:
            np.savez(fh, **arrays)

    @classmethod
    def load(cls, path) -> "NetSpec":
        import json
        with np.load(path, allow_pickle=False) as z:
            head = json.loads(bytes(z["header"]).decode())
            if head.get("format") != "ievm-netspec-1":
                raise ValueError(f"{path}: not an ievm NetSpec blob")
            net = cls(dtype=int(head["dtype"]), in_c=int(head["in_c"]), in_h=int(head["in_h"]), in_w=int(head["in_w"]),
                      num_classes=int(head["num_classes"]), in_scale=float(head["in_scale"]), in_zp=int(head["in_zp"]),
                      tensor_names={int(k): v for k, v in head["tensor_names"].items()})
            for i, row in enumerate(head["layers"]):
                kw = {f: (float(v) if f.endswith("_scale") else v) for f, v in row.items()}
                for f in ("weight", "bias", "w_scale"):
                    key = f"L{i}.{f}"
                    kw[f] = np.array(z[key]) if key in z.files else None
                net.layers.append(LayerSpec(**kw))
        _check_order(net)
        return net

    def conv_macs_per_image(self) -> int:
        """Algorithmic MACs per image on real (un-padded) channels."""
        h, w = self.in_h, self.in_w
        shape = {0: (h, w)}
        total = 0
        for L in self.layers:
            ih, iw = shape[L.in_tensor]
            if L.op == OP_CONV:
                oh = (ih + 2 * L.pad - L.ksize) // L.stride + 1
                ow = (iw + 2 * L.pad - L.ksize) // L.stride + 1
                total += oh * ow * L.cout * L.cin * L.ksize * L.ksize
                shape[L.out_tensor] = (oh, ow)
            elif L.op == OP_MAXPOOL:
                shape[L.out_tensor] = ((ih - 1) // 2 + 1, (iw - 1) // 2 + 1)
            elif L.op == OP_ADD_RELU:
                shape[L.out_tensor] = (ih, iw)
            else:
                total += L.cin * L.cout
        return total


# --------------------------------------------------------------------------------------------- INT8

def _qconv_fields(mod):
    w = mod.weight()
    if w.qscheme() not in (torch.per_channel_affine, torch.per_channel_symmetric):
        raise ValueError("expected per-channel quantized weights")
    if int(w.q_per_channel_zero_points().abs().max()) != 0:
        raise ValueError("expected symmetric (zero-point 0) weights")
    b = mod.bias()
    cout = w.shape[0]
    return dict(
        weight=np.ascontiguousarray(w.int_repr().numpy()),
        w_scale=w.q_per_channel_scales().to(torch.float32).numpy().copy(),
        bias=(b.detach().float().numpy().copy() if b is not None else np.zeros(cout, np.float32)),
        out_scale=float(mod.scale), out_zp=int(mod.zero_point))


def from_converted(gm, in_hw=(224, 224)) -> NetSpec:
    """Flatten an FX-converted quantized ResNet-style GraphModule (quantization/engines.py:118)."""
    import torch.ao.nn.intrinsic.quantized as nniq
    import torch.ao.nn.quantized as nnq

    mods = dict(gm.named_modules())
    net = NetSpec(dtype=DTYPE_I8, in_h=in_hw[0], in_w=in_hw[1])
    tid: Dict[str, int] = {}        # fx node name -> tensor id
    qp: Dict[int, tuple] = {}       # tensor id -> (scale, zp)
    ch: Dict[int, int] = {}
    producer: Dict[int, LayerSpec] = {}
    next_id = 1
    pending_pool = None

    def attr(node):
        return getattr(gm, node.target)

    for node in gm.graph.nodes:
        if node.op in ("placeholder", "get_attr", "output"):
            continue
        if node.op == "call_function" and node.target is torch.quantize_per_tensor:
            net.in_scale, net.in_zp = float(attr(node.args[1])), int(attr(node.args[2]))
            tid[node.name] = 0
            qp[0] = (net.in_scale, net.in_zp)
            ch[0] = net.in_c
            net.tensor_names[0] = node.name
            continue
        if node.op == "call_module":
            mod = mods[node.target]
            src = tid[node.args[0].name]
            if isinstance(mod, (nnq.Conv2d, nniq.ConvReLU2d)):
                f = _qconv_fields(mod)
                cout, cin, kh, kw = f["weight"].shape
                if kh != kw or mod.stride[0] != mod.stride[1] or mod.padding[0] != mod.padding[1] or mod.groups != 1 \
                        or tuple(mod.dilation) != (1, 1):
                    raise ValueError(f"{node.target}: unsupported conv geometry")
                L = LayerSpec(op=OP_CONV, name=node.target, in_tensor=src, out_tensor=next_id, cin=cin, cout=cout,
                              ksize=kh, stride=int(mod.stride[0]), pad=int(mod.padding[0]),
                              relu=isinstance(mod, nniq.ConvReLU2d), in_scale=qp[src][0], in_zp=qp[src][1], **f)
                net.layers.append(L)
                tid[node.name] = next_id
                qp[next_id] = (L.out_scale, L.out_zp)
                ch[next_id] = cout
                producer[next_id] = L
                net.tensor_names[next_id] = node.target
                next_id += 1
            elif isinstance(mod, torch.nn.MaxPool2d):
                k = mod.kernel_size if isinstance(mod.kernel_size, int) else mod.kernel_size[0]
                s = mod.stride if isinstance(mod.stride, int) else mod.stride[0]
                p = mod.padding if isinstance(mod.padding, int) else mod.padding[0]
                if (k, s, p) != (3, 2, 1):
                    raise ValueError("only MaxPool2d(3, 2, 1) is supported")
                net.layers.append(LayerSpec(op=OP_MAXPOOL, name=node.target, in_tensor=src, out_tensor=next_id,
                                            cin=ch[src], cout=ch[src]))
                tid[node.name] = next_id
                qp[next_id] = qp[src]
                ch[next_id] = ch[src]
                net.tensor_names[next_id] = node.target
                next_id += 1
            elif isinstance(mod, torch.nn.AdaptiveAvgPool2d):
                pending_pool = src
                tid[node.name] = src
            elif isinstance(mod, nnq.Linear):
                if pending_pool is None:
                    raise ValueError("Linear without a preceding AdaptiveAvgPool2d")
                w, b = mod._packed_params._weight_bias()
                net.num_classes = int(w.shape[0])
                net.layers.append(LayerSpec(
                    op=OP_HEAD, name=node.target, in_tensor=pending_pool, cin=int(w.shape[1]), cout=int(w.shape[0]),
                    weight=np.ascontiguousarray(w.int_repr().numpy()),
                    w_scale=w.q_per_channel_scales().to(torch.float32).numpy().copy(),
                    bias=b.detach().float().numpy().copy(), in_scale=qp[pending_pool][0], in_zp=qp[pending_pool][1],
                    out_scale=float(mod.scale), out_zp=int(mod.zero_point)))
                tid[node.name] = -1
            else:
                raise ValueError(f"unsupported module {type(mod).__name__} at {node.target}")
            continue
        if node.op == "call_function" and node.target is torch.ops.quantized.add_relu:
            a, b = tid[node.args[0].name], tid[node.args[1].name]
            scale, zp = float(attr(node.args[2])), int(attr(node.args[3]))

            # Fold the add into the epilogue of a conv operand (a + b is commutative in IEEE float, so
            # either operand gives identical results); prefer the first argument (the block's main path).
            def ok(n):
                L = producer.get(tid[n.name])
                return L is not None and not L.relu and L.res_tensor < 0 and len(n.users) == 1
            if ok(node.args[0]):
                t, other = a, b
            elif ok(node.args[1]):
                t, other = b, a
            else:
                raise ValueError(f"{node.name}: no single-use conv operand to fuse the add into")
            L = producer[t]
            net.layers.remove(L)        # run it after the other operand's producer (e.g. the downsample conv)
            net.layers.append(L)
            L.res_tensor = other
            L.res_scale, L.res_zp = qp[other]
            L.add_scale, L.add_zp = scale, zp
            tid[node.name] = t
            qp[t] = (scale, zp)
            net.tensor_names[t] = node.name
            continue
        if node.op == "call_function" and node.target is torch.flatten:
            tid[node.name] = tid[node.args[0].name]
            continue
        if node.op == "call_method" and node.target == "dequantize":
            continue
        raise ValueError(f"unsupported graph node {node.op}:{node.target}")
    _check_order(net)
    return net


def _check_order(net: NetSpec) -> None:
    seen = {0}
    for L in net.layers:
        if L.in_tensor not in seen or (L.res_tensor >= 0 and L.res_tensor not in seen):
            raise ValueError(f"layer {L.name} consumes a tensor that is produced later")
        if L.op != OP_HEAD:
            seen.add(L.out_tensor)


def from_quantized_state_dict(sd, in_hw=(224, 224)) -> NetSpec:
    """Rebuild the INT8 net from the state_dict saved at quantization/main.py:306-308
    (torchvision BasicBlock ResNet naming; strides follow the torchvision topology)."""
    net = NetSpec(dtype=DTYPE_I8, in_h=in_hw[0], in_w=in_hw[1])
    net.in_scale = float(sd["conv1_input_scale_0"])
    net.in_zp = int(sd["conv1_input_zero_point_0"])

    def conv(name, src, src_qp, stride, pad, relu, out_id):
        w = sd[name + ".weight"]
        b = sd.get(name + ".bias")
        cout, cin, k, _ = w.shape
        return LayerSpec(op=OP_CONV, name=name, in_tensor=src, out_tensor=out_id, cin=cin, cout=cout, ksize=k,
                         stride=stride, pad=pad, relu=relu,
                         weight=np.ascontiguousarray(w.int_repr().numpy()),
                         w_scale=w.q_per_channel_scales().to(torch.float32).numpy().copy(),
                         bias=(b.detach().float().numpy().copy() if b is not None else np.zeros(cout, np.float32)),
                         in_scale=src_qp[0], in_zp=src_qp[1],
                         out_scale=float(sd[name + ".scale"]), out_zp=int(sd[name + ".zero_point"]))

    nid = 1
    stem = conv("conv1", 0, (net.in_scale, net.in_zp), 2, 3, True, nid)
    net.layers.append(stem)
    net.tensor_names[0] = "quantize_per_tensor"
    net.tensor_names[nid] = "conv1"
    cur, cur_qp = nid, (stem.out_scale, stem.out_zp)
    nid += 1
    net.layers.append(LayerSpec(op=OP_MAXPOOL, name="maxpool", in_tensor=cur, out_tensor=nid, cin=stem.cout, cout=stem.cout))
    net.tensor_names[nid] = "maxpool"
    cur = nid
    nid += 1
    add_idx = 0
    li = 1
    while f"layer{li}.0.conv1.weight" in sd:
        bi = 0
        while f"layer{li}.{bi}.conv1.weight" in sd:
            p = f"layer{li}.{bi}"
            has_down = f"{p}.downsample.0.weight" in sd
            stride = 2 if (bi == 0 and li > 1) else 1
            c1 = conv(f"{p}.conv1", cur, cur_qp, stride, 1, True, nid)
            net.layers.append(c1)
            net.tensor_names[nid] = c1.name
            t1 = nid
            t2 = nid + 1            # same numbering as the FX walk: conv1, conv2, downsample
            nid += 2
            res, res_qp = cur, cur_qp
            if has_down:
                ds = conv(f"{p}.downsample.0", cur, cur_qp, stride, 0, False, nid)
                net.layers.append(ds)
                net.tensor_names[nid] = ds.name
                res, res_qp = nid, (ds.out_scale, ds.out_zp)
                nid += 1
            c2 = conv(f"{p}.conv2", t1, (c1.out_scale, c1.out_zp), 1, 1, False, t2)
            c2.res_tensor = res
            c2.res_scale, c2.res_zp = res_qp
            c2.add_scale = float(sd[f"layer{li}_{bi}_relu_scale_0"])
            c2.add_zp = int(sd[f"layer{li}_{bi}_relu_zero_point_0"])
            net.layers.append(c2)
            net.tensor_names[t2] = "add_relu" if add_idx == 0 else f"add_relu_{add_idx}"
            add_idx += 1
            cur, cur_qp = t2, (c2.add_scale, c2.add_zp)
            bi += 1
        li += 1
    w, b = sd["fc._packed_params._packed_params"]
    net.num_classes = int(w.shape[0])
    net.layers.append(LayerSpec(op=OP_HEAD, name="fc", in_tensor=cur, cin=int(w.shape[1]), cout=int(w.shape[0]),
                                weight=np.ascontiguousarray(w.int_repr().numpy()),
                                w_scale=w.q_per_channel_scales().to(torch.float32).numpy().copy(),
                                bias=b.detach().float().numpy().copy(), in_scale=cur_qp[0], in_zp=cur_qp[1],
                                out_scale=float(sd["fc.scale"]), out_zp=int(sd["fc.zero_point"])))
    return net


# --------------------------------------------------------------------------------------------- FP16

def _fold_bn(conv, bn):
    """Fold eval-mode BatchNorm into the conv in fp32 (from the parameters as stored, i.e. already
    rounded to fp16 by ``.half()``); weights are rounded to fp16 once, the bias stays fp32."""
    w = conv.weight.detach().float()
    gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
    mean, var = bn.running_mean.detach().float(), bn.running_var.detach().float()
    s = gamma / torch.sqrt(var + bn.eps)
    wf = (w * s.view(-1, 1, 1, 1)).half()
    bias = beta - mean * s
    if conv.bias is not None:
        bias = bias + conv.bias.detach().float() * s
    return np.ascontiguousarray(wf.numpy()), bias.numpy().astype(np.float32)


def from_half_module(model, in_hw=(224, 224)) -> NetSpec:
    """Flatten a (``.half()``-cast or fp32) torchvision ResNet with BasicBlock / Bottleneck blocks."""
    from torchvision.models.resnet import BasicBlock, Bottleneck

    net = NetSpec(dtype=DTYPE_F16, in_h=in_hw[0], in_w=in_hw[1])
    nid = 1

    def conv(name, c, bn, src, relu):
        nonlocal nid
        w, b = _fold_bn(c, bn)
        if c.kernel_size[0] != c.kernel_size[1] or c.groups != 1:
            raise ValueError(f"{name}: unsupported conv geometry")
        L = LayerSpec(op=OP_CONV, name=name, in_tensor=src, out_tensor=nid, cin=c.in_channels, cout=c.out_channels,
                      ksize=c.kernel_size[0], stride=c.stride[0], pad=c.padding[0], relu=relu, weight=w, bias=b)
        net.layers.append(L)
        net.tensor_names[nid] = name
        nid += 1
        return L

    stem = conv("conv1", model.conv1, model.bn1, 0, True)
    net.tensor_names[0] = "x"
    cur = stem.out_tensor
    net.layers.append(LayerSpec(op=OP_MAXPOOL, name="maxpool", in_tensor=cur, out_tensor=nid, cin=stem.cout, cout=stem.cout))
    net.tensor_names[nid] = "maxpool"
    cur = nid
    nid += 1
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(model, f"layer{li}")):
            p = f"layer{li}.{bi}"
            res = cur
            if blk.downsample is not None:
                res = conv(f"{p}.downsample", blk.downsample[0], blk.downsample[1], cur, False).out_tensor
            if isinstance(blk, BasicBlock):
                t = conv(f"{p}.conv1", blk.conv1, blk.bn1, cur, True).out_tensor
                last = conv(f"{p}.conv2", blk.conv2, blk.bn2, t, True)
            elif isinstance(blk, Bottleneck):
                t = conv(f"{p}.conv1", blk.conv1, blk.bn1, cur, True).out_tensor
                t = conv(f"{p}.conv2", blk.conv2, blk.bn2, t, True).out_tensor
                last = conv(f"{p}.conv3", blk.conv3, blk.bn3, t, True)
            else:
                raise ValueError(f"unsupported block type {type(blk).__name__}")
            last.res_tensor = res
            cur = last.out_tensor
    fc = model.fc
    net.num_classes = fc.out_features
    net.layers.append(LayerSpec(op=OP_HEAD, name="fc", in_tensor=cur, cin=fc.in_features, cout=fc.out_features,
                                weight=np.ascontiguousarray(fc.weight.detach().half().numpy()),
                                bias=fc.bias.detach().float().numpy().copy()))
    return net


# --------------------------------------------------------------------------------- calibration (FP16)

POINT_POOLED, POINT_LOGITS = "pooled", "logits"      # observation points that are not workspace tensors


def from_prepared(prepared, in_hw=(224, 224)):
    """Flatten the observer-instrumented float module returned by ``prepare_fx`` (quantization/engines.py:109,
    quantization/main.py:232: conv+BN(+ReLU) already fused, one observer call after every quantizable node).

    Returns ``(net, plan)``: ``net`` is an FP16 ``NetSpec`` in which every ``add -> relu`` pair is its own
    ``OP_ADD_RELU`` layer (so the conv output in front of the add exists as a tensor), and ``plan`` lists the observer
    calls of one forward in graph order as ``(observer module name, point)`` with ``point`` a tensor id,
    ``POINT_POOLED`` (the avgpool / flatten output) or ``POINT_LOGITS``.  Observer modules that appear under several
    names (prepare_fx shares one instance between a max-pool / avg-pool / flatten and its input) are listed once per
    call, as the reference's calibration forward calls them."""
    import operator

    import torch.ao.nn.intrinsic as nni
    from torch.ao.quantization.observer import ObserverBase

    net = NetSpec(dtype=DTYPE_F16, in_h=in_hw[0], in_w=in_hw[1])
    tid: Dict[str, object] = {}      # fx node name -> tensor id | ("add", a, b) | (POINT_POOLED, src) | POINT_LOGITS
    ch: Dict[int, int] = {}
    plan = []
    nid = 1

    def tensor_of(node):
        t = tid[node.name]
        if not isinstance(t, int):
            raise ValueError(f"{node.name}: expected a feature-map tensor, got {t!r}")
        return t

    def conv_layer(name, conv, src, relu):
        nonlocal nid
        if conv.kernel_size[0] != conv.kernel_size[1] or conv.groups != 1 or tuple(conv.dilation) != (1, 1) \
                or conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1]:
            raise ValueError(f"{name}: unsupported conv geometry")
        bias = conv.bias.detach().float().numpy().copy() if conv.bias is not None else np.zeros(conv.out_channels, np.float32)
        L = LayerSpec(op=OP_CONV, name=name, in_tensor=src, out_tensor=nid, cin=conv.in_channels, cout=conv.out_channels,
                      ksize=conv.kernel_size[0], stride=conv.stride[0], pad=conv.padding[0], relu=relu,
                      weight=np.ascontiguousarray(conv.weight.detach().half().numpy()), bias=bias)
        net.layers.append(L)
        net.tensor_names[nid] = name
        ch[nid] = conv.out_channels
        nid += 1
        return L.out_tensor

    def add_relu_layer(name, pending):
        nonlocal nid
        _, a, b = pending
        net.layers.append(LayerSpec(op=OP_ADD_RELU, name=name, in_tensor=a, res_tensor=b, out_tensor=nid, cin=ch[a], cout=ch[a],
                                    relu=True))
        net.tensor_names[nid] = name
        ch[nid] = ch[a]
        nid += 1
        return nid - 1

    for node in prepared.graph.nodes:
        if node.op == "placeholder":
            tid[node.name] = 0
            ch[0] = net.in_c
            net.tensor_names[0] = node.name
        elif node.op == "output":
            pass
        elif node.op == "call_module":
            mod = prepared.get_submodule(node.target)     # (named_modules() lists a shared observer once)
            src = tid[node.args[0].name]
            if isinstance(mod, ObserverBase):
                point = src if isinstance(src, int) else (src[0] if isinstance(src, tuple) else src)
                if point == "add":
                    raise ValueError(f"{node.target}: an observer on a bare add (no ReLU behind it) is not supported")
                plan.append((node.target, point))
                tid[node.name] = src
            elif isinstance(mod, nni.ConvReLU2d):
                tid[node.name] = conv_layer(node.target, mod[0], tensor_of(node.args[0]), True)
            elif isinstance(mod, torch.nn.Conv2d):
                tid[node.name] = conv_layer(node.target, mod, tensor_of(node.args[0]), False)
            elif isinstance(mod, torch.nn.MaxPool2d):
                k = mod.kernel_size if isinstance(mod.kernel_size, int) else mod.kernel_size[0]
                st = mod.stride if isinstance(mod.stride, int) else mod.stride[0]
                pd = mod.padding if isinstance(mod.padding, int) else mod.padding[0]
                if (k, st, pd) != (3, 2, 1):
                    raise ValueError("only MaxPool2d(3, 2, 1) is supported")
                t = tensor_of(node.args[0])
                net.layers.append(LayerSpec(op=OP_MAXPOOL, name=node.target, in_tensor=t, out_tensor=nid, cin=ch[t], cout=ch[t]))
                net.tensor_names[nid] = node.target
                ch[nid] = ch[t]
                tid[node.name] = nid
                nid += 1
            elif isinstance(mod, torch.nn.ReLU):
                if not (isinstance(src, tuple) and src[0] == "add"):
                    raise ValueError(f"{node.target}: a ReLU that is neither fused into a conv nor behind an add")
                tid[node.name] = add_relu_layer(node.name, src)
            elif isinstance(mod, torch.nn.AdaptiveAvgPool2d):
                tid[node.name] = (POINT_POOLED, tensor_of(node.args[0]))
            elif isinstance(mod, torch.nn.Linear):
                if not (isinstance(src, tuple) and src[0] == POINT_POOLED):
                    raise ValueError("Linear without a preceding AdaptiveAvgPool2d")
                net.num_classes = mod.out_features
                net.layers.append(LayerSpec(op=OP_HEAD, name=node.target, in_tensor=src[1], cin=mod.in_features,
                                            cout=mod.out_features,
                                            weight=np.ascontiguousarray(mod.weight.detach().half().numpy()),
                                            bias=mod.bias.detach().float().numpy().copy()))
                tid[node.name] = POINT_LOGITS
            else:
                raise ValueError(f"unsupported module {type(mod).__name__} at {node.target} (was the model in eval() mode "
                                 "when prepare_fx fused conv + BN?)")
        elif node.op == "call_function" and node.target in (operator.add, operator.iadd, torch.add):
            tid[node.name] = ("add", tensor_of(node.args[0]), tensor_of(node.args[1]))
        elif node.op == "call_function" and node.target in (torch.relu, torch.nn.functional.relu):
            src = tid[node.args[0].name]
            if not (isinstance(src, tuple) and src[0] == "add"):
                raise ValueError(f"{node.name}: a ReLU that is neither fused into a conv nor behind an add")
            tid[node.name] = add_relu_layer(node.name, src)
        elif node.op == "call_function" and node.target is torch.flatten:
            tid[node.name] = tid[node.args[0].name]
        else:
            raise ValueError(f"unsupported graph node {node.op}:{node.target}")
    if not net.layers or net.layers[-1].op != OP_HEAD:
        raise ValueError("the prepared graph does not end in AdaptiveAvgPool2d + Linear")
    _check_order(net)
    return net, plan
