"""Mini ONNX interpreter on torch-CPU (fp32) for the five shipped AirSLAM graphs.

TEST / TOOLING INFRASTRUCTURE ONLY.  It executes the reference's model files
(/root/reference/output/*.onnx -- the arithmetic spec of SURVEY.md §8a G1-G5)
node by node so that (a) tools/make_golden.py can freeze golden vectors under
tests/golden/ and (b) the hand-restated oracle in oracle/ can be pinned against
the graphs themselves.  It only runs where /root/reference exists (this
container), never on the GPU box, and never on the product path.

Covers the union of op types of the five graphs (59 types).  `emul` mode rounds
both operands of Conv/MatMul/Gemm/Einsum to fp16 and accumulates in fp32 (what a
tensor-core kernel computes up to summation order).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_reader  # noqa: F401  (package-relative when imported as tools.onnx_interp)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    x = np.asarray(x)
    if x.ndim == 0:
        return torch.tensor(x.item(), dtype=torch.from_numpy(x.reshape(1)).dtype)
    return torch.from_numpy(np.ascontiguousarray(x))


_ONNX2TORCH = {1: torch.float32, 2: torch.uint8, 3: torch.int8, 5: torch.int16, 6: torch.int32,
               7: torch.int64, 9: torch.bool, 10: torch.float16, 11: torch.float64}


class Interp:
    def __init__(self, graph, emul=None, record=None):
        """graph: onnx_reader.Graph; emul: None | 'fp16' | 'bf16'; record: set of tensor names to keep."""
        self.g = graph
        self.emul = emul
        self.record = record
        self.init = {k: _t(v) for k, v in graph.init.items()}

    def _r(self, x):
        if self.emul is None or not x.is_floating_point():
            return x
        dt = torch.float16 if self.emul == "fp16" else torch.bfloat16
        return x.to(dt).to(torch.float32)

    def run(self, feeds):
        env = dict(self.init)
        for k, v in feeds.items():
            env[k] = _t(v)
        kept = {}
        for n in self.g.nodes:
            ins = [env[i] if i != "" else None for i in n.inputs]
            fn = getattr(self, "op_" + n.op, None)
            if fn is None:
                raise NotImplementedError(n.op)
            outs = fn(n, *ins)
            if not isinstance(outs, (tuple, list)):
                outs = (outs,)
            for name, val in zip(n.outputs, outs):
                env[name] = val
                if self.record is not None and name in self.record:
                    kept[name] = val
        res = {o: env[o] for o in self.g.outputs}
        res.update(kept)
        return res

    # ---- elementwise -----------------------------------------------------------------
    def op_Abs(self, n, x): return x.abs()
    def op_Neg(self, n, x): return -x
    def op_Relu(self, n, x): return F.relu(x)
    def op_Sigmoid(self, n, x): return torch.sigmoid(x)
    def op_Sin(self, n, x): return torch.sin(x)
    def op_Cos(self, n, x): return torch.cos(x)
    def op_Tan(self, n, x): return torch.tan(x)
    def op_Erf(self, n, x): return torch.erf(x)
    def op_Log(self, n, x): return torch.log(x)
    def op_Sqrt(self, n, x): return torch.sqrt(x)
    def op_Floor(self, n, x): return torch.floor(x)
    def op_Identity(self, n, x): return x
    def op_Not(self, n, x): return ~x
    def op_And(self, n, a, b): return a & b
    def op_Or(self, n, a, b): return a | b
    def op_Equal(self, n, a, b): return a == b
    def op_Greater(self, n, a, b): return a > b
    def op_Less(self, n, a, b): return a < b
    def op_Add(self, n, a, b): return a + b
    def op_Sub(self, n, a, b): return a - b
    def op_Mul(self, n, a, b): return a * b
    def op_Min(self, n, *xs):
        r = xs[0]
        for x in xs[1:]:
            r = torch.minimum(r, x)
        return r
    def op_Max(self, n, *xs):
        r = xs[0]
        for x in xs[1:]:
            r = torch.maximum(r, x)
        return r

    def op_Div(self, n, a, b):
        if not a.is_floating_point() and not b.is_floating_point():
            return torch.div(a, b, rounding_mode="trunc")
        return a / b

    def op_Mod(self, n, a, b):
        if n.attrs.get("fmod", 0):
            return torch.fmod(a, b)
        return torch.remainder(a, b)

    def op_Pow(self, n, a, b):
        return torch.pow(a, b.to(a.dtype) if a.is_floating_point() else b)

    def op_Clip(self, n, x, lo=None, hi=None):
        if lo is None and "min" in n.attrs:
            lo = torch.tensor(n.attrs["min"])
        if hi is None and "max" in n.attrs:
            hi = torch.tensor(n.attrs["max"])
        if lo is not None:
            x = torch.maximum(x, lo.to(x.dtype))
        if hi is not None:
            x = torch.minimum(x, hi.to(x.dtype))
        return x

    def op_Where(self, n, c, a, b): return torch.where(c, a, b)

    def op_Cast(self, n, x):
        return x.to(_ONNX2TORCH[n.attrs["to"]])

    # ---- shape / movement ------------------------------------------------------------
    def op_Constant(self, n):
        if "value" in n.attrs:
            return _t(n.attrs["value"])
        if "value_float" in n.attrs:
            return torch.tensor(n.attrs["value_float"], dtype=torch.float32)
        if "value_int" in n.attrs:
            return torch.tensor(n.attrs["value_int"], dtype=torch.int64)
        if "value_ints" in n.attrs:
            return torch.tensor(n.attrs["value_ints"], dtype=torch.int64)
        raise NotImplementedError(n.attrs.keys())

    def op_ConstantOfShape(self, n, shape):
        v = n.attrs.get("value")
        val = _t(v).reshape(-1)[0] if v is not None else torch.tensor(0.0)
        return torch.full([int(s) for s in shape.tolist()], val.item(), dtype=val.dtype)

    def op_Shape(self, n, x):
        return torch.tensor(list(x.shape), dtype=torch.int64)

    def op_Reshape(self, n, x, shape):
        s = [int(v) for v in shape.tolist()]
        s = [x.shape[i] if v == 0 else v for i, v in enumerate(s)]
        return x.reshape(s)

    def op_Flatten(self, n, x):
        ax = n.attrs.get("axis", 1)
        lead = int(np.prod(x.shape[:ax])) if ax > 0 else 1
        return x.reshape(lead, -1)

    def op_Transpose(self, n, x):
        perm = n.attrs.get("perm")
        if perm is None:
            perm = list(range(x.dim()))[::-1]
        return x.permute(perm).contiguous()

    def _axes(self, n, axes_in):
        if axes_in is not None:
            return [int(a) for a in axes_in.tolist()]
        return list(n.attrs.get("axes", []))

    def op_Unsqueeze(self, n, x, axes=None):
        ax = self._axes(n, axes)
        nd = x.dim() + len(ax)
        ax = sorted(a % nd for a in ax)
        for a in ax:
            x = x.unsqueeze(a)
        return x

    def op_Squeeze(self, n, x, axes=None):
        ax = self._axes(n, axes)
        if not ax:
            return x.squeeze()
        for a in sorted((a % x.dim() for a in ax), reverse=True):
            x = x.squeeze(a)
        return x

    def op_Concat(self, n, *xs):
        return torch.cat(list(xs), dim=n.attrs["axis"])

    def op_Expand(self, n, x, shape):
        s = [int(v) for v in shape.tolist()]
        tgt = torch.broadcast_shapes(tuple(x.shape), tuple(s))
        return x.expand(tgt)

    def op_Slice(self, n, x, starts=None, ends=None, axes=None, steps=None):
        if starts is None:
            starts, ends = n.attrs["starts"], n.attrs["ends"]
            axes = n.attrs.get("axes")
        else:
            starts, ends = starts.tolist(), ends.tolist()
            axes = axes.tolist() if axes is not None else None
            steps = steps.tolist() if steps is not None else None
        if axes is None:
            axes = list(range(len(starts)))
        if steps is None:
            steps = [1] * len(starts)
        idx = [slice(None)] * x.dim()
        for s, e, a, st in zip(starts, ends, axes, steps):
            a = a % x.dim()
            d = x.shape[a]
            if st > 0:
                s = max(0, min(d, s + d if s < 0 else s))
                e = max(0, min(d, e + d if e < 0 else e))
                idx[a] = slice(s, e, st)
            else:
                # negative step: torch cannot slice backwards, so gather explicit indices
                s = max(-1, min(d - 1, s + d if s < 0 else s))
                e = -1 if e < -d else max(-1, min(d - 1, e + d if e < 0 else e))
                sel = torch.arange(s, e, st, dtype=torch.int64)
                x = torch.index_select(x, a, sel)
        return x[tuple(idx)]

    def op_Gather(self, n, x, idx):
        ax = n.attrs.get("axis", 0) % x.dim()
        idx = idx.to(torch.int64)
        idx = torch.where(idx < 0, idx + x.shape[ax], idx)
        if idx.dim() == 0:
            return x.select(ax, int(idx))
        out = torch.index_select(x, ax, idx.reshape(-1))
        return out.reshape(list(x.shape[:ax]) + list(idx.shape) + list(x.shape[ax + 1:]))

    def op_GatherElements(self, n, x, idx):
        ax = n.attrs.get("axis", 0)
        idx = idx.to(torch.int64)
        idx = torch.where(idx < 0, idx + x.shape[ax], idx)
        return torch.gather(x, ax, idx)

    def op_ScatterElements(self, n, x, idx, upd):
        ax = n.attrs.get("axis", 0)
        assert n.attrs.get("reduction", "none") == "none"
        out = x.clone()
        idx = idx.to(torch.int64)
        # sequential semantics (last writer wins) -- torch.scatter on CPU is sequential too, but be explicit
        if out.dim() == 1:
            o = out.numpy()
            o[idx.numpy()] = upd.numpy()  # numpy fancy assignment: last value wins for duplicates
            return torch.from_numpy(o)
        if out.dim() == 2 and ax == 0 and idx.shape[1] == 1:
            o = out.numpy()
            o[idx.numpy()[:, 0], 0] = upd.numpy()[:, 0]
            return torch.from_numpy(o)
        return out.scatter(ax, idx, upd)

    def op_Range(self, n, a, b, c):
        return torch.arange(a.item(), b.item(), c.item(), dtype=a.dtype)

    def op_Resize(self, n, x, roi=None, scales=None, sizes=None):
        assert n.attrs.get("mode", "nearest") == "nearest"
        assert n.attrs.get("coordinate_transformation_mode") == "asymmetric"
        assert n.attrs.get("nearest_mode", "round_prefer_floor") == "floor"
        sc = scales.tolist()
        assert sc[0] == 1 and sc[1] == 1 and sc[2] == 2 and sc[3] == 2
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    # ---- reductions ------------------------------------------------------------------
    def _red_axes(self, n, x, axes_in):
        ax = self._axes(n, axes_in)
        if not ax:
            ax = list(range(x.dim()))
        return [a % x.dim() for a in ax], bool(n.attrs.get("keepdims", 1))

    def op_ReduceSum(self, n, x, axes=None):
        ax, kd = self._red_axes(n, x, axes)
        return x.sum(dim=ax, keepdim=kd)

    def op_ReduceMin(self, n, x, axes=None):
        ax, kd = self._red_axes(n, x, axes)
        return x.amin(dim=ax, keepdim=kd)

    def op_ReduceL2(self, n, x, axes=None):
        ax, kd = self._red_axes(n, x, axes)
        return torch.sqrt((x * x).sum(dim=ax, keepdim=kd))

    def op_ReduceLogSumExp(self, n, x, axes=None):
        ax, kd = self._red_axes(n, x, axes)
        return torch.logsumexp(x, dim=ax, keepdim=kd)

    def op_ArgMin(self, n, x):
        ax = n.attrs.get("axis", 0)
        kd = bool(n.attrs.get("keepdims", 1))
        assert not n.attrs.get("select_last_index", 0)
        # first index on ties: numpy argmin guarantees first occurrence
        r = torch.from_numpy(np.argmin(x.numpy(), axis=ax))
        return r.unsqueeze(ax) if kd else r

    def op_TopK(self, n, x, k):
        ax = n.attrs.get("axis", -1) % x.dim()
        assert n.attrs.get("largest", 1) == 1 and ax == x.dim() - 1
        k = int(k.reshape(-1)[0])
        # value desc, index asc on ties: stable descending sort
        xs = x.reshape(-1, x.shape[-1]).numpy()
        order = np.argsort(-xs, axis=-1, kind="stable")[:, :k]
        vals = np.take_along_axis(xs, order, axis=-1)
        shp = list(x.shape[:-1]) + [k]
        return torch.from_numpy(vals).reshape(shp), torch.from_numpy(order.astype(np.int64)).reshape(shp)

    def op_Softmax(self, n, x):
        return torch.softmax(x, dim=n.attrs.get("axis", -1))

    def op_LogSoftmax(self, n, x):
        return torch.log_softmax(x, dim=n.attrs.get("axis", -1))

    def op_LayerNormalization(self, n, x, w, b=None):
        ax = n.attrs.get("axis", -1) % x.dim()
        return F.layer_norm(x, x.shape[ax:], w, b, eps=n.attrs.get("epsilon", 1e-5))

    # ---- contractions ----------------------------------------------------------------
    def op_Conv(self, n, x, w, b=None):
        pads = n.attrs.get("pads", [0] * (2 * (x.dim() - 2)))
        strides = n.attrs.get("strides", [1] * (x.dim() - 2))
        dil = n.attrs.get("dilations", [1] * (x.dim() - 2))
        grp = n.attrs.get("group", 1)
        nd = x.dim() - 2
        assert pads[:nd] == pads[nd:]
        x, w = self._r(x), self._r(w)
        if nd == 2:
            return F.conv2d(x, w, b, stride=strides, padding=pads[:2], dilation=dil, groups=grp)
        return F.conv1d(x, w, b, stride=strides, padding=pads[:1], dilation=dil, groups=grp)

    def op_MatMul(self, n, a, b):
        return torch.matmul(self._r(a), self._r(b))

    def op_Gemm(self, n, a, b, c=None):
        a, b = self._r(a), self._r(b)
        if n.attrs.get("transA", 0):
            a = a.t()
        if n.attrs.get("transB", 0):
            b = b.t()
        y = n.attrs.get("alpha", 1.0) * (a @ b)
        if c is not None:
            y = y + n.attrs.get("beta", 1.0) * c
        return y

    def op_Einsum(self, n, *xs):
        return torch.einsum(n.attrs["equation"], *[self._r(x) for x in xs])

    def op_MaxPool(self, n, x):
        k = n.attrs["kernel_shape"]
        s = n.attrs.get("strides", [1, 1])
        p = n.attrs.get("pads", [0, 0, 0, 0])
        assert p[:2] == p[2:] and not n.attrs.get("ceil_mode", 0)
        return F.max_pool2d(x, k, s, p[:2])


def run_model(path, feeds, emul=None, record=None):
    from . import onnx_reader as R
    g = R.load(path)
    with torch.no_grad():
        return Interp(g, emul=emul, record=record).run(feeds)
