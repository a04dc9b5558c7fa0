"""Host mirror of the reference's src/ml package: the Tensor type and the ops that sit on the
forward path, with the same exported names, argument meaning and error behaviour
(Go `(nil, error)` becomes a raised MlError).  Arithmetic is done by liblnb.so's CUDA kernels
through the op-level C-ABI -- exactly what the cgo shims of INTEGRATION.md bind.

Reference: src/ml/tensor.go:11-57, datatype.go:11-34, operations_impl.go, activations.go.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import LNB_ACC_FAST, LNB_ACC_STRICT, check, lib, ptr


class MlError(ValueError):
    pass


@dataclass(frozen=True)
class DataType:
    Name: str
    ItemSize: int
    np_dtype: object

    def __str__(self):
        return self.Name


DT_BF16 = DataType("BF16", 2, np.uint16)       # datatype.go:11-34
DT_F32 = DataType("Float32", 4, np.float32)
DT_INT32 = DataType("Int32", 4, np.int32)
DT_COMPLEX = DataType("Complex", 8, np.complex64)

# accumulation order used by the shims (reference order by default)
ACC_MODE = LNB_ACC_STRICT


def f32_to_bf16_bits(x) -> np.ndarray:
    """dtype.Float32ToBFloat16bits: truncation (src/dtype/bfloat16.go:59-61)"""
    x = np.ascontiguousarray(x, np.float32)
    return (x.view(np.uint32) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b) -> np.ndarray:
    """dtype.BFloat16bitsToFloat32 (src/dtype/bfloat16.go:55-57)"""
    b = np.ascontiguousarray(b, np.uint16)
    return (b.astype(np.uint32) << 16).view(np.float32)


class Tensor:
    """ml.Tensor (tensor.go:11-19): Size + DataType + contiguous row-major RawData."""

    def __init__(self, data: np.ndarray, dtype: DataType, name: str = ""):
        if data.dtype != dtype.np_dtype:
            raise MlError(f"raw data dtype {data.dtype} does not match {dtype}")
        self.RawData = np.ascontiguousarray(data)
        self.DataType = dtype
        self.Name = name

    @property
    def Size(self):
        return list(self.RawData.shape)

    @classmethod
    def from_f32(cls, values, dtype: DataType = DT_BF16, name: str = "") -> "Tensor":
        v = np.asarray(values, np.float32)
        if dtype is DT_BF16:
            return cls(f32_to_bf16_bits(v).reshape(v.shape), DT_BF16, name)
        if dtype is DT_F32:
            return cls(v.copy(), DT_F32, name)
        raise MlError(f"unsupported tensor datatype {dtype}")

    def to_f32_array(self) -> np.ndarray:
        if self.DataType is DT_BF16:
            return bf16_bits_to_f32(self.RawData).reshape(self.RawData.shape)
        return self.RawData.astype(np.float32)

    def GetElementCount(self) -> int:
        return int(self.RawData.size)

    def Reshape(self, size) -> "Tensor":  # tensor.go:386-397
        if int(np.prod(size)) != self.RawData.size:
            raise MlError(f"shape {self.Size} cannot be reshaped to {list(size)}")
        return Tensor(self.RawData.reshape(size), self.DataType, self.Name)

    def Slice(self, loc_start, loc_end) -> "Tensor":
        """tensor.go:266-343: [start, end) per given leading dimension; a leading dimension with
        start == end selects that single index and is dropped from the result; returns a copy."""
        if len(loc_start) != len(loc_end):
            raise MlError(f"locStart {len(loc_start)} and locEnd {len(loc_end)} don't have same dimensions")
        if len(loc_start) == 0 or len(loc_start) > self.RawData.ndim:
            raise MlError(f"locStart {len(loc_start)} and tensor \"{self.Name}\" {self.RawData.ndim} don't have "
                          "compatible dimensions")
        idx, leading = [], True
        for d, (a, b) in enumerate(zip(loc_start, loc_end)):
            n = self.RawData.shape[d]
            if a < 0 or a > n or b < 0 or b > n or b - a < 0:
                raise MlError("incompatible locStart, locEnd values and tensor")
            if leading and a == b and d < self.RawData.ndim - 1:
                idx.append(a)
            else:
                leading = False
                idx.append(slice(a, b))
        return Tensor(np.ascontiguousarray(self.RawData[tuple(idx)]), self.DataType, self.Name)

    def SetSlice(self, loc_start, loc_end, val: "Tensor") -> None:
        """tensor.go:345-384: writes `val` over the region Slice(loc_start, loc_end) would return (same leading-index
        rule); sizes and data types must agree"""
        if val.DataType is not self.DataType:
            raise MlError("tensors are not in same data type")
        if len(loc_start) != len(loc_end) or len(loc_start) == 0 or len(loc_start) > self.RawData.ndim:
            raise MlError("incompatible locStart, locEnd values and tensor")
        idx, leading = [], True
        for d, (a, b) in enumerate(zip(loc_start, loc_end)):
            n = self.RawData.shape[d]
            if a < 0 or a > n or b < 0 or b > n or b - a < 0:
                raise MlError("incompatible locStart, locEnd values and tensor")
            if leading and a == b and d < self.RawData.ndim - 1:
                idx.append(a)
            else:
                leading = False
                idx.append(slice(a, b))
        region = self.RawData[tuple(idx)]
        if list(region.shape) != val.Size:
            raise MlError(f"incompatible sizes: region {list(region.shape)} and value {val.Size}")
        region[...] = val.RawData

    def Transpose(self, dim1: int, dim2: int) -> "Tensor":
        """tensor.go:542-598: swaps two dimensions and returns a contiguous copy"""
        nd = self.RawData.ndim
        if not (0 <= dim1 < nd and 0 <= dim2 < nd):
            raise MlError(f"dimensions {dim1}, {dim2} out of range for a {nd}-D tensor")
        return Tensor(np.ascontiguousarray(np.swapaxes(self.RawData, dim1, dim2)), self.DataType, self.Name)

    def ToFloat32(self) -> "Tensor":  # tensor.go:430-454
        if self.DataType is DT_F32:
            return self
        if self.DataType is not DT_BF16:
            raise MlError(f"unsupported tensor datatype {self.DataType}")
        return Tensor(self.to_f32_array(), DT_F32, self.Name)

    def ToBFloat16(self) -> "Tensor":  # tensor.go:456-480 (truncating)
        if self.DataType is DT_BF16:
            return self
        if self.DataType is not DT_F32:
            raise MlError(f"unsupported tensor datatype {self.DataType}")
        return Tensor(f32_to_bf16_bits(self.RawData).reshape(self.RawData.shape), DT_BF16, self.Name)

    def Item(self):
        if self.RawData.size != 1:
            raise MlError("Item() needs a one-element tensor")
        return self.RawData.reshape(-1)[0].item()


def _same_dtype(a: Tensor, b: Tensor):
    if a.DataType is not b.DataType:  # checkSameDataType
        raise MlError(f"tensors are not in same data type: {a.DataType} and {b.DataType}")


def LinearTransformation(input: Tensor, weights: Tensor) -> Tensor:
    """ml.LinearTransformation (operations_impl.go:427-447): input [S,K] x weights [N,K]^T."""
    _same_dtype(input, weights)
    if input.RawData.ndim != 2 or weights.RawData.ndim != 2:
        raise MlError("LinearTransformation needs two matrices")
    S, K = input.RawData.shape
    N, Kw = weights.RawData.shape
    if K != Kw:
        raise MlError(f"columns size {K} of input tensor ({input.Size}) should be equal with {Kw} input features "
                      f"count of weights tensor ({weights.Size})")
    if input.DataType is not DT_BF16:
        raise MlError(f"unsupported tensor datatype {input.DataType}")
    out = np.empty((S, N), np.uint16)
    check(lib.lnb_op_linear_bf16(ptr(input.RawData, _capi.u16p), ptr(weights.RawData, _capi.u16p), ptr(out, _capi.u16p),
                                 S, K, N, ACC_MODE))
    return Tensor(out, DT_BF16)


def MatMul(input: Tensor, other: Tensor) -> Tensor:
    """ml.MatMul (operations_impl.go:449-476): [..,M,K] x [..,K,N]."""
    _same_dtype(input, other)
    a, b = input.RawData, other.RawData
    if a.ndim < 2 or b.ndim < 2:
        raise MlError("MatMul needs at least 2-d tensors")
    if a.shape[-1] != b.shape[-2]:
        raise MlError(f"columns size {a.shape[-1]} of input tensor ({input.Size}) should be equal with rows size "
                      f"{b.shape[-2]} of other tensor ({other.Size})")
    if a.shape[:-2] != b.shape[:-2]:
        raise MlError(f"first parts of dimensions are not compatible;  {list(a.shape[:-2])} of input tensor and "
                      f"{list(b.shape[:-2])} of other tensor")
    if input.DataType is not DT_BF16:
        raise MlError(f"unsupported tensor datatype {input.DataType}")
    B = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    M, K, N = a.shape[-2], a.shape[-1], b.shape[-1]
    out = np.empty(a.shape[:-2] + (M, N), np.uint16)
    check(lib.lnb_op_matmul_bf16(ptr(a, _capi.u16p), ptr(b, _capi.u16p), ptr(out, _capi.u16p), B, M, K, N))
    return Tensor(out, DT_BF16)


def _binary(a: Tensor, b: Tensor, fn) -> Tensor:
    _same_dtype(a, b)
    if a.DataType is not DT_BF16:
        raise MlError(f"unsupported tensor datatype {a.DataType}")
    if a.RawData.shape != b.RawData.shape:  # the forward path only uses same-shape Add / Multiply
        raise MlError(f"tensors are not broadcastable here: {a.Size} and {b.Size}")
    out = np.empty(a.RawData.shape, np.uint16)
    check(fn(ptr(a.RawData, _capi.u16p), ptr(b.RawData, _capi.u16p), ptr(out, _capi.u16p), a.RawData.size))
    return Tensor(out, DT_BF16)


def Add(input: Tensor, other: Tensor) -> Tensor:  # operations_impl.go:307-335
    return _binary(input, other, lib.lnb_op_add_bf16)


def MultiplyElementwise(input: Tensor, other: Tensor) -> Tensor:  # operations_impl.go:367-395
    return _binary(input, other, lib.lnb_op_mul_bf16)


def Silu(input: Tensor) -> Tensor:  # activations.go:27-50
    if input.DataType is not DT_BF16:
        raise MlError(f"unsupported tensor datatype {input.DataType}")
    out = np.empty(input.RawData.shape, np.uint16)
    check(lib.lnb_op_silu_bf16(ptr(input.RawData, _capi.u16p), ptr(out, _capi.u16p), input.RawData.size))
    return Tensor(out, DT_BF16)


def Softmax(input: Tensor, dim: int) -> Tensor:  # operations_impl.go:478-511
    if dim != input.RawData.ndim - 1:
        raise MlError("currenlty Softmax supports only last dimension of input tensor as dim argument")
    if input.DataType is not DT_F32:
        raise MlError(f"unsupported tensor datatype {input.DataType}")
    cols = input.RawData.shape[-1]
    rows = input.RawData.size // cols
    out = np.empty(input.RawData.shape, np.float32)
    check(lib.lnb_op_softmax_f32(ptr(input.RawData, _capi.f32p), ptr(out, _capi.f32p), rows, cols))
    return Tensor(out, DT_F32)


def Argmax(input: Tensor, dim: int) -> Tensor:  # operations_impl.go:513-548
    if dim != len(input.Size) - 1:
        raise MlError("currenlty Argmax supports only last dimension of input tensor as dim argument")
    if input.DataType is not DT_F32:
        raise MlError(f"unsupported tensor datatype {input.DataType}")
    if isinstance(input, DeviceLogits) and input.on_device():
        # the logits never left HBM: argmax them there (4 bytes come back), no re-upload of [rows, vocab] floats
        out = np.empty(input.rows, np.int32)
        check(lib.lnb_session_logits_argmax(input.ctx.h, input.generation, input.row0, input.rows, ptr(out, _capi.i32p)))
        return Tensor(out, DT_INT32)
    cols = input.RawData.shape[-1]
    rows = input.RawData.size // cols
    out = np.empty(input.RawData.shape[:-1], np.int32)
    check(lib.lnb_op_argmax_f32(ptr(input.RawData, _capi.f32p), rows, cols, ptr(out, _capi.i32p)))
    return Tensor(out, DT_INT32)


class DeviceLogits(Tensor):
    """The f32 logits tensor LlamaTransformer.Forward returns (llamatransformer.go:175), still in HBM: Size / DataType /
    Slice / ml.Argmax work on the handle (lnb_forward_device); the first access to RawData copies the rows to the host
    (lnb_session_logits_read) and from then on it is an ordinary host tensor.  The handle is valid until the
    InferenceContext's next Forward -- exactly the lifetime generateTokensInternal needs (inference.go:202-216)."""

    def __init__(self, ctx, generation: int, row0: int, rows: int, vocab: int, name: str = ""):
        self.ctx, self.generation, self.row0, self.rows, self.vocab = ctx, generation, row0, rows, vocab
        self._host = None
        self.DataType = DT_F32
        self.Name = name

    def on_device(self) -> bool:
        return self._host is None and self.ctx.h and self.ctx.generation == self.generation

    @property
    def RawData(self):
        if self._host is None:
            if not (self.ctx.h and self.ctx.generation == self.generation):
                raise MlError("logits tensor outlived the Forward call that produced it (its rows were never read)")
            out = np.empty((self.rows, self.vocab), np.float32)
            check(lib.lnb_session_logits_read(self.ctx.h, self.generation, self.row0, self.rows, ptr(out, _capi.f32p)))
            self._host = out
        return self._host

    @RawData.setter
    def RawData(self, v):
        self._host = v

    @property
    def Size(self):
        return [self.rows, self.vocab]

    def GetElementCount(self) -> int:
        return self.rows * self.vocab

    def Slice(self, loc_start, loc_end) -> "Tensor":
        # row ranges keep the handle (tensor.go:266-343 semantics: one leading dimension, start < end)
        if self._host is None and len(loc_start) == 1 and len(loc_end) == 1 and 0 <= loc_start[0] < loc_end[0] <= self.rows:
            return DeviceLogits(self.ctx, self.generation, self.row0 + loc_start[0], loc_end[0] - loc_start[0], self.vocab, self.Name)
        return Tensor(self.RawData, DT_F32, self.Name).Slice(loc_start, loc_end)


def Fwd_Get_Rows(embedding: Tensor, tokens: Tensor) -> Tensor:  # operations_impl.go:142-173
    if embedding.RawData.ndim != 2:
        raise MlError("embedding is not a matrix")
    if tokens.RawData.ndim != 1:
        raise MlError("tokens is not a vector")
    if tokens.DataType is not DT_INT32:
        raise MlError(f"tensor is not in data type {DT_INT32}: \"{tokens.Name}\" is {tokens.DataType}")
    S = tokens.RawData.shape[0]
    V, D = embedding.RawData.shape
    out = np.empty((S, D), np.uint16)
    check(lib.lnb_op_get_rows_bf16(ptr(embedding.RawData, _capi.u16p), ptr(tokens.RawData, _capi.i32p),
                                   ptr(out, _capi.u16p), S, V, D))
    return Tensor(out, DT_BF16)


def Full(size, dtype: DataType, fill_value) -> Tensor:  # operations_impl.go:55-64
    return Tensor(np.full(size, fill_value, dtype.np_dtype), dtype)


# ---- init-time builders (host-side in the reference as well: they only run inside precomputeFreqsCis / prepare,
# src/model/llamatransformer.go:115-143,694-751) ---------------------------------------------------------------

def _fill_value(dtype: DataType, v: float):
    if dtype is DT_BF16:
        return int(f32_to_bf16_bits(np.float32(v)).reshape(-1)[0])
    if dtype is DT_F32:
        return np.float32(v)
    raise MlError(f"unsupported tensor datatype {dtype}")


def Zeros(size, dtype: DataType) -> Tensor:  # operations_impl.go:66-77
    return Full(size, dtype, _fill_value(dtype, 0.0))


def Ones(size, dtype: DataType) -> Tensor:  # operations_impl.go:79-90
    return Full(size, dtype, _fill_value(dtype, 1.0))


def ZerosLike(input: Tensor) -> Tensor:  # :92-94
    return Zeros(input.Size, input.DataType)


def OnesLike(input: Tensor) -> Tensor:  # :96-98
    return Ones(input.Size, input.DataType)


def ARange(start: int, end: int, step: int, dtype: DataType) -> Tensor:  # operations_impl.go:11-24
    if start >= end:
        raise MlError(f"start value {start} must be less than end value {end} in ARange")
    vals = np.arange(start, end, step, dtype=np.int64).astype(np.float32)   # SetItem_FromFloat32(float32(val))
    if dtype is DT_BF16:
        return Tensor(f32_to_bf16_bits(vals), DT_BF16)
    if dtype is DT_F32:
        return Tensor(vals, DT_F32)
    raise MlError(f"unsupported tensor datatype {dtype}")


def Outer(vec1: Tensor, vec2: Tensor) -> Tensor:  # operations_impl.go:26-52
    if vec1.RawData.ndim != 1 or vec2.RawData.ndim != 1:
        raise MlError("tensor must be a vector")                     # checkIsVector
    _same_dtype(vec1, vec2)
    prod = (vec1.to_f32_array()[:, None] * vec2.to_f32_array()[None, :]).astype(np.float32)   # valF32 := row * col
    return Tensor.from_f32(prod, vec1.DataType)


def Polar(abs: Tensor, angle: Tensor) -> Tensor:  # operations_impl.go:100-140
    if abs.Size != angle.Size:
        raise MlError(f"tensors are not in same shape: {abs.Size} and {angle.Size}")
    _same_dtype(abs, angle)
    if abs.RawData.ndim != 2:
        raise MlError("tensor must be a matrix")                     # "Currently only 2D matrices are supported"
    a, t = abs.to_f32_array().astype(np.float64), angle.to_f32_array().astype(np.float64)
    return Tensor(((a * np.cos(t)) + 1j * (a * np.sin(t))).astype(np.complex64), DT_COMPLEX)


def TriangularUpper(input: Tensor, diagonal: int) -> Tensor:  # operations_impl.go:175-195
    if input.RawData.ndim != 2:
        raise MlError("tensor must be a matrix")
    i, j = np.indices(input.RawData.shape)
    return Tensor(np.where(j - i >= diagonal, input.RawData, np.zeros_like(input.RawData)), input.DataType)
