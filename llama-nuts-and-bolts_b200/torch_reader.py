"""Host mirror of src/torch: TorchModelReader over the native .pth reader / writer of liblnb.so.

The parsing (zip central directory, pickle, storage lookup) and the mapping live in C++ (csrc/pth.cpp);
this module only shapes the result like the reference's API: `TorchModelReader(path).Load()` returns an
ordered name -> tensor mapping whose arrays alias the read-only mmap (src/torch/torchmodelreader.go:39-66,
src/torch/types.go:51-56).  PyTorch itself is NOT used here -- tests use torch.save / torch.load only as an
independent second opinion on the file format."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib

DT_NAMES = ("bfloat16", "float16", "float32", "float64", "int8", "uint8", "int16", "int32", "int64", "bool")
# numpy views of the raw bytes; bf16 has no numpy dtype -> uint16 bit patterns, like ml.Tensor's BF16
_NP = (np.uint16, np.float16, np.float32, np.float64, np.int8, np.uint8, np.int16, np.int32, np.int64, np.bool_)
PTH_BF16, PTH_F16, PTH_F32 = 0, 1, 2


class TorchTensor:
    """one entry of the checkpoint: name, dtype, Size, and RawData aliasing the file mapping"""

    def __init__(self, name, dtype, size, file_offset, nbytes, contiguous, data):
        self.Name, self.DataType, self.Size = name, DT_NAMES[dtype], tuple(size)
        self.dtype_code, self.file_offset, self.nbytes, self.contiguous = dtype, file_offset, nbytes, contiguous
        self.RawData = data        # np.ndarray over the mmap (read-only), shaped like Size when contiguous

    def __repr__(self):
        return f"TorchTensor({self.Name!r}, {self.DataType}, {list(self.Size)})"


class _Mapping:
    """owns the lnb_pth handle (file mapping); released when the reader AND every array handed out are gone --
    the reference never unmaps at all (src/common/memorymapper_unix.go:47-60)"""

    def __init__(self, path: str):
        self.h = C.c_void_p()
        check(lib.lnb_pth_open(path.encode(), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib.lnb_pth_close(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


class TorchModelReader:
    """torch.NewTorchModelReader(modelFilePath) / .Load() / .Close()"""

    def __init__(self, modelFilePath: str):
        self.modelFilePath = modelFilePath
        self._map = _Mapping(modelFilePath)

    def Load(self) -> dict[str, TorchTensor]:
        if self._map is None:
            raise _capi.LnbError(-4, "reader is closed")
        h = self._map.h
        out: dict[str, TorchTensor] = {}
        n = check(lib.lnb_pth_tensor_count(h))
        name, dt, nd = C.c_char_p(), C.c_int(), C.c_int()
        shape = (C.c_int64 * 8)()
        off, nb = C.c_int64(), C.c_int64()
        for i in range(n):
            rc = check(lib.lnb_pth_tensor_info(h, i, C.byref(name), C.byref(dt), C.byref(nd), shape, C.byref(off), C.byref(nb)))
            size = [int(shape[k]) for k in range(nd.value)]
            if nb.value:
                buf = (C.c_uint8 * nb.value).from_address(lib.lnb_pth_tensor_data(h, i))
                buf._owner = self._map                     # the array keeps the mapping alive
                raw = np.frombuffer(buf, np.uint8)
            else:
                raw = np.zeros(0, np.uint8)
            raw.flags.writeable = False
            arr = raw.view(_NP[dt.value])
            if rc == 0:
                arr = arr.reshape(size)
            key = name.value.decode(errors="surrogateescape")
            out[key] = TorchTensor(key, dt.value, size, off.value, nb.value, rc == 0, arr)
        return out

    def Close(self):
        self._map = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Close()


class TorchModelWriter:
    """the missing half in the reference: writes a .pth that torch.load AND the reference's reader accept"""

    def __init__(self, path: str):
        self.h = C.c_void_p()
        check(lib.lnb_pth_writer_create(path.encode(), C.byref(self.h)))

    def Add(self, name: str, arr: np.ndarray, dtype_code: int | None = None):
        arr = np.asarray(arr)
        if not arr.flags["C_CONTIGUOUS"]:          # (np.ascontiguousarray would turn a 0-d array into 1-d)
            arr = arr.copy(order="C")
        if dtype_code is None:
            dtype_code = {np.dtype(np.uint16): PTH_BF16, np.dtype(np.float16): PTH_F16, np.dtype(np.float32): PTH_F32,
                          np.dtype(np.float64): 3, np.dtype(np.int8): 4, np.dtype(np.uint8): 5, np.dtype(np.int16): 6,
                          np.dtype(np.int32): 7, np.dtype(np.int64): 8, np.dtype(np.bool_): 9}[arr.dtype]
        shape = (C.c_int64 * max(1, arr.ndim))(*arr.shape)
        check(lib.lnb_pth_writer_add(self.h, name.encode(), dtype_code, arr.ctypes.data_as(C.c_void_p), shape, arr.ndim))

    def Finish(self):
        h, self.h = self.h, C.c_void_p()
        check(lib.lnb_pth_writer_finish(h))


def write_synthetic_checkpoint(path: str, args_c: "_capi.ModelArgsC", seed: int):
    """consolidated.00.pth with the bits of lnb_model_init_synthetic (the .pth writer)"""
    check(lib.lnb_pth_write_synthetic(path.encode(), C.byref(args_c), seed))
