// Package torch -- cgo shim for the checkpoint reader: the zip / pickle parsing and the file mapping are done
// by liblnb.so (csrc/pth.cpp); Load keeps the reference's signature and returns tensors whose RawData aliases
// the library's read-only mapping, exactly like the reference's own mmap slices.
//
// Integration points in the reference (signatures unchanged):
//   NewTorchModelReader   src/torch/torchmodelreader.go:21-33   -> lnb_pth_open
//   (*TorchModelReader).Load   :39-66                            -> lnb_pth_tensor_count / _info / _data
//   (*TorchModelReader).Close  :35-37                            -> lnb_pth_close (after the device upload)
// With the model-level boundary in place the tensors need not cross into Go at all:
//   model.LoadModelEx  src/model/loader.go:22-41  can call C.lnb_model_load_pth(dm.h, path, &n) instead, which
//   copies every tensor NewLlamaTransformer asks for from the mapping straight into HBM.
// NOT compiled in this repository (no Go toolchain in the build image; see INTEGRATION.md).
package torch

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../llama-nuts-and-bolts_b200 -llnb -lcudart -ldl
#include <stdlib.h>
#include "lnb.h"
*/
import "C"

import (
	"fmt"
	"unsafe"

	"github.com/adalkiran/llama-nuts-and-bolts/src/ml"
	"github.com/adalkiran/llama-nuts-and-bolts/src/pickle"
)

type TorchModelReader struct {
	modelFilePath string
	h             *C.lnb_pth
}

func lnbErr(rc C.int) error {
	if rc >= 0 {
		return nil
	}
	return fmt.Errorf("%s", C.GoString(C.lnb_last_error()))
}

func NewTorchModelReader(modelFilePath string) (*TorchModelReader, error) {
	cpath := C.CString(modelFilePath)
	defer C.free(unsafe.Pointer(cpath))
	r := &TorchModelReader{modelFilePath: modelFilePath}
	if err := lnbErr(C.lnb_pth_open(cpath, &r.h)); err != nil {
		return nil, err
	}
	return r, nil
}

func (tmr *TorchModelReader) Close() error {
	rc := C.lnb_pth_close(tmr.h)
	tmr.h = nil
	return lnbErr(rc)
}

func (tmr *TorchModelReader) Load() (*pickle.PickleDict[*ml.Tensor], error) {
	modelTensors := pickle.NewPickleDict[*ml.Tensor]()
	n := int(C.lnb_pth_tensor_count(tmr.h))
	for i := 0; i < n; i++ {
		var name *C.char
		var dtype, ndim C.int
		var shape [8]C.int64_t
		var off, nbytes C.int64_t
		rc := C.lnb_pth_tensor_info(tmr.h, C.int(i), &name, &dtype, &ndim, &shape[0], &off, &nbytes)
		if err := lnbErr(rc); err != nil {
			return nil, err
		}
		if dtype != C.LNB_PTH_BF16 { // the reference knows torch.BFloat16Storage only (src/torch/types.go:9-21)
			return nil, fmt.Errorf("tensor \"%s\": unsupported storage type", C.GoString(name))
		}
		size := make([]int, int(ndim))
		for d := range size {
			size[d] = int(shape[d])
		}
		// C memory owned by the library's mapping (not a Go pointer): legal to keep until Close
		raw := unsafe.Slice((*byte)(C.lnb_pth_tensor_data(tmr.h, C.int(i))), int(nbytes))
		t := ml.NewTensor(C.GoString(name), size, nil, ml.DT_BF16, raw)
		modelTensors.Set(t.Name, t)
	}
	return modelTensors, nil
}
