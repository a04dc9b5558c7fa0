// Package ml -- cgo shims that route the hot-path ops of src/ml to liblnb.so.
//
// Drop these files next to the reference's src/ml (same package) and delete the Go bodies of the
// functions they replace; every exported signature stays exactly as in the reference
// (src/ml/operations_impl.go:427,449,478,513; activations.go:27).  NOT compiled in this repository:
// the build image has no Go toolchain (see INTEGRATION.md).
package ml

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../llama-nuts-and-bolts_b200 -llnb -lcudart -ldl
#include "lnb.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// AccMode selects LNB_ACC_STRICT (reference order, bit-identical) or LNB_ACC_FAST.
var AccMode = C.int(C.LNB_ACC_STRICT)

func lnbErr(rc C.int) error {
	if rc >= 0 {
		return nil
	}
	return fmt.Errorf("lnb: %s", C.GoString(C.lnb_last_error()))
}

func u16(t *Tensor) *C.uint16_t { return (*C.uint16_t)(unsafe.Pointer(&t.RawData[0])) }
func f32(t *Tensor) *C.float    { return (*C.float)(unsafe.Pointer(&t.RawData[0])) }

// linearTransformation_BF16 replaces src/ml/operations_lineartransform.go:145-207.
func linearTransformation_BF16(input *Tensor, weights *Tensor) (*Tensor, error) {
	s, k, n := input.Size[0], input.Size[1], weights.Size[0]
	dst := NewEmptyTensor([]int{s, n}, DT_BF16)
	// cgo pointer rule: the library copies everything it needs before returning
	if err := lnbErr(C.lnb_op_linear_bf16(u16(input), u16(weights), u16(dst), C.int(s), C.int(k), C.int(n), AccMode)); err != nil {
		return nil, err
	}
	return dst, nil
}

// matMul_BF16 replaces src/ml/operations_matmul.go:136-182.
func matMul_BF16(input *Tensor, other *Tensor) (*Tensor, error) {
	nd := len(input.Size)
	m, k, n := input.Size[nd-2], input.Size[nd-1], other.Size[nd-1]
	b := 1
	for _, d := range input.Size[:nd-2] {
		b *= d
	}
	dstSize := append(append([]int{}, input.Size[:nd-2]...), m, n)
	dst := NewEmptyTensor(dstSize, DT_BF16)
	if err := lnbErr(C.lnb_op_matmul_bf16(u16(input), u16(other), u16(dst), C.int(b), C.int(m), C.int(k), C.int(n))); err != nil {
		return nil, err
	}
	return dst, nil
}

// Softmax replaces src/ml/operations_impl.go:478-511 (F32 rows, last dimension).
func softmaxF32(input *Tensor, dst *Tensor) error {
	cols := input.Size[len(input.Size)-1]
	rows := input.GetElementCount() / cols
	return lnbErr(C.lnb_op_softmax_f32(f32(input), f32(dst), C.int(rows), C.int(cols)))
}

// Argmax replaces src/ml/operations_impl.go:513-548.
func argmaxF32(input *Tensor, dst *Tensor) error {
	cols := input.Size[len(input.Size)-1]
	rows := input.GetElementCount() / cols
	return lnbErr(C.lnb_op_argmax_f32(f32(input), C.int(rows), C.int(cols), (*C.int32_t)(unsafe.Pointer(&dst.RawData[0]))))
}

// Silu / Add / MultiplyElementwise (same-shape BF16): activations.go:27-50, operations_impl.go:307-395.
func siluBF16(input *Tensor, dst *Tensor) error {
	return lnbErr(C.lnb_op_silu_bf16(u16(input), u16(dst), C.int64_t(input.GetElementCount())))
}
func addBF16(a, b, dst *Tensor) error {
	return lnbErr(C.lnb_op_add_bf16(u16(a), u16(b), u16(dst), C.int64_t(a.GetElementCount())))
}
func mulBF16(a, b, dst *Tensor) error {
	return lnbErr(C.lnb_op_mul_bf16(u16(a), u16(b), u16(dst), C.int64_t(a.GetElementCount())))
}
