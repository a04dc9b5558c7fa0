// Package inference -- cgo shim for the tokenizer half of the engine: the vocabulary (tiktoken.Load +
// model.NewVocabulary's tables), the split regexp, the byte pair merge and the chat template run in liblnb.so
// (csrc/tokenizer.cpp); the method set of InferenceEngine is unchanged.
//
// Integration points in the reference (signatures unchanged):
//   (*InferenceEngine).TokenizeString      src/inference/tokenize.go:175-193 -> lnb_tokenize_string
//   (*InferenceEngine).Tokenize            src/inference/tokenize.go:27-95   -> lnb_tokenize_prompt
//   (*InferenceEngine).TokenBatchToString  src/inference/tokenize.go:239-258 -> lnb_detokenize (bytes; the emoji
//                                          annotation of src/inference/emoji.go stays in Go)
//   model.loadVocab                         src/model/loader.go:84-96         -> lnb_vocab_load (kept next to the Go
//                                          Vocabulary, which still serves IdToToken / StopTokenIds to the console UI)
// The generate loop (generateTokensInternal, inference.go:173-254) is untouched: it calls Transformer.Forward, whose
// body is the one C call of go/model/lnb_cgo.go.
// NOT compiled in this repository (no Go toolchain in the build image; see INTEGRATION.md).
package inference

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

	"github.com/adalkiran/llama-nuts-and-bolts/src/model"
)

type deviceVocab struct{ h *C.lnb_vocab }

func lnbErr(rc C.int) error {
	if rc >= 0 {
		return nil
	}
	return fmt.Errorf("%s", C.GoString(C.lnb_last_error()))
}

func loadDeviceVocab(vocabFilePath string) (*deviceVocab, error) {
	cpath := C.CString(vocabFilePath)
	defer C.free(unsafe.Pointer(cpath))
	dv := &deviceVocab{}
	return dv, lnbErr(C.lnb_vocab_load(cpath, &dv.h))
}

// body of (*InferenceEngine).TokenizeString
func (dv *deviceVocab) tokenizeString(text string) ([]model.TokenId, error) {
	out := make([]model.TokenId, len(text)+8) // a piece never yields more tokens than it has bytes
	var n C.int
	ctext := C.CString(text)
	defer C.free(unsafe.Pointer(ctext))
	rc := C.lnb_tokenize_string(dv.h, ctext, C.int64_t(len(text)), (*C.int32_t)(unsafe.Pointer(&out[0])), C.int(len(out)), &n)
	if err := lnbErr(rc); err != nil {
		return nil, err
	}
	return out[:int(n)], nil
}

// body of (*InferenceEngine).Tokenize
func (dv *deviceVocab) tokenize(promptParts []PromptPart) ([]model.TokenId, error) {
	k := len(promptParts)
	headers := make([]*C.char, k+1)
	contents := make([]*C.char, k+1)
	capacity := 32
	for i, p := range promptParts {
		headers[i] = C.CString(p.Header)
		contents[i] = C.CString(p.Content)
		defer C.free(unsafe.Pointer(headers[i]))
		defer C.free(unsafe.Pointer(contents[i]))
		capacity += len(p.Header) + len(p.Content) + 16
	}
	out := make([]model.TokenId, capacity)
	var n C.int
	rc := C.lnb_tokenize_prompt(dv.h, &headers[0], &contents[0], C.int(k), (*C.int32_t)(unsafe.Pointer(&out[0])), C.int(capacity), &n)
	if err := lnbErr(rc); err != nil {
		return nil, err
	}
	return out[:int(n)], nil
}
