// lnb_bench -- times the unmodified reference (adalkiran/llama-nuts-and-bolts) on a model directory written by
// `lnb_generate --write-synthetic`.  Copy to <reference>/cmd/lnb_bench/main.go and `go run ./cmd/lnb_bench <dir>`.
// Uses only the reference's exported API; not compiled in this repository (no Go toolchain in the image).
package main

import (
	"fmt"
	"os"
	"time"

	"github.com/adalkiran/llama-nuts-and-bolts/src/common"
	"github.com/adalkiran/llama-nuts-and-bolts/src/inference"
	"github.com/adalkiran/llama-nuts-and-bolts/src/model"
)

func main() {
	if len(os.Args) < 2 {
		fmt.Fprintln(os.Stderr, "usage: lnb_bench <modelDir> [sequenceLength=136]")
		os.Exit(2)
	}
	seqLen := 136
	if len(os.Args) > 2 {
		fmt.Sscanf(os.Args[2], "%d", &seqLen)
	}
	var err error
	if common.GLogger, err = common.NewLogger(os.Stdout, nil); err != nil {
		panic(err)
	}
	defer common.GLogger.Close()

	llamaModel, err := model.LoadModel(os.Args[1])
	if err != nil {
		panic(err)
	}
	defer llamaModel.Free()

	inferenceArgs := common.NewInferenceArgs()
	inferenceArgs.SequenceLength = seqLen
	engine := inference.NewInferenceEngine(llamaModel, inferenceArgs, func(format string, v ...any) {})

	// the fixed prompt of BASELINE.json's workload (SURVEY 8d); the first id is <|begin_of_text|>
	prompt := []model.TokenId{128000, 9906, 11, 856, 836, 374, 220, 16}
	if llamaModel.ModelArgs.VocabSize < 128256 { // the `tiny` architecture of lnb_generate --write-synthetic <dir> tiny
		prompt = []model.TokenId{1, 50, 999, 7, 300, 12, 64, 2}
	}
	start := time.Now()
	var first time.Time
	ids := make([]model.TokenId, 0, seqLen)
	partsCh, errCh := engine.GenerateString(prompt)
	for partsCh != nil || errCh != nil {
		select {
		case part, ok := <-partsCh:
			if !ok {
				partsCh = nil
				continue
			}
			if part.IsResendOfWaiting {
				continue
			}
			if len(ids) == 0 {
				first = time.Now()
			}
			ids = append(ids, part.TokenId)
		case err, ok := <-errCh:
			if !ok {
				errCh = nil
				continue
			}
			if err != nil {
				panic(err)
			}
		}
	}
	end := time.Now()
	fmt.Print("tokens:")
	for _, id := range ids {
		fmt.Printf(" %d", id)
	}
	fmt.Println()
	decode := end.Sub(first).Seconds()
	fmt.Printf("generated %d tokens; prefill+first token %.1f s; decode %.4f tokens/s (reference Go CPU path)\n",
		len(ids), first.Sub(start).Seconds(), float64(len(ids)-1)/decode)
}
