"""Profiling driver: synthetic 8B model, one prefill + N decode steps through the C-ABI without
CUDA graphs (so ncu sees every kernel).  Usage: python tools/prof_decode.py [fast|strict] [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lnb_b200 as L

mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
acc = L._capi.LNB_ACC_FAST if mode == "fast" else L._capi.LNB_ACC_STRICT
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc)
first, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.PROMPT_8, np.int32), 0)
toks, ms, g = ctx.decode_run(first, 8, n, use_graph=False)
print(mode, "tokens", first, list(toks), "ms/token", ms / n)
ctx.close(); m.Free()
