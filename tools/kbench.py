"""Quick kernel-level timing on the synthetic 8B model: every projection kernel alone (CUDA
events, all layers back to back) and a 32-step graph decode, for both accumulation modes.
Usage: python tools/kbench.py [fast,strict]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lnb_b200 as L

modes = (sys.argv[1] if len(sys.argv) > 1 else "fast,strict").split(",")
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
names = {0: "wqkv", 1: "wo", 2: "w13", 3: "w2", 4: "lm_head"}
for mode in modes:
    acc = L._capi.LNB_ACC_FAST if mode == "fast" else L._capi.LNB_ACC_STRICT
    ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc)
    first, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.PROMPT_8, np.int32), 0)
    tot = 0.0
    out = []
    for k, nm in names.items():
        ms, nb, nl = ctx.bench_kernel(k, reps=3)
        out.append(f"{nm} {ms * 1e3:7.2f} us {nb / ms / 1e6:7.1f} GB/s")
        tot += ms * (1 if k == 4 else 32)
    ctx.decode_run(first, 8, 16, use_graph=True)
    toks, ms, g = ctx.decode_run(first, 8, 64, use_graph=True)
    print(f"[{mode}] " + " | ".join(out))
    print(f"[{mode}] sum of projection kernels/token {tot:.3f} ms; graph decode {ms / 64:.4f} ms/token = {64e3 / ms:.1f} tok/s "
          f"(graph={g}) tokens[:6]={list(toks[:6])}")
    ctx.close()
m.Free()
