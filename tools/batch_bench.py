"""Config 5 of BASELINE.json on the GPUs at hand: 8 concurrent prompts decoded together (one pass over the
weights per step).  Prints aggregate decode tokens/s.  Usage: python tools/batch_bench.py [fast|strict] [n_seq=8] [steps=32]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import lnb_b200 as L

mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
n_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 32
acc = L._capi.LNB_ACC_FAST if mode == "fast" else L._capi.LNB_ACC_STRICT
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc, n_seq=n_seq)
cur, pos = [], []
for b in range(n_seq):
    ctx.set_active_sequence(b)
    nxt, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.batch_prompt(b), np.int32), 0)
    cur.append(nxt); pos.append(8)
for _ in range(3):  # warm-up steps (positions advance like in a real run)
    nxt, _ = ctx.forward_batch(cur, pos)
    cur = list(nxt); pos = [p + 1 for p in pos]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    nxt, _ = ctx.forward_batch(cur, pos)
    cur = list(nxt); pos = [p + 1 for p in pos]
dt = time.perf_counter() - t0
print(f"batch={n_seq} {mode}: {dt / steps * 1e3:.3f} ms per step -> {n_seq * steps / dt:.1f} tokens/s aggregate "
      f"(host-driven steps, {n_seq} tokens back per step); last tokens {cur}")
ctx.close(); m.Free()
