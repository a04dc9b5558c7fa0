"""Where the batch engine spends its time (LNB_ENGINE_PROF=1): microseconds of consumer thread 0 per section and step, mean
and max over the CTAs.  Usage: python tools/batch_prof.py [fast,strict] [n_seq=8]"""
import os
import sys

os.environ["LNB_ENGINE_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lnb_b200 as L

modes = (sys.argv[1] if len(sys.argv) > 1 else "fast,strict").split(",")
n_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
MHZ = 1965.0
STEPS = 16
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
for mode in modes:
    acc = L._capi.LNB_ACC_FAST if mode == "fast" else L._capi.LNB_ACC_STRICT
    ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc, n_seq=n_seq)
    cur, pos = [], []
    for b in range(n_seq):
        ctx.set_active_sequence(b)
        nxt, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.batch_prompt(b), np.int32), 0)
        cur.append(nxt); pos.append(8)
    for _ in range(3):
        nxt, _ = ctx.forward_batch(cur, pos)
        cur = list(nxt); pos = [p + 1 for p in pos]
    ctx.engine_profile(batch=True)
    for _ in range(STEPS):
        nxt, _ = ctx.forward_batch(cur, pos)
        cur = list(nxt); pos = [p + 1 for p in pos]
    prof = ctx.engine_profile(batch=True)
    tiles = prof.pop("tiles")
    tot = sum(v[0] for v in prof.values())
    print(f"[{mode}] n_seq={n_seq}: {tot / STEPS / MHZ / 1e3:.3f} ms per step accounted, {tiles[0] / STEPS:.0f} tiles per CTA and step", flush=True)
    for k, (mean, mx) in prof.items():
        print(f"   {k:>20s}: mean {mean / STEPS / MHZ:9.1f} us   max {mx / STEPS / MHZ:9.1f} us   ({mean / max(tiles[0], 1):7.0f} cycles per tile)", flush=True)
    ctx.close()
m.Free()
