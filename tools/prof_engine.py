"""ncu target: ONE launch of the persistent decode engine on the synthetic 8B model (n steps), after a prefill through the
kernel chain.  Usage: python tools/prof_engine.py [fast|strict] [n_steps]   (LNB_ENGINE=1 for fast)"""
import os
import sys

os.environ.setdefault("LNB_ENGINE", "1")
os.environ.setdefault("LNB_P2P_TIMEOUT_MS", "0")     # ncu replays the kernel with serialisation: no in-kernel deadlines
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lnb_b200 as L

mode = sys.argv[1] if len(sys.argv) > 1 else "strict"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
acc = L._capi.LNB_ACC_FAST if mode == "fast" else L._capi.LNB_ACC_STRICT
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc)
first, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.PROMPT_8, np.int32), 0)
toks, ms, g = ctx.decode_run(first, 8, n)
print(mode, "tokens", first, list(toks), "ms/token", ms / n)
ctx.close(); m.Free()
