"""decode tok/s of the persistent engine for a few L2-prefetch windows (LNB_ENGINE_PF_KB), both modes; one process per
setting because the window is read once per process.  Usage: python tools/engine_sweep.py [kb,kb,...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys; sys.path.insert(0, %r)
import numpy as np, lnb_b200 as L
m = L.model.LoadSyntheticModel(dict(L.synth.LLAMA31_8B))
out = []
for mode, acc in (("fast", L._capi.LNB_ACC_FAST), ("strict", L._capi.LNB_ACC_STRICT)):
    ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(136), max_rows=8, acc_mode=acc)
    first, _ = m.Transformer.forward_argmax(ctx, np.array(L.synth.PROMPT_8, np.int32), 0)
    ctx.decode_run(first, 8, 16)
    toks, ms, _ = ctx.decode_run(first, 8, 127)
    out.append("%%s %%.1f tok/s (%%.3f ms)" %% (mode, 127e3 / ms, ms / 127))
    ctx.close()
print(" | ".join(out))
""" % ROOT
for kb in (sys.argv[1] if len(sys.argv) > 1 else "0,128,256,512").split(","):
    env = dict(os.environ, LNB_ENGINE_PF_KB=kb)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    print(f"PF_KB={kb:>4s}: {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)
