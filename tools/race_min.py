import os, sys
os.environ["LNB_ENGINE"] = "1"; os.environ["LNB_P2P_TIMEOUT_MS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lnb_b200 as L
from tests.helpers import host_tensors
mode = sys.argv[1] if len(sys.argv) > 1 else "strict"
args = dict(L.synth.TINY)
gm = L.model.LoadModelFromTensors(args, host_tensors(args, 7))
acc = L._capi.LNB_ACC_STRICT if mode == "strict" else L._capi.LNB_ACC_FAST
ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(24), max_rows=8, acc_mode=acc)
first, _ = gm.Transformer.forward_argmax(ctx, np.array([1, 50, 999], np.int32), 0)
toks, ms, _ = ctx.decode_run(first, 3, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
print(mode, first, list(toks))
ctx.close(); gm.Free()
