"""Writes tests/golden/oracle_regression.json: SHA-256 digests of what the CPU oracle computes for fixed seeded inputs.
The oracle is the checker of every GPU parity test; this anchor makes an accidental change of the oracle itself visible
(tests/test_oracle_golden.py::test_oracle_outputs_did_not_drift).  Run: python tests/golden/make_oracle_regression.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.helpers import host_tensors, oracle_model  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression.json")


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute() -> dict:
    import lnb_b200 as L
    args = dict(L.synth.TINY)
    tensors = host_tensors(args, 1234)
    om = oracle_model(args, tensors)
    prompt = np.array([3, 77, 1000, 5, 9, 512, 64, 1], np.int32)
    s = om.new_session(24)
    logits_prefill = s.forward(prompt, 0)
    logits_decode = s.forward(np.array([int(np.argmax(logits_prefill[-1]))], np.int32), len(prompt))
    k0, v0 = (np.array(c) for c in s.cache(0))          # copies: cache() returns views into the session
    tokens = om.generate(prompt, 24, stop_ids=(10**9,))
    s.close()
    om.close()
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((3, 4096)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
    w = (rng.standard_normal((32, 4096)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
    freqs, cis = O.rope_table(128, 256, 500000.0, True)
    return {
        "tiny_weights": digest(np.concatenate([t.reshape(-1) for t in tensors.values()])),
        "tiny_logits_prefill": digest(logits_prefill), "tiny_logits_decode": digest(logits_decode),
        "tiny_cache_k0": digest(k0), "tiny_cache_v0": digest(v0), "tiny_tokens": [int(t) for t in tokens],
        "rms_scale_3x4096": digest(O.rms_scale(x, 1e-5)), "rmsnorm_3x4096": digest(O.rmsnorm(x, w[0], 1e-5)),
        "linear_bf16_3x4096x32": digest(O.linear_bf16(x, w)),
        "rope_freqs_128_scaled": digest(freqs), "rope_cis_128x256_scaled": digest(cis), "silu_table": digest(O.silu_table_bf16()),
    }


if __name__ == "__main__":
    with open(OUT, "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", OUT)
