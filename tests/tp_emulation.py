"""Tensor-parallel forward of the tiny model written with the oracle's per-op functions and a
torch.distributed all-reduce -- the host-side twin of what liblnb.so does per rank (one fp32
all-reduce after Wo and one after w2).  Used by the world_size-2 gloo test on CPU."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O


def shard(tensors, args, name, rank, tp, shard_window):
    r0, c0, nr, nc = shard_window(args, name, rank, tp)
    t = tensors[name]
    return np.ascontiguousarray(t if t.ndim == 1 else t[r0:r0 + nr, c0:c0 + nc])


def tp_forward(args, tensors, tokens, start_pos, caches, rank, tp, allreduce, shard_window):
    """returns this rank's slice of the last-row logits [vocab/tp] (f32)"""
    D, hd = args["dim"], args["head_dim"]
    nh_l, nkv_l = args["n_heads"] // tp, args["n_kv_heads"] // tp
    S = len(tokens)
    _, cis = O.rope_table(hd, args["max_seq_len"] * 2, args["rope_theta"], bool(args["use_scaled_rope"]))
    sh = lambda n: shard(tensors, args, n, rank, tp, shard_window)
    x = tensors["tok_embeddings.weight"][tokens]
    for l in range(args["n_layers"]):
        pre = f"layers.{l}."
        xn = O.rmsnorm(x, tensors[pre + "attention_norm.weight"], args["norm_eps"])
        q = O.linear_bf16(xn, sh(pre + "attention.wq.weight")).reshape(S, nh_l, hd)
        k = O.linear_bf16(xn, sh(pre + "attention.wk.weight")).reshape(S, nkv_l, hd)
        v = O.linear_bf16(xn, sh(pre + "attention.wv.weight")).reshape(S, nkv_l, hd)
        q, k = O.rope_apply(q, cis, start_pos), O.rope_apply(k, cis, start_pos)
        ck, cv = caches[l]
        ck[start_pos:start_pos + S], cv[start_pos:start_pos + S] = k, v
        T = start_pos + S
        o = O.attention(q, ck, cv, T, S > 1)
        part = O.linear_bf16_f32out(o, sh(pre + "attention.wo.weight"))           # K-split partial, untruncated
        tot = allreduce(part)
        h1 = O.add_bf16(x, O.bf16_bits(tot).reshape(S, D))
        hn = O.rmsnorm(h1, tensors[pre + "ffn_norm.weight"], args["norm_eps"])
        g = O.silu_bf16(O.linear_bf16(hn, sh(pre + "feed_forward.w1.weight")))
        u = O.linear_bf16(hn, sh(pre + "feed_forward.w3.weight"))
        m = O.mul_bf16(g, u)
        part = O.linear_bf16_f32out(m, sh(pre + "feed_forward.w2.weight"))
        tot = allreduce(part)
        x = O.add_bf16(h1, O.bf16_bits(tot).reshape(S, D))
    xf = O.rmsnorm(x, tensors["norm.weight"], args["norm_eps"])
    lg = O.linear_bf16(xf[-1:], sh("output.weight"))
    return O.bf16_to_f32(lg)[0]
