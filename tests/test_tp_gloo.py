"""world_size-2 `gloo` test (CPU): the multi-process host logic of the tensor-parallel path --
rendezvous on 127.0.0.1, the library's shard windows, one fp32 all-reduce after Wo and one after
w2, vocab-sharded logits + (max, index) argmax merge -- reproduces the oracle's TP emulation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lnb_b200 as L
        from oracle import oracle as O
        from tests.helpers import host_tensors, oracle_model
        from tests.tp_emulation import tp_forward
        O.lib().orc_set_num_threads(2)
        args = dict(L.synth.TINY)
        tensors = host_tensors(args, 4321)
        nkv_l, hd, seq = args["n_kv_heads"] // world, args["head_dim"], 16
        caches = [(np.zeros((seq, nkv_l, hd), np.uint16), np.zeros((seq, nkv_l, hd), np.uint16)) for _ in range(args["n_layers"])]

        def allreduce(part):
            t = torch.from_numpy(part.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)           # 2 ranks: a+b is commutative -> bit-exact
            return t.numpy()

        prompt = np.array([5, 900, 33, 7, 64], np.int32)
        toks, pos, cur = [], 0, prompt
        for _ in range(4):
            lg = tp_forward(args, tensors, cur, pos, caches, rank, world, allreduce, L.synth.shard_window)
            # greedy argmax across the vocab shards: (value, lowest global index) like lnb's key merge
            v_l = args["vocab_size"] // world
            i = O.argmax_f32(np.ascontiguousarray(lg))
            cand = torch.tensor([float(lg[i]), float(-(rank * v_l + i))], dtype=torch.float64)
            allc = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(allc, cand)
            best = max(allc, key=lambda c: (c[0].item(), c[1].item()))
            nxt = int(-best[1].item())
            full = [torch.zeros(v_l) for _ in range(world)]
            dist.all_gather(full, torch.from_numpy(lg.copy()))
            toks.append((nxt, torch.cat(full).numpy()))
            pos += len(cur)
            cur = np.array([nxt], np.int32)
        if rank == 0:
            om = oracle_model(args, tensors)
            sess = om.new_session(seq)
            pos, cur, ok = 0, prompt, True
            for nxt, logits in toks:
                exp = sess.forward(cur, pos, all_rows=False, tp=world)[0]
                ok &= bool(np.array_equal(exp, logits)) and (O.argmax_f32(exp) == nxt)
                pos += len(cur)
                cur = np.array([nxt], np.int32)
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_tp2_gloo_matches_oracle_tp_emulation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, "a rank failed"
    assert q.get(timeout=5) is True
