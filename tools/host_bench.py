"""Host-side measurements of the widened rows, no GPU needed:
  tokenizer  -- MB/s and tokens/s of lnb_tokenize_string on a multilingual corpus (synthetic 2 k-merge vocabulary)
               next to tiktoken on the same vocabulary / pattern
  checkpoint -- lnb_pth_write_synthetic GB/s (generator + CRC-32 + file write) and lnb_pth_open latency (zip index +
               unpickle of 291 tensors) on a 4-layer slice of the 8B architecture (2.3 GB)
Usage: python tools/host_bench.py [scratch dir=/tmp]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tiktoken

import lnb_b200 as L
from lnb_b200.torch_reader import TorchModelReader, write_synthetic_checkpoint
from lnb_b200.vocabulary import Load
from tests.test_tokenizer_cpu import CORPUS, GO_PAT, train_bpe, write_model

scratch = sys.argv[1] if len(sys.argv) > 1 else tempfile.gettempdir()
ranks = train_bpe(CORPUS * 4, 2000)
path = os.path.join(scratch, "host_bench_tokenizer.model")
write_model(path, ranks)
v = Load(path)
text = CORPUS * 400                                            # ~0.6 MB
b = text.encode()
t0 = time.perf_counter(); ids = v.TokenizeString(text); dt = time.perf_counter() - t0
enc = tiktoken.Encoding("hb", pat_str=GO_PAT, mergeable_ranks=ranks, special_tokens={})
t0 = time.perf_counter(); ref = enc.encode_ordinary(text); dt_ref = time.perf_counter() - t0
assert ids == ref
print(f"tokenizer: {len(b) / 1e6:.2f} MB -> {len(ids)} tokens; lnb {len(b) / 1e6 / dt:.1f} MB/s ({len(ids) / dt / 1e6:.2f} M tok/s), "
      f"tiktoken {len(b) / 1e6 / dt_ref:.1f} MB/s; identical ids")
t0 = time.perf_counter(); out = v.TokenBatchToBytes(ids); dt = time.perf_counter() - t0
assert out == b
print(f"detokenizer: {len(ids) / dt / 1e6:.1f} M tok/s")
os.remove(path)

args = dict(L.synth.LLAMA31_8B, n_layers=4)
p = os.path.join(scratch, "host_bench.pth")
t0 = time.perf_counter(); write_synthetic_checkpoint(p, L.synth.args_c(args), 1); dt = time.perf_counter() - t0
size = os.path.getsize(p)
print(f"checkpoint writer: {size / 1e9:.2f} GB in {dt:.1f} s = {size / 1e9 / dt:.2f} GB/s (generator + CRC-32 + write)")
t0 = time.perf_counter()
with TorchModelReader(p) as r:
    n = len(r.Load())
dt = time.perf_counter() - t0
print(f"checkpoint reader: open + index + unpickle + {n} tensor views in {dt * 1e3:.1f} ms (data stays in the page cache)")
os.remove(p)
