"""the native .pth reader / writer (csrc/pth.cpp) against PyTorch itself.

PyTorch is the independent second opinion on the file format: files written by torch.save must come back
bit-identical through lnb_pth_*, files written by lnb_pth_writer_* must load with torch.load(weights_only=True),
and the pickle we emit may only use the opcodes the reference's unpickler dispatches
(src/pickle/pickledispatch.go:52-77) so that the unmodified Go loader can read the synthetic checkpoint."""
import ctypes as C
import json
import os
import pickletools
import subprocess
import sys
import zipfile

import numpy as np
import pytest
import torch

import lnb_b200 as L
from lnb_b200.torch_reader import TorchModelReader, TorchModelWriter, write_synthetic_checkpoint
from tests.helpers import host_tensors

# opcodes registered in src/pickle/pickledispatch.go:52-77
REFERENCE_OPCODES = {"PROTO", "EMPTY_DICT", "BINPUT", "MARK", "BINUNICODE", "GLOBAL", "BININT", "BINSTRING", "TUPLE", "BINPERSID",
                     "BININT1", "BININT2", "TUPLE1", "TUPLE2", "TUPLE3", "NEWTRUE", "NEWFALSE", "EMPTY_TUPLE", "REDUCE", "BINGET",
                     "LONG_BINPUT", "STOP", "SHORT_BINSTRING", "SETITEMS"}


def bits(t: torch.Tensor) -> np.ndarray:
    return t.view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.numpy()


def state_dict():
    g = torch.Generator().manual_seed(0)
    return {"tok_embeddings.weight": torch.randn(64, 32, generator=g).bfloat16(), "norm.weight": torch.randn(32, generator=g).bfloat16(),
            "layers.0.attention.wq.weight": torch.randn(32, 32, generator=g).bfloat16(), "aux.f32": torch.randn(3, 5, generator=g),
            "aux.f16": torch.randn(2, 2, generator=g).half(), "aux.i64": torch.arange(7), "aux.u8": torch.arange(5, dtype=torch.uint8),
            "aux.bool": torch.tensor([True, False, True]), "scalar": torch.tensor(3.5), "empty": torch.zeros(0, 4).bfloat16()}


@pytest.mark.parametrize("proto", [2, 3, 4, 5])
def test_reads_what_torch_save_writes(tmp_path, proto):
    sd = state_dict()
    p = str(tmp_path / "a.pth")
    torch.save(sd, p, pickle_protocol=proto)
    with TorchModelReader(p) as r:
        t = r.Load()
        assert list(t) == list(sd)                                   # dict order preserved (PickleDict keys)
        for k, v in sd.items():
            assert t[k].Size == tuple(v.shape) and t[k].contiguous
            assert t[k].DataType == str(v.dtype).replace("torch.", "")
            assert np.array_equal(t[k].RawData.reshape(v.shape), bits(v)), k
            assert not t[k].RawData.flags.writeable                  # aliases the read-only mapping
            raw = open(p, "rb").read()
            assert raw[t[k].file_offset:t[k].file_offset + t[k].nbytes] == bits(v).tobytes()


def test_reads_module_state_dicts_views_and_flags_non_contiguous(tmp_path):
    lin = torch.nn.Linear(4, 3).bfloat16()
    p = str(tmp_path / "b.pth")
    torch.save(lin.state_dict(), p)                                   # OrderedDict + BUILD(_metadata)
    with TorchModelReader(p) as r:
        t = r.Load()
        assert np.array_equal(t["weight"].RawData, bits(lin.weight.detach())) and t["bias"].Size == (3,)
    base = torch.arange(24, dtype=torch.float32)
    p = str(tmp_path / "c.pth")
    torch.save({"v1": base[4:12].view(2, 4), "v2": base[12:], "nc": base.view(4, 6).t(), "n": 5, "nested": {"x": base}}, p)
    with TorchModelReader(p) as r:
        t = r.Load()
        assert set(t) == {"v1", "v2", "nc"}                           # non-tensor entries are skipped
        assert np.array_equal(t["v1"].RawData, base[4:12].view(2, 4).numpy())     # storage offset honoured
        assert np.array_equal(t["v2"].RawData, base[12:].numpy())
        assert not t["nc"].contiguous and t["nc"].Size == (6, 4)
        assert t["nc"].nbytes == 24 * 4 and t["nc"].RawData.size == 24     # the touched span of the storage, not numel


def test_strided_views_never_map_past_their_storage(tmp_path):
    """an expanded view has numel >> storage: the reader must hand out the storage span only (round-1 advisor
    finding: nbytes = numel * itemsize mapped 512 MiB over a 1.8 KB file and segfaulted on first touch)"""
    p = str(tmp_path / "x.pth")
    torch.save({"x": torch.zeros(1, dtype=torch.bfloat16).expand(65536, 4096), "col": torch.arange(12.).view(3, 4)[:, 1]}, p)
    with TorchModelReader(p) as r:
        t = r.Load()
        assert t["x"].Size == (65536, 4096) and not t["x"].contiguous
        assert t["x"].nbytes == 2 and int(t["x"].RawData.sum()) == 0       # one element reachable; touching it is safe
        assert not t["col"].contiguous and t["col"].nbytes == 9 * 4        # 1 + 2*4 elements from the offset
        assert np.array_equal(t["col"].RawData[::4], np.array([1., 5., 9.], np.float32))


@pytest.mark.parametrize("force_zip64", [False, True])
def test_torch_load_reads_what_the_writer_writes(tmp_path, force_zip64):
    arrs = {"x.bf16": np.arange(12, dtype=np.uint16).reshape(3, 4) + 0x3F80, "y.f32": np.linspace(0, 1, 10, dtype=np.float32),
            "z.i64": np.arange(6, dtype=np.int64).reshape(1, 2, 3), "s": np.float32(2.5).reshape(()), "e": np.zeros((0, 3), np.uint16),
            "big": (np.arange(70000, dtype=np.uint32) % 65536).astype(np.uint16)}
    p = str(tmp_path / "w.pth")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import lnb_b200 as L\n"
            "from lnb_b200.torch_reader import TorchModelWriter\n"
            "d = np.load(%r)\nw = TorchModelWriter(%r)\n[w.Add(k, d[k]) for k in d.files]\nw.Finish()\n")
    np.savez(str(tmp_path / "in.npz"), **arrs)
    env = dict(os.environ, LNB_PTH_FORCE_ZIP64="1" if force_zip64 else "0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code % (root, str(tmp_path / "in.npz"), p)], check=True, env=env)
    back = torch.load(p, weights_only=True)                            # the strict loader: only allow-listed globals
    assert list(back) == list(arrs)
    for k, a in arrs.items():
        assert tuple(back[k].shape) == a.shape and np.array_equal(bits(back[k]), a), k
    z = zipfile.ZipFile(p)
    assert z.testzip() is None                                         # CRC-32 of every entry
    names = z.namelist()
    assert "archive/data.pkl" in names and "archive/version" in names and "archive/data/0" in names
    raw = open(p, "rb").read()
    for info in z.infolist():
        if info.filename.startswith("archive/data/"):
            assert info.compress_type == zipfile.ZIP_STORED
            hdr = info.header_offset
            start = hdr + 30 + int.from_bytes(raw[hdr + 26:hdr + 28], "little") + int.from_bytes(raw[hdr + 28:hdr + 30], "little")
            assert start % 64 == 0                                     # torch's storage alignment
    ops = [op.name for op, _, _ in pickletools.genops(z.read("archive/data.pkl"))]
    assert set(ops) <= REFERENCE_OPCODES and "TUPLE3" not in ops        # TUPLE3 is mis-read by the reference (:236-240)
    with TorchModelReader(p) as r:                                      # and our own reader (zip64 records included)
        t = r.Load()
        for k, a in arrs.items():
            assert np.array_equal(t[k].RawData.reshape(a.shape), a), k


def test_synthetic_checkpoint_has_the_reference_inventory_and_the_generator_bits(tmp_path):
    args = dict(L.synth.TINY)
    p = str(tmp_path / "consolidated.00.pth")
    write_synthetic_checkpoint(p, L.synth.args_c(args), 77)
    sd = torch.load(p, weights_only=True)
    shapes = L.synth.tensor_shapes(args)
    assert set(sd) == set(shapes) and len(sd) == 3 + 9 * args["n_layers"]
    assert list(sd)[0] == "tok_embeddings.weight" and list(sd)[-2:] == ["norm.weight", "output.weight"]   # Meta's order
    with TorchModelReader(p) as r:
        mine = r.Load()
    oracle_bits = host_tensors(args, 77)                               # the ORACLE's generator
    for name, shape in shapes.items():
        assert tuple(sd[name].shape) == tuple(shape) and sd[name].dtype == torch.bfloat16
        assert np.array_equal(bits(sd[name]), oracle_bits[name]), name
        assert np.array_equal(mine[name].RawData, oracle_bits[name]), name
    ops = {op.name for op, _, _ in pickletools.genops(zipfile.ZipFile(p).read("archive/data.pkl"))}
    assert ops <= REFERENCE_OPCODES


def test_params_json_follows_modelargs_go(tmp_path):
    # the params.json Meta ships with Llama-3.1-8B (docs/04-LOADING-MODEL-ARGS.md)
    meta = {"dim": 4096, "n_layers": 32, "n_heads": 32, "n_kv_heads": 8, "vocab_size": 128256, "ffn_dim_multiplier": 1.3,
            "multiple_of": 1024, "norm_eps": 1e-05, "rope_theta": 500000.0, "use_scaled_rope": True}
    (tmp_path / "params.json").write_text(json.dumps(meta))
    d = L.model.load_model_args(str(tmp_path))
    assert (d["dim"], d["n_layers"], d["n_heads"], d["n_kv_heads"], d["head_dim"], d["ffn_dim"], d["vocab_size"]) == \
        (4096, 32, 32, 8, 128, 14336, 128256)
    assert d["use_scaled_rope"] == 1 and d["rope_theta"] == 500000.0 and abs(d["norm_eps"] - 1e-5) < 1e-12 and d["max_seq_len"] == 2048
    # defaults of NewModelArgs (modelargs.go:29-44): n_kv_heads -> n_heads, no multiplier, multiple_of 256, no vocab
    (tmp_path / "params.json").write_text('{"dim": 512, "n_heads": 8, "n_layers": 2, "ffn_dim_multiplier": null, "unknown": [1, {"a": "b}"}], "s": "x"}')
    d = L.model.load_model_args(str(tmp_path), max_seq_len=64)
    assert d["n_kv_heads"] == 8 and d["head_dim"] == 64 and d["vocab_size"] == -1 and d["use_scaled_rope"] == 0 and d["max_seq_len"] == 64
    hidden = int(2 * (4 * 512) / 3)
    assert d["ffn_dim"] == 256 * ((hidden + 255) // 256)
    (tmp_path / "params.json").write_text('{"dim": 4096,')
    with pytest.raises(L._capi.LnbError):
        L.model.load_model_args(str(tmp_path))
    with pytest.raises(L._capi.LnbError):
        L.model.load_model_args(str(tmp_path / "missing"))


def test_malformed_checkpoints_are_errors_not_crashes(tmp_path):
    sd = {"a": torch.ones(4, 4).bfloat16()}
    good = str(tmp_path / "g.pth")
    torch.save(sd, good)
    raw = open(good, "rb").read()

    def expect_error(data, what):
        p = str(tmp_path / "bad.pth")
        open(p, "wb").write(data)
        with pytest.raises(L._capi.LnbError) as e:
            TorchModelReader(p)
        assert what in str(e.value), str(e.value)

    expect_error(b"", "empty")
    expect_error(raw[:100], "zip")
    expect_error(raw[:-30], "zip")
    with pytest.raises(L._capi.LnbError):
        TorchModelReader(str(tmp_path / "does_not_exist.pth"))
    # a zip without data.pkl (torchmodelreader.go:48-50)
    p = str(tmp_path / "nopkl.pth")
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("archive/version", "3\n")
    with pytest.raises(L._capi.LnbError) as e:
        TorchModelReader(p)
    assert "no .pkl file found" in str(e.value)
    # a class the reference does not know either (findClassTorch)
    import pickle
    p = str(tmp_path / "cls.pth")
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("archive/data.pkl", pickle.dumps({"a": np.float32}, protocol=2) if False else b"\x80\x02cfoo\nbar\n)R.")
    with pytest.raises(L._capi.LnbError) as e:
        TorchModelReader(p)
    assert 'unknown class "foo.bar" not found' in str(e.value)
    # compressed storages cannot be mapped
    p = str(tmp_path / "deflated.pth")
    with zipfile.ZipFile(good) as zin, zipfile.ZipFile(p, "w", zipfile.ZIP_DEFLATED) as zout:
        for info in zin.infolist():
            zout.writestr(info.filename, zin.read(info.filename), zipfile.ZIP_STORED if info.filename.endswith(".pkl") else zipfile.ZIP_DEFLATED)
    with pytest.raises(L._capi.LnbError) as e:
        TorchModelReader(p)
    assert "compressed" in str(e.value)


def test_fuzzed_checkpoints_never_crash_the_reader(tmp_path):
    """truncations and byte flips of a valid file: every outcome must be a clean error or a clean parse
    (run in a child process so that a crash would be seen as a failure, not take pytest down)"""
    sd = {f"t{i}": torch.randn(8, 8).bfloat16() for i in range(4)}
    good = str(tmp_path / "g.pth")
    torch.save(sd, good)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, random
sys.path.insert(0, %r)
import lnb_b200 as L
from lnb_b200.torch_reader import TorchModelReader
raw = bytearray(open(%r, "rb").read())
rnd = random.Random(1)
n_ok = n_err = 0
for trial in range(600):
    b = bytearray(raw)
    kind = trial %% 3
    if kind == 0:
        b = b[:rnd.randrange(len(b))]
    elif kind == 1:
        for _ in range(rnd.randrange(1, 6)):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
    else:   # corrupt the pickle / directory regions specifically
        lo = b.find(b"archive/data.pkl")
        for _ in range(3):
            i = rnd.randrange(lo, min(len(b), lo + 400)) if rnd.random() < 0.5 else rnd.randrange(max(0, len(b) - 600), len(b))
            b[i] = rnd.randrange(256)
    p = %r
    open(p, "wb").write(b)
    try:
        with TorchModelReader(p) as r:
            for t in r.Load().values():
                if t.nbytes: t.RawData.reshape(-1)[-1]     # touch the last byte: in-bounds or the reader lied
        n_ok += 1
    except L._capi.LnbError:
        n_err += 1
print("ok", n_ok, "err", n_err)
''' % (root, good, str(tmp_path / "f.pth"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "err" in out.stdout


def test_sanitized_mutation_fuzz_of_the_native_reader(tmp_path):
    """tests/native/pth_fuzz.cpp + csrc/pth.cpp under AddressSanitizer / UBSan: 8000 mutants, zero findings"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pth_fuzz")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-I", os.path.join(root, "llama-nuts-and-bolts_b200", "csrc"), os.path.join(root, "tests", "native", "pth_fuzz.cpp"),
                           os.path.join(root, "llama-nuts-and-bolts_b200", "csrc", "pth.cpp"), "-o", exe])
    good = str(tmp_path / "good.pth")
    sd = {f"t{i}": torch.randn(8, 8).bfloat16() for i in range(4)}
    sd["view"] = torch.arange(24.0)[4:12].view(2, 4)
    torch.save(sd, good)
    out = subprocess.run([exe, good, str(tmp_path / "scratch.pth"), "8000", "5"], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert "parsed" in out.stdout and "rejected" in out.stdout


def test_cpp_host_cli_writes_a_model_directory_without_a_gpu(tmp_path):
    """host/lnb_generate --write-synthetic: params.json + consolidated.00.pth through the C++ host mirror (host-only);
    the files satisfy PyTorch and derive TINY's widths the reference's way"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "host"), "-s"])
    subprocess.check_call([os.path.join(root, "host", "lnb_generate"), "--write-synthetic", str(tmp_path), "tiny"])
    d = L.model.load_model_args(str(tmp_path), max_seq_len=64)
    tiny = dict(L.synth.TINY)
    assert {k: d[k] for k in ("dim", "n_layers", "n_heads", "n_kv_heads", "head_dim", "ffn_dim", "vocab_size")} == \
        {k: tiny[k] for k in ("dim", "n_layers", "n_heads", "n_kv_heads", "head_dim", "ffn_dim", "vocab_size")}
    # the stand-in tokenizer.model makes the directory pass checkModelArgs: vocabulary length == VocabSize
    from lnb_b200 import vocabulary
    v = vocabulary.Load(str(tmp_path / "tokenizer.model"))
    assert len(v) == tiny["vocab_size"] and v.BeginOfSentenceId == tiny["vocab_size"] - 256
    ids = v.TokenizeString("Hi é!\n")
    assert ids == list("Hi é!\n".encode()) and v.TokenBatchToString(ids) == "Hi é!\n"
    v.close()
    sd = torch.load(str(tmp_path / "consolidated.00.pth"), weights_only=True)
    expect = host_tensors(tiny, 7)
    assert set(sd) == set(expect)
    for k in ("tok_embeddings.weight", "layers.1.feed_forward.w2.weight", "norm.weight"):
        assert np.array_equal(bits(sd[k]), expect[k])


@pytest.mark.skipif(os.environ.get("LNB_SLOW_TESTS") != "1", reason="writes a 5 GiB file; set LNB_SLOW_TESTS=1")
def test_archive_beyond_4gib_uses_real_zip64_offsets(tmp_path):
    """five 1 GiB tensors: local-header offsets pass 4 GiB, so the central directory carries zip64 extras and the
    zip64 end records are mandatory -- the shape of the 16 GB consolidated.00.pth.  torch.load(mmap=True) and the
    native reader must both return the data (last verified by hand: 5.0 GiB, identical)."""
    p = str(tmp_path / "big.pth")
    base = np.random.default_rng(0).integers(0, 65536, size=(1 << 27), dtype=np.uint16)

    def block(i):
        return np.concatenate([np.roll(base, i * 4 + j) for j in range(4)]).reshape(16384, 32768)
    w = TorchModelWriter(p)
    for i in range(5):
        w.Add(f"layers.{i}.big.weight", block(i))
    w.Add("tail.weight", np.arange(10, dtype=np.uint16))
    w.Finish()
    assert os.path.getsize(p) > 5 * 2**30
    sd = torch.load(p, weights_only=True, mmap=True)
    with TorchModelReader(p) as r:
        ts = r.Load()
        assert ts["tail.weight"].file_offset > 5 * 2**30
        for i in range(5):
            exp = block(i)
            assert np.array_equal(sd[f"layers.{i}.big.weight"].view(torch.uint16).numpy(), exp)
            assert np.array_equal(ts[f"layers.{i}.big.weight"].RawData, exp)
        assert np.array_equal(ts["tail.weight"].RawData, np.arange(10, dtype=np.uint16))
