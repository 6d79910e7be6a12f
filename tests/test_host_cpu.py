"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host logic, CLI parsing."""
import ctypes
import os
import re

import numpy as np
import pytest

import nnacousticmodeling_b200 as nn
from nnacousticmodeling_b200 import _native
from oracle import nnam_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    _native.build()
    return _native.lib()


def test_library_exports_every_symbol_in_header(built_lib):
    hdr = open(os.path.join(ROOT, "include", "nnam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nnam_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(raw, sym), f"{sym} declared in include/nnam_b200.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS)
    assert built_lib.nnam_abi_version() == _native.ABI_VERSION


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(nn.NnamError, match="no CPU or PyTorch fallback"):
        _native.lib()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nnacousticmodeling_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src, f"{f} must not reference the oracle"


def test_feature_transform_parser_and_adaptation(golden_dir):
    ft = nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform"))
    ref = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    assert ft["shape"] == ref["shape"] and ft["shifts"] == ref["shifts"]
    assert np.array_equal(ft["addShift"], ref["addShift"]) and np.array_equal(ft["rescale"], ref["rescale"])
    for net, spl in (("lstm", 0), ("gru", 0), ("tdnn", 8), ("ff", 5)):
        a = nn.adapt_transform(ft, net, spl, nn.is_nn_recurrent(net))
        b = O.select_transform_for_network(ref, net, spl)
        assert a["shape"] == b["shape"] and a["shifts"] == b["shifts"]
        assert np.array_equal(a["addShift"], b["addShift"]) and np.array_equal(a["rescale"], b["rescale"])
    assert ft["shape"] == [440, 40]  # adapt_transform must not modify its input


def test_get_nn_dispatch_and_param_layout():
    kinds = {"ff": nn.MLP, "lstm": nn.LSTM, "zoneoutlstm": nn.ZoneoutLSTM, "zoneoutdropoutlstm": nn.ZoneoutDropoutLSTM,
             "peepholelstm": nn.PeepholeLSTM, "gru": nn.GRU, "mgrurelu": nn.NetMGRU, "mgrurelur": nn.NetMGRU}
    for name, cls in kinds.items():
        drop = [0.5, 0.5] if name == "zoneoutlstm" else ([0, 0.5, 0.5] if name == "zoneoutdropoutlstm" else [0])
        m = nn.get_nn(name, 2, [16], 39, nn.F.relu, [5], drop)
        assert type(m) is cls and m.network == name
        assert nn.is_nn_recurrent(name) == m.recurrent
        m.init_params(12)
        ref = O.init_mlp(np.random.default_rng(0), 12, 16, 2, 39) if name == "ff" else \
            O.init_recurrent(np.random.default_rng(0), name, 12, 16, 2, 39)
        assert {k: v.shape for k, v in m.params.items()} == {k: v.shape for k, v in ref.items()}
    assert np.all(nn.get_nn("lstm", 1, [4], 3, "relu", [5]).init_params(5).params["layer_0/upward/b"][2::4] == 1)
    t = nn.get_nn("tdnn", 4, [8, 8, 8, 8], 39, nn.F.relu, [5, 5, 5, 5])
    assert t.input_win_size == 17
    with pytest.raises(SystemExit):
        nn.get_nn("nope", 1, [4], 3, "relu", [5])
    with pytest.raises(nn.NnamError):
        nn.get_nn("ff", 2, [16], 39, "relu", [5]).load_params({"layer_0/W": np.zeros((16, 12))})


def test_partitions_cover_everything():
    off = O.synth_set(3, 57)[1]
    for parts in (1, 2, 4, 8, 57, 64):
        sh = nn.partition_utterances(off, parts)
        assert sh[0][0] == 0 and sh[-1][1] == 57 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        fr = nn.partition_frames(int(off[-1]), parts)
        assert fr[0][0] == 0 and fr[-1][1] == off[-1] and all(a[1] == b[0] for a, b in zip(fr, fr[1:]))
    loads = [off[b] - off[a] for a, b in nn.partition_utterances(off, 8)]
    assert max(loads) < 1.3 * (off[-1] / 8)


def test_lab_writer_matches_reference_format(golden_dir, tmp_path):
    want = open(os.path.join(golden_dir, "sample.lab"), "rb").read()
    y = np.load(os.path.join(golden_dir, "head.npz"))["logsoftmax"][:4]
    nn.saveBin(str(tmp_path / "a.lab"), y)
    assert (tmp_path / "a.lab").read_bytes() == want
    assert np.array_equal(nn.loadBin(str(tmp_path / "a.lab")), y)


def test_product_synth_matches_the_oracles_generator():
    """bench.py's product arm builds its workload through the package (nnacousticmodeling_b200.synth); the tests and the
    CPU baseline use the oracle's generator: same seed, same data."""
    from nnacousticmodeling_b200 import synth
    for seed, utts, iv, total in ((3, 17, 0, None), (4, 40, 12, 12345)):
        a, b = synth.synth_set(seed, utts, 40, iv, total=total), O.synth_set(seed, utts, 40, iv, total=total)
        for u, v in zip(a, b):
            assert (u is None and v is None) or np.array_equal(u, v)
    assert synth.TIMIT_TRAIN_FRAMES == 1124823


def test_precision_grammar():
    from nnacousticmodeling_b200.engine import Precision
    from nnacousticmodeling_b200.ops import SPLIT_A, SPLIT_AW, SPLIT_NONE, SPLIT_W
    p = Precision("bf16+a:0/6+w:-1")
    assert [p.nsplit(l, 7) for l in range(7)] == [SPLIT_A, SPLIT_NONE, SPLIT_NONE, SPLIT_NONE, SPLIT_NONE, SPLIT_NONE,
                                                  SPLIT_AW]
    assert Precision("bf16+w").nsplit(3, 7) == SPLIT_W and Precision("fp32").nsplit(0, 7) == SPLIT_AW
    assert Precision("fp16").nsplit(0, 7) == SPLIT_NONE and not Precision("fp16").split
    for bad in ("int8", "fp16+a", "bf16+x"):
        with pytest.raises(nn.NnamError):
            Precision(bad)


def test_npy_file_sink_writes_what_np_save_writes(tmp_path):
    sink_cls = __import__("importlib").import_module("nnacousticmodeling_b200.predict").NpyFileSink
    y = np.random.default_rng(0).standard_normal((1000, 39)).astype(np.float32)
    s = sink_cls(str(tmp_path / "a.npy"), 1000, 39)
    for r0 in range(0, 1000, 300):
        s.write(r0, min(r0 + 300, 1000), y[r0:r0 + 300])
    s.close()
    np.save(str(tmp_path / "b.npy"), y)
    assert np.array_equal(np.load(str(tmp_path / "a.npy")), y)
    assert open(str(tmp_path / "a.npy"), "rb").read() == open(str(tmp_path / "b.npy"), "rb").read()
    bad = sink_cls(str(tmp_path / "c.npy"), 10, 39)
    with pytest.raises(nn.NnamError):
        bad.write(5, 10, y[:5])


def test_transfer_stats_probe_then_keep_the_best_fraction(monkeypatch):
    """The compact fraction of the transfer mix is probed -- model, 0 (float32 rows only), a third candidate, the first
    again -- and the best measured one is kept; NNAM_TRANSFER_PROBE=0 uses the model's value directly."""
    from nnacousticmodeling_b200.engine import _TransferStats
    monkeypatch.delenv("NNAM_TRANSFER_PROBE", raising=False)
    st = _TransferStats(4)
    assert st.probe
    seen = []
    for _ in range(8):
        x = st.compact_fraction()
        seen.append(x)
        rate = {True: 11e6, False: 9e6}[x == 0.0] if x in (0.0, seen[0]) else 12e6
        st.update([], None, 500000, 500000 / rate)
    assert seen[1] == 0.0 and len(set(seen[:3])) == 3
    assert seen[3] == seen[0]                           # the first candidate again: its first pass was a cold one
    assert all(x == seen[2] for x in seen[4:])          # the third candidate measured best and is kept
    assert len(st.tried) == 3
    monkeypatch.setenv("NNAM_TRANSFER_PROBE", "0")
    st2 = _TransferStats(16)
    assert not st2.probe and st2.compact_fraction() == st2.model_fraction()


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/nnam_b200.h against the ctypes table of _native.py: same number of arguments, and the
    same KIND per argument (pointer / integer width / float), so a changed C signature cannot silently drift from the
    binding (ctypes would pass garbage, not fail)."""
    from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p
    hdr = open(os.path.join(ROOT, "include", "nnam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b([A-Za-z_][\w \*]*?)\b(nnam_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) == len(_native._SIGNATURES)

    def kind_of_c(decl):
        decl = decl.strip()
        if decl in ("void", ""):
            return None
        if "*" in decl:
            return "ptr"
        t = re.sub(r"\b(const|unsigned|signed)\b", "", decl).split()
        base = " ".join(t[:-1]) if len(t) > 1 else t[0]
        return {"int": "i32", "long long": "i64", "float": "f32"}[base]

    def kind_of_ctypes(t):
        if t in (c_int,):
            return "i32"
        if t is c_longlong:
            return "i64"
        if t is c_float:
            return "f32"
        if t in (c_void_p, c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is not None:
            return "ptr"
        raise AssertionError(f"unexpected ctypes type {t}")

    for ret, name, args in protos:
        res, argtypes = _native._SIGNATURES[name]
        c_kinds = [k for k in (kind_of_c(a) for a in args.split(",")) if k is not None]
        py_kinds = [kind_of_ctypes(t) for t in argtypes]
        assert c_kinds == py_kinds, f"{name}: header {c_kinds} vs ctypes {py_kinds}"
        assert ("*" in ret) == (res in (c_char_p, c_void_p)), name
