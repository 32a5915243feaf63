"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host packing logic,
state_dict compatibility with the reference layout, loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT
from flowchain_iccv2023_b200 import _lib, packing, synth


def header_symbols():
    text = open(os.path.join(ROOT, "include", "flowchain_b200.h")).read()
    return sorted(set(re.findall(r"\b(fc_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert set(syms) == set(_lib.SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s


def test_blob_size_matches_c_side():
    lib = _lib.load()
    for spec in (synth.FlowSpec(), synth.FlowSpec(pred_len=5, cond_size=9, hidden_size=40, n_hidden=1, n_blocks=2)):
        d = _lib.ModelDesc(spec.pred_len, spec.input_size, spec.cond_size, spec.hidden_size, spec.n_hidden, spec.n_blocks)
        blob = packing.flatten_state(spec, synth.make_params(spec, 0))
        assert blob.dtype == np.float32 and blob.size == lib.fc_weight_blob_floats(ctypes.byref(d))


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    d = _lib.ModelDesc(12, 2, 16, 64, 2, 3)
    ptr = ctypes.c_void_p()
    assert lib.fc_create(ctypes.byref(ptr), ctypes.byref(d), 0) != 0
    assert "no CPU fallback" in _lib.last_error(None)


def test_bad_desc_rejected():
    lib = _lib.load()
    ptr = ctypes.c_void_p()
    d = _lib.ModelDesc(12, 3, 16, 64, 2, 3)
    assert lib.fc_create(ctypes.byref(ptr), ctypes.byref(d), 0) != 0
    assert "input_size" in _lib.last_error(None)


def test_mirror_state_dict_layout_and_cpu_refusal():
    from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond
    spec, params, g = load_golden("default_eth")
    f = fastpredNF_CIF_separate_cond(2, 3, 64, 2, 16, flow_architecture="realNVP", pred_len=12)
    assert sum(p.numel() for p in f.parameters()) == 1089352
    assert set(f.state_dict().keys()) == set(params.keys())
    f.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    blob = packing.flatten_state(spec, f.state_dict())
    np.testing.assert_array_equal(blob, packing.flatten_state(spec, params))
    f.eval()                              # inference mode: fused kernels only, and they need a B200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.log_prob(torch.zeros(1, 2), torch.zeros(1, 12, 2), torch.zeros(1, 12, 16))
    with torch.no_grad():                 # train() mode without grad is still inference
        f.train()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            f.log_prob(torch.zeros(1, 2), torch.zeros(1, 12, 2), torch.zeros(1, 12, 16))
    with pytest.raises(ValueError):
        bad = dict(params)
        bad["var"] = np.zeros((3, 2), np.float32)
        packing.flatten_state(spec, bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_reference_state_dict_keys_match_mirror():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    ref = make_golden.import_reference()
    rf = ref.fastpredNF_CIF_separate_cond(input_size=2, n_blocks=3, n_hidden=2, hidden_size=64, cond_label_size=16,
                                          flow_architecture="realNVP", pred_len=12)
    from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond
    f = fastpredNF_CIF_separate_cond(2, 3, 64, 2, 16, flow_architecture="realNVP", pred_len=12)
    assert list(rf.state_dict().keys()) == list(f.state_dict().keys())
    f.load_state_dict(rf.state_dict(), strict=True)          # a reference checkpoint loads unchanged


def test_synth_is_deterministic():
    a = synth.make_params(synth.FlowSpec(pred_len=2), 3)
    b = synth.make_params(synth.FlowSpec(pred_len=2), 3)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert synth.DEFAULT_SPEC.flops_separate_row() == 2136704
    assert synth.DEFAULT_SPEC.flops_sequential_row() == 12616448


def test_wrapper_load_rejects_mismatched_flow_keys(tmp_path):
    """model.load tolerates only `encoder.*` differences (the reference loads strictly, fastpredNF.py:237-238): a
    checkpoint whose flow.* section is truncated or renamed must raise instead of leaving random-init weights."""
    from flowchain_iccv2023_b200.model import fastpredNF_TP
    m = fastpredNF_TP(encoder=lambda d: None, pred_len=3, n_blocks=1, n_hidden=1, hidden_size=16, cond_size=4)
    path = tmp_path / "ckpt.pt"
    m.save(epoch=7, path=path)
    assert m.load(path) == 7
    state = dict(m.state_dict())
    state["encoder.some.weight"] = torch.zeros(3)                   # a reference encoder's keys are ignored
    torch.save({"epoch": 8, "state": state, "optim_state": {}}, path)
    assert m.load(path) == 8
    state.pop("flow.var")
    torch.save({"epoch": 9, "state": state, "optim_state": {}}, path)
    with pytest.raises(RuntimeError, match="flow.var"):
        m.load(path)
    state = {k.replace("flow.", "model."): v for k, v in m.state_dict().items()}
    torch.save({"epoch": 9, "state": state, "optim_state": {}}, path)
    with pytest.raises(RuntimeError, match="does not match"):
        m.load(path)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: ONE JSON line on stdout (the driver parses it); everything else -- library banners included,
    which is why bench.py points file descriptor 1 at stderr for the run -- goes to stderr.  The reference arm runs on the
    host cores, so the contract can be checked here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"] and d["unit"]


def test_bench_own_arm_refuses_without_gpu():
    """No CPU fallback: without a CUDA device the product arm exits non-zero and prints nothing on stdout."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode != 0 and r.stdout == "" and "no CPU fallback" in r.stderr
