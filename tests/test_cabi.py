"""The C-ABI library loads and exports every symbol include/mrclip.h declares (no compute on CPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mrclip_b200 import _cabi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mrclip.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrclip_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mrclip.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes prototype in _cabi.SIGNATURES"
    assert set(_cabi.SIGNATURES) == set(names)


def test_geometry_entry_points_need_no_gpu():
    lib = _cabi.load()
    assert lib.mrclip_version() >= 100
    assert [lib.mrclip_padded_dim(d) for d in (1, 8, 9, 512, 768)] == [8, 8, 16, 512, 768]
    assert [lib.mrclip_padded_cols(n) for n in (1, 128, 129, 32768)] == [256, 256, 256, 32768]
    assert lib.mrclip_workspace_bytes(0, 10, 10) == 0
    small = lib.mrclip_workspace_bytes(256, 256, 512)
    big = lib.mrclip_workspace_bytes(4096, 32768, 768)
    assert 0 < small < big < (8 << 30)
    g = lib.mrclip_fwd_col_granule(4096, 32768)
    assert g % 128 == 0 and 4096 % g == 0 and (g & (g - 1)) == 0


def test_argument_errors_are_reported_not_fatal():
    lib = _cabi.load()
    rc = lib.mrclip_pack_bf16(None, 7, 4, 4, 4, None, 8, None)
    assert rc < 0 and b"dtype" in lib.mrclip_last_error()
    shape = _cabi.Shape(0, 4, 4, 0)
    rc = lib.mrclip_clip_fwd_tiles(None, None, shape, 8, None, 0, 4, None, None)
    assert rc < 0 and b"empty shape" in lib.mrclip_last_error()
    with pytest.raises(_cabi.MrclipError):
        _cabi.check(rc)


def test_product_refuses_to_run_without_gpu():
    import torch
    import mrclip_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mrclip_b200.set_engine(None)
    loss = mrclip_b200.ClipLoss()
    x = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    with pytest.raises(RuntimeError, match="no CPU (fallback|path)"):
        loss(x.requires_grad_(True), x.clone(), torch.tensor(10.0))


def test_step_descriptor_structs_match_the_library():
    """The ctypes twins of mrclip_step / mrclip_peer (mrclip_b200/step.py) have the size the library was compiled with,
    the step entries reject a NULL / inconsistent descriptor with an error code (no crash, no compute on CPU), and the
    pure planning entry answers without a GPU."""
    from mrclip_b200.step import Peer, Step
    lib = _cabi.load()
    assert ctypes.sizeof(Step) == lib.mrclip_step_struct_bytes()
    assert ctypes.sizeof(Peer) == lib.mrclip_peer_struct_bytes()
    assert lib.mrclip_peer_block_bytes() == 4096 and lib.mrclip_step_small_floats() >= 128 + 64 * 64
    assert lib.mrclip_step_forward(None, None, 0, 0, None, 0, 0, None, None, 0, 0, None, None) < 0
    assert b"NULL descriptor" in lib.mrclip_last_error()
    st = Step()
    st.shape = _cabi.Shape(4096, 32768, 768, 3 * 4096)
    st.ld, st.kind, st.local_loss = 768, 0, 1
    st.peer = Peer(8, 3, None, None, None, None, None, None, None, 1, None, None, None)
    assert lib.mrclip_step_uses_fwd_ds(ctypes.byref(st)) == 1          # large multi-rank local loss: forward-side d scale
    st.local_loss = 0
    assert lib.mrclip_step_uses_fwd_ds(ctypes.byref(st)) == 0          # global loss: entropy sums
    st.local_loss, st.shape = 1, _cabi.Shape(512, 4096, 512, 3 * 512)
    assert lib.mrclip_step_uses_fwd_ds(ctypes.byref(st)) == 0          # n*N < 2^22
    st.shape = _cabi.Shape(4096, 32768, 768, 0)                        # label_offset must be rank * n
    assert lib.mrclip_step_backward(ctypes.byref(st), None, None, 0.0, None, 0, 0, None, 0, 0, None, None, None) < 0
