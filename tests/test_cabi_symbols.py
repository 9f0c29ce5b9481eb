"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO
from audiotokenization_b200 import _cabi


def declared_symbols():
    names = set()
    inc = os.path.join(REPO, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(bc_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _cabi.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert names == set(_cabi._SIGNATURES), (names ^ set(_cabi._SIGNATURES))
    assert lib.bc_abi_version() == 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_device_is_an_error_not_a_fallback():
    lib = _cabi.load_library()
    rc = lib.bc_device_info(0, None, None, None, None)
    assert rc == -3 and b"no CUDA device" in lib.bc_last_error()


def test_argument_validation_without_touching_the_device():
    lib = _cabi.load_library()
    assert lib.bc_conv1d_fwd(None, None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0, None) == -1
    assert b"null pointer" in lib.bc_last_error()
    assert lib.bc_snake_fwd(None, None, None, None, None, 1, 1, 1, 0, None) == -1
    assert lib.bc_vq_encode(None, None, None, None, None, None, None, 1, 1, 8, 8, None) == -1
    assert lib.bc_lstm_recurrent_fwd(None, None, None, None, None, 1, 1, 32, None) == -1
    assert lib.bc_lstm_workspace_bytes(3, 64) == 3 * 64 * 32 * 4   # h (two step parities) + c
    assert lib.bc_lstm_packed_whh_floats(64) == 4 * 64 * 64


def test_host_side_lstm_packer_matches_module_packing():
    from audiotokenization_b200.vq.module import _LSTMParams
    lib = _cabi.load_library()
    H = 32
    p = _LSTMParams(H, H, 1)
    w_hh = p.weight_hh_l0.detach().contiguous()
    out = torch.empty(4 * H * H)
    rc = lib.bc_lstm_pack_whh(w_hh.data_ptr(), out.data_ptr(), H)
    assert rc == 0
    _, _, w_rec = p.packed(0)
    assert torch.equal(out.view_as(w_rec), w_rec)


def test_cpu_tensor_raises():
    from audiotokenization_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.snake(torch.zeros(1, 4, 8), torch.ones(8), torch.ones(8))
    from audiotokenization_b200.vq import BigCodecEncoder
    from audiotokenization_b200 import configs
    enc = BigCodecEncoder(**configs.get_config("tiny")["codec_encoder"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 1, 400))
