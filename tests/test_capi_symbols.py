"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU driver and exports
every symbol include/otb200.h declares.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def libpath():
    from opentransformer_b200 import build
    return build.build()


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'otb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(otb_[a-z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/otb200.h but not exported'


def test_python_binding_covers_the_header():
    from opentransformer_b200 import exported_symbols
    assert exported_symbols() == _header_symbols()


def test_version_and_error_reporting(libpath):
    from opentransformer_b200 import _lib
    lib = _lib.lib()
    assert lib.otb_version() == 1
    t1 = ctypes.c_int()
    assert lib.otb_conv_geometry(1000, 80, t1, None, None, None) == 0 and t1.value == 499
    assert lib.otb_conv_geometry(3, 80, None, None, None, None) != 0       # too short -> status, not abort
    assert b'otb_conv_geometry' in lib.otb_last_error()


def test_conv_geometry_matches_reference_formula(libpath):
    from opentransformer_b200 import ops
    for T, F in [(1000, 80), (1000, 40), (777, 83), (90, 20), (7, 1)]:
        t1, f1, t2, f2 = ops.conv_geometry(T, F)
        assert t1 == (T - 3) // 2 + 1 and t2 == (t1 - 3) // 2 + 1          # frontend/conv.py:10-11, pad 0
        assert f1 == (F + 2 - 3) // 2 + 1 and f2 == (f1 + 2 - 3) // 2 + 1  # pad 1


def test_persistent_decode_host_arithmetic(libpath):
    """Host-only parts of the persistent decode entry points: workspace size (no GPU needed) and argument checks."""
    from opentransformer_b200 import _lib
    lib = _lib.lib()
    ws = lib.otb_decode_persistent_workspace
    base = ws(320, 6, 60, 32, 10, 4234)                 # benchmark geometry: 32 utterances x beam 10, 3 row groups
    assert base > 0
    groups, ldv = 3, (4234 + 31) // 32 * 32
    # dominated by the fp32 logits (128 rows per group) and the 16 fp32 feed-forward partial products per group
    assert base >= groups * 128 * ldv * 4 + 16 * groups * 128 * 256 * 4
    assert ws(640, 6, 60, 64, 10, 4234) > base          # more utterances -> more row groups
    assert ws(320, 6, 60, 32, 10, 8000) > base          # larger vocabulary -> larger logits buffer
    assert ws(320, 6, 60, 32, 17, 4234) == -1           # beam > 16 is not supported by the kernel
    assert ws(0, 6, 60, 32, 10, 4234) == -1
    assert lib.otb_set_decode_barrier(2) != 0 and b'otb_set_decode_barrier' in lib.otb_last_error()
    for kind in (1, 0, -1):
        assert lib.otb_set_decode_barrier(kind) == 0


def test_product_refuses_cpu_tensors(libpath):
    import torch
    from opentransformer_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.linear(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))


def test_state_dict_keys_match_reference_names():
    """Key names / shapes of the shipped transformer config (probed from the reference)."""
    from opentransformer_b200.model import SpeechToText
    params = {'frontend_type': 'conv', 'encoder_type': 'transformer', 'decoder_type': 'transformer',
              'frontend': dict(input_size=80, output_size=256, in_channel=1, mid_channel=64, out_channel=128,
                               kernel_size=[[3, 3], [3, 3]], stride=[2, 2], dropout=0.0, act_func_type='relu',
                               front_end_layer_norm=False),
              'encoder': dict(d_model=256, n_heads=4, d_ff=2048, n_blocks=2, pos_dropout=0.0, slf_attn_dropout=0.0,
                              ffn_dropout=0.0, residual_dropout=0.1, normalize_before=False, concat_after=False,
                              activation='glu', relative_positional=False),
              'decoder': dict(vocab_size=4234, d_model=256, n_heads=4, d_ff=2048, memory_dim=256, n_blocks=1,
                              pos_dropout=0.0, slf_attn_dropout=0.0, src_attn_dropout=0.0, ffn_dropout=0.0,
                              residual_dropout=0.1, activation='glu', normalize_before=False, concat_after=False,
                              share_embedding=True),
              'ctc_weight': 0.0, 'smoothing': 0.1}
    m = SpeechToText(params)
    enc, dec, fe = m.encoder.state_dict(), m.decoder.state_dict(), m.frontend.state_dict()
    assert tuple(enc['blocks.1.slf_attn.qvk_proj.weight'].shape) == (768, 256)
    assert tuple(enc['blocks.0.feed_forward.w_1.weight'].shape) == (4096, 256)
    assert tuple(dec['blocks.0.src_attn.vk_proj.weight'].shape) == (512, 256)
    assert tuple(fe['conv2.conv_layer.weight'].shape) == (128, 64, 3, 3)
    assert tuple(fe['output_layer.weight'].shape) == (256, 2560)
    assert dec['embedding.weight'].data_ptr() == dec['output_layer.weight'].data_ptr()   # tied
