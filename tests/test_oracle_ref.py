"""The oracle port against the REFERENCE ITSELF, where oracle/_ref exists (the reference's own Python modules, byte-compiled
by oracle/build_ref.py in the build container; git-ignored, travels with gpurun).  Complements tests/test_oracle_golden.py
(committed fixtures): here the reference code runs live on fresh random models.  Nothing reads /root/reference."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import beam_search as obs      # noqa: E402
from oracle import build_ref               # noqa: E402
from oracle import speech_model as om      # noqa: E402

pytestmark = pytest.mark.skipif(build_ref.load() is None, reason='oracle/_ref not built (no reference tree in this environment)')


def _params(normalize_before, activation):
    return {'type': 'speech2text', 'frontend_type': 'conv', 'encoder_type': 'transformer', 'decoder_type': 'transformer',
            'frontend': dict(input_size=20, output_size=32, in_channel=1, mid_channel=4, out_channel=8,
                             kernel_size=[[3, 3], [3, 3]], stride=[2, 2], dropout=0.0, act_func_type='relu',
                             front_end_layer_norm=False),
            'encoder': dict(d_model=32, n_heads=4, d_ff=48, n_blocks=2, pos_dropout=0.0, slf_attn_dropout=0.0,
                            ffn_dropout=0.0, residual_dropout=0.0, normalize_before=normalize_before, concat_after=False,
                            activation=activation, relative_positional=False),
            'decoder': dict(vocab_size=40, d_model=32, n_heads=4, d_ff=48, memory_dim=32, n_blocks=2, pos_dropout=0.0,
                            slf_attn_dropout=0.0, src_attn_dropout=0.0, ffn_dropout=0.0, residual_dropout=0.0,
                            activation=activation, normalize_before=normalize_before, concat_after=False,
                            share_embedding=True),
            'ctc_weight': 0.0, 'smoothing': 0.1}


@pytest.mark.parametrize('normalize_before,activation', [(False, 'glu'), (True, 'relu')])
def test_port_equals_live_reference(normalize_before, activation):
    from otrans.model import End2EndModel
    from otrans.recognize.speech2text import SpeechToTextRecognizer
    params = _params(normalize_before, activation)
    torch.manual_seed(4321)
    ref = End2EndModel['speech2text'](params).eval()
    with torch.no_grad():
        ref.decoder.embedding.weight.mul_(0.15)
        ref.decoder.output_layer.bias[1] = 1.5          # hypotheses end at different steps
    sd = {k: v.detach().clone().float() for k, v in ref.state_dict().items() if k.split('.')[0] in ('frontend', 'encoder', 'decoder')}
    g = torch.Generator().manual_seed(5)
    B, T, beam, max_len = 3, 64, 3, 9
    lens = torch.tensor([64, 50, 37])
    mask = torch.arange(T)[None] < lens[:, None]
    x = torch.randn(B, T, 20, generator=g) * mask.unsqueeze(2)
    with torch.no_grad():
        # encoder states: reference modules vs the port
        fx, fm = ref.frontend(x, mask)
        mem_ref, mm_ref, _ = ref.encoder(fx, fm)
        mem, mm = om.encode(x, mask, sd, params)
        assert torch.equal(mm, mm_ref)
        torch.testing.assert_close(mem[mm], mem_ref[mm_ref], rtol=1e-4, atol=2e-5)
        # the whole recognize(): ids bit-exact, scores to fp32 tolerance
        rec = SpeechToTextRecognizer(ref, beam_width=beam, nbest=beam, max_len=max_len, idx2unit={i: str(i) for i in range(40)},
                                     penalty=0.6, lamda=5, ngpu=0)
        hyps, scores_ref = rec.recognize(x, mask)
        nb, ns, _, _ = obs.recognize(x, mask, sd, params, beam=beam, nbest=beam, max_len=max_len, penalty=0.6, lamda=5)
    for b in range(B):
        for r in range(beam):
            port = [int(t) for t in nb[b, r].tolist()]
            port = port[:port.index(1)] if 1 in port else port
            want = [int(t) for t in hyps[b][r].split()] if hyps[b][r] else []
            assert port == want, (b, r, port, want)
    torch.testing.assert_close(ns, scores_ref.float(), rtol=1e-4, atol=1e-4)
