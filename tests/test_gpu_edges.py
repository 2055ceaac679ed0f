"""Edge cases of the hot path on the GPU: smallest geometries, single utterance, very ragged batches, extreme beam / length
settings, error behaviour of the C ABI (status codes -> RuntimeError, never an abort, never a silent CPU path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from opentransformer_b200 import ops
    from opentransformer_b200.model import SpeechToText
    from opentransformer_b200.recognize import SpeechToTextRecognizer
    DEV = torch.device('cuda:0')

from oracle import beam_search as obs
from oracle import speech_model as om
from tests.test_gpu_model import _params, _build, _batch, _rel, _valid, REL_L2_STATES


def test_minimal_and_single_utterance_geometries():
    """T = 7 is the shortest input the two stride-2 convs accept (T' = 1); B = 1; F = 40 (the YAML as shipped)."""
    params = _params(n_enc=1, n_dec=1, F=40)
    model, sd = _build(params)
    for B, T, lens in [(1, 7, [7]), (1, 8, [8]), (2, 23, [23, 9]), (1, 1000, [1000])]:
        x, mask = _batch(B, T, 40, lens)
        mem_ref, mmask = om.encode(x, mask, sd, params)
        with torch.no_grad():
            mem, l, _, T2 = model.encode_bf16(x.to(DEV), mask.to(DEV))
        assert T2 == mem_ref.shape[1] and l.cpu().tolist() == mmask.sum(1).tolist()
        r = _rel(_valid(mem.float().view(B, T2, -1).cpu(), mmask), _valid(mem_ref, mmask))
        assert r < REL_L2_STATES, (B, T, r)


def test_very_ragged_batch_and_short_memories_in_beam_search():
    """One utterance has a single valid encoder frame, another the full length: masks, cross-attention over 1 key,
    and the search on a batch whose utterances end at different steps."""
    params = _params(n_enc=1, n_dec=2)
    model, sd = _build(params)
    with torch.no_grad():
        model.decoder.output_layer.bias[1] = 1.0
    sd['decoder.output_layer.bias'] = model.decoder.output_layer.bias.detach().float().cpu().clone()
    B, beam, max_len = 4, 3, 7
    x, mask = _batch(B, 120, 80, [120, 11, 64, 7])
    nb_ref, ns_ref, _, _ = obs.recognize(x, mask, sd, params, beam=beam, nbest=2, max_len=max_len, penalty=0.6, lamda=5)
    for persistent in (False, True):
        rec = SpeechToTextRecognizer(model, beam_width=beam, nbest=2, max_len=max_len, penalty=0.6, lamda=5, ngpu=1,
                                     persistent=persistent)
        p, s, n = rec.recognize_ids(x.to(DEV), mask.to(DEV))
        assert torch.isfinite(s).all() and p.shape[:2] == (B, 2) and n <= max_len
        torch.testing.assert_close(s.cpu(), ns_ref, rtol=3e-2, atol=0.3)


@pytest.mark.parametrize('beam,max_len', [(1, 3), (16, 4), (2, 128)])
def test_extreme_beam_and_length_settings(beam, max_len):
    params = _params(n_enc=1, n_dec=1)
    model, sd = _build(params)
    x, mask = _batch(2, 64, 80, [64, 40])
    rec = SpeechToTextRecognizer(model, beam_width=beam, nbest=1, max_len=max_len, penalty=0.0, lamda=5, ngpu=1)
    p, s, n = rec.recognize_ids(x.to(DEV), mask.to(DEV))
    assert n == max_len and p.shape == (2, 1, max_len) and torch.isfinite(s).all()
    assert int(p.min()) >= 0 and int(p.max()) < 4234


def test_errors_are_python_exceptions_not_aborts():
    a = torch.randn(8, 64).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.linear(a, a)                                           # CPU tensors
    ad = a.to(DEV)
    with pytest.raises(RuntimeError, match='K must be a multiple of 8|aligned'):
        ops.linear(ad[:, :60].contiguous(), ad[:, :60].contiguous())
    with pytest.raises(RuntimeError, match='LayerNorm'):
        ops.linear(ad, ad, None, ops.EPI_RESID_LN, resid=ad[:, :8].contiguous())   # missing gamma / beta
    with pytest.raises(RuntimeError, match='beam'):
        st = ops.BeamState(1, 17, 4, DEV)
        st.step(torch.zeros(17, 32, device=DEV), 32)
    with pytest.raises((RuntimeError, ValueError)):
        ops.attention(ad, ad, ad, 1, 1, 8, 8, q_col0=4)            # misaligned column offset
    torch.cuda.synchronize()                                       # the context is still healthy
    assert float(ops.linear(ad, ad).float().abs().sum()) > 0


def test_training_on_ragged_batch_with_all_pad_tail():
    """Targets whose tails are PAD (ignored by the loss) and inputs of very different lengths: finite loss / gradients,
    zero gradient for the embedding rows of unused tokens except through the tied output layer."""
    from opentransformer_b200 import train
    from tests.test_gpu_train import _params as tparams, _build as tbuild
    params = tparams(n_enc=1, n_dec=1, tied=False)
    model, sd = tbuild(params)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 96, 80, generator=g)
    lens = torch.tensor([96, 15, 50])
    mask = torch.arange(96)[None] < lens[:, None]
    x = x * mask.unsqueeze(2)
    tgt = torch.tensor([[1, 5, 6, 7, 1, 0, 0, 0], [1, 9, 1, 0, 0, 0, 0, 0], [1, 4, 4, 4, 4, 4, 4, 1]])
    with torch.no_grad():
        loss, grads = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV))
    assert torch.isfinite(loss)
    for n, gr in grads.items():
        assert torch.isfinite(gr).all(), n
    used = torch.unique(tgt[:, :-1])
    emb_g = grads['decoder.embedding.weight'].cpu()
    unused = torch.ones(4234, dtype=torch.bool)
    unused[used] = False
    assert float(emb_g[unused].abs().max()) == 0.0 and float(emb_g[used].abs().max()) > 0
