"""Device fbank + CMVN (opentransformer_b200/features.py, csrc/fbank.cu) against the reference's feature extractor: the
reference calls torchaudio.compliance.kaldi.fbank(wavform, num_mel_bins, sample_frequency, dither=0.0) (data/audio.py:117-120)
and `normalization` (audio.py:22-24); torchaudio is its un-vendored pip dependency, so parity is anchored on that function."""
import math

import pytest
import torch

ta = pytest.importorskip('torchaudio')
kaldi = ta.compliance.kaldi


def test_mel_banks_and_window_equal_kaldi_definition():
    from opentransformer_b200.features import mel_banks, povey_window
    for nb in (40, 80):
        ref, _ = kaldi.get_mel_banks(nb, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
        bank, rng = mel_banks(nb, 512, 16000)
        torch.testing.assert_close(bank, ref, rtol=1e-3, atol=1e-5)      # mel points in double here, fp32 in torchaudio
        for f in range(nb):
            lo, hi = int(rng[f, 0]), int(rng[f, 1])
            assert float(bank[f, :lo].abs().sum()) == 0.0 and float(bank[f, hi:].abs().sum()) == 0.0 and hi > lo
    torch.testing.assert_close(povey_window(400), kaldi._feature_window_function('povey', 400, 0.42, torch.device('cpu'), torch.float32))


def _waves():
    g = torch.Generator().manual_seed(5)
    out = []
    for n, f0 in ((16000, 220.0), (23456, 997.0), (399, 50.0), (8123, 3300.0)):
        t = torch.arange(n) / 16000.0
        w = 0.3 * torch.sin(2 * math.pi * f0 * t) + 0.1 * torch.sin(2 * math.pi * 2.7 * f0 * t + 1.0) + 0.02 * torch.randn(n, generator=g) + 0.05
        out.append(w)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize('num_mel', [80, 40])
def test_fbank_kernel_matches_torchaudio_kaldi_fbank(num_mel):
    from opentransformer_b200.features import FbankExtractor
    waves = _waves()
    fe = FbankExtractor(num_mel_bins=num_mel, normalization=False)
    feats, lens, mask = fe(waves)
    assert lens.cpu().tolist() == [98, 145, 0, 49]
    for b, w in enumerate(waves):
        n = int(lens[b])
        if n == 0:
            assert float(feats[b].abs().max()) == 0.0 and not bool(mask[b].any())
            continue
        ref = kaldi.fbank(w.unsqueeze(0), num_mel_bins=num_mel, sample_frequency=16000.0, dither=0.0)
        assert ref.shape == (n, num_mel)
        got = feats[b, :n].cpu()
        err = float((got - ref).abs().max())
        print(f'fbank utt {b}: {n} frames, max |diff| of the log-mel energies {err:.2e} (range {float(ref.min()):.1f} .. {float(ref.max()):.1f})')
        torch.testing.assert_close(got, ref, rtol=1e-3, atol=2e-3)
        assert float(feats[b, n:].abs().max()) == 0.0 if n < feats.shape[1] else True
        assert mask[b].sum().item() == n


@pytest.mark.gpu
def test_utterance_and_global_cmvn():
    from opentransformer_b200.features import FbankExtractor
    waves = _waves()
    raw, lens, _ = FbankExtractor(num_mel_bins=80, normalization=False)(waves)
    norm, _, _ = FbankExtractor(num_mel_bins=80, normalization=True)(waves)
    for b in range(len(waves)):
        n = int(lens[b])
        if n == 0:
            continue
        x = raw[b, :n].cpu()
        std, mean = torch.std_mean(x)                     # data/audio.py:22-24
        torch.testing.assert_close(norm[b, :n].cpu(), (x - mean) / std, rtol=1e-4, atol=1e-4)
        assert float(norm[b, n:].abs().max()) == 0.0 if n < norm.shape[1] else True
    gm, gs = torch.linspace(-1, 1, 80), torch.linspace(0.5, 2.0, 80)
    gn, _, _ = FbankExtractor(num_mel_bins=80, normalization=True, global_mean=gm, global_std=gs)(waves)
    n = int(lens[1])
    torch.testing.assert_close(gn[1, :n].cpu(), (raw[1, :n].cpu() - gm) / gs, rtol=1e-5, atol=1e-5)
