"""SpecAugment for device-resident batches (otrans/data/augment.py:9-41; called per utterance by data/audio.py:136 with
the defaults 2 / 2 / 0.3 / 0.05 / 100).  The band positions are drawn on the host with exactly the reference's RNG call
order (np.random.uniform then random.randint per mask, frequency masks first), the zeroing runs as one CUDA kernel over
the whole batch (otb_spec_augment) instead of numpy slicing inside DataLoader workers."""
import ctypes
import random

import numpy as np
import torch

from . import _lib
from ._lib import check


def draw_bands(tau, v, freq_mask_num=2, time_mask_num=2, freq_mask_rate=0.3, time_mask_rate=0.05, max_mask_time_len=100):
    """[(f0, f)] * freq_mask_num + [(t0, t)] * time_mask_num for one utterance of `tau` valid frames x `v` bins."""
    freq_para = int(v * freq_mask_rate)
    time_para = min(int(tau * time_mask_rate), max_mask_time_len)
    bands = []
    for _ in range(freq_mask_num):
        f = int(np.random.uniform(low=0.0, high=freq_para))
        bands += [random.randint(0, v - f), f]
    for _ in range(time_mask_num):
        t = int(np.random.uniform(low=0.0, high=time_para))
        bands += [random.randint(0, tau - t), t]
    return bands


def spec_augment_(x, lengths, freq_mask_num=2, time_mask_num=2, **kw):
    """In place on x f32 [B,T,F] (CUDA); lengths = valid frames per utterance (the reference augments each utterance's
    un-padded feature matrix)."""
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError('spec_augment_: x must be a contiguous f32 CUDA tensor (no CPU fallback)')
    B, T, F = x.shape
    rows = [draw_bands(int(n), F, freq_mask_num, time_mask_num, **kw) for n in lengths]
    bands = torch.tensor(rows, dtype=torch.int32).to(x.device, non_blocking=True)
    check(_lib.lib().otb_spec_augment(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(bands.data_ptr()), B, T, F, freq_mask_num,
                                      time_mask_num, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'otb_spec_augment')
    return x
