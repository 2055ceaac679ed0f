"""Generator of tests/golden/spec_augment.pt: the REFERENCE's otrans.data.augment.spec_augment (data/augment.py:9-41) run in
the build container on seeded features, under fixed `random` / `numpy.random` seeds.

    PYTHONPATH=/root/reference python tests/golden/make_spec_augment.py [out.pt]

The fixture pins the RNG call order (np.random.uniform for the width, random.randint for the offset, frequency masks before
time masks) that opentransformer_b200/augment.py:draw_bands must reproduce.
"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get('OTRANS_REFERENCE', '/root/reference'))
from otrans.data.augment import spec_augment  # noqa: E402  (the real reference function)

LENS, SEED_FEAT, SEED_PY, SEED_NP = [300, 211, 257], 2, 11, 12


def main(out_path):
    g = torch.Generator().manual_seed(SEED_FEAT)
    feats = []
    for i, n in enumerate(LENS):
        x = torch.randn(300, 80, generator=g) + 3.0 if i == 0 else torch.randn(n, 80, generator=g) + 3.0
        feats.append(x[:n].clone())
    random.seed(SEED_PY)
    np.random.seed(SEED_NP)
    aug = [spec_augment(f.clone()) for f in feats]          # the reference masks in place and returns its argument
    torch.save({'lens': LENS, 'seed_py': SEED_PY, 'seed_np': SEED_NP, 'features': feats, 'augmented': aug}, out_path)
    return feats, aug


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, 'spec_augment.pt')
    main(out)
    print('wrote', out)
