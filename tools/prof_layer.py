"""Run the kernels of one encoder layer (cfg-2 shapes) a few times -- target for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from opentransformer_b200.model import SpeechToText

dev = torch.device('cuda:0')
params = bench.model_params()
params['encoder']['n_blocks'] = 1
params['decoder']['n_blocks'] = 1
torch.manual_seed(0)
model = SpeechToText(params).eval().to(dev)
x, mask = bench.synthetic_batch(32, 0)
x, mask = x.to(dev), mask.to(dev)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        model.encode_bf16(x, mask)
torch.cuda.synchronize()
print('done')
