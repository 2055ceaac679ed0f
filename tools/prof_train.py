"""Run eager training micro-steps (cfg-5 shapes, 2 encoder + 1 decoder layers) -- target for `ncu --set full` on the backward kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from opentransformer_b200 import train
from opentransformer_b200.model import SpeechToText

dev = torch.device('cuda:0')
params = bench.train_params()
params['encoder']['n_blocks'] = 2
params['decoder']['n_blocks'] = 1
torch.manual_seed(0)
model = SpeechToText(params).to(dev).train()
x, mask = bench.synthetic_batch(32, 0)
t = bench.synthetic_targets(32, 1)
x, mask, t = x.to(dev), mask.to(dev), t.to(dev)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        loss, g = train.forward_backward(model, x, mask, t)
torch.cuda.synchronize()
print('done', float(loss))
