"""Run a few eager (non-graph) beam-decode steps at cfg-3 shapes -- target for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from opentransformer_b200.model import SpeechToText
from opentransformer_b200.recognize import BeamDecoder

dev = torch.device('cuda:0')
params = bench.model_params()
params['encoder']['n_blocks'] = 1
torch.manual_seed(0)
model = SpeechToText(params).eval().to(dev)
with torch.no_grad():
    model.decoder.output_layer.bias[1] = -1e4
x, mask = bench.synthetic_batch(32, 0)
x, mask = x.to(dev), mask.to(dev)
with torch.no_grad():
    mem, lens, B, T2 = model.encode_bf16(x, mask)
    bd = BeamDecoder(model.decoder, B, 10, T2, 60, dev, use_graph=False)
    bd.keep_logp = False
    bd.setup(mem, lens)
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        bd.step()
torch.cuda.synchronize()
print('done')
