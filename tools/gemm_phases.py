"""Print per-phase clock64 deltas of single GEMM launches (debug aid, GPU only)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opentransformer_b200 import ops, _lib
dev = torch.device('cuda:0')
L = _lib.lib()
buf = torch.zeros(148 * 8 + 148 * 2 * 64 + 148 * 16, dtype=torch.int64, device=dev)
names = ['entry', 'setup_done', 'first_tma_issued', 'first_full', 'mma_committed', 'tfull_seen', 'epi_done', 'exit']


def run(tag, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    L.otb_debug_gemm_timing(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    L.otb_debug_gemm_timing(None)
    kbt = buf[148 * 8:148 * 8 + 148 * 2 * 64].view(148, 2, 64).cpu()
    epi = buf[148 * 8 + 148 * 2 * 64:].view(148, 16).cpu()
    b = buf[:148 * 8].view(148, 8).cpu()
    live = b[:, 0] > 0
    b = b[live]
    d = (b - b[:, :1]).float()
    print(f'{tag}: {e0.elapsed_time(e1)*1e3:.1f} us by events, {int(live.sum())} CTAs; cycles since entry (mean over CTAs / max):')
    for i, n in enumerate(names):
        print(f'    {n:18s} {d[:, i].mean():9.0f} {d[:, i].max():9.0f}')
    t0 = int(buf[0])
    iss = [int(v) - t0 for v in kbt[0, 0] if int(v) > 0][:20]
    ful = [int(v) - t0 for v in kbt[0, 1] if int(v) > 0][:20]
    print('    CTA0 TMA issue  times:', iss)
    print('    CTA0 full seen  times:', ful)
    ev = epi[epi[:, 6] > 0][:, :7].float()
    if len(ev):
        dd = (ev[:, 1:] - ev[:, :-1]).mean(0).tolist()
        print('    2nd-tile epilogue phases (cycles, mean over CTAs): setup+bar %.0f | wait accumulator %.0f | tmem->math->stage %.0f | bar %.0f | copy-out %.0f | bar %.0f  (total %.0f)' % (*dd, float((ev[:, 6] - ev[:, 0]).mean())))


def mk(M, N, K, epi, **kw):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N * (2 if epi == ops.EPI_GLU else 1), K, device=dev).to(torch.bfloat16)
    b = torch.randn(w.shape[0], device=dev)
    extra = {}
    if epi in (ops.EPI_RESID_LN, ops.EPI_RESID):
        extra['resid'] = torch.randn(M, N, device=dev).to(torch.bfloat16)
    if epi == ops.EPI_RESID_LN:
        extra['gamma'] = torch.ones(N, device=dev); extra['beta'] = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    return lambda: ops.linear(a, w, b, epi, out=out, **extra)


for (M, N, K) in [(7968, 256, 2048), (7968, 2048, 2048), (7968, 768, 1024), (128, 64, 2048), (128, 256, 2048)]:
    run(f'BIAS M={M} N={N} K={K} (per-k-block cycles = (mma_committed - first_full) / {K // 64})', mk(M, N, K, ops.EPI_BIAS))
run('decode out-proj  M=320 N=256 K=256  RESID_LN', mk(320, 256, 256, ops.EPI_RESID_LN))
run('decode w2        M=320 N=256 K=2048 RESID_LN', mk(320, 256, 2048, ops.EPI_RESID_LN))
run('decode qkv       M=320 N=768 K=256  BIAS    ', mk(320, 768, 256, ops.EPI_BIAS))
run('decode w1 glu    M=320 N=2048 K=256 GLU     ', mk(320, 2048, 256, ops.EPI_GLU))
run('encoder w1 glu   M=7968 N=2048 K=256 GLU    ', mk(7968, 2048, 256, ops.EPI_GLU))
run('encoder qkv      M=7968 N=768 K=256 BIAS    ', mk(7968, 768, 256, ops.EPI_BIAS))
run('encoder out-proj M=7968 N=256 K=256 RESID_LN', mk(7968, 256, 256, ops.EPI_RESID_LN))
run('encoder w2       M=7968 N=256 K=2048 RESID_LN', mk(7968, 256, 2048, ops.EPI_RESID_LN))
