"""GPU parity of the training step against the oracle's train step
(oracle/train_step.py, pinned to the real reference's trainer by tests/golden/train_step_postnorm_glu.pt):
loss and every parameter gradient within the stated bf16 tolerance, clip + scheduler + Adam update, autograd seam."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from opentransformer_b200 import train
    from opentransformer_b200.model import SpeechToText
    DEV = torch.device('cuda:0')

from oracle import train_step as ot

REL_L2_GRAD = 4e-2        # per-parameter gradient, bf16 activations / gradients vs the fp32 oracle
REL_L2_GRAD_ALL = 2e-2    # all gradients concatenated


def _params(n_enc=2, n_dec=2, tied=True):
    return {'type': 'speech2text', 'frontend_type': 'conv', 'encoder_type': 'transformer', 'decoder_type': 'transformer',
            'frontend': dict(input_size=80, output_size=256, in_channel=1, mid_channel=64, out_channel=128,
                             kernel_size=[[3, 3], [3, 3]], stride=[2, 2], dropout=0.0, act_func_type='relu',
                             front_end_layer_norm=False),
            'encoder': dict(d_model=256, n_heads=4, d_ff=2048, n_blocks=n_enc, pos_dropout=0.0, slf_attn_dropout=0.0,
                            ffn_dropout=0.0, residual_dropout=0.0, normalize_before=False, concat_after=False,
                            activation='glu', relative_positional=False),
            'decoder': dict(vocab_size=4234, d_model=256, n_heads=4, d_ff=2048, memory_dim=256, n_blocks=n_dec,
                            pos_dropout=0.0, slf_attn_dropout=0.0, src_attn_dropout=0.0, ffn_dropout=0.0,
                            residual_dropout=0.0, activation='glu', normalize_before=False, concat_after=False,
                            share_embedding=tied),
            'ctc_weight': 0.0, 'smoothing': 0.1}


def _build(params, seed=1234):
    torch.manual_seed(seed)
    model = SpeechToText(params)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'norm' in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith('.bias'):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        model.decoder.embedding.weight.mul_(0.12)      # keeps the logits O(1): the loss is then sensitive to every layer
    sd = {}
    for part in ('frontend', 'encoder', 'decoder'):
        for k, v in getattr(model, part).state_dict().items():
            sd[f'{part}.{k}'] = v.detach().clone().float()
    return model.to(DEV), sd


def _batch(B=3, T=200, L=12, lens=(200, 160, 181), seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, 80, generator=g)
    lens = torch.tensor(lens)
    mask = torch.arange(T)[None] < lens[:, None]
    x = x * mask.unsqueeze(2)
    tgt = torch.randint(3, 4234, (B, L), generator=g)
    tgt[:, 0] = 1
    tgt[0, L - 3:] = torch.tensor([1, 0, 0])
    tgt[1, L - 1] = 1
    tgt[2, L // 2:] = torch.tensor([1] + [0] * (L - L // 2 - 1))
    return x, mask, tgt


def _rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-20))


@pytest.mark.parametrize('tied', [True, False])
def test_gradients_match_oracle(tied):
    params = _params(tied=tied)
    model, sd = _build(params)
    x, mask, tgt = _batch()
    loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params)
    with torch.no_grad():
        loss, g = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV))
    names = [n for n, _ in model.named_parameters()]
    assert set(names) == set(g_ref) == set(g), (set(names) ^ set(g_ref), set(names) ^ set(g))
    worst = []
    for n in names:
        assert g[n].shape == g_ref[n].shape, n
        assert torch.isfinite(g[n]).all(), n
        worst.append((_rel(g[n], g_ref[n]), n))
    worst.sort(reverse=True)
    cat = torch.cat([g[n].reshape(-1).cpu() for n in names]), torch.cat([g_ref[n].reshape(-1) for n in names])
    r_all = _rel(*cat)
    print(f'loss gpu {float(loss):.5f} oracle {float(loss_ref):.5f}; all grads rel_l2 {r_all:.3e}; worst: '
          + ', '.join(f'{n} {r:.2e}' for r, n in worst[:5]))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    assert r_all < REL_L2_GRAD_ALL
    assert worst[0][0] < REL_L2_GRAD, worst[:3]


@pytest.mark.parametrize('which', ['both', 'encoder_only', 'decoder_only'])
def test_prenorm_gradients_match_oracle(which):
    """normalize_before=True as the reference writes it (encoder/transformer.py:41-63, decoder/transformer.py:54-90,151-152):
    the residual is taken AFTER the norm and the stack ends with one more LayerNorm.  Loss and every parameter gradient
    (the extra `encoder.norm` / `decoder.after_norm` included) against the oracle's autograd."""
    params = _params(tied=True)
    params['encoder']['normalize_before'] = which in ('both', 'encoder_only')
    params['decoder']['normalize_before'] = which in ('both', 'decoder_only')
    model, sd = _build(params)
    x, mask, tgt = _batch()
    loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params)
    with torch.no_grad():
        loss, g = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV))
    names = [n for n, _ in model.named_parameters()]
    assert set(names) == set(g_ref) == set(g), (set(names) ^ set(g_ref), set(names) ^ set(g))
    if which != 'decoder_only':
        assert 'encoder.norm.weight' in g
    if which != 'encoder_only':
        assert 'decoder.after_norm.weight' in g
    worst = sorted(((_rel(g[n], g_ref[n]), n) for n in names), reverse=True)
    r_all = _rel(torch.cat([g[n].reshape(-1).cpu() for n in names]), torch.cat([g_ref[n].reshape(-1) for n in names]))
    print(f'pre-norm ({which}): loss gpu {float(loss):.5f} oracle {float(loss_ref):.5f}; all grads rel_l2 {r_all:.3e}; worst: '
          + ', '.join(f'{n} {r:.2e}' for r, n in worst[:4]))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    assert r_all < REL_L2_GRAD_ALL
    assert worst[0][0] < REL_L2_GRAD, worst[:3]


@pytest.mark.parametrize('prenorm', [False, True])
def test_relu_feed_forward_gradients_match_oracle(prenorm):
    """activation='relu' (the reference's default, ffn.py:38-41): ReLU fused into the w_1 GEMM, backward from the hidden
    activations alone; post- and pre-norm."""
    params = _params(tied=False)
    for part in ('encoder', 'decoder'):
        params[part]['activation'] = 'relu'
        params[part]['normalize_before'] = prenorm
    model, sd = _build(params)
    x, mask, tgt = _batch()
    loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params)
    with torch.no_grad():
        loss, g = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV))
    names = [n for n, _ in model.named_parameters()]
    assert set(names) == set(g_ref) == set(g)
    worst = sorted(((_rel(g[n], g_ref[n]), n) for n in names), reverse=True)
    r_all = _rel(torch.cat([g[n].reshape(-1).cpu() for n in names]), torch.cat([g_ref[n].reshape(-1) for n in names]))
    print(f'relu ffn (prenorm={prenorm}): loss gpu {float(loss):.5f} oracle {float(loss_ref):.5f}; all grads rel_l2 {r_all:.3e}; '
          f'worst: ' + ', '.join(f'{n} {r:.2e}' for r, n in worst[:3]))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    assert r_all < REL_L2_GRAD_ALL
    # ReLU's derivative is a step: hidden units whose pre-activation lies within the bf16 noise of zero take the other branch
    # than in the fp32 oracle (measured: ~6e-2 on the decoder's w_1 with 36 target rows, and 4.2e-2 on the query projection
    # upstream of it): 1e-1 for w_1, 6e-2 for the other parameters of this test (the GLU tests keep 4e-2 everywhere)
    for r, n in worst:
        assert r < (1e-1 if 'feed_forward.w_1' in n else 6e-2), (n, r)


def test_autograd_seam_fills_param_grads():
    """model.train(); loss, _ = model(inputs, targets); loss.backward()  -- the reference's calling convention."""
    params = _params(n_enc=1, n_dec=1)
    model, sd = _build(params)
    x, mask, tgt = _batch()
    model.train()
    loss, aux = model({'inputs': x.to(DEV), 'mask': mask.to(DEV)}, {'targets': tgt.to(DEV), 'targets_length': None})
    assert aux is None and loss.requires_grad
    (loss * 0.5).backward()
    _, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params)
    for n, p in model.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n
    r = _rel(torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()]),
             0.5 * torch.cat([g_ref[n].reshape(-1) for n, _ in model.named_parameters()]))
    assert r < REL_L2_GRAD_ALL
    model.eval()
    with torch.no_grad():
        l2, _ = model({'inputs': x.to(DEV), 'mask': mask.to(DEV)}, {'targets': tgt.to(DEV), 'targets_length': None})
    assert abs(float(l2) - float(loss)) < 2e-2 * abs(float(loss))


@pytest.mark.parametrize('use_graph', [False, True])
def test_fused_trainer_step_matches_oracle_step(use_graph):
    """Two optimizer steps with gradient accumulation over 2 micro-batches each: weights after clip + lr(3), lr(4) +
    Adam must follow the oracle's (= the reference trainer's) update."""
    params = _params(n_enc=1, n_dec=1)
    model, sd = _build(params)
    model.train()
    tr = train.FusedTrainer(model, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6, clip_grad=0.05, model_size=256,
                            warmup_steps=12000, accum_steps=2, use_graph=use_graph)
    names = [n for n, _ in model.named_parameters()]
    w = {n: sd[n].clone() for n in names}
    m = {n: torch.zeros_like(w[n]) for n in names}
    v = {n: torch.zeros_like(w[n]) for n in names}
    for opt_step in (1, 2):
        acc = None
        for micro in range(2):
            x, mask, tgt = _batch(seed=10 * opt_step + micro)
            loss = tr.step(x.to(DEV), mask.to(DEV), tgt.to(DEV))
            sd_now = dict(sd)
            sd_now.update({n: w[n] for n in names})
            if params['decoder']['share_embedding']:
                sd_now['decoder.output_layer.weight'] = w['decoder.embedding.weight']
            loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd_now, params)
            assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
            acc = {n: g_ref[n] / 2 for n in names} if acc is None else {n: acc[n] + g_ref[n] / 2 for n in names}
        lr = ot.transformer_lr(ot.first_step_index() + opt_step - 1, 256, 12000)
        assert abs(lr - tr.lr()) < 1e-12
        ot.clip_and_adam(w, acc, m, v, opt_step, lr, 0.05)
        torch.cuda.synchronize()
        num = sum(float(((p.detach().cpu() - w[n]) ** 2).sum()) for n, p in model.named_parameters())
        den = sum(float(((w[n] - sd[n]) ** 2).sum()) for n in names)
        print(f'optimizer step {opt_step}: ||w_gpu - w_oracle|| / ||w_oracle - w_init|| = {(num / den) ** 0.5:.3e}')
        assert (num / den) ** 0.5 < 0.3       # Adam normalises every coordinate to ~lr*sign(g): bf16 noise flips the sign of near-zero gradients only
    # parameters are views into the flat buffer; state_dict keys / shapes unchanged (checkpoint format, speech2text.py:71-87)
    assert set(k for k in model.state_dict()) >= set(names)


def test_residual_dropout_gradients_with_replayed_masks():
    """residual_dropout = 0.1 as shipped (egs/aishell/conf/transformer_baseline.yaml): the CUDA path draws counter-based
    masks in the residual GEMM epilogues and replays them in the backward.  torch's Philox masks cannot be reproduced,
    so the oracle is run with the PRODUCT's masks (exported per dropout site): loss and every gradient must then match
    like in the deterministic case, the masks must have the right rate, differ between sites and between seeds, and a
    shipped-config model must train through the public seam."""
    from oracle import speech_model as om
    from opentransformer_b200 import ops
    params = _params(n_enc=2, n_dec=2)
    params['encoder']['residual_dropout'] = 0.1
    params['decoder']['residual_dropout'] = 0.1
    model, sd = _build(params)
    x, mask, tgt = _batch()
    B, T2, L, p = x.shape[0], ops.conv_geometry(x.shape[1], 80)[2], tgt.shape[1] - 1, 0.1
    seed = torch.tensor([12345], dtype=torch.int32, device=DEV)
    masks = {}
    for site, part, rate in train.dropout_sites(model):
        assert rate == p
        rows = B * T2 if part == 'encoder' else B * L
        m = ops.dropout_mask(rows, 256, p, seed, site).cpu()
        i = (site - train.ENC_SITE) // 2 if part == 'encoder' else (site - train.DEC_SITE) // 3
        k = (site - train.ENC_SITE) % 2 if part == 'encoder' else (site - train.DEC_SITE) % 3
        masks[(f'{part}.blocks.{i}', k)] = m.view(B, -1, 256).float()
    keep = torch.stack([m.mean() for m in masks.values()])
    assert float((keep - 0.9).abs().max()) < 0.01, keep.tolist()
    ms = list(masks.values())
    assert not torch.equal(ms[0], ms[1]), 'different sites must draw different masks'
    seed2 = torch.tensor([12346], dtype=torch.int32, device=DEV)
    assert not torch.equal(ops.dropout_mask(B * T2, 256, p, seed2, 0).cpu().view(B, -1, 256).float(), ms[0])
    om.set_dropout(lambda site, t: t * masks[site] / (1.0 - p))
    try:
        loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params)
    finally:
        om.set_dropout(None)
    loss_nodrop, _ = ot.loss_and_grads(x, mask, tgt, sd, params)
    with torch.no_grad():
        loss, g = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV), drop_seed=seed)
    names = [n for n, _ in model.named_parameters()]
    worst = sorted(((_rel(g[n], g_ref[n]), n) for n in names), reverse=True)
    r_all = _rel(torch.cat([g[n].reshape(-1).cpu() for n in names]), torch.cat([g_ref[n].reshape(-1) for n in names]))
    print(f'residual dropout 0.1, replayed masks: loss gpu {float(loss):.5f} oracle {float(loss_ref):.5f} (without dropout '
          f'{float(loss_nodrop):.5f}); all grads rel_l2 {r_all:.3e}; worst ' + ', '.join(f'{n} {r:.2e}' for r, n in worst[:3]))
    assert abs(float(loss_ref) - float(loss_nodrop)) > 1e-4, 'the replayed masks must actually change the network'
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    assert r_all < REL_L2_GRAD_ALL and worst[0][0] < REL_L2_GRAD, worst[:3]
    # public seam: model.train(); loss.backward() -- two calls draw different masks (the seed advances on the device)
    l1, _ = model({'inputs': x.to(DEV), 'mask': mask.to(DEV)}, {'targets': tgt.to(DEV), 'targets_length': None})
    l2, _ = model({'inputs': x.to(DEV), 'mask': mask.to(DEV)}, {'targets': tgt.to(DEV), 'targets_length': None})
    l1.backward()
    assert float(l1) != float(l2) and all(p_.grad is not None for p_ in model.parameters())
    # FusedTrainer with the CUDA graph: replays must not freeze the mask
    tr = train.FusedTrainer(model, accum_steps=1, use_graph=True)
    losses = [float(tr.step(x.to(DEV), mask.to(DEV), tgt.to(DEV))) for _ in range(4)]
    print('FusedTrainer losses with dropout (graph from the 2nd step on):', losses)
    assert len(set(losses)) == 4


def test_other_dropout_rates_raise_in_training():
    params = _params(n_enc=1, n_dec=1)
    params['encoder']['ffn_dropout'] = 0.1
    model, _ = _build(params)
    x, mask, tgt = _batch()
    with pytest.raises(NotImplementedError):
        train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV))


def test_joint_ctc_loss_and_gradients_match_oracle():
    """ctc_weight 0.3 (model/speech2text.py:30-36,60-72): (1 - w) * attention + w * CTC; the CTC head's gradients and its
    contribution to the encoder's, against the oracle (pinned to the reference by tests/golden/joint_ctc_postnorm_glu.pt)."""
    params = _params(n_enc=2, n_dec=1)
    params['ctc_weight'] = 0.3
    params['encoder_output_size'] = 256
    model, sd = _build(params)
    for k, v in model.assistor.state_dict().items():
        sd[f'assistor.{k}'] = v.detach().clone().float().cpu()
    x, mask, tgt = _batch()
    tlen = torch.tensor([(int((tgt[b, 1:] != 0).sum())) for b in range(tgt.shape[0])], dtype=torch.int32)
    loss_ref, g_ref = ot.loss_and_grads(x, mask, tgt, sd, params, truth_length=tlen)
    with torch.no_grad():
        loss, g, lc = train.forward_backward(model.train(), x.to(DEV), mask.to(DEV), tgt.to(DEV), truth_length=tlen.to(DEV),
                                             return_ctc=True)
    names = [n for n, _ in model.named_parameters()]
    assert set(names) == set(g_ref) == set(g), (set(names) ^ set(g_ref), set(names) ^ set(g))
    worst = sorted(((_rel(g[n], g_ref[n]), n) for n in names), reverse=True)
    r_all = _rel(torch.cat([g[n].reshape(-1).cpu() for n in names]), torch.cat([g_ref[n].reshape(-1) for n in names]))
    print(f'joint CTC: loss gpu {float(loss):.5f} oracle {float(loss_ref):.5f} (ctc part {float(lc):.5f}); all grads rel_l2 {r_all:.3e}; '
          'worst ' + ', '.join(f'{n} {r:.2e}' for r, n in worst[:3]))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    assert r_all < REL_L2_GRAD_ALL and worst[0][0] < REL_L2_GRAD, worst[:3]
    # public seam, eval mode: (loss, {'CTCLoss': ..}) like the reference
    model.eval()
    with torch.no_grad():
        l2, aux = model({'inputs': x.to(DEV), 'mask': mask.to(DEV)}, {'targets': tgt.to(DEV), 'targets_length': tlen})
    assert abs(float(l2) - float(loss_ref)) < 2e-2 * abs(float(loss_ref)) and 'CTCLoss' in aux
