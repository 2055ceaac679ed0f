"""Generate golden fixtures by running the REAL reference (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so these fixtures --
seeded inputs, the reference-initialised weights and the reference's own outputs -- are what pins
oracle/ to the reference.  Model sizes are deliberately tiny so that weights + activations fit in
a few hundred KB; every code path of the hot path is exercised (post-/pre-norm, relu/glu,
abs/rel-pos transformer encoder, conformer, decoder, loss, beam search with per-step traces).
"""
import copy
import os
import sys

sys.dont_write_bytecode = True
sys.path[:0] = ['/root/reference', '/root/reference/otrans/module']  # non-package import in ffn.py:9

import torch  # noqa: E402
import yaml  # noqa: E402
from otrans.model import End2EndModel  # noqa: E402
from otrans.recognize.speech2text import SpeechToTextRecognizer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_params(encoder_type, **enc_over):
    base = 'conformer_baseline.yaml' if encoder_type == 'conformer' else 'transformer_baseline.yaml'
    with open('/root/reference/egs/aishell/conf/' + base) as f:
        p = yaml.load(f, Loader=yaml.FullLoader)['model']
    p = copy.deepcopy(p)
    d = 32
    p['frontend'].update(input_size=20, output_size=d, mid_channel=4, out_channel=8)
    if encoder_type == 'conformer':
        p['encoder'].update(d_model=d, d_ff=24, nblocks=2, n_heads=4, residual_dropout=0.0, cov_kernel_size=5)
    else:
        p['encoder'].update(d_model=d, d_ff=48, n_blocks=2, n_heads=4, residual_dropout=0.0)
    p['encoder'].update(enc_over)
    p['decoder'].update(vocab_size=40, d_model=d, d_ff=48, memory_dim=d, n_blocks=2, n_heads=4,
                        residual_dropout=0.0)
    return p


def make_batch(seed, b, t, f, lens):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, t, f, generator=g)
    lens = torch.tensor(lens)
    mask = torch.arange(t).unsqueeze(0) < lens.unsqueeze(1)
    x = x * mask.unsqueeze(2)          # collate zero-pads (otrans/data/loader.py:81)
    return x, mask


def run_case(name, params, dec_over=None, beam=3, max_len=8):
    params = copy.deepcopy(params)
    if dec_over:
        params['decoder'].update(dec_over)
    torch.manual_seed(1234)
    model = End2EndModel['speech2text'](params)
    model.eval()
    # randomise BatchNorm running stats / LN affine so they are not the identity
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'norm' in n and n.endswith('weight'):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            if 'norm' in n and n.endswith('bias'):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            if n.endswith('.bias') and 'norm' not in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        for n, buf in model.named_buffers():
            if n.endswith('running_mean'):
                buf.add_(0.3 * torch.randn(buf.shape, generator=g))
            if n.endswith('running_var'):
                buf.mul_(1.0 + 0.5 * torch.rand(buf.shape, generator=g))
        # random init + tied N(0,1) embeddings make <S/E> top-1 at step 1 with
        # near-one-hot posteriors; shrink the tied table so hypotheses end at different steps and the
        # finished-hypothesis masking (speech2text.py:156-192) is exercised end to end.
        model.decoder.embedding.weight.mul_(0.12)
        model.decoder.output_layer.bias[1] = 2.2

    x, mask = make_batch(0, 3, 90, params['frontend']['input_size'], [90, 61, 74])
    out = {'params': params}
    sd = {}
    for part in ('frontend', 'encoder', 'decoder'):
        for k, v in getattr(model, part).state_dict().items():
            sd[f'{part}.{k}'] = v.clone()
    out['state_dict'] = sd
    out['x'], out['mask'] = x, mask
    with torch.no_grad():
        fe, fmask = model.frontend(x, mask)
        out['frontend_out'], out['frontend_mask'] = fe, fmask
        mem, mmask, _ = model.encoder(fe, fmask)
        out['memory'], out['memory_mask'] = mem, mmask
        gt = torch.Generator().manual_seed(3)
        tgt = torch.randint(3, params['decoder']['vocab_size'], (3, 9), generator=gt)
        tgt[:, 0] = 1
        tgt[0, 7:] = torch.tensor([1, 0])   # EOS then PAD
        tgt[1, 5:] = torch.tensor([1, 0, 0, 0])
        tgt[2, 8] = 1
        out['targets'] = tgt
        logits, _ = model.decoder(tgt[:, :-1], mem, mmask)
        out['logits'] = logits
        loss, _ = model({'inputs': x, 'mask': mask}, {'targets': tgt, 'targets_length': None})
        out['loss'] = loss
        lp, _, _ = model.decoder.inference(tgt[:, :4], mem, mmask, None)
        out['inference_log_probs'] = lp

        # beam search through the real Recognizer, tracing every step's integer decisions
        rec = SpeechToTextRecognizer(model, lm=None, beam_width=beam, nbest=2, max_len=max_len,
                                     idx2unit=None, penalty=0.6, lamda=5, ngpu=0)
        steps = []
        orig = rec.decode_step

        def traced(preds, memory, memory_mask, cache, scores, flag):
            r = orig(preds, memory, memory_mask, cache, scores, flag)
            steps.append({'preds': r[0].clone(), 'scores': r[2].clone(), 'flag': r[3].clone()})
            return r
        rec.decode_step = traced
        rec.nbest_translate = lambda nbest_preds: nbest_preds    # keep ids (idx2unit is None)
        nbest_preds, nbest_scores = rec.recognize(x, mask)
        out['beam'] = {'beam': beam, 'nbest': 2, 'max_len': max_len, 'penalty': 0.6, 'lamda': 5,
                       'steps': steps, 'nbest_preds': nbest_preds, 'nbest_scores': nbest_scores}
    path = os.path.join(HERE, name + '.pt')
    torch.save(out, path)
    print(name, os.path.getsize(path) // 1024, 'KiB', 'loss', float(out['loss']),
          'steps', len(steps), 'nbest0', nbest_preds[0, 0].tolist())


def beam_step_cases():
    """decode_step integer behaviour on synthetic log-probs incl. finished hypotheses (no model)."""
    from otrans.recognize.speech2text import mask_finished_preds, mask_finished_scores
    g = torch.Generator().manual_seed(11)
    cases = []
    for (b, beam, v, l) in [(2, 3, 17, 1), (3, 4, 50, 5), (4, 10, 4234, 7), (1, 1, 9, 3)]:
        n = b * beam
        lp = torch.log_softmax(torch.randn(n, v, generator=g) * 3, dim=-1)
        preds = torch.randint(2, v, (n, l), generator=g)
        preds[:, 0] = 1
        scores = -torch.rand(n, 1, generator=g) * 5
        flag = torch.rand(n, 1, generator=g) < 0.3
        if l > 1:
            preds[flag.view(-1), -1] = 1
        if l == 1:
            scores = torch.tensor([0.0] + [float('-inf')] * (beam - 1)).repeat(b).unsqueeze(1)
            flag = torch.zeros_like(flag)

        class _M:  # minimal stand-in so the real decode_step can run on given log-probs
            pass
        rec = SpeechToTextRecognizer.__new__(SpeechToTextRecognizer)
        rec.beam_width, rec.lm, rec.lm_weight = beam, None, 0.0
        rec.decode = lambda p, m, mm, c, _lp=lp: (_lp.clone(), None, None)
        np_, _, ns, nf = rec.decode_step(preds.clone(), None, None, {'decoder': None, 'lm': None},
                                         scores.clone(), flag.clone())
        cases.append({'beam': beam, 'log_probs': lp, 'preds': preds, 'scores': scores, 'flag': flag,
                      'new_preds': np_, 'new_scores': ns, 'new_flag': nf})
    torch.save(cases, os.path.join(HERE, 'beam_step_cases.pt'))
    print('beam_step_cases', len(cases))


def lm_case():
    """TransformerLanguageModel.predict of the real reference on a tiny config (model/lm.py:93-163)."""
    from otrans.model import LanguageModel
    params = dict(vocab_size=40, num_blocks=2, d_model=32, n_heads=4, d_ff=48, residual_dropout=0.0, smoothing=0.1,
                  share_embedding=True)
    torch.manual_seed(4321)
    lm = LanguageModel['transformer_lm'](params).eval()
    with torch.no_grad():
        lm.embedding.weight.mul_(0.3)
    g = torch.Generator().manual_seed(5)
    toks = torch.randint(2, 40, (4, 7), generator=g)
    toks[:, 0] = 1
    with torch.no_grad():
        out = {'params': params, 'state_dict': {'lm.' + k: v.clone() for k, v in lm.state_dict().items()},
               'tokens': toks, 'last': lm.predict(toks, last_frame=True), 'all': lm.predict(toks, last_frame=False)}
    torch.save(out, os.path.join(HERE, 'small_transformer_lm.pt'))
    print('small_transformer_lm', tuple(out['last'].shape), tuple(out['all'].shape))


def train_case():
    """One inner step of Trainer.train_one_epoch (train/trainer.py:206-234) through the REAL reference:
    model.train() forward -> loss.backward() -> clip_grad_norm_(5) -> TransformerScheduler.step -> Adam.step.
    Dropout rates are 0 (gradient parity is defined without the stochastic residual dropout)."""
    from otrans.train.scheduler import TransformerScheduler
    params = small_params('transformer')
    for part in ('frontend', 'encoder', 'decoder'):
        for k in list(params[part]):
            if 'dropout' in k:
                params[part][k] = 0.0
    torch.manual_seed(1234)
    model = End2EndModel['speech2text'](params)
    model.train()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'norm' in n:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith('.bias'):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        model.decoder.embedding.weight.mul_(0.12)
    x, mask = make_batch(0, 3, 90, params['frontend']['input_size'], [90, 61, 74])
    gt = torch.Generator().manual_seed(3)
    tgt = torch.randint(3, params['decoder']['vocab_size'], (3, 9), generator=gt)
    tgt[:, 0] = 1
    tgt[0, 7:] = torch.tensor([1, 0])
    tgt[1, 5:] = torch.tensor([1, 0, 0, 0])
    tgt[2, 8] = 1
    sd = {}
    for part in ('frontend', 'encoder', 'decoder'):
        for k, v in getattr(model, part).state_dict().items():
            sd[f'{part}.{k}'] = v.clone()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6, amsgrad=False)
    sched = TransformerScheduler(opt, model_size=32, warmup_steps=100, factor=1.0)
    loss, _ = model({'inputs': x, 'mask': mask}, {'targets': tgt, 'targets_length': None})
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)      # small max_norm so that clipping is active
    sched.step()
    opt.step()
    after = {n: p.detach().clone() for n, p in model.named_parameters()}
    out = {'params': params, 'state_dict': sd, 'x': x, 'mask': mask, 'targets': tgt, 'loss': loss.detach(), 'grads': grads,
           'grad_norm': torch.as_tensor(float(grad_norm)), 'clip': 0.5, 'lr': sched.lr, 'after': after,
           'adam': dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6), 'sched': dict(model_size=32, warmup_steps=100)}
    path = os.path.join(HERE, 'train_step_postnorm_glu.pt')
    torch.save(out, path)
    print('train_step', os.path.getsize(path) // 1024, 'KiB loss', float(loss), 'grad_norm', float(grad_norm), 'lr', sched.lr)


def ctc_case():
    """Joint-CTC training forward + backward through the REAL reference (model/speech2text.py:30-36,60-72 -> CTCAssistor,
    model/ctc.py:12-52): loss = (1 - w) * attention loss + w * nn.CTCLoss(blank 0, zero_infinity) and every gradient."""
    params = small_params('transformer')
    for part in ('frontend', 'encoder', 'decoder'):
        for k in list(params[part]):
            if 'dropout' in k:
                params[part][k] = 0.0
    params['ctc_weight'] = 0.3
    params['encoder_output_size'] = params['encoder']['d_model']
    torch.manual_seed(1234)
    model = End2EndModel['speech2text'](params)
    model.train()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'norm' in n:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith('.bias'):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        model.decoder.embedding.weight.mul_(0.12)
    x, mask = make_batch(0, 3, 90, params['frontend']['input_size'], [90, 61, 74])
    gt = torch.Generator().manual_seed(3)
    tgt = torch.randint(3, params['decoder']['vocab_size'], (3, 9), generator=gt)
    tgt[:, 0] = 1
    tgt[0, 7:] = torch.tensor([1, 0])
    tgt[1, 5:] = torch.tensor([1, 0, 0, 0])
    tgt[2, 8] = 1
    tgt[2, 3] = tgt[2, 2]                      # a repeated label: CTC must pass through the blank between them
    tlen = torch.tensor([7, 5, 8], dtype=torch.int32)      # labels incl. <S/E> (collate: targets_length + 1, data/loader.py:94)
    sd = {}
    for part in ('frontend', 'encoder', 'decoder'):
        for k, v in getattr(model, part).state_dict().items():
            sd[f'{part}.{k}'] = v.clone()
    for k, v in model.assistor.state_dict().items():
        sd[f'assistor.{k}'] = v.clone()
    loss, aux = model({'inputs': x, 'mask': mask}, {'targets': tgt, 'targets_length': tlen})
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    out = {'params': params, 'state_dict': sd, 'x': x, 'mask': mask, 'targets': tgt, 'targets_length': tlen,
           'loss': loss.detach(), 'ctc_loss': torch.as_tensor(aux['CTCLoss']), 'grads': grads}
    path = os.path.join(HERE, 'joint_ctc_postnorm_glu.pt')
    torch.save(out, path)
    print('joint_ctc', os.path.getsize(path) // 1024, 'KiB loss', float(loss), 'ctc', aux['CTCLoss'])


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'train':
        train_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'ctc':
        ctc_case()
        sys.exit(0)
    lm_case()
    run_case('small_transformer_postnorm_glu', small_params('transformer'))
    run_case('small_transformer_prenorm_relu',
             small_params('transformer', normalize_before=True, activation='relu'),
             dec_over={'normalize_before': True, 'activation': 'relu'})
    run_case('small_transformer_relpos', small_params('transformer', relative_positional=True))
    run_case('small_conformer', small_params('conformer'))
    beam_step_cases()
