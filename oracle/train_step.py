"""ORACLE (test infrastructure only -- never imported by opentransformer_b200/): CPU fp32 restatement of one inner
training step of the reference, otrans/train/trainer.py:206-234:

    loss, _ = model(inputs, targets)           model/speech2text.py:39-64
    loss.backward()                            torch autograd through the functional restatement in speech_model.py
    clip_grad_norm_(parameters, clip)          trainer.py:221
    scheduler.step(); optimizer.step()         scheduler.py:50-54,137-138 ; torch.optim.Adam (L2 weight decay)

Pinned against the real reference by tests/golden/train_step_postnorm_glu.pt (tests/test_oracle_golden.py).
"""
import torch

from . import speech_model as om

TIED = ('decoder.embedding.weight', 'decoder.output_layer.weight')


def loss_and_grads(x, mask, truth, sd, params, truth_length=None):
    """-> (loss, {name: grad}) with the reference's parameter names; a tied output layer contributes to
    'decoder.embedding.weight' (decoder/transformer.py:156-158) and has no entry of its own."""
    leaf = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    tied = params['decoder'].get('share_embedding', False)
    if tied:
        leaf[TIED[1]] = leaf[TIED[0]]
    loss, _ = om.model_forward_loss(x, mask, truth, leaf, params, truth_length=truth_length)
    loss.backward()
    grads = {}
    for k, v in leaf.items():
        if not v.requires_grad or (tied and k == TIED[1]):
            continue
        grads[k] = v.grad if v.grad is not None else torch.zeros_like(v)
    return loss.detach(), grads


def transformer_lr(step, model_size, warmup_steps, factor=1.0):
    """TransformerScheduler.get_step_lr (scheduler.py:137-138)."""
    return factor * model_size ** (-0.5) * min(step ** (-0.5), step * warmup_steps ** (-1.5))


def first_step_index():
    """BaseScheduler starts at global_step = 1 and its constructor already calls step() once (scheduler.py:22,42-46);
    the trainer calls scheduler.step() again before the first optimizer.step() (trainer.py:232) -> lr(step = 3)."""
    return 3


def clip_and_adam(weights, grads, m, v, t, lr, clip, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6):
    """In-place: clip_grad_norm_ over ALL grads, then torch.optim.Adam's update rule; t = 1-based step count.
    Returns the un-clipped total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = min(1.0, float(clip) / (float(total) + 1e-6)) if clip > 0 else 1.0
    b1, b2 = betas
    for k, w in weights.items():
        if k not in grads:
            continue
        g = grads[k] * coef + weight_decay * w
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        denom = v[k].sqrt() / (1 - b2 ** t) ** 0.5 + eps
        w -= lr / (1 - b1 ** t) * m[k] / denom
    return total
