"""CPU oracle for the OpenTransformer hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, fp32, torch-CPU restatement of the reference algorithm (frontend -> encoder ->
decoder -> loss).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path (``opentransformer_b200``)
never does.

Every function takes a flat ``state_dict`` (the reference's own checkpoint key names, e.g.
``blocks.3.slf_attn.qvk_proj.weight``) plus a ``prefix`` and restates what the matching reference
module computes.  Citations are ``path:line`` under the reference tree.

Parity pinning: the reference ships no golden vectors, so this oracle is pinned
against the reference *executed in the build container* -- ``tests/golden/make_golden.py`` imports
the real ``otrans`` package, runs it on seeded inputs and commits inputs/weights/outputs as
fixtures; ``tests/test_oracle_golden.py`` replays them through this file.

``policy``: ``None`` reproduces the reference exactly (fp32 everywhere).  ``'bf16'`` additionally
rounds GEMM operands to bfloat16 at the points where the CUDA path stores bf16 (documented in
DESIGN.md) so that beam-search token ids can be compared bit-exactly; it is an emulation of the
product's rounding, the arithmetic is still the reference's.
"""
import math

import torch
import torch.nn.functional as F

PAD, BOS, EOS = 0, 1, 1  # otrans/data/__init__.py:7-10


# ----------------------------------------------------------------------------------------------
# rounding policy helpers
# ----------------------------------------------------------------------------------------------
def _r(x, policy):
    """Round to bf16 and back when the bf16 storage policy is emulated."""
    if policy == 'bf16':
        return x.to(torch.bfloat16).to(torch.float32)
    return x


# Training-mode residual dropout (encoder/transformer.py:32-33,54,61; decoder/transformer.py:36-38).  torch's own Philox
# masks cannot be reproduced by another implementation, so parity tests REPLAY the product's masks: set_dropout(fn) installs
# fn(site, tensor) -> tensor (e.g. tensor * keep / (1 - p)); None (default) = eval mode / rate 0.  Sites: 'encoder.blocks.i'
# x {0: after self-attention, 1: after the feed-forward}, 'decoder.blocks.i' x {0, 1, 2} (self, cross, feed-forward).
_DROP = None


def set_dropout(fn):
    global _DROP
    _DROP = fn


def _drop(prefix, k, t):
    return t if _DROP is None else _DROP((prefix, k), t)


def linear(x, sd, prefix, policy=None, bias=True):
    """nn.Linear: x @ W^T + b (operands optionally bf16-rounded, fp32 accumulate)."""
    w = sd[prefix + '.weight']
    b = sd.get(prefix + '.bias') if bias else None
    y = _r(x, policy) @ _r(w, policy).t()
    if b is not None:
        y = y + b
    return y


def layer_norm(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], eps)


# ----------------------------------------------------------------------------------------------
# frontend: otrans/frontend/conv.py
# ----------------------------------------------------------------------------------------------
def conv_out_len(n, k=3, s=2, pad=0):
    return (n + 2 * pad - k) // s + 1


def conv2d_layer(x, mask, sd, prefix, kernel=(3, 3), stride=2, policy=None):
    """Conv2dLayer.forward (frontend/conv.py:50-76) + return_output_mask (:78-83).

    relu(conv2d(x)) with padding (0, kw//2): time is NOT padded, frequency is.
    mask' = mask[:, kh//2::stride][:, :t_out].
    """
    kh, kw = kernel
    w = sd[prefix + '.conv_layer.weight']
    b = sd[prefix + '.conv_layer.bias']
    y = F.conv2d(_r(x, policy), _r(w, policy), b, stride=stride, padding=(0, kw // 2))
    y = torch.relu(y)
    mask = mask[:, kh // 2::stride][:, :y.shape[2]]
    return y, mask


def conv_frontend(x, mask, sd, prefix='', policy=None, layer_norm_out=False):
    """ConvFrontEnd.forward (frontend/conv.py:131-153).

    x [B,T,F] -> unsqueeze(1) -> conv1 -> conv2 -> [B,T',C*F'] (channel-major feature index
    c*F'+f, conv.py:145) -> Linear.  Neither input nor output is masked (conv.py:148 is a comment).
    """
    y = x.unsqueeze(1)
    y, mask = conv2d_layer(y, mask, sd, prefix + 'conv1', policy=policy)
    y = _r(y, policy)
    y, mask = conv2d_layer(y, mask, sd, prefix + 'conv2', policy=policy)
    b, c, t, f = y.shape
    y = y.transpose(1, 2).reshape(b, t, c * f)
    y = linear(y, sd, prefix + 'output_layer', policy)
    if layer_norm_out:
        y = layer_norm(y, sd, prefix + 'layer_norm')
    return y, mask


# ----------------------------------------------------------------------------------------------
# positional encoding: otrans/module/pos.py
# ----------------------------------------------------------------------------------------------
def sinusoid_table(positions, d):
    """_embedding_from_positions (pos.py:30-42): PE[p,2i]=sin(p*w_i), PE[p,2i+1]=cos(p*w_i),
    w_i = exp(-2i*ln(1e4)/d); positions is a 1-D tensor (may be negative for rel-pos)."""
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    ang = positions.float().unsqueeze(-1) * div
    pe = torch.zeros(positions.numel(), d)
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


def abs_posenc(x):
    """PositionalEncoding.forward, scale_learnable False branch (pos.py:44-57): x*sqrt(d)+PE."""
    t, d = x.shape[1], x.shape[2]
    pe = sinusoid_table(torch.arange(t), d)
    return x * math.sqrt(d) + pe.unsqueeze(0), pe.unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# attention: otrans/module/attention.py
# ----------------------------------------------------------------------------------------------
def _split_heads(x, h):
    b, t, d = x.shape
    return x.reshape(b, t, h, d // h).transpose(1, 2)  # [B,h,T,dk]


def attention_core(q, k, v, mask, policy=None, extra_scores=None):
    """scores/sqrt(dk) -> masked_fill(~mask,-inf) -> softmax -> @V -> merge heads
    (attention.py:80, 34-41).  q,k,v [B,h,T,dk]; mask broadcastable to [B,1,T1,T2]."""
    dk = q.shape[-1]
    scores = _r(q, policy) @ _r(k, policy).transpose(2, 3)
    if extra_scores is not None:
        scores = scores + extra_scores
    scores = scores / math.sqrt(dk)
    if mask is not None:
        scores = scores.masked_fill(~mask, float('-inf'))
    w = torch.softmax(scores, dim=-1)
    ctx = _r(w, policy) @ _r(v, policy)
    b, h, t, _ = ctx.shape
    return ctx.transpose(1, 2).reshape(b, t, h * dk), w


def mha_self(x, mask, sd, prefix, n_heads, policy=None):
    """MultiHeadedSelfAttention.forward (attention.py:60-84). qvk_proj split order is Q,K,V (:73).
    mask [B,1,T] (encoder) or [B,T,T] (decoder) -> unsqueeze(1)."""
    d = x.shape[-1]
    qkv = _r(linear(x, sd, prefix + '.qvk_proj', policy), policy)
    q, k, v = torch.split(qkv, d, dim=-1)
    ctx, w = attention_core(_split_heads(q, n_heads), _split_heads(k, n_heads),
                            _split_heads(v, n_heads), mask.unsqueeze(1) if mask is not None else None,
                            policy)
    return linear(_r(ctx, policy), sd, prefix + '.output_proj', policy), w


def mha_cross(x, memory, memory_mask, sd, prefix, n_heads, policy=None):
    """MultiHeadedCrossAttention.forward (attention.py:119-144). vk_proj split order is K,V (:134).
    memory_mask [B,1,T2]."""
    d = x.shape[-1]
    q = _r(linear(x, sd, prefix + '.q_proj', policy), policy)
    kv = _r(linear(memory, sd, prefix + '.vk_proj', policy), policy)
    k, v = torch.split(kv, d, dim=-1)
    ctx, w = attention_core(_split_heads(q, n_heads), _split_heads(k, n_heads),
                            _split_heads(v, n_heads), memory_mask.unsqueeze(1), policy)
    return linear(_r(ctx, policy), sd, prefix + '.output_proj', policy), w


def mha_self_relpos(x, mask, pos, sd, prefix, n_heads, policy=None):
    """MultiHeadedSelfAttentionWithRelPos.forward (attention.py:217-253) + _RelPosBias (:196-215).

    AC = (q+u) k^T ; BD[i,j] = (q_i+v) . P[j-i+T-1] with P = pos_proj(PE[-(T-1)..T-1]);
    scores = (AC+BD)/sqrt(dk).  Quirk: enable_output_proj == dropout_rate, so with the
    shipped slf_attn_dropout 0.0 there is NO output projection (attention.py:178).
    """
    b, t, d = x.shape
    dk = d // n_heads
    qkv = _r(linear(x, sd, prefix + '.qvk_proj', policy), policy)
    q, k, v = torch.split(qkv, d, dim=-1)
    q = q.reshape(b, t, n_heads, dk)
    kh, vh = _split_heads(k, n_heads), _split_heads(v, n_heads)
    p = _r(linear(pos, sd, prefix + '.pos_proj', policy, bias=False), policy)  # [1,2T-1,d]
    p = p.reshape(pos.shape[0], -1, n_heads, dk).transpose(1, 2)          # [1,h,2T-1,dk]
    u = sd[prefix + '.posu']  # [1,1,h,dk]
    vb = sd[prefix + '.posv']
    qu = _r(q + u, policy).transpose(1, 2)
    qv = _r(q + vb, policy).transpose(1, 2)
    full = qv @ _r(p, policy).transpose(-2, -1)                            # [B,h,T,2T-1]
    idx = (torch.arange(t)[None, :] - torch.arange(t)[:, None]) + (t - 1)  # j-i+T-1
    bd = torch.gather(full, 3, idx.reshape(1, 1, t, t).expand(b, n_heads, t, t))
    ac = qu @ _r(kh, policy).transpose(-2, -1)
    scores = (ac + bd) / math.sqrt(dk)
    if mask is not None:
        scores = scores.masked_fill(~mask.unsqueeze(1), float('-inf'))
    w = torch.softmax(scores, dim=-1)
    ctx = _r(w, policy) @ _r(vh, policy)
    ctx = ctx.transpose(1, 2).reshape(b, t, d)
    if (prefix + '.output_proj.weight') in sd:
        ctx = linear(_r(ctx, policy), sd, prefix + '.output_proj', policy)
    return ctx, w


# ----------------------------------------------------------------------------------------------
# feed forward: otrans/module/ffn.py:24-41
# ----------------------------------------------------------------------------------------------
def _act(x, activation):
    if activation == 'glu':
        return F.glu(x)          # first half * sigmoid(second half)
    if activation == 'relu':
        return torch.relu(x)
    if activation == 'gelu':
        return F.gelu(x)
    if activation == 'tanh':
        return torch.tanh(x)
    if activation == 'swish':
        return x * torch.sigmoid(x)
    raise ValueError(activation)


def ffn(x, sd, prefix, activation, policy=None):
    """PositionwiseFeedForward.forward (ffn.py:38-41): w_2(act(w_1(x)))."""
    h = _act(linear(x, sd, prefix + '.w_1', policy), activation)
    return linear(_r(h, policy), sd, prefix + '.w_2', policy)


# ----------------------------------------------------------------------------------------------
# encoders
# ----------------------------------------------------------------------------------------------
def transformer_encoder_layer(x, mask, pos, sd, prefix, n_heads, activation, normalize_before,
                              relative_positional, policy=None):
    """TransformerEncoderLayer.forward (encoder/transformer.py:41-65).  In the pre-norm variant the
    residual is taken AFTER the norm (:42-44), which is reproduced as is."""
    if normalize_before:
        x = layer_norm(x, sd, prefix + '.norm1')
    res = x
    if relative_positional:
        a, _ = mha_self_relpos(_r(x, policy), mask, pos, sd, prefix + '.slf_attn', n_heads, policy)
    else:
        a, _ = mha_self(_r(x, policy), mask, sd, prefix + '.slf_attn', n_heads, policy)
    x = res + _drop(prefix, 0, a)
    if not normalize_before:
        x = layer_norm(x, sd, prefix + '.norm1')
    if normalize_before:
        x = layer_norm(x, sd, prefix + '.norm2')
    res = x
    x = res + _drop(prefix, 1, ffn(_r(x, policy), sd, prefix + '.feed_forward', activation, policy))
    if not normalize_before:
        x = layer_norm(x, sd, prefix + '.norm2')
    return _r(x, policy)


def transformer_encoder(x, mask, sd, prefix, n_blocks, n_heads, activation='glu',
                        normalize_before=False, relative_positional=False, policy=None,
                        return_layers=False):
    """TransformerEncoder.forward (encoder/transformer.py:114-134)."""
    if relative_positional:
        t = x.shape[1]
        pos = sinusoid_table(torch.arange(-(t - 1), t), x.shape[2]).unsqueeze(0)
        y = x
    else:
        y, pos = abs_posenc(x)
    y = _r(y, policy)
    layers = []
    for i in range(n_blocks):
        y = transformer_encoder_layer(y, mask.unsqueeze(1), pos, sd, f'{prefix}blocks.{i}', n_heads,
                                      activation, normalize_before, relative_positional, policy)
        layers.append(y)
    if normalize_before:
        y = layer_norm(y, sd, prefix + 'norm')
    return (y, mask, layers) if return_layers else (y, mask)


def conformer_conv_module(x, mask, sd, prefix, policy=None, bn_eps=1e-5):
    """ConformerConvolutionModule.forward (module/conformer.py:36-57), eval-mode BatchNorm.
    Linear d->2d, GLU, zero pads, depthwise Conv1d(k, pad (k-1)/2), BN(running stats), swish,
    Linear d->d, zero pads."""
    m = mask.unsqueeze(2)
    y = F.glu(linear(x, sd, prefix + '.pointwise_conv1', policy))
    y = _r(y.masked_fill(~m, 0.0), policy)
    w = sd[prefix + '.depthwise_conv.weight']          # [C,1,k]
    b = sd.get(prefix + '.depthwise_conv.bias')
    k = w.shape[-1]
    y = F.conv1d(y.transpose(1, 2), w, b, padding=(k - 1) // 2, groups=w.shape[0])
    y = F.batch_norm(y, sd[prefix + '.batch_norm.running_mean'], sd[prefix + '.batch_norm.running_var'],
                     sd[prefix + '.batch_norm.weight'], sd[prefix + '.batch_norm.bias'], False, 0.0, bn_eps)
    y = y * torch.sigmoid(y)
    y = _r(y.transpose(1, 2), policy)
    y = linear(y, sd, prefix + '.pointwise_conv2', policy)
    return y.masked_fill(~m, 0.0)


def conformer_block(x, mask, pos, sd, prefix, n_heads, activation, ffn_scale=0.5, macaron=True,
                    relative_positional=True, conv_first=False, policy=None):
    """ConformerEncoderBlock.forward (encoder/conformer.py:75-89) with residual_dropout 0.

    Quirk reproduced: the post-FFN is never applied, only post_ffn_norm then final_norm (:87-89).
    """
    if macaron:
        x = x + ffn_scale * ffn(_r(layer_norm(x, sd, prefix + '.macaron_ffn_norm'), policy), sd,
                                prefix + '.pre_ffn', activation, policy)

    def attn(x):
        y = _r(layer_norm(x, sd, prefix + '.mha_norm'), policy)
        if relative_positional:
            a, _ = mha_self_relpos(y, mask.unsqueeze(1), pos, sd, prefix + '.mha', n_heads, policy)
        else:
            a, _ = mha_self(y, mask.unsqueeze(1), sd, prefix + '.mha', n_heads, policy)
        return x + a

    def conv(x):
        y = _r(layer_norm(x, sd, prefix + '.conv_norm'), policy)
        return x + conformer_conv_module(y, mask, sd, prefix + '.conv', policy)

    x = attn(conv(x)) if conv_first else conv(attn(x))
    x = layer_norm(x, sd, prefix + '.post_ffn_norm')
    return _r(layer_norm(x, sd, prefix + '.final_norm'), policy)


def conformer_encoder(x, mask, sd, prefix, n_blocks, n_heads, activation='glu', ffn_scale=0.5,
                      macaron=True, positional_encoding=True, relative_positional=True,
                      conv_first=False, policy=None, return_layers=False):
    """ConformerEncoder.forward (encoder/conformer.py:151-164)."""
    pos = None
    y = x
    if positional_encoding:
        if relative_positional:
            t = x.shape[1]
            pos = sinusoid_table(torch.arange(-(t - 1), t), x.shape[2]).unsqueeze(0)
        else:
            y, pos = abs_posenc(x)
    y = _r(y, policy)
    layers = []
    for i in range(n_blocks):
        y = conformer_block(y, mask, pos, sd, f'{prefix}blocks.{i}', n_heads, activation, ffn_scale,
                            macaron, relative_positional, conv_first, policy)
        layers.append(y)
    return (y, mask, layers) if return_layers else (y, mask)


# ----------------------------------------------------------------------------------------------
# decoder: otrans/decoder/transformer.py
# ----------------------------------------------------------------------------------------------
def transformer_decoder(targets, memory, memory_mask, sd, prefix, n_blocks, n_heads,
                        activation='glu', normalize_before=False, policy=None):
    """TransformerDecoder.forward (decoder/transformer.py:161-183): embedding -> x*sqrt(d)+PE ->
    causal tril mask only (decoder/utils.py:7-11, no target-pad mask) -> n x [self, cross, ffn]
    -> output_layer.  Returns logits [B,L,V]."""
    emb = sd[prefix + 'embedding.weight']
    x = _r(emb, policy)[targets]
    x, _ = abs_posenc(x)
    x = _r(x, policy)
    L = targets.shape[1]
    causal = torch.tril(torch.ones(L, L)).bool().unsqueeze(0).expand(targets.shape[0], L, L)
    mm = memory_mask.unsqueeze(1)
    for i in range(n_blocks):
        p = f'{prefix}blocks.{i}'
        if normalize_before:
            x = layer_norm(x, sd, p + '.norm1')
        res = x
        a, _ = mha_self(_r(x, policy), causal, sd, p + '.slf_attn', n_heads, policy)
        x = res + _drop(p, 0, a)
        if not normalize_before:
            x = layer_norm(x, sd, p + '.norm1')
        if normalize_before:
            x = layer_norm(x, sd, p + '.norm2')
        res = x
        a, _ = mha_cross(_r(x, policy), memory, mm, sd, p + '.src_attn', n_heads, policy)
        x = res + _drop(p, 1, a)
        if not normalize_before:
            x = layer_norm(x, sd, p + '.norm2')
        if normalize_before:
            x = layer_norm(x, sd, p + '.norm3')
        res = x
        x = res + _drop(p, 2, ffn(_r(x, policy), sd, p + '.feed_forward', activation, policy))
        if not normalize_before:
            x = layer_norm(x, sd, p + '.norm3')
        x = _r(x, policy)
    if normalize_before:
        x = layer_norm(x, sd, prefix + 'after_norm')
    return linear(x, sd, prefix + 'output_layer', policy)


def decoder_inference(preds, memory, memory_mask, sd, prefix, **kw):
    """TransformerDecoder.inference (decoder/transformer.py:185-208): FULL forward over the prefix,
    log_softmax of the last position."""
    logits = transformer_decoder(preds, memory, memory_mask, sd, prefix, **kw)
    return F.log_softmax(logits[:, -1, :], dim=-1)


def transformer_lm_log_probs(targets, sd, prefix, n_blocks, n_heads, last_frame=True, policy=None):
    """TransformerLanguageModel.predict (model/lm.py:143-163): embedding -> x*sqrt(d)+PE -> post-norm encoder
    layers (GLU) under the causal tril mask (lm.py:14-18) -> output_project -> log_softmax."""
    x = _r(sd[prefix + 'embedding.weight'], policy)[targets]
    x, _ = abs_posenc(x)
    x = _r(x, policy)
    L = targets.shape[1]
    causal = torch.tril(torch.ones(L, L)).bool().unsqueeze(0).expand(targets.shape[0], L, L)
    for i in range(n_blocks):
        x = transformer_encoder_layer(x, causal, None, sd, f'{prefix}blocks.{i}', n_heads, 'glu', False, False, policy)
    logits = linear(x, sd, prefix + 'output_project', policy)
    if last_frame:
        return F.log_softmax(logits[:, -1, :].unsqueeze(1), dim=-1)
    return F.log_softmax(logits, dim=-1)


# ----------------------------------------------------------------------------------------------
# loss: otrans/module/loss.py:21-48
# ----------------------------------------------------------------------------------------------
def label_smoothing_loss(logits, target, smoothing=0.1):
    """sum_v KL(conf || softmax) per token, PAD(0) tokens zeroed, / #non-pad."""
    v = logits.shape[-1]
    flat = logits.reshape(-1, v)
    tgt = target.reshape(-1)
    conf = torch.full_like(flat, smoothing / (v - 1))
    conf.scatter_(1, tgt.unsqueeze(1), 1 - smoothing)
    logp = F.log_softmax(flat, dim=-1)
    # F.kl_div(logp, conf, 'none') = xlogy(conf, conf) - conf * logp  (0 log 0 = 0: loss.py:43 stays finite at smoothing 0 / 1)
    per_tok = torch.sum(torch.xlogy(conf, conf) - conf * logp, dim=-1)
    pad = tgt == PAD
    return torch.sum(per_tok.masked_fill(pad, 0.0)) / torch.sum(~pad)


# ----------------------------------------------------------------------------------------------
# whole-model helpers driven by the reference's YAML `model` dict
# ----------------------------------------------------------------------------------------------
def encode(x, mask, sd, params, policy=None, return_layers=False):
    """frontend + encoder as SpeechToTextRecognizer.encode drives them (recognize/speech2text.py:24-33).
    `sd` holds 'frontend.*' and 'encoder.*' keys."""
    fe = params['frontend']
    y, m = conv_frontend(x, mask, sd, 'frontend.', policy, fe.get('front_end_layer_norm', False))
    fe_out = y
    ep = params['encoder']
    if params['encoder_type'] == 'transformer':
        out = transformer_encoder(_r(y, policy), m, sd, 'encoder.', ep['n_blocks'], ep['n_heads'],
                                  ep.get('activation', 'relu'), ep.get('normalize_before', False),
                                  ep.get('relative_positional', False), policy, return_layers)
    elif params['encoder_type'] == 'conformer':
        out = conformer_encoder(_r(y, policy), m, sd, 'encoder.', ep.get('nblocks', 12), ep['n_heads'],
                                ep.get('activation', 'glu'), ep.get('ffn_scale', 0.5),
                                ep.get('macaron_style', True), ep.get('positional_encoding', True),
                                ep.get('relative_positional', True), ep.get('conv_first', False),
                                policy, return_layers)
    else:
        raise ValueError(params['encoder_type'])
    if return_layers:
        return out[0], out[1], fe_out, out[2]
    return out


def decoder_kwargs(params):
    dp = params['decoder']
    return dict(n_blocks=dp['n_blocks'], n_heads=dp['n_heads'], activation=dp.get('activation', 'relu'),
                normalize_before=dp.get('normalize_before', True))


def ctc_loss(memory, memory_mask, targets_out, targets_length, sd, prefix='assistor.', blank=0, policy=None):
    """SpeechToText.compute_ctc_loss -> CTCAssistor.forward (model/speech2text.py:66-69, model/ctc.py:33-52):
    nn.CTCLoss(blank, reduction 'mean', zero_infinity=True) on log_softmax(output_layer(memory)), lengths from the mask."""
    logits = linear(memory, sd, prefix + 'output_layer', policy)
    log_probs = F.log_softmax(logits, dim=-1)
    mem_len = memory_mask.sum(dim=-1)
    return F.ctc_loss(log_probs.transpose(0, 1), targets_out, mem_len, targets_length, blank=blank, reduction='mean',
                      zero_infinity=True)


def model_forward_loss(inputs, mask, truth, sd, params, policy=None, truth_length=None, return_parts=False):
    """SpeechToText.forward (model/speech2text.py:39-64): loss of decoder(truth[:, :-1]) vs truth[:, 1:]; with
    ctc_weight > 0 mixed with the joint-CTC loss of the encoder states, (1 - w) * att + w * ctc."""
    memory, mmask = encode(inputs, mask, sd, params, policy)
    logits = transformer_decoder(truth[:, :-1], memory, mmask, sd, 'decoder.', policy=policy,
                                 **decoder_kwargs(params))
    loss = label_smoothing_loss(logits, truth[:, 1:], params.get('smoothing', 0.1))
    w = params.get('ctc_weight', 0.0)
    if w > 0:
        lc = ctc_loss(memory, mmask, truth[:, 1:], truth_length, sd, policy=policy)
        total = (1 - w) * loss + w * lc
        return (total, logits, lc) if return_parts else (total, logits)
    return (loss, logits, None) if return_parts else (loss, logits)
