"""CPU oracle for the batched beam search  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates SpeechToTextRecognizer.recognize / decode_step (otrans/recognize/speech2text.py:39-153)
and mask_finished_scores / mask_finished_preds (:156-192) on plain tensors.  Integer outputs
(token ids, parent rows, top-k offsets) are the bit-exact parity target for the CUDA beam kernel.
"""
import torch

from .speech_model import EOS, BOS, decoder_inference, decoder_kwargs, encode


def beam_step(log_probs, preds, scores, flag, beam, lm_log_probs=None, lm_weight=0.0, trace=None):
    """One decode_step after the decoder has produced log_probs (speech2text.py:102-153).

    log_probs f32 [N,V], preds i64 [N,l], scores f32 [N,1], flag bool [N,1], N = B*beam.
    Returns (preds [N,l+1], scores [N,1], flag [N,1]).
      (i)   optional + lm_weight * lm_log_probs                                  (:102-105)
      (ii)  top-k(beam) over V per hypothesis                                    (:112)
      (iii) finished hyps: candidate scores -> [0,-inf,...], tokens -> EOS        (:114-115,156-192)
      (iv)  scores[N,1] + cand -> view [B, beam*beam]                            (:118-119)
      (v)   top-k(beam) per utterance                                            (:122)
      (vi)  flat idx = b*beam^2 + offset; token = cand_tok[idx]; parent = idx//beam;
            new prefix = cat(preds[parent], token)                               (:126-140)
      (vii) flag = last == EOS                                                   (:146)
    """
    n = scores.shape[0]
    b = n // beam
    if lm_log_probs is not None:
        log_probs = log_probs + lm_weight * lm_log_probs
    k_scores, k_tok = log_probs.topk(beam)
    fin = flag.view(n, 1)
    col = torch.arange(beam).view(1, beam)
    k_scores = torch.where(fin & (col > 0), torch.full_like(k_scores, float('-inf')), k_scores)
    k_scores = torch.where(fin & (col == 0), torch.zeros_like(k_scores), k_scores)
    k_tok = torch.where(fin.expand(n, beam), torch.full_like(k_tok, EOS), k_tok)
    cand = (scores + k_scores).view(b, beam * beam)
    new_scores, offs = torch.topk(cand, k=beam)
    flat = (torch.arange(b).view(b, 1) * beam * beam + offs).view(-1)
    tok = k_tok.reshape(-1)[flat]
    parent = torch.div(flat, beam, rounding_mode='floor')
    new_preds = torch.cat((preds[parent], tok.view(-1, 1)), dim=1)
    new_flag = (new_preds[:, -1] == EOS).view(-1, 1)
    if trace is not None:
        trace.append({'k_tok': k_tok.clone(), 'offs': offs.clone(), 'parent': parent.clone(),
                      'tok': tok.clone(), 'scores': new_scores.reshape(-1).clone()})
    return new_preds, new_scores.reshape(-1, 1), new_flag


def beam_finalize(preds, scores, beam, nbest, penalty, lamda):
    """Tail of recognize (speech2text.py:70-91): lengths = #(tok != EOS) (BOS==EOS so BOS is not
    counted), scores /= ((lamda+len)/(lamda+1))**penalty applied ONCE at the end, sort desc,
    gather, strip BOS, take nbest."""
    n = preds.shape[0]
    b = n // beam
    scores = scores.view(b, beam).clone()
    p = preds.view(b, beam, -1)
    lengths = torch.sum(torch.ne(p, EOS).float(), dim=-1)
    if penalty:
        scores = scores / torch.pow((lamda + lengths) / (lamda + 1), penalty)
    s_scores, order = torch.sort(scores, dim=-1, descending=True)
    idx = (order + torch.arange(b).view(b, 1) * beam).view(-1)
    s_preds = preds[idx].view(b, beam, -1)
    k = min(beam, nbest)
    return s_preds[:, :k, 1:], s_scores[:, :k]


def beam_search_from_memory(memory, memory_mask, sd, params, beam=10, nbest=1, max_len=60,
                            penalty=0.6, lamda=5, policy=None, trace=None):
    """recognize() after encode (speech2text.py:49-91): tile memory x beam, init preds=BOS and
    scores=[0,-inf...], loop <= max_len with early break when every hypothesis ended."""
    b, t, d = memory.shape
    bm = memory.unsqueeze(1).repeat(1, beam, 1, 1).view(b * beam, t, d)
    bmask = memory_mask.unsqueeze(1).repeat(1, beam, 1).view(b * beam, t)
    preds = torch.full((b * beam, 1), BOS, dtype=torch.long)
    scores = torch.tensor([0.0] + [float('-inf')] * (beam - 1)).repeat(b).unsqueeze(1)
    flag = torch.zeros_like(scores, dtype=torch.bool)
    kw = decoder_kwargs(params)
    for _ in range(max_len):
        lp = decoder_inference(preds, bm, bmask, sd, 'decoder.', policy=policy, **kw)
        preds, scores, flag = beam_step(lp, preds, scores, flag, beam, trace=trace)
        if int(flag.sum()) == b * beam:
            break
    return beam_finalize(preds, scores, beam, nbest, penalty, lamda) + (preds, scores)


def recognize(x, mask, sd, params, **kw):
    """encode + beam search; returns (nbest ids [B,nbest,L], nbest scores, raw preds, raw scores)."""
    policy = kw.get('policy')
    memory, mmask = encode(x, mask, sd, params, policy)
    return beam_search_from_memory(memory, mmask, sd, params, **kw)
