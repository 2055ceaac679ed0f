"""Model assembly with the reference's registries and checkpoint format.

Mirror of otrans/model/speech2text.py:14-87, otrans/model/__init__.py:6-9 and the three
string-keyed registries (frontend/__init__.py:8-12, encoder/__init__.py:10-13,
decoder/__init__.py:8-10).  A YAML `model:` section of the reference builds the B200 model
unchanged; checkpoints written by the reference load through load_model() key for key.
"""
import torch
import torch.nn as nn

from . import ops
from .modules import ConformerEncoder, ConvFrontEnd, TransformerDecoder, TransformerEncoder, _lengths

BuildFrontEnd = {'conv': ConvFrontEnd}
BuildEncoder = {'transformer': TransformerEncoder, 'conformer': ConformerEncoder}
BuildDecoder = {'transformer': TransformerDecoder}


class SpeechToText(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.frontend = BuildFrontEnd[params['frontend_type']](**params['frontend'])
        self.encoder = BuildEncoder[params['encoder_type']](**params['encoder'])
        self.decoder = BuildDecoder[params['decoder_type']](**params['decoder'])
        self.ctc_weight = params.get('ctc_weight', 0.0)
        if self.ctc_weight > 0.0:
            raise NotImplementedError('joint CTC is out of scope (SURVEY.md 8f row 4)')
        self.smoothing = params.get('smoothing', 0.0)

    def encode_bf16(self, inputs, mask):
        """frontend + encoder on the fused path -> (memory bf16 [B*T',D], lengths i32 [B], B, T')."""
        fe, enc = self.frontend, self.encoder
        B, T, _ = inputs.shape
        _, _, T2, _ = ops.conv_geometry(T, fe.input_size)
        lengths = _lengths(fe.output_mask(mask))
        if enc.fuse_abs_posenc():
            scale, table = enc.pos_emb.scale_and_table(T2, inputs.device)
            x, _ = fe.forward_bf16(inputs, scale, table)
        else:
            x, _ = fe.forward_bf16(inputs)
            x = enc.apply_posenc_bf16(x, B, T2)
        return enc.forward_bf16(x, B, T2, lengths), lengths, B, T2

    def forward_logits(self, inputs, mask, targets_in):
        """encoder + teacher-forced decoder (the forward half of speech2text.py:39-56) -> logits f32 [B,L,V]."""
        mem, lengths, B, T2 = self.encode_bf16(inputs, mask)
        L = targets_in.shape[1]
        logits = self.decoder.forward_bf16(targets_in, mem, lengths, B, L, T2)
        return logits.view(B, L, -1)[:, :, :self.decoder.vocab_size]

    def forward(self, inputs, targets):
        """SpeechToText.forward (model/speech2text.py:39-58) -> (loss, None).  In train mode with grad enabled the loss
        carries the hand-written backward (train.py): loss.backward() fills .grad of the fp32 parameters."""
        truth = targets['targets']
        if self.training and torch.is_grad_enabled():
            from . import train
            return train.loss_with_grad(self, inputs['inputs'], inputs['mask'], truth), None
        logits_pad = self._logits_padded(inputs['inputs'], inputs['mask'], truth[:, :-1].contiguous())
        loss, _ = ops.ls_cross_entropy(logits_pad, truth[:, 1:].contiguous(), self.decoder.vocab_size, self.smoothing)
        return loss, None

    def _logits_padded(self, inputs, mask, targets_in):
        mem, lengths, B, T2 = self.encode_bf16(inputs, mask)
        return self.decoder.forward_bf16(targets_in, mem, lengths, B, targets_in.shape[1], T2)

    def save_checkpoint(self, params, name):
        torch.save({'params': params, 'frontend': self.frontend.state_dict(),
                    'encoder': self.encoder.state_dict(), 'decoder': self.decoder.state_dict()}, name)

    def load_model(self, chkpt):
        self.frontend.load_state_dict(chkpt['frontend'])
        self.encoder.load_state_dict(chkpt['encoder'])
        self.decoder.load_state_dict(chkpt['decoder'])

    def load_flat_state_dict(self, sd):
        """Load a flat {'frontend.x': ..} dict (test fixtures)."""
        for part in ('frontend', 'encoder', 'decoder'):
            sub = {k[len(part) + 1:]: v for k, v in sd.items() if k.startswith(part + '.')}
            getattr(self, part).load_state_dict(sub)

    def set_epoch(self, epoch):
        pass


End2EndModel = {'speech2text': SpeechToText}
