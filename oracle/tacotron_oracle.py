"""CPU oracle for the Tacotron-2 training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a from-scratch restatement, in elementary torch tensor arithmetic (matmul, exp, tanh,
explicit time loops; float32 or float64), of the algorithm the reference executes through
``torch.nn`` modules.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package must never import ``oracle``.

Parity pinning: ``tests/golden/make_golden.py`` runs the *unmodified* reference modules (imported
read-only from /root/reference in the build container) on seeded inputs with a recorded dropout-mask
tape and stores inputs, weights, masks, outputs and gradients in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.  The reference ships
no tests / golden vectors of its own for this path, so those fixtures are the pin.

Every function cites the reference lines it restates (paths relative to the reference root).
Weights are looked up in a flat ``state_dict`` using the reference's parameter names, so reference
checkpoints (and the golden fixtures) load directly.

Randomness: the reference draws dropout masks inside ``F.dropout``; here every mask is an explicit
input (a "mask tape", dict of 0/1 float tensors).  A missing / ``None`` entry means "no dropout at
this site" (equivalent to p = 0).  Tape keys:
    enc{j}              encoder block j output mask
    prenet0, prenet1    [B, T+1, P] masks of the time-batched prenet pass (training)
    teacher             bool [T]   (True = feed ground truth at step i)
    att_h, att_c        [T, B, D]  attention-LSTM regulariser masks (dropout cell uses att_h only)
    gen_h, gen_c        [T, B, D]  generator-LSTM masks
    step_prenet0/1      [T, B, P]  prenet masks of free-running steps (used where teacher[i] is False)
    post{j}             postnet block j output mask
"""
import math
import torch


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------

def lengths_to_mask(lengths, max_length=None):
    """utils/__init__.py:7-10 -- mask[b, l] = l < lengths[b]."""
    ml = int(lengths.max()) if max_length is None else int(max_length)
    return torch.arange(ml, device=lengths.device)[None, :] < lengths[:, None]


def _sigmoid(x):
    return 1.0 / (1.0 + torch.exp(-x))


# Optional operand quantiser used to model the library's bf16 perf mode (operands rounded to bf16, fp32+ accumulate):
# tests set QUANT = bf16_round so that the comparison isolates kernel bugs from the (expected) bf16 rounding.
QUANT = None


def bf16_round(x):
    return x.to(torch.bfloat16).to(x.dtype)


def _q(x):
    return x if QUANT is None else QUANT(x)


def _linear(x, w, b=None):
    y = _q(x) @ _q(w).transpose(0, 1)
    return y if b is None else y + b


def apply_dropout(x, keep, p):
    """F.dropout in training mode with an explicit keep mask: x * keep / (1 - p)."""
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


def conv1d_same(x, w, dilation=1):
    """ConstantPad1d((k-1)*dil//2) + Conv1d(padding=0, bias=False)  (modules/layers.py:72-75).

    x [B, Cin, L], w [Cout, Cin, k] (cross-correlation), odd k. Returns [B, Cout, L].
    """
    B, Cin, L = x.shape
    Cout, _, k = w.shape
    pad = (k - 1) * dilation // 2
    xp = torch.zeros(B, Cin, L + 2 * pad, dtype=x.dtype, device=x.device)
    xp[:, :, pad:pad + L] = x
    y = torch.zeros(B, Cout, L, dtype=x.dtype, device=x.device)
    for t in range(k):
        # y[b, o, l] += sum_i w[o, i, t] * xp[b, i, l + t*dil]
        y = y + torch.einsum('oi,bil->bol', w[:, :, t], xp[:, :, t * dilation:t * dilation + L])
    return y


def batch_norm_train(x, gamma, beta, eps):
    """F.batch_norm(training=True) on [B, C, L]: biased variance over (B, L) incl. padded positions.

    Returns (y, mean, biased_var).  (modules/layers.py:78, modules/generated.py:94-96)
    """
    mean = x.mean(dim=(0, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2), keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    y = y * gamma.view(1, -1, 1) + beta.view(1, -1, 1)
    return y, mean.view(-1), var.view(-1)


def batch_norm_eval(x, gamma, beta, running_mean, running_var, eps):
    y = (x - running_mean.view(1, -1, 1)) / torch.sqrt(running_var.view(1, -1, 1) + eps)
    return y * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def running_stats_update(running_mean, running_var, mean, biased_var, count, momentum=0.1):
    """torch batch-norm running-stat rule: unbiased variance, momentum 0.1."""
    unbiased = biased_var * (count / max(count - 1, 1))
    return ((1 - momentum) * running_mean + momentum * mean,
            (1 - momentum) * running_var + momentum * unbiased)


# ----------------------------------------------------------------------------------------------
# LSTM cells  (modules/layers.py:18-47 over torch.nn.LSTMCell, gate order i, f, g, o)
# ----------------------------------------------------------------------------------------------

def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    z = _linear(x, w_ih, b_ih) + _linear(h, w_hh, b_hh)
    D = h.shape[1]
    i, f, g, o = z[:, :D], z[:, D:2 * D], z[:, 2 * D:3 * D], z[:, 3 * D:]
    c_new = _sigmoid(f) * c + _sigmoid(i) * torch.tanh(g)
    h_new = _sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def regularised_cell(kind, training, x, h, c, w, rates, keep_h=None, keep_c=None):
    """kind 'dropout' -> DropoutLSTMCell (layers.py:44-47); 'zoneout' -> ZoneoutLSTMCell (layers.py:26-34).

    w = (w_ih, w_hh, b_ih, b_hh);  rates = (rate_h, rate_c).
    """
    h_new, c_new = lstm_cell(x, h, c, *w)
    zh, zc = rates
    if kind == 'zoneout':
        if training:
            h_out = (1 - zh) * apply_dropout(h_new - h, keep_h, zh) + h if keep_h is not None else (1 - zh) * (h_new - h) + h
            c_out = (1 - zc) * apply_dropout(c_new - c, keep_c, zc) + c if keep_c is not None else (1 - zc) * (c_new - c) + c
        else:
            h_out = zh * h + (1 - zh) * h_new
            c_out = zc * c + (1 - zc) * c_new
        return h_out, c_out
    # dropout cell: only h is regularised, c passes through
    if training:
        h_new = apply_dropout(h_new, keep_h, zh)
    return h_new, c_new


# ----------------------------------------------------------------------------------------------
# Prenet (modules/tacotron2.py:37-46): 2 x (Linear -> ReLU -> dropout, ALWAYS on)
# ----------------------------------------------------------------------------------------------

def prenet(sd, prefix, x, p, keep0=None, keep1=None):
    keeps = [keep0, keep1]
    j = 0
    while f'{prefix}._layers.{j}.weight' in sd:
        x = _linear(x, sd[f'{prefix}._layers.{j}.weight'], sd[f'{prefix}._layers.{j}.bias'])
        x = torch.clamp(x, min=0.0)
        x = apply_dropout(x, keeps[j] if j < len(keeps) else None, p)
        j += 1
    return x


# ----------------------------------------------------------------------------------------------
# Location-sensitive attention (modules/attention.py:23-28, 39-45, 67-86)
# ----------------------------------------------------------------------------------------------

def attention_reset(sd, prefix, memory):
    """AttentionBase.reset: memT = memory . Wm^T; cum = 0; ctx = 0."""
    B, L, M = memory.shape
    memT = _q(_linear(memory, sd[f'{prefix}._memory.weight']))     # the perf mode also stores the projection in bf16
    cum = torch.zeros(B, L, dtype=memory.dtype)
    ctx = torch.zeros(B, M, dtype=memory.dtype)
    return memT, cum, ctx


def attention_step(sd, prefix, query, memory, memT, cum, mask):
    """One LocationSensitiveAttention.forward call.  Returns (context, weights, new_cum)."""
    Wq = sd[f'{prefix}._query.weight']                 # [A, D]
    Wc = sd[f'{prefix}._loc_features.weight']          # [C, 1, K]
    Wl = sd[f'{prefix}._location.weight']              # [A, C]
    bias = sd[f'{prefix}._bias']                       # [1, A]
    v = sd[f'{prefix}._energy.weight']                 # [1, A]
    B, L = cum.shape
    C, _, K = Wc.shape
    half = (K - 1) // 2
    q = query @ Wq.transpose(0, 1)                     # [B, A]  (kept in fp32 in both precision modes)
    cum_pad = torch.zeros(B, L + 2 * half, dtype=cum.dtype)
    cum_pad[:, half:half + L] = cum
    # f[b, c, l] = sum_k Wc[c, 0, k] * cum[b, l + k - half]   (zero outside [0, L))
    f = torch.zeros(B, C, L, dtype=cum.dtype)
    for k in range(K):
        f = f + Wc[:, 0, k].view(1, C, 1) * cum_pad[:, None, k:k + L]
    loc = f.transpose(1, 2) @ Wl.transpose(0, 1)       # [B, L, A]
    s = q[:, None, :] + memT + loc + bias.view(1, 1, -1)
    e = (torch.tanh(s) * v.view(1, 1, -1)).sum(dim=2)   # [B, L]
    e = torch.where(mask, e, torch.full_like(e, float('-inf')))
    e_max = e.max(dim=1, keepdim=True).values
    ex = torch.exp(e - e_max)
    w = ex / ex.sum(dim=1, keepdim=True)
    ctx = (w[:, :, None] * _q(memory)).sum(dim=1)      # bmm(w[B,1,L], memory[B,L,M])
    return ctx, w, cum + w


# ----------------------------------------------------------------------------------------------
# Decoder (modules/tacotron2.py:148-209)
# ----------------------------------------------------------------------------------------------

def _cell_weights(sd, prefix):
    return (sd[f'{prefix}.weight_ih'], sd[f'{prefix}.weight_hh'], sd[f'{prefix}.bias_ih'], sd[f'{prefix}.bias_hh'])


def decoder_memory(sd, hp, encoded, speaker, language, prefix='_decoder'):
    """tacotron2.py:143-146,158-161 -- concat speaker / language embeddings to the encoder output."""
    mem = encoded
    key = f'{prefix}._speaker_embedding.weight'
    if hp.multi_speaker and key in sd:
        mem = torch.cat((mem, sd[key][speaker]), dim=-1)
    key = f'{prefix}._language_embedding.weight'
    if hp.multi_language and key in sd:
        mem = torch.cat((mem, sd[key][language]), dim=-1)
    return mem


def decoder_forward(sd, hp, encoded, mask, target, speaker, language, tape, training=True, prefix='_decoder',
                    max_frames=None):
    """Decoder._decode.  target [B, N, T] or None (inference, B == 1).  Returns (spec[B,T,N], stop[B,T], align[B,T,L]).

    tape: see module docstring. tape['teacher'] must be given when target is not None.
    """
    tape = tape or {}
    dt = encoded.dtype
    B = encoded.shape[0]
    N = hp.num_mels
    D = hp.decoder_dimension
    kind = hp.decoder_regularization
    rates = (hp.zoneout_hidden, hp.zoneout_cell) if kind == 'zoneout' else (hp.dropout_hidden, 0.0)
    memory = decoder_memory(sd, hp, encoded, speaker, language, prefix)
    att = f'{prefix}._attention'
    memT, cum, ctx = attention_reset(sd, att, memory)
    h_att = torch.zeros(B, D, dtype=dt); c_att = torch.zeros(B, D, dtype=dt)
    h_gen = torch.zeros(B, D, dtype=dt); c_gen = torch.zeros(B, D, dtype=dt)
    frame = torch.zeros(B, N, dtype=dt)
    inference = target is None
    if not inference:
        T = target.shape[2]
        tgt = torch.cat((torch.zeros(B, 1, N, dtype=dt), target.transpose(1, 2)), dim=1)   # [B, T+1, N]
        tgt = prenet(sd, f'{prefix}._prenet', tgt, hp.dropout, tape.get('prenet0'), tape.get('prenet1'))
        teacher = tape['teacher']
    else:
        T = hp.max_output_length if max_frames is None else max_frames
    w_att = _cell_weights(sd, f'{prefix}._attention_lstm')
    w_gen = _cell_weights(sd, f'{prefix}._generator_lstm')
    Wf, bf = sd[f'{prefix}._frame_prediction.weight'], sd[f'{prefix}._frame_prediction.bias']
    Ws, bs = sd[f'{prefix}._stop_prediction.weight'], sd[f'{prefix}._stop_prediction.bias']

    def tm(name, i):
        t = tape.get(name)
        return None if t is None else t[i]

    specs, stops, aligns = [], [], []
    stop_frames = -1
    for i in range(T):
        if inference or not bool(teacher[i]):
            prev = prenet(sd, f'{prefix}._prenet', frame, hp.dropout, tm('step_prenet0', i), tm('step_prenet1', i))
        else:
            prev = tgt[:, i]
        h_att, c_att = regularised_cell(kind, training, torch.cat((prev, ctx), dim=1), h_att, c_att, w_att, rates,
                                        tm('att_h', i), tm('att_c', i))
        ctx, w, cum = attention_step(sd, att, h_att, memory, memT, cum, mask)
        h_gen, c_gen = regularised_cell(kind, training, torch.cat((h_att, ctx), dim=1), h_gen, c_gen, w_gen, rates,
                                        tm('gen_h', i), tm('gen_c', i))
        proto = torch.cat((h_gen, ctx), dim=1)
        frame = _linear(proto, Wf, bf)
        stop = _linear(proto, Ws, bs)
        specs.append(frame); stops.append(stop[:, 0]); aligns.append(w)
        if inference and bool(_sigmoid(stop)[0, 0] >= 0.5):
            # tacotron2.py:201-207: first hit arms a countdown of hp.stop_frames further hits
            if stop_frames == -1:
                stop_frames = hp.stop_frames
                continue
            stop_frames -= 1
            if stop_frames == 0:
                break
    return torch.stack(specs, dim=1), torch.stack(stops, dim=1), torch.stack(aligns, dim=1)


# ----------------------------------------------------------------------------------------------
# Conv blocks, postnet, vanilla encoder (modules/layers.py:66-86, tacotron2.py:66-76, encoder.py:30-45)
# ----------------------------------------------------------------------------------------------

def _act(name, x):
    if name == 'relu':
        return torch.clamp(x, min=0.0)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'sigmoid':
        return _sigmoid(x)
    return x


def conv_block(sd, prefix, x, activation, p, keep, training, stats=None, eps=1e-5):
    """ConvBlock: pad -> Conv1d(no bias) -> BatchNorm1d -> act -> Dropout(p).  x [B, C, L]."""
    w = sd[f'{prefix}._block.1.weight']
    y = conv1d_same(x, w)
    g, b = sd[f'{prefix}._block.2.weight'], sd[f'{prefix}._block.2.bias']
    if training:
        y, mean, var = batch_norm_train(y, g, b, eps)
        if stats is not None:
            stats[prefix] = (mean, var, y.shape[0] * y.shape[2])
    else:
        y = batch_norm_eval(y, g, b, sd[f'{prefix}._block.2.running_mean'], sd[f'{prefix}._block.2.running_var'], eps)
    y = _act(activation, y)
    if training:
        y = apply_dropout(y, keep, p)
    return y


def postnet(sd, hp, x, tape, training=True, prefix='_postnet', stats=None):
    """Postnet.forward (tacotron2.py:72-76): 5 conv blocks (tanh x4, identity) + residual."""
    tape = tape or {}
    residual = x
    nb = hp.postnet_blocks
    for j in range(nb):
        act = 'tanh' if j < nb - 1 else 'identity'
        x = conv_block(sd, f'{prefix}._convs.{j}', x, act, hp.dropout, tape.get(f'post{j}'), training, stats)
    return x + residual


def bilstm_packed(sd, prefix, x, lengths):
    """nn.LSTM(bidirectional, batch_first) on a packed sequence (encoder.py:41-44).

    x [B, L, E]; per-sample lengths; outputs beyond the length are exact zeros; the reverse
    direction starts at each sample's own last valid token.  Returns [B, L, 2H].
    """
    B, L, _ = x.shape
    outs = []
    for suffix in ('', '_reverse'):
        w_ih, w_hh = sd[f'{prefix}.weight_ih_l0{suffix}'], sd[f'{prefix}.weight_hh_l0{suffix}']
        b_ih, b_hh = sd[f'{prefix}.bias_ih_l0{suffix}'], sd[f'{prefix}.bias_hh_l0{suffix}']
        H = w_hh.shape[1]
        h = torch.zeros(B, H, dtype=x.dtype); c = torch.zeros(B, H, dtype=x.dtype)
        out = [None] * L
        order = range(L) if suffix == '' else range(L - 1, -1, -1)
        for l in order:
            valid = (l < lengths).to(x.dtype).view(B, 1)
            h_new, c_new = lstm_cell(x[:, l], h, c, w_ih, w_hh, b_ih, b_hh)
            h = valid * h_new + (1 - valid) * h
            c = valid * c_new + (1 - valid) * c
            out[l] = valid * h_new
        outs.append(torch.stack(out, dim=1))
    return torch.cat(outs, dim=2)


def vanilla_encoder(sd, hp, embedded, lengths, tape, training=True, prefix='_encoder', stats=None):
    """Encoder.forward (encoder.py:35-45)."""
    tape = tape or {}
    x = embedded.transpose(1, 2)
    for j in range(hp.encoder_blocks):
        x = conv_block(sd, f'{prefix}._convs.{j}', x, 'relu', hp.dropout, tape.get(f'enc{j}'), training, stats)
    return bilstm_packed(sd, f'{prefix}._lstm', x.transpose(1, 2), lengths)


# ----------------------------------------------------------------------------------------------
# Generated (language-grouped, highway) convolutional encoder
#   modules/generated.py:34-42, 71-96; modules/layers.py:124-131, 171-178; modules/encoder.py:180-221
# ----------------------------------------------------------------------------------------------

GENERATED_BLOCKS = ([(1, 1, 'relu', False), (1, 1, 'identity', False)]
                    + [(3, 3 ** i, 'identity', True) for i in range(4)]
                    + [(3, 3 ** i, 'identity', True) for i in range(4)]
                    + [(3, 1, 'identity', True)] * 2 + [(1, 1, 'identity', True)] * 2)


def generated_kernel(sd, prefix, e):
    """Conv1dGenerated weight generation: (e . Wb^T + bb) . Wk^T + bk  -> [G, Cout*Cin*k] (generated.py:38-39)."""
    eb = _linear(e, sd[f'{prefix}._bottleneck.weight'], sd[f'{prefix}._bottleneck.bias'])
    return _linear(eb, sd[f'{prefix}._kernel.weight'], sd[f'{prefix}._kernel.bias'])


def generated_conv_block(sd, prefix, e, x, G, k, dil, activation, highway, p, keep, training, stats=None, eps=1e-8):
    """ConvBlockGenerated / HighwayConvBlockGenerated on x [Q, G*Cin, L] -> [Q, G*Cout, L]."""
    Q, GC, L = x.shape
    Cin = GC // G
    kern = generated_kernel(sd, f'{prefix}._convolution', e)          # [G, Cout'*Cin*k]
    Coutp = kern.shape[1] // (Cin * k)
    ys = []
    for g in range(G):
        wg = kern[g].view(Coutp, Cin, k)
        ys.append(conv1d_same(x[:, g * Cin:(g + 1) * Cin], wg, dil))
    y = torch.cat(ys, dim=1)                                           # [Q, G*Cout', L]
    eb = _linear(e, sd[f'{prefix}._regularizer._bottleneck.weight'], sd[f'{prefix}._regularizer._bottleneck.bias'])
    aff = _linear(eb, sd[f'{prefix}._regularizer._affine.weight'], sd[f'{prefix}._regularizer._affine.bias'])
    gamma = aff[:, :Coutp].reshape(-1)
    beta = aff[:, Coutp:].reshape(-1)
    if training:
        y, mean, var = batch_norm_train(y, gamma, beta, eps)
        if stats is not None:
            stats[prefix] = (mean, var, Q * L)
    else:
        y = batch_norm_eval(y, gamma, beta, sd[f'{prefix}._regularizer.running_mean'],
                            sd[f'{prefix}._regularizer.running_var'], eps)
    y = _act(activation, y)
    if training:
        y = apply_dropout(y, keep, p)
    if not highway:
        return y
    C = Coutp // 2
    y = y.view(Q, G, 2, C, L)
    gate = _sigmoid(y[:, :, 0]).reshape(Q, G * C, L)
    val = y[:, :, 1].reshape(Q, G * C, L)
    return val * gate + x * (1.0 - gate)


def generated_encoder(sd, hp, embedded, tape, training=True, prefix='_encoder', language_weights=None, stats=None,
                      dropout=0.05):
    """GeneratedConvolutionalEncoder.forward (encoder.py:196-221).

    embedded [B, L, F] with B % G == 0 and sample b belonging to language b % G; or, for inference
    with ``language_weights`` [1, L, G], a single sample expanded to all groups and mixed per character.
    """
    tape = tape or {}
    G = sd[f'{prefix}._embedding.weight'].shape[0]
    x = embedded
    mixing = language_weights is not None and language_weights.shape[0] == 1
    if mixing:
        x = x.expand(G, -1, -1)
    e = sd[f'{prefix}._embedding.weight']                              # Embedding(arange(G))
    B, L, F = x.shape
    x = x.transpose(1, 2).reshape(B // G, G * F, L)
    for j, (k, dil, act, highway) in enumerate(GENERATED_BLOCKS):
        x = generated_conv_block(sd, f'{prefix}._layers.{j}', e, x, G, k, dil, act, highway, dropout,
                                 tape.get(f'enc{j}'), training, stats)
    x = x.reshape(B, -1, L).transpose(1, 2)
    if mixing:
        norm = language_weights / language_weights.sum(2, keepdim=True)[0]     # normaliser from position 0 only
        out = torch.zeros(1, L, x.shape[2], dtype=x.dtype)
        for g in range(G):
            out[0] = out[0] + norm[0, :, g].reshape(-1, 1) * x[g]
        x = out
    return x


# ----------------------------------------------------------------------------------------------
# Adversarial classifier (modules/classifier.py:6-69)
# ----------------------------------------------------------------------------------------------

class _GradReverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, clip):
        ctx.scale, ctx.clip = scale, clip
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.scale * g.clamp(-ctx.clip, ctx.clip), None, None


def reversal_classifier(sd, hp, encoded, prefix='_reversal_classifier', scale=1.0):
    x = _GradReverse.apply(encoded, scale, hp.reversal_gradient_clipping)
    x = _linear(x, sd[f'{prefix}._classifier.0.weight'], sd[f'{prefix}._classifier.0.bias'])
    return _linear(x, sd[f'{prefix}._classifier.1.weight'], sd[f'{prefix}._classifier.1.bias'])


# ----------------------------------------------------------------------------------------------
# Whole model (modules/tacotron2.py:355-385) and loss (tacotron2.py:439-485)
# ----------------------------------------------------------------------------------------------

def tacotron_forward(sd, hp, text, text_length, target, target_length, speakers, languages, tape, training=True,
                     stats=None):
    """Tacotron.forward; returns the reference's 6-tuple."""
    L = text.shape[1]
    if speakers is not None and speakers.dim() == 1:
        speakers = speakers[:, None].expand(-1, L)
    if languages is not None and languages.dim() == 1:
        languages = languages[:, None].expand(-1, L)
    embedded = sd['_embedding.weight'][text]
    if hp.encoder_type == 'generated':
        encoded = generated_encoder(sd, hp, embedded, tape, training, stats=stats)
    elif hp.encoder_type == 'simple':
        encoded = vanilla_encoder(sd, hp, embedded, text_length, tape, training, stats=stats)
    else:
        raise NotImplementedError(hp.encoder_type)
    spk_pred = reversal_classifier(sd, hp, encoded) if hp.reversal_classifier else None
    mask = lengths_to_mask(text_length, L)
    spec, stop, align = decoder_forward(sd, hp, encoded, mask, target, speakers, languages, tape, training)
    pre = spec.transpose(1, 2)
    post = postnet(sd, hp, pre, tape, training, stats=stats)
    tmask = lengths_to_mask(target_length, target.shape[2])
    stop = torch.where(tmask, stop, torch.full_like(stop, 1000.0))
    tm = tmask[:, None, :].to(pre.dtype)
    return post * tm, pre * tm, stop, align, spk_pred, encoded


def guided_attention_loss(align, input_lengths, target_lengths, g):
    """TacotronLoss._guided_attention (tacotron2.py:439-457), closed form instead of the meshgrid loop."""
    B, T, L = align.shape
    dt = align.dtype
    f = torch.arange(T, dtype=dt)[None, :, None]
    l = torch.arange(L, dtype=dt)[None, None, :]
    tl = target_lengths.to(dt)[:, None, None]
    il = input_lengths.to(dt)[:, None, None]
    w = 1.0 - torch.exp(-((l / il - f / tl) ** 2) / (2.0 * g * g))
    valid = (f < tl) & (l < il)
    w = torch.where(valid, w, torch.zeros_like(w))
    loss = (w * align).sum(dim=(1, 2))
    return (loss / target_lengths.to(dt)).mean()


def tacotron_loss(hp, g, text_length, target_length, pre, pre_target, post, post_target, stop, stop_target, align,
                  speaker=None, speaker_prediction=None, guided=True):
    """TacotronLoss.forward (tacotron2.py:459-485) for the 'reversal' classifier type."""
    dt = pre.dtype
    losses = {
        'mel_pre': 2.0 * ((pre - pre_target) ** 2).mean(),
        'mel_pos': ((post - post_target) ** 2).mean(),
    }
    # BCE-with-logits, pos_weight 100: -[100 * y * log s(x) + (1-y) * log(1 - s(x))]
    x, y = stop, stop_target.to(dt)
    log_sig = -torch.nn.functional.softplus(-x)
    log_one_minus = -torch.nn.functional.softplus(x)
    losses['stop_token'] = (-(100.0 * y * log_sig + (1 - y) * log_one_minus)).mean() / (hp.num_mels + 2)
    if hp.reversal_classifier and speaker_prediction is not None:
        B, L, S = speaker_prediction.shape
        ml = int(text_length.max())
        imask = lengths_to_mask(text_length, ml)
        logp = torch.log_softmax(speaker_prediction[:, :ml], dim=2)
        tgt = speaker[:, None].expand(B, ml)
        nll = -logp.gather(2, tgt[:, :, None])[:, :, 0]
        losses['lang_class'] = (nll * imask.to(dt)).sum() / imask.sum() * hp.reversal_classifier_w / (hp.num_mels + 2)
    if guided and hp.guided_attention_loss:
        losses['guided_att'] = guided_attention_loss(align, text_length, target_length, g)
    return sum(losses.values()), losses
