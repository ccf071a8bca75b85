"""Test-only differentiable restatement (PyTorch CPU, fp64) of the decoder training
forward + loss, used to obtain reference GRADIENTS by autograd.

It is validated against the NumPy oracle's forward (oracle/comic_oracle.py
`training_decode` + `caption_loss`) in tests/test_oracle_known_answers.py, so the
chain is: reference graph -> NumPy oracle (forward) == this module (forward) ->
autograd gradients -> CUDA backward.  Nothing here is imported by the product.

Follows: common/ops_rnn.py:183-243 (rnn_decoder_training), :531-565, :660-755;
src/model_base.py:325-417 (_train_caption_model / _loss_regularisation), :501-528,
:651-689.
"""
import numpy as np
import torch

DEC = 'Model/decoder/rnn_decoder/'
ATT = DEC + 'multi_add_attention/'


def _cell_scope(c):
    return (DEC + 'rnn_init_input/basic_lstm_cell/' if c.rnn_init_method == 'first_input'
            else DEC + 'basic_lstm_cell/')


def to_params(W, dtype=torch.float64):
    """name -> leaf tensor (requires_grad) for every decoder variable."""
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True)
            for k, v in W.items() if k.startswith('Model/decoder')}


def legacy_head(PH, mixed5c):
    """--legacy image embedding (src/model_base.py:80-91; oracle/inception_v1_oracle.encoder): tanh(LN(pool)) . W with
    pool = the 7x7 average of Mixed_5c, layer norm over the 1024 channels with epsilon 1e-12.  PH: leaf tensors of
    Model/encoder/LN_tanh/{gamma, beta} and Model/encoder/im_embed/weight."""
    ENC = 'Model/encoder/'
    g = PH[ENC + 'LN_tanh/gamma']
    pool = torch.as_tensor(np.asarray(mixed5c), dtype=g.dtype).mean(dim=(1, 2))
    mu = pool.mean(-1, keepdim=True)
    var = ((pool - mu) ** 2).mean(-1, keepdim=True)
    y = (pool - mu) / torch.sqrt(var + 1e-12) * g + PH[ENC + 'LN_tanh/beta']
    return torch.tanh(y) @ PH[ENC + 'im_embed/weight']


def _lstm(P, c, x, cp, hp):
    g = torch.cat([x, hp], 1) @ P[_cell_scope(c) + 'kernel'] + P[_cell_scope(c) + 'bias']
    i, j, f, o = g.chunk(4, 1)
    cn = cp * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
    hn = torch.tanh(cn) * torch.sigmoid(o)
    return cn, hn


def _drop(x, keep, mask):
    if mask is None or keep >= 1.0:
        return x
    return (x / keep) * mask


def training_loss(P, c, im_embed, fm, captions, masks=None, keeps=(1.0, 1.0, 1.0), rewards=None,
                  fm_requires_grad=False, dtype=torch.float64):
    """Returns (total, xe, map, reg, aux) as fp64 torch scalars; `aux` holds logits [T,B,V]
    (imputed) and attention maps [B,H,T_run,M].  captions [B,L] int (PAD = -1)."""
    dt = dtype
    # fm / im_embed may be torch tensors still attached to the CNN graph (cnn_finetune)
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), dtype=dt)
    H, R = c.attn_num_heads, c.rnn_size
    in_keep, out_keep, att_keep = keeps
    m = masks or {}
    cap = np.asarray(captions, np.int64)
    wmask = np.sign((cap[:, 1:] + 1).astype(np.float64))
    lens = wmask.sum(1).astype(np.int64)
    clipped = np.maximum(cap, 0)
    inputs = clipped[:, :-1] if c.token_type == 'word' else cap[:, :-1]
    targets = clipped[:, 1:]
    B, T = inputs.shape
    fm_t = t(fm)
    im_t = t(im_embed)
    if fm_requires_grad:
        fm_t.requires_grad_(True)
        im_t.requires_grad_(True)
    E = P[DEC + 'embedding_map']
    V = E.shape[0]
    keys = fm_t @ P[DEC + 'memory_layer/kernel']
    proj = c.cnn_fm_projection
    vals = keys if proj == 'tied' else (fm_t @ P[DEC + 'value_layer/kernel'] if proj == 'independent' else fm_t)
    M = keys.shape[1]
    dv = vals.shape[-1] // H
    # init state (first_input: model_base.py:675-686)
    if c.rnn_init_method == 'first_input':
        x0 = _drop(im_t @ P[DEC + 'rnn_init_input/projection/weight'], in_keep,
                   None if 'init_in' not in m else t(m['init_in']))
        z = torch.zeros((B, R), dtype=dt)
        cs, hs = _lstm(P, c, x0, z, z)
    else:
        hs = im_t @ P[DEC + 'rnn_initial_state/weight']
        cs = torch.zeros_like(hs)
    ctx = torch.zeros((B, R if c.attn_context_layer else vals.shape[-1]), dtype=dt)

    def embed(ids):
        ids = np.asarray(ids)
        if c.token_type == 'word':
            return E[torch.as_tensor(ids)]
        valid = (ids >= 0) & (ids < V)
        out = E[torch.as_tensor(np.where(valid, ids, 0))]
        return out * torch.as_tensor(valid.astype(np.float64)).to(dt)[:, None]

    T_run = int(lens.max()) if B else 0
    outs, hist = [], []
    for step in range(T_run):
        fin = torch.as_tensor(lens <= step)[:, None]
        x = torch.cat([embed(inputs[:, step]), ctx], 1)
        x = _drop(x, in_keep, None if 'inp' not in m else t(m['inp'][step]))
        cn, hn = _lstm(P, c, x, cs, hs)
        hout = _drop(hn, out_keep, None if 'out' not in m else t(m['out'][step]))
        logits = hout @ P[DEC + 'output_projection/kernel'] + P[DEC + 'output_projection/bias']
        q = hout @ P[ATT + 'query_layer/kernel']
        if c.attn_alignment_method == 'add_LN':
            u = keys + q[:, None, :]
            mu = u.mean(-1, keepdim=True)
            var = ((u - mu) ** 2).mean(-1, keepdim=True)
            y = (u - mu) / torch.sqrt(var + 1e-12) * P[ATT + 'LN_tanh/gamma'] + P[ATT + 'LN_tanh/beta']
            z = torch.tanh(y) * P[ATT + 'attention_v']
            s = z.reshape(B, M, H, R // H).sum(-1).permute(0, 2, 1) / P[DEC + 'softmax_temperature']
        else:                                                                 # MultiHeadDot, ops_rnn.py:603-632
            z = keys * q[:, None, :]
            s = z.reshape(B, M, H, R // H).sum(-1).permute(0, 2, 1) / float(np.sqrt(R / H))
        if c.attn_probability_fn == 'softmax':
            al = torch.softmax(s, -1)
        else:
            sg = torch.sigmoid(s)
            al = sg / sg.sum(-1, keepdim=True)
        al = _drop(al, att_keep, None if 'att' not in m else t(m['att'][step]))
        cnew = torch.einsum('nhm,nmhd->nhd', al, vals.reshape(B, M, H, dv)).reshape(B, -1)
        if c.attn_context_layer:                                              # ops_rnn.py:734-739
            cnew = cnew @ P[DEC + 'a_layer/kernel']
        hist.append(al)
        outs.append(torch.where(fin, torch.zeros_like(logits), logits))       # impute_finished
        cs = torch.where(fin, cs, cn)
        hs = torch.where(fin, hs, hn)
        ctx = torch.where(fin, ctx, cnew)
    logits = torch.stack(outs)                                                # [T_run,B,V]
    if T_run < T:
        logits = torch.cat([logits, logits[-1:].repeat(T - T_run, 1, 1)], 0)
    lb = logits.permute(1, 0, 2)                                              # [B,T,V]
    lp = torch.log_softmax(lb, -1)
    xent = -lp.gather(-1, torch.as_tensor(targets)[..., None])[..., 0] * t(wmask)
    if rewards is None:
        xe = xent.sum() / (t(wmask).sum() + 1e-12)
    else:
        per = xent.sum(1) / (t(wmask).sum(1) + 1e-12)
        xe = (per * t(rewards)).mean()
    am = torch.stack(hist)                                                    # [T_run,B,H,M]
    am = am.permute(1, 2, 0, 3)                                               # [B,H,T_run,M]
    map_loss = torch.zeros((), dtype=dt)
    if c.rnn_map_loss_scale > 0:
        map_loss = ((1.0 - am.sum(1)) ** 2).mean() * c.rnn_map_loss_scale     # axis 1 = heads (model_base.py:360)
    reg = torch.zeros((), dtype=dt)
    if c.l2_decay > 0:
        for v in P.values():
            reg = reg + (v ** 2).sum() / 2 * c.l2_decay
    total = xe + map_loss + reg
    return total, xe, map_loss, reg, dict(logits=logits, attn=am, fm=fm_t, im=im_t)


# ---------------------------------------------------------------------------
# Encoder (cnn_finetune): differentiable InceptionV1 restatement.
# Follows common/nets/inception_v1.py:29-339 under inception_arg_scope
# (common/nets/inception_utils.py:32-82) with is_training=False
# (src/model_base.py:71-77): conv (no bias, TF SAME) -> BN with MOVING statistics,
# no gamma, eps 1e-3 -> ReLU.  Trainable: conv kernels and BN betas.
# ---------------------------------------------------------------------------
CNN = 'Model/encoder/cnn/InceptionV1/'


def _same_pad(n, k, s):
    out = -(-n // s)
    pad = max((out - 1) * s + k - n, 0)
    return pad // 2, pad - pad // 2


def cnn_params(W, dtype=torch.float64):
    """Trainable CNN leaves (weights, betas) + constant moving statistics."""
    P = {}
    for k, v in W.items():
        if not k.startswith(CNN):
            continue
        tr = k.endswith('/weights') or k.endswith('/beta')
        P[k] = torch.tensor(np.asarray(v), dtype=dtype, requires_grad=tr)
    return P


def _cbr(P, x, scope, stride=1):
    import torch.nn.functional as F
    w = P[CNN + scope + '/weights']                                            # HWIO
    k = w.shape[0]
    pt, pb = _same_pad(x.shape[2], k, stride)
    pl, pr = _same_pad(x.shape[3], k, stride)
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w.permute(3, 2, 0, 1), stride=stride)
    s = torch.rsqrt(P[CNN + scope + '/BatchNorm/moving_variance'] + 1e-3)
    sh = P[CNN + scope + '/BatchNorm/beta'] - P[CNN + scope + '/BatchNorm/moving_mean'] * s
    return torch.relu(y * s[None, :, None, None] + sh[None, :, None, None])


def _maxpool(x, k, s):
    import torch.nn.functional as F
    pt, pb = _same_pad(x.shape[2], k, s)
    pl, pr = _same_pad(x.shape[3], k, s)
    return F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float('-inf')), k, s)


def encoder_forward(P, images):
    """images [B,224,224,3] NHWC -> (im_embed [B,1024], fm [B,196,832]) differentiable in P."""
    from comic_b200 import weights as wts
    x = torch.as_tensor(np.asarray(images), dtype=next(iter(P.values())).dtype).permute(0, 3, 1, 2)
    x = _cbr(P, x, 'Conv2d_1a_7x7', 2)
    x = _maxpool(x, 3, 2)
    x = _cbr(P, x, 'Conv2d_2b_1x1')
    x = _cbr(P, x, 'Conv2d_2c_3x3')
    x = _maxpool(x, 3, 2)
    fm = None
    for b in wts.BLOCKS:
        if len(b) == 3:
            x = _maxpool(x, b[1], b[2])
            continue
        sc = wts.block_conv_scopes(b[0])
        b0 = _cbr(P, x, sc[0])
        b1 = _cbr(P, _cbr(P, x, sc[1]), sc[2])
        b2 = _cbr(P, _cbr(P, x, sc[3]), sc[4])
        b3 = _cbr(P, _maxpool(x, 3, 1), sc[5])
        x = torch.cat([b0, b1, b2, b3], 1)
        if b[0] == 'Mixed_4f':
            fm = x.permute(0, 2, 3, 1).reshape(x.shape[0], 196, 832)
    im_embed = x.mean(dim=(2, 3))
    return im_embed, fm
