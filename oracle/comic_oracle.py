"""ORACLE (test infrastructure, not product code) -- COMIC decoder, beam search,
greedy search, teacher-forced decode and losses, restated in NumPy.

Follows the reference graph builders
  common/ops_rnn.py:49-112   rnn_decoder_beam_search
  common/ops_rnn.py:115-180  rnn_decoder_search
  common/ops_rnn.py:183-243  rnn_decoder_training
  common/ops_rnn.py:246-280  split_heads / combine_heads
  common/ops_rnn.py:403-565  MultiHeadAttV3 / MultiHeadAddLN
  common/ops_rnn.py:603-632  MultiHeadDot
  common/ops_rnn.py:635-803  MultiHeadAttentionWrapperV3
  common/ops_rnn.py:807-845  BeamSearchDecoderMultiHead
  src/model_base.py:109-314  _decoder_rnn / _decoder_post_process
  src/model_base.py:325-417  _train_caption_model / _loss_regularisation
  src/model_base.py:501-594  _process_inputs / embeddings
  src/model_base.py:599-757  _signorm / _get_rnn_init / _rnn_dynamic_decoder
and, because the arithmetic underneath lives in an un-vendored third party --
TensorFlow 1.9.0 (README.md:48; tf.contrib.seq2seq BeamSearchDecoder,
dynamic_decode, BasicDecoder, helpers, gather_tree; tf.contrib.rnn
BasicLSTMCell / DropoutWrapper; tf.contrib.layers layer_norm / dropout) --
restates TF r1.9's published algorithms for those ops.

PARITY UNPINNED: the reference ships no test, golden vector or fixture for the
decoder / beam search / loss, and TF 1.9 + Python 2.7 cannot run here.  What
pins this oracle is listed in tests/test_oracle_known_answers.py (parameter
counts from README.md:219-233, structural invariants of beam search, an
independent PyTorch-CPU eager cross-check, fp64 finite differences for the
backward).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import numpy as np

DEC = 'Model/decoder/rnn_decoder/'
ATT = DEC + 'multi_add_attention/'


def number_to_base(n, base):
    """common/ops.py:25-40."""
    if base < 2:
        raise ValueError('Base cannot be less than 2.')
    if n < 0:
        sign = -1
        n *= sign
    elif n == 0:
        return [0]
    else:
        sign = 1
    digits = []
    while n:
        digits.append(sign * int(n % base))
        n //= base
    return digits[::-1]


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def layer_norm(x, gamma, beta, eps=1e-12):
    """contrib layer_norm, last axis, biased variance, eps 1e-12
    (common/ops.py:241-275)."""
    dt = x.dtype
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    inv = (1.0 / np.sqrt(var + dt.type(eps))) * gamma
    return x * inv + (beta - mean * inv)


def split_heads(x, H):
    """[N, M, C] -> [N, H, M, C/H]  (common/ops_rnn.py:246-261)."""
    N, M, C = x.shape
    return x.reshape(N, M, H, C // H).transpose(0, 2, 1, 3)


def combine_heads(x):
    """[N, H, L, d] -> [N, L, H*d]  (common/ops_rnn.py:264-280)."""
    N, H, L, d = x.shape
    return x.transpose(0, 2, 1, 3).reshape(N, L, H * d)


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def log_softmax(x):
    """tf.nn.log_softmax: shifted = x - max; shifted - log(sum(exp(shifted)))."""
    shifted = x - x.max(axis=-1, keepdims=True)
    return shifted - np.log(np.exp(shifted).sum(axis=-1, keepdims=True))


def dropout(x, keep, mask):
    """tf.nn.dropout with an explicit 0/1 mask: div(x, keep) * mask."""
    if mask is None or keep >= 1.0:
        return x
    return (x / x.dtype.type(keep)) * mask.astype(x.dtype)


def tile_batch(x, k):
    """tf.contrib.seq2seq.tile_batch: repeat-interleave on axis 0."""
    return np.repeat(x, k, axis=0)


class Decoder(object):
    """The attention-LSTM decoder cell of one sequence batch
    (MultiHeadAttentionWrapperV3 over BasicLSTMCell, with the mechanism's
    memory already bound)."""

    def __init__(self, W, config, dtype=np.float32):
        c = self.c = config
        self.dt = np.dtype(dtype)
        self.W = {k: np.asarray(v, dtype=self.dt) for k, v in W.items() if k.startswith('Model/decoder')}
        self.R = c.rnn_size
        self.H = c.attn_num_heads
        self.Wd = c.rnn_word_size
        if c.token_type == 'radix':
            self.V = c.radix_base + 2
            self.go, self.eos = c.radix_base, c.radix_base + 1       # model_base.py:701-703
        else:
            self.V = len(c.itow)
            self.go, self.eos = c.wtoi['<GO>'], c.wtoi['<EOS>']      # :705-706
        if c.attn_alignment_method not in ('add_LN', 'dot'):
            raise ValueError('Invalid alignment method.')            # model_base.py:133-138
        if c.attn_probability_fn not in ('softmax', 'sigmoid'):
            raise ValueError('Invalid alignment method.')            # :140-145
        if c.rnn_name != 'LSTM':
            raise NotImplementedError('oracle restates rnn_name=LSTM only')
        self.cell_scope = (DEC + 'rnn_init_input/basic_lstm_cell/'
                           if c.rnn_init_method == 'first_input' else DEC + 'basic_lstm_cell/')

    # -- D0: MultiHeadAttV3.__init__ (ops_rnn.py:408-477) --------------------
    def setup_memory(self, fm):
        """fm [N, M, C] (already tiled by beam, model_base.py:130-131)."""
        fm = fm.astype(self.dt)
        self.keys = fm @ self.W[DEC + 'memory_layer/kernel']        # [N,M,R]
        proj = self.c.cnn_fm_projection
        if proj == 'tied':
            vals = self.keys
        elif proj == 'independent':
            vals = fm @ self.W[DEC + 'value_layer/kernel']
        else:
            vals = fm
        self.values_split = split_heads(vals, self.H)                # [N,H,M,dv]
        self.N, self.M = fm.shape[0], fm.shape[1]
        self.A = (fm.shape[-1] if (proj is None and not self.c.attn_context_layer) else self.R)

    # -- D3: BasicLSTMCell ----------------------------------------------------
    def lstm(self, x, c_prev, h_prev):
        K = self.W[self.cell_scope + 'kernel']
        b = self.W[self.cell_scope + 'bias']
        g = np.concatenate([x, h_prev], axis=1) @ K + b
        i, j, f, o = np.split(g, 4, axis=1)
        c_new = c_prev * sigmoid(f + self.dt.type(1.0)) + sigmoid(i) * np.tanh(j)
        h_new = np.tanh(c_new) * sigmoid(o)
        return c_new, h_new

    # -- D1: _get_rnn_init (model_base.py:651-689) ----------------------------
    def init_state(self, im_embed, in_mask=None, in_keep=1.0):
        e = im_embed.astype(self.dt)
        N = e.shape[0]
        if self.c.rnn_init_method == 'project_hidden':
            h = e @ self.W[DEC + 'rnn_initial_state/weight']
            c = np.zeros_like(h)
        elif self.c.rnn_init_method == 'first_input':
            x = e @ self.W[DEC + 'rnn_init_input/projection/weight']
            x = dropout(x, in_keep, in_mask)            # DropoutWrapper'd cell in train mode
            z = np.zeros((N, self.R), self.dt)
            c, h = self.lstm(x, z, z)
        else:
            raise ValueError('Invalid RNN init method specified.')
        return c, h

    # -- D2: zero_state (ops_rnn.py:776-803) ----------------------------------
    def zero_state(self, cell_state):
        c, h = cell_state
        N = c.shape[0]
        return dict(c=c, h=h, attention=np.zeros((N, self.A), self.dt),
                    alignments=np.zeros((N, self.H * self.M), self.dt), time=0)

    # -- D4: the attention mechanism call -------------------------------------
    def alignments(self, query):
        q = query @ self.W[ATT + 'query_layer/kernel']               # ops_rnn.py:545
        if self.c.attn_alignment_method == 'add_LN':
            s = self.keys + q[:, None, :]                            # :548
            s = np.tanh(layer_norm(s, self.W[ATT + 'LN_tanh/gamma'], self.W[ATT + 'LN_tanh/beta']))  # :549
            s = s * self.W[ATT + 'attention_v']                      # :550
            s = split_heads(s, self.H).sum(axis=3)                   # :551-552  [N,H,M]
            s = s / self.W[DEC + 'softmax_temperature']              # :554-562
        else:                                                        # MultiHeadDot :603-632
            s = self.keys * q[:, None, :]
            s = split_heads(s, self.H).sum(axis=3)
            s = s / np.sqrt(self.dt.type(self.R / self.H))
        if self.c.attn_probability_fn == 'softmax':
            return softmax(s, axis=-1)
        sg = sigmoid(s)                                              # _signorm model_base.py:599-603
        return sg / sg.sum(axis=-1, keepdims=True)

    # -- D3-D6: MultiHeadAttentionWrapperV3.call (ops_rnn.py:660-755) ---------
    def call(self, inputs, state, in_mask=None, out_mask=None, att_mask=None,
             in_keep=1.0, out_keep=1.0, att_keep=1.0):
        x = np.concatenate([inputs.astype(self.dt), state['attention']], axis=1)   # :673
        x = dropout(x, in_keep, in_mask)
        c_new, h_new = self.lstm(x, state['c'], state['h'])                        # :675
        cell_output = dropout(h_new, out_keep, out_mask)
        al = self.alignments(cell_output)                                          # :692-694 [N,H,M]
        al = dropout(al, att_keep, att_mask)                                       # :696-701
        ctx = np.einsum('nhm,nhmd->nhd', al, self.values_split)                    # :703-715
        attention = ctx.reshape(ctx.shape[0], -1)                                  # combine_heads :716
        if self.c.attn_context_layer:
            attention = attention @ self.W[DEC + 'a_layer/kernel']                 # :734-739
        al_flat = al.reshape(al.shape[0], -1)                                      # :742
        new_state = dict(c=c_new, h=h_new, attention=attention, alignments=al_flat,
                         time=state['time'] + 1)
        return cell_output, new_state, al_flat

    def output_layer(self, h):
        return h @ self.W[DEC + 'output_projection/kernel'] + self.W[DEC + 'output_projection/bias']

    # -- D8: embeddings (model_base.py:557-594) -------------------------------
    def embed(self, ids):
        E = self.W[DEC + 'embedding_map']
        ids = np.asarray(ids)
        if self.c.token_type == 'word':
            return E[ids]                               # embedding_lookup
        valid = (ids >= 0) & (ids < self.V)             # one_hot -> zero row off-range
        out = E[np.where(valid, ids, 0)]
        return out * valid[..., None].astype(self.dt)

    def max_iterations(self):
        """model_base.py:709-714."""
        c = self.c
        m = c.infer_max_length
        if c.token_type == 'radix':
            m *= len(number_to_base(len(c.wtoi), c.radix_base))
        elif c.token_type == 'char':
            m *= 5
        return m


# ---------------------------------------------------------------------------
# TF r1.9 beam_search_ops.gather_tree (CPU functor) and
# beam_search_decoder.gather_tree_from_array
# ---------------------------------------------------------------------------
def gather_tree(step_ids, parent_ids, max_sequence_lengths, end_token):
    T, B, K = parent_ids.shape
    beams = np.full((T, B, K), end_token, dtype=step_ids.dtype)
    for b in range(B):
        L = min(T, int(max_sequence_lengths[b]))
        if L <= 0:
            continue
        for k in range(K):
            beams[L - 1, b, k] = step_ids[L - 1, b, k]
            parent = parent_ids[L - 1, b, k]
            for level in range(L - 2, -1, -1):
                if parent < 0 or parent > K:
                    raise ValueError('Saw invalid parent id %d' % parent)
                beams[level, b, k] = step_ids[level, b, parent]
                parent = parent_ids[level, b, parent]
            finished = False
            for t in range(L):
                if finished:
                    beams[t, b, k] = end_token
                elif beams[t, b, k] == end_token:
                    finished = True
    return beams


def gather_tree_from_array(t, parent_ids, sequence_length):
    """t [T, B*K, D] (or [T,B,K,D]); parent_ids [T,B,K]; sequence_length [B,K].
    Out-of-range sorted ids (the beam_width+1 sentinel surviving the final
    `where`) gather zeros, as tf.gather_nd does on GPU."""
    T, B, K = parent_ids.shape
    beam_ids = np.tile(np.arange(K, dtype=np.int32)[None, None, :], (T, B, 1))
    mask = (np.arange(T)[:, None, None] < sequence_length[None, :, :]).astype(np.int32)
    masked = beam_ids * mask + (1 - mask) * (K + 1)
    max_len = sequence_length.max(axis=1).astype(np.int32)
    sorted_ids = gather_tree(masked, parent_ids, max_len, K + 1)
    sorted_ids = np.where(mask.astype(bool), sorted_ids, beam_ids)
    src = t.reshape(T, B, K, -1)
    out = np.zeros_like(src)
    for tt in range(T):
        for b in range(B):
            for k in range(K):
                s = sorted_ids[tt, b, k]
                if 0 <= s < K:
                    out[tt, b, k] = src[tt, b, s]
    return out.reshape(t.shape), sorted_ids


def length_penalty(lengths, w):
    """_length_penalty: static 0 -> 1.0; else (5+len)^w / 6^w in fp32."""
    if w == 0:
        return np.float32(1.0)
    w = np.float32(w)
    return np.power(np.float32(5.0) + lengths.astype(np.float32), w) / np.power(np.float32(6.0), w)


def beam_search_step(logits, log_probs, finished, lengths, k, eos, lpw):
    """TF r1.9 `_beam_search_step` on logits [B,k,V].  Returns
    (scores[B,k], word[B,k] i32, parent[B,k] i32, new_log_probs, new_finished,
    new_lengths i64, total_probs)."""
    B, K, V = logits.shape
    dt = logits.dtype
    lp = log_softmax(logits)
    fin_row = np.full((V,), np.finfo(dt).min, dtype=dt)
    fin_row[eos] = 0
    lp = np.where(finished[:, :, None], fin_row[None, None, :], lp)      # _mask_probs
    total = log_probs[:, :, None] + lp
    add = np.ones((V,), np.int64)
    add[eos] = 0
    new_len = add[None, None, :] * (~finished)[:, :, None].astype(np.int64) + lengths[:, :, None]
    pen = length_penalty(new_len, lpw)
    scores = total / pen if lpw != 0 else total
    flat = scores.reshape(B, K * V)
    # nn.top_k: sorted descending; equal values keep the lower index first.
    idx = np.argsort(-flat, axis=1, kind='stable')[:, :k]
    # argsort of -x treats -0.0/0.0 and NaN like TF does not matter here (no NaN).
    top = np.take_along_axis(flat, idx, axis=1)
    word = (idx % V).astype(np.int32)
    parent = (idx // V).astype(np.int32)
    new_lp = np.take_along_axis(total.reshape(B, K * V), idx, axis=1)
    prev_fin = np.take_along_axis(finished, parent, axis=1)
    new_fin = prev_fin | (word == eos)
    new_lengths = np.take_along_axis(lengths, parent, axis=1) + (~prev_fin).astype(np.int64)
    return top, word, parent, new_lp, new_fin, new_lengths, total


def beam_search_decode(dec, im_embed, fm, beam, lpw=0.0, max_it=None, return_trace=False):
    """rnn_decoder_beam_search (ops_rnn.py:49-112) + dynamic_decode
    (impute_finished=False) + BeamSearchDecoder.finalize with
    reorder_tensor_arrays=True.

    im_embed [B,E], fm [B,M,C] (untiled).  Returns dict with
      predicted_ids [T,B,k] i32, scores [T,B,k], parent_ids [T,B,k] i32,
      step_ids [T,B,k] i32 (pre gather_tree), lengths [B,k] i64,
      alignment_history [T,B*k,H*M] (reordered), raw_history (unreordered)."""
    B = im_embed.shape[0]
    k = beam
    if max_it is None:
        max_it = dec.max_iterations()
    dec.setup_memory(tile_batch(fm, k))                                  # model_base.py:130-131
    state = dec.zero_state(dec.init_state(tile_batch(im_embed, k)))
    dt = dec.dt
    log_probs = np.full((B, k), -np.inf, dt); log_probs[:, 0] = 0
    finished = np.ones((B, k), bool); finished[:, 0] = False
    lengths = np.zeros((B, k), np.int64)
    inputs = dec.embed(np.full((B * k,), dec.go, np.int32))
    out_scores, out_ids, out_par, hist, trace = [], [], [], [], []
    t = 0
    loop_finished = finished.copy()
    if max_it <= 0:
        loop_finished[:] = True
    while not loop_finished.all():
        cell_out, new_state, al = dec.call(inputs, state)
        hist.append(al)                                                  # written at index `time`
        logits = dec.output_layer(cell_out).reshape(B, k, -1)
        top, word, parent, log_probs, finished, lengths, total = beam_search_step(
            logits, log_probs, finished, lengths, k, dec.eos, lpw)
        if return_trace:
            trace.append(dict(logits=logits, total=total, h=cell_out, attention=new_state['attention']))
        gidx = (np.arange(B)[:, None] * k + parent).reshape(-1)          # state gather by parent
        state = dict(c=new_state['c'][gidx], h=new_state['h'][gidx],
                     attention=new_state['attention'][gidx],
                     alignments=new_state['alignments'][gidx], time=new_state['time'])
        out_scores.append(top); out_ids.append(word); out_par.append(parent)
        if finished.all():
            inputs = dec.embed(np.full((B * k,), dec.go, np.int32))
        else:
            inputs = dec.embed(word.reshape(-1))
        t += 1
        loop_finished = finished | (t >= max_it)                         # tracks_own_finished
    T = t
    step_ids = np.stack(out_ids) if T else np.zeros((0, B, k), np.int32)
    parents = np.stack(out_par) if T else np.zeros((0, B, k), np.int32)
    scores = np.stack(out_scores) if T else np.zeros((0, B, k), dt)
    raw_hist = np.stack(hist) if T else np.zeros((0, B * k, dec.H * dec.M), dt)
    max_len = lengths.max(axis=1).astype(np.int32)
    predicted = gather_tree(step_ids, parents, max_len, dec.eos)
    hist_sorted, sorted_ids = gather_tree_from_array(raw_hist, parents, lengths)
    out = dict(predicted_ids=predicted, scores=scores, parent_ids=parents, step_ids=step_ids,
               lengths=lengths, alignment_history=hist_sorted, raw_history=raw_hist,
               sorted_beam_ids=sorted_ids, final_state=state, T=T)
    if return_trace:
        out['trace'] = trace
    return out


def greedy_decode(dec, im_embed, fm, max_it=None):
    """rnn_decoder_search(greedy_search=True) (ops_rnn.py:115-180):
    GreedyEmbeddingHelper + BasicDecoder + dynamic_decode(impute_finished=False).
    Outputs are NOT masked after EOS."""
    B = im_embed.shape[0]
    if max_it is None:
        max_it = dec.max_iterations()
    dec.setup_memory(fm)
    state = dec.zero_state(dec.init_state(im_embed))
    finished = np.zeros((B,), bool)
    if max_it <= 0:
        finished[:] = True
    inputs = dec.embed(np.full((B,), dec.go, np.int32))
    ids, logits_l, hist = [], [], []
    t = 0
    while not finished.all():
        cell_out, state, al = dec.call(inputs, state)
        hist.append(al)
        logits = dec.output_layer(cell_out)
        sample = logits.argmax(axis=-1).astype(np.int32)                 # first max
        dec_fin = sample == dec.eos
        inputs = (dec.embed(np.full((B,), dec.go, np.int32)) if dec_fin.all() else dec.embed(sample))
        ids.append(sample); logits_l.append(logits)
        t += 1
        finished = dec_fin | finished | (t >= max_it)
    T = t
    return dict(ids=np.stack(ids) if T else np.zeros((0, B), np.int32),
                logits=np.stack(logits_l) if T else np.zeros((0, B, dec.V), dec.dt),
                alignment_history=np.stack(hist) if T else np.zeros((0, B, dec.H * dec.M), dec.dt),
                final_state=state, T=T)


def process_inputs(captions, token_type):
    """_process_inputs (model_base.py:501-528): captions [B,L] int32 PAD=-1 ->
    (inputs [B,L-1], targets [B,L-1], masks [B,L-1] f32, lens [B] i32)."""
    cap = np.asarray(captions, np.int32)
    masks = np.sign((cap[:, 1:] + 1).astype(np.float32))
    lens = masks.sum(axis=1).astype(np.int32)
    clipped = np.maximum(cap, 0)
    if token_type == 'word':
        inputs = clipped[:, :-1]
    else:
        inputs = cap[:, :-1]
    targets = clipped[:, 1:]
    return inputs, targets, masks, lens


def training_decode(dec, im_embed, fm, dec_inputs, lens, masks=None, keeps=(1.0, 1.0, 1.0),
                    save=False):
    """rnn_decoder_training (ops_rnn.py:183-243): TrainingHelper + BasicDecoder
    + dynamic_decode(impute_finished=True).  dec_inputs [B,T] ids.

    masks: optional dict(init_in [B,W+A], inp [T,B,W+A], out [T,B,R],
    att [T,B,H,M]) of 0/1 dropout masks; keeps = (in_keep, out_keep, att_keep).
    Returns logits [T,B,V] (padded by repeating the last executed step,
    ops_rnn.py:237-241), ids [T,B], alignment_history [T_run,B,H*M]."""
    B, T = dec_inputs.shape
    in_keep, out_keep, att_keep = keeps
    m = masks or {}
    dec.setup_memory(fm)
    emb = dec.embed(dec_inputs).transpose(1, 0, 2)                       # [T,B,W] model_base.py:587-593
    state = dec.zero_state(dec.init_state(im_embed, m.get('init_in'), in_keep))
    lens = np.asarray(lens, np.int32)
    finished = (lens == 0)
    zero_in = np.zeros_like(emb[0])
    inputs = zero_in if finished.all() else emb[0]
    outs, ids, hist, saved = [], [], [], []
    t = 0
    while not finished.all():
        cell_out, new_state, al = dec.call(
            inputs, state,
            in_mask=None if 'inp' not in m else m['inp'][t],
            out_mask=None if 'out' not in m else m['out'][t],
            att_mask=None if 'att' not in m else m['att'][t],
            in_keep=in_keep, out_keep=out_keep, att_keep=att_keep)
        logits = dec.output_layer(cell_out)
        sample = logits.argmax(axis=-1).astype(np.int32)
        next_fin = (t + 1 >= lens)
        nxt = zero_in if next_fin.all() else emb[min(t + 1, T - 1)]
        # impute_finished=True: zero outputs / copy state through for rows already finished
        f = finished[:, None]
        outs.append(np.where(f, 0, logits).astype(dec.dt))
        ids.append(np.where(finished, 0, sample).astype(np.int32))
        hist.append(al)                                                  # TensorArray passes through
        state = dict(c=np.where(f, state['c'], new_state['c']),
                     h=np.where(f, state['h'], new_state['h']),
                     attention=np.where(f, state['attention'], new_state['attention']),
                     alignments=np.where(f, state['alignments'], new_state['alignments']),
                     time=new_state['time'])
        inputs = nxt
        t += 1
        finished = next_fin | finished
    T_run = t
    logits = np.stack(outs)
    idsa = np.stack(ids)
    if T_run < T:                                                        # ops_rnn.py:237-241
        logits = np.concatenate([logits, np.tile(logits[-1:], (T - T_run, 1, 1))], axis=0)
        idsa = np.concatenate([idsa, np.tile(idsa[-1:], (T - T_run, 1))], axis=0)
    return dict(logits=logits, ids=idsa, alignment_history=np.stack(hist), T_run=T_run)


def post_process_beam(res, H, k, top_beam=True):
    """_decoder_post_process, beam branch (model_base.py:277-288, 296-313)."""
    pred, scores = res['predicted_ids'], res['scores']
    if top_beam:
        output_ids = pred[:, :, 0].T                                     # [B,T]
        logits = scores[:, :, 0].T
    else:
        output_ids = pred.transpose(2, 1, 0)                             # [k,B,T]
        logits = scores.transpose(2, 1, 0)
    hist = res['alignment_history']                                      # [T, B*k, H*M]
    T = hist.shape[0]
    am = hist.reshape(T, -1, k, hist.shape[2])[:, :, 0, :]               # top beam
    am = am.reshape(T, am.shape[1], H, -1).transpose(1, 2, 0, 3)         # [B,H,T,M]
    return logits, output_ids, am


def post_process_plain(logits, ids, hist, H):
    """_decoder_post_process, greedy/train branch (model_base.py:289-293, 307-313)."""
    T = hist.shape[0]
    am = hist.reshape(T, hist.shape[1], H, -1).transpose(1, 2, 0, 3)
    return logits.transpose(1, 0, 2), ids.T, am


def sequence_loss(logits, targets, weights, average_across_batch=True):
    """tf.contrib.seq2seq.sequence_loss with batch-major logits [B,T,V]."""
    lp = log_softmax(logits)
    xent = -np.take_along_axis(lp, targets[..., None], axis=-1)[..., 0]
    xent = xent * weights
    eps = logits.dtype.type(1e-12)
    if average_across_batch:
        return xent.sum() / (weights.sum() + eps)
    return xent.sum(axis=1) / (weights.sum(axis=1) + eps)


def caption_loss(logits_bt, targets, masks, attn_maps, W_train, config, rewards=None):
    """_train_caption_model (model_base.py:325-383): returns (total, xe, map, reg)."""
    dt = logits_bt.dtype
    if rewards is None:
        xe = sequence_loss(logits_bt, targets, masks.astype(dt))
    else:
        per = sequence_loss(logits_bt, targets, masks.astype(dt), average_across_batch=False)
        xe = (per * rewards.astype(dt)).mean()
    map_loss = dt.type(0)
    if config.rnn_map_loss_scale > 0:
        flat = attn_maps.sum(axis=1)                       # axis 1 of [B,H,T,M] = heads (model_base.py:360)
        map_loss = ((1.0 - flat) ** 2).mean() * dt.type(config.rnn_map_loss_scale)
    reg = dt.type(0)
    if config.l2_decay > 0:
        for v in W_train.values():
            reg = reg + (np.asarray(v, dt) ** 2).sum() / 2 * dt.type(config.l2_decay)
    return xe + map_loss + reg, xe, map_loss, reg


def cosine_lr(step, max_step, lr_start, lr_end):
    """_create_cosine_lr (model_base.py:809-820)."""
    s = np.float32(step / max_step)
    s = np.float32(1.0) + np.cos(np.minimum(np.float32(1.0), s) * np.float32(np.pi))
    return np.float32((lr_start - lr_end) * s / 2 + lr_end)


def adam_step(theta, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-2):
    """tf.train.AdamOptimizer dense update (epsilon-hat form), step t >= 1."""
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m = m + (g - m) * (1 - beta1)
    v = v + (g * g - v) * (1 - beta2)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v
