"""GPU parity of the training path (csrc/train.cu through the C ABI): losses and every
decoder-variable gradient against fp64 autograd of the torch restatement
(tests/torch_ref.py, itself pinned to the NumPy oracle's forward on CPU).
Tolerance 1e-3 relative to the largest entry of each gradient tensor."""
import numpy as np
import pytest

from _common import comic_config, word_config, make_weights, fake_features, rel_err
from test_oracle_known_answers import _train_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch_mod():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch


def _reference(c, W, im, fm, caps, masks, keeps, rewards):
    import torch_ref as TR
    P = TR.to_params(W)
    tot, xe, mp, reg, aux = TR.training_loss(P, c, im, fm, caps, masks, keeps, rewards)
    tot.backward()
    grads = {k: v.grad.numpy() for k, v in P.items()}
    return dict(total=float(tot), xe=float(xe), map=float(mp), reg=float(reg), grads=grads,
                logits=aux['logits'].detach().numpy(), attn=aux['attn'].detach().numpy())


CASES = {
    'comic256_xe_dropout': (lambda: comic_config(train_mode='decoder'), True, False),
    'comic256_scst_dropout': (lambda: comic_config(train_mode='scst'), True, True),
    'comic256_xe_nodrop': (lambda: comic_config(train_mode='decoder'), False, False),
    'word_none_h1_xe': (lambda: word_config(n_words=300, train_mode='decoder'), True, False),
    'independent_h4': (lambda: comic_config(train_mode='decoder', cnn_fm_projection='independent', attn_num_heads=4),
                       True, False),
    # the cell variants of src/model_base.py:599-603, 622-667 and common/ops_rnn.py:603-632, 734-739
    'dot': (lambda: comic_config(train_mode='decoder', attn_alignment_method='dot'), True, False),
    'sigmoid': (lambda: comic_config(train_mode='decoder', attn_probability_fn='sigmoid'), True, False),
    'context_layer': (lambda: comic_config(train_mode='decoder', attn_context_layer=True), True, False),
    'project_hidden': (lambda: comic_config(train_mode='decoder', rnn_init_method='project_hidden'), True, False),
    'dot_sigmoid_ctx_ph_indep_h4': (lambda: comic_config(train_mode='scst', attn_alignment_method='dot',
                                                         attn_probability_fn='sigmoid', attn_context_layer=True,
                                                         rnn_init_method='project_hidden',
                                                         cnn_fm_projection='independent', attn_num_heads=4), True, True),
    'word_none_ctx_sigmoid': (lambda: word_config(n_words=300, train_mode='decoder', attn_context_layer=True,
                                                  attn_probability_fn='sigmoid'), True, False),
}


@pytest.mark.parametrize('name', list(CASES))
def test_gradients_match_autograd(torch_mod, name):
    from comic_b200.train import Trainer
    from comic_b200 import weights as wts
    mk, dropout, scst = CASES[name]
    c = mk()
    W, im, fm, caps, masks, keeps = _train_case(c, B=4, L=8, seed=3, dropout=dropout)
    rewards = np.array([0.4, -0.3, 1.2, 0.05], np.float32) if scst else None
    ref = _reference(c, W, im, fm, caps, masks, keeps, rewards)
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    eng.set_precision('f32')
    dmasks = None
    if masks is not None:
        dmasks = dict(init_in=eng.to_dev(masks['init_in']), inp=eng.to_dev(masks['inp']), out=eng.to_dev(masks['out']),
                      att=eng.to_dev(masks['att'].reshape(masks['att'].shape[0], masks['att'].shape[1], -1)))
    out = tr.forward_backward(eng.to_dev(fm), eng.to_dev(im), caps, rewards, dmasks, keeps, want_logits=True,
                              want_attn=True)
    loss = out['loss'].cpu().numpy()
    assert abs(loss[1] - ref['xe']) < 1e-4 * max(1.0, abs(ref['xe']))
    assert abs(loss[2] - ref['map']) < 1e-4 * max(1e-3, abs(ref['map']))
    assert abs(loss[3] - ref['reg']) < 1e-4 * max(1e-3, abs(ref['reg']))
    assert abs(loss[0] - ref['total']) < 1e-4 * max(1.0, abs(ref['total']))
    T_run = out['T_run']
    lg = out['logits'].cpu().numpy().transpose(1, 0, 2)                 # [T,B,V]
    assert rel_err(lg, ref['logits']) < 2e-4
    assert rel_err(out['attn'].cpu().numpy(), ref['attn']) < 2e-4
    worst = {}
    for vname in wts.decoder_shapes(c):
        g = tr.gradient(vname).cpu().numpy().reshape(ref['grads'][vname].shape)
        worst[vname] = rel_err(g, ref['grads'][vname])
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, (bad, worst)


@pytest.mark.parametrize('name', ['comic256_xe_dropout', 'project_hidden', 'dot_sigmoid_ctx_ph_indep_h4',
                                  'word_none_ctx_sigmoid'])
def test_encoder_output_gradients_match_autograd(torch_mod, name):
    """comic_train_encoder_grads (what cnn_finetune feeds the CNN backward): d loss / d fm and d loss / d im_embed."""
    import torch_ref as TR
    from comic_b200.train import Trainer
    mk, dropout, scst = CASES[name]
    c = mk()
    W, im, fm, caps, masks, keeps = _train_case(c, B=4, L=8, seed=3, dropout=dropout)
    rewards = np.array([0.4, -0.3, 1.2, 0.05], np.float32) if scst else None
    P = TR.to_params(W)
    tot, _, _, _, aux = TR.training_loss(P, c, im, fm, caps, masks, keeps, rewards, fm_requires_grad=True)
    tot.backward()
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    eng.set_precision('f32')
    dmasks = dict(init_in=eng.to_dev(masks['init_in']), inp=eng.to_dev(masks['inp']), out=eng.to_dev(masks['out']),
                  att=eng.to_dev(masks['att'].reshape(masks['att'].shape[0], masks['att'].shape[1], -1)))
    out = tr.forward_backward(eng.to_dev(fm), eng.to_dev(im), caps, rewards, dmasks, keeps)
    dfm, demb = eng.train_encoder_grads(4, out['T_run'])
    assert rel_err(dfm.cpu().numpy(), aux['fm'].grad.numpy()) < 1e-3
    assert rel_err(demb.cpu().numpy(), aux['im'].grad.numpy()) < 1e-3


@pytest.mark.parametrize('name', ['comic256_xe_dropout', 'dot', 'context_layer', 'dot_sigmoid_ctx_ph_indep_h4'])
def test_training_forward_through_the_fused_attention_kernel(torch_mod, name):
    """From 48 images on the training forward takes attn_fused_kernel (one launch for scores, softmax, dropout and context)
    instead of the sliced kernels the small parity cases run: forced here on the 4-row case, it must leave the same
    tape, i.e. the same losses and gradients up to summation order (signorm keeps the sliced kernels by design)."""
    from comic_b200.train import Trainer
    mk, dropout, scst = CASES[name]
    c = mk()
    W, im, fm, caps, masks, keeps = _train_case(c, B=4, L=8, seed=3, dropout=dropout)
    rewards = np.array([0.4, -0.3, 1.2, 0.05], np.float32) if scst else None
    res = []
    for min_images in (1, 10000):
        tr = Trainer(c, W, with_cnn=False)
        eng = tr.engine
        eng.set_precision('f32')
        eng.set_option('fused_attn_min_images', min_images)
        dmasks = dict(init_in=eng.to_dev(masks['init_in']), inp=eng.to_dev(masks['inp']), out=eng.to_dev(masks['out']),
                      att=eng.to_dev(masks['att'].reshape(masks['att'].shape[0], masks['att'].shape[1], -1)))
        n0 = eng.launch_count()
        out = tr.forward_backward(eng.to_dev(fm), eng.to_dev(im), caps, rewards, dmasks, keeps, want_logits=True)
        res.append((out['loss'].cpu().numpy(), tr.grads.cpu().numpy().copy(), eng.launch_count() - n0))
    (l1, g1, n1), (l2, g2, n2) = res
    if c.attn_probability_fn == 'softmax':
        assert n1 < n2, 'the fused kernel replaces two launches per step'
    np.testing.assert_allclose(l1, l2, rtol=2e-5, atol=1e-6)
    assert rel_err(g1, g2) < 2e-5


def test_training_forward_on_the_k_split_tensor_path(torch_mod):
    """48 rows in the default precision: the teacher-forced forward runs its gate / [logits | query] GEMMs on the tensor path
    with the K loop split over CTAs (partials summed by the LSTM kernel that also writes the gate tape).  Same losses and
    gradients as the all-FFMA f32 mode, within the bf16x3 budget."""
    from comic_b200.train import Trainer
    c = comic_config(train_mode='decoder')
    W, im, fm, caps, masks, keeps = _train_case(c, B=48, L=7, seed=9, dropout=True)
    res = []
    for prec in ('split', 'f32'):
        tr = Trainer(c, W, with_cnn=False)
        eng = tr.engine
        eng.set_precision(prec)
        dmasks = dict(init_in=eng.to_dev(masks['init_in']), inp=eng.to_dev(masks['inp']), out=eng.to_dev(masks['out']),
                      att=eng.to_dev(masks['att'].reshape(masks['att'].shape[0], masks['att'].shape[1], -1)))
        out = tr.forward_backward(eng.to_dev(fm), eng.to_dev(im), caps, None, dmasks, keeps, want_logits=True)
        res.append((out['loss'].cpu().numpy(), out['logits'].cpu().numpy(), tr.grads.cpu().numpy().copy()))
    (l1, lg1, g1), (l2, lg2, g2) = res
    np.testing.assert_allclose(l1, l2, rtol=1e-4, atol=1e-6)
    assert rel_err(lg1, lg2) < 2e-4
    assert rel_err(g1, g2) < 1e-3


def test_legacy_head_trains_in_decoder_mode(torch_mod):
    """--legacy, train_mode=decoder (the only mode the reference trains legacy models in, src/train.py:242, 253): the
    image-embedding head LN_tanh + im_embed is trainable (src/model_base.py:80-91).  Every decoder gradient and the three
    head gradients against fp64 autograd of one graph Mixed_5c -> head -> decoder -> loss."""
    import torch
    import torch_ref as TR
    from comic_b200.train import Trainer
    from comic_b200 import weights as wts
    c = comic_config(train_mode='decoder', legacy=True)
    W, _im, fm, caps, masks, keeps = _train_case(c, B=4, L=8, seed=3, dropout=True)
    rng = np.random.default_rng(12)
    m5c = np.maximum(rng.standard_normal((4, 7, 7, 1024)), 0).astype(np.float32)
    P = TR.to_params(W)
    PH = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in W.items()
          if k.startswith(wts.ENC) and not k.startswith(wts.CNN)}
    assert len(PH) == 3
    im_t = TR.legacy_head(PH, m5c)
    tot, xe, mp, reg, aux = TR.training_loss(P, c, im_t, fm, caps, masks, keeps)
    for v in PH.values():                                  # the head's variables are in tvars: L2 applies (model_base.py:368-380)
        tot = tot + (v ** 2).sum() / 2 * c.l2_decay
        reg = reg + (v ** 2).sum() / 2 * c.l2_decay
    tot.backward()
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    eng.set_precision('f32')
    dmasks = dict(init_in=eng.to_dev(masks['init_in']), inp=eng.to_dev(masks['inp']), out=eng.to_dev(masks['out']),
                  att=eng.to_dev(masks['att'].reshape(masks['att'].shape[0], masks['att'].shape[1], -1)))
    out = tr.forward_backward(eng.to_dev(fm), eng.to_dev(im_t.detach().numpy().astype(np.float32)), caps, None, dmasks, keeps,
                              mixed5c=eng.to_dev(m5c))
    loss = out['loss'].cpu().numpy()
    assert abs(loss[1] - float(xe)) < 1e-4 * max(1.0, abs(float(xe)))
    assert abs(loss[3] - float(reg)) < 1e-4 * max(1e-3, abs(float(reg)))
    worst = {}
    for vname in wts.decoder_shapes(c):
        worst[vname] = rel_err(tr.gradient(vname).cpu().numpy().reshape(P[vname].shape), P[vname].grad.numpy())
    for vname, v in PH.items():
        worst[vname] = rel_err(tr.gradient(vname).cpu().numpy().reshape(v.shape), v.grad.numpy())
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, (bad, worst)
    with pytest.raises(NotImplementedError):               # src/train.py:242, 253
        Trainer(comic_config(train_mode='scst', legacy=True), W, with_cnn=False)


def test_legacy_model_train_step_end_to_end(torch_mod):
    """Trainer.step on a --legacy model: images -> frozen CNN (Mixed_5c kept) -> head -> decoder fwd + bwd -> Adam; the
    head's variables move, and the next encode uses them."""
    from _common import images
    from comic_b200.train import Trainer
    from comic_b200 import weights as wts
    c = comic_config(train_mode='decoder', legacy=True, max_step=100)
    W = make_weights(c, seed=5)
    _, _, _, caps, _, _ = _train_case(c, B=2, L=7, seed=4, dropout=False)
    tr = Trainer(c, W)
    img = tr.engine.to_dev(images(2, seed=3))
    emb0, _ = tr.engine.encode(img)
    w0 = tr.variable(wts.ENC + 'im_embed/weight').clone()
    losses = [float(tr.step(img, caps, lr=1e-2)['loss'][1]) for _ in range(3)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert float((tr.variable(wts.ENC + 'im_embed/weight') - w0).abs().max()) > 0
    emb1, _ = tr.engine.encode(img)
    assert float((emb1 - emb0).abs().max()) > 0


def test_adam_and_l2_match_oracle(torch_mod):
    import comic_oracle as O
    from comic_b200.engine import Engine
    torch = torch_mod
    c = comic_config()
    eng = Engine(c)
    rng = np.random.default_rng(0)
    n = 10007
    th = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    d_th, d_m, d_v = eng.to_dev(th), eng.to_dev(m), eng.to_dev(v)
    th64, m64, v64 = th.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for step in range(1, 4):
        g = rng.standard_normal(n).astype(np.float32) * 0.1
        eng.adam_step(d_th, eng.to_dev(g), d_m, d_v, 1e-2, step, eps=1e-2, grad_scale=0.5)
        th64, m64, v64 = O.adam_step(th64, g.astype(np.float64) * 0.5, m64, v64, 1e-2, step, eps=1e-2)
    assert rel_err(d_th.cpu().numpy(), th64) < 1e-6
    assert rel_err(d_m.cpu().numpy(), m64) < 1e-5
    reg = torch.zeros(1, device=eng.device)
    gbuf = eng.to_dev(np.zeros(n, np.float32))
    eng.l2_regularise(d_th, gbuf, 1e-5, reg)
    assert abs(float(reg.item()) - 0.5e-5 * float((th64 ** 2).sum())) < 1e-6 * float((th64 ** 2).sum())
    assert rel_err(gbuf.cpu().numpy(), 1e-5 * th64) < 1e-5


def test_dropout_masks_are_seeded_bernoulli(torch_mod):
    from comic_b200.engine import Engine
    eng = Engine(comic_config())
    a = eng.dropout_masks((41, 32, 768), 0.65, seed=7, stream_id=1).cpu().numpy()
    b = eng.dropout_masks((41, 32, 768), 0.65, seed=7, stream_id=1).cpu().numpy()
    c2 = eng.dropout_masks((41, 32, 768), 0.65, seed=8, stream_id=1).cpu().numpy()
    assert set(np.unique(a)) <= {0.0, 1.0}
    np.testing.assert_array_equal(a, b)
    assert (a != c2).mean() > 0.3
    assert abs(a.mean() - 0.65) < 5e-3
    assert abs(np.corrcoef(a.reshape(-1)[:-1], a.reshape(-1)[1:])[0, 1]) < 0.01


def test_recurrent_dropout_masks_are_shared_over_rows_and_steps(torch_mod):
    """rnn_recurr_dropout=True: DropoutWrapper(variational_recurrent=True) draws one input and one output mask per
    optimiser step (leading dimension 1 in TF r1.9), used by every row, every time step and the rnn-init call."""
    from comic_b200.train import Trainer
    c = comic_config(train_mode='decoder', rnn_recurr_dropout=True)
    W, im, fm, caps, _, _ = _train_case(c, B=3, L=7, seed=2, dropout=False)
    from comic_b200.train import process_inputs
    tr = Trainer(c, W, with_cnn=False)
    T_run = int(process_inputs(caps, c.token_type)[3].max())         # executed steps of this batch (tokens + <EOS>)
    masks, keeps = tr.make_masks(3, T_run, seed=11)
    inp, out, init = masks['inp'].cpu().numpy(), masks['out'].cpu().numpy(), masks['init_in'].cpu().numpy()
    assert inp.shape == (T_run, 3, 768) and out.shape == (T_run, 3, 512) and init.shape == (3, 768)
    assert (inp == inp[0, 0]).all() and (out == out[0, 0]).all() and (init == inp[0, 0]).all()
    assert 0.5 < inp[0, 0].mean() < 0.8 and set(np.unique(inp)) <= {0.0, 1.0}
    m2, _ = tr.make_masks(3, T_run, seed=12)
    assert (m2['inp'].cpu().numpy()[0, 0] != inp[0, 0]).any()
    out1 = tr.forward_backward(tr.engine.to_dev(fm), tr.engine.to_dev(im), caps, None, masks, keeps)
    assert np.isfinite(float(out1['loss'][0]))
    short, _ = tr.make_masks(3, T_run - 1, seed=11)
    with pytest.raises(ValueError):                       # masks for fewer steps than the batch executes are refused
        tr.forward_backward(tr.engine.to_dev(fm), tr.engine.to_dev(im), caps, None, short, keeps)


def test_momentum_sgd_and_per_variable_clipping(torch_mod):
    """--optimiser sgd (tf.train.MomentumOptimizer(lr, 0.9), src/model_base.py:868-880) and clip_gradient_norm
    (slim clip_gradient_norms: each variable by its own norm) against NumPy on the Trainer's own gradients."""
    from comic_b200.train import Trainer
    c = comic_config(train_mode='decoder', optimiser='sgd', clip_gradient_norm=0.05, max_step=100)
    W, im, fm, caps, _, _ = _train_case(c, B=4, L=8, seed=6, dropout=False)
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    fm_d, im_d = eng.to_dev(fm), eng.to_dev(im)
    theta = tr.params.cpu().numpy().astype(np.float64)
    acc = np.zeros_like(theta)
    clipped_any = False
    for it in range(2):
        tr.forward_backward(fm_d, im_d, caps)
        g = tr.grads.cpu().numpy().astype(np.float64)
        for name, (o, n, _) in tr.offsets.items():
            nrm = np.sqrt((g[o:o + n] ** 2).sum())
            if nrm > 0.05:
                g[o:o + n] *= 0.05 / nrm
                clipped_any = True
        acc = 0.9 * acc + g
        theta = theta - 1e-2 * acc
        tr.apply_gradients(lr=1e-2)
        assert rel_err(tr.params.cpu().numpy(), theta) < 1e-5
        assert rel_err(tr.adam_m.cpu().numpy(), acc) < 1e-5
        theta = tr.params.cpu().numpy().astype(np.float64)
        acc = tr.adam_m.cpu().numpy().astype(np.float64)
    assert clipped_any


def test_training_steps_reduce_loss(torch_mod):
    """A few optimiser steps on one fixed batch (teacher forcing, dropout off): XE loss goes down,
    and the packed decoder copies follow the updated variables (greedy decode changes)."""
    from comic_b200.train import Trainer
    c = comic_config(train_mode='decoder', max_step=100)
    W, im, fm, caps, _, _ = _train_case(c, B=6, L=9, seed=5, dropout=False)
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    fm_d, im_d = eng.to_dev(fm), eng.to_dev(im)
    losses = []
    for _ in range(6):
        out = tr.forward_backward(fm_d, im_d, caps)
        losses.append(float(out['loss'][1].item()))
        tr.apply_gradients(lr=5e-3)
    assert losses[-1] < losses[0] - 0.05, losses
    assert all(np.isfinite(losses))


def _synthetic_refs(B, n_words=1000, seed=0):
    """5 random 10-word reference sentences per image over a 1,000-word vocabulary."""
    rng = np.random.default_rng(seed)
    return [[' '.join('w%d' % w for w in rng.integers(0, n_words - 3, size=10)) for _ in range(5)] for _ in range(B)]


def test_scst_step_end_to_end(torch_mod):
    """train_fn.py:218-256 on the GPU path: greedy + beam-k sampling equals the oracle's decodes,
    rewards = weighted CIDEr-D + BLEU-4 of the decoded strings, weighted-XE step runs."""
    import comic_oracle as O
    import inception_v1_oracle as I
    from comic_b200 import scst as S
    from comic_b200.train import Trainer
    from _common import images
    c = comic_config(train_mode='scst', scst_beam_size=3, n_words=1000, max_step=50)
    W = make_weights(c)
    B = 2
    img = images(B, seed=12)
    refs = _synthetic_refs(B)
    df = {'document_frequency': S.compute_doc_freq(refs), 'ref_len': B}
    scorer = S.CaptionScorer(df, dict(ciderD=c.scst_weight_ciderD, bleu=c.scst_weight_bleu))
    tr = Trainer(c, W)
    eng = tr.engine
    before = tr.params.clone()
    cap_beam, cap_greedy, im_embed, fm = S.sample_captions(eng, c, eng.to_dev(img), 3, max_length=4)
    o_emb, o_fm, _ = I.encoder(img, W, c)
    rb = O.beam_search_decode(O.Decoder(W, c), o_emb, o_fm, 3, 0.0, 8)
    rg = O.greedy_decode(O.Decoder(W, c), o_emb, o_fm, 8)
    np.testing.assert_array_equal(cap_beam.cpu().numpy(), rb['predicted_ids'].transpose(2, 1, 0))
    np.testing.assert_array_equal(cap_greedy.cpu().numpy(), rg['ids'].T)
    out = S.scst_step(tr, scorer, eng.to_dev(img), refs, seed=3, lr=1e-3)
    assert out['rewards'].shape == (B * 3,) and np.isfinite(out['rewards']).all()
    np.testing.assert_allclose(out['rewards'], out['sc_sample'] - out['sc_greedy'], rtol=1e-6)
    loss = out['loss'].cpu().numpy()
    assert np.isfinite(loss).all()
    assert tr.global_step == 1 and not torch_mod.equal(before, tr.params)


def test_caption_model_train_and_eval_surface(torch_mod):
    from comic_b200.model import CaptionModel
    from _common import images
    c = comic_config(train_mode='decoder', max_step=20)
    W = make_weights(c)
    m_train = CaptionModel(c, 'train', weights=W)
    m_eval = CaptionModel(c, 'eval', reuse=True, share=m_train)
    m_infer = CaptionModel(c, 'infer', reuse=True, share=m_train)
    _, _, _, caps, _, _ = _train_case(c, B=3, L=7, seed=1, dropout=False)
    img = images(3, seed=2)
    p0 = float(m_eval.eval_step(img, caps).item())
    for _ in range(3):
        ppl, gs = m_train.train_step(img, caps, seed=11, lr=3e-3)
    assert gs == 3
    p1 = float(m_eval.eval_step(img, caps).item())
    assert p1 < p0
    preds, attn = m_infer.run(img)                       # shares the updated variables
    assert preds.shape[0] == 3 and attn.shape[:2] == (3, 8)


def test_default_train_step_applies_dropout_and_eval_is_forward_only(torch_mod):
    """`train_step(images, captions)` regularises like the reference's train graph (DropoutWrapper +
    attention-map dropout on by default); `eval_step` runs forward only -- it works under train_mode=cnn_finetune
    and leaves the trainer's gradient buffer alone."""
    from comic_b200.model import CaptionModel
    from _common import images
    c = comic_config(train_mode='cnn_finetune', max_step=20)
    W = make_weights(c)
    _, _, _, caps, _, _ = _train_case(c, B=2, L=6, seed=3, dropout=False)
    img = images(2, seed=4)
    m = CaptionModel(c, 'train', weights=W)
    m_eval = CaptionModel(c, 'eval', reuse=True, share=m)
    tr = m.trainer
    # same weights, same data: the default call draws masks, dropout=False does not
    eng = tr.engine
    im_embed, fm = eng.encode(eng.to_dev(img))
    plain = float(tr.forward_backward(fm, im_embed, np.asarray(caps), forward_only=True)['loss'][1].item())
    ev = float(m_eval.eval_step(img, caps).item())
    assert abs(ev - plain) < 1e-6 * max(1.0, abs(plain))
    tr.grads.fill_(3.0)
    m_eval.eval_step(img, caps)
    assert float(tr.grads.min().item()) == 3.0 and float(tr.grads.max().item()) == 3.0
    ppl_default, _ = m.train_step(img, caps, lr=0.0)
    ppl_nodrop, _ = m.train_step(img, caps, lr=0.0, dropout=False)
    assert abs(float(ppl_nodrop.item()) - plain) < 2e-4 * max(1.0, abs(plain))
    assert abs(float(ppl_default.item()) - plain) > 1e-3 * max(1.0, abs(plain))


def test_restore_model_from_a_v2_checkpoint(torch_mod, tmp_path):
    """CaptionModel.restore_model with a TF V2 checkpoint path: variables, Adam slots and global_step come back
    (src/model_base.py:422-490, resume case) and the restored model decodes like the one that wrote it."""
    from comic_b200.model import CaptionModel
    from comic_b200 import checkpoint as ck
    from _common import images
    c = comic_config(train_mode='decoder', max_step=20, infer_max_length=2)
    W = make_weights(c)
    m = CaptionModel(c, 'train', weights=W)
    _, _, _, caps, _, _ = _train_case(c, B=2, L=6, seed=3, dropout=False)
    img = images(2, seed=4)
    m.train_step(img, caps, seed=1, lr=1e-3)
    tr = m.trainer
    tensors = tr.variables_numpy()
    for name, (o, n, shp) in tr.offsets.items():
        tensors[name + '/Adam'] = tr.adam_m[o:o + n].cpu().numpy().reshape(shp)
        tensors[name + '/Adam_1'] = tr.adam_v[o:o + n].cpu().numpy().reshape(shp)
    tensors['global_step'] = np.array(tr.global_step, np.int64)
    prefix = str(tmp_path / 'model_compact-1')
    ck.write_v2(prefix, tensors, with_data_crc=False)
    c2 = comic_config(train_mode='decoder', max_step=20, infer_max_length=2, checkpoint_path=str(tmp_path), resume_training=True)
    m2 = CaptionModel(c2, 'train', weights=make_weights(c2, seed=99))
    info = m2.restore_model()
    assert info['mode'] == 'resume' and m2.trainer.global_step == 1
    assert torch_mod.equal(m2.trainer.params, tr.params) and torch_mod.equal(m2.trainer.adam_m, tr.adam_m)
    a = CaptionModel(c, 'infer', reuse=True, share=m).run(img)
    b = CaptionModel(c2, 'infer', reuse=True, share=m2).run(img)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_char_tokens_through_input_manager_train_eval_infer(torch_mod, tmp_path):
    """token_type=char end to end -- InputManager_Char batches -> train steps -> the validation
    perplexity loop of src/train_fn.py:320-338 -> beam-search inference -> id_to_caption."""
    from comic_b200 import inputs, scst
    from comic_b200.model import CaptionModel
    from test_inputs import _dataset, _config
    from _common import images
    _dataset(tmp_path, n_train=32, n_valid=32)
    c = _config(tmp_path, 'char', train_mode='decoder', infer_max_length=4)
    loader = lambda paths: images(len(paths), seed=len(paths[0]))
    man = inputs.get_input_manager(c, image_loader=loader)
    assert c.vocab_size == 40 and c.max_step == int(32 / 4 * 3)
    W = make_weights(c)
    assert W['Model/decoder/rnn_decoder/embedding_map'].shape[0] == 40
    m = CaptionModel(c, 'train', weights=W)
    m_eval = CaptionModel(c, 'eval', reuse=True, share=m)
    valid = list(man.batches('valid', epochs=1))
    assert len(valid) >= 1
    p0 = inputs.run_eval_loop(m_eval, valid)
    n = 0
    for img, caps in man.batches('train', epochs=2):
        assert caps.min() >= -1 and caps.max() == 39
        m.train_step(img, caps, seed=5, lr=3e-3, dropout=False)
        n += 1
    assert n >= 4
    p1 = inputs.run_eval_loop(m_eval, valid)
    assert np.isfinite(p0) and np.isfinite(p1) and p1 < p0
    m_inf = CaptionModel(c, 'infer', reuse=True, share=m)
    preds, _ = m_inf.run(images(2, seed=3))
    assert preds.shape == (2, 4 * 5) and preds.max() <= 39            # char: infer_max_length x 5 steps
    caps = scst.id_to_caption(preds, c)
    assert len(caps) == 2 and all(set(s.replace('<GO>', '')) <= set(' 0123456789abcdefghijklmnopqrstuvwxyz') for s in caps)


def test_cuda_graph_replay_of_fwd_bwd_is_bit_identical(torch_mod):
    """Trainer replays the teacher-forced forward + backward as one CUDA graph from the third call of a shape on: loss and
    the flat gradient buffer must equal the eager launches bit for bit (fixed-order reductions), with new inputs copied
    into the graph's static buffers on every replay, dropout masks included."""
    from comic_b200.train import Trainer
    c = comic_config(train_mode='decoder', max_step=100)
    W, im, fm, caps, _, _ = _train_case(c, B=5, L=8, seed=7, dropout=False)
    rng = np.random.default_rng(3)
    runs = {}
    for graphed in (False, True):
        tr = Trainer(c, W, with_cnn=False)
        tr.cuda_graph = graphed
        eng = tr.engine
        got = []
        for i in range(5):
            scale = 1.0 + 0.1 * i                       # different inputs every call
            fm_d, im_d = eng.to_dev(fm * scale), eng.to_dev(im * scale)
            masks, keeps = tr.make_masks(5, int((caps[:, 1:] >= 0).sum(1).max()), 100 + i)
            out = tr.forward_backward(fm_d, im_d, caps, None, masks, keeps)
            got.append((out['loss'].clone(), tr.grads.clone()))
        if graphed:
            assert any(e['graph'] is not None for e in tr._graphs.values())
        runs[graphed] = got
    for (l0, g0), (l1, g1) in zip(runs[False], runs[True]):
        assert torch_mod.equal(l0, l1) and torch_mod.equal(g0, g1)
    assert not torch_mod.equal(runs[True][3][1], runs[True][4][1])
