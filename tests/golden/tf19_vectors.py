"""Known-answer vectors of TensorFlow r1.9's own unit tests for the pieces of `tf.contrib.seq2seq` this path restates.

Provenance: tensorflow/contrib/seq2seq/python/kernel_tests/beam_search_decoder_test.py (class TestBeamStep:
`test_step`, `test_step_with_eos`; batch 2, beam 3, vocab 5, end_token 0, length_penalty_weight 0.6) and
beam_search_ops_test.py (GatherTreeTest.testGatherTreeOne; end_token 10) at branch r1.9.  TensorFlow is not vendored in
the reference and there is no network here, so inputs AND expected outputs below were restated from the public test
source rather than copied from a file on this machine.  What makes the restatement trustworthy: inputs and expected
arrays were written down independently, and the NumPy oracle -- itself written from TF's semantics, not from these
numbers -- reproduces every expected array exactly (tests/test_oracle_known_answers.py); a misremembered input or output
would not survive that.  These are the only reference-side known answers that exist for the beam-search step.
"""
import numpy as np

BATCH, BEAM, VOCAB, END_TOKEN, LENGTH_PENALTY = 2, 3, 5, 0, 0.6


def _logits(b1_beam1_tok2, b1_beam2_tok2):
    x = np.full((BATCH, BEAM, VOCAB), 0.0001, np.float32)
    x[0, 0, 2] = 1.9
    x[0, 0, 3] = 2.1
    x[0, 1, 3] = 3.1
    x[0, 1, 4] = 0.9
    x[1, 0, 1] = 0.5
    x[1, 1, 2] = b1_beam1_tok2
    x[1, 2, 2] = b1_beam2_tok2
    x[1, 2, 3] = 0.2
    return x


# TestBeamStep.test_step: nothing finished, every beam at length 2, log_probs = log_softmax(ones)
STEP = dict(
    logits=_logits(2.7, 10.0),
    finished=np.zeros((BATCH, BEAM), bool),
    lengths=np.full((BATCH, BEAM), 2, np.int64),
    predicted_ids=np.array([[3, 3, 2], [2, 2, 1]], np.int32),
    parent_ids=np.array([[1, 0, 0], [2, 1, 0]], np.int32),
    next_lengths=np.array([[3, 3, 3], [3, 3, 3]], np.int64),
    next_finished=np.zeros((BATCH, BEAM), bool),
    # expected_log_probs[b, i] = initial[b, parent] + log_softmax(logits)[b, parent, id]
)

# TestBeamStep.test_step_with_eos: beam (0, 1) and beam (1, 2) already finished
STEP_WITH_EOS = dict(
    logits=_logits(5.7, 1.0),
    finished=np.array([[False, True, False], [False, False, True]]),
    lengths=np.array([[2, 1, 2], [2, 2, 1]], np.int64),
    predicted_ids=np.array([[0, 3, 2], [2, 0, 1]], np.int32),
    parent_ids=np.array([[1, 0, 0], [1, 2, 0]], np.int32),
    next_lengths=np.array([[1, 3, 3], [3, 1, 3]], np.int64),
    next_finished=np.array([[True, False, False], [False, True, False]]),
)

# GatherTreeTest.testGatherTreeOne ([batch, time, beam] in the test file; time-major here)
GATHER_TREE = dict(
    end_token=10,
    step_ids=np.transpose(np.array([[[1, 2, 3], [4, 5, 6], [7, 8, 9], [-1, -1, -1]]], np.int32), (1, 0, 2)),
    parent_ids=np.transpose(np.array([[[0, 0, 0], [0, 1, 1], [2, 1, 2], [-1, -1, -1]]], np.int32), (1, 0, 2)),
    max_sequence_lengths=np.array([3], np.int32),
    expected=np.transpose(np.array([[[2, 2, 2], [6, 5, 6], [7, 8, 9], [10, 10, 10]]], np.int32), (1, 0, 2)),
)


def initial_log_probs():
    """nn_ops.log_softmax(array_ops.ones([batch, beam]))"""
    x = np.ones((BATCH, BEAM), np.float32)
    return (x - np.log(np.exp(x).sum(axis=1, keepdims=True))).astype(np.float32)
