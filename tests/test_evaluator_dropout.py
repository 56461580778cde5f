"""Evaluator and dropout behaviours the reference pins down in its unit tests
(/root/reference/tests/unit/test_evaluator.py:50-124, test_dropout.py:58-131): a short last
minibatch (batch_size < allocated rows) gets zero error rows however the buffer was poisoned, error
counts / confusion matrix / max error sums, and the inverted-dropout statistics."""
import numpy

from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.normalization import NoneNormalizer
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops import dropout
from veles.znicz_b200.workflow import evaluator

RS = numpy.random.RandomState(21)


def test_softmax_evaluator_short_batch_and_metrics():
    batch, classes, valid = 25, 7, 20
    logits = RS.uniform(-1, 1, (batch, classes))
    e = numpy.exp(logits - logits.max(axis=1, keepdims=True))
    probs = (e / e.sum(axis=1, keepdims=True)).astype(numpy.float32)
    labels = RS.randint(0, classes, batch).astype(numpy.int32)
    labels[3] = -1                                        # unlabeled sample: ignored
    ev = evaluator.EvaluatorSoftmax(DummyWorkflow(), compute_confusion_matrix=True)
    ev.output = Array(probs.copy())
    ev.max_idx = Array(probs.argmax(axis=1).astype(numpy.int32))
    ev.labels = Array(labels.copy())
    ev.batch_size = valid
    ev.initialize(device=None)
    ev.err_output.map_write()
    ev.err_output.mem[...] = 1.0e30                       # poisoned: every row must be rewritten
    ev.run()
    onehot = numpy.zeros_like(probs)
    ok = labels >= 0
    onehot[numpy.arange(batch)[ok], labels[ok]] = 1
    gold = (probs - onehot) / valid
    gold[valid:] = 0
    gold[3] = 0
    assert numpy.abs(ev.err_output.mem - gold).max() < 1e-6
    rows = [i for i in range(valid) if labels[i] >= 0]
    wrong = sum(int(probs[i].argmax() != labels[i]) for i in rows)
    assert int(ev.n_err.mem[0]) == wrong and int(ev.n_err.mem[1]) == len(rows)
    cm = ev.confusion_matrix.mem
    assert int(cm.sum()) == len(rows)
    i = rows[0]
    assert cm[probs[i].argmax(), labels[i]] >= 1
    assert abs(float(ev.max_err_output_sum.mem[0]) -
               max(numpy.abs(gold[i]).sum() for i in rows)) < 1e-6


def test_mse_evaluator_short_batch_and_rmse():
    batch, size, valid = 12, 30, 9
    out = RS.uniform(-1, 1, (batch, size)).astype(numpy.float32)
    tgt = RS.uniform(-1, 1, (batch, size)).astype(numpy.float32)
    ev = evaluator.EvaluatorMSE(DummyWorkflow())
    ev.output, ev.target = Array(out.copy()), Array(tgt.copy())
    ev.batch_size = valid
    ev.normalizer = NoneNormalizer()
    ev.normalizer.analyze(None)
    ev.initialize(device=None)
    ev.err_output.map_write()
    ev.err_output.mem[...] = 1.0e30
    ev.run()
    gold = (out - tgt) / valid
    gold[valid:] = 0
    assert numpy.abs(ev.err_output.mem - gold).max() < 1e-6
    per_sample = numpy.sqrt(((out - tgt)[:valid] ** 2).mean(axis=1))    # root = True: RMSE
    assert numpy.allclose(ev.mse.mem[:valid], per_sample, atol=1e-5)
    assert abs(float(ev.metrics.mem[0]) - per_sample.sum()) < 1e-4
    assert abs(float(ev.metrics.mem[1]) - per_sample.max()) < 1e-5
    assert abs(float(ev.metrics.mem[2]) - per_sample.min()) < 1e-5


def test_dropout_statistics_and_eval_mode():
    wf = DummyWorkflow()
    d = dropout.DropoutForward(wf, dropout_ratio=0.4)
    x = numpy.ones((200, 500), numpy.float32)
    d.input = Array(x.copy())
    d.minibatch_class = 2                                  # TRAIN
    d.initialize(device=None)
    d.run()
    y = d.output.mem
    kept = y != 0
    assert abs(kept.mean() - 0.6) < 0.01                   # keep probability 1 - p
    assert numpy.allclose(y[kept], 1.0 / 0.6, atol=1e-5)   # inverted dropout: E[y] = x
    assert abs(y.mean() - 1.0) < 0.02
    b = dropout.DropoutBackward(wf, dropout_ratio=0.4)
    b.mask = d.mask
    b.err_output = Array(numpy.full_like(x, 2.0))
    b.minibatch_class = 2
    b.initialize(device=None)
    b.run()
    assert numpy.allclose(b.err_input.mem, 2.0 * d.mask.mem, atol=1e-5)
    d.minibatch_class = 1                                  # VALID: identity
    d.run()
    assert numpy.array_equal(d.output.mem, x)
    d.minibatch_class = 2
    d.run()
    assert not numpy.array_equal(d.output.mem != 0, kept)  # a new mask every run
