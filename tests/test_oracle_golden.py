"""Pins oracle/lattice_oracle.py against the reference.

(1) Known-answer vectors restated from the reference's own tests
    (file:line cited per test); (2) tests/golden/*.npz produced by running the
    unmodified reference (tests/golden/make_golden.py).  CPU only.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest

from conftest import GOLDEN_DIR, golden_files
from oracle import lattice_oracle as O

SR = [('Real', O.REAL), ('Log', O.LOG), ('MaxTropical', O.MAXTROPICAL)]


def _load(name):
  return np.load(os.path.join(GOLDEN_DIR, name))


def _split(table):
  return np.ascontiguousarray(table[..., 0]), np.ascontiguousarray(table[..., 1:])


def _align(g):
  k = int(g['k'])
  return (0, True) if k < 0 else (k, False)


# ---- known answers from the reference test-suite ---------------------------

def test_log_plus_known_answer():
  # tests/semirings_test.py:197-201
  npt.assert_allclose(O.sr_plus(O.LOG, np.float32(2), np.float32(3)), 3.31326169, rtol=1e-6)
  npt.assert_allclose(O.sr_sum(O.LOG, np.array([2, 3], np.float32), 0), 3.31326169, rtol=1e-6)


def test_maxtropical_tie_gradients():
  # tests/semirings_test.py:226-247
  a = np.array([[1., 2., 3.], [0., 2., 4.]], np.float32)
  ga, gb = O.maximum_grad(a[0], a[1], np.ones(3, np.float32))
  npt.assert_array_equal(np.stack([ga, gb]), [[1, 1, 0], [0, 0, 1]])
  npt.assert_array_equal(O.max_grad(a, 0, np.ones(3, np.float32)), [[1, 1, 0], [0, 0, 1]])


def test_empty_sums():
  # tests/semirings_test.py:57-63, :181-189
  for _, sr in SR:
    npt.assert_array_equal(O.sr_sum(sr, np.zeros([3, 0], np.float32), 1),
                           np.full([3], O.sr_zero(sr)))
    assert O.sr_sum(sr, np.zeros([3, 0], np.float32), 0).shape == (0,)


def test_full_ngram_known_answers():
  # tests/contexts_test.py:41-170
  c0 = O.FullNGram(3, 0)
  assert c0.shape() == (1, 3)
  npt.assert_array_equal(c0.next_state([0, 0, 0], [0, 1, 2]), [0, 0, 0])
  npt.assert_array_equal(c0.next_state([0, 1, 2], [0, 0, 0]), [0, 1, 2])
  npt.assert_array_equal(c0.forward_reduce(np.arange(6.).reshape(2, 1, 3), O.REAL), [[3], [12]])
  npt.assert_array_equal(c0.backward_broadcast(np.array([[1.], [2.]])), [[[1, 1, 1]], [[2, 2, 2]]])
  c1 = O.FullNGram(2, 1)
  assert c1.shape() == (3, 2)
  npt.assert_array_equal(c1.next_state([0, 1, 2], [1, 2, 1]), [1, 2, 1])
  npt.assert_array_equal(c1.forward_reduce(np.arange(6.).reshape(3, 2), O.REAL), [0, 6, 9])
  npt.assert_array_equal(c1.backward_broadcast(np.arange(3.)), [[1, 2]] * 3)
  c2 = O.FullNGram(3, 2)
  assert c2.shape() == (13, 3)
  npt.assert_array_equal(c2.next_state([0, 1, 3, 4, 12], [1, 2, 3, 1, 2]), [1, 5, 12, 4, 11])
  npt.assert_array_equal(c2.next_state([0, 1, 3, 4, 12], [0] * 5), [0, 1, 3, 4, 12])
  npt.assert_array_equal(
      c2.forward_reduce(np.arange(39.).reshape(1, 13, 3), O.REAL),
      [[0, 0, 1, 2] + [i * 4 + 54 for i in range(3, 12)]])
  npt.assert_array_equal(
      c2.backward_broadcast(np.arange(13.).reshape(1, 13)),
      [[[1, 2, 3]] + [[4, 5, 6], [7, 8, 9], [10, 11, 12]] * 4])
  npt.assert_array_equal(c2.walk_states([2, 3, 1]), [0, 2, 9, 10])
  npt.assert_array_equal(c2.walk_states([2, 0, 0, 3, 1]), [0, 2, 2, 2, 9, 10])


def test_frame_dependent_hand_expansion():
  # tests/alignments_test.py:49-67, :93-140, :171-185
  rng = np.random.RandomState(0)
  ctx = O.FullNGram(2, 1)
  alpha, blank, beta = rng.rand(3), rng.rand(3), rng.rand(3)
  lex = rng.rand(3, 2)
  z = rng.rand()
  nxt = O.frame_forward(alpha, blank, lex, ctx, O.REAL)
  npt.assert_allclose(nxt, [alpha[0] * blank[0],
                            alpha[1] * blank[1] + np.sum(alpha * lex[:, 0]),
                            alpha[2] * blank[2] + np.sum(alpha * lex[:, 1])])
  nb, bm, lm = O.frame_backward(np.log(alpha), np.log(blank), np.log(lex),
                                np.log(beta), np.log(np.array(z)), ctx)
  npt.assert_allclose(np.exp(nb), [
      blank[p] * beta[p] + lex[p, 0] * beta[1] + lex[p, 1] * beta[2] for p in range(3)], rtol=1e-6)
  npt.assert_allclose(bm, alpha * blank * beta / z, rtol=1e-6)
  npt.assert_allclose(lm, [[alpha[p] * lex[p, y] * beta[y + 1] / z for y in range(2)]
                           for p in range(3)], rtol=1e-6)
  a4, b4, l4 = rng.rand(4), rng.rand(4), rng.rand(4)
  npt.assert_allclose(O.frame_string_forward(a4, b4, l4, O.REAL), [
      a4[0] * b4[0], a4[1] * b4[1] + a4[0] * l4[0], a4[2] * b4[2] + a4[1] * l4[1],
      a4[3] * b4[3] + a4[2] * l4[2]])
  # tests/alignments_test.py:27-37
  npt.assert_array_equal(O.shift_down(np.array([[1., 2, 3], [4, 5, 6]]), O.LOG),
                         [[-np.inf, 1, 2], [-np.inf, 4, 5]])


def test_lattice_frame_dependent_known_answer():
  # tests/lattices_test.py:181-288
  b, t, c, v = 3, 2, 3, 2
  table = 1 + np.arange(b * t * c * (1 + v), dtype=np.float32).reshape(b, t, c, 1 + v)
  table *= np.array([[-1, 1], [1, -1], [1, 1]], np.float32)[:, :, None, None]
  blank, lex = _split(table)
  ctx = O.FullNGram(v, 1)
  nf = np.array([2, 1, 0])
  lse = lambda xs: np.log(np.sum(np.exp(np.array(xs, np.float64))))
  den = lse([-1 + 10, -1 + 11, -1 + 12, -2 + 13, -2 + 14, -2 + 15, -3 + 16, -3 + 17, -3 + 18])
  npt.assert_allclose(O.lattice_forward(blank, lex, nf, ctx, O.MAXTROPICAL)[0], [15, 21, 0])
  npt.assert_allclose(O.lattice_forward(blank, lex, nf, ctx, O.REAL)[0],
                      [(-1) * 33 + (-2) * 42 + (-3) * 51, 60, 1])
  npt.assert_allclose(O.lattice_forward(blank, lex, nf, ctx, O.LOG)[0],
                      [den, lse([19, 20, 21]), 0], rtol=1e-6)
  dist, gb, gl, labels = O.viterbi(blank, lex, nf, ctx)
  npt.assert_allclose(dist, [15, 21, 0])
  # True paths: utterance 0 takes labels [2, 2] (-3, +18); utterance 1 takes label 2
  # (+21) in its single frame.  The reference's shortest_path reports [[1,1],[0,0],[0,0]]
  # (tests/lattices_test.py:238-242) because of SURVEY D4 (y-1) and D5 (batch mix-up);
  # its path WEIGHTS [15, 21, 0] are the trustworthy part and are what we pin.
  npt.assert_array_equal(labels, [[2, 2], [2, 0], [0, 0]])
  labels_in = np.array([[1, 2, 0], [2, 1, 0], [1, 2, 0]])
  nl = np.array([1, 1, 0])
  for sr, exp in [(O.MAXTROPICAL, [11, 21, 0]), (O.REAL, [-11 - 26, 21, 1]),
                  (O.LOG, [lse([10, 11]), 21, 0])]:
    npt.assert_allclose(O.lattice_string_forward(blank, lex, nf, labels_in, nl, ctx, sr),
                        exp, rtol=1e-6)
    npt.assert_array_equal(
        O.lattice_string_forward(blank, lex, nf, labels_in, np.array([3, 2, 1]), ctx, sr),
        np.full([3], O.sr_zero(sr)))
  loss, _, _ = O.lattice_loss_and_grads(blank, lex, nf, labels_in, nl, ctx)
  npt.assert_allclose(loss, [den - lse([10, 11]), lse([19, 20, 21]) - 21, 0], rtol=1e-6, atol=1e-6)


# ---- golden fixtures generated from the reference -------------------------

@pytest.mark.parametrize('fname', golden_files('lattice_'))
def test_lattice_golden(fname):
  g = _load(fname)
  ctx = O.FullNGram(int(g['vocab']), int(g['context_size']))
  k, fd = _align(g)
  nf, labels, nl = g['num_frames'], g['labels'], g['num_labels']
  for name, sr in SR:
    table = g['Real_table'] if name == 'Real' else g['table']
    blank, lex = _split(table)
    dist, alphas = O.lattice_forward(blank, lex, nf, ctx, sr, k, fd)
    npt.assert_allclose(dist, g[f'{name}_dist'], rtol=2e-5, atol=1e-6, err_msg=name)
    npt.assert_allclose(alphas, g[f'{name}_alphas'], rtol=2e-5, atol=1e-5, err_msg=name)
    sd = O.lattice_string_forward(blank, lex, nf, labels, nl, ctx, sr, k, fd)
    npt.assert_allclose(sd, g[f'{name}_string'], rtol=2e-5, atol=1e-6, err_msg=name)
  blank, lex = _split(g['table'])
  # Log denominators: both reference oracles (patched autograd, alignment.backward loop).
  log_z, gb, gl = O.lattice_marginals(blank, lex, nf, ctx, k, fd)
  for key in ['Log_dist_grad', 'Log_marginals_fb']:
    npt.assert_allclose(gb, g[key][..., 0], rtol=2e-4, atol=2e-6, err_msg=key)
    npt.assert_allclose(gl, g[key][..., 1:], rtol=2e-4, atol=2e-6, err_msg=key)
  # free invariant: FrameDependent marginals sum to num_frames.
  if fd:
    npt.assert_allclose(gb.sum((1, 2)) + gl.sum((1, 2, 3)), np.clip(nf, 0, blank.shape[1]), rtol=1e-4)
  # numerator gradient (patched reference autograd)
  bw, lw, states, safe = O.gather_string_weights(blank, lex, labels, ctx)
  num, gbw, glw = O.string_marginals(bw, lw, nf, nl, k, fd)
  sb, sl = O.scatter_string_grads(gbw, glw, states, safe, blank.shape, lex.shape)
  npt.assert_allclose(sb, g['Log_string_grad'][..., 0], rtol=2e-4, atol=2e-6)
  npt.assert_allclose(sl, g['Log_string_grad'][..., 1:], rtol=2e-4, atol=2e-6)
  loss, _, _ = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, ctx, k, fd)
  npt.assert_allclose(loss, g['loss'], rtol=2e-5, atol=2e-5)
  # MaxTropical: one-hot Viterbi gradient (reference autograd as shipped).
  dist, vb, vl, _ = O.viterbi(blank, lex, nf, ctx, k, fd)
  npt.assert_allclose(dist, g['MaxTropical_dist'], rtol=1e-6)
  npt.assert_array_equal(vb, g['MaxTropical_dist_grad'][..., 0])
  npt.assert_array_equal(vl, g['MaxTropical_dist_grad'][..., 1:])
  # Real gradient (plain reference autograd).
  rb, rl = _split(g['Real_table'])
  dist, gb, gl = O.real_lattice_grads(rb, rl, nf, ctx, k, fd)
  npt.assert_allclose(gb, g['Real_dist_grad'][..., 0], rtol=2e-4, atol=1e-6)
  npt.assert_allclose(gl, g['Real_dist_grad'][..., 1:], rtol=2e-4, atol=1e-6)


@pytest.mark.parametrize('fname', golden_files('frameops_'))
def test_frame_ops_golden(fname):
  g = _load(fname)
  ctx = O.FullNGram(int(g['vocab']), int(g['context_size']))
  k, fd = _align(g)
  for name, sr in SR:
    npt.assert_allclose(
        O.frame_forward(g['alpha'], g['blank'], g['lexical'], ctx, sr, k, fd),
        g[f'{name}_forward'], rtol=1e-5, atol=1e-6)
    npt.assert_allclose(
        O.frame_string_forward(g['salpha'], g['sblank'], g['slex'], sr, k, fd),
        g[f'{name}_string_forward'], rtol=1e-5, atol=1e-6)
    npt.assert_allclose(ctx.forward_reduce(g['lexical'], sr),
                        g[f'{name}_forward_reduce'], rtol=1e-5, atol=1e-6)
  nb, bm, lm = O.frame_backward(g['alpha'], g['blank'], g['lexical'], g['beta'],
                                g['log_z'], ctx, k, fd)
  npt.assert_allclose(nb, g['backward_next_beta'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(bm, g['backward_blank_marginal'], rtol=1e-5, atol=1e-7)
  npt.assert_allclose(lm, g['backward_lexical_marginal'], rtol=1e-5, atol=1e-7)
  npt.assert_array_equal(ctx.backward_broadcast(g['beta']), g['backward_broadcast'])
  npt.assert_array_equal(ctx.next_state_table(), g['next_state_table'])


@pytest.mark.parametrize('fname', golden_files('joint_'))
def test_joint_golden(fname):
  g = _load(fname)
  blank, lexical = O.joint_weights(g['cache'], g['frame'], g['w_ctx'], g['w_frame'],
                                   g['w_blank'], g['b_blank'], g['w_vocab'], g['b_vocab'])
  npt.assert_allclose(blank, g['blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(lexical, g['lexical'], rtol=1e-5, atol=1e-6)
  s = g['state']
  npt.assert_allclose(blank[np.arange(len(s)), s], g['state_blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(lexical[np.arange(len(s)), s], g['state_lexical'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('fname', golden_files('jointlattice_'))
def test_joint_lattice_golden(fname):
  """The whole GNAT loss of the reference with its JointWeightFn inside the lattice
  (tests/golden/make_golden.py: joint_lattice_case), in the tensor-core shape envelope: loss
  value of the reference as shipped, parameter gradients from its patched Log autograd."""
  g = _load(fname)
  f64 = lambda k: g[k].astype(np.float64)
  v = int(g['vocab'])
  blank, lex = O.joint_weights(f64('cache'), f64('frames'), f64('w_ctx'), f64('w_frame'),
                               f64('w_blank')[0], float(g['b_blank'][0]), f64('w_vocab'),
                               f64('b_vocab'))
  k = int(g['k'])
  loss, gb, gl = O.lattice_loss_and_grads(blank, lex, g['num_frames'], g['labels'],
                                          g['num_labels'], O.FullNGram(v, int(g['context_size'])),
                                          max(k, 0), k < 0)
  npt.assert_allclose(loss, g['loss'], rtol=1e-5)
  npt.assert_allclose(loss, g['loss_patched'], rtol=1e-5)
  joint = np.tanh((f64('cache') @ f64('w_ctx').T)[None, None] +
                  (f64('frames') @ f64('w_frame').T)[:, :, None, :])          # [B,T,C,H]
  dpre = (gl @ f64('w_vocab') + gb[..., None] * f64('w_blank')[0]) * (1.0 - joint * joint)
  want = {
      'w_vocab': np.einsum('btcv,btch->vh', gl, joint), 'b_vocab': gl.sum(axis=(0, 1, 2)),
      'w_blank': np.einsum('btc,btch->h', gb, joint)[None], 'b_blank': gb.sum()[None],
      'w_ctx': dpre.sum(axis=(0, 1)).T @ f64('cache'),
      'w_frame': np.einsum('bth,btd->hd', dpre.sum(axis=2), f64('frames')),
      'cache': dpre.sum(axis=(0, 1)) @ f64('w_ctx'),
  }
  for name, w in want.items():
    ref = g['grad_' + name]
    scale = np.abs(ref).max()
    # (FrameLabelDependent: every path takes exactly one blank arc per frame, so the b_blank
    # gradient of den - num is exactly 0 and the reference holds round-off noise: hence the atol)
    assert np.abs(w - ref).max() <= 2e-5 * scale + 1e-5, (name, np.abs(w - ref).max() / scale)


# ---- the C restatement (oracle/lattice_oracle.c) agrees with the numpy oracle ----

@pytest.mark.parametrize('vnk', [(5, 1, -1), (3, 2, -1), (4, 0, -1), (3, 1, 2), (2, 2, 3), (16, 1, -1)])
def test_c_oracle_matches_numpy_oracle(vnk):
  from oracle import c_oracle
  if not c_oracle.available():
    import __graft_entry__ as ge
    ge.build_oracle()
  v, n, k = vnk
  rng = np.random.RandomState(v * 100 + n * 10 + k + 1)
  c = sum(v**i for i in range(n + 1))
  b, t, u = 4, 9, 5
  tab = rng.randn(b, t, c, 1 + v).astype(np.float32)
  nf = np.array([9, 6, 3, 0])
  lab = rng.randint(1, v + 1, (b, u))
  nl = np.array([4, 2, 5, 0])     # utterance 2: 5 labels in 3 frames (FrameDependent: unreachable)
  kk, fd = (0, True) if k < 0 else (k, False)
  with np.errstate(all='ignore'):
    ol, ogb, ogl = O.lattice_loss_and_grads(
        tab[..., 0].astype(np.float64), tab[..., 1:].astype(np.float64), nf, lab, nl,
        O.FullNGram(v, n), kk, fd)
  cl, cgb, cgl, _, _ = c_oracle.lattice_loss_and_grads(tab[..., 0], tab[..., 1:], nf, lab, nl, v, n, k)
  npt.assert_array_equal(np.isfinite(cl), np.isfinite(ol))
  fin = np.isfinite(ol)
  npt.assert_allclose(cl[fin], ol[fin], rtol=1e-5, atol=1e-5)
  npt.assert_allclose(cgb, ogb, rtol=1e-4, atol=2e-5)
  npt.assert_allclose(cgl, ogl, rtol=1e-4, atol=2e-5)


def test_next_state_table_known_answers():
  # tests/contexts_test.py:189-239: NextStateTable built from FullNGram(3, 2)
  table = O.FullNGram(3, 2).next_state_table()
  assert table.shape == (13, 3)
  ctx = O.NextStateTable(table)
  assert ctx.shape() == (13, 3) and ctx.start() == 0
  npt.assert_array_equal(ctx.next_state([0, 1, 3, 4, 12], [1, 2, 3, 1, 2]), [1, 5, 12, 4, 11])
  npt.assert_array_equal(ctx.next_state([0, 1, 3, 4, 12], [0, 0, 0, 0, 0]), [0, 1, 3, 4, 12])
  npt.assert_array_equal(
      ctx.forward_reduce(np.arange(39, dtype=np.float64).reshape(1, 13, 3), O.REAL),
      [[0, 0, 1, 2, 3 * 4 + 54, 4 * 4 + 54, 5 * 4 + 54, 6 * 4 + 54, 7 * 4 + 54, 8 * 4 + 54,
        9 * 4 + 54, 10 * 4 + 54, 11 * 4 + 54]])
  npt.assert_array_equal(ctx.backward_broadcast(np.arange(13).reshape(1, 13)),
                         [[[1, 2, 3]] + [[4, 5, 6], [7, 8, 9], [10, 11, 12]] * 4])
  npt.assert_array_equal(ctx.walk_states([2, 3, 1]), [0, 2, 9, 10])
  npt.assert_array_equal(ctx.walk_states([2, 0, 0, 3, 1]), [0, 2, 2, 2, 9, 10])
  # every semiring agrees with the closed-form FullNGram reduction on the same DFA
  rng = np.random.RandomState(0)
  w = rng.randn(2, 13, 3)
  full = O.FullNGram(3, 2)
  for sr in (O.REAL, O.LOG, O.MAXTROPICAL):
    npt.assert_allclose(ctx.forward_reduce(w, sr), full.forward_reduce(w, sr), rtol=1e-12)
