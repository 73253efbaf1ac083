"""CPU tests (no GPU needed): host-side index math and error behaviour of the
API mirror, the C-ABI library (loads, exports every symbol of
include/last_lattice.h, validates arguments before touching CUDA), and the
no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import numpy.testing as npt
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT, golden_files


@pytest.fixture(scope='module')
def lt():
  import __graft_entry__ as ge
  ge.build()
  import last_torch_b200
  return last_torch_b200


def test_import_surface(lt):
  # /root/reference/last_torch/__init__.py:18-22
  for name in ['alignments', 'contexts', 'semirings', 'weight_fns', 'RecognitionLattice']:
    assert hasattr(lt, name)
  for name in ['Real', 'Log', 'MaxTropical', 'Cartesian', 'Expectation', 'LogLogExpectation',
               'value_shape', 'value_dtype']:
    assert hasattr(lt.semirings, name)


def test_capi_exports_every_declared_symbol(lt):
  from last_torch_b200 import _native
  header = open(os.path.join(ROOT, 'include', 'last_lattice.h')).read()
  declared = set(re.findall(r'\b(lt_[a-z0-9_]+)\s*\(', header))
  assert declared, 'no declarations found'
  handle = ctypes.CDLL(_native.LIB_PATH)
  for sym in sorted(declared):
    assert hasattr(handle, sym), f'{sym} is declared in last_lattice.h but not exported'
  # and the ctypes signatures cover exactly the declared symbols
  assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
  assert _native.lib().lt_version() >= 100


def test_capi_validates_arguments_without_a_gpu(lt):
  from last_torch_b200 import _native as N
  L = N.lib()
  # vocab_size <= 0: same wording as contexts.py:174-176
  rc = L.lt_lattice_forward(N.LOG, 0, 1, -1, None, None, None, 1, 1, None, None, None, None, None,
                            None, None, 0, None)
  assert rc == 1
  assert b'vocab_size should be > 0' in L.lt_last_error()
  rc = L.lt_lattice_forward(N.LOG, 4, -1, -1, None, None, None, 1, 1, None, None, None, None,
                            None, None, None, 0, None)
  assert rc == 1 and b'context_size should be >= 0' in L.lt_last_error()
  rc = L.lt_lattice_forward(7, 4, 1, -1, None, None, None, 1, 1, None, None, None, None, None,
                            None, None, 0, None)
  assert rc == 1 and b'unknown semiring' in L.lt_last_error()
  rc = L.lt_lattice_backward(N.MAXTROPICAL, 4, 1, -1, None, None, None, 1, 1, None, None, None,
                             None, None, None, None, 0, None)
  assert rc == 1 and b'lt_viterbi_backtrace' in L.lt_last_error()
  with pytest.raises(ValueError, match='vocab_size should be > 0'):
    N.check(L.lt_lattice_forward(N.LOG, 0, 1, -1, None, None, None, 1, 1, None, None, None, None,
                                 None, None, None, 0, None), 'lt_lattice_forward')


def test_table_kernel_path_query_without_a_gpu(lt, monkeypatch):
  """lt_table_lattice_cluster: which NextStateTable lattices run on the cluster kernels
  (csrc/lattice_table2.cu) and with how many CTAs per utterance -- host logic only."""
  from last_torch_b200 import _native as N
  L = N.lib()
  assert L.lt_get_option(b'LT_TABLE_V1') == 0 and L.lt_get_option(b'LT_TABLE_CLUSTER') == 0
  assert L.lt_get_option(b'LT_NO_SUCH_OPTION') == -1
  assert L.lt_set_option(b'LT_NO_SUCH_OPTION', 1) == 1          # LT_ERR_INVALID_ARGUMENT
  for backward in (0, 1):
    assert L.lt_table_lattice_cluster(257, 256, -1, backward) == 8   # configs[1] geometry
    assert L.lt_table_lattice_cluster(20, 8, -1, backward) == 1      # one small slab
    assert L.lt_table_lattice_cluster(257, 256, 2, backward) == 0    # FrameLabelDependent
    assert L.lt_table_lattice_cluster(300, 17, -1, backward) == 0    # rows are not 16-byte multiples
    assert L.lt_table_lattice_cluster(0, 8, -1, backward) == 0
    assert L.lt_table_lattice_cluster(1025, 32, -1, backward) == 4   # smallest cluster with slabs <= 40 KB
    assert L.lt_table_lattice_cluster(4161, 64, -1, backward) == 0   # C > 2048 / two slabs exceed 227 KB
  with N.option('LT_TABLE_CLUSTER', 2):
    assert L.lt_table_lattice_cluster(20, 8, -1, 0) == 2
    assert L.lt_table_lattice_cluster(257, 256, -1, 0) == 0          # two slabs of 129 rows do not fit
    with N.option('LT_TABLE_V1', 1):
      assert L.lt_table_lattice_cluster(20, 8, -1, 0) == 0
  assert L.lt_table_lattice_cluster(20, 8, -1, 0) == 1


def test_missing_library_fails_loudly(lt, monkeypatch):
  from last_torch_b200 import _native
  monkeypatch.setattr(_native, '_lib', None)
  monkeypatch.setattr(_native, 'LIB_PATH', '/nonexistent/liblast_lattice.so')
  with pytest.raises(_native.NativeLibraryError, match='no CPU or eager fallback'):
    _native.lib()


def test_no_cpu_fallback(lt):
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    lt.semirings.Log.plus(torch.zeros([2]), torch.zeros([2]))
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    lt.semirings.MaxTropical.sum(torch.zeros([2, 3]), dim=0)
  table = torch.zeros([1, 2, 3, 3])
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(2, 1), alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    lattice._forward(cache=None, frames=torch.zeros([1, 2, 1]), num_frames=torch.tensor([2]),
                     semiring=lt.semirings.Log)
  with pytest.raises(NotImplementedError, match='Real, Log and MaxTropical'):
    lt.semirings.kernel_id(lt.semirings.LogLogExpectation)


def test_full_ngram_index_math(lt):
  # tests/contexts_test.py:26-170 (integer index math runs wherever the tensors live)
  C = lt.contexts
  with pytest.raises(ValueError, match='vocab_size should be > 0'):
    C.FullNGram(vocab_size=0, context_size=1)
  with pytest.raises(ValueError, match='context_size should be >= 0'):
    C.FullNGram(vocab_size=1, context_size=-1)
  c0 = C.FullNGram(3, 0)
  assert (c0.num_states(), c0.shape(), c0.start()) == (1, (1, 3), 0)
  npt.assert_array_equal(c0.next_state(torch.Tensor([0, 0, 0]), torch.Tensor([0, 1, 2])), [0, 0, 0])
  npt.assert_array_equal(c0.next_state(torch.Tensor([0, 1, 2]), torch.Tensor([0, 0, 0])), [0, 1, 2])
  npt.assert_array_equal(c0.backward_broadcast(torch.Tensor([[1], [2]])),
                         [[[1, 1, 1]], [[2, 2, 2]]])
  c1 = C.FullNGram(2, 1)
  assert (c1.num_states(), c1.shape()) == (3, (3, 2))
  npt.assert_array_equal(c1.next_state(torch.Tensor([0, 1, 2]), torch.Tensor([1, 2, 1])), [1, 2, 1])
  npt.assert_array_equal(c1.backward_broadcast(torch.arange(3)), [[1, 2]] * 3)
  npt.assert_array_equal(c1.forward_reduce(torch.arange(6.).reshape(3, 2), lt.semirings.Real),
                         [0, 6, 9])
  with pytest.raises(ValueError, match=r'weights\.shape\[-2:\] should be \(3, 2\)'):
    c1.forward_reduce(torch.zeros([3, 4]), lt.semirings.Real)
  with pytest.raises(ValueError, match=r'weights\.shape\[-1\] should be 3'):
    c1.backward_broadcast(torch.zeros([4]))
  c2 = C.FullNGram(3, 2)
  assert (c2.num_states(), c2.shape()) == (13, (13, 3))
  npt.assert_array_equal(
      c2.next_state(torch.Tensor([0, 1, 3, 4, 12]), torch.Tensor([1, 2, 3, 1, 2])),
      [1, 5, 12, 4, 11])
  npt.assert_array_equal(
      c2.forward_reduce(torch.arange(39.).reshape(1, 13, 3), lt.semirings.Real),
      [[0, 0, 1, 2] + [i * 4 + 54 for i in range(3, 12)]])
  npt.assert_array_equal(c2.backward_broadcast(torch.arange(13).reshape(1, 13)),
                         [[[1, 2, 3]] + [[4, 5, 6], [7, 8, 9], [10, 11, 12]] * 4])
  assert c2.walk_states(torch.zeros([2, 3, 4], dtype=torch.int32)).shape == (2, 3, 5)
  npt.assert_array_equal(c2.walk_states(torch.Tensor([2, 3, 1])), [0, 2, 9, 10])
  npt.assert_array_equal(c2.walk_states(torch.Tensor([2, 0, 0, 3, 1])), [0, 2, 2, 2, 9, 10])
  from oracle import lattice_oracle as O
  for v, n in [(3, 2), (4, 1), (2, 3), (5, 0)]:
    npt.assert_array_equal(C.FullNGram(v, n).next_state_table(), O.FullNGram(v, n).next_state_table())


def test_alignment_topology_and_real_semiring_frame_ops(lt):
  # tests/alignments_test.py:27-47, :211-221 and the Real-semiring expansions :49-67, :171-185
  A, S = lt.alignments, lt.semirings
  fd = A.FrameDependent()
  assert (fd.num_states(), fd.start(), fd.blank_next(0), fd.lexical_next(0)) == (1, 0, 0, 0)
  assert fd.topological_visit() == [0]
  fld = A.FrameLabelDependent(max_expansions=2)
  assert fld.num_states() == 3 and fld.topological_visit() == [0, 1, 2]
  assert [fld.lexical_next(i) for i in range(3)] == [1, 2, None]
  assert [fld.blank_next(i) for i in range(3)] == [0, 0, 0]
  npt.assert_array_equal(A.shift_down(torch.Tensor([[1, 2, 3], [4, 5, 6]]), S.Log),
                         [[-np.inf, 1, 2], [-np.inf, 4, 5]])
  context = lt.contexts.FullNGram(vocab_size=2, context_size=1)
  alpha, blank, lexical = torch.rand([3]), torch.rand([3]), torch.rand([3, 2])
  nxt = fd.forward(alpha, [blank], [lexical], context, S.Real)
  npt.assert_allclose(nxt, [alpha[0] * blank[0],
                            alpha[1] * blank[1] + torch.sum(alpha * lexical[:, 0]),
                            alpha[2] * blank[2] + torch.sum(alpha * lexical[:, 1])], rtol=1e-6)
  a4, b4, l4 = torch.rand([4]), torch.rand([4]), torch.rand([4])
  npt.assert_allclose(fd.string_forward(a4, [b4], [l4], S.Real),
                      [a4[0] * b4[0], a4[1] * b4[1] + a4[0] * l4[0], a4[2] * b4[2] + a4[1] * l4[1],
                       a4[3] * b4[3] + a4[2] * l4[2]], rtol=1e-6)
  with pytest.raises(ValueError, match='blank should be'):
    fd.forward(alpha, [blank, blank], [lexical], context, S.Real)
  with pytest.raises(ValueError, match='lexical should be'):
    fd.string_forward(a4, [b4], [l4, l4], S.Real)


def test_semiring_host_behaviour(lt):
  S = lt.semirings
  assert S.value_shape({'a': torch.zeros([1, 2]), 'b': torch.ones([1, 2])}) == (1, 2)
  with pytest.raises(ValueError, match='No common shape can be derived for an empty PyTree'):
    S.value_shape(None)
  with pytest.raises(ValueError, match='A semiring value must consist of ndarrays of a common shape'):
    S.value_shape({'a': torch.zeros([1, 2]), 'b': torch.ones([2])})
  npt.assert_array_equal(S.Real.times(torch.Tensor([2]), torch.Tensor([3])), 6)
  npt.assert_array_equal(S.Real.sum(torch.Tensor([2, 3]), dim=0), 5)
  npt.assert_array_equal(S.Log.times(torch.Tensor([2]), torch.Tensor([3])), 5)
  npt.assert_array_equal(S.Log.zeros([2]), [-np.inf, -np.inf])
  npt.assert_array_equal(S.MaxTropical.ones([2]), [0, 0])
  # empty-axis sums need no kernel (semirings.py:216-220)
  npt.assert_array_equal(S.Log.sum(torch.zeros([3, 0]), dim=1), S.Log.zeros([3]))
  with pytest.raises(ValueError, match='Invalid reduction axis'):
    S.Log.sum(torch.zeros([2, 3]), dim=2)
  with pytest.raises(ValueError, match='Only int axis'):
    S.MaxTropical.sum(torch.zeros([2, 3]), dim=None)


def test_weight_fns_host_side(lt):
  W = lt.weight_fns
  # tests/weight_fns_test.py:25-41 goldens for the normalisers
  blank = torch.Tensor([2, 7]); lexical = torch.Tensor([[0, 1], [3, 5]])
  nb, nl = W.hat_normalize(blank, lexical)
  npt.assert_allclose(torch.exp(nb) + torch.exp(nl).sum(-1), [1, 1], rtol=1e-6)
  nb, nl = W.log_softmax_normalize(blank, lexical)
  npt.assert_allclose(torch.exp(nb) + torch.exp(nl).sum(-1), [1, 1], rtol=1e-6)
  # TableWeightFn lookup (tests/weight_fns_test.py:48-83)
  table = torch.arange(5 * 4 * 3).reshape([5, 4, 3]).float()
  fn = W.TableWeightFn(table)
  frame = torch.Tensor([[2, 0.5]])[0]
  b, l = fn(None, frame)
  npt.assert_array_equal(b, table[2, :, 0]); npt.assert_array_equal(l, table[2, :, 1:])
  b, l = fn(None, frame, torch.tensor(3))
  npt.assert_array_equal(b, table[2, 3, 0]); npt.assert_array_equal(l, table[2, 3, 1:])
  bt = torch.arange(2 * 5 * 4 * 3).reshape([2, 5, 4, 3]).float()
  fnb = W.TableWeightFn(bt)
  frames = torch.Tensor([[[1], [4], [0]], [[2], [2], [3]]])
  ab, al = fnb.all_frames(None, frames)
  assert ab.shape == (2, 3, 4) and al.shape == (2, 3, 4, 2)
  npt.assert_array_equal(ab[1, 2], bt[1, 3, :, 0]); npt.assert_array_equal(al[0, 1], bt[0, 4, :, 1:])
  with pytest.raises(ValueError, match='frame should have batch_dims'):
    fnb(None, torch.zeros([3, 1]))
  assert W.NullCacher()() is None
  emb = W.SharedEmbCacher(num_context_states=7, embedding_size=5)
  assert emb().shape == (7, 5) and len(list(emb.parameters())) == 1


@pytest.mark.parametrize('fname', golden_files('joint_'))
def test_joint_weight_fn_per_frame_matches_reference(lt, fname):
  """JointWeightFn.forward (the per-frame reference API, plain tensor ops) against
  the reference body run with injected weights (tests/golden/make_golden.py)."""
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  h, e = g['w_ctx'].shape
  d = g['w_frame'].shape[1]
  fn = lt.weight_fns.JointWeightFn(vocab_size=int(g['vocab']), hidden_size=h, embedding_size=e,
                                   feature_size=d)
  assert len(list(fn.parameters())) == 6        # registered once, not re-created per call (D6)
  with torch.no_grad():
    fn.context_projection.weight.copy_(torch.tensor(g['w_ctx']))
    fn.blank_projection.weight.copy_(torch.tensor(g['w_frame']))
    fn.joint_projection_to_blank.weight.copy_(torch.tensor(g['w_blank'])[None])
    fn.joint_projection_to_blank.bias.copy_(torch.tensor(g['b_blank']).reshape(1))
    fn.joint_projection_to_vocab.weight.copy_(torch.tensor(g['w_vocab']))
    fn.joint_projection_to_vocab.bias.copy_(torch.tensor(g['b_vocab']))
  blank, lexical = fn(torch.tensor(g['cache']), torch.tensor(g['frame']))
  npt.assert_allclose(blank.detach(), g['blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(lexical.detach(), g['lexical'], rtol=1e-5, atol=1e-6)
  sb, sl = fn(torch.tensor(g['cache']), torch.tensor(g['frame']), torch.tensor(g['state']))
  npt.assert_allclose(sb.detach(), g['state_blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(sl.detach(), g['state_lexical'], rtol=1e-5, atol=1e-6)
  b2, l2 = fn(torch.tensor(g['cache']), torch.tensor(g['frame']))
  npt.assert_array_equal(b2.detach(), blank.detach())   # deterministic across calls


def test_lattice_shape_errors(lt):
  # tests/lattices_test.py:68-89 (raised on the host before any kernel runs)
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(2, 1), alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(torch.zeros([4, 6, 3, 3])),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  frames = torch.rand([4, 6, 8]); nf = torch.Tensor([6, 3, 2, 1])
  labels = torch.ones([4, 4]); nl = torch.Tensor([4, 3, 1, 2])
  with pytest.raises(ValueError, match='frames and num_frames have different batch_dims'):
    lattice(frames=frames[:1], num_frames=nf, labels=labels, num_labels=nl)
  with pytest.raises(ValueError, match='labels and num_frames have different batch_dims'):
    lattice(frames=frames, num_frames=nf, labels=labels[:1], num_labels=nl)
  with pytest.raises(ValueError, match='num_labels and num_frames have different batch_dims'):
    lattice(frames=frames, num_frames=nf, labels=labels, num_labels=nl[:1])
  with pytest.raises(ValueError, match='The length of blank_mask should be equal to 1'):
    lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=lt.semirings.Log,
                     blank_mask=[torch.zeros([1]), torch.zeros([1])])


def test_next_state_table_host_logic():
  """tests/contexts_test.py:176-239 (everything that is integer plumbing)."""
  import last_torch_b200 as lt
  with pytest.raises(ValueError, match='next_state_table should have a non-zero size'):
    lt.contexts.NextStateTable(torch.zeros([1, 0], dtype=torch.int32))
  with pytest.raises(ValueError, match='next_state_table should have a non-zero size'):
    lt.contexts.NextStateTable(torch.zeros([0, 1], dtype=torch.int32))
  with pytest.raises(ValueError, match='next_state_table should have shape'):
    lt.contexts.NextStateTable(torch.zeros([1], dtype=torch.int32))
  with pytest.raises(ValueError, match='next_state_table should be an int32 ndarray'):
    lt.contexts.NextStateTable(torch.zeros([2, 3]))
  table = lt.contexts.FullNGram(vocab_size=3, context_size=2).next_state_table()
  assert table.shape == (13, 3)
  ctx = lt.contexts.NextStateTable(table.to(torch.int32))
  assert tuple(ctx.shape()) == (13, 3) and ctx.start() == 0
  npt.assert_array_equal(
      ctx.next_state(torch.Tensor([0, 1, 3, 4, 12]), torch.Tensor([1, 2, 3, 1, 2])),
      [1, 5, 12, 4, 11])
  npt.assert_array_equal(
      ctx.next_state(torch.Tensor([0, 1, 3, 4, 12]), torch.Tensor([0, 0, 0, 0, 0])),
      [0, 1, 3, 4, 12])
  npt.assert_array_equal(ctx.backward_broadcast(torch.arange(13).reshape((1, 13))),
                         [[[1, 2, 3]] + [[4, 5, 6], [7, 8, 9], [10, 11, 12]] * 4])
  assert ctx.walk_states(torch.zeros([2, 3, 4], dtype=torch.int32)).shape == (2, 3, 5)
  npt.assert_array_equal(ctx.walk_states(torch.Tensor([2, 3, 1])), [0, 2, 9, 10])
  npt.assert_array_equal(ctx.walk_states(torch.Tensor([2, 0, 0, 3, 1])), [0, 2, 2, 2, 9, 10])
  with pytest.raises(ValueError, match=r'weights\.shape\[-2:\] should be torch.Size\(\[13, 3\]\)'):
    ctx.forward_reduce(torch.zeros([4, 3]), lt.semirings.Real)
  with pytest.raises(ValueError, match=r'weights\.shape\[-1\] should be 13'):
    ctx.backward_broadcast(torch.zeros([4]))
  # CSR of incoming arcs: every arc once, grouped by destination, ascending inside a group
  t, off, arcs = ctx.kernel_tables('cpu')
  assert off[0] == 0 and off[-1] == 39 and sorted(arcs.tolist()) == list(range(39))
  flat = t.reshape(-1)
  for q in range(13):
    grp = arcs[off[q]:off[q + 1]]
    assert bool((flat[grp.long()] == q).all())
    assert grp.tolist() == sorted(grp.tolist())


def test_shared_rnn_cacher():
  """tests/weight_fns_test.py:120-193: state ordering with a fake RNN cell."""
  import last_torch_b200 as lt
  pad, start = -2, -1

  class FakeRNNCell(torch.nn.RNNCellBase):
    def __init__(self, input_size, hidden_size, bias, num_chunks, device=None, dtype=None):
      super().__init__(input_size, hidden_size, bias, num_chunks, device, dtype)

    def forward(self, inputs, carry=None):
      if carry is None:
        carry = torch.full((1, self.hidden_size), pad)
      carry = torch.concat(([carry[..., 1:], inputs[..., :1]]), dim=-1)
      return carry, carry

  embeddings = torch.broadcast_to(torch.Tensor([start, 1, 2, 3]).unsqueeze(-1), (4, 6))
  cacher = lt.weight_fns.SharedRNNCacher(
      vocab_size=3, context_size=2, rnn_size=4, rnn_embedding_size=6,
      rnn_cell=FakeRNNCell(input_size=3, hidden_size=4, bias=False, num_chunks=1))
  cacher.embedding = torch.nn.Embedding.from_pretrained(embeddings)
  npt.assert_array_equal(cacher(), [
      [pad, pad, pad, start],
      [pad, pad, start, 1], [pad, pad, start, 2], [pad, pad, start, 3],
      [pad, start, 1, 1], [pad, start, 1, 2], [pad, start, 1, 3],
      [pad, start, 2, 1], [pad, start, 2, 2], [pad, start, 2, 3],
      [pad, start, 3, 1], [pad, start, 3, 2], [pad, start, 3, 3]])
  cacher = lt.weight_fns.SharedRNNCacher(
      vocab_size=3, context_size=0, rnn_size=4, rnn_embedding_size=6,
      rnn_cell=FakeRNNCell(input_size=3, hidden_size=4, bias=False, num_chunks=1))
  cacher.embedding = torch.nn.Embedding.from_pretrained(embeddings)
  npt.assert_array_equal(cacher(), [[pad, pad, pad, start]])
  # the default cell is registered once (trainable, deterministic across calls: SURVEY D7)
  cacher = lt.weight_fns.SharedRNNCacher(vocab_size=3, context_size=2, rnn_size=4,
                                         rnn_embedding_size=6)
  assert len(list(cacher.parameters())) == 5
  out1, out2 = cacher(), cacher()
  assert out1.shape == (13, 4)
  npt.assert_array_equal(out1.detach(), out2.detach())
