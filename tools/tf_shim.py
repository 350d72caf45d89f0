"""A stand-in for the slice of the TensorFlow 1.x graph API that code/score/score.py and
code/slice_models/slice_model.py use, so that the reference's OWN model classes can be executed unmodified here
(TensorFlow 1.x cannot be installed: no wheel for Python 3.12, no network).  FIXTURE TOOLING ONLY (tools/make_golden.py).

What this pins and what it does not.  The reference's source decides the WIRING - which tensors are looked up,
concatenated, tiled, fed to which layer, in which order the variables are created and how they are named, what train() and
eval() feed and fetch.  The stand-in decides what each op computes; those op semantics are the TF-1.x ones the oracle
documents (oracle/score_ref.py header) and are themselves anchored to TensorFlow's published unit-test constants
(tests/test_tf_known_answers.py on the oracle; tests/test_tf_shim.py on this file's own graph API: GRUCell through
dynamic_rnn, AdamOptimizer.minimize, log_loss, batch_normalization, l2_loss, sequence_mask, layer auto-numbering).  A tensor here is a lazy graph
node (like tf.Tensor): ops build nodes, Session.run evaluates the fetches for a feed_dict.  Static shapes
(get_shape().as_list()) come from evaluating every node once on zero inputs with batch size 2 while the graph is built.

Ops are torch (CPU) so that minimize() can differentiate the loss the reference built.
"""
from __future__ import annotations

import contextlib
import sys
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import score_ref as ref   # op semantics shared with the oracle: GRU cell, Adam kernel, constants

int32, float32 = "int32", "float32"
_DUMMY_B = 2


class _Graph(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.variables = []          # creation order
        self.by_name = {}
        self.layer_counts = {}
        self.dtype = torch.float32
        self.dropout_masks = None    # optional iterator of injected 0/1 masks (deterministic fixtures)
        self.drawn_masks = []        # masks actually used by the last run (keep_prob < 1)


G = _Graph()


class Shape(object):
    def __init__(self, dims):
        self.dims = dims

    def as_list(self):
        return list(self.dims)


class Tensor(object):
    """lazy node: value = fn(*input values)"""

    def __init__(self, fn, inputs=(), name=None, dummy=None):
        self.fn, self.inputs, self.name = fn, tuple(inputs), name
        self.dummy = dummy if dummy is not None else fn(*[_dummy(i) for i in self.inputs])

    def get_shape(self):
        d = list(self.dummy.shape)
        if d and isinstance(self, Tensor) and not isinstance(self, Variable):
            d[0] = None if self._batch_first() else d[0]
        return Shape(d)

    def _batch_first(self):
        return self.dummy.dim() > 0 and self.dummy.shape[0] == _DUMMY_B and getattr(self, "batched", True)

    # operators of tf.Tensor used by the reference
    def __add__(self, o): return _binary(torch.add, self, o)
    def __radd__(self, o): return _binary(torch.add, o, self)
    def __sub__(self, o): return _binary(torch.sub, self, o)
    def __rsub__(self, o): return _binary(torch.sub, o, self)
    def __mul__(self, o): return _binary(torch.mul, self, o)
    def __rmul__(self, o): return _binary(torch.mul, o, self)
    def __neg__(self): return Tensor(lambda a: -a, [self])
    def __getitem__(self, idx): return Tensor(lambda a: a[idx], [self])
    __hash__ = object.__hash__


class Placeholder(Tensor):
    def __init__(self, dtype, shape, name=None):
        self.ph_dtype, self.ph_shape = dtype, list(shape)
        dims = [_DUMMY_B if s is None else s for s in self.ph_shape]
        dummy = torch.zeros(dims, dtype=torch.int64 if dtype == int32 else G.dtype)
        if dtype != int32 and not dims:
            dummy = torch.ones([], dtype=G.dtype)     # scalar floats (keep_prob, lr): 1 keeps the dummy pass finite
        Tensor.__init__(self, None, (), name=name, dummy=dummy)
        self.batched = bool(self.ph_shape) and self.ph_shape[0] is None

    def convert(self, value):
        a = np.asarray(value)
        if self.ph_dtype == int32:
            return torch.from_numpy(a.astype(np.int64))     # feeding an int32 placeholder casts (float dummy rows of the loader)
        return torch.tensor(np.asarray(a, np.float64), dtype=G.dtype)


class Variable(Tensor):
    def __init__(self, name, shape, init, trainable=True):
        self.value = torch.zeros(list(shape), dtype=G.dtype)
        self.trainable, self.init = trainable, init
        Tensor.__init__(self, None, (), name=name + ":0", dummy=self.value)
        self.batched = False
        G.variables.append(self)
        G.by_name[name] = self


def _dummy(x):
    return x.dummy if isinstance(x, Tensor) else x


def _wrap_const(x):
    if isinstance(x, Tensor):
        return x
    return x


def _binary(fn, a, b):
    ins = [x for x in (a, b) if isinstance(x, Tensor)]
    if isinstance(a, Tensor) and isinstance(b, Tensor):
        return Tensor(lambda x, y: fn(x, y), [a, b])
    if isinstance(a, Tensor):
        return Tensor(lambda x: fn(x, torch.as_tensor(b, dtype=x.dtype) if not isinstance(b, torch.Tensor) else b), [a])
    return Tensor(lambda y: fn(torch.as_tensor(a, dtype=y.dtype) if not isinstance(a, torch.Tensor) else a, y), [b])


# ------------------------------------------------------------------------------------------ module-level API
def reset_default_graph():
    G.reset()


@contextlib.contextmanager
def name_scope(name):
    yield          # tf.name_scope does not prefix variables created with get_variable / tf.layers (score.py:20,43,204)


def placeholder(dtype, shape, name=None):
    return Placeholder(dtype, shape, name)


class truncated_normal_initializer(object):
    pass


def get_variable(name, shape, initializer=None):
    return Variable(name, shape, initializer)


def constant(value, shape=None):
    return Tensor(lambda: torch.full(list(shape), float(value), dtype=G.dtype), [], dummy=torch.full(list(shape), float(value), dtype=G.dtype))


def concat(values, axis):
    return Tensor(lambda *v: torch.cat(v, dim=axis), values)


def reshape(x, shape):
    if isinstance(shape, Tensor):
        return Tensor(lambda a, s: a.reshape([int(k) for k in s]), [x, shape])
    return Tensor(lambda a: a.reshape(list(shape)), [x])


def expand_dims(x, axis):
    return Tensor(lambda a: a.unsqueeze(axis), [x])


def tile(x, multiples):
    return Tensor(lambda a: a.repeat(*multiples), [x])


def reduce_sum(x, axis=None):
    return Tensor(lambda a: a.sum() if axis is None else a.sum(dim=axis), [x])


def reduce_mean(x, axis=None):
    return Tensor(lambda a: a.mean() if axis is None else a.mean(dim=axis), [x])


def sequence_mask(lengths, maxlen, dtype=float32):
    return Tensor(lambda l: (torch.arange(maxlen).unsqueeze(0) < l.unsqueeze(1)).to(G.dtype), [lengths])


def equal(a, b):
    return Tensor(lambda x, y: x == y, [a, b])


def ones_like(x):
    return Tensor(lambda a: torch.ones_like(a), [x])


def where(cond, a, b):
    return Tensor(lambda c, x, y: torch.where(c, x, y), [cond, a, b])


def sigmoid(x):
    return Tensor(torch.sigmoid, [x])


def log(x):
    return Tensor(torch.log, [x])


def clip_by_value(x, lo, hi):
    return Tensor(lambda a: a.clamp(lo, hi), [x])


def trainable_variables():
    return [v for v in G.variables if v.trainable]


def global_variables():
    return list(G.variables)


class _NN(object):
    @staticmethod
    def embedding_lookup(params, ids):
        return Tensor(lambda p, i: p[i], [params, ids])

    relu = staticmethod(lambda x: Tensor(torch.relu, [x]))
    sigmoid = staticmethod(sigmoid)

    @staticmethod
    def softmax(x):
        return Tensor(lambda a: torch.softmax(a, dim=-1), [x])

    @staticmethod
    def l2_loss(v):
        return Tensor(lambda a: (a * a).sum() / 2, [v])

    @staticmethod
    def dropout(x, keep_prob, name=None):
        def fn(a, kp):
            kp = float(kp)
            if kp >= 1.0:
                return a        # tf.nn.dropout with keep_prob 1 returns its input
            if G.dropout_masks is not None:
                m = next(G.dropout_masks).to(a.dtype)
            else:
                m = (torch.rand(a.shape) < kp).to(a.dtype)
            G.drawn_masks.append(m)
            return a * m / kp   # x / keep_prob * mask
        return Tensor(fn, [x, keep_prob])

    @staticmethod
    def dynamic_rnn(cell, inputs, sequence_length=None, dtype=None, scope=None):
        """tf.nn.dynamic_rnn(GRUCell(H)): variables <scope>/gru_cell/{gates,candidate}/{kernel,bias} (gate bias init 1),
        outputs zeroed and state copied through beyond sequence_length; returns (outputs, final state)"""
        H = cell.num_units
        in_dim = inputs.dummy.shape[-1]
        pre = scope + "/gru_cell/"
        gk = Variable(pre + "gates/kernel", [in_dim + H, 2 * H], "glorot")
        gb = Variable(pre + "gates/bias", [2 * H], "ones")
        ck = Variable(pre + "candidate/kernel", [in_dim + H, H], "glorot")
        cb = Variable(pre + "candidate/bias", [H], "zeros")
        both = Tensor(lambda x, l, a, b, c, d: ref.gru_dynamic_rnn(x, l, a, b, c, d, H), [inputs, sequence_length, gk, gb, ck, cb],
                      dummy=(torch.zeros(_DUMMY_B, inputs.dummy.shape[1], H, dtype=G.dtype), torch.zeros(_DUMMY_B, H, dtype=G.dtype)))
        outs = Tensor(lambda t: t[0], [both], dummy=both.dummy[0])
        last = Tensor(lambda t: t[1], [both], dummy=both.dummy[1])
        return outs, last


nn = _NN()


def _unique_layer_name(base):
    n = G.layer_counts.get(base, 0)
    G.layer_counts[base] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


class _Layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, use_bias=True, name=None):
        lname = name if name is not None else _unique_layer_name("dense")
        in_dim = inputs.dummy.shape[-1]
        k = Variable(lname + "/kernel", [in_dim, units], "glorot")
        ins = [inputs, k]
        if use_bias:
            ins.append(Variable(lname + "/bias", [units], "zeros"))
        out = Tensor(lambda x, w, *b: torch.matmul(x, w) + (b[0] if b else 0), ins)
        return activation(out) if activation is not None else out

    @staticmethod
    def batch_normalization(inputs, name=None, training=False):
        assert training is False        # the reference never passes training=True: inference mode forever (score.py:69)
        lname = name if name is not None else _unique_layer_name("batch_normalization")
        F = inputs.dummy.shape[-1]
        gamma = Variable(lname + "/gamma", [F], "ones")
        beta = Variable(lname + "/beta", [F], "zeros")
        mean = Variable(lname + "/moving_mean", [F], "zeros", trainable=False)
        var = Variable(lname + "/moving_variance", [F], "ones", trainable=False)

        def fn(x, g, b, m, v):
            inv = g / torch.sqrt(v + ref.BN_EPS)
            return x * inv + (b - m * inv)
        return Tensor(fn, [inputs, gamma, beta, mean, var])


layers = _Layers()


class _Losses(object):
    @staticmethod
    def log_loss(labels, predictions, epsilon=ref.LOGLOSS_EPS):
        def fn(y, p):
            y = y.to(p.dtype)
            return (-y * torch.log(p + epsilon) - (1 - y) * torch.log(1 - p + epsilon)).mean()
        return Tensor(fn, [labels, predictions])


losses = _Losses()


class GRUCell(object):
    def __init__(self, num_units):
        self.num_units = num_units


class _TrainStep(Tensor):
    def __init__(self, opt, loss):
        Tensor.__init__(self, None, (), dummy=torch.zeros([]))
        self.opt, self.loss = opt, loss


class _Adam(object):
    def __init__(self, learning_rate):
        self.lr = learning_rate
        self.state = None

    def minimize(self, loss):
        return _TrainStep(self, loss)


class _Train(object):
    AdamOptimizer = _Adam

    class Saver(object):
        def save(self, sess, save_path=None):
            raise NotImplementedError

        restore = save


train = _Train()


class Session(object):
    """sess.run(fetches, feed_dict): every fetch is evaluated against the PRE-update variables, then a fetched train step
    applies its update (what TF does for sess.run([loss, train_step]))"""

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        feed = {}
        for k, v in (feed_dict or {}).items():
            feed[k] = k.convert(v)
        memo = {}
        steps = [f for f in fl if isinstance(f, _TrainStep)]
        for v in G.variables:
            v._leaf = v.value.detach().clone().requires_grad_(bool(steps) and v.trainable)
        G.drawn_masks = []

        def ev(node):
            if not isinstance(node, Tensor):
                return node
            if id(node) in memo:
                return memo[id(node)]
            if isinstance(node, Placeholder):
                if node not in feed:
                    raise KeyError("placeholder %s was not fed" % node.name)
                out = feed[node]
            elif isinstance(node, Variable):
                out = node._leaf
            elif isinstance(node, _TrainStep):
                out = None
            else:
                out = node.fn(*[ev(i) for i in node.inputs])
            memo[id(node)] = out
            return out

        results = [ev(f) for f in fl]
        for st in steps:
            loss = ev(st.loss)
            tv = trainable_variables()
            grads = torch.autograd.grad(loss, [v._leaf for v in tv], allow_unused=True)
            params = {v.name[:-2]: v.value for v in tv}
            gd = {v.name[:-2]: (torch.zeros_like(v.value) if g is None else g.detach()) for v, g in zip(tv, grads)}
            if st.opt.state is None:
                st.opt.state = ref.AdamState({v.name[:-2]: v.value for v in G.variables})
            ref.adam_apply(params, gd, st.opt.state, float(ev(st.opt.lr)))     # TF's ApplyAdam, oracle/score_ref.py
            self.last_grads = gd
        out = []
        for f, r in zip(fl, results):
            if isinstance(f, _TrainStep):
                out.append(None)
            elif isinstance(r, torch.Tensor):
                a = r.detach().numpy()
                out.append(a if a.ndim else a[()])
            else:
                out.append(r)
        return out[0] if single else out


def gradients(sess, loss, feed_dict):
    """d loss / d every trainable variable for a feed (what tf.gradients(loss, tf.trainable_variables()) would fetch):
    -> (loss value, {name: gradient})"""
    feed = {k: k.convert(v) for k, v in feed_dict.items()}
    memo = {}
    for v in G.variables:
        v._leaf = v.value.detach().clone().requires_grad_(v.trainable)

    def ev(node):
        if not isinstance(node, Tensor):
            return node
        if id(node) in memo:
            return memo[id(node)]
        if isinstance(node, Placeholder):
            out = feed[node]
        elif isinstance(node, Variable):
            out = node._leaf
        else:
            out = node.fn(*[ev(i) for i in node.inputs])
        memo[id(node)] = out
        return out

    val = ev(loss)
    tv = trainable_variables()
    grads = torch.autograd.grad(val, [v._leaf for v in tv], allow_unused=True)
    return val.detach(), {v.name[:-2]: (torch.zeros_like(v.value) if g is None else g.detach()) for v, g in zip(tv, grads)}


def set_variables(values, dtype=torch.float32):
    """assign every variable by its TF name (without ':0'); returns [(name, shape)] in creation order"""
    spec = []
    for v in G.variables:
        n = v.name[:-2]
        v.value = values[n].detach().clone().to(dtype)
        assert list(v.value.shape) == list(v.dummy.shape), (n, v.value.shape, v.dummy.shape)
        spec.append((n, tuple(v.value.shape)))
    return spec
