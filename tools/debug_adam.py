import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import parity_util as pu
from test_parity_gpu import _oracle_adam_on_cuda_grads
from oracle import score_ref as ref
shape = pu.SHAPES["tiny"]; lr, lam = 5e-4, 1e-4
batch = pu.make_batch(shape, seed=21)
exp_p, exp_st = _oracle_adam_on_cuda_grads(shape, batch, lr, lam)
cfg, params, m = pu.make_models(shape, adam_mode="dense")
m.train(None, batch, lr, lam, keep_prob=1.0)
for name, _ in m.tensor_names():
    a = m.get_tensor(name).reshape(-1); b = exp_p[name].numpy().reshape(-1)
    line = "%-45s var neq=%d maxdiff=%.3e" % (name, int((a != b).sum()), float(np.abs(a - b).max()))
    if name not in ref.NON_TRAINABLE:
        am = m.get_tensor(name + "/Adam").reshape(-1); bm = exp_st.m[name].numpy().reshape(-1)
        av = m.get_tensor(name + "/Adam_1").reshape(-1); bv = exp_st.v[name].numpy().reshape(-1)
        line += " | m neq=%d rel=%.2e | v neq=%d rel=%.2e" % (int((am != bm).sum()), float(np.abs(am-bm).max()/(np.abs(bm).max()+1e-30)), int((av != bv).sum()), float(np.abs(av-bv).max()/(np.abs(bv).max()+1e-30)))
    print(line)
print("---- element-level check")
name = "dense_3/kernel"
cfg2, params2, m2 = pu.make_models(shape, adam_mode="dense")
old = params2[name].numpy().reshape(-1).copy()
a = m.get_tensor(name).reshape(-1); b = exp_p[name].numpy().reshape(-1)
mm = exp_st.m[name].numpy().reshape(-1); vv = exp_st.v[name].numpy().reshape(-1)
f = np.float32
alpha = f(f(lr) * np.sqrt(f(1) - f(0.999), dtype=np.float32) / (f(1) - f(0.9)))
idx = np.nonzero(a != b)[0][:5]
for i in idx:
    num = f(mm[i] * alpha); den = f(np.sqrt(vv[i], dtype=np.float32) + f(1e-8)); q = f(num / den); r = f(old[i] - q)
    print(i, "old", repr(old[i]), "m", repr(mm[i]), "v", repr(vv[i]), "cuda", repr(a[i]), "torch", repr(b[i]), "numpy", repr(r), "q", repr(q))
t = torch.tensor
for i in idx:
    q_t = (t(mm[i]) * t(alpha)) / (torch.sqrt(t(vv[i])) + t(f(1e-8)))
    print("torch scalar q", repr(float(q_t)), "vec path:", repr(float(((torch.from_numpy(mm) * t(alpha)) / (torch.sqrt(torch.from_numpy(vv)) + t(f(1e-8))))[i])))
