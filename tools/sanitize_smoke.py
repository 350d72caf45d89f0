"""Tiny end-to-end run for compute-sanitizer (memcheck): fwd+bwd, 2 train steps, eval."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from score_b200.model import SCORE
from score_b200.synth import SHAPES, make_batch
for mode in ("dense", "lazy", "sparse"):
    sh = SHAPES["tiny"]
    m = SCORE(*sh.ctor_args(), adam_mode=mode, use_graph=False)
    b = make_batch(sh, seed=1)
    print(mode, "fwdbwd loss", m.forward_backward(b, 1e-4))
    for i in range(3):
        print(mode, "train loss", m.train(None, make_batch(sh, seed=2 + i), 5e-4, 1e-4))
    p, l, loss = m.eval(None, b, 1e-4)
    print(mode, "eval loss", loss, p[:3])
    m.close()
print("SANITIZE_SMOKE_OK")
