"""Verbose parity report (run on the GPU box): CUDA path vs CPU oracle, every intermediate."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import parity_util as pu  # noqa: E402


def main():
    names = sys.argv[1:] or ["tiny", "tiny_tb"]
    for name in names:
        shape = pu.SHAPES[name]
        batch = pu.make_batch(shape, seed=11)
        rep = pu.forward_backward_report(shape, batch)
        pu.print_report("forward/backward %s B=%d" % (name, batch[0].shape[0]), rep)
        rep = pu.forward_backward_report(shape, batch, keep_prob=0.8)
        pu.print_report("forward/backward %s keep_prob=0.8" % name, rep)
        batches = [pu.make_batch(shape, seed=20 + i) for i in range(3)]
        for mode in ("dense", "lazy", "sparse"):
            rep = pu.train_steps_report(shape, batches, adam_mode=mode)
            pu.print_report("3 train steps %s adam=%s" % (name, mode), rep, tol=1e-3)


if __name__ == "__main__":
    main()
