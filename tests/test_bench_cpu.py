"""bench.py contract checks that run without a GPU: the reference arm prints exactly one JSON line with the agreed
keys, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--workload", "tiny_tb", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "train_samples_per_sec" and j["unit"] == "samples/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_product_arm_fails_loudly_without_cuda():
    r = _run("--workload", "tiny_tb", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_rooflines_assembly_runs_on_cpu():
    """the roofline objects of the bench line (algorithmic bytes, fractions, ceilings) from stand-in step statistics"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from score_b200.synth import SHAPES
    stats = {"positions": 494592, "live": 334232, "unique_rows": 128337}
    probes = {"coatt_fwd": (1.4, 100), "emb_update": (2.8, 100), "sort": (0.0, 0)}
    g, s = bench.rooflines(SHAPES["taobao"], stats, probes)
    assert g["bytes_per_launch"] == 334232 * 68 and s["bytes_per_launch"] == 334232 * 68 + 128337 * 384
    assert abs(g["frac"] - g["achieved"] / g["peak"]) < 1e-12 and 0 < g["frac"] < 1 and 0 < s["frac"] < 1
    assert g["random_access_ceiling"]["frac_at_this_launch_size"] == 0.28 and s["random_access_ceiling"]["frac"] == 0.41
    g2, s2 = bench.rooflines(SHAPES["large_vocab_shard"], stats, {})
    assert g2["frac"] is None and g2["random_access_ceiling"] is None
    import json
    json.dumps([g, s, g2, s2])


def test_both_arms_describe_the_same_config_at_every_n():
    """the reference arm's config is the product arm's config at that N (global batch = per-GPU batch x N, same texts): the
    driver compares the two lines of a run; the scheme the product arm ran is reported outside `config`"""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from score_b200.synth import SHAPES
    shape = SHAPES["taobao"]
    for n in (1, 2, 8):
        ours = argparse.Namespace(gpus=n, zipf=0.0, adam_mode="lazy", no_graph=False, par="single" if n == 1 else "sharded")
        ref = argparse.Namespace(gpus=n, zipf=0.0, adam_mode="lazy", no_graph=False)
        a = bench.workload_config(shape, ours, shape.batch * n)
        b = bench.workload_config(shape, ref, shape.batch * max(1, ref.gpus))
        assert a == b and a["global_batch"] == 1024 * n and a["per_gpu_batch"] == 1024
        info = bench.scheme_info(ours, {"dp": 0.5, "sharded": 0.4} if n > 1 else None)
        assert info["chosen"] == ours.par and info["description"]
    r = _run("--impl", "reference", "--gpus", "2", "--workload", "tiny_tb", "--steps", "1", "--warmup", "1")
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["n_gpus"] == 2 and j["config"]["global_batch"] == 2 * j["config"]["per_gpu_batch"]


def test_line_guard_emits_once_and_rescues_the_measured_line():
    """bench.LineGuard: one JSON line whatever happens - a rescue while the extra large-vocab leg is still running prints
    the measured line with the leg marked unfinished (rank 0) and leaves with status 0 (every rank); afterwards nothing
    else is written"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod3", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for rank in (0, 1):
        r, w = os.pipe()
        codes = []
        g = bench.LineGuard(w, rank, 480, exit_fn=codes.append)
        g.rescue()                                   # not armed: nothing happens
        assert codes == [] and not g.done
        g.arm({"metric": "train_samples_per_sec", "value": 1.0})
        g.rescue()
        assert codes == [0]
        g.emit({"metric": "second line"})            # the late main thread must not add a line
        os.close(w)
        out = os.read(r, 1 << 16).decode()
        os.close(r)
        if rank == 0:
            lines = out.splitlines()
            assert len(lines) == 1
            j = json.loads(lines[0])
            assert j["value"] == 1.0 and "not finished" in j["large_vocab"]["error"]
        else:
            assert out == ""
    # normal end: armed, leg finishes, disarmed, line emitted once; a late timer does nothing
    r, w = os.pipe()
    codes = []
    g = bench.LineGuard(w, 0, 480, exit_fn=codes.append)
    g.arm({"value": 2.0}); g.disarm()
    g.emit({"value": 2.0, "large_vocab": {"value": 3.0}})
    g.rescue()
    os.close(w)
    assert codes == [] and json.loads(os.read(r, 1 << 16).decode())["large_vocab"]["value"] == 3.0
    os.close(r)
