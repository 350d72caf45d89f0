"""Capacity bounds of the run-descriptor lists of the scatter kernel (csrc/scatter.cu: emb_runs_kernel ->
emb_update_kernel), checked on the CPU: the constants and the capacity formulas are read from the source, the
classification of emb_runs_kernel is restated here, and adversarial run-length mixes must fit the buffers the host
allocates from those formulas (model.cu: 4 * emb_runs_long_cap(N) ints, emb_runs_part_cap(N) slots)."""
import os
import re

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

SRC = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "score_b200", "csrc", "scatter.cu")).read()


def const(name):
    return int(re.search(r"constexpr int %s = (\d+);" % name, SRC).group(1))


RUN_S_MAX, RUN_CHUNK, RUN_M_FLOOR, RUN_M_DEFAULT = const("RUN_S_MAX"), const("RUN_CHUNK"), const("RUN_M_FLOOR"), const("RUN_M_DEFAULT")


def long_cap(n):
    m = re.search(r"int64_t emb_runs_long_cap\(int64_t n\) \{ return (.*?); \}", SRC).group(1)
    assert m == "n / (RUN_S_MAX + 1) + n / (RUN_M_FLOOR + 1) + n / RUN_CHUNK + 16", m
    return n // (RUN_S_MAX + 1) + n // (RUN_M_FLOOR + 1) + n // RUN_CHUNK + 16


def part_cap(n):
    m = re.search(r"int64_t emb_runs_part_cap\(int64_t n\) \{ return (.*?); \}", SRC).group(1)
    assert m == "2 * (n / RUN_CHUNK) + 16", m
    return 2 * (n // RUN_CHUNK) + 16


def classify(run_lengths, m_max):
    """emb_runs_kernel: S (<= 4 entries), M (5 .. m_max), L (> m_max; beyond RUN_CHUNK one descriptor + one slot per chunk)"""
    n_s = n_m = n_l = slots = 0
    for c in run_lengths:
        if c <= RUN_S_MAX:
            n_s += 1
        elif c <= m_max:
            n_m += 1
        elif c <= RUN_CHUNK:
            n_l += 1
        else:
            k = -(-c // RUN_CHUNK)
            n_l += k
            slots += k
    return n_s, n_m, n_l, slots


def check(run_lengths, zeros, m_max):
    n = int(sum(run_lengths)) + zeros           # positions with key 0 are in the sorted list but form no run
    n_s, n_m, n_l, slots = classify(run_lengths, m_max)
    assert n_s <= n                                   # runs buffer: 8 * N ints = one 32-byte descriptor per position
    assert n_m + n_l <= long_cap(n), (n, n_m, n_l)    # M grows from the front, L from the back of ONE list
    assert slots <= part_cap(n), (n, slots)


def test_constants_are_what_the_bounds_assume():
    assert RUN_S_MAX == 4 and RUN_M_FLOOR >= RUN_S_MAX + 1 and RUN_M_DEFAULT >= RUN_M_FLOOR and RUN_CHUNK >= 256


@pytest.mark.parametrize("m_max", [16, 32, 512, 4096])
def test_adversarial_mixes_fit(m_max):
    for n in (1, 4, 5, 17, 2048, 2049, 4097, 100000, 494592):
        check([n], 0, m_max)                                              # one run
        check([5] * (n // 5) + [n % 5] * (1 if n % 5 else 0), 0, m_max)   # as many tier-M runs as possible
        k = max(RUN_M_FLOOR, m_max) + 1
        check([k] * (n // k) + ([n % k] if n % k else []), 0, m_max)      # as many one-chunk tier-L runs as possible
        c = RUN_CHUNK + 1
        check([c] * (n // c) + ([n % c] if n % c else []), 0, m_max)      # as many two-chunk runs as possible
        check([1] * n, n // 3, m_max)


@settings(max_examples=300, deadline=None)
@given(st.lists(st.one_of(st.integers(1, 40), st.integers(2000, 9000), st.integers(1, 300000)), min_size=1, max_size=60),
       st.integers(0, 5000), st.sampled_from([16, 32, 512]))
def test_random_mixes_fit(run_lengths, zeros, m_max):
    check(run_lengths, zeros, m_max)
