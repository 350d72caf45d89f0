"""Build compile-time variants of the library next to the default one (score_b200/libscore_b200.<tag>.so) for A/B runs
on the GPU box:  SCORE_B200_LIB=score_b200/libscore_b200.<tag>.so python bench.py ...   Usage: build_variants.py KC:NST ..."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from score_b200 import build as b  # noqa: E402

for spec in sys.argv[1:]:
    kc, nst = spec.split(":")
    out = os.path.join(b.HERE, "libscore_b200.kc%s_nst%s.so" % (kc, nst))
    print(b.build_library(force=True, defines=["SCORE_TL_KC=" + kc, "SCORE_TL_NST=" + nst], variant_path=out))
