"""CPU check of the tensor-core filter's exactness argument (tools/tc_filter_model.py restates the kernel's operand splits and
margin test in numpy): over adversarial clouds no query that the margin test trusts may have its true argmin outside the chunk
the filter picked.  The same sweep WITHOUT the absolute (fp16-subnormal) margin term does produce such queries -- the hole that
tests/test_gpu_tc_hypothesis.py found in round 2 -- which pins the term's purpose."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def test_margin_constants_match_the_kernel():
    import tc_filter_model as M
    src = open(os.path.join(ROOT, "3d-pointcloudreconstruction_b200", "csrc", "chamfer_nn_tc.cu")).read()
    assert "2.0e-6f * S" in src and "__fmaf_rn(Seff, 2.5e-6f, 1.5e-7f * (0.5f + __fmaf_rn(2.6f, qn, 0.9f * rho)))" in src
    assert (float(M.RHO_SLACK), float(M.MARGIN_REL), float(M.MARGIN_ABS)) == (np.float32(2.0e-6), np.float32(2.5e-6), np.float32(1.5e-7))


def test_no_trusted_query_leaves_its_chunk_on_adversarial_clouds():
    import tc_filter_model as M
    from test_gpu_tc_hypothesis import KINDS, adversarial_cloud
    rng = np.random.default_rng(7)
    with_term = without_term = trusted = 0
    for ex in range(21):
        kq = KINDS[ex % len(KINDS)]
        kt = KINDS[int(rng.integers(len(KINDS)))] if ex % 3 else kq
        n, m = int(rng.integers(1, 1100)), int(rng.integers(1, 1100))
        x, y = adversarial_cloud(kq, rng, 1, n)[0], adversarial_cloud(kt, rng, 1, m)[0]
        for qs, ts in ((x, y), (y, x)):
            tr, vio = M.model(qs, ts, abs_term=True)
            with_term += int(vio.sum()); trusted += int(tr.sum())
            without_term += int(M.model(qs, ts, abs_term=False)[1].sum())
    assert with_term == 0 and trusted > 1000
    assert without_term > 0      # the relative margin alone is unsound once the split's lo terms are fp16 subnormals


def test_the_round2_counterexample():
    """mixed_scales b=1 n=53 m=1 seed=0: the case hypothesis shrank to on the B200."""
    import tc_filter_model as M
    from test_gpu_tc_hypothesis import adversarial_cloud
    rng = np.random.default_rng(0)
    x = adversarial_cloud("mixed_scales", rng, 1, 53)[0]
    y = adversarial_cloud("mixed_scales", rng, 1, 1)[0]
    assert M.model(y, x, abs_term=False)[1].any() and not M.model(y, x, abs_term=True)[1].any()
