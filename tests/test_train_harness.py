"""CPU checks of the synthetic train-step harness (tools/train_step.py, BASELINE configs[3]): the load generator has the
reference generator's interface (models/repvgg_edge_nose_NEW_cmlp.py:253-336: three [B,3,N] clouds of 128 / 256 / 1024 points),
its hierarchical decoder adds offsets to the coarser level like the reference, and the parameter count is the reference's
(fc1_1 = Linear(1024, 131072) dominates) plus the 224x224 edge Linear."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_generator_shapes_and_parameter_count():
    import train_step as T
    torch.manual_seed(0)
    for image in (128, 224):
        g = T.Generator(image).eval()
        with torch.no_grad():
            p1, p2, p3 = g(torch.randn(2, 3, image, image))
        assert p1.shape == (2, 3, 128) and p2.shape == (2, 3, 256) and p3.shape == (2, 3, 1024)
        assert p3.is_contiguous() and not p3.transpose(2, 1).is_contiguous()     # callers pass the transposed VIEW (train.py:163)
        # level k+1 = level k (repeated) + offsets: the mean of each group of children minus its parent is the mean offset
        n = sum(p.numel() for p in g.parameters())
        edge = 3 * (image // 4) ** 2 * 1000 + 1000
        assert abs(n - (183_575_976 - (3 * 56 * 56 * 1000 + 1000) + edge)) == 0
    assert 3 * 32 * 32 == 3072                                                   # the reference's hard-wired edge Linear(3072, 1000) is the 128x128 case


def test_reference_arm_of_the_harness_is_declared():
    import train_step as T
    assert callable(T.run) and callable(T.load_ops)
    src = open(os.path.join(ROOT, "tools", "train_step.py")).read()
    assert "oracle" in src and "DistributedDataParallel" in src and "no_sync" in src
