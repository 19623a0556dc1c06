"""Exactness evidence for the tensor-core filter: a hypothesis-driven sweep of
chamfer_nn_tc_kernel against the CPU oracle over random shapes and adversarial coordinate distributions -- mixtures of scales
1e-6 ... 1e6 inside one cloud, one far outlier plus a tight cluster (scaled coordinates land in the fp16 SUBNORMAL range of
the split, chamfer_nn_tc.cu split_h), coordinates on the fp16 split boundaries, constant axes.  dist AND idx bit-exact."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

KINDS = ("mixed_scales", "outlier_cluster", "split_boundary", "constant_axes", "subnormal_scaled", "planar_lattice", "two_far_clusters")


def adversarial_cloud(kind, rng, b, n):
    if kind == "mixed_scales":          # every point has its own magnitude between 1e-6 and 1e6
        e = rng.uniform(-6, 6, size=(b, n, 1))
        p = rng.standard_normal((b, n, 3)) * 10.0 ** e
    elif kind == "outlier_cluster":     # a tight cluster and ONE far outlier: after the power-of-two scale the cluster's
        p = 1e-4 * rng.standard_normal((b, n, 3)) + rng.uniform(-1, 1, size=(b, 1, 3))   # coordinates are ~1e-7 ... 1e-9
        p[:, rng.integers(0, n)] = rng.uniform(500, 5000, size=3) * rng.choice([-1, 1], size=3)
    elif kind == "split_boundary":      # values whose fp16 hi part sits on a rounding boundary (k + 1/2) * 2^-11
        k = rng.integers(-2048, 2048, size=(b, n, 3))
        p = (k + 0.5) * 2.0 ** -11 + rng.choice([0.0, 2.0 ** -24, -2.0 ** -24], size=(b, n, 3))
    elif kind == "constant_axes":       # y and z constant: the search is 1-D, massive near-ties
        p = np.zeros((b, n, 3))
        p[..., 0] = rng.integers(0, 4 * n, size=(b, n)) / 7.0
        p[..., 1] = 0.25
        p[..., 2] = -3.0
    elif kind == "subnormal_scaled":    # extent ~1 but most coordinates within 2^-20 of the centre: lo parts are fp16 subnormals
        p = 2.0 ** -20 * rng.standard_normal((b, n, 3))
        p[:, :min(4, n)] = rng.uniform(-1, 1, size=(b, min(4, n), 3))
    elif kind == "planar_lattice":      # integer lattice in a plane: exact ties everywhere
        p = np.zeros((b, n, 3))
        p[..., :2] = rng.integers(0, 12, size=(b, n, 2))
    else:                               # two clusters 1e3 apart, each of size 1e-2
        c = rng.choice([0.0, 1000.0], size=(b, n, 1))
        p = c + 1e-2 * rng.standard_normal((b, n, 3))
    return np.ascontiguousarray(p, np.float32)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(kind_q=st.sampled_from(KINDS), kind_t=st.sampled_from(KINDS + ("same",)), b=st.integers(1, 5), n=st.integers(1, 2600),
       m=st.integers(1, 2600), seed=st.integers(0, 2 ** 31 - 1))
def test_tensor_core_kernel_sweep(pkg, oracle, cuda, kind_q, kind_t, b, n, m, seed):
    rng = np.random.default_rng(seed)
    x = adversarial_cloud(kind_q, rng, b, n)
    if kind_t == "same":                # targets drawn from the same distribution as the queries, some exact copies
        y = adversarial_cloud(kind_q, rng, b, m)
        k = min(n, m) // 2
        y[:, :k] = x[:, :k]
    else:
        y = adversarial_cloud(kind_t, rng, b, m)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    for variant in (3, 1):              # the tensor-core kernel, and the FFMA kernel (same scheme, fp32 filter)
        old = pkg._lib.lib.psd_chamfer_nn_variant(variant)
        try:
            out = pkg.chamfer_3DDist()(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda))
            torch.cuda.synchronize()
        finally:
            pkg._lib.lib.psd_chamfer_nn_variant(old)
        for got, w, name in zip(out, want, ("dist1", "dist2", "idx1", "idx2")):
            g = got.cpu().numpy()
            same = (g.view(np.uint32) == w.view(np.uint32)) if g.dtype == np.float32 else (g == w)
            assert same.all(), f"variant {variant} {kind_q}/{kind_t} b={b} n={n} m={m} seed={seed}: {name} differs at {np.argwhere(~same)[:3].tolist()}"


@pytest.mark.parametrize("kind", KINDS)
def test_tensor_core_kernel_adversarial_full_tile(pkg, oracle, cuda, kind):
    """The same distributions at the full resident size (2048 targets, several clouds per CTA) and in multi-tile mode."""
    rng = np.random.default_rng(11)
    for b, n, m in ((6, 2048, 2048), (1, 700, 4500)):
        x, y = adversarial_cloud(kind, rng, b, n), adversarial_cloud(kind, rng, b, m)
        old = pkg._lib.lib.psd_chamfer_nn_variant(3)
        try:
            out = pkg.chamfer_3DDist()(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda))
            torch.cuda.synchronize()
        finally:
            pkg._lib.lib.psd_chamfer_nn_variant(old)
        want = oracle.chamfer_forward(x, y, nthreads=8)
        for got, w in zip(out, want):
            g = got.cpu().numpy()
            assert ((g.view(np.uint32) == w.view(np.uint32)) if g.dtype == np.float32 else (g == w)).all(), (kind, b, n, m)
