"""Run-time compiled program kernels on hardware (opt-in until they have run on a B200: this file is NOT
selected by `-m gpu`; run it with `python -m pytest tests/test_zz_jit_hw.py` on a GPU box).
A statement group must give the same bits through the run-time kernel (FSB_OPT_JIT=1) as through the generic
program kernel (0): both evaluate the same statements per element, and reductions fold per-thread partials of
the same grid-stride loop only when the grids agree -- so vectors are compared exactly, reductions to rounding."""
import numpy as np
import pytest

from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.jit_hw


@pytest.fixture()
def gpu():
    if F.device_count() == 0:
        pytest.skip("no CUDA device")
    c = F.Context(0)
    yield c
    c.close()


def _group(ctx, vecs):
    x, y, z, w = vecs
    w.linear_sum(2.0, x, -3.0, y)
    w.multiply(w, z)
    w.add_scalar(w, 0.25)
    z.axpy(0.5, w, z)
    t1, t2 = w.dot_token(z), z.sumsq_token()
    m = w.max()
    return w.download(), z.download(), ctx.get(t1), ctx.get(t2), m


@pytest.mark.parametrize("layout", ["flat", "box"])
def test_runtime_kernel_equals_generic_kernel(gpu, layout):
    rng = np.random.default_rng(0)
    if layout == "flat":
        n = 100003
        make = lambda: gpu.vector(n)
    else:
        ext, lo, hi = (67, 45, 31), (1, 1, 1), (66, 44, 30)
        make = lambda: gpu.box_vector(ext, lo, hi)
    results = []
    for jit in (0, 1):
        gpu.set_option("jit", jit)
        vecs = [make() for _ in range(4)]
        rng = np.random.default_rng(1)
        for v in vecs:
            v.upload(rng.standard_normal(v.n))
        gpu.reset_stats()
        results.append(_group(gpu, vecs))
        assert (gpu.stat("jit_groups") > 0) == bool(jit)
        for v in vecs:
            v.destroy()
    gpu.set_option("jit", 0)
    (w0, z0, d0, s0, m0), (w1, z1, d1, s1, m1) = results
    assert np.array_equal(w0, w1) and np.array_equal(z0, z1) and m0 == m1
    assert abs(d0 - d1) <= 1e-12 * abs(d0) and abs(s0 - s1) <= 1e-12 * abs(s0)


def test_multi_component_bicgstab_same_iterates(gpu):
    nn = 32
    n = nn ** 3
    A0 = F.ParCSR.stencil(gpu, 7, nn, nn, nn)
    A1 = F.ParCSR.stencil(gpu, 7, nn, nn, nn, 1e-3, 1.0)
    rng = np.random.default_rng(3)
    b = np.concatenate([np.zeros(n), rng.random(n)])
    out = []
    for jit in (0, 1):
        gpu.set_option("jit", jit)
        gpu.reset_stats()
        x, info, hist = H.solve_multi2(gpu, A0, A1, b, np.full(2 * n, 2.0), solver="bicgstab", rtol=1e-8, maxiter=300,
                                       history_cap=300)
        out.append((x, info.iters, hist, gpu.stat("jit_groups")))
    gpu.set_option("jit", 0)
    assert out[1][3] > 0 and out[0][3] == 0
    assert abs(out[0][1] - out[1][1]) <= 1
    k = min(len(out[0][2]), len(out[1][2]), 40)
    assert np.allclose(out[0][2][:k], out[1][2][:k], rtol=1e-8)
    A0.destroy(); A1.destroy()
