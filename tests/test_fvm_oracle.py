"""The oracle's restatement of the reference's finite-volume diffusion operator (physics/volume_diffusion/diffusion.hh:84-207)
against the reference's own known answers (physics/test/fvm_diffusion.cc:169-210: zero flux, source only, boundary sink on
the 8^3 unit cube) and against the CSR entries fsb_parcsr_create_box_fvm documents (include/fsb.h)."""
import numpy as np
import pytest

import oracle as O


def unit_cube(n):
    dx = 1.0 / n
    return n + 2, dx ** 3, [dx * dx * (1.0 / dx)] * 3, (1, 1, 1), (n + 1,) * 3


def neumann(u):  # boundary/neumann.hh:57-76 with zero flux: the boundary layer mirrors the adjacent cell
    u = u.copy()
    u[0], u[-1] = u[1], u[-2]
    u[:, 0], u[:, -1] = u[:, 1], u[:, -2]
    u[:, :, 0], u[:, :, -1] = u[:, :, 1], u[:, :, -2]
    return u


def dirichlet(u, value):  # boundary/dirichlet.hh:57-63: the boundary layer holds the value
    u = u.copy()
    u[0] = u[-1] = value
    u[:, 0] = u[:, -1] = value
    u[:, :, 0] = u[:, :, -1] = value
    return u


def test_reference_known_answers():
    E, vol, k, lo, hi = unit_cube(8)
    ones = np.ones((E, E, E))
    inner = (slice(1, -1),) * 3
    # "zero flux": beta = 1, alpha = 0, Neumann boundaries, x = 1  ->  y = 0
    v = O.fvm_diffusion_apply(neumann(ones), ones, [ones] * 3, 1.0, 0.0, vol, k, lo, hi)
    assert np.abs(v[inner]).max() < 1e-12
    # "source only": beta = 0, alpha = 1, a = 1, x = 1  ->  y = alpha vol
    v = O.fvm_diffusion_apply(neumann(ones), ones, [ones] * 3, 0.0, 1.0, vol, k, lo, hi)
    assert np.abs(v[inner] - vol).max() < 1e-12
    # "boundary sink": Dirichlet value 1e-9, x = 1: cells next to the boundary lose to it (0 < y < 1), the others see nothing
    v = O.fvm_diffusion_apply(dirichlet(ones, 1e-9), ones, [ones] * 3, 1.0, 0.0, vol, k, lo, hi)[inner]
    edge = np.ones_like(v, dtype=bool)
    edge[1:-1, 1:-1, 1:-1] = False
    assert np.abs(v[~edge]).max() < 1e-12 and np.all(v[edge] < 1.0) and np.all(v[edge] > 0.0)


def assembled_rows(a, bface, beta, alpha, vol, kface, lo, hi):
    """dense restatement of the entries include/fsb.h documents for fsb_parcsr_create_box_fvm, applied to u"""
    dim = a.ndim
    box = tuple(slice(lo[dim - 1 - d], hi[dim - 1 - d]) for d in range(dim))

    def apply(u):
        diag = np.zeros_like(a[box])
        out = np.zeros_like(a[box])
        terms = []
        for ax in range(dim):
            npax = dim - 1 - ax
            lower = tuple(slice(s.start - 1, s.stop - 1) if d == npax else s for d, s in enumerate(box))
            upper = tuple(slice(s.start + 1, s.stop + 1) if d == npax else s for d, s in enumerate(box))
            fu, fl = bface[ax][box] * kface[ax], bface[ax][lower] * kface[ax]
            diag = diag + (fu + fl)
            terms.append((ax, -(beta * fl), u[lower], -(beta * fu), u[upper]))
        for ax, cl, ul, cu, uu in reversed(terms):  # entries ascend in storage offset: z-, y-, x-, centre, x+, y+, z+
            out = out + cl * ul
        out = out + (beta * diag + (alpha * vol) * a[box]) * u[box]
        for ax, cl, ul, cu, uu in terms:
            out = out + cu * uu
        return out
    return apply


@pytest.mark.parametrize("shape", [(9,), (7, 6), (6, 5, 7)])
def test_assembled_entries_reproduce_the_matrix_free_operator(shape):
    dim = len(shape)
    rng = np.random.default_rng(dim)
    ext = tuple(n + 2 for n in shape)  # x first
    npshape = ext[::-1]
    lo, hi = (1,) * dim, tuple(n + 1 for n in shape)
    a, u = rng.uniform(0.5, 2.0, npshape), rng.standard_normal(npshape)
    bface = [rng.uniform(0.5, 2.0, npshape) for _ in range(dim)]
    kface = list(rng.uniform(0.5, 2.0, dim))
    beta, alpha, vol = 1.3, 0.7, 0.11
    ref = O.fvm_diffusion_apply(u, a, bface, beta, alpha, vol, kface, lo, hi)
    box = tuple(slice(lo[dim - 1 - d], hi[dim - 1 - d]) for d in range(dim))
    got = assembled_rows(a, bface, beta, alpha, vol, kface, lo, hi)(u)
    scale = np.abs(ref[box]).max()
    assert np.abs(got - ref[box]).max() <= 1e-13 * max(scale, 1.0) * 50
