"""Statement programs on the host side (flecsolve_b200/csrc/program.h, fuser.cu): the canonical form that decides
which kernel a queued statement group gets, and the guarantee that every group flecsolve's own Krylov loops
produce has an ahead-of-time kernel (a missing registration would not break results -- the generic kernel takes
over -- but it would silently halve the bandwidth of that step; this is the regression guard).  Pure host code."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from flecsolve_b200 import _lib as F

SET, SCALE, LIN2, MUL, DIV, RECIP, ABS, ADDS = range(8)
DOT, ASUM, AMAX, MIN, MAX, POWSUM = range(16, 22)
READS_X = {SCALE, LIN2, MUL, DIV, RECIP, ABS, ADDS, DOT, ASUM, AMAX, MIN, MAX, POWSUM}
READS_Y = {LIN2, MUL, DIV, DOT}


def info(stmts, dev=False):
    raw = np.array(stmts, dtype=np.int32).ravel()
    out, canon = np.zeros(6, dtype=np.int32), np.zeros(6 * len(stmts), dtype=np.int32)
    rc = F.lib().fsb_debug_program_info(raw.ctypes.data_as(C.POINTER(C.c_int32)), len(stmts), int(dev),
                                        out.ctypes.data_as(C.POINTER(C.c_int32)), canon.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, F.lib().fsb_last_error().decode()
    return dict(registered=bool(out[0]), nv=int(out[1]), ns=int(out[2]), nr=int(out[3]), load=int(out[4]), store=int(out[5]),
                canon=canon.reshape(-1, 6).tolist())


# vector names as the solver templates use them; any distinct small integers would do
x, r, p, w, z, dinv, v, s, t, rt, ph, sh, b = range(13)

KRYLOV_GROUPS = {
    # cg.hh: { x += a p ; r -= a w ; |r|^2 }   { z = dinv r ; r.z }   { z = r ; r.z }   { p = b p + z }
    "cg update": ([(LIN2, x, p, x), (LIN2, r, w, r), (DOT, -1, r, r)], False),
    "cg jacobi + r.z": ([(MUL, z, dinv, r), (DOT, -1, r, z)], False),
    "cg identity + r.z": ([(SCALE, z, r, -1), (DOT, -1, r, z)], False),
    "cg direction": ([(LIN2, p, z, p)], False),
    "residual + norm": ([(LIN2, r, b, r), (DOT, -1, r, r)], False),
    # cg_device.hh: the whole update in one pass, and the direction with a device-resident beta
    "cg_device update (jacobi)": ([(LIN2, x, p, x), (LIN2, r, w, r), (DOT, -1, r, r), (MUL, z, dinv, r), (DOT, -1, r, z)], True),
    "cg_device update (identity)": ([(LIN2, x, p, x), (LIN2, r, w, r), (DOT, -1, r, r), (SCALE, z, r, -1), (DOT, -1, r, z)], True),
    "cg_device direction": ([(LIN2, p, z, p)], True),
    # cg_sr.hh: the whole vector side of a single-reduction CG iteration in one pass (z plays u, t plays s = A p)
    "cg_sr update (jacobi)": ([(LIN2, p, z, p), (LIN2, t, w, t), (LIN2, x, p, x), (LIN2, r, t, r), (DOT, -1, r, r),
                               (MUL, z, dinv, r), (DOT, -1, r, z)], True),
    "cg_sr update (identity)": ([(LIN2, p, z, p), (LIN2, t, w, t), (LIN2, x, p, x), (LIN2, r, t, r), (DOT, -1, r, r),
                                 (SCALE, z, r, -1), (DOT, -1, r, z)], True),
    # gmres.hh: modified Gram-Schmidt step, normalise + store (+ precondition)
    "gmres mgs step": ([(LIN2, v, p, v), (DOT, -1, v, w)], False),
    "gmres normalise + copy": ([(SCALE, v, v, -1), (SCALE, w, v, -1)], False),
    "gmres normalise + copy + jacobi": ([(SCALE, v, v, -1), (SCALE, w, v, -1), (MUL, z, dinv, w)], False),
    # bicgstab.hh: direction (+ preconditioner), half step, full update
    "bicgstab direction + jacobi": ([(LIN2, p, v, p), (LIN2, p, r, p), (MUL, ph, dinv, p)], False),
    "bicgstab direction + identity": ([(LIN2, p, v, p), (LIN2, p, r, p), (SCALE, ph, p, -1)], False),
    "bicgstab half step": ([(LIN2, s, v, r), (DOT, -1, s, s)], False),
    "bicgstab update": ([(LIN2, x, ph, x), (LIN2, x, sh, x), (LIN2, r, t, s), (DOT, -1, r, r)], False),
    # the same loops on a two-component vec::multi (every call fans out per component; reductions per component)
    "cg update, 2 components": ([(LIN2, x, p, x), (LIN2, 20 + x, 20 + p, 20 + x), (LIN2, r, w, r), (LIN2, 20 + r, 20 + w, 20 + r),
                                 (DOT, -1, 20 + r, 20 + r), (DOT, -1, r, r)], False),
    "cg identity + r.z, 2 components": ([(SCALE, z, r, -1), (SCALE, 20 + z, 20 + r, -1), (DOT, -1, 20 + r, 20 + z), (DOT, -1, r, z)], False),
    "bicgstab half step, 2 components": ([(LIN2, s, v, r), (LIN2, 20 + s, 20 + v, 20 + r), (DOT, -1, 20 + s, 20 + s), (DOT, -1, s, s)], False),
    # fcg: d, q recurrences + iterate and residual updates + norm
    "fcg update": ([(LIN2, 30, v, 30), (LIN2, 31, w, 31), (LIN2, x, 30, x), (LIN2, r, 31, r), (DOT, -1, r, r)], False),
}


@pytest.mark.parametrize("name", sorted(KRYLOV_GROUPS))
def test_groups_of_the_krylov_loops_have_ahead_of_time_kernels(name):
    stmts, dev = KRYLOV_GROUPS[name]
    assert info(stmts, dev)["registered"], name


def test_single_statements_in_every_alias_pattern_are_registered():
    for op in (LIN2, MUL, DIV):
        for zz, xx, yy in ((2, 0, 1), (1, 0, 1), (0, 0, 1), (1, 0, 0), (0, 0, 0)):
            if op in (LIN2, MUL) and (zz, xx, yy) == (0, 0, 1):
                continue  # the C ABI normalises z == x to the y operand for commutative statements
            assert info([(op, zz, xx, yy)])["registered"], (op, zz, xx, yy)
    for op in (SCALE, RECIP, ABS, ADDS):
        assert info([(op, 1, 0, -1)])["registered"] and info([(op, 0, 0, -1)])["registered"]
    assert info([(SET, 0, -1, -1)])["registered"]
    for op in (ASUM, AMAX, MIN, MAX, POWSUM):
        assert info([(op, -1, 0, -1)])["registered"]
    assert info([(DOT, -1, 0, 1)])["registered"] and info([(DOT, -1, 0, 0)])["registered"]


def test_canonical_form_numbers_slots_by_first_appearance():
    i = info([(LIN2, 7, 3, 7), (LIN2, 9, 5, 9), (DOT, -1, 9, 9)])  # CG's update with arbitrary names
    assert i["canon"] == [[LIN2, 1, 0, 1, 0, 1], [LIN2, 3, 2, 3, 2, 3], [DOT, 0, 3, 3, -1, -1]]
    assert (i["nv"], i["ns"], i["nr"]) == (4, 4, 1)
    assert i["load"] == 0b1111 and i["store"] == 0b1010  # p, x, w, r are read; x and r are written back
    # a value produced inside the group is not loaded from memory
    j = info([(SCALE, 1, 0, -1), (MUL, 2, 1, 1)])
    assert j["load"] == 0b001 and j["store"] == 0b110


@st.composite
def groups(draw):
    n = draw(st.integers(1, 8))
    ids = st.integers(0, 5)
    out = []
    for _ in range(n):
        op = draw(st.sampled_from([SET, SCALE, LIN2, MUL, DIV, RECIP, ABS, ADDS, DOT, ASUM, AMAX, MIN, MAX]))
        xx = draw(ids) if op in READS_X else -1
        yy = draw(ids) if op in READS_Y else -1
        zz = -1 if op >= 16 else draw(ids)
        out.append((op, zz, xx, yy))
    if sum(1 for o in out if o[0] >= 16) > 4:
        out = [o for o in out if o[0] < 16] or [(SET, 0, -1, -1)]
    return out


@settings(max_examples=200, deadline=None)
@given(groups(), st.permutations(list(range(6))), st.integers(1, 40))
def test_canonical_form_ignores_the_names_of_vectors(stmts, perm, shift):
    """two groups that differ only by an injective renaming of their vectors are the same program; and the
    canonical form is idempotent"""
    ren = lambda k: k if k < 0 else perm[k] * 3 + shift
    other = [(op, ren(zz), ren(xx), ren(yy)) for op, zz, xx, yy in stmts]
    a, b_ = info(stmts), info(other)
    assert a == b_
    again = info([(c[0], c[1] if c[0] < 16 else -1, c[2], c[3]) for c in a["canon"]])
    assert again["canon"] == a["canon"]


def test_aliasing_changes_the_program():
    assert info([(LIN2, 2, 0, 1)])["canon"] != info([(LIN2, 1, 0, 1)])["canon"]
    assert info([(LIN2, 1, 0, 1)])["canon"] != info([(LIN2, 1, 0, 0)])["canon"]
    assert info([(DOT, -1, 0, 0)])["canon"] != info([(DOT, -1, 0, 1)])["canon"]
