"""Replays a call trace of the CPU stand-in (FSB_STANDIN_TRACE) through the grouping rules of the deferred queue
(flecsolve_b200/csrc/fuser.cu: flush) on the host: which statement groups -- i.e. kernel launches -- does a solver
or integrator produce, and which of them have ahead-of-time kernels?  Test / analysis infrastructure."""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from flecsolve_b200 import _lib as F

LIN2, MUL, DOT = 2, 3, 16
MAXS = 12
NAMES = {0: "set", 1: "scale", 2: "lin2", 3: "mul", 4: "div", 5: "recip", 6: "abs", 7: "adds",
         16: "dot", 17: "asum", 18: "amax", 19: "min", 20: "max", 21: "powsum"}


def program_info(stmts, dev):
    raw = np.array(stmts, dtype=np.int32).ravel()
    out, canon = np.zeros(6, dtype=np.int32), np.zeros(6 * len(stmts), dtype=np.int32)
    rc = F.lib().fsb_debug_program_info(raw.ctypes.data_as(C.POINTER(C.c_int32)), len(stmts), int(dev),
                                        out.ctypes.data_as(C.POINTER(C.c_int32)), canon.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        return None
    return bool(out[0]), tuple(tuple(r[:4]) for r in canon.reshape(-1, 6).tolist())


def describe(canon):
    return " ; ".join(f"{NAMES[op]}(" + ",".join(f"v{k}" for k in ((z, x, y) if op < 16 else (x, y)) if k >= 0) + ")"
                      for op, z, x, y in canon)


def replay(lines):
    """-> Counter {(description, registered, dev): launches}, total launches"""
    launches = collections.Counter()
    queue = []
    ids = {}

    def vid(p):
        return -1 if p == 0 else ids.setdefault(p, len(ids))

    def flush():
        q, i = list(queue), 0
        queue.clear()
        while i < len(q):
            if q[i][0] == "SPMV":
                fused = None
                for j in range(i + 1, len(q)):
                    if q[j][0] == "SPMV":
                        if q[j][2] == q[i][2]:
                            break
                        continue
                    if q[j][0] != "RED" or q[j][1] != DOT:
                        break
                    if q[i][2] not in (q[j][2], q[j][3]):
                        continue
                    u = q[j][3] if q[j][2] == q[i][2] else q[j][2]
                    if not any(q[k][0] == "SPMV" and q[k][2] == u for k in range(i + 1, j)):
                        fused = j
                    break
                if fused is not None:
                    q.insert(i + 1, q.pop(fused))
                launches[("spmv + dot" if fused is not None else "spmv", True, False)] += 1
                i += 2 if fused is not None else 1
                continue
            j = i
            while j < len(q) and j - i < MAXS and q[j][0] != "SPMV" and q[j][4] == q[i][4]:
                if q[j][0] == "EW" and q[j][5] and any(q[m][0] == "RED" and q[m][5] > 0 for m in range(i, j)):
                    break  # coefficient produced by a reduction of this run: next launch
                j += 1
            n = j - i
            while n >= 1:
                stmts = [(e[1], e[2], e[3], e[4 - 1] if False else 0) for e in q[i:i + n]]  # placeholder, rebuilt below
                stmts = []
                for e in q[i:i + n]:
                    if e[0] == "EW":
                        stmts.append((e[1], e[2], e[3], e[6]))
                    else:
                        stmts.append((e[1], -1, e[2], e[3]))
                dev = any((e[0] == "EW" and e[5]) or e[-1] for e in q[i:i + n])
                info = program_info(stmts, dev)
                if info is not None:
                    break
                n -= 1
            registered, canon = info
            launches[(describe(canon), registered, bool(dev))] += 1
            i += n
        return

    for line in lines:
        t = line.split()
        if not t:
            continue
        if t[0] == "EW":
            op, z, x, y, n, dev = int(t[1]), vid(int(t[2])), vid(int(t[3])), vid(int(t[4])), int(t[5]), int(t[6])
            if op in (LIN2, MUL) and z == x and z != y:
                x, y = y, x
            # (kind, op, z, x, n, devcoef, y, armed)
            queue.append(("EW", op, z, x, n, dev, y, 0))
        elif t[0] == "RED":
            op, x, y, n, store, armed = int(t[1]), vid(int(t[2])), vid(int(t[3])), int(t[4]), int(t[5]), int(t[6])
            queue.append(("RED", op, x, y, n, store, armed))
        elif t[0] == "SPMV":
            queue.append(("SPMV", vid(int(t[1])), vid(int(t[2])), int(t[3])))
        elif t[0] == "FLUSH":
            flush()
        if len(queue) >= 64:
            flush()
    flush()
    return launches
