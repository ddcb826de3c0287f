"""tests/expr_eval.py — TEST INFRASTRUCTURE: a numpy interpreter for spectrum-expression programs (include/tsb200.h,
TSB_X_*), vectorised over (channel, bin).  It restates the opcode semantics independently of the device interpreter
(csrc/tsb_spectral.cu: GenExpr) so that the CPU suite can check the expression compiler and the transcriptions of the
reference's hacks (tinyspec-cling_b200/expr.py) against the oracle's direct restatements (oracle/orc.py: hack_spectrum)."""
import numpy as np


def _cdiv(a, b):
    with np.errstate(all="ignore"):
        q = np.abs(a) // np.where(b == 0, 1, np.abs(b))
        q = np.where((a < 0) == (b < 0), q, -q)
    return np.where(b == 0, 0, q)


def evaluate(prog, names, t, frame_counter, ch, n):
    """-> complex128 [ch, n] (bin 0 evaluated like any other; the kernel never asks for it)"""
    it, dt = prog.tables(n)
    ip, dp = prog.params([t], n)
    i = np.broadcast_to(np.arange(n, dtype=np.int64)[None, :], (ch, n))
    c = np.broadcast_to(np.arange(ch, dtype=np.int64)[:, None], (ch, n))
    full = lambda v, ty: np.full((ch, n), v, ty)
    st = []
    with np.errstate(all="ignore"):
        for w in prog.ops:
            op, arg = names[int(w) & 255], int(w) >> 8
            if op == "I_BIN": st.append(i)
            elif op == "I_CH": st.append(c)
            elif op == "I_FRAME": st.append(full(frame_counter, np.int64))
            elif op == "I_N": st.append(full(n, np.int64))
            elif op == "I_CONST": st.append(full(prog.iconst[arg], np.int64))
            elif op == "I_TAB": st.append(np.broadcast_to(it[arg].astype(np.int64)[None, :], (ch, n)))
            elif op == "I_PARAM": st.append(full(ip[0, arg], np.int64))
            elif op == "D_CONST": st.append(full(prog.dconst[arg], np.float64))
            elif op == "D_TAB": st.append(np.broadcast_to(dt[arg][None, :], (ch, n)))
            elif op == "D_PARAM": st.append(full(dp[0, arg], np.float64))
            elif op in ("I_ADD", "I_SUB", "I_MUL", "I_DIV", "I_MOD", "I_AND", "I_OR", "I_XOR", "I_SHL", "I_SHR", "I_EQ", "I_LT"):
                b, a = st.pop(), st.pop()
                st.append({"I_ADD": lambda: a + b, "I_SUB": lambda: a - b, "I_MUL": lambda: a * b, "I_DIV": lambda: _cdiv(a, b),
                           "I_MOD": lambda: np.where(b == 0, 0, a - _cdiv(a, b) * b), "I_AND": lambda: a & b, "I_OR": lambda: a | b,
                           "I_XOR": lambda: a ^ b, "I_SHL": lambda: a << (b & 63), "I_SHR": lambda: a >> (b & 63),
                           "I_EQ": lambda: (a == b).astype(np.int64), "I_LT": lambda: (a < b).astype(np.int64)}[op]())
            elif op == "I_NEG": st.append(-st.pop())
            elif op == "I_NOT": st.append(~st.pop())
            elif op == "I_WRAP32": st.append(st.pop().astype(np.int32).astype(np.int64))
            elif op == "I2D": st.append(st.pop().astype(np.float64))
            elif op == "D2I": st.append(np.trunc(st.pop()).astype(np.int64))
            elif op in ("D_ADD", "D_SUB", "D_MUL", "D_DIV", "D_FMOD", "D_POW"):
                b, a = st.pop(), st.pop()
                st.append({"D_ADD": lambda: a + b, "D_SUB": lambda: a - b, "D_MUL": lambda: a * b, "D_DIV": lambda: a / b,
                           "D_FMOD": lambda: np.fmod(a, b), "D_POW": lambda: np.power(a, b)}[op]())
            elif op in ("D_NEG", "D_ABS", "D_SIN", "D_COS", "D_EXP", "D_SQRT", "D_FLOOR"):
                a = st.pop()
                st.append({"D_NEG": lambda: -a, "D_ABS": lambda: np.abs(a), "D_SIN": lambda: np.sin(a), "D_COS": lambda: np.cos(a),
                           "D_EXP": lambda: np.exp(a), "D_SQRT": lambda: np.sqrt(a), "D_FLOOR": lambda: np.floor(a)}[op]())
            elif op == "PICK": st.append(st[-1 - arg])
            elif op == "POLAR":
                ph, mag = st.pop(), st.pop()
                st.append(mag * np.cos(ph))
                st.append(mag * np.sin(ph))
            else:
                raise ValueError(op)
    assert len(st) == 2
    return st[0] + 1j * st[1]
