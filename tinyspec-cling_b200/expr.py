"""expr.py — spectrum expressions for the fused spectral synthesis (TSB_GEN_EXPR, include/tsb200.h).

The reference's spectral_bytebeat hacks (hacks/spectral_bytebeat/*.cpp, hacks/readme.cpp:12-18) fill a spectrum in a
host loop, one formula per bin, e.g. voidbubbles.cpp:9-13:

    int r = (int)(t*32)*n+i;
    buf[c][i] = sin(2.0*M_PI/64.*(c+((int)pow(i,1.1))&r/3|r/250|r/136|r>>16))/pow(i+1, 0.7)*32;

Here the formula is written once with the operators below, compiled to a postfix program and evaluated by the
generator of kernel A for every (frame, channel, bin) — the spectrum never exists in host memory or HBM:

    i, c = X.bin(), X.channel()
    r = X.iparam(lambda t, n: wrap32(int(t * 32) * n)) + i
    ipow = X.itab(lambda i: int(math.pow(i, 1.1)))
    mag = X.sin((2.0 * math.pi / 64.) * X.f64(((c + ipow) & (r // 3)) | (r // 250) | (r // 136) | (r >> 16))) \
          / X.dtab(lambda i: math.pow(i + 1, 0.7)) * 32.0
    prog = X.compile(mag, 0.0)

What depends on the bin only (`itab`, `dtab`) or on the frame only (`iparam`, `dparam`: functions of the callback's
time t and the frame size n) is evaluated HERE, on the host, with Python's math module, i.e. the C library the
reference itself calls — so a truncation like (int)pow(i,1.1) cannot come out differently; the device is left with
exact integer arithmetic and the sin / cos / exp / fmod / pow of the part in which bin and time are coupled.

Integers are int64 on the device; C's `int` is 32 bits: wrap explicitly (`.i32()`) where the hack relies on it
(andshuffle.cpp's i*i).  `//` and `%` are C's truncating operators, not Python's flooring ones; x//0 and x%0 give 0.
"""
import math

import numpy as np

# opcodes: include/tsb200.h (TSB_X_*), checked against the header by tests/test_host_logic.py
OPS = ["", "I_BIN", "I_CH", "I_FRAME", "I_N", "I_CONST", "I_TAB", "I_PARAM", "D_CONST", "D_TAB", "D_PARAM",
       "I_ADD", "I_SUB", "I_MUL", "I_DIV", "I_MOD", "I_AND", "I_OR", "I_XOR", "I_SHL", "I_SHR", "I_EQ", "I_LT",
       "I_NEG", "I_NOT", "I_WRAP32", "I2D", "D2I",
       "D_ADD", "D_SUB", "D_MUL", "D_DIV", "D_FMOD", "D_POW",
       "D_NEG", "D_ABS", "D_SIN", "D_COS", "D_EXP", "D_SQRT", "D_FLOOR", "PICK", "POLAR"]
CODE = {name: k for k, name in enumerate(OPS) if name}
MAX_OPS, MAX_CONST, MAX_TABS, STACK = 128, 24, 4, 16


def wrap32(v):
    """(int)v of C for an integer that overflowed 32 bits (two's complement)"""
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def cdiv(a, b):
    """C's truncating integer division"""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


class Node:
    """A typed expression node ('i' = integer, 'd' = double)."""

    def __init__(self, ty, op, args=(), payload=None):
        self.ty, self.op, self.args, self.payload = ty, op, tuple(args), payload

    # ---- coercion
    @staticmethod
    def lift(v, like=None):
        if isinstance(v, Node):
            return v
        if isinstance(v, (bool, int, np.integer)):
            return Node("i", "I_CONST", payload=int(v))
        if isinstance(v, (float, np.floating)):
            return Node("d", "D_CONST", payload=float(v))
        raise TypeError(f"cannot use {v!r} in a spectrum expression")

    def f64(self):
        """(double)x"""
        return self if self.ty == "d" else Node("d", "I2D", [self])

    def i64(self):
        """(int)x: C truncation"""
        return self if self.ty == "i" else Node("i", "D2I", [self])

    def i32(self):
        """wrap an integer to C's 32-bit int"""
        assert self.ty == "i"
        return Node("i", "I_WRAP32", [self])

    # ---- arithmetic: int op int -> int; anything with a double -> double (C's usual conversions)
    def _bin(self, other, iop, dop, swap=False):
        other = Node.lift(other)
        a, b = (other, self) if swap else (self, other)
        if a.ty == "i" and b.ty == "i":
            if iop is None:
                raise TypeError(f"{dop} needs doubles")
            return Node("i", iop, [a, b])
        if dop is None:
            raise TypeError(f"{iop} needs integers: convert with .i64()")
        return Node("d", dop, [a.f64(), b.f64()])

    def __add__(self, o): return self._bin(o, "I_ADD", "D_ADD")
    def __radd__(self, o): return self._bin(o, "I_ADD", "D_ADD", True)
    def __sub__(self, o): return self._bin(o, "I_SUB", "D_SUB")
    def __rsub__(self, o): return self._bin(o, "I_SUB", "D_SUB", True)
    def __mul__(self, o): return self._bin(o, "I_MUL", "D_MUL")
    def __rmul__(self, o): return self._bin(o, "I_MUL", "D_MUL", True)
    def __truediv__(self, o): return self._bin(o, "I_DIV", "D_DIV")      # C's `/`: integer division for two ints
    def __rtruediv__(self, o): return self._bin(o, "I_DIV", "D_DIV", True)
    def __floordiv__(self, o): return self._bin(o, "I_DIV", None)
    def __rfloordiv__(self, o): return self._bin(o, "I_DIV", None, True)
    def __mod__(self, o): return self._bin(o, "I_MOD", "D_FMOD")
    def __rmod__(self, o): return self._bin(o, "I_MOD", "D_FMOD", True)
    def __and__(self, o): return self._bin(o, "I_AND", None)
    def __rand__(self, o): return self._bin(o, "I_AND", None, True)
    def __or__(self, o): return self._bin(o, "I_OR", None)
    def __ror__(self, o): return self._bin(o, "I_OR", None, True)
    def __xor__(self, o): return self._bin(o, "I_XOR", None)
    def __rxor__(self, o): return self._bin(o, "I_XOR", None, True)
    def __lshift__(self, o): return self._bin(o, "I_SHL", None)
    def __rshift__(self, o): return self._bin(o, "I_SHR", None)
    def eq(self, o): return self._bin(o, "I_EQ", None)
    def lt(self, o): return self._bin(o, "I_LT", None)
    def __neg__(self): return Node(self.ty, "I_NEG" if self.ty == "i" else "D_NEG", [self])
    def __invert__(self): return Node("i", "I_NOT", [self]) if self.ty == "i" else NotImplemented
    def __abs__(self): return Node("d", "D_ABS", [self.f64()])


def bin(): return Node("i", "I_BIN")            # noqa: A001 — the hack's `i`
def channel(): return Node("i", "I_CH")         # the hack's `c`
def frame_counter(): return Node("i", "I_FRAME")
def size(): return Node("i", "I_N")             # the hack's `n` = out.size
def itab(fn): return Node("i", "I_TAB", payload=fn)      # fn(i) -> int, evaluated on the host for every bin
def dtab(fn): return Node("d", "D_TAB", payload=fn)      # fn(i) -> float
def iparam(fn): return Node("i", "I_PARAM", payload=fn)  # fn(t, n) -> int, evaluated on the host for every frame
def dparam(fn): return Node("d", "D_PARAM", payload=fn)  # fn(t, n) -> float
def f64(x): return Node.lift(x).f64()
def i64(x): return Node.lift(x).i64()
def _un(op, x): return Node("d", op, [Node.lift(x).f64()])
def sin(x): return _un("D_SIN", x)
def cos(x): return _un("D_COS", x)
def exp(x): return _un("D_EXP", x)
def sqrt(x): return _un("D_SQRT", x)
def floor(x): return _un("D_FLOOR", x)
def fabs(x): return _un("D_ABS", x)
def fmod(a, b): return Node("d", "D_FMOD", [Node.lift(a).f64(), Node.lift(b).f64()])
def pow(a, b): return Node("d", "D_POW", [Node.lift(a).f64(), Node.lift(b).f64()])  # noqa: A001


class Program:
    """A compiled expression: the postfix words, the constants and the host functions behind tables / parameters."""

    def __init__(self, ops, iconst, dconst, itabs, dtabs, iparams, dparams):
        self.ops = np.asarray(ops, np.int32)
        self.iconst = np.asarray(iconst, np.int64)
        self.dconst = np.asarray(dconst, np.float64)
        self.itabs, self.dtabs, self.iparams, self.dparams = itabs, dtabs, iparams, dparams
        self._tables = {}

    def tables(self, n):
        """per-bin tables for frame size n: (int32 [n_itab, n], float64 [n_dtab, n]); bin 0 is never evaluated"""
        if n in self._tables:
            return self._tables[n]
        it = np.zeros((len(self.itabs), n), np.int32)
        dt = np.zeros((len(self.dtabs), n), np.float64)
        for k, fn in enumerate(self.itabs):
            it[k, 1:] = [wrap32(int(fn(i))) for i in range(1, n)]
        for k, fn in enumerate(self.dtabs):
            dt[k, 1:] = [float(fn(i)) for i in range(1, n)]
        self._tables[n] = (it, dt)
        return it, dt

    def params(self, times, n):
        """per-frame parameters for the callback times `times`: (int64 [F, n_iparam], float64 [F, n_dparam])"""
        ip = np.zeros((len(times), len(self.iparams)), np.int64)
        dp = np.zeros((len(times), len(self.dparams)), np.float64)
        for f, t in enumerate(times):
            for k, fn in enumerate(self.iparams):
                ip[f, k] = int(fn(float(t), n))
            for k, fn in enumerate(self.dparams):
                dp[f, k] = float(fn(float(t), n))
        return ip, dp

    def disassemble(self):
        return " ".join(OPS[w & 255] + (f"({w >> 8})" if OPS[w & 255] in ("I_CONST", "I_TAB", "I_PARAM", "D_CONST", "D_TAB",
                                                                         "D_PARAM", "PICK") else "") for w in self.ops)


def polar(mag, phase):
    """std::polar(mag, phase) / mag*exp(cplx(0, phase)): returns the pair (re, im) for compile()"""
    return ("polar", Node.lift(mag).f64(), Node.lift(phase).f64())


def compile(re, im=0.0):  # noqa: A001
    """Postfix program leaving (re, im).  `re` may be the pair polar() returns (then `im` is ignored); a value used as
    both components (the hacks' cplx(mag, mag)) is evaluated once and duplicated."""
    ops, iconst, dconst, itabs, dtabs, iparams, dparams = [], [], [], [], [], [], []

    def const_index(pool, v, same):
        for k, u in enumerate(pool):
            if same(u, v):
                return k
        pool.append(v)
        if len(pool) > MAX_CONST:
            raise ValueError("too many constants in a spectrum expression")
        return len(pool) - 1

    def fn_index(pool, fn, what):
        for k, u in enumerate(pool):
            if u is fn:
                return k
        pool.append(fn)
        if len(pool) > MAX_TABS:
            raise ValueError(f"too many {what} in a spectrum expression (at most {MAX_TABS})")
        return len(pool) - 1

    def emit(node):
        for a in node.args:
            emit(a)
        arg = 0
        if node.op == "I_CONST":
            arg = const_index(iconst, node.payload, lambda u, v: u == v)
        elif node.op == "D_CONST":
            arg = const_index(dconst, node.payload, lambda u, v: u == v and math.copysign(1, u) == math.copysign(1, v))
        elif node.op == "I_TAB":
            arg = fn_index(itabs, node.payload, "integer tables")
        elif node.op == "D_TAB":
            arg = fn_index(dtabs, node.payload, "double tables")
        elif node.op == "I_PARAM":
            arg = fn_index(iparams, node.payload, "integer frame parameters")
        elif node.op == "D_PARAM":
            arg = fn_index(dparams, node.payload, "double frame parameters")
        ops.append(CODE[node.op] | (arg << 8))

    if isinstance(re, tuple) and re and re[0] == "polar":
        emit(re[1])
        emit(re[2])
        ops.append(CODE["POLAR"])
    else:
        re_n, im_n = Node.lift(re).f64(), Node.lift(im).f64()
        emit(re_n)
        if im_n is re_n:
            ops.append(CODE["PICK"])
        else:
            emit(im_n)
    if len(ops) > MAX_OPS:
        raise ValueError(f"spectrum expression too long ({len(ops)} words, at most {MAX_OPS})")
    return Program(ops, iconst, dconst, itabs, dtabs, iparams, dparams)


# ---- the reference's stateless spectral_bytebeat hacks, transcribed operator by operator ---------------------------------
# C's precedence: * / %  >  + -  >  << >>  >  &  >  ^  >  |.  `c` is a size_t there: arithmetic with it is unsigned 64-bit,
# the same bits as int64 for the values these formulas reach.
def voidbubbles():
    """hacks/spectral_bytebeat/voidbubbles.cpp:9-13 (frame 1<<12, hop ratio 0.5, BASELINE cfg5's pattern)"""
    i, c = bin(), channel()
    r = iparam(lambda t, n: wrap32(int(t * 32) * n)) + i
    ipow = itab(lambda i: int(math.pow(i, 1.1)))
    k = ((c + ipow) & (r // 3)) | (r // 250) | (r // 136) | (r >> 16)
    mag = sin((2.0 * math.pi / 64.) * f64(k)) / dtab(lambda i: math.pow(i + 1, 0.7)) * 32.0
    return dict(frame=1 << 12, hop_ratio=0.5, program=compile(mag, 0.0))


def acidbattery():
    """hacks/spectral_bytebeat/acidbattery.cpp:9-17 (frame 1<<11, hop ratio 0.5)"""
    i = bin()
    r = iparam(lambda t, n: wrap32(int(t * 32) * n)) + i
    k = ((r >> 14) ^ (r >> 17)) % 64 + 2
    arg = dtab(lambda i: math.pow(i, 0.1)) * f64(k) * 16.0 + f64((r & (r >> 13)) % k) / (1.0 + f64(((r // 3) & (r // 1023)) % k))
    mag = sin(arg) / dtab(lambda i: math.pow(i + 1, 1.0)) * 32.0
    return dict(frame=1 << 11, hop_ratio=0.5, program=compile(mag, mag))


def andshuffle():
    """hacks/spectral_bytebeat/andshuffle.cpp:9-14 (frame 1<<17, hop ratio 0.5); i*i overflows C's int above 46340"""
    i, c, n = bin(), channel(), size()
    r = iparam(lambda t, n: cdiv(wrap32(int(t * 8) * n), 2)) + n - i
    j = r // 128 + c
    m = ((((j // 5) & (j // 7) & ((j // 64) ^ (j // 127))) ^ (j >> 16)) % 128)
    mag = sin(1.0 + fmod(f64((i * i).i32()), 1.0 + f64(m))) / dtab(lambda i: math.pow(i, 0.75)) * 32.0
    return dict(frame=1 << 17, hop_ratio=0.5, program=compile(mag, mag))


def brokenconvo():
    """hacks/spectral_bytebeat/brokenconvo.cpp:9-11 (frame 1<<15, hop ratio 0.5)"""
    i, c = bin(), channel()
    r = iparam(lambda t, n: cdiv(wrap32(int(t * math.pi) * n), 2)) + i
    m = (r // 77) & (r // 38) & (r // 24 + c)
    mag = sin(2.0 * math.pi * (f64(i) / (1.0 + fmod(1.0 + f64(m), f64(i + 1))))) / dtab(lambda i: math.pow(i + 1, 0.75)) * 64.0
    return dict(frame=1 << 15, hop_ratio=0.5, program=compile(mag, 0.0))


def ghostacid():
    """hacks/spectral_bytebeat/ghostacid.cpp:9-12 (frame 1<<14, hop ratio 0.5)"""
    i, c = bin(), channel()
    r = iparam(lambda t, n: wrap32(int(t * 4) * n)) + i
    den = 1.0 + f64((r // 97) ^ (r & (r // 2)))
    mag = sin(math.pi * fmod(f64(i // 24 + c), 1.0 + f64(r) / den / 7.0)) / dtab(lambda i: math.pow(i + 1, 1.0)) * 128.0
    return dict(frame=1 << 14, hop_ratio=0.5, program=compile(mag, mag))


HACKS = {"voidbubbles": voidbubbles, "acidbattery": acidbattery, "andshuffle": andshuffle, "brokenconvo": brokenconvo,
         "ghostacid": ghostacid}
