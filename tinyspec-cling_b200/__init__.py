"""tinyspec-cling_b200 — B200 (sm_100a) backend for tinyspec-cling's overlap-add spectral hot path.

The product is libtsb200.so (csrc/, C ABI in include/tsb200.h).  This Python package is the thin
ctypes binding used by tests and bench.py; the C++ facade that is source-compatible with the
reference's rtinc/ headers lives in rtinc/.  There is no CPU fallback: importing works without a
GPU (the host-side schedule logic is usable), every device call fails loudly.

The directory name carries a hyphen, so import it through the repo-root shim: `import tsb200`.
"""
from . import expr
from .api import (Context, Plan, Stream, TimePar, TsbError, abi_version, bind_host_to_gpu, expr_check, granular_schedule, hop_fixed, lib_path, schedule_const, shift_map,
                  F32, F64, PROG_PV_SHIFT, PROG_GAIN, WIN_NONE, WIN_HANN, WIN_SQRT_HANN)

__all__ = ["Context", "Plan", "Stream", "TimePar", "TsbError", "abi_version", "bind_host_to_gpu", "expr", "expr_check", "granular_schedule", "hop_fixed", "lib_path", "schedule_const", "shift_map",
           "F32", "F64", "PROG_PV_SHIFT", "PROG_GAIN", "WIN_NONE", "WIN_HANN", "WIN_SQRT_HANN"]
