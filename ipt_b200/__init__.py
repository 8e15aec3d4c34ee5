"""ipt_b200 — B200-native trace loop for dimalit/ipt ("Intuitive Path Tracer").

The product is the CUDA library `ipt_b200/lib/libipt_b200.so` behind the C ABI of `include/ipt_b200.h`
(+ the C++ host classes of `ipt_b200/host/` that plug it behind the reference's tracer_interfaces.h).
This Python package is only the harness glue used by tests/ and bench.py: a ctypes binding (`capi`) and the
nvcc build recipe (`build`). There is no CPU fallback: without the built library, importing `capi` raises.
"""
__all__ = ["build", "capi"]
