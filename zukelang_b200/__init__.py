"""zukelang_b200 — B200 (sm_100a) prover hot path for zukelang.

The package is a thin host-side mirror of the reference's OCaml interfaces
(``Curve.Bls12_381``, ``QAP.eval``, ``Groth16.Make(C)``, ``Pinocchio.Make(C)``)
over the C ABI of ``libzkb200.so`` (``include/zkb200.h``).  All arithmetic runs
in hand-written CUDA kernels; there is no CPU fallback — importing ``_lib``
without the built library, or calling it without a CUDA device, raises.
"""

__version__ = "0.1.0"
