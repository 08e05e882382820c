"""CPU oracle for the recsys-lite hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU (NumPy/SciPy) restatement of the reference's
arithmetic for the path named in BASELINE.json (ALS half-steps, SVD reconstruct,
masked top-N, observed-entry RMSE).  It exists so the CUDA path can be checked
bit-for-bit / within the stated fp tolerance on machines where ``/root/reference``
is not present (the GPU box).

Rules (enforced by tests/test_abi.py::test_product_never_imports_oracle):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import this package;
  * nothing under ``recsys_lite_b200/`` imports it -- the product path has no CPU
    fallback and raises when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * ALS: the reference's own tests hold no numeric vector for ALS (tests/test_algo.py:698-711
    checks keys/shape only), so the restatement is pinned against outputs of the reference
    itself, generated in the build container by ``tests/golden/make_golden.py`` and committed
    under ``tests/golden/``; where ``/root/reference`` is importable the tests additionally
    compare live and require bitwise equality.
  * top_n / svd_reconstruct / rmse: pinned against the literal known-answer vectors of
    tests/test_algo.py (cited per test) and against the same generated fixtures.
"""

from . import als, compare, generators, metrics, svd, topn  # noqa: F401
