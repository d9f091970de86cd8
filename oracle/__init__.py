"""CPU oracle for the BEV input-generation path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or as the timed
CPU baseline.  The product (``visiontransformer-intention-prediction_b200``) never imports it.

Two layers:

* ``oracle.c_oracle``   -- ctypes binding of ``bevgen_oracle.c`` (plain scalar C, explicit
  ``fma()`` chain, own restatement of OpenCV's line / fillPoly / warpAffine / resize
  arithmetic).  Deterministic on any host; this is THE parity checker.
* ``oracle.numpy_port`` -- the reference's numpy/OpenCV call sequence restated with the
  same library calls (cost profile of the reference); this is the timed CPU baseline
  (``cpu_baseline.kind = "port"``).

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference's own functions executed in the build container
(``tests/test_oracle_vs_reference.py``, skipped where ``/root/reference`` is absent) and
against fixtures generated from those functions (``tests/golden/``, made by
``oracle/make_golden.py``).
"""
from . import c_oracle  # noqa: F401
