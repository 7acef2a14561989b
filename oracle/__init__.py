"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Evoformer trunk hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU baseline.  The product
(``alphafold2_b200``) never imports this package and raises if its CUDA library
is missing.
"""
