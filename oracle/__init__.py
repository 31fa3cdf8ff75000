"""CPU oracle for the mini-frame GP log-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the path that is shipped.

Parity status: PINNED against the reference itself.  The reference ships no
golden vectors (its only test is ``assert True``,
miniframe/tests/test_kernels.py:1-4), so ``tests/golden/make_golden.py`` runs
the unmodified reference (imported from /root/reference through
``oracle/ref_shim.py``) and commits its inputs/outputs as fixtures under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against
those fixtures and, when /root/reference is present, against the live
reference.
"""
