"""CPU oracle for the MaskAttn-UNet hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker or as the timed
CPU baseline.  The product path (``maskunet_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the restatement is pinned against outputs of the reference's
own classes, AST-loaded from ``/root/reference`` in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``.
"""
