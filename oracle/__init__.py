"""CPU oracle for the ipp-marl environment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  The product path (``ipp_marl_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests, golden vectors or KATs for this
path ("parity unpinned" by the reference's own tests).  The oracle is pinned
instead against outputs of the *unmodified reference itself*, imported from
``/root/reference`` in the build container by ``oracle/make_golden.py`` (which
is committed, together with the fixtures it wrote to ``tests/golden/``).
"""
