"""CPU oracle for the face-vae training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
reported CPU baseline -- never as the thing shipped.  The product path
(``face_vae_b200``) fails loudly when its CUDA library is missing; it never
routes through this package.

Parity pin: the reference ships no tests or golden vectors,
so this restatement is pinned against outputs of the reference's own classes,
executed on CPU fp32 from ``/root/reference`` by ``tests/golden/make_golden.py``
and committed as fixtures under ``tests/golden/`` (see ``tests/test_oracle_golden.py``).
"""
