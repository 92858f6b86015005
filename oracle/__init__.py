"""CPU oracle for the SESA chunked-separation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  The product package
(``sesa_audio_separation_b200``) never imports this package and has no CPU fallback.

Parity status: PINNED against the reference itself.  ``oracle/make_golden.py`` imports the
unmodified reference from ``/root/reference`` (through the stand-in modules in ``oracle/shims``
for third-party packages that are not installed) and writes ``tests/golden/*.npz``; the
restatements here are checked against those vectors by ``tests/test_oracle_golden.py``.
Two pieces of arithmetic live in third-party packages absent from the reference tree and are
restated in ``oracle/third_party.py`` from their published algorithms (un-pinned upstream):
``rotary_embedding_torch.RotaryEmbedding`` and ``librosa.filters.mel`` — for those two the
status is "parity unpinned" (no upstream test vectors exist here); everything else is pinned.
"""
