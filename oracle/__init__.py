"""oracle/ — CPU restatement of the reference's vecalign hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``speech-vecalign_b200/`` may import, link or execute anything from here; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker or the timed CPU baseline.

* ``dp_core_oracle.c`` / ``core.py``  — plain-C port of ``svecalign/vecalign/dp_core.pyx``.
* ``vecalign_oracle.py``             — numpy port of ``svecalign/vecalign/dp_utils.py``.
* ``margin_oracle.py``               — numpy port of ``svecalign/postprocess/score_align.py:118-161`` (exact flat search
  in place of the faiss index; pinned to the margin scores the reference shipped for its example).
* ``_ref/`` (git-ignored, built by ``make -C oracle ref``) — the reference's own Cython core
  compiled from ``/root/reference`` in place; ``ref_loader.py`` imports it (and, in the build
  container only, the reference's Python driver straight from ``/root/reference``).

Parity status: PINNED — see ``tests/test_oracle_golden.py`` (live comparison with the real reference in the
build container, the reference's compiled core wherever ``_ref`` travelled, and the committed fixtures) and
``tests/golden/`` (reference outputs committed for the GPU box).
"""
