"""TEST INFRASTRUCTURE ONLY — CPU/torch restatements of the reference's STN warp stage.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / reported baseline.

Parity status: **unpinned at the kornia boundary** (the reference ships no tests or golden
vectors and kornia itself is absent from this image); pinned for everything the reference's
own files compute (``models/losses.py``, ``utils/dataset.py`` loaders and the
``Reconstructor.warp/transform_poi/predict`` tails, executed from ``/root/reference`` with
``oracle/kornia_stub.py`` standing in for kornia — see ``tools/make_golden.py``).
"""
