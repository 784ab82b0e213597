"""CPU oracle for the TSPN tracklet-pair stage.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's pair stage
(`/root/reference/lib/modeling`, `/root/reference/lib/evaluation/common.py`) so that
the CUDA path in ``tspn_b200`` can be checked against it.  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``tspn_b200``) never imports ``oracle`` and fails loudly when its CUDA library is
missing.

Parity pinning
--------------
The reference ships no tests and no golden vectors, but it is a
Python reference, so every arithmetic function on the path that *can* execute was
run in the build container from ``/root/reference`` (import stubs for the absent
``dlib`` / ``IPython`` only) by ``tests/golden/make_golden.py``; its outputs on
seeded inputs are committed under ``tests/golden/`` and this oracle is checked
against them by ``tests/test_oracle_golden.py``:

* pinned by reference outputs: ``cubic_iou`` / ``_intersect`` / ``_union``
  (trajectory.py:85-141), ``viou`` (evaluation/common.py:65-106), ``_traj_iou``
  (association.py:35-48), ``PPNHead.forward`` + ``PPN._forward_test``
  (relpn/ppn.py:79-112), ``RelationPredictor.forward`` (model.py:76-88),
  ``DPNHead.forward`` (relpn/dpn.py:55-73), ``BaseModel._forward_test``
  (model.py:53-65), ``normalize`` (utils/miscellaneous.py:32-35);
* PARITY UNPINNED by the reference (it has no code for them; the oracle *defines*
  them, see DESIGN.md "[SPEC] items"): per-frame pair geometry channels, tIoU and
  overlap window, the pooled 3x1000 relative block, anchors restated from
  anchor_generator.py:48-104 (the reference's own generator crashes on numpy>=1.24),
  span decode, the stable top-K tie rule, and the exact-order fp32 arithmetic of
  ``oracle/exact``.

Modules
-------
``geometry``  pair enumeration, vIoU variants V1/V2/V3, per-frame pair geometry
``features``  L1 normalisation, pooled relative block, [P, F] feature layout
``heads``     PPNHead, stable top-K, RelationPredictor, DPNHead, anchors, span decode,
              predict.py top-K post-processing
``exact``     C restatement (gcc) of the fixed-order fp32 arithmetic used for the
              bit-exact checks (pair indices, top-K selection, span frame bounds)
"""
