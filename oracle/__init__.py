"""CPU oracle for the sgmcmc sampler hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and there only as the checker / reported CPU baseline.

PARITY UNPINNED: the reference (STOR-i/sgmcmc, R + TensorFlow 1.x + TFP) ships no
golden vectors or known-answer tests for this path, and
neither R nor TensorFlow exist in this image, so the oracle cannot be checked
against outputs of the reference itself.  It is pinned instead by (i) two
independent restatements that must agree (closed-form float64 numpy vs an
op-by-op float32 torch-autograd replica of the reference's TensorFlow graph) and
(ii) the reference's own statistical acceptance bounds
(tests/testthat/test1dGauss.R:92-95, test3dGauss.R:95-100).
"""
