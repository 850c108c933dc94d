"""CPU oracle for the cutoffvgg16 hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / reported CPU baseline.

PARITY UNPINNED: the reference (gchetty/ab-line-classifier) ships no tests, golden
vectors or saved models for this path, and its arithmetic lives in TensorFlow 2.9.0 /
Keras (requirements.txt:13), which cannot be imported in this image.  The oracle is
therefore a restatement of the published Keras semantics, cross-checked by two
independent implementations (torch-CPU fp32 and NumPy direct loops) and by hand-computed
known-answer tests, not by outputs of the reference itself.
"""
