"""CPU oracle for the hybrid-similarity -> top-K path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the timed CPU baseline.  The product package
(``tvbingefriend_recommendation_service_b200``) never imports this package and has no CPU
fallback: it raises if ``libtvbf.so`` (the sm_100a CUDA library) is missing.

What is restated here (numpy / scipy, float64 exactly as the reference promotes its inputs):

* ``oracle.cosine``          -- ``sklearn.metrics.pairwise.cosine_similarity`` (third-party, not
  vendored under /root/reference; pinned scikit-learn 1.7.2 in the reference's poetry.lock:5998;
  1.9.0 in this image), the only arithmetic the reference's path calls
  (ml/similarity_computer.py:41,58,86; scripts/populate_database.py:180-186).
* ``oracle.reference_paths`` -- variant A ``SimilarityComputer`` (ml/similarity_computer.py:30-190),
  variant B production loop (scripts/populate_database.py:170-218) and variant C service matrix
  path (services/content_based_service.py:161-236).
* ``oracle.compare``         -- the tie-aware top-K comparator (the reference's
  ``np.argsort(...)[::-1]`` has no defined tie order).
* ``oracle.real_reference``  -- runs the UNMODIFIED reference functions from /root/reference
  (with ``sqlalchemy`` / the Azure storage package stubbed) to pin the restatement and to
  generate the committed fixtures in ``tests/golden`` (``oracle/make_golden.py``).

Parity status: PINNED.  The reference's own tests hold no golden top-K vector
(tests/test_scripts/test_populate_database.py:148-205 asserts only stats keys), so the pin is
(a) the toy known answers of tests/test_ml/test_similarity_computer.py restated in
``tests/test_oracle.py`` and (b) outputs of the real reference functions executed in the build
container, committed as ``tests/golden/*.npz`` together with the generating script.
"""
