"""CPU oracle for the registration-and-fusion hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on numpy + scipy, the
arithmetic of the reference's hot path (multiview-stitcher @ 629f72d) so the
CUDA engine can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
``multiview_stitcher_b200`` never does.

Parity pinning (see DESIGN.md "Oracle"):

* ``oracle.fusion`` is pinned against the reference's *own* ``fuse_np`` /
  ``transform_sim`` / ``get_blending_weights`` / ``content_based`` code, run in
  the build container with the missing data-model packages stubbed
  (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and against the
  reference's known-answer tests (``_tests/test_fusion.py``).
* ``oracle.registration`` drives the reference's own
  ``phase_correlation_registration`` candidate loop in the same way, but
  scikit-image (0.26.0, pinned in the reference's ``uv.lock``) is absent from
  this image, so ``oracle.skimage_restated`` restates its published algorithm
  (``phase_cross_correlation``, ``rescale_intensity``,
  ``structural_similarity``).  That part is anchored on the reference's
  artificial-ground-truth test (``_tests/test_registration.py:262-336``,
  0.1 px) and exact Fourier-shift self checks; SSIM / Spearman *values* are
  "parity unpinned" beyond that.
"""
