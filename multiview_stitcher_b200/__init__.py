"""B200-native registration-and-fusion engine behind multiview-stitcher's
extension API (hot path only).

Public surface mirrors the reference's names for this path:

* ``fusion.fuse`` / ``fusion.fuse_np`` / ``fusion.weighted_average_fusion`` /
  ``fusion.max_fusion`` / ``fusion.simple_average_fusion``
* ``registration.phase_correlation_registration`` (``pairwise_reg_func``) /
  ``registration.pairwise_executor`` = ``pairs.pairwise_executor`` (hook A), with
  ``pairs.PairPlan`` / ``pairs.register_views`` underneath
* ``hooks.*`` (``fusion_func`` / ``weights_func`` on resampled stacks),
  ``batch.BatchFuser`` (``batch_options["batch_func"]``, hook C)
* ``hooks.content_based_dct`` / ``fusion.multi_view_deconvolution`` (the other built-in
  weights / fusion methods)
* ``pyramid.build_pyramid`` (output resolution levels), ``ngff_io.*`` (Zarr v2 / OME-Zarr 0.4
  chunk encode + write and read + decode on the device), ``distributed.*`` (one process per GPU)

Everything computes on the GPU through ``libmvs_b200.so`` (C ABI,
include/mvs_b200.h); there is no CPU fallback.
"""

from ._lib import EngineError, EngineUnavailable  # noqa: F401

__version__ = "0.1.0"
