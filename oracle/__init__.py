"""CPU oracle for the SLEAP batched-frame inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sleap_b200/`` may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker or as the timed
CPU baseline -- never as the shipped product path.

The oracle is a NumPy float32 (+ torch-CPU fp32 for the convolutions, SciPy for the
linear-sum-assignment, NetworkX for the edge ordering -- the same third-party
routines the reference calls) restatement of:

  sleap/nn/peak_finding.py            -> oracle/peak_finding.py
  sleap/nn/paf_grouping.py            -> oracle/paf_grouping.py
  sleap/nn/data/instance_cropping.py  -> oracle/tf_ops.py (bbox helpers)
  sleap/nn/data/normalization.py, resizing.py -> oracle/preprocess.py
  sleap/nn/data/confidence_maps.py, edge_maps.py -> oracle/synth.py
  sleap/nn/architectures/{unet,encoder_decoder,hourglass}.py, heads.py, model.py
                                      -> oracle/convnet.py
  sleap/nn/inference.py (layer call()s) -> oracle/layers.py

Parity pinning: TensorFlow is not installable in the build container, so the
reference itself cannot be run.  The oracle is pinned against every analytic
known-answer vector in the reference's own tests for this path (SURVEY.md
Appendix C; see tests/test_oracle_*.py).  TF-kernel-internal semantics
(crop_and_resize, dilation2d, resize, round-half-even, ...) are restated from the
TF 2.7 kernel definitions and flagged in each docstring; the numeric output of the
conv network is pinned by no reference test ("parity unpinned" for conv numerics --
see DESIGN.md).
"""
