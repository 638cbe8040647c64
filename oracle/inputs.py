"""Seeded synthetic inputs: moved to dpmn_b200/synth.py (numpy-only generators shared by bench.py, the tests and the
fixture generators); re-exported here for the oracle-side scripts and tests."""
from dpmn_b200.synth import image_stream, prior_branch1, prior_branch2, residuals  # noqa: F401
