"""Deterministic synthetic parameters: moved to dpmn_b200/synth.py; re-exported here for the oracle-side scripts and tests."""
from dpmn_b200.synth import synth_params, synth_value  # noqa: F401
