"""Seeded inputs of the five BASELINE.json configs -- re-exported from recsys_lite_b200.synth (TEST INFRASTRUCTURE).

The generators carry no arithmetic of the hot path, so they live with the product (bench.py needs them on the
GPU leg, which must not import oracle/); the oracle-side name is kept for the tests.
"""
from recsys_lite_b200.synth import *  # noqa: F401,F403
from recsys_lite_b200.synth import CONFIGS  # noqa: F401
