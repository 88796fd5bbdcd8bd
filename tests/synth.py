"""Seeded synthetic inputs — shared with bench.py; the definitions live in the package."""
from multiviewstitch_b200.synth import *  # noqa: F401,F403
