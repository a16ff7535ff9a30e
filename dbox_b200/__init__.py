"""dbox_b200 — a B200-native implementation of dbox's per-step world pipeline (b2World.Step) behind a C ABI.

Layout: csrc/ holds the CUDA kernels and the C-ABI library (libdbox_b200.so); world.py mirrors the reference's
b2World / b2Body / b2Fixture API over that ABI; scenes.py builds BASELINE.json's configs.
"""
from ._abi import *  # noqa: F401,F403  (constants + POD structs)
from .world import *  # noqa: F401,F403
