"""Parameter protocol shared by the CNVM and CNTM parameter classes (reference: sponet/parameters.py:1-20).

The reference's pickle helpers (`save_params` / `load_params`) are out of scope: parameter
objects here are plain dataclasses and pickle like any other.
"""
from __future__ import annotations

from typing import Protocol


class Parameters(Protocol):
    num_opinions: int
    num_agents: int
