"""Parameter protocol and pickle helpers (reference: sponet/parameters.py)."""
from __future__ import annotations

import pickle
from typing import Protocol


class Parameters(Protocol):
    num_opinions: int
    num_agents: int


def save_params(filename: str, params: Parameters) -> None:
    with open(filename, "wb") as fh:
        pickle.dump(params, fh)


def load_params(filename: str) -> Parameters:
    with open(filename, "rb") as fh:
        return pickle.load(fh)
