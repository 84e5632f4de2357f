"""CNTM parameter container (reference: sponet/cntm/parameters.py:5-25)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass()
class CNTMParameters:
    """Noisy threshold model on a network.

    At rate ``r`` a node of opinion m in {0, 1} looks at its neighbours and switches to 1-m when the
    share of neighbours holding 1-m reaches ``threshold_m(1-m)``; at rate ``r_tilde`` it draws a
    random opinion.
    """

    network: object  # nx.Graph
    r: float
    r_tilde: float
    threshold_01: float
    threshold_10: float

    def __post_init__(self):
        self.num_agents = self.network.number_of_nodes()
        self.num_opinions = 2
