"""Metropolis-Hastings samplers (host-side mirror).  ref: sampler/metropolis_hastings.py:25-373"""

from abc import abstractmethod
from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from openmcmc_b200.sampler.sampler import MCMCSampler


@dataclass
class AcceptRate:
    """Acceptance-rate bookkeeping.  ref: metropolis_hastings.py:25-66 (counts are summed over chains)."""

    def __init__(self):
        self.count = {"accept": 0, "proposal": 0}

    @property
    def acceptance_rate(self) -> float:
        return self.count["accept"] / self.count["proposal"] * 100

    def get_acceptance_rate(self) -> str:
        if self.count["proposal"] == 0:
            return "No proposals"
        return f"Acceptance rate {self.acceptance_rate:.0f}%"

    def increment_accept(self):
        self.count["accept"] += 1

    def increment_proposal(self):
        self.count["proposal"] += 1


@dataclass
class MetropolisHastings(MCMCSampler):
    """ref: metropolis_hastings.py:69-173.  The accept/reject step (log U < log alpha, strict; NaN rejects) runs
    inside the proposal kernels; per-chain accept/proposal counters live on the device."""

    step: np.ndarray = field(default_factory=lambda: np.array([0.2], ndmin=2), init=True)
    accept_rate: AcceptRate = field(default_factory=lambda: AcceptRate(), init=False)

    def _collect_accept(self, plan):
        ctx = plan.ctx(self)
        cnt = ctx.get("counters")
        if cnt is not None:
            c = cnt.cpu().numpy()
            self.accept_rate.count["accept"] += int(c[:, 0].sum())
            self.accept_rate.count["proposal"] += int(c[:, 1].sum())
            self.accept_counts = c.copy()
            cnt.zero_()

    def _after_sample(self, plan):
        self._collect_accept(plan)
