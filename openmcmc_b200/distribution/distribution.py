"""Distribution base class, Gamma, Poisson, Uniform (host-side mirror).  ref: distribution/distribution.py:28-523

Objects are declarative (they name state entries); `log_p`, `grad_log_p` and `rvs` on host dict states are evaluated
by the CUDA kernels behind `hostcalls` — there is no numpy/scipy arithmetic here.  Derivatives of Gamma / Poisson:
the reference only has central finite differences; the device evaluates the same stencil
(`method="fd"`, the parity path) or the analytic derivative (`method="analytic"`, what the samplers use).
"""

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Union

import numpy as np

from openmcmc_b200.parameter import Identity, LinearCombination, MixtureParameterVector


@dataclass
class Distribution(ABC):
    """ref: distribution.py:28-198"""

    response: str

    @abstractmethod
    def log_p(self, state: dict, by_observation: bool = False):
        """Log-density of state[self.response]."""

    @abstractmethod
    def rvs(self, state: dict, n: int = 1):
        """Random draws (p x n)."""

    @property
    @abstractmethod
    def _dist_params(self) -> list:
        """Parameter labels excluding the response."""

    @property
    def param_list(self) -> list:
        """ref: distribution.py:79-88"""
        return [self.response] + self._dist_params

    def grad_log_p(self, state: dict, param: str, hessian_required: bool = True, method: str = "fd"):
        """Gradient of the POSITIVE log-pdf and Hessian of the NEGATIVE log-pdf w.r.t. `param`.

        ref: distribution.py:90-122 (default = central finite differences, step 1e-4, :124-198).
        """
        from openmcmc_b200 import hostcalls

        return hostcalls.grad_log_p(self, state, param, hessian_required, method)


def _as_parameter(value, allowed, what):
    if isinstance(value, str):
        value = Identity(value)
    if not isinstance(value, allowed):
        raise TypeError(f"{what} expected to be one of [Identity, LinearCombination, MixtureParameterVector]")
    return value


@dataclass
class Gamma(Distribution):
    """Gamma(shape, rate).  ref: distribution.py:201-278"""

    shape: Union[str, Identity, LinearCombination, MixtureParameterVector]
    rate: Union[str, Identity, LinearCombination, MixtureParameterVector]

    def __post_init__(self):
        allowed = (Identity, LinearCombination, MixtureParameterVector)
        self.shape = _as_parameter(self.shape, allowed, "shape")
        self.rate = _as_parameter(self.rate, allowed, "rate")

    @property
    def _dist_params(self) -> list:
        return self.shape.get_param_list() + self.rate.get_param_list()

    def log_p(self, state: dict, by_observation: bool = False):
        """ref: distribution.py:241-261"""
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def rvs(self, state, n: int = 1):
        """ref: distribution.py:263-278"""
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)


@dataclass
class Uniform(Distribution):
    """Uniform on a hyper-rectangle.  ref: distribution.py:377-458"""

    domain_response_lower: Union[float, np.ndarray] = 0.0
    domain_response_upper: Union[float, np.ndarray] = 1.0

    def __post_init__(self):
        self.domain_response_lower = np.array(self.domain_response_lower, ndmin=2, dtype=np.float64)
        if self.domain_response_lower.shape[0] == 1:
            self.domain_response_lower = self.domain_response_lower.T
        self.domain_response_upper = np.array(self.domain_response_upper, ndmin=2, dtype=np.float64)
        if self.domain_response_upper.shape[0] == 1:
            self.domain_response_upper = self.domain_response_upper.T

    @property
    def _dist_params(self) -> list:
        return []

    def log_p(self, state: dict, by_observation: bool = False):
        """ref: distribution.py:422-442"""
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def rvs(self, state, n: int = 1):
        """ref: distribution.py:444-458"""
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)


@dataclass
class Poisson(Distribution):
    """Poisson(rate).  ref: distribution.py:461-523"""

    rate: Union[str, Identity, LinearCombination, MixtureParameterVector]

    def __post_init__(self):
        self.rate = _as_parameter(self.rate, (Identity, LinearCombination, MixtureParameterVector), "rate")

    @property
    def _dist_params(self) -> list:
        return self.rate.get_param_list()

    def log_p(self, state: dict, by_observation: bool = False):
        """ref: distribution.py:490-508"""
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def rvs(self, state: dict, n: int = 1):
        """ref: distribution.py:510-523"""
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)


@dataclass
class Categorical(Distribution):
    """Categorical allocation prior: response in {0, ..., n_cat - 1}, prob of shape (1 or p, n_cat).
    ref: distribution.py:282-374.  log_p = sum_i log prob[i, z_i] (the multinomial(n = 1) log-pmf of :318-345)."""

    prob: Union[str, Identity]

    def __post_init__(self):
        if isinstance(self.prob, str):
            self.prob = Identity(self.prob)
        if not isinstance(self.prob, Identity):
            raise TypeError("prob expected to be Identity")

    @property
    def _dist_params(self) -> list:
        return self.prob.get_param_list()

    def log_p(self, state: dict, by_observation: bool = False):
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def rvs(self, state, n: int = 1):
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)
