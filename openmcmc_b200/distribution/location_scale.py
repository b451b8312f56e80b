"""Location-scale distributions: Normal, NullDistribution (host-side mirror).  ref: distribution/location_scale.py

LogNormal: response-branch log_p / gradient / Hessian run as an MH term on the device.
"""

from abc import ABC
from dataclasses import dataclass
from typing import Union

import numpy as np

from openmcmc_b200.distribution.distribution import Distribution
from openmcmc_b200.parameter import (
    Identity,
    LinearCombination,
    MixtureParameterMatrix,
    MixtureParameterVector,
    ScaledMatrix,
)


@dataclass
class LocationScale(Distribution, ABC):
    """ref: location_scale.py:31-62"""

    mean: Union[str, Identity, LinearCombination, MixtureParameterVector]  # incl. LinearCombinationWithTransform
    precision: Union[str, Identity, ScaledMatrix, MixtureParameterMatrix]

    @property
    def _dist_params(self) -> list:
        return self.mean.get_param_list() + self.precision.get_param_list()

    def __post_init__(self):
        if isinstance(self.mean, str):
            self.mean = Identity(self.mean)
        if not isinstance(self.mean, (Identity, LinearCombination, MixtureParameterVector)):
            raise TypeError("mean expected to be one of [Identity, LinearCombination, MixtureParameterVector]")
        if isinstance(self.precision, str):
            self.precision = Identity(self.precision)
        if not isinstance(self.precision, (Identity, ScaledMatrix, MixtureParameterMatrix)):
            raise TypeError("precision expected to be one of [Identity, ScaledMatrix, MixtureParameterMatrix]")


class NullDistribution(LocationScale):
    """log_p = 0, zero gradient / Hessian.  ref: location_scale.py:65-123"""

    def log_p(self, state: dict, by_observation: bool = False) -> float:
        return 0.0

    def grad_log_p(self, state: dict, param: str, hessian_required: bool = True, method: str = "fd"):
        shape = np.shape(state[param])
        if hessian_required:
            return np.zeros(shape), np.zeros((shape[0], shape[0]))
        return np.zeros(shape)

    def rvs(self, state: dict, n: int = 1) -> None:
        return None


@dataclass
class Normal(LocationScale):
    """Multivariate normal in precision form, optionally truncated.  ref: location_scale.py:126-272"""

    domain_response_lower: np.ndarray = None
    domain_response_upper: np.ndarray = None

    def log_p(self, state: dict, by_observation: bool = False):
        """Non-truncated Gaussian log-density; -inf outside the domain.  ref: location_scale.py:145-167"""
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def grad_log_p(self, state: dict, param: str, hessian_required: bool = True, method: str = "analytic"):
        """Analytic response / linear-mean branches.  ref: location_scale.py:190-250"""
        from openmcmc_b200 import hostcalls

        return hostcalls.grad_log_p(self, state, param, hessian_required, method)

    def rvs(self, state: dict, n: int = 1):
        """ref: location_scale.py:252-272"""
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)


@dataclass
class LogNormal(LocationScale):
    """Multivariate log-normal in precision form.  ref: location_scale.py:275-418

    log_p = MVN log-pdf at log(x) minus sum(log x) (:296-303); as the response of an MH-sampled parameter the gradient is
    -(1 + Q r)/x and the Hessian diag(1/x) Q diag(1/x) - diag((1 + Q r)/x^2), r = log x - mean (:340-343, :383-399).
    """

    def log_p(self, state: dict, by_observation: bool = False):
        from openmcmc_b200 import hostcalls

        return hostcalls.log_p(self, state, by_observation)

    def grad_log_p(self, state: dict, param: str, hessian_required: bool = True, method: str = "analytic"):
        from openmcmc_b200 import hostcalls

        return hostcalls.grad_log_p(self, state, param, hessian_required, method)

    def rvs(self, state: dict, n: int = 1):
        from openmcmc_b200 import hostcalls

        return hostcalls.rvs(self, state, n)
