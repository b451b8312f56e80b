"""Parameter specifications (host-side mirror of the reference's `openmcmc.parameter`).

ref: parameter.py:26-538.  These classes are declarative: they name the state entries a distribution parameter is
built from.  The sweep-plan compiler (engine.py) pattern-matches them onto CUDA kernels; the per-call `predictor`
helpers below run on the device through libomc as well (no numpy arithmetic on the host).
"""

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Union


@dataclass
class Parameter(ABC):
    """Abstract base class for parameter.  ref: parameter.py:26-71"""

    @abstractmethod
    def predictor(self, state: dict):
        """Evaluate the parameter from `state` (host dict of arrays); computed on the device."""

    @abstractmethod
    def get_param_list(self) -> list:
        """All state entries used by the parameter."""

    @abstractmethod
    def get_grad_param_list(self) -> list:
        """State entries the gradient is defined for."""


@dataclass
class Identity(Parameter):
    """f = x.  ref: parameter.py:74-141"""

    form: str

    def predictor(self, state: dict):
        return state[self.form]

    def get_param_list(self) -> list:
        return [self.form]

    def get_grad_param_list(self) -> list:
        return [self.form]


@dataclass
class LinearCombination(Parameter):
    """f = sum_i state[form[k_i]] @ state[k_i].  ref: parameter.py:144-228"""

    form: dict

    def predictor(self, state: dict):
        return self.predictor_conditional(state)

    def predictor_conditional(self, state: dict, term_to_exclude: Union[str, list] = None):
        """ref: parameter.py:174-197; the matrix-vector products run in omc_linear_predictor."""
        from openmcmc_b200 import hostcalls

        if term_to_exclude is None:
            term_to_exclude = []
        if isinstance(term_to_exclude, str):
            term_to_exclude = [term_to_exclude]
        terms = [(prefactor, prm) for prm, prefactor in self.form.items() if prm not in term_to_exclude]
        return hostcalls.linear_predictor(state, terms)

    def get_param_list(self) -> list:
        return list(self.form.keys()) + list(self.form.values())

    def get_grad_param_list(self) -> list:
        return list(self.form.keys())


@dataclass
class LinearCombinationWithTransform(LinearCombination):
    """f = sum_i X_i @ (exp(theta_i) if transform[theta_i] else theta_i).  ref: parameter.py:231-297."""

    transform: dict = None

    def predictor_conditional(self, state: dict, term_to_exclude: Union[str, list] = None):
        """ref: parameter.py:255-281; the products (and the exp) run in omc_linear_predictor."""
        from openmcmc_b200 import hostcalls

        if term_to_exclude is None:
            term_to_exclude = []
        if isinstance(term_to_exclude, str):
            term_to_exclude = [term_to_exclude]
        terms = [(prefactor, prm, bool(self.transform[prm])) for prm, prefactor in self.form.items()
                 if prm not in term_to_exclude]
        return hostcalls.linear_predictor(state, terms)


@dataclass
class ScaledMatrix(Parameter):
    """f = scalar * matrix.  ref: parameter.py:300-373"""

    matrix: str
    scalar: str

    def predictor(self, state: dict):
        from openmcmc_b200 import hostcalls

        return hostcalls.scaled_matrix(state, self.matrix, self.scalar)

    def get_param_list(self) -> list:
        return [self.scalar, self.matrix]

    def get_grad_param_list(self) -> list:
        return [self.scalar]

    def precision_unscaled(self, state: dict, _):
        """ref: parameter.py:362-373"""
        return state[self.matrix]


@dataclass
class MixtureParameter(Parameter, ABC):
    """ref: parameter.py:376-417.  Mixture parameters are SURVEY §8 f2 ("next"); declared so models that use them fail
    at plan-compile time with a clear message rather than at import."""

    param: str
    allocation: str

    def get_param_list(self) -> list:
        return [self.param, self.allocation]

    def predictor(self, state: dict):
        """param[allocation] (parameter.py:437-446; the Matrix flavour wraps it in a sparse diagonal, :494-504): a plain
        index gather of constant-shaped host arrays, no arithmetic."""
        import numpy as np

        return np.asarray(state[self.param])[np.asarray(state[self.allocation]).astype(int).flatten()]


@dataclass
class MixtureParameterVector(MixtureParameter):
    """ref: parameter.py:420-471"""

    def get_grad_param_list(self) -> list:
        return [self.param]


@dataclass
class MixtureParameterMatrix(MixtureParameter):
    """ref: parameter.py:474-538"""

    def get_grad_param_list(self) -> list:
        return []
