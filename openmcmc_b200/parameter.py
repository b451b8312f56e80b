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

    def grad(self, state: dict, param: str):
        """[n_param x n_data] Jacobian of the predictor with respect to `param`.  ref: parameter.py:61-71.  These are
        structural: a prefactor matrix, an identity, a one-hot allocation pattern -- read off the state, not computed."""
        raise NotImplementedError(f"{type(self).__name__}.grad")


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

    def grad(self, state: dict, param: str):
        """ref: parameter.py:125-141: the identity for the parameter itself, zeros otherwise."""
        import numpy as np

        value = np.asarray(state[self.form])
        if value.ndim > 1 and value.shape[1] > 1:
            raise ValueError("Gradient in Identity should not be used for variables 2D and above.")
        p = value.size
        return np.eye(p) if param == self.form else np.zeros(shape=(p, p))


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

    def grad(self, state: dict, param: str):
        """ref: parameter.py:218-228: the transposed prefactor of `param`."""
        return state[self.form[param]].T


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

    def grad(self, state: dict, param: str):
        """ref: parameter.py:282-297: exp(theta) * X' for a transformed term (the scaling runs on the device)."""
        X = state[self.form[param]]
        if not self.transform[param]:
            return X.T
        from openmcmc_b200 import hostcalls

        return hostcalls.scale_columns(X, state[param]).T


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

    def grad(self, state: dict, param: str):
        """ref: parameter.py:349-360: the un-scaled matrix."""
        return state[self.matrix]

    def precision_unscaled(self, state: dict, _):
        """ref: parameter.py:362-373"""
        return state[self.matrix]


@dataclass
class MixtureParameter(Parameter, ABC):
    """ref: parameter.py:376-417.  Mean / precision of a mixture Normal, indexed by an allocation vector (SURVEY §8 f2;
    the device path: csrc/mixture.cu, engine.MixtureNormal)."""

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

    def grad(self, state: dict, param: str):
        """ref: parameter.py:448-463: the one-hot pattern [n_param x n_data] of the allocations."""
        import numpy as np

        alloc = np.asarray(state[self.allocation]).astype(int).reshape(-1)
        return (np.arange(np.asarray(state[param]).size).reshape(-1, 1) == alloc.reshape(1, -1)).astype(np.float64)


@dataclass
class MixtureParameterMatrix(MixtureParameter):
    """ref: parameter.py:474-538"""

    def get_grad_param_list(self) -> list:
        return []
