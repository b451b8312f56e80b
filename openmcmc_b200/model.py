"""Model: dictionary-like collection of distributions (host-side mirror of `openmcmc.model`).  ref: model.py:19-112"""

from dataclasses import dataclass


@dataclass
class Model(dict):
    """self.keys() are the distribution responses; values the Distribution objects.  ref: model.py:19-39"""

    def __init__(self, distributions: list, response: dict = None):
        dist_dict = {}
        for dist in distributions:
            dist_dict[dist.response] = dist
        super().__init__(dist_dict)
        self.response = response

    def conditional(self, param: str):
        """Sub-model of the distributions that depend on `param`.  ref: model.py:41-55"""
        return Model([dst for dst in self.values() if param in dst.param_list])

    def log_p(self, state: dict):
        """Sum of the member log-densities (each evaluated on the device).  ref: model.py:57-70"""
        log_prob = 0
        for dst in self.values():
            log_prob += dst.log_p(state)
        return log_prob

    def grad_log_p(self, state: dict, param: str, hessian_required: bool = True):
        """Sum of member gradients (and Hessians of the NEGATIVE log-pdf).  ref: model.py:72-112"""
        grad_sum = None
        hess_sum = None
        for dist in self.values():
            out = dist.grad_log_p(state, param, hessian_required=hessian_required)
            if hessian_required:
                grad_sum = out[0] if grad_sum is None else grad_sum + out[0]
                hess_sum = out[1] if hess_sum is None else hess_sum + out[1]
            else:
                grad_sum = out if grad_sum is None else grad_sum + out
        if hessian_required:
            return grad_sum, hess_sum
        return grad_sum
