"""Stand-ins for the slices of gpytorch the reference's kernel classes rely on.

The reference subclasses ``gpytorch.kernels.Kernel`` and uses ``gpytorch.constraints.GreaterThan``
(kernel_utils/kernels_sphere.py:1-3, :30-60).  When gpytorch is importable its own classes are used, so the kernels
plug into gpytorch / botorch models unchanged.  gpytorch is not part of this image, so otherwise the minimal
equivalents below provide the same surface: ``register_parameter / register_constraint / register_prior /
initialize``, the ``raw_<name>_constraint`` attribute convention, ``lengthscale``, ``batch_shape`` and ``__call__``
dispatching to ``forward``.  The transforms follow gpytorch: ``GreaterThan(lb).transform(r) = lb + softplus(r)``,
``Positive().transform(r) = softplus(r)``.
"""
import math

import torch

try:  # pragma: no cover - exercised only where gpytorch is installed
    import gpytorch as _gpytorch
    from gpytorch.constraints import GreaterThan, Positive
    from gpytorch.kernels import Kernel, ScaleKernel
    HAVE_GPYTORCH = True
except Exception:  # gpytorch absent: minimal equivalents
    _gpytorch = None
    HAVE_GPYTORCH = False

    def _inv_softplus(x):
        # log(exp(x) - 1), stable for large x
        return x + torch.log(-torch.expm1(-x))

    class Interval(torch.nn.Module):
        def __init__(self, lower_bound, upper_bound):
            super().__init__()
            self.lower_bound = torch.as_tensor(float(lower_bound))
            self.upper_bound = torch.as_tensor(float(upper_bound))

    class GreaterThan(Interval):
        """transform(raw) = lower_bound + softplus(raw)."""

        def __init__(self, lower_bound):
            super().__init__(lower_bound, math.inf)

        def transform(self, tensor):
            return torch.nn.functional.softplus(tensor) + self.lower_bound.to(tensor)

        def inverse_transform(self, tensor):
            return _inv_softplus(tensor - self.lower_bound.to(tensor))

    class Positive(GreaterThan):
        def __init__(self):
            super().__init__(0.0)

    class Kernel(torch.nn.Module):
        """The subset of gpytorch.kernels.Kernel the reference's kernels use."""
        has_lengthscale = False

        def __init__(self, has_lengthscale=False, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                     lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
            super().__init__()
            self._batch_shape = torch.Size(batch_shape)
            self.ard_num_dims = ard_num_dims
            self.active_dims = active_dims
            self.eps = eps
            self._priors = {}
            has_lengthscale = has_lengthscale or self.has_lengthscale
            self.has_lengthscale = has_lengthscale
            if has_lengthscale:
                num = 1 if ard_num_dims is None else ard_num_dims
                self.register_parameter('raw_lengthscale',
                                        torch.nn.Parameter(torch.zeros(*self._batch_shape, 1, num)))
                self.register_constraint('raw_lengthscale', lengthscale_constraint or Positive())
                if lengthscale_prior is not None:
                    self.register_prior('lengthscale_prior', lengthscale_prior, lambda: self.lengthscale,
                                        lambda v: self._set_lengthscale(v))

        # -- gpytorch.Module conventions ---------------------------------------------------------------
        @property
        def batch_shape(self):
            return self._batch_shape

        def register_parameter(self, name, parameter):  # gpytorch names the second argument `parameter`
            super().register_parameter(name, parameter)

        def register_constraint(self, param_name, constraint):
            self.add_module(param_name + '_constraint', constraint)

        def register_prior(self, name, prior, param_or_closure, setting_closure=None):
            self._priors[name] = (prior, param_or_closure, setting_closure)

        def initialize(self, **kwargs):
            for name, value in kwargs.items():
                param = getattr(self, name)
                value = torch.as_tensor(value).to(param)
                with torch.no_grad():
                    param.copy_(value.expand_as(param))
            return self

        @property
        def lengthscale(self):
            if not self.has_lengthscale:
                return None
            return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

        @lengthscale.setter
        def lengthscale(self, value):
            self._set_lengthscale(value)

        def _set_lengthscale(self, value):
            if not torch.is_tensor(value):
                value = torch.as_tensor(value).to(self.raw_lengthscale)
            self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(value))

        def __call__(self, x1, x2=None, diag=False, **params):
            if x2 is None:
                x2 = x1
            if x1.dim() == 1:
                x1 = x1.unsqueeze(1)
            if x2.dim() == 1:
                x2 = x2.unsqueeze(1)
            return self.forward(x1, x2, diag=diag, **params)

    class ScaleKernel(Kernel):
        """outputscale * base_kernel, outputscale = softplus(raw_outputscale) (gpytorch.kernels.ScaleKernel)."""

        def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
            super().__init__(**kwargs)
            self.base_kernel = base_kernel
            self.register_parameter('raw_outputscale', torch.nn.Parameter(torch.zeros(tuple(self._batch_shape))))
            self.register_constraint('raw_outputscale', outputscale_constraint or Positive())
            if outputscale_prior is not None:
                self.register_prior('outputscale_prior', outputscale_prior, lambda: self.outputscale,
                                    lambda v: self._set_outputscale(v))

        @property
        def outputscale(self):
            return self.raw_outputscale_constraint.transform(self.raw_outputscale)

        @outputscale.setter
        def outputscale(self, value):
            self._set_outputscale(value)

        def _set_outputscale(self, value):
            if not torch.is_tensor(value):
                value = torch.as_tensor(value).to(self.raw_outputscale)
            self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(value))

        def forward(self, x1, x2, diag=False, **params):
            k = self.base_kernel.forward(x1, x2, diag=diag, **params)
            s = self.outputscale
            return k * s.to(k).view(*s.shape, 1, 1) if s.dim() > 0 else k * s.to(k)


class GammaPrior:
    """Gamma(concentration, rate) prior with the two attributes and the ``log_prob`` the reference's model set-up uses
    (``gpytorch.priors.torch_priors.GammaPrior(2.0, 0.15)``, gabo_sphere.py:131-137).  gpytorch's own class is accepted
    wherever this one is (duck-typed on ``concentration`` / ``rate``)."""

    def __init__(self, concentration, rate):
        self.concentration = torch.as_tensor(float(concentration))
        self.rate = torch.as_tensor(float(rate))

    def log_prob(self, x):
        x = torch.as_tensor(x, dtype=torch.float64)
        c, r = self.concentration.double(), self.rate.double()
        return c * torch.log(r) + (c - 1.0) * torch.log(x) - r * x - torch.lgamma(c)

