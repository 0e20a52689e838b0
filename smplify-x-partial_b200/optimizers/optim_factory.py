"""``create_optimizer`` with the reference's signature (smplifyx/optimizers/optim_factory.py:
27-65).  ``lbfgsls`` (every shipped config) and ``adam`` return a ``DeviceOptimizer``: a
description of the optimiser whose iterations run inside the CUDA fit kernel
(csrc/sfx_core.cuh: lbfgs_step / strong_wolfe / adam_step), one launch per
``FittingMonitor.run_fitting`` call.  ``lbfgs`` (torch's), ``rmsprop`` and ``sgd`` return the
torch optimiser, driven from the host through the device closure.
"""
import torch.optim as optim


class DeviceOptimizer(object):
    """Optimiser settings + parameter list; state lives on the device for the duration of a
    stage (the reference builds a new optimiser per stage too, fit_single_frame.py:561)."""

    def __init__(self, params, kind, lr, max_iter=None, beta1=0.9, beta2=0.999, eps=1e-8,
                 tolerance_grad=1e-5, tolerance_change=1e-9, history_size=100):
        self.param_list = list(params)
        self.param_groups = [{'params': self.param_list, 'lr': lr}]
        self.kind, self.lr, self.max_iter = kind, lr, max_iter
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.tolerance_grad, self.tolerance_change = tolerance_grad, tolerance_change
        self.history_size = history_size

    def zero_grad(self, set_to_none=False):
        for p in self.param_list:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    def step(self, closure):
        raise RuntimeError('a DeviceOptimizer steps inside FittingMonitor.run_fitting (one CUDA '
                           'launch per stage); it has no host-side step()')


def create_optimizer(parameters, optim_type='lbfgs', lr=1e-3, momentum=0.9, use_nesterov=True,
                     beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, weight_decay=0.0,
                     centered=False, rmsprop_alpha=0.99, maxiters=20, gtol=1e-6, ftol=1e-9,
                     **kwargs):
    """-> (optimizer, create_graph) exactly like the reference."""
    if optim_type == 'adam':
        if weight_decay != 0.0:
            raise ValueError('adam on the device has no weight decay (the reference passes 0)')
        return (DeviceOptimizer(parameters, 'adam', lr, beta1=beta1, beta2=beta2, eps=epsilon),
                False)
    elif optim_type == 'lbfgs':
        return (optim.LBFGS(parameters, lr=lr, max_iter=maxiters), False)
    elif optim_type == 'lbfgsls':
        return (DeviceOptimizer(parameters, 'lbfgsls', lr, max_iter=maxiters), False)
    elif optim_type == 'rmsprop':
        return (optim.RMSprop(parameters, lr=lr, eps=epsilon, alpha=rmsprop_alpha,
                              weight_decay=weight_decay, momentum=momentum, centered=centered),
                False)
    elif optim_type == 'sgd':
        return (optim.SGD(parameters, lr=lr, momentum=momentum, weight_decay=weight_decay,
                          nesterov=use_nesterov), False)
    else:
        raise ValueError('Optimizer {} not supported!'.format(optim_type))
