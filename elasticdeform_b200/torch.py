"""PyTorch wrapper, mirroring reference elasticdeform/torch.py (cited as ``ref:LINE``).

Same call signature and return conventions as ``elasticdeform.torch.deform_grid``
(ref:33-66) and the same autograd contract (gradient w.r.t. the inputs only,
``None`` for the displacement, ref:29) -- but CUDA tensors never leave the device:
the reference round-trips through ``.cpu().numpy()`` in both directions
(ref:13-16, ref:25-29).
"""
from __future__ import absolute_import

import torch

from .deform_grid import deform_grid as _deform_grid
from .deform_grid import deform_grid_gradient as _deform_grid_gradient


class ElasticDeform(torch.autograd.Function):
    """``ElasticDeform.apply(displacement, args, kwargs, *xs)`` -> tuple of deformed tensors (ref:5-30).

    The inputs stay where they are: CUDA tensors are handed to the C-ABI as device pointers on the
    current stream, and the outputs / input gradients are CUDA tensors.  Only the gradient with respect
    to the inputs exists (``None`` for the displacement and the two argument containers)."""

    @staticmethod
    def forward(ctx, displacement, deform_args, deform_kwargs, *xs):
        ctx.save_for_backward(displacement)
        ctx.call = (tuple(deform_args), dict(deform_kwargs))
        ctx.input_shapes = [tuple(x.shape) for x in xs]
        outputs = _deform_grid([x.detach() for x in xs], displacement.detach(), *ctx.call[0], **ctx.call[1])
        return tuple(outputs)

    @staticmethod
    def backward(ctx, *dys):
        (displacement,) = ctx.saved_tensors
        args, kwargs = ctx.call
        grads = _deform_grid_gradient([dy.detach() for dy in dys], displacement.detach(), *args,
                                      X_shape=ctx.input_shapes, **kwargs)
        return (None, None, None) + tuple(grads)


def deform_grid(X, displacement, *args, **kwargs):
    """
    Elastic deformation with a deformation grid, wrapped for PyTorch (ref:33-66).

    Parameters
    ----------
    X : torch.Tensor or list/tuple of torch.Tensors
        input image or list of input images (CUDA tensors stay on the device)
    displacement : torch.Tensor (or array-like)
        displacement vectors for each control point

    Returns
    -------
    torch.Tensor, or a tuple of tensors if a list/tuple was given.

    See ``elasticdeform_b200.deform_grid`` for the other parameters.
    """
    single = not isinstance(X, (list, tuple))
    inputs = [X] if single else list(X)
    outputs = ElasticDeform.apply(torch.as_tensor(displacement), args, kwargs, *inputs)
    return outputs[0] if single else outputs
