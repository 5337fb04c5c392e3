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
    @staticmethod
    def forward(ctx, displacement, deform_args, deform_kwargs, *xs):
        ctx.save_for_backward(displacement)
        ctx.deform_args = deform_args
        ctx.deform_kwargs = deform_kwargs
        ctx.x_shapes = [tuple(x.shape) for x in xs]

        ys = _deform_grid([x.detach() for x in xs], displacement.detach(),
                             *deform_args, **deform_kwargs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        displacement, = ctx.saved_tensors
        dxs = _deform_grid_gradient([dy.detach() for dy in dys], displacement.detach(),
                                       *ctx.deform_args, X_shape=ctx.x_shapes, **ctx.deform_kwargs)
        return (None, None, None) + tuple(dxs)


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
    if not isinstance(X, (list, tuple)):
        X_list = [X]
    else:
        X_list = X
    displacement = torch.as_tensor(displacement)
    y = ElasticDeform.apply(displacement, args, kwargs, *X_list)

    if isinstance(X, (list, tuple)):
        return y
    else:
        return y[0]
