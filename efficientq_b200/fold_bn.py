"""BN folding before calibration (reference src/models/fold_bn.py:14-80).

W <- W * gamma / sqrt(var + eps),  b <- beta - gamma * mean / sqrt(var + eps) (+ scaled old
bias); every folded BN is replaced by an identity, so every conv ends up with a bias.
Host-side, one-off; it defines W0 / b0 of the ADMM problem and must match exactly.
"""
import torch
import torch.nn as nn


class StraightThrough(nn.Module):
    def forward(self, x):
        return x


def _is_bn(m):
    return isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d))


def _absorbs(m):
    return isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear))


def fold_pair(conv, bn):
    std = torch.sqrt(bn.running_var + bn.eps)
    shape = (conv.out_channels,) + (1,) * (conv.weight.dim() - 1)
    if bn.affine:
        w = conv.weight.data * (bn.weight / std).view(shape)
        shift = bn.bias - bn.weight * bn.running_mean / std
        b = bn.weight * conv.bias / std + shift if conv.bias is not None else shift
    else:
        w = conv.weight.data / std.view(shape)
        shift = -bn.running_mean / std
        b = conv.bias / std + shift if conv.bias is not None else shift
    if conv.bias is None:
        conv.bias = nn.Parameter(b.detach().clone())
    else:
        conv.bias.data = b.detach()
    conv.weight.data = w.detach()


def search_fold_and_remove_bn(model):
    """Depth-first walk; a BN directly following an absorbing layer (in registration order,
    also across container boundaries) is folded into it and replaced by identity."""
    model.eval()
    prev = None
    for name, child in model.named_children():
        if _is_bn(child) and _absorbs(prev):
            fold_pair(prev, child)
            setattr(model, name, StraightThrough())
        elif _absorbs(child):
            prev = child
        else:
            prev = search_fold_and_remove_bn(child)
    return prev
