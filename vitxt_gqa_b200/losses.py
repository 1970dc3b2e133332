"""Registered losses `pos_bce_loss` and `InfoNCE` on the fused CUDA kernels (K7).

Drop-in for reference pythia/modules/losses.py:322-385: same registry keys, same
`forward(sample_list, model_output)` signature, scalar result.  Each loss is a
`torch.autograd.Function` over its fused forward / backward kernels, so
`loss.backward()` of the reference trainer reaches the model's backward schedule.
"""
import torch
from torch import nn

from . import lib as _lib
from .pythia_api import registry


def _ws(B, T, device):
    n = int(_lib.get_lib().loss_workspace_bytes(B, T))
    return torch.empty(n, device=device, dtype=torch.uint8)


def _dev_f32(t, device):
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, targets, mask):
        dev = scores.device
        B, T, N = scores.shape
        scores = scores.contiguous()
        out = torch.empty(1, device=dev, dtype=torch.float32)
        _lib.get_lib().pos_bce_loss(scores.data_ptr(), targets.data_ptr(), mask.data_ptr(), B, T, N,
                                    _ws(B, T, dev).data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        ctx.save_for_backward(scores, targets, mask)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        scores, targets, mask = ctx.saved_tensors
        B, T, N = scores.shape
        go = grad_out.detach().reshape(1).float().contiguous()
        d = torch.empty_like(scores)
        _lib.get_lib().pos_bce_loss_bwd(scores.data_ptr(), targets.data_ptr(), mask.data_ptr(), B, T, N, go.data_ptr(),
                                        d.data_ptr(), 0, torch.cuda.current_stream(scores.device).cuda_stream)
        return d, None, None


class _NCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, pos, neg, temperature):
        dev = ref.device
        B, T, N = ref.shape
        out = torch.empty(1, device=dev, dtype=torch.float32)
        _lib.get_lib().info_nce_loss(ref.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, T, N, float(temperature),
                                     _ws(B, T, dev).data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        ctx.save_for_backward(ref, pos, neg)
        ctx.temperature = float(temperature)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        ref, pos, neg = ctx.saved_tensors
        B, T, N = ref.shape
        dev = ref.device
        L = _lib.get_lib()
        go = grad_out.detach().reshape(1).float().contiguous()
        ws = torch.empty(int(L.loss_bwd_workspace_bytes(B, T)), device=dev, dtype=torch.uint8)
        d = [torch.empty_like(ref) for _ in range(3)]
        L.info_nce_loss_bwd(ref.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, T, N, ctx.temperature, ws.data_ptr(),
                            go.data_ptr(), d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), 0,
                            torch.cuda.current_stream(dev).cuda_stream)
        return d[0], d[1], d[2], None


@registry.register_loss("pos_bce_loss")
class POSBCEWithMaskLoss(nn.Module):
    """sum(BCEWithLogits(pos_scores, targets) * loss_mask) / max(sum(loss_mask), 1)
    (reference losses.py:329-343)."""

    def forward(self, sample_list, model_output):
        scores = model_output["pos_scores"]
        assert scores.dim() == 3 and scores.is_cuda, "pos_bce_loss runs on the CUDA scores of the B200 model"
        dev = scores.device
        B, T, N = scores.shape
        targets = _dev_f32(sample_list["targets"], dev)
        mask = _dev_f32(sample_list["train_loss_mask"], dev)
        assert mask.dim() == 2
        return _BCEFn.apply(scores, targets, mask)


@registry.register_loss("InfoNCE")
class InfoNCE(nn.Module):
    """Per-sample 2-way InfoNCE between cos(ref,pos) and cos(ref,neg) of the last-dim-normalised
    score tensors, temperature from the forward default (reference losses.py:346-385, Q19)."""

    def __init__(self, temperature=0.1, reduction="mean", negative_mode="unpaired"):
        super().__init__()
        self.temperature, self.reduction, self.negative_mode = temperature, reduction, negative_mode

    def forward(self, sample_list, model_output, temperature=0.1, reduction="mean", negative_mode="paired"):
        ref, pos, neg = (model_output[k].contiguous() for k in ("ref_scores", "pos_scores", "neg_scores"))
        assert ref.is_cuda and reduction == "mean"
        return _NCEFn.apply(ref, pos, neg, float(temperature))
