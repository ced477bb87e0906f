"""`BcosUtilMixin` / `explanation_mode` -- mirror of reference bcos/common.py:23-436 (plotting excluded).

`explain` runs the forward under explanation mode, back-propagates the chosen logit with
`Tensor.backward(inputs=[x])` (our modules' autograd nodes launch the explain-dgrad kernels) and returns the dynamic
linear weights and the contribution map `(x * x.grad).sum(1)`.
"""
from __future__ import annotations

import warnings
from typing import Any, Dict

import torch
import torch.nn.functional as F
from torch import Tensor, nn

__all__ = ["BcosUtilMixin", "explanation_mode", "gradient_to_image", "gradient_to_image_batch"]


class explanation_mode:
    """Context manager / decorator putting every module with `set_explanation_mode` into explanation mode
    (bcos/common.py:347-384; the module list is cached on first entry like the reference does)."""

    def __init__(self, model: nn.Module):
        self.model = model
        self.expl_modules = None

    def find_expl_modules(self) -> None:
        self.expl_modules = [m for m in self.model.modules() if hasattr(m, "set_explanation_mode")]

    def __enter__(self):
        if self.expl_modules is None:
            self.find_expl_modules()
        for m in self.expl_modules:
            m.set_explanation_mode(True)

    def __exit__(self, exc_type, exc_val, exc_tb):
        for m in self.expl_modules:
            m.set_explanation_mode(False)

    def __call__(self, fn):
        def wrapped(*a, **k):
            with self:
                return fn(*a, **k)
        return wrapped


def gradient_to_image(image: Tensor, linear_mapping: Tensor, smooth: int = 15, alpha_percentile: float = 99.5):
    """RGBA explanation [H, W, 4] from a 6-channel image and its dynamic linear weights (bcos/common.py:387-436), computed
    by `bcosk_explanation_rgba` on the tensor's CUDA device and returned as a NUMPY array like the reference's (so
    `plt.imshow(model.explain(x)["explanation"])` works unchanged).  Device tensors / batches: `gradient_to_image_batch`."""
    return gradient_to_image_batch(image[None], linear_mapping[None], smooth, alpha_percentile)[0].cpu().numpy()


def gradient_to_image_batch(images: Tensor, linear_mappings: Tensor, smooth: int = 15,
                            alpha_percentile: float = 99.5) -> Tensor:
    """[nb, 6, H, W] inputs and dynamic linear weights -> [nb, H, W, 4] RGBA, one launch sequence for the batch."""
    from . import _lib as L
    if not images.is_cuda:
        raise L.BcoskError("gradient_to_image runs on a CUDA device (sm_100a); move the tensors there")
    L.require_device()
    nb, c, h, w = images.shape
    if c != 6 or linear_mappings.shape != images.shape:
        raise ValueError("expected [nb, 6, H, W] image and linear mapping")
    x = images.detach().float().contiguous()
    g = linear_mappings.detach().float().contiguous()
    out = torch.empty(nb, h, w, 4, dtype=torch.float32, device=images.device)
    tmp = torch.empty(2 * nb * h * w + nb, dtype=torch.float32, device=images.device)
    L.explanation_rgba(g, x, smooth, alpha_percentile, tmp, out)
    return out


def localisation_scores(attributions: Tensor, cell: int, smooth: int = 0, neg: bool = False) -> Tensor:
    """Per-target region fractions of a grid image on the device (interpretability/analyses/localisation.py:306-388):
    attributions [T, C, H, W] (e.g. x * dynamic weights per target, `ResNetPlan.explain_targets`) -> [T, regions], regions
    in the reference's column-major order; the localisation metric of the target in region t is `scores[t, t]`."""
    from . import _lib as L
    if not attributions.is_cuda:
        raise L.BcoskError("localisation_scores runs on a CUDA device (sm_100a); move the tensors there")
    L.require_device()
    if attributions.dim() != 4:
        raise ValueError("expected [T, C, H, W] attributions")
    a = attributions.detach().float().contiguous()
    nt, _, h, w = a.shape
    regions = (h // cell) * (w // cell)
    out = torch.empty(nt, regions, dtype=torch.float32, device=a.device)
    tmp = torch.empty(2 * nt * h * w + nt * regions, dtype=torch.float32, device=a.device)
    L.localisation_scores(a, smooth, cell, neg, tmp, out)
    return out


def text_localisation_target(out: Tensor, zeroshot_weight: Tensor, attn_unpool: bool, pool_cosine: float = 1,
                             norm_max_cosine: bool = False) -> Tensor:
    """The scalar that `compute_attributions` back-propagates (interpretability/analyses/text_localisation.py:80-105):
    cosine of the (per-token, with attn_unpool) image embedding with a text embedding [D, 1]; tokens are pooled by
    `mean(cos * |cos|^(p-1))` (p = pool_cosine > 1), by the arg-max token (p = 0) or by the plain mean (p = 1).
    Plain torch on the caller's device: a few hundred scalars; the encoder before it is the kernel path."""
    img_features = out / out.norm(dim=-1, keepdim=True)
    logits = img_features @ zeroshot_weight
    if attn_unpool:
        logits = logits.reshape(-1, 1)
        if pool_cosine == 0:
            num_features = logits.shape[0]
            logits = logits.reshape(-1, num_features)
            mask = torch.zeros_like(logits)
            mask[torch.arange(logits.shape[0], device=logits.device), logits.argmax(dim=1)] = 1.0
            logits = (logits * mask.detach()).reshape(1, num_features)
        if norm_max_cosine:
            logits = logits / logits.abs().detach().max(dim=0, keepdim=True)[0]
        if pool_cosine > 1:
            logits = logits * torch.pow(logits, pool_cosine - 1).abs().detach()
        logits = logits.mean(dim=0)
    if logits.dim() == 1:
        logits = logits.unsqueeze(0)
    return logits.max(1).values


class BcosUtilMixin:
    """Explanation helpers for models made of B-cos modules (bcos/common.py:23-344)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__explanation_mode_ctx = explanation_mode(self)  # noqa

    def explanation_mode(self):
        return self.__explanation_mode_ctx

    def explain(self, in_tensor: Tensor, idx=None, **grad2img_kwargs) -> Dict[str, Any]:
        """bcos/common.py:92-188 (batch size 1, like the reference)."""
        if in_tensor.ndim == 3:
            raise ValueError("Expected 4-dimensional input tensor")
        if in_tensor.shape[0] != 1:
            raise ValueError("Expected batch size of 1")
        if not in_tensor.requires_grad:
            warnings.warn("Input tensor did not require grad! Has been set automatically to True!")
            in_tensor.requires_grad = True
        if self.training:  # noqa
            warnings.warn("Model is in training mode! This might lead to unexpected results! Use model.eval()!")
        result = dict()
        with torch.enable_grad(), self.explanation_mode():
            out = self(in_tensor)  # noqa
            pred_out = out.max(1)
            result["prediction"] = pred_out.indices.item()
            if idx is None:
                to_be_explained_logit = pred_out.values
                result["explained_class_idx"] = pred_out.indices.item()
            else:
                to_be_explained_logit = out[0, idx]
                result["explained_class_idx"] = idx
            to_be_explained_logit.backward(inputs=[in_tensor])
        result["dynamic_linear_weights"] = in_tensor.grad
        result["contribution_map"] = (in_tensor * in_tensor.grad).sum(1)
        result["explanation"] = gradient_to_image(in_tensor[0].detach(), in_tensor.grad[0], **grad2img_kwargs)
        return result

    def explain_batch(self, in_tensor: Tensor) -> Dict[str, Tensor]:
        """Batched form used by the reference's ExplanationsLogger / Captum path."""
        xb = in_tensor.detach().clone().requires_grad_(True)
        with torch.enable_grad(), self.explanation_mode():
            out = self(xb)  # noqa
            out.max(1).values.sum().backward(inputs=[xb])
        return {"logits": out.detach(), "prediction": out.argmax(1), "dynamic_linear_weights": xb.grad,
                "contribution_map": (xb.detach() * xb.grad).sum(1)}

    def attribute(self, image: Tensor, target: int, **kwargs) -> Tensor:
        """Input x gradient of one target under explanation mode == Captum InputXGradient (bcos/common.py:280-317,
        interpretability/explanation_methods/explainers/captum.py:29-32)."""
        x = image.detach().clone().requires_grad_(True)
        with torch.enable_grad(), self.explanation_mode():
            out = self(x)  # noqa
            out[:, target].sum().backward(inputs=[x])
        return (x.detach() * x.grad)

    def attribute_selection(self, image: Tensor, targets, **kwargs) -> Tensor:
        """bcos/common.py:319-344."""
        return torch.cat([self.attribute(image, t) for t in targets], dim=0)
