"""Fused execution plans over the C-ABI kernels (CUDA only; no fallback)."""
from .resnet import PipelinedExplainer, ResNetPlan  # noqa: F401
