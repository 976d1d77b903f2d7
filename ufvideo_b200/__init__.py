"""ufvideo_b200 -- B200-native drop-in for UFVideo's object (region) encoder hot path.

Public surface mirrors the reference's ``ufvideo/model/layer.py``:
``build_region_encoder``, ``MaskExtractor``, ``MaskPooling``, ``token_merge``.
Importing the package does not load the CUDA library; the first operator call does, and raises
if ``libufv_b200.so`` has not been built (there is no CPU fallback).
"""
from .layer import MaskExtractor, MaskPooling, build_region_encoder, token_merge  # noqa: F401

__all__ = ["MaskExtractor", "MaskPooling", "build_region_encoder", "token_merge"]
