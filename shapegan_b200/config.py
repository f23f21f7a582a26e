"""Precision / runtime configuration of the shapegan_b200 hot path.

precision 'bf16'  : bf16 operands, fp32 accumulation (throughput mode; BASELINE.json quotes its metric in bf16)
precision 'fp32x' : every operand is fed to the tensor cores as a hi/lo bf16 split and each product as
                    hi*hi + hi*lo + lo*hi with fp32 accumulation (~2^-17 relative per product): the mode in which
                    outputs match the reference's fp32 CPU path within 1e-3 (tests/test_parity_gpu.py).
"""
import os

_PRECISION = os.environ.get('SG_B200_PRECISION', 'bf16')


def set_precision(name):
    global _PRECISION
    if name not in ('bf16', 'fp32x'):
        raise ValueError("precision must be 'bf16' or 'fp32x'")
    _PRECISION = name


def precision():
    return _PRECISION


def planes():
    return 2 if _PRECISION == 'fp32x' else 1
