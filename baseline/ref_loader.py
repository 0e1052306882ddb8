"""Import single modules of the UNMODIFIED reference from baseline/_ref/scratchpad.

The reference cannot be imported wholesale in this image: ``scratchpad/__init__`` and the package ``__init__``
files pull in the whole engine (``triteia``, ``outlines``, server code, ...).  The hot path's own
files, however, have few dependencies.  This loader registers *empty* package objects for the reference's package
directories (so no ``__init__.py`` runs) and stubs for the handful of engine modules the path's files import for
logging / configuration only:

    scratchpad.utils                         logger, TorchMemorySaverAdapter (no-op), envs, current_platform
    scratchpad.utils.logger                  logger
    scratchpad.model_executor.custom_op      CustomOp (forward -> forward_native)

Everything else -- ``memory/pool.py``, ``model_executor/forward_info.py``, ``nn/attention/backend.py``,
``radix_attention.py``, ``flashinfer_backend.py``, ``triton_attn/*.py``, ``server/args.py``,
``nn/layers/rotary_embedding.py``, ``speculative/spec_info.py`` -- is executed from the installed files, unedited.
"""
from __future__ import annotations

import contextlib
import importlib
import logging
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PKG = os.path.join(HERE, "_ref", "scratchpad")

_PACKAGES = (
    "scratchpad", "scratchpad.memory", "scratchpad.model_executor", "scratchpad.model_executor.speculative",
    "scratchpad.nn", "scratchpad.nn.attention", "scratchpad.nn.attention.triton_attn", "scratchpad.nn.layers",
    "scratchpad.server", "scratchpad.scheduler",
)


def available() -> bool:
    return os.path.exists(os.path.join(REF_PKG, "nn", "attention", "flashinfer_backend.py"))


def _register_stubs() -> None:
    if "scratchpad" in sys.modules and getattr(sys.modules["scratchpad"], "_b200_ref_stub", False):
        return
    if not available():
        raise ImportError(f"{REF_PKG} is not installed: run `python baseline/install_ref.py` in the build container")
    for name in _PACKAGES:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_PKG, *name.split(".")[1:])]
        m._b200_ref_stub = True
        sys.modules[name] = m
    logger = logging.getLogger("scratchpad_ref")

    class TorchMemorySaverAdapter:  # utils/mem_saver.py: the disabled adapter is a no-op context
        @staticmethod
        def create(enable: bool):
            return TorchMemorySaverAdapter()

        def region(self):
            return contextlib.nullcontext()

        def pause(self):
            pass

        def resume(self):
            pass

    utils = types.ModuleType("scratchpad.utils")
    utils.__path__ = []  # a package with nothing importable below it except what is registered here
    utils.logger = logger
    utils.TorchMemorySaverAdapter = TorchMemorySaverAdapter
    utils.envs = types.SimpleNamespace()
    utils.current_platform = types.SimpleNamespace(is_cuda=lambda: True, is_rocm=lambda: False, is_cpu=lambda: False)
    sys.modules["scratchpad.utils"] = utils
    ulog = types.ModuleType("scratchpad.utils.logger")
    ulog.logger = logger
    sys.modules["scratchpad.utils.logger"] = ulog

    import torch.nn as nn

    class CustomOp(nn.Module):  # model_executor/custom_op.py:5-33 without the platform probing
        def __init__(self, *a, **k):
            super().__init__()
            self._forward_method = self.forward_native

        def forward(self, *a, **k):
            return self._forward_method(*a, **k)

    cop = types.ModuleType("scratchpad.model_executor.custom_op")
    cop.CustomOp = CustomOp
    sys.modules["scratchpad.model_executor.custom_op"] = cop


def load(name: str):
    """``load("memory.pool")`` -> the reference's ``scratchpad.memory.pool`` module."""
    _register_stubs()
    return importlib.import_module("scratchpad." + name)


def triton_kernels():
    """(decode_attention module, extend_attention module) of ``scratchpad/nn/attention/triton_attn``."""
    import torch

    if not torch.cuda.is_available():
        # extend_attention.py:7 reads the device capability at import
        torch.cuda.get_device_capability = lambda *a, **k: (8, 0)  # type: ignore[assignment]
    load("nn.attention.triton_attn.prefill_attention")
    return load("nn.attention.triton_attn.decode_attention"), load("nn.attention.triton_attn.extend_attention")


def flashinfer_backend_module():
    """The reference's ``scratchpad/nn/attention/flashinfer_backend.py`` over the installed FlashInfer."""
    from . import fi_env

    fi_env.patch_jit()
    return load("nn.attention.flashinfer_backend")
