"""shmgan_b200 -- B200 (sm_100a) kernels behind SHMGAN's hot path, plus the host-side mirror of the reference's
Python surface (ShmGANwithSSpecSeg.py: build_generator / build_discriminator / train_step; SpecSeg.py: SpecSeg).

The compute lives in ``libshmgan.so`` (C ABI in ``include/shmgan.h``, sources in ``shmgan_b200/csrc``); this package only
orchestrates launches with torch CUDA tensors as buffer carriers.  There is no CPU fallback: importing the ops without
the built library raises.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
