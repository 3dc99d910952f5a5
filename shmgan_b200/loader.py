"""Dataset-loader contract of datasetLoader.py:19-170, device side.

The reference zips five `image_dataset_from_directory` streams (folders I0, I60, I90, I150, ED -- or I0, I45, I90, I135, ED --
:29-33), each `batch_size=1`, resized to (image_size, image_size) with bilinear interpolation, `/255.0` (:60) and flipped
vertically when `random_flip` is False (:61).  File decoding stays on the host (PNG/JPEG decode is control-plane work, out of
scope); this module takes the DECODED uint8 arrays, stages them through pinned memory and runs resize + scale + flip as one
kernel per folder (`shm_load_u8_bilinear`), yielding exactly the 5-tuple `train_step(orig0, orig45, orig90, orig135, origED)` takes.
"""
from __future__ import annotations

import os
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

FOLDERS_PSD = ("I0", "I60", "I90", "I150", "ED")         # datasetLoader.py:29-33 (PSD polar dataset)
FOLDERS_SHMGAN = ("I0", "I45", "I90", "I135", "ED")      # datasetLoader.py:22-26 (commented alternative)
_EXT = (".bmp", ".gif", ".jpeg", ".jpg", ".png")         # keras image_dataset_from_directory's allow-list


def list_folder(path: str) -> List[str]:
    """File order of image_dataset_from_directory(shuffle=False): sorted walk, allow-listed extensions."""
    return [os.path.join(path, f) for f in sorted(os.listdir(path)) if f.lower().endswith(_EXT)]


def decode_rgb_u8(path: str) -> np.ndarray:
    """Host-side decode to uint8 [H,W,3] RGB (color_mode='rgb').  Uses whichever decoder this image has."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            return np.asarray(im.convert("RGB"), dtype=np.uint8)
    except ImportError:
        import cv2
        return cv2.cvtColor(cv2.imread(path, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


class PolarimetricLoader:
    """Iterates 5-tuples of [B,S,S,3] fp32 device tensors in the reference's zip order.

    sources: five sequences of decoded uint8 [H,W,3] arrays (or five folder paths).  batch_size images of equal source size are
    stacked per step (the reference uses batch_size = 1, datasetLoader.py:57); `repeat` = num_epochs (:161)."""

    def __init__(self, sources: Sequence, image_size: int, batch_size: int = 1, random_flip: bool = False, repeat: int = 1,
                 est_diffuse: bool = False):
        # random_flip: the reference's map lambda `x if self.random_flip else flip_up_down(x)` (datasetLoader.py:61) is traced ONCE, when
        # datasetLoad runs, with self.random_flip == 0.0 (ShmGANwithSSpecSeg.py:203; the per-epoch redraw at :983 never reaches the traced
        # graph): the training streams are therefore ALWAYS flipped vertically.  The default here reproduces that effective behaviour
        # (random_flip=False -> flip); test.py:96,121 has the flip commented out, so inference callers pass random_flip=True (no flip).
        # est_diffuse (main.py:36 declares the flag, nothing reads it): four polarisation streams only; the fifth (ED) is the
        # pseudo-diffuse min-of-4 of the DECODED uint8 images (utils.py:68-123 works on the originals), computed on the device
        # and then sent through the same resize / scale / flip kernel as a pre-populated ED folder would be
        self.est_diffuse = est_diffuse
        assert len(sources) == (4 if est_diffuse else 5), "streams: I0, I45|I60, I90, I135|I150 (, ED unless est_diffuse)"
        self.streams = [([decode_rgb_u8(p) for p in list_folder(s)] if isinstance(s, str) else list(s)) for s in sources]
        n = {len(s) for s in self.streams}
        assert len(n) == 1, "the five folders must hold the same number of images (tf.data.Dataset.zip truncates silently)"
        self.length_dataset = n.pop()                      # datasetLoader.py:165
        self.image_size, self.batch_size, self.flip, self.repeat = image_size, batch_size, (not random_flip), repeat
        self._pinned = {}
        self._copy_stream = None

    def __len__(self) -> int:
        return self.repeat * ((self.length_dataset + self.batch_size - 1) // self.batch_size)

    def _load_batch(self, i: int, slot: int):
        """Stages batch i (five uint8 stacks) through pinned buffer set `slot` and runs the resize / scale / flip kernels on the copy
        stream; returns (tensors, event recorded when they are ready)."""
        stream = self._copy_stream
        out, raw = [], []
        with torch.cuda.stream(stream):
            for k, s in enumerate(self.streams):
                arr = np.stack(s[i:i + self.batch_size])
                pin = self._pinned.get((slot, k))
                if pin is None or pin.shape != arr.shape:
                    pin = torch.empty(arr.shape, dtype=torch.uint8, pin_memory=True)
                    self._pinned[(slot, k)] = pin
                pin.copy_(torch.from_numpy(arr))
                dev = pin.cuda(non_blocking=True)
                out.append(ops.load_u8_images(dev, self.image_size, self.flip))
                if self.est_diffuse:
                    raw.append(dev)
            if self.est_diffuse:
                out.append(ops.load_u8_images(ops.pseudo_diffuse_min4(*raw), self.image_size, self.flip))
            ev = torch.cuda.Event()
            ev.record(stream)
        return tuple(out), ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        """Double-buffered: while the consumer trains on batch i, batch i + 1 is copied and resized on a side stream (two pinned
        buffer sets; a set is rewritten only after the copies that read it have completed)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._pinned = {}
        starts = [i for _ in range(self.repeat) for i in range(0, self.length_dataset, self.batch_size)]
        pending = None
        for n, i in enumerate(starts):
            if pending is None:
                pending = self._load_batch(i, n & 1)
            batch, ev = pending
            ev.synchronize()                                 # the pinned set of this batch may now be reused two batches later
            pending = self._load_batch(starts[n + 1], (n + 1) & 1) if n + 1 < len(starts) else None
            torch.cuda.current_stream().wait_event(ev)
            for t in batch:
                t.record_stream(torch.cuda.current_stream())
            yield batch
