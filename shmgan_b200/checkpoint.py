"""Checkpoint / resume with the reference's surface (ShmGANwithSSpecSeg.py:939-951, :1125-1134; test.py:164-169):

    ckpt = Checkpoint(generator=net.G, discriminator=net.D, specseg=net.SpecSeg, host=net)
                                                                 # optimizer state lives with each network's ParamStore; `specseg` makes a
                                                                 # run self-contained (the reference re-reads specsegv3_chkpt.h5, :931);
                                                                 # `host` = the step / RNG counters a bit-reproducible resume needs
    manager = CheckpointManager(ckpt, checkpoint_dir, max_to_keep=3)
    ckpt.restore(manager.latest_checkpoint).expect_partial()
    path = manager.save()

Format: one `.npz` per save, entries `<net>/<variable>` in the reference's (Keras) layouts and variable names, plus
`<net>/adam_m/<variable>`, `<net>/adam_v/<variable>` and `<net>/step` (Keras Adam `iterations`), so that a TensorFlow run of the
reference can be diffed against it (shmgan_b200/keras_names.py maps the variable names); an object with host_state() /
load_host_state() (the ShmGANwithSSpecSeg instance: step counter = Philox offsets of GaussianNoise / Dropout, the drop-bit RNG, D's call
counter, the running standardisation scale) is stored as one JSON string under `<tag>/host_state`; a `checkpoint` index file lists the kept paths, newest last (TF's CheckpointManager does the
same).  Host-side control plane: no kernels involved beyond the device<->host copies of the flat parameter buffers."""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import numpy as np
import torch


def _store(model):
    net = getattr(model, "net", model)
    return net.store


class _Status:
    def __init__(self, missing: List[str]):
        self.missing = missing

    def expect_partial(self):
        return self

    def assert_consumed(self):
        assert not self.missing, "entries missing from the checkpoint: %s" % self.missing[:5]
        return self


class Checkpoint:
    def __init__(self, **models):
        self.models = models

    def state(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        for tag, model in self.models.items():
            if hasattr(model, "host_state"):
                out["%s/host_state" % tag] = np.asarray(json.dumps(model.host_state()))
                continue
            st = _store(model)
            for k, t in st.export().items():
                out["%s/%s" % (tag, k)] = t.numpy()
            if st.m is not None:
                m, v = st.m.cpu(), st.v.cpu()
                for k, (o, n, shape) in st.offsets.items():
                    if o + n <= st.n_train and k in st.gviews:
                        out["%s/adam_m/%s" % (tag, k)] = m[o:o + n].view(shape).numpy()
                        out["%s/adam_v/%s" % (tag, k)] = v[o:o + n].view(shape).numpy()
            out["%s/step" % tag] = np.asarray(st.step, dtype=np.int64)
        return out

    def write(self, path: str) -> str:
        if not path.endswith(".npz"):
            path += ".npz"
        tmp = path + ".tmp.npz"
        np.savez(tmp, **self.state())
        os.replace(tmp, path)
        return path

    def restore(self, path: Optional[str]) -> _Status:
        """`path` None (no checkpoint yet) is a no-op, like tf.train.Checkpoint.restore(None)."""
        missing: List[str] = []
        if path is None:
            return _Status(missing)
        with np.load(path) as z:
            for tag, model in self.models.items():
                if hasattr(model, "load_host_state"):
                    key = "%s/host_state" % tag
                    if key in z.files:
                        model.load_host_state(json.loads(str(z[key])))
                    else:
                        missing.append(key)
                    continue
                st = _store(model)
                named = {}
                for k in st.offsets:
                    key = "%s/%s" % (tag, k)
                    if key in z.files:
                        named[k] = torch.from_numpy(z[key])
                    else:
                        missing.append(key)
                st.load(named)
                if hasattr(model, "loaded") and len(named) == len(st.offsets):
                    model.loaded = True                      # SpecSeg: trained weights are in place (model.py::_require_mask_weights)
                if st.m is not None:
                    m, v = st.m.cpu(), st.v.cpu()
                    for k, (o, n, shape) in st.offsets.items():
                        km, kv = "%s/adam_m/%s" % (tag, k), "%s/adam_v/%s" % (tag, k)
                        if km in z.files and kv in z.files:
                            m[o:o + n] = torch.from_numpy(z[km]).reshape(-1)
                            v[o:o + n] = torch.from_numpy(z[kv]).reshape(-1)
                        elif k in st.gviews:
                            missing.append(km)
                    st.m.copy_(m)
                    st.v.copy_(v)
                if "%s/step" % tag in z.files:
                    st.step = int(z["%s/step" % tag])
        return _Status(missing)


class CheckpointManager:
    """tf.train.CheckpointManager(ckpt, directory, max_to_keep) (:945): numbered saves, an index file, pruning of old saves."""

    def __init__(self, checkpoint: Checkpoint, directory: str, max_to_keep: int = 3, checkpoint_name: str = "ckpt"):
        self.checkpoint, self.directory, self.max_to_keep, self.name = checkpoint, directory, max_to_keep, checkpoint_name
        os.makedirs(directory, exist_ok=True)
        self._index = os.path.join(directory, "checkpoint")
        self.checkpoints: List[str] = []
        if os.path.exists(self._index):
            with open(self._index) as f:
                self.checkpoints = [os.path.join(directory, l.strip()) for l in f if l.strip()]
            self.checkpoints = [p for p in self.checkpoints if os.path.exists(p)]

    @property
    def latest_checkpoint(self) -> Optional[str]:
        return self.checkpoints[-1] if self.checkpoints else None

    def save(self) -> str:
        nxt = 1
        if self.checkpoints:
            nxt = int(os.path.basename(self.checkpoints[-1]).rsplit("-", 1)[1].split(".")[0]) + 1
        path = self.checkpoint.write(os.path.join(self.directory, "%s-%d" % (self.name, nxt)))
        self.checkpoints.append(path)
        while self.max_to_keep and len(self.checkpoints) > self.max_to_keep:
            old = self.checkpoints.pop(0)
            if os.path.exists(old):
                os.remove(old)
        with open(self._index, "w") as f:
            f.write("\n".join(os.path.basename(p) for p in self.checkpoints) + "\n")
        return path
