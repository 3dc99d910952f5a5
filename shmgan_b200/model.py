"""Host-side mirror of the reference's Python surface for the hot path (class ShmGANwithSSpecSeg, ShmGANwithSSpecSeg.py:96-875;
inference body test.py:218-297; SpecSeg.py:27-98).  Same names, argument meaning and published attributes, so main.py /
test.py style callers drop in; every number is produced by libshmgan kernels (no torch math, no CPU fallback).

What differs from the reference, on purpose:
  * tensors are torch CUDA tensors (NHWC fp32 in [0,1]) instead of tf tensors;
  * the batch may be > 1: one train_step call batches the 5 cyclic generator passes (5B images) and the 12 discriminator passes
    (2B with noise/dropout + 10B without) into a handful of large launches; per-image statistics (Q5), batch-mean losses (Q6);
  * the SpecSeg mask is a LIVE input of the attention branch (Q1); `live_mask=False` reproduces the as-written graph;
  * random draws are explicit: `drop_bits` (5 Bernoulli(0.5) bits, :509-521), `TARGET_LABELS` (:986), noise / dropout seeds;
  * precision: dtype "fp32" = exact-fp32 SIMT kernels (parity mode), "bf16" = tcgen05 tensor-core kernels, fp32 master weights.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch

from . import losses as LS
from . import _lib as L
from . import nets, ops
from .ops import _p, _stream, call

__all__ = ["ShmGANwithSSpecSeg", "SpecSeg", "default_args"]


def default_args(**kw):
    """The argparse defaults of main.py:30-70 (only the flags the hot path reads), overridable by keyword."""
    a = dict(c_dim=5, image_size=128, batch_size=1, num_epochs=200, num_iteration_decay=100000, g_lr=2e-5, d_lr=2e-5,
             n_critic=5, beta1=0.5, beta2=0.99, d_repeat_num=6, mode="train", data_dir="", model_save_dir="./models",
             checkpoint_save_dir="./checkpoints", result_dir="./results", log_dir="./logs/train", log_step=1,
             checkpoint_save_step=10, filter_size=64)
    a.update(kw)
    return SimpleNamespace(**a)


def _f32(t: torch.Tensor) -> torch.Tensor:
    return ops.cast(t.contiguous(), torch.float32)


class _GeneratorModel:
    """What build_generator returns: callable like the Keras model, G(x[B,S,S,10] fp32, training) -> [B,S,S,1] fp32.
    `mask` ([B or 1,S,S,1], SpecSeg probabilities) feeds the attention branch (live-mask mode)."""

    def __init__(self, net: nets.Generator):
        self.net = net
        self.trainable = True

    def __call__(self, x, training=False, mask=None):
        n = self.net
        attn = None
        if mask is not None and n.live_mask:
            attn, _ = n.attention(ops.cast(mask.contiguous(), n.dtype))
        y = n.forward(ops.cast(x.contiguous(), n.dtype), attn)
        return _f32(y)

    @property
    def trainable_variables(self):
        return [v for k, v in self.net.store.views.items() if k in self.net.store.gviews]


class _DiscriminatorModel:
    """What build_discriminator returns: D(x[B,S,S,3] fp32, training) -> (real/fake map [B,S/32,S/32,1], logits [B,5]).
    training=True draws GaussianNoise(0.1) and Dropout(0.2) (:352, :363) from the counter-based RNG with `seed`."""

    def __init__(self, net: nets.Discriminator):
        self.net = net
        self.trainable = True
        self.calls = 0

    def __call__(self, x, training=False, mask=None, seed=0):
        n = self.net
        attn = None
        if mask is not None and n.live_mask:
            attn, _ = n.attention(ops.cast(mask.contiguous(), n.dtype))
        xd = ops.cast(x.contiguous(), n.dtype)
        noise = keep = None
        if training:
            B, S = x.shape[0], x.shape[1]
            noise = ops.rng_normal(tuple(x.shape), seed, 2 * self.calls << 32, 0.1, n.dtype)
            keep = ops.rng_keep((B, S // 32, S // 32, n.blocks[-1].conv.cout), seed, (2 * self.calls + 1) << 32, 1.0 - n.dropout, n.dtype)
            self.calls += 1
        rf, cls = n.forward(xd, attn, noise, keep)
        return _f32(rf), cls

    @property
    def trainable_variables(self):
        return [v for k, v in self.net.store.views.items() if k in self.net.store.gviews]


class SpecSeg:
    """SpecSeg(H, W, C) of SpecSeg.py:27; `.predict(x[B,S,S,1], verbose=0)` -> sigmoid probabilities [B,S,S,1] fp32.
    The reference loads specsegv3_chkpt.h5 (ShmGANwithSSpecSeg.py:930-931, test.py:155-156; absent from the checkout).  Until one of
    the `load*` methods has run the weights are the seeded reference INITIALISERS (`loaded` is False) and the mask is that of an
    untrained network: ShmGANwithSSpecSeg refuses to train / infer on it unless built with allow_random_specseg=True."""

    def __init__(self, H, W, C=1, dtype=torch.float32, seed=44, tensor_core=True):
        assert C == 1, "SpecSeg takes a single-channel (Y) image"
        assert H % 16 == 0 and W % 16 == 0, "SpecSeg needs sides divisible by 16 (4 pooling levels)"
        self.net = nets.SpecSegNet(dtype, seed=seed, tensor_core=tensor_core)
        self.loaded = False

    def load(self, named):
        """Keras-layout tensors keyed by this repo's names (shmgan_b200/keras_names.py maps them to the Keras variable names)."""
        self.net.store.load(named)
        self.loaded = True

    def load_keras_weights(self, weights):
        """`weights` = `keras_model.get_weights()` of SpecSeg.py:27-98 (a list of arrays in layer-creation order: kernel, bias per
        conv; gamma, beta, moving_mean, moving_variance per BatchNormalization)."""
        from .keras_names import specseg_keras_names
        names = list(specseg_keras_names())
        if len(weights) != len(names):
            raise ValueError("SpecSeg has %d weight arrays, got %d" % (len(names), len(weights)))
        named = {}
        for k, wgt in zip(names, weights):
            t = torch.as_tensor(wgt)
            want = tuple(self.net.store.offsets[k][2])
            if tuple(t.shape) != want:
                raise ValueError("SpecSeg weight %s: shape %s, expected %s" % (k, tuple(t.shape), want))
            named[k] = t
        self.load(named)

    def load_npz(self, path):
        """An .npz keyed either by this repo's names (c1a.w, bn1.gamma ...) or by the Keras variable names (conv2d/kernel:0 ...)."""
        import numpy as np
        from .keras_names import specseg_keras_names
        k2r = {v: k for k, v in specseg_keras_names().items()}
        named = {}
        with np.load(path) as z:
            for key in z.files:
                name = key if key in self.net.store.offsets else k2r.get(key, k2r.get(key + ":0"))
                if name is None:
                    raise KeyError("SpecSeg.load_npz: unknown entry %r" % key)
                named[name] = torch.from_numpy(z[key])
        missing = [k for k in self.net.store.offsets if k not in named]
        if missing:
            raise KeyError("SpecSeg.load_npz: missing %s" % missing[:4])
        self.load(named)

    def load_h5(self, path):
        """specsegv3_chkpt.h5 (a Keras full-model HDF5, ShmGANwithSSpecSeg.py:931).  Needs h5py, which this image does not ship:
        convert on a machine that has it (`np.savez(out, **{w.name: w.numpy() for w in model.weights})`) and use load_npz."""
        try:
            import h5py
        except ImportError as ex:
            raise RuntimeError("h5py is not installed: export the Keras weights to .npz and use SpecSeg.load_npz") from ex
        from .keras_names import specseg_keras_names
        named = {}
        with h5py.File(path, "r") as f:
            root = f["model_weights"] if "model_weights" in f else f
            flat = {}
            root.visititems(lambda n, o: flat.__setitem__(n, o) if hasattr(o, "shape") else None)
            for name, kname in specseg_keras_names().items():
                hits = [n for n in flat if n.endswith(kname) or n.endswith(kname.split(":")[0])]
                if len(hits) != 1:
                    raise KeyError("SpecSeg.load_h5: %d datasets match %s" % (len(hits), kname))
                named[name] = torch.from_numpy(flat[hits[0]][()])
        self.load(named)

    def predict(self, x, verbose=0):
        y = self.net.predict(ops.cast(x.contiguous(), self.net.dtype))
        return _f32(y)


# loss attributes published by train_step (ShmGANwithSSpecSeg.py:669-844); they are read back from the device on first access
_LOSS_ATTRS = frozenset((
    "D1_RealFake_loss", "D3_RealFake_cyc", "D1_classification_loss", "D3_classification_loss", "D2_RealFake_target", "D4_RealFake_cyc",
    "D4_classification_loss", "G_gan_loss", "G_clsf_loss", "L1_loss_Gen", "ssim_cyc_loss", "Spec_loss", "content_loss", "style_loss",
    "total_NST_loss", "total_Generator_loss", "total_Discriminator_loss", "total_Classification_loss", "loss_values"))


class ShmGANwithSSpecSeg:
    """ShmGANwithSSpecSeg(args) (ShmGANwithSSpecSeg.py:96).  Extra keyword arguments select the B200 execution mode."""

    def __init__(self, args, dtype: str = "fp32", live_mask: bool = True, tensor_core: bool = True, seed: int = 25,
                 process_group=None, device: Optional[int] = None, allow_random_specseg: bool = False):
        self.c_dim = 5                                      # :192 (args.c_dim is overridden)
        self.image_size = args.image_size
        self.batch_size = args.batch_size
        self.num_epochs = getattr(args, "num_epochs", 1)
        self.num_iteration_decay = getattr(args, "num_iteration_decay", 100000)
        self.g_lr, self.d_lr = args.g_lr, getattr(args, "d_lr", args.g_lr)   # d_lr is ignored by the reference too (:169-174)
        self.n_critic = getattr(args, "n_critic", 5)
        self.beta1, self.beta2 = args.beta1, args.beta2
        self.d_repeat_num = getattr(args, "d_repeat_num", 6)
        self.mode = getattr(args, "mode", "train")
        for k in ("data_dir", "model_save_dir", "checkpoint_save_dir", "result_dir", "log_dir", "log_step", "checkpoint_save_step"):
            setattr(self, k, getattr(args, k, None))
        self.filter_size = args.filter_size
        self.seed = seed
        self.randomness = 0.50                              # :159
        self.dropout_amnt = 0.2                             # :160
        self.TARGET_LABELS = 0.90                           # :161; the train loop redraws it per step (:986)
        self.use_lsgan = True
        self.gradmapD, self.gradmapG = {}, {}
        self.epoch = 0
        # datasetLoader.py:42-44.  stddev_arr is the only one the reference reads back (:548, test.py:246): a device-side running mean
        # (ops.RunningMean, created by build()); mean_arr / variance_arr are write-only there and not kept
        self.stddev_arr, self.mean_arr, self.variance_arr = None, [], []
        self.allow_random_specseg = allow_random_specseg
        assert dtype in ("fp32", "bf16")
        assert self.image_size % 32 == 0, "image_size must be a multiple of 32 (5 stride-2 discriminator blocks)"
        self.dtype = torch.float32 if dtype == "fp32" else torch.bfloat16
        self.live_mask = live_mask
        self.tensor_core = tensor_core and dtype == "bf16"
        self.pg = process_group
        if device is not None:
            torch.cuda.set_device(device)
        self.specular_candidate = ops.zeros((1, self.image_size, self.image_size, 1), torch.float32)   # :206
        self.G = self.D = self.SpecSeg = None
        self._rng = random.Random(seed)
        self.drop_bits: Optional[Sequence[bool]] = None     # set to pin the 5 Bernoulli draws of the next step
        self.d_noise = self.d_keep = None                   # set to pin the GaussianNoise / Dropout draws ([2B,...] tensors)
        self.noise_seed = seed
        self.step_count = 0
        self.table = LS.LossTable()
        self._reducer = None
        self.dp_overlap = True                              # all-reduce G's gradient ranges while the last backward sweep still runs
        self._pending_losses = None                         # (pinned host copy of the loss table, event) of the last train_step
        self._loss_host = [None, None]
        # the discriminator's weight-gradient sweep (12 passes of small, bandwidth-heavy layers) is independent of the generator-loss
        # sweep (D dgrad-only -> G backward): it runs on a side stream so that its bandwidth kernels fill the gaps of the
        # tensor-bound generator kernels; likewise D(xA) forward runs beside the cyclic generator forward
        self.overlap = True
        self._side = None
        # CUDA-graph replay of train_step (bf16 / fp32 alike, with or without the data-parallel all-reduce): the ~600 launches of a step are
        # captured once per (drop bits, batch) and replayed with one cudaGraphLaunch; everything that moves from step to step (Philox offsets, Adam's lr_t, TARGET_LABELS)
        # lives in a small device block the kernels read when they run.  Off by default; `net.cuda_graph = True` turns it on.
        self.cuda_graph = False
        self.graph_warmup = 2                               # eager steps before the first capture (lazy allocations, function attributes)
        self._graphs, self._graph_pool, self._g_in, self._gp = {}, None, None, None
        self._graph_eager_steps = 0

    # -- model builders (same names as the reference) --------------------------------------------------------------
    def build_generator(self):
        """build_generator (:228-327) with the live mask-attention branch (:404-412)."""
        net = nets.Generator(self.filter_size, self.live_mask, self.dtype, seed=42, tensor_core=self.tensor_core)
        return _GeneratorModel(net)

    def build_discriminator(self):
        """build_discriminator (:343-380)."""
        net = nets.Discriminator(self.image_size, self.filter_size, self.live_mask, self.dtype, seed=43,
                                 tensor_core=self.tensor_core, dropout=self.dropout_amnt)
        return _DiscriminatorModel(net)

    def build(self):
        """What train() does before its loop (:911-931): G, D and the SpecSeg mask network."""
        if self.G is None:
            self.G = self.build_generator()
        if self.D is None:
            self.D = self.build_discriminator()
        if self.SpecSeg is None:
            self.SpecSeg = SpecSeg(self.image_size, self.image_size, 1, self.dtype, tensor_core=self.tensor_core)
        if self.stddev_arr is None:
            self.stddev_arr = ops.RunningMean()
        return self

    def _require_mask_weights(self):
        """The reference cannot run without specsegv3_chkpt.h5 (:930-931, test.py:155-156): the mask feeds both attention branches and
        Spec_loss.  Seeded random SpecSeg weights are for benchmarks / parity tests only and must be asked for."""
        if self.live_mask and not self.SpecSeg.loaded and not self.allow_random_specseg:
            raise RuntimeError("SpecSeg holds its random initialisers, not trained weights: call net.SpecSeg.load / load_npz / "
                               "load_keras_weights (specsegv3_chkpt.h5 of the reference), restore a checkpoint that contains SpecSeg, "
                               "or construct ShmGANwithSSpecSeg(..., allow_random_specseg=True)")

    def __getattr__(self, name):
        # only reached when normal lookup fails: the loss scalars of the last step are materialised on first access
        if name in _LOSS_ATTRS and self.__dict__.get("_pending_losses") is not None:
            self._finish_losses()
            return self.__dict__[name]
        if name == "gen_rgb_output" and self.__dict__.get("gen_rgb") is not None:
            # :550 / test.py:249: yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255) = gen_rgb * mean(stddev_arr) * 255 (yuv_to_rgb is linear)
            out = self.stddev_arr.scaled(self.gen_rgb.contiguous(), 255.0)
            self.__dict__["gen_rgb_output"] = out
            return out
        raise AttributeError(name)

    # -- preprocessing ------------------------------------------------------------------------------------------------
    def custom_per_image_standardization(self, image):
        """:1271-1309 on a YUV tensor is served by `yuv_standardize` on the RGB tensor (rgb->yuv and the divide are fused);
        this entry keeps the reference name for callers that hold RGB (test.py:218)."""
        self.build()
        yuv, scale = ops.yuv_standardize(image.contiguous())
        self.stddev_arr.append(scale)
        return yuv

    def calculate_estimate_diffuse(self, i0, i45, i90, i135):
        """utils.py:68-123: pseudo-diffuse = per-pixel, per-channel min over the four polarisation images."""
        return ops.pseudo_diffuse_min4(i0.contiguous(), i45.contiguous(), i90.contiguous(), i135.contiguous())

    def calcDOP(self, I0_Ych, I45_Ych, I90_Ych, I135_Ych):
        """:1157-1169: degree of polarisation from the four Y planes (Stokes S0, S1, S2; divide_no_nan)."""
        return ops.dop(_f32(I0_Ych), _f32(I45_Ych), _f32(I90_Ych), _f32(I135_Ych))

    def _y_plane(self, yuv, dtype):
        """Y channel [B,S,S,1] of a yuv tensor, converted to `dtype` (:486-490)."""
        B, S = yuv.shape[0], yuv.shape[1]
        out = ops.new((B, S, S, 1), dtype)
        ops.cast_into(yuv[..., 0:1], out)
        return out

    # -- the hot path ---------------------------------------------------------------------------------------------------
    def train_step(self, orig0, orig45, orig90, orig135, origED):
        """train_step (:467-875).  Five [B,S,S,3] fp32 CUDA tensors in [0,1]; returns None and publishes the reference's
        attributes (gen_Y, gen_rgb, cyc_gen*_rgb, specular_candidate, the loss scalars ...)."""
        self.build()
        self._require_mask_weights()
        for k in _LOSS_ATTRS:                               # the previous step's scalars (read or not) are stale now
            self.__dict__.pop(k, None)
        self.__dict__.pop("gen_rgb_output", None)
        origs = [t.contiguous() for t in (orig0, orig45, orig90, orig135, origED)]
        B, S = origs[0].shape[0], origs[0].shape[1]
        assert S == self.image_size and all(tuple(t.shape) == (B, S, S, 3) and t.dtype == torch.float32 for t in origs)
        bits = list(self.drop_bits) if self.drop_bits is not None else [self._rng.random() < self.randomness for _ in range(5)]
        if self.cuda_graph and self.d_noise is None and ops.PROF is None:
            self._train_step_graph(origs, bits)
        else:
            self._train_step_body(origs, bits, None)
            self._start_loss_readback()
        return None

    # -- CUDA-graph mode -------------------------------------------------------------------------------------------------
    class _GraphParams:
        """Device block of the per-step scalars: f = [lr_t(D), lr_t(G), T, labels(0, 0, 0, 0, T)] fp32, i = [noise offset, keep offset] int64."""

        def __init__(self):
            self.f = ops.zeros((8,), torch.float32)
            self.i = ops.zeros((2,), torch.int64)
            self.lr_D, self.lr_G, self.T, self.labels = self.f[0:1], self.f[1:2], self.f[2:3], self.f[3:8]
            self.off_noise, self.off_keep = self.i[0:1], self.i[1:2]

    _PUBLISHED = ("specular_candidate", "gen_input", "gen_Y", "gen_rgb", "averageCbCr", "ds_yuv", "cyc_Y", "cyc_gen0_rgb", "cyc_gen45_rgb",
                  "cyc_gen90_rgb", "cyc_gen135_rgb", "cyc_genED_rgb", "RealFake_gen_D1", "label_gen_D1", "RealFake_target_D2",
                  "label_target_D2", "RealFake_cyc_D3", "label_cyc_D3", "RealFake_orig_D4", "label_orig0_D4", "label_orig45_D4",
                  "label_orig90_D4", "label_orig135_D4", "label_origED_D4", "ssim_values")

    def _write_graph_params(self):
        """This step's scalars, computed on the host exactly as the eager path computes them, copied into the device block (a pageable
        source: cudaMemcpyAsync stages it at call time, so the host may run ahead of the GPU)."""
        G, D, T = self.G.net.store, self.D.net.store, float(self.TARGET_LABELS)
        f = torch.tensor([D.lr_t(self.g_lr, self.beta1, self.beta2), G.lr_t(self.g_lr, self.beta1, self.beta2), T, 0.0, 0.0, 0.0, 0.0, T],
                         dtype=torch.float32)
        i = torch.tensor([(4 * self.step_count) << 32, (4 * self.step_count + 2) << 32], dtype=torch.int64)
        self._gp.f.copy_(f, non_blocking=True)
        self._gp.i.copy_(i, non_blocking=True)

    def _train_step_graph(self, origs, bits):
        B = origs[0].shape[0]
        if self._gp is None:
            self._gp = ShmGANwithSSpecSeg._GraphParams()
        if self._g_in is None or tuple(self._g_in[0].shape) != tuple(origs[0].shape):
            self._g_in = [ops.new(tuple(o.shape), torch.float32) for o in origs]
            self._graphs.clear()                            # captured against the old input buffers
        for dst, src in zip(self._g_in, origs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)           # cudaMemcpyAsync, device to device
        self._write_graph_params()
        in_graph_opt = self._reducer is None                # data parallel: all-reduce + clip + Adam follow the replay as ordinary launches
        # everything the recorded launch sequence or its baked-in scalars depend on
        key = (tuple(bool(b) for b in bits), B, in_graph_opt, self.overlap, self.noise_seed, self.dropout_amnt, self.beta1, self.beta2)
        entry = self._graphs.get(key)
        if entry is None and self._graph_eager_steps < max(2, self.graph_warmup):    # (layers join the batched weight refresh in step 2)
            self._graph_eager_steps += 1                    # same code path as the capture, executed eagerly
            self._train_step_body(self._g_in, bits, self._gp)
            self._start_loss_readback()
            return
        G, D = self.G.net.store, self.D.net.store
        if entry is None:
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launches()
            saved = (self.step_count, [(st.step, st.version, getattr(st, "_finalized", False)) for st in (D, G)])
            try:
                # thread_local: other threads (NCCL's watchdog) may keep calling the CUDA API while this thread captures
                with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                    self._train_step_body(self._g_in, bits, self._gp, with_opt=in_graph_opt)   # records the launches; the host bookkeeping runs here
            except Exception as ex:                         # e.g. a collective that cannot be captured: stay eager from here on
                import warnings
                warnings.warn("train_step: CUDA-graph capture failed (%s: %s); continuing eagerly" % (type(ex).__name__, ex))
                self.cuda_graph = False
                self.step_count = saved[0]
                for st, (a, b, c) in zip((D, G), saved[1]):
                    st.step, st.version, st._finalized = a, b, c
                if self._reducer is not None:
                    self._reducer.pending = []
                torch.cuda.synchronize()
                self._train_step_body(origs, bits, None)
                self._start_loss_readback()
                return
            entry = (g, {k: self.__dict__[k] for k in self._PUBLISHED if k in self.__dict__}, L.launches() - n0)
            self._graphs[key] = entry
            g.replay()
        else:
            g, pub, nk = entry
            g.replay()
            L.count_replayed(nk)
            # the bookkeeping _train_step_body does on the host
            self.__dict__.update(pub)
            self.last_drop_bits = list(bits)
            for st in (D, G):
                if in_graph_opt:
                    st.step += 1
                    st.version += 1
                st._finalized = True
            self.step_count += 1
        if not in_graph_opt:
            self._reduce_and_step()
        self._start_loss_readback()

    def release_graphs(self):
        """Drops the captured graphs, their memory pool and the static input buffers (the next train_step captures afresh)."""
        self._graphs.clear()
        self._graph_pool = self._g_in = None
        self._graph_eager_steps = 0

    def _reduce_and_step(self):
        """Data-parallel tail of a replayed step: all-reduce of both flat gradient buffers (NCCL's stream), then clip + Adam on the average.
        (Measured at 8 GPUs, profiles/r02_negative_results.txt #5: reducing after the sweep costs < 0.2 ms against the overlapped form.)"""
        G, D, r = self.G.net.store, self.D.net.store, self._reducer
        r.reduce_async(D.grad)
        r.reduce_async(G.grad)
        r.wait()
        gscale = 1.0 / r.world
        D.adam_step(self.g_lr, self.beta1, self.beta2, 1e-7, 1.0, gscale)
        G.adam_step(self.g_lr, self.beta1, self.beta2, 1e-7, 1.0, gscale)

    def _train_step_body(self, origs, bits, gp, with_opt=True):
        """One step on the current stream.  gp: the device block of per-step scalars (graph mode) or None (host scalars in the launches).
        with_opt = False stops after the backward sweeps: no gradient all-reduce, no clip + Adam (`_reduce_and_step` does them) -- what a
        data-parallel step records into its CUDA graph, so that no NCCL kernel ever sits inside a graph."""
        reducer = self._reducer if with_opt else None
        ops.arena_begin(whole=gp is not None)
        G, D = self.G.net, self.D.net
        G.store.refresh_tc_all()                            # bf16 weight copies of every layer, one launch per network
        D.store.refresh_tc_all()
        B, S = origs[0].shape[0], origs[0].shape[1]
        dt, f32 = self.dtype, torch.float32
        npix = B * S * S
        T = float(self.TARGET_LABELS) if gp is None else gp.T
        self.last_drop_bits = bits
        tab = self.table
        tab.zero()

        # ---- preprocessing (:480-506) and the mask (:492, outside the tape)
        scales = ops.new((5, B), f32)
        ds = [ops.yuv_standardize(o, scales[k])[0] for k, o in enumerate(origs)]
        self.stddev_arr.append(scales)                      # :1306 (five appends per step in the reference)
        avg = ops.avg_cbcr(ds)
        mask = self.SpecSeg.net.predict(self._y_plane(ds[2], dt))
        self.specular_candidate = _f32(mask)
        g_attn = g_attn_saved = d_attn = d_attn_saved = None
        if self.live_mask:
            g_attn, g_attn_saved = G.attention(mask)
            d_attn, d_attn_saved = D.attention(mask)

        # ---- G(1) (:509-553)
        # input layout per buffer (10 plain, 16 thin first layer, 64 zero-padded first layer): whatever forward() wants for THAT batch
        g_cin, g_cin5 = G.in_channels(B, S, S), G.in_channels(5 * B, S, S)
        gen_in = ops.new((B, S, S, g_cin), dt)
        ops.assemble_input([None if bits[k] else ds[k] for k in range(5)], [3] * 5, 4, gen_in)
        gen_Y_lp, tape1 = G.forward(gen_in, g_attn, save=True)
        gen_Y = _f32(gen_Y_lp)
        d_cin = D.in_channels(2 * B, S, S)                  # the discriminator takes the plain 3-channel image in every mode
        xA = ops.new((2 * B, S, S, d_cin), dt)              # D batch A = [gen_rgb | origED], training=True (:559-563)
        xB = ops.new((10 * B, S, S, d_cin), dt)             # D batch B = [5 cyc_rgb | 5 orig], training=False (:627-642)
        to_d = lambda src, out: ops.cast_into(src, out)
        if dt == f32:
            gen_rgb = xA[:B]
            ops.yuv2rgb(gen_Y, avg, gen_rgb, None)
        else:
            gen_rgb = ops.new((B, S, S, 3), f32)
            ops.yuv2rgb(gen_Y, avg, gen_rgb, xA[:B])
        to_d(origs[4], out=xA[B:])

        # ---- D on [gen_rgb | origED] (training=True, :559-563) on the side stream, beside the cyclic generator pass
        s32 = S // 32
        if self.d_noise is not None:
            noise, keep = ops.cast(self.d_noise.contiguous(), dt), ops.cast(self.d_keep.contiguous(), dt)
        else:
            noise = ops.rng_normal((2 * B, S, S, 3), self.noise_seed, (4 * self.step_count) << 32 if gp is None else gp.off_noise, 0.1, dt)
            keep = ops.rng_keep((2 * B, s32, s32, D.blocks[-1].conv.cout), self.noise_seed,
                                (4 * self.step_count + 2) << 32 if gp is None else gp.off_keep, 1.0 - self.dropout_amnt, dt)
        side = self._side_stream()
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                rfA_lp, clsA, tapeA = D.forward(xA, d_attn, noise, keep, save=True)

        # ---- cyclic G passes, batched as 5B images: pass k = images [kB, (k+1)B) (:576-624)
        cyc_in = ops.new((5 * B, S, S, g_cin5), dt)
        for k in range(5):
            srcs, lds = [], []
            for j in range(5):
                if j == k:
                    srcs.append(None); lds.append(0)
                elif bits[j]:
                    srcs.append(gen_Y); lds.append(1)
                else:
                    srcs.append(ds[j]); lds.append(3)
            ops.assemble_input(srcs, lds, k, cyc_in[k * B:(k + 1) * B])
        cyc_Y_lp, tape5 = G.forward(cyc_in, g_attn, save=True)
        cyc_Y = _f32(cyc_Y_lp)
        if dt == f32:
            cyc_rgb = xB[:5 * B]
            ops.yuv2rgb(cyc_Y, avg, cyc_rgb, None)
        else:
            cyc_rgb = ops.new((5 * B, S, S, 3), f32)
            ops.yuv2rgb(cyc_Y, avg, cyc_rgb, xB[:5 * B])
        for k in range(5):
            to_d(origs[k], out=xB[(5 + k) * B:(6 + k) * B])

        # ---- D passes
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)    # D(xA) ran beside the cyclic generator pass
        else:
            rfA_lp, clsA, tapeA = D.forward(xA, d_attn, noise, keep, save=True)
        rfB_lp, clsB, tapeB = D.forward(xB, d_attn, None, None, save=True)
        rfA, rfB = _f32(rfA_lp), _f32(rfB_lp)
        nrf = B * s32 * s32

        # ---- losses (:669-844): values into the table, seed gradients for the two backward sweeps
        # D-loss seeds: total_D + total_Cls = (D1_cls + D3_cls)/6 + (D2_rf + D4_rf)/6 + 10.5 D4_cls (+ NST, no D dependence)
        dD_rfA, dD_clsA = ops.zeros_like(rfA), ops.zeros_like(clsA)
        dD_rfB, dD_clsB = ops.zeros_like(rfB), ops.zeros_like(clsB)
        # G-loss seeds through D: total_G contains (D1_rf + D3_rf)/6
        dG_rfA, dG_rfB = ops.new((B, s32, s32, 1), f32), ops.new((5 * B, s32, s32, 1), f32)
        sixth = 1.0 / 6.0
        LS.lsgan(rfA[:B], T, tab.slot("D1_rf"), 1.0, dG_rfA, sixth)                                   # :677
        LS.lsgan(rfB[:5 * B], T, tab.slot("D3_rf"), 5.0, dG_rfB, 5.0 * sixth)                        # :669-674 (sum of 5 means)
        # D2_rf = sqd(rf_tgt, T) + mean(rf_gen^2) (:721); it enters total_D twice (D2 + D4, :728,:837-840)
        LS.lsgan(rfA[B:], T, tab.slot("D2_rf"), 1.0, dD_rfA[B:], 2.0 * sixth)
        LS.lsgan(rfA[:B], 0.0, tab.slot("D2_rf"), 1.0, dD_rfA[:B], 2.0 * sixth)
        LS.lsgan(rfB[5 * B:], T, tab.slot("D4_rf_only"), 5.0, dD_rfB[5 * B:], 5.0 * sixth)           # :723-727
        LS.lsgan(rfB[:5 * B], 0.0, tab.slot("D4_rf_only"), 5.0, dD_rfB[:5 * B], 5.0 * sixth)
        LS.softmax_ce(clsA[:B], [0, 0, 0, 0, T] if gp is None else gp.labels, tab.slot("D1_cls"), 1.0, dD_clsA[:B], sixth)   # :702
        for k in range(5):
            onehot = [1.0 if j == k else 0.0 for j in range(5)]
            LS.softmax_ce(clsB[k * B:(k + 1) * B], onehot, tab.slot("D3_cls"), 1.0, dD_clsB[k * B:(k + 1) * B], sixth)      # :695-700
            LS.softmax_ce(clsB[(5 + k) * B:(6 + k) * B], onehot, tab.slot("D4_cls"), 1.0, dD_clsB[(5 + k) * B:(6 + k) * B], 10.5)  # :709-714

        # image-space terms of total_G = ... + 10 L1 + 10 ssim + 10 NST (:829-832)
        d_gen_rgb = ops.new((B, S, S, 3), f32)
        d_cyc_rgb = ops.new((5 * B, S, S, 3), f32)
        LS.l1(gen_rgb, origs[4], tab.slot("L1_G1"), 1.0, d_gen_rgb, 10.0 / 5.0)                      # :744,:751
        for k in range(5):
            w = 10.0 * (10.0 if k == 4 else 0.2)
            LS.l1(cyc_rgb[k * B:(k + 1) * B], origs[k], tab.slot("L1_c%d" % k), 1.0, d_cyc_rgb[k * B:(k + 1) * B], w)   # :745-751
        d_cyc_Y = ops.zeros((5 * B, S, S, 1), f32, cyc_Y.device)
        self.ssim_values = []
        for k in range(5):
            Yk, dYk = cyc_Y[k * B:(k + 1) * B], d_cyc_Y[k * B:(k + 1) * B]
            if not bits[k]:                                                                            # :774-778
                w = 10.0 * (10.0 if k == 4 else 1.0) / 5.0
                self.ssim_values.append(LS.ssim_term(Yk, avg, ds[k], tab.slot("ssim%d" % k), 1.0, dYk, w))
            LS.spec(Yk, avg, ds[k], self.specular_candidate, tab.slot("spec%d" % k), 1.0)            # :792-796 (value only)
        Y4, dY4 = cyc_Y[4 * B:], d_cyc_Y[4 * B:]
        LS.mse_ycc(Y4, avg, ds[0], tab.slot("content"), 1.0, dY4, 10.0)                               # :814 (vs ds1, as written)
        LS.style(Y4, avg, ds[4], S, tab.slot("style"), 1.0, dY4, 10.0 * 100.0)                        # :817-826

        # ---- backward: D weight gradients (:859), then the generator loss through D (dgrad only) and G (:868)
        D.store.zero_grad()
        G.store.zero_grad()
        d_dattn = ops.zeros_like(d_attn) if self.live_mask else None

        def d_weight_sweep():
            D.backward(tapeA, dD_rfA, dD_clsA, wgrad=True, need_dx=False, dattn=d_dattn, attn_nb=B)
            D.backward(tapeB, dD_rfB, dD_clsB, wgrad=True, need_dx=False, dattn=d_dattn, attn_nb=B)
            if self.live_mask:
                D.attention_backward(d_attn_saved, d_dattn)
            D.store.finalize_grads()
            if reducer is not None:
                reducer.reduce_async(D.store.grad)

        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                d_weight_sweep()
        else:
            d_weight_sweep()
        dxA = D.backward(tapeA, dG_rfA, None, n=B, wgrad=False, need_dx=True)
        dxB = D.backward(tapeB, dG_rfB, None, n=5 * B, wgrad=False, need_dx=True)
        lpA, lpB = (None, None) if dt == f32 else (dxA, dxB)
        if dt == f32:
            ops.axpy(1.0, dxA, d_gen_rgb)
            ops.axpy(1.0, dxB, d_cyc_rgb)
        ops.yuv2rgb_bwd(d_cyc_rgb, lpB, d_cyc_Y, accumulate=True)
        g_dattn = [ops.zeros_like(a) for a in g_attn] if self.live_mask else None
        any_dropped = any(bits)
        d_cyc_in = G.backward(tape5, ops.cast(d_cyc_Y, dt), g_dattn, attn_nb=B, need_dx=any_dropped)
        d_gen_Y = ops.new((B, S, S, 1), f32)
        ops.yuv2rgb_bwd(d_gen_rgb, lpA, d_gen_Y, accumulate=False)
        if any_dropped:                                    # gen_Y feeds the dropped slots of the cyclic inputs (:576-580)
            for k in range(5):
                slots = [j for j in range(5) if j != k and bits[j]]
                if slots:
                    ops.assemble_bwd(d_cyc_in[k * B:(k + 1) * B], slots, d_gen_Y)
        # data parallel: this is the LAST pass that touches G's gradients, and it finishes the decoder first -- every range of the flat
        # gradient buffer is all-reduced (NCCL's stream) as soon as its last weight-gradient kernel is enqueued, while the rest of the
        # backward keeps computing; only the encoder / attention head of the buffer is reduced after the sweep
        hook = None
        if reducer is not None and self.dp_overlap:
            def hook(stage):
                lo, hi = G.grad_range(stage)
                reducer.reduce_async(G.store.grad, lo, hi)
        G.backward(tape1, ops.cast(d_gen_Y, dt), g_dattn, attn_nb=B, need_dx=False, hook=hook)
        if self.live_mask:
            G.attention_backward(g_attn_saved, g_dattn)
        G.store.finalize_grads()
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if reducer is not None:
            if hook is not None:
                hook("head")
            else:
                reducer.reduce_async(G.store.grad)           # dp_overlap = False: one reduction of the whole buffer after the sweep
            reducer.wait()

        # ---- clip_by_value(+-1) + Adam (:860-871); both optimisers use g_lr (:169-174)
        if with_opt:
            gscale = 1.0 if reducer is None else 1.0 / reducer.world
            D.store.adam_step(self.g_lr, self.beta1, self.beta2, 1e-7, 1.0, gscale, lr_dev=None if gp is None else gp.lr_D)
            G.store.adam_step(self.g_lr, self.beta1, self.beta2, 1e-7, 1.0, gscale, lr_dev=None if gp is None else gp.lr_G)
        self.step_count += 1

        # ---- published tensors / scalars (reference attribute names)
        self.gen_input, self.gen_Y, self.gen_rgb = gen_in[..., :10], gen_Y, gen_rgb
        self.averageCbCr = avg
        self.ds_yuv = ds
        self.cyc_Y = [cyc_Y[k * B:(k + 1) * B] for k in range(5)]
        rgbs = [cyc_rgb[k * B:(k + 1) * B] for k in range(5)]
        self.cyc_gen0_rgb, self.cyc_gen45_rgb, self.cyc_gen90_rgb, self.cyc_gen135_rgb, self.cyc_genED_rgb = rgbs
        self.RealFake_gen_D1, self.label_gen_D1 = rfA[:B], clsA[:B]
        self.RealFake_target_D2, self.label_target_D2 = rfA[B:], clsA[B:]
        self.RealFake_cyc_D3 = [rfB[k * B:(k + 1) * B] for k in range(5)]
        self.label_cyc_D3 = [clsB[k * B:(k + 1) * B] for k in range(5)]
        self.RealFake_orig_D4 = [rfB[(5 + k) * B:(6 + k) * B] for k in range(5)]
        (self.label_orig0_D4, self.label_orig45_D4, self.label_orig90_D4, self.label_orig135_D4,
         self.label_origED_D4) = [clsB[(5 + k) * B:(6 + k) * B] for k in range(5)]

    def _start_loss_readback(self):
        """Asynchronous device->host copy of the loss table into pinned memory; the scalars are published when first read
        (`net.total_Generator_loss` ...), so a caller that does not look at them every step never stalls the stream."""
        i = self.step_count & 1
        if self._loss_host[i] is None:
            self._loss_host[i] = torch.empty(self.table.buf.numel(), dtype=torch.float32, pin_memory=True)
        self._loss_host[i].copy_(self.table.buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending_losses = (self._loss_host[i], ev)

    def _finish_losses(self):
        host, ev = self._pending_losses
        ev.synchronize()
        self._pending_losses = None
        self._publish_losses({k: float(host[i]) for k, i in LS.IDX.items()})

    def _side_stream(self):
        if not self.overlap:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def _publish_losses(self, v):
        """Totals of :669-844 from the per-term table (one device->host copy per step)."""
        self.loss_values = dict(v)
        L1 = (v["L1_c0"] + v["L1_c1"] + v["L1_c2"] + v["L1_c3"] + v["L1_G1"]) / 5.0 + v["L1_c4"] * 10.0
        ssim = (v["ssim0"] + v["ssim1"] + v["ssim2"] + v["ssim3"] + v["ssim4"] * 10.0) / 5.0
        spec = (v["spec0"] + v["spec1"] + v["spec2"] + v["spec3"]) / 5.0 + v["spec4"] * 5.0
        nst = 100.0 * v["style"] + v["content"]
        D4_rf = v["D4_rf_only"] + v["D2_rf"]
        self.D1_RealFake_loss, self.D3_RealFake_cyc = v["D1_rf"], v["D3_rf"]
        self.D1_classification_loss, self.D3_classification_loss = v["D1_cls"], v["D3_cls"]
        self.D2_RealFake_target = v["D2_rf"]
        self.D4_RealFake_cyc, self.D4_classification_loss = D4_rf, v["D4_cls"]
        self.G_gan_loss = (v["D3_rf"] + v["D1_rf"]) / 6.0
        self.G_clsf_loss = (v["D3_cls"] + v["D1_cls"]) / 6.0
        self.L1_loss_Gen, self.ssim_cyc_loss, self.Spec_loss = L1, ssim, spec
        self.content_loss, self.style_loss, self.total_NST_loss = v["content"], v["style"], nst
        self.total_Generator_loss = (v["D1_rf"] + v["D3_rf"]) / 6.0 + 10.0 * L1 + 10.0 * ssim + 10.0 * nst
        self.total_Discriminator_loss = ((v["D1_cls"] + v["D3_cls"]) / 6.0 + (v["D2_rf"] + D4_rf) / 6.0
                                         + 0.5 * v["D4_cls"] + 10.0 * nst)
        self.total_Classification_loss = 10.0 * (v["D4_cls"] + nst)

    _INFER_PUBLISHED = ("specular_candidate", "gen_Y", "gen_rgb", "cyc_rgb", "cyc_gen0_rgb", "cyc_gen45_rgb", "cyc_gen90_rgb", "cyc_gen135_rgb",
                        "cyc_genED_rgb")

    def inference_step(self, rgb, cyclic: bool = False):
        """The per-image body of test.py:218-297: standardise -> SpecSeg mask -> G1 with only slot 0 populated and the ED
        one-hot plane -> yuv->rgb with the image's own CbCr.  Returns gen_rgb [B,S,S,3] fp32 (and publishes gen_Y, mask).
        With net.cuda_graph the ~190 launches are captured once per (batch shape, cyclic, weight version) and replayed; the returned /
        published tensors are then rewritten by the next call."""
        self.build()
        self._require_mask_weights()
        self.__dict__.pop("gen_rgb_output", None)
        rgb = rgb.contiguous()
        if not self.cuda_graph or ops.PROF is not None:
            return self._inference_body(rgb, cyclic)
        key = ("infer", tuple(rgb.shape), bool(cyclic), self.G.net.store.version, self.SpecSeg.net.store.version)
        entry = self._graphs.get(key)
        if entry is None and not self._graphs.get(("infer-warm", tuple(rgb.shape), bool(cyclic))):
            self._graphs[("infer-warm", tuple(rgb.shape), bool(cyclic))] = True       # first call: eager (lazy allocations, weight re-layout)
            return self._inference_body(rgb, cyclic)
        if entry is None:
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            static = ops.new(tuple(rgb.shape), torch.float32)
            static.copy_(rgb, non_blocking=True)
            self._inference_body(static, cyclic)            # THIS call's result, eagerly; the weight copies of this version are now in place
            mine = {k: self.__dict__[k] for k in self._INFER_PUBLISHED if k in self.__dict__}
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launches()
            with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                self._inference_body(static, cyclic)        # recorded, not executed
            entry = (g, {k: self.__dict__[k] for k in self._INFER_PUBLISHED if k in self.__dict__}, L.launches() - n0, static)
            for old in [k for k in self._graphs if k[0] == "infer" and k[1:3] == key[1:3]]:
                del self._graphs[old]                       # graphs recorded against older weights
            self._graphs[key] = entry
            self.__dict__.update(mine)
            return self.gen_rgb
        g, pub, nk, static = entry
        if static.data_ptr() != rgb.data_ptr():
            static.copy_(rgb, non_blocking=True)
        g.replay()
        L.count_replayed(nk)
        self.__dict__.update(pub)
        return self.gen_rgb

    def _inference_body(self, rgb, cyclic):
        ops.arena_begin(whole=torch.cuda.is_current_stream_capturing())
        G = self.G.net
        rgb = rgb.contiguous()
        B, S = rgb.shape[0], rgb.shape[1]
        dt = self.dtype
        yuv, scale = ops.yuv_standardize(rgb)
        self.stddev_arr.append(scale)                                                                  # :1306 via test.py:218
        mask = self.SpecSeg.net.predict(self._y_plane(yuv, dt))
        self.specular_candidate = _f32(mask)
        attn = G.attention(mask, infer=True)[0] if self.live_mask else None
        gin = ops.new((B, S, S, G.in_channels(B, S, S, infer=True)), dt)
        ops.assemble_input([yuv, None, None, None, None], [3, 0, 0, 0, 0], 4, gin)                   # test.py:227-235
        self.gen_Y = _f32(G.forward(gin, attn))
        cbcr = ops.new((B, S, S, 2), torch.float32)
        ops.cast_into(yuv[..., 1:3], cbcr)                                                             # test.py:224
        self.gen_rgb = ops.new((B, S, S, 3), torch.float32)
        ops.yuv2rgb(self.gen_Y, cbcr, self.gen_rgb, None)                                              # test.py:244-250
        if cyclic:                                          # test.py:252-284 (Q11: the R channel stands in for Y)
            R = ops.new((B, S, S, 1), torch.float32)
            ops.cast_into(self.gen_rgb[..., 0:1], R)
            cin = ops.new((5 * B, S, S, G.in_channels(5 * B, S, S, infer=True)), dt)
            for k in range(5):
                srcs = [None if j == k else R for j in range(5)]
                ops.assemble_input(srcs, [1] * 5, k, cin[k * B:(k + 1) * B])
            cy = _f32(G.forward(cin, attn))
            crgb = ops.new((5 * B, S, S, 3), torch.float32)
            ops.yuv2rgb(cy, cbcr, crgb, None)
            self.cyc_rgb = [crgb[k * B:(k + 1) * B] for k in range(5)]
            self.cyc_gen0_rgb, self.cyc_gen45_rgb, self.cyc_gen90_rgb, self.cyc_gen135_rgb, self.cyc_genED_rgb = self.cyc_rgb   # test.py:293-297
        return self.gen_rgb

    # -- data parallel ----------------------------------------------------------------------------------------------------
    def enable_data_parallel(self, process_group=None, bucket_mb: float = 25.0):
        """Shard the batch over the ranks of `process_group` (NCCL): gradients are summed in buckets on a side stream,
        overlapped with the rest of the backward; clip + Adam run on the average."""
        from .parallel import GradReducer
        self.build()
        self._reducer = GradReducer(process_group, bucket_mb)
        stores = [self.G.net.store, self.D.net.store, self.SpecSeg.net.store]
        self._reducer.broadcast_params([st.flat for st in stores])
        for st in stores:
            st.version += 1                                  # bf16 tensor-core copies made before the broadcast are stale
        # one global batch = independent GaussianNoise / Dropout draws per sample: every rank draws from its own Philox stream
        # (the 5 drop bits and T stay shared: they are per-step scalars of the whole batch)
        self.noise_seed = self.seed + 1000003 * self._reducer.rank
        return self

    # -- host-side state a resumed run needs beside the parameters (checkpoint.py) ------------------------------------------
    def host_state(self) -> dict:
        return {"step_count": self.step_count, "rng": list(self._rng.getstate()[1]), "rng_gauss": self._rng.getstate()[2],
                "d_calls": 0 if self.D is None else self.D.calls, "noise_seed": self.noise_seed, "epoch": self.epoch,
                "stddev": None if self.stddev_arr is None else self.stddev_arr.state()}

    def load_host_state(self, st: dict):
        self.build()
        self.step_count = int(st["step_count"])
        self._rng.setstate((3, tuple(int(v) for v in st["rng"]), st.get("rng_gauss")))
        self.D.calls = int(st.get("d_calls", 0))
        self.noise_seed = int(st.get("noise_seed", self.noise_seed))
        self.epoch = int(st.get("epoch", 0))
        if st.get("stddev") is not None:
            self.stddev_arr.load_state(st["stddev"])
