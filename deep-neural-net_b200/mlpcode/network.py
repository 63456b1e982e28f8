"""`Network` for the B200 path — same public surface as mlpcode/network.py:22-484.

Construction, `compile`, `train`, `predict`, `predict_classes`, `evaluate`, `get_accuracy`,
`addCallbacks` / `clearCallbacks`, `load_weights`, `loadBatchNormParameters`, `weights`, `biases`,
`batchNormParams` behave as in the reference; the per-batch work is delegated to the device layers
in mlpcode.layers.  Differences, all forced by the environment and documented in DESIGN.md:

  * only binarized, batch-normalised, bias-free networks run (the BNN configuration);
  * checkpoints are `.npz` files with the reference's HDF5 key layout (h5py is not in the image);
  * `train_step(X, y)` exposes the body of the private `__updateBatch` (network.py:372-392) as the
    public per-batch call that bench.py and the data-parallel driver use;
  * with torch.distributed initialised, `set_data_parallel()` shards nothing itself — the caller
    feeds each rank its batch shard — but makes batch-norm statistics, the loss gradient and dW
    global, so N ranks reproduce the single-GPU (and reference) step.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Union

import numpy as np
import torch

from . import checkpoint
from . import kernels as K_
from . import loss as loss_mod
from .activation import ACTIVATION_FUNCTIONS
from .activation import ActivationFuncs as af
from .callbacks import Callback
from .device import DeviceArray, asarray, asnumpy, to_tensor, zeros
from .layers import BatchNormLayer, BinaryLayer, LinearLayer, _NoComm
from .loss import LOSS_DERIVATES, LOSS_FUNCS
from .loss import LossFuncs as lf
from .optim import LRScheduler

ModelLayer = Union[LinearLayer, BinaryLayer, BatchNormLayer]
MODELDIR = Path(__file__).resolve().parent / "models"


class Network(object):
    def __init__(self, units: List[int], useGpu=False, binarized=False, useBias=False,
                 useBatchNorm=False):
        if useBatchNorm and useBias:
            print("Won't use bias with BatchNorm, as it will automatically get cancelled in the "
                  "normalization step")
            useBias = False
        self.isBinarized = binarized
        self.useBias = useBias
        self.useGpu = useGpu
        self.useBatchNorm = useBatchNorm
        self.unitList = list(units)
        layer_cls = BinaryLayer if binarized else LinearLayer
        self._layers: List[ModelLayer] = [
            layer_cls(n_out, n_in, gpu=useGpu, useBias=useBias, batchNorm=useBatchNorm)
            for n_out, n_in in zip(units[1:], units[:-1])
        ]
        for i, layer in enumerate(self._layers):
            layer.layer_index = i
            layer.input_layer = self._layers[i - 1] if i > 0 else None
            layer.output_layer = self._layers[i + 1] if i + 1 < len(self._layers) else None
        if self._layers:
            self._layers[0].skip_input_grad = True
        self._lossF: lf = None
        self._outAf: af = None
        self._lr: LRScheduler = None
        self._compiled = False
        self._comm = _NoComm()
        self._graph_on = False
        self._graphs = {}
        self._dw_overlap = None

    # ------------------------------------------------------------------ parameters
    @property
    def isCompiled(self) -> bool:
        return self._compiled

    @property
    def weights(self):
        return [layer.weights for layer in self._layers]

    @property
    def biases(self):
        if self.useBias:
            return [layer.bias for layer in self._layers]
        return [None for _ in self.unitList[:-1]]

    @property
    def batchNormParams(self):
        assert self.useBatchNorm
        cols = ([], [], [], [])
        for layer in self._layers:
            for dst, v in zip(cols, layer.batchNormParams):
                dst.append(v)
        return cols

    def _copyBnParams(self):
        return tuple([p.copy() for p in group] for group in self.batchNormParams)

    @property
    def linearLayers(self):
        return [layer for layer in self._layers if isinstance(layer, LinearLayer)]

    def load_weights(self, weights, biases=None):
        if not self.useBias:
            biases = [None] * len(weights)
        assert len(weights) == len(self._layers) == len(biases)
        for layer, w, b in zip(self.linearLayers, weights, biases):
            layer.load_parameters(w, b)

    def loadBatchNormParameters(self, gammas, betas, mus, sigmas):
        n = len(self._layers)
        assert len(gammas) == len(betas) == len(mus) == len(sigmas) == n
        for layer, g, b, m, s in zip(self._layers, gammas, betas, mus, sigmas):
            layer.loadBatchNormParams(g, b, m, s)

    # ------------------------------------------------------------------ callbacks
    def addCallbacks(self, callbacks: Union[Callback, List[Callback]], num_layers: int = None):
        if num_layers is None:
            num_layers = len(self._layers)
        assert num_layers > 0
        if isinstance(callbacks, list):
            assert len(callbacks) == num_layers
        else:
            callbacks = [callbacks] * num_layers
        for layer, cb in zip(self._layers[:num_layers], callbacks):
            layer.addCallback(cb)

    def clearCallbacks(self):
        for layer in self._layers:
            layer.clearCallbacks()

    # ------------------------------------------------------------------ data parallel
    def set_data_parallel(self, comm) -> None:
        """comm: object with `.world` and `.allreduce_sum(tensor)` (see mlpcode.parallel)."""
        self._comm = comm
        if hasattr(comm, "enable_peer_exchange"):
            comm.enable_peer_exchange(max(self.unitList[1:]), n_sites=2 * len(self._layers))
        for i, layer in enumerate(self._layers):
            layer.set_comm(comm)
            if getattr(layer, "batchNorm", False):
                layer.batchNormLayer.exchange_site = 2 * i
        if hasattr(comm, "enable_dw_exchange") and comm.enable_dw_exchange(
                [layer.weights_shape for layer in self._layers]):
            for i, layer in enumerate(self._layers):
                layer.attach_dw_exchange(comm.dw_exchange, i)

    # ------------------------------------------------------------------ checkpoints (.npz bridge)
    @staticmethod
    def fromModel(filePth, useGpu=False, binarized=False):
        """Load a checkpoint with the key layout of the reference's HDF5 files (network.py:157-228: units,
        useBias, useBatchNorm, weights/weights_i, gammas/gammas_i, ...), stored as `.npz` because h5py is
        unavailable — see mlpcode.checkpoint.  A packed-only file (sign words, no latent weights) loads as
        latent weights of +-1."""
        filePth = Path(filePth)
        assert filePth.exists()
        ck = checkpoint.read_checkpoint(filePth)
        useBias = ck["useBias"]
        if ck["useBatchNorm"] and useBias:
            print("Setting useBias to false cause of batchnorm")
            useBias = False
        nn = Network(ck["units"], useGpu=useGpu, useBias=useBias, binarized=binarized,
                     useBatchNorm=ck["useBatchNorm"])
        nn.load_weights(ck["weights"], ck["biases"] if useBias else None)
        if ck["useBatchNorm"]:
            nn.loadBatchNormParameters(*[ck[g] for g in checkpoint.BN_GROUPS])
        return nn

    def save_weights(self, modelName: str, binarized=False, directory=None, latent=True):
        """network.py:428-472 with `.npz` for `.hdf5` (same members; no timestamp in the name so that scripts can
        find their file).  binarized=True — in the reference only a file-name suffix — also stores the packed
        sign words `packed_weights/packed_weights_i`; latent=False then drops the float32 weights (a 1-bit-per-
        weight deployment file that fromModel reads back as +-1)."""
        directory = Path(directory) if directory is not None else MODELDIR
        path = directory / f"{modelName}{'_binarized' if binarized else ''}.npz"
        bn = [[asnumpy(v) for v in group] for group in self.batchNormParams] if self.useBatchNorm else None
        biases = [asnumpy(b) for b in self.biases] if self.useBias else None
        checkpoint.write_checkpoint(path, self.unitList, self.useBias, self.useBatchNorm,
                                    [asnumpy(w) for w in self.weights], biases, bn,
                                    packed=bool(binarized), latent=latent)
        print(f"Saving model to {path}")
        return path

    # ------------------------------------------------------------------ compile
    def compile(self, lr: Union[LRScheduler, float] = 1e-3, hiddenAf: af = af.sigmoid,
                outAf: af = af.softmax, lossF: lf = lf.cross_entropy) -> None:
        assert hiddenAf in ACTIVATION_FUNCTIONS
        assert outAf in ACTIVATION_FUNCTIONS
        assert lossF in LOSS_FUNCS
        if self.isBinarized and hiddenAf != af.sign:
            print(f"Changing hidden activation function to {af.sign} for BNN")
            hiddenAf = af.sign
        for layer in self._layers[:-1]:
            layer.build(hiddenAf)
        self._layers[-1].build(outAf)
        self._lossF, self._outAf = lossF, outAf
        if isinstance(lr, float):
            self._lr = LRScheduler(alpha=lr)
        elif isinstance(lr, LRScheduler):
            self._lr = lr
        else:
            raise ValueError("Invalid value for learning rate (only float or LRScheduler allowed)")
        self._compiled = True

    # ------------------------------------------------------------------ training
    def train(self, trainX, trainY, epochs: int, batch_size=1, shuffle=True, valX=None, valY=None,
              save_best_params=True):
        if not self._compiled:
            print("\nEXCEPTION: Must compile the model before running train")
            return
        trainX = asarray(trainX, np.float32)
        trainY = asarray(trainY)
        if self._lossF == lf.hinge:   # one-hot {0, 1} -> {-1, +1}; the reference edits trainY in place
            yt = trainY.t             # (network.py:301-303), here a copy carries the change
            trainY = DeviceArray(torch.where(yt == 0, torch.full_like(yt, -1), yt))
        n = len(trainX)
        best_acc, best_epoch, best = -1.0, -1, None
        costs, train_accs, val_accs = [], [], []
        has_val = valX is not None
        if has_val:
            valX = asarray(valX, np.float32)
            valY = asarray(valY)
            if valY.shape[0] != valY.size:  # one-hot -> labels (network.py:298-299)
                valY = valY.argmax(axis=1)
        print(f"\n\nStarting training (binarized: {self.isBinarized})")
        print("=" * 20)
        print()
        X_all, Y_all = trainX.t.contiguous(), trainY.t.contiguous()
        if Y_all.ndim == 1:
            Y_all = Y_all.unsqueeze(1)
        order = None        # composed permutation: the reference's trainX = trainX[p, :] of every epoch so far
        step = n if batch_size == -1 else batch_size
        bad = torch.zeros(1, dtype=torch.int32, device=X_all.device)
        for epoch in range(epochs):
            if shuffle:
                # network.py:312-315 shuffles the whole data set (a 2 x n x F byte copy per epoch); here only the
                # index vector is permuted and bnn_gather_rows assembles each mini-batch straight into the
                # training step's static buffers.  X[p0][p1] == X[p0[p1]], so the batches are the reference's.
                p = self._draw_permutation(n, X_all.device)
                order = p if order is None else order[p]
            # the epoch stays on the device: every batch's mean loss is parked in a device vector and
            # the host reads them once per epoch (the reference syncs on float(cost.mean()) per batch,
            # network.py:378); graph replays reuse one loss buffer, hence the copy right after each step
            n_batches = (n + step - 1) // step
            batch_costs = torch.zeros(n_batches, dtype=torch.float32, device=X_all.device)
            for i, k in enumerate(range(0, n, step)):
                if order is None:
                    rows = self.train_step_async(X_all[k:k + step], Y_all[k:k + step])
                else:
                    rows = self.train_step_async(None, None, gather=(X_all, Y_all, order[k:k + step], bad))
                batch_costs[i] = rows.mean().t
            self._lr.step()
            cost = float(batch_costs.double().sum()) / n_batches
            assert int(bad.item()) == 0, "permutation index outside the training set"
            acc = self.get_accuracy(trainX, trainY.argmax(axis=1))   # a count: independent of the row order
            train_accs.append(acc)
            costs.append(cost)
            line = f"Epoch {epoch + 1} / {epochs}\tAccuracy : {acc:0.3f}%"
            if has_val:
                acc = self.get_accuracy(valX, valY)
                val_accs.append(acc)
                line += f"\tVal Acc: {acc:0.3f}%"
            print(line + f"\tLoss: {cost:.02f}")
            if save_best_params and acc > best_acc:
                best_acc, best_epoch = acc, epoch
                best = ([w.copy() for w in self.weights],
                        self._copyBnParams() if self.useBatchNorm else None)
        if save_best_params and best is not None:
            kind = "Validation" if has_val else "Training"
            print("\nBest {0} Accuracy:\t{1:.03f}% (epoch: {2})".format(kind, float(best_acc),
                                                                       best_epoch))
            print("Switching to best params\n")
            self.load_weights(best[0])
            if self.useBatchNorm:
                self.loadBatchNormParameters(*best[1])
        history = {"loss": costs, "accuracy": train_accs}
        if has_val:
            history["val_accuracy"] = val_accs
        return history

    def _draw_permutation(self, n: int, device) -> torch.Tensor:
        """Row permutation of one epoch as a device int64 vector.  `shuffle_source` (a callable n -> index array,
        e.g. np.random.RandomState(seed).permutation) replays a given stream — the reference draws from NumPy's
        global one (network.py:313) — otherwise torch.randperm on the device."""
        src = getattr(self, "shuffle_source", None)
        if src is None:
            return torch.randperm(n, device=device)
        p = to_tensor(np.ascontiguousarray(src(n))).to(device=device, dtype=torch.int64).contiguous()
        assert p.numel() == n
        return p

    def _gather_batch(self, gather, Xb=None, yb=None):
        """Assemble the mini-batch rows gather = (X_all, Y_all, idx, bad) into (Xb, yb) (allocated if None)."""
        X_all, Y_all, idx, bad = gather
        idx = idx.contiguous()
        if Xb is None:
            Xb = torch.empty((idx.numel(), X_all.shape[1]), dtype=X_all.dtype, device=X_all.device)
            yb = torch.empty((idx.numel(), Y_all.shape[1]), dtype=Y_all.dtype, device=Y_all.device)
        K_.gather_rows(X_all, idx, Xb, bad)
        K_.gather_rows(Y_all, idx, yb, bad)
        return Xb, yb

    # ------------------------------------------------------------------ CUDA-graph replay of a step
    def enable_step_graph(self, on: bool = True) -> None:
        """Replay each training step as ONE CUDA graph launch instead of ~90 kernel launches.

        A step is a fixed sequence of libbnn_b200 kernels over fixed buffers; the only things that
        vary are the batch (copied into static device buffers) and, in data-parallel runs, the
        exchange epoch (a device word the kernels read).  The first two steps of a given (batch shape,
        lr, global batch, parameter identity) run eagerly to populate the workspaces, the third is
        captured (copy-engine all-gather on its side stream included) and every later one is a
        replay.  Loading new parameters or changing lr re-captures.  Results are identical to eager
        execution: the same kernels with the same arguments."""
        self._graph_on = bool(on)
        self._graphs = {}

    def _param_version(self) -> int:
        v = 0
        for layer in self._layers:
            v += getattr(layer, "version", 0)
            if getattr(layer, "batchNorm", False):
                v += getattr(layer.batchNormLayer, "version", 0)
        return v

    def _graph_step(self, X, y, lr, global_batch, consumed=None, gather=None):
        if self._comm.world > 1 and (getattr(self._comm, "dw_exchange", None) is None
                                     or getattr(self._comm, "peer_exchange", None) is None):
            return None  # NCCL collectives in the step: stay eager
        if gather is not None:   # shapes of the batch the gather will assemble
            rows = gather[2].numel()
            shape_x, shape_y, ydt = (rows, gather[0].shape[1]), (rows, gather[1].shape[1]), gather[1].dtype
            if gather[0].dtype != torch.float32:
                return None
        else:
            Xt, yt = to_tensor(X, torch.float32), to_tensor(y)
            shape_x, shape_y, ydt = tuple(Xt.shape), tuple(yt.shape), yt.dtype
        key = (shape_x, shape_y, ydt, float(lr), global_batch, self._param_version())
        st = self._graphs.get(key)
        if st is None:
            # graphs of other parameter identities are stale; of the rest keep the two most recent
            # (each captured step pins its own pool of activations, and an lr schedule would
            # otherwise grow one graph per lr value)
            keep = [(k, v) for k, v in self._graphs.items() if k[-1] == key[-1]][-2:]
            self._graphs = dict(keep)
            st = self._graphs[key] = {"warm": 0}
        if st["warm"] < 2:
            st["warm"] += 1
            return None
        if "graph" not in st:
            if gather is not None:
                st["X"], st["y"] = self._gather_batch(gather)
            else:
                st["X"], st["y"] = Xt.clone(), yt.clone()
            if consumed is not None:
                consumed.record(torch.cuda.current_stream())
            launches0 = K_.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                st["cost"] = self._train_step_eager(DeviceArray(st["X"]), DeviceArray(st["y"]), lr,
                                                    global_batch)
            st["graph"], st["launches"] = g, K_.launch_count - launches0
            K_.launch_count = launches0
        else:
            if gather is not None:   # the permuted rows land directly in the graph's static batch buffers
                self._gather_batch(gather, st["X"], st["y"])
            else:
                st["X"].copy_(Xt, non_blocking=True)
                st["y"].copy_(yt, non_blocking=True)
            if consumed is not None:  # the caller's batch buffers are free again: the step runs on copies
                consumed.record(torch.cuda.current_stream())
        st["graph"].replay()
        K_.launch_count += st["launches"]
        return st["cost"]

    def train_step_async(self, X, y, lr: float = None, global_batch: int = None, consumed=None, gather=None):
        """One SGD step on one batch (the body of `__updateBatch`, network.py:372-392) without the
        host read of the loss; returns the device array of per-sample losses.  Data-parallel runs
        with unequal shards pass the global batch size.  consumed: optional CUDA event recorded on the
        current stream at the point after which X and y are no longer read (graph replays work on
        copies, so that is right after the copies; eager steps read them until the end)."""
        lr = self._lr.value if lr is None else lr
        if self._graph_on and K_.gemm_events is None:
            cost = self._graph_step(X, y, lr, global_batch, consumed, gather)  # records `consumed` itself
            if cost is not None:
                return cost
        if gather is not None:   # gather: (X_all, Y_all, idx, bad) — the batch is rows idx of the data set
            X, y = (DeviceArray(t) for t in self._gather_batch(gather))
        cost = self._train_step_eager(X, y, lr, global_batch)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())
        return cost

    def _train_step_eager(self, X, y, lr: float, global_batch: int = None):
        if self._comm.world > 1:
            self._comm.global_batch = global_batch
        # side stream for the weight-gradient chain: single GPU, and data-parallel runs whose dW exchange is
        # the peer-memory one (an NCCL all-reduce has its own stream and handle)
        if (self._dw_overlap is None and os.environ.get("BNN_OVERLAP_DW", "1") != "0"
                and (self._comm.world == 1 or getattr(self._comm, "dw_exchange", None) is not None)):
            from .layers import DwOverlap
            self._dw_overlap = DwOverlap()
            for layer in self._layers:
                layer._dw_overlap = self._dw_overlap
        self._comm.begin_step()
        output = self.predict(X, isTraining=True)
        loss_mod.GLOBAL_BATCH_ROWS = self._comm.global_rows(X.shape[0])
        cost = LOSS_FUNCS[self._lossF](output, y)
        fused = self._lossF == lf.cross_entropy and self._outAf in (af.softmax, af.sigmoid)
        delta = LOSS_DERIVATES[self._lossF](output, y, with_softmax=fused)
        delta = self._layers[-1].backwards(delta, lr, activDeriv=not fused)
        for layer in reversed(self._layers[:-1]):
            delta = layer.backwards(delta, lr)
        # data-parallel: the peer-memory exchange is completed here (copy-engine stream joined, every
        # rank's rows landed); an NCCL all-reduce still in flight is completed lazily, when that
        # layer's weights are next read
        self._comm.finish_step()
        if self._dw_overlap is not None:
            self._dw_overlap.join()  # the side-stream dW GEMMs have updated W before the step ends
        return cost

    def train_step(self, X, y, lr: float = None, global_batch: int = None) -> float:
        return float(self.train_step_async(X, y, lr, global_batch).mean())

    _Network__updateBatch = lambda self, batch, lr: self.train_step(batch[0], batch[1], lr)  # noqa: E731

    def train_stream(self, batches, lr: float = None, global_batch: int = None) -> List[float]:
        """Run one SGD step per (X, y) host batch and return the per-batch mean losses, pipelined:
        the host->device copy of batch i+1 (from pinned memory, on a copy stream, into the other of two
        device buffers) overlaps the kernels of batch i, and the loss of batch i is read back while
        batch i+1 runs.  Every batch is still copied in and every loss still copied out; only the
        serialisation of `train_step` (copy, compute, blocking read) is removed."""
        compute = torch.cuda.current_stream()
        st = getattr(self, "_stream_state", None)
        if st is None:  # created once: stream, events, pinned loss words and device buffers are reused
            st = dict(copier=torch.cuda.Stream(), ready=[torch.cuda.Event(), torch.cuda.Event()],
                      freed=[torch.cuda.Event(), torch.cuda.Event()],
                      loss_done=[torch.cuda.Event(), torch.cuda.Event()], dev=[None, None],
                      loss_host=[torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)])
            self._stream_state = st
        copier, ready, freed, loss_done = st["copier"], st["ready"], st["freed"], st["loss_done"]
        dev, loss_host = st["dev"], st["loss_host"]
        for e in freed:
            e.record(compute)
        rows = [0, 0]
        losses: List[float] = []

        def stage(slot, batch):
            X, y = batch
            X = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X))
            y = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
            with torch.cuda.stream(copier):
                copier.wait_event(freed[slot])  # the step that last read this buffer has finished
                if dev[slot] is None or dev[slot][0].shape != X.shape or dev[slot][1].shape != y.shape:
                    dev[slot] = (torch.empty(X.shape, dtype=torch.float32, device="cuda"),
                                 torch.empty(y.shape, dtype=torch.int8, device="cuda"))
                dev[slot][0].copy_(X, non_blocking=True)
                dev[slot][1].copy_(y, non_blocking=True)
                ready[slot].record(copier)

        it = iter(batches)
        try:
            stage(0, next(it))
        except StopIteration:
            return losses
        i = 0
        more = True
        while more:
            slot = i & 1
            try:
                stage(slot ^ 1, next(it))  # queued before this step's kernels so the two overlap
            except StopIteration:
                more = False
            compute.wait_event(ready[slot])
            Xd, Yd = dev[slot]
            cost = self.train_step_async(Xd, Yd, lr, global_batch, consumed=freed[slot])
            loss_host[slot].copy_(cost._sum, non_blocking=True)
            loss_done[slot].record(compute)
            rows[slot] = Xd.shape[0]
            if i > 0:  # read the previous step's loss while this one runs
                prev = slot ^ 1
                loss_done[prev].synchronize()
                losses.append(float(loss_host[prev][0]) / rows[prev])
            i += 1
        last = (i - 1) & 1
        loss_done[last].synchronize()
        losses.append(float(loss_host[last][0]) / rows[last])
        return losses

    # ------------------------------------------------------------------ inference
    def predict(self, X, isTraining=False):
        a = X
        for layer in self._layers:
            a = layer.forward(a, isTrain=isTraining)
        return a

    def predict_classes(self, X):
        return self.predict(X, isTraining=False).argmax(axis=1)

    def evaluate(self, X, y, batch_size=-1):
        """Number of correctly classified rows; y holds class labels (not one-hot)."""
        labels = to_tensor(y).reshape(-1).to(torch.int32)
        correct = zeros((1,), torch.int32)
        n = X.shape[0]
        step = n if batch_size == -1 else batch_size
        for k in range(0, n, step):
            bx = X[k:k + step]
            scores = self.predict(bx, isTraining=False)
            st = scores.t
            K_.count_correct(st, st.shape[0], st.shape[1], labels[k:k + step].contiguous(), correct)
        return int(correct.item())

    def get_accuracy(self, X, y, batch_size=-1):
        return self.evaluate(X, y, batch_size=batch_size) * 100.0 / X.shape[0]

    @staticmethod
    def get_batches(X, y, batch_size: int, n: int):
        if batch_size == -1:
            return [(X, y)]
        return ((X[k:k + batch_size], y[k:k + batch_size]) for k in range(0, n, batch_size))
