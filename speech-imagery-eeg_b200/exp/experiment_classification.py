"""Classification experiment with the reference's surface (Experiment(args).train()/validation()/test(),
exp/experiment_classification.py:85-1138) on the B200-native shapelet path.

Differences from the reference that are deliberate:
  * multi-GPU is one process per GPU + flat NCCL gradient all-reduce (exp/parallel.py) instead of
    nn.DataParallel; `--multi_gpu` is accepted and means "use WORLD_SIZE ranks if launched by torchrun"
  * the per-step `loss.item()` host sync (:343) is replaced by a device-side running sum read once per epoch
  * batches go host->device through pinned memory on a side stream, one batch ahead (DevicePrefetcher)
Loss, optimiser, beta schedule, early stopping and checkpoint keys are the reference's.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

from data_provider.data_factory import data_provider
from exp.parallel import DevicePrefetcher, FlatGradAllReduce, init_distributed
from models.FullyConvNet import FullyConvNetwork
from models.InterpGN import InterpGN, dnn_dict
from models.Shapelet import DistThresholdSBM, ShapeBottleneckModel
from utils.shapelet_util import ClassificationResult
from utils.tools import EarlyStopping, convert_to_hms


def compute_beta(epoch, max_epoch, schedule='cosine'):
    """Weight of the auxiliary shapelet-expert CE (reference :19-26)."""
    if schedule == 'cosine':
        return 0.5 * (1 + np.cos(np.pi * epoch / max_epoch))
    if schedule == 'linear':
        return 1 - epoch / max_epoch
    return 1


def compute_shapelet_score(shapelet_distances, cls_weights, y_pred, y_true):
    """Mean class-relevant shapelet score over correctly classified samples (reference :29-34)."""
    score = shapelet_distances @ nn.functional.relu(cls_weights.T) / shapelet_distances.shape[-1]
    ok = y_pred == y_true
    return score[ok].gather(-1, y_true[ok].unsqueeze(1)).mean().item()


def get_dnn_model(configs):
    return dnn_dict[configs.dnn_type](configs)


class Experiment(object):
    model_dict = {'InterpGN': InterpGN, 'SBM': ShapeBottleneckModel, 'LTS': DistThresholdSBM, 'DNN': get_dnn_model}

    def __init__(self, args, load_data=True, loaders=None, independent=False):
        """load_data=False (bench.py) skips the loaders; args.seq_len/enc_in/num_class must then be set.
        loaders=((train_ds, train_dl), (val_ds, val_dl), (test_ds, test_dl)) injects prepared splits (LOSO folds);
        independent=True trains this rank's own model: no gradient all-reduce, no parameter broadcast, every rank
        logs and checkpoints (exp/loso.py)."""
        self.args = args
        self.rank, self.local_rank, self.world = init_distributed()
        self.independent = independent
        args.rank, args.world_size = (0, 1) if independent else (self.rank, self.world)
        self.is_main = self.rank == 0 or independent
        if not torch.cuda.is_available():
            raise RuntimeError("the ign_b200 training path needs a CUDA device (no CPU fallback)")
        self.device = torch.device('cuda', self.local_rank)
        torch.cuda.set_device(self.device)

        if loaders is not None:
            (self.train_data, self.train_loader), (self.val_data, self.val_loader), \
                (self.test_data, self.test_loader) = loaders
            self._get_params_from_data()
        elif load_data:
            self.train_data, self.train_loader = data_provider(args, flag="train")
            self.val_data, self.val_loader = data_provider(args, flag="val")
            self.test_data, self.test_loader = data_provider(args, flag="test")
            self._get_params_from_data()
        self.model = self._build_model().to(self.device)
        # same Adam(lr) as the reference (:283); the fused CUDA implementation is one kernel per step (capturable: its
        # step counter lives on the device, so the whole step can be replayed from a CUDA graph — --cuda_graph)
        self.use_graph = bool(getattr(args, "cuda_graph", False))
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.args.lr, fused=True, capturable=self.use_graph)
        self._graph = None         # (CUDAGraph, static x, y, mask, loss, beta) once captured
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(self.optimizer, T_0=self.args.train_epochs)
        self.grads = FlatGradAllReduce(self.model, 1 if independent else self.world)
        a = self.args
        self.checkpoint_dir = "./checkpoints/{}/{}/dnn-{}_seed-{}_k-{}_div-{}_reg-{}_eps-{}_beta-{}_dfunc-{}_cls-{}".format(
            a.model, a.dataset + ("-synthetic" if getattr(a, "data_source", getattr(a, "data", "")) == "synthetic" else ""),
            a.dnn_type, a.seed, a.num_shapelet, a.lambda_div, a.lambda_reg, a.epsilon,
            a.beta_schedule, a.distance_func, a.sbm_cls)
        if self.is_main and load_data:
            os.makedirs(self.checkpoint_dir, exist_ok=True)
        self.epoch_stop = 0
        self.history = []          # (epoch, train_loss, val_loss, val_acc) per finished epoch of train()

    def log(self, *a):
        if self.is_main:
            print(*a)
            sys.stdout.flush()

    def _get_params_from_data(self):
        """seq_len / enc_in / num_class are injected from the dataset (reference :166-249)."""
        ds = self.train_data
        # one padded length for every split: the longest series of train / val / test (variable-length UEA archives
        # such as JapaneseVowels have longer TEST series; the collate functions read args.seq_len per batch)
        self.args.seq_len = max(int(getattr(d, "max_seq_len", getattr(d, "seq_len", 845)))
                                for d in (self.train_data, self.val_data, self.test_data))
        self.args.enc_in = int(getattr(ds, "enc_in", 122))
        self.args.num_class = int(getattr(ds, "num_class", 3))
        self.args.pred_len, self.args.label_len = 0, 0

    def _build_model(self):
        a = self.args
        if a.model not in self.model_dict:
            raise ValueError(f"model {a.model!r} not available here: {sorted(self.model_dict)}")
        if a.model in ('SBM', 'LTS'):       # reference :264-269
            fr = [0.05, 0.1, 0.2, 0.3, 0.5, 0.8]
            return self.model_dict[a.model](configs=a, num_shapelet=[a.num_shapelet] * len(fr), shapelet_len=fr)
        if a.model == 'DNN':
            return get_dnn_model(a)
        return self.model_dict[a.model](a)  # InterpGN: constructor defaults K=[5]*4 (reference :276-277)

    def print_args(self):
        if self.is_main:
            for k in sorted(vars(self.args)):
                print(f"  {k}: {getattr(self.args, k)}")

    # ------------------------------------------------------------------ one optimisation step
    def _to_device(self, batch_x, label, padding_mask):
        x = batch_x.float().to(self.device, non_blocking=True)
        y = label.long().squeeze(-1).to(self.device, non_blocking=True)
        m = padding_mask.float().to(self.device, non_blocking=True)
        return x, y, m

    def _loss(self, x, y, mask, epoch, reduction='mean', gating_value=None, train=True, beta=None):
        a = self.args
        with torch.autocast(device_type='cuda', dtype=torch.bfloat16, enabled=a.amp):
            if a.model == 'DNN':
                logits = self.model(x, mask, None, None)
                return nn.functional.cross_entropy(logits, y, reduction=reduction), logits, None
            if a.model == 'InterpGN' and not train:
                logits, info = self.model(x, mask, None, None, gating_value=gating_value)
            else:
                logits, info = self.model(x, mask, None, None)
            loss = nn.functional.cross_entropy(logits, y, reduction=reduction) + info.loss.mean()
            if train and a.model == 'InterpGN':
                if beta is None:            # (a device scalar under graph replay: refreshed per epoch without re-capture)
                    beta = compute_beta(epoch, a.train_epochs, a.beta_schedule)
                loss = loss + beta * nn.functional.cross_entropy(info.shapelet_preds, y)
        return loss, logits, info

    def train_step(self, x, y, mask, epoch, step_index):
        """fwd + bwd (+ all-reduce + Adam on accumulation boundaries).  Returns the detached loss (device).
        With --cuda_graph, full-size batches replay the whole step (forward, backward, Adam, weight clamp, gradient
        reset: ~150 kernel launches at config 2, ~130 at config 1) from ONE captured CUDA graph — small configurations
        such as the UEA archives are bound by launch latency, not by the GPU (3.0 ms per step at JapaneseVowels shape)."""
        if self.use_graph and self._graph_applicable(x):
            return self._graphed_step(x, y, mask, epoch, step_index)
        return self._eager_step(x, y, mask, epoch, step_index)

    def _graph_applicable(self, x):
        a = self.args
        if a.gradient_accumulation_steps != 1 or a.gradient_clip > 0 or (self.world > 1 and not self.independent) \
                or getattr(a, "lr_decay", False):
            return False            # accumulation / clipping / the NCCL exchange keep the eager path; so does a learning-
                                    # rate schedule (the captured Adam kernel carries the rate it was captured with)
        return self._graph is None or tuple(x.shape) == tuple(self._graph[1].shape)     # ragged last batch: eager

    def _graphed_step(self, x, y, mask, epoch, step_index):
        a = self.args
        beta = float(compute_beta(epoch, a.train_epochs, a.beta_schedule))
        if self._graph is None:
            sx, sy, sm = x.clone(), y.clone(), mask.clone()
            sbeta = torch.full((), beta, device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                   # warm-up on a side stream, as graph capture requires
                for _ in range(3):
                    self._eager_step(sx, sy, sm, epoch, step_index, beta=sbeta)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                sloss = self._eager_step(sx, sy, sm, epoch, step_index, beta=sbeta)
            self._graph = (graph, sx, sy, sm, sloss, sbeta)
            # (the three warm-up steps are real optimisation steps on this batch; the capture pass only records)
        graph, sx, sy, sm, sloss, sbeta = self._graph
        sx.copy_(x, non_blocking=True); sy.copy_(y, non_blocking=True); sm.copy_(mask, non_blocking=True)
        sbeta.fill_(beta)
        graph.replay()
        return sloss.clone()

    def _eager_step(self, x, y, mask, epoch, step_index, beta=None):
        a = self.args
        boundary = step_index % a.gradient_accumulation_steps == 0
        if boundary:
            self.grads.arm()
        loss, _, _ = self._loss(x, y, mask, epoch, beta=beta)
        if a.gradient_accumulation_steps > 1:
            loss = loss / a.gradient_accumulation_steps
        loss.backward()
        if boundary:
            self.grads.finish()
            if a.gradient_clip > 0:
                nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=a.gradient_clip)
            self.optimizer.step()
            if a.pos_weight:
                self.model.step()
            self.grads.zero_grad()
        return loss.detach()

    def train(self):
        a = self.args
        torch.set_float32_matmul_precision('medium')     # reference :297
        t0 = time.time()
        stopper = EarlyStopping(patience=a.patience, verbose=self.is_main, delta=0, is_main=self.is_main)
        step = 0
        self.grads.zero_grad()
        for epoch in range(a.train_epochs):
            self.model.train()
            if len(self.train_loader) == 0:
                continue
            run = torch.zeros((), device=self.device)
            nstep = 0
            for x, y, m in DevicePrefetcher(self.train_loader, self.device):
                step += 1
                run += self.train_step(x, y, m, epoch, step)
                nstep += 1
            train_loss = float(run) / max(1, nstep)       # one host sync per epoch
            val_loss, val_acc = self.validation()
            self.history.append((epoch, train_loss, val_loss, val_acc))
            remain = (time.time() - t0) * (a.train_epochs - epoch) / (epoch + 1)
            if (epoch + 1) % a.log_interval == 0:
                self.log(f"Epoch {epoch + 1}/{a.train_epochs} | Train Loss {train_loss:.4f} | Val Loss {val_loss:.4f} "
                         f"| Val Acc {val_acc:.4f} | Time Rem {convert_to_hms(remain)}")
            if a.lr_decay:
                self.scheduler.step()
            if epoch >= a.min_epochs:
                stopper(-val_acc, self.model, self.checkpoint_dir)
            self.epoch_stop = epoch
            if stopper.early_stop:
                self.log("Early stopping")
                break
        if self.world > 1 and not self.independent:
            torch.distributed.barrier()
        best = os.path.join(self.checkpoint_dir, 'checkpoint.pth')
        if os.path.exists(best):
            self.model.load_state_dict(torch.load(best, map_location=self.device))
        return self.model

    # ------------------------------------------------------------------ evaluation
    def _evaluate(self, loader, gating_value=None, keep=False):
        losses, preds, trues = [], [], []
        extra = {k: [] for k in ("x", "p", "d", "eta", "shapelet_preds", "dnn_preds")}
        self.model.eval()
        with torch.no_grad():
            for batch in loader:
                x, y, m = self._to_device(*batch)
                loss, logits, info = self._loss(x, y, m, 0, reduction='none', gating_value=gating_value, train=False)
                losses.append(loss.flatten().float()); preds.append(logits.float()); trues.append(y)
                if keep:
                    extra["x"].append(x.cpu())
                    if info is not None:
                        for k in ("p", "d", "eta", "shapelet_preds", "dnn_preds"):
                            v = getattr(info, k)
                            if v is not None:
                                extra[k].append(v.float().cpu())
        self.model.train()
        if not losses:
            return float('inf'), 0.0, None, None, extra
        loss = torch.cat(losses).mean().item()
        preds, trues = torch.cat(preds), torch.cat(trues).flatten()
        pred_cls = torch.softmax(preds, dim=1).argmax(dim=1)
        acc = float((pred_cls == trues).float().mean())
        return loss, acc, pred_cls.cpu(), trues.cpu(), extra

    def validation(self):
        loss, acc, _, _, _ = self._evaluate(self.val_loader)
        return loss, acc

    def test(self, save_csv=True, result_dir=None):
        """Same forward with the test-time hard gate (reference :974).  Returns (loss, ClassificationResult, None)."""
        loss, acc, pred_cls, trues, extra = self._evaluate(self.test_loader, gating_value=self.args.gating_value, keep=True)
        if pred_cls is None:
            return float('inf'), None, None
        cat = lambda k: torch.cat(extra[k], dim=0) if extra[k] else None
        res = ClassificationResult(x_data=cat("x"), trues=trues, preds=pred_cls, loss=loss, accuracy=acc,
                                   p=cat("p"), d=cat("d"), shapelet_preds=cat("shapelet_preds"),
                                   eta=cat("eta"), dnn_preds=cat("dnn_preds"))
        if self.args.model != 'DNN':
            sbm = self.model.sbm if self.args.model == 'InterpGN' else self.model
            res.w = sbm.output_layer.weight.detach().cpu()
            res.shapelets = sbm.get_shapelets()
        n_cls = int(self.args.num_class)
        self.log(f"Test loss {loss:.4f} | accuracy {acc * 100:.2f}% | random baseline {100.0 / n_cls:.2f}% "
                 f"| samples {len(trues)}")
        return loss, res, None
