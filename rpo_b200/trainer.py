"""Host-side mirror of the reference's Dassl trainer plugin (trainers/rpo.py:24-39, 235-357): the class
`RPO` with `check_cfg`, `build_model`, `forward_backward`, `parse_batch_train` and `load_model`,
registered into Dassl's TRAINER_REGISTRY under the same name, so the reference's `train.py`
(`import trainers.rpo` at train.py:31, `build_trainer(cfg)` at :163) runs unchanged once
`trainers/rpo.py` re-exports this module (see INTEGRATION.md and integration/rpo.py).

Only the model behind `self.model` changes: `rpo_b200.model.CustomCLIP` (sm_100a kernels through
the C ABI) instead of the PyTorch-op implementation.  The trainer itself stays host glue, as in the
reference: Dassl builds the optimizer and scheduler over `prompt_learner`, `loss.backward()` reaches
the two prompt Parameters through `rpo_b200.model._RpoLoss`, `optim.step()` is torch's.

Deliberate differences from trainers/rpo.py (none changes a result):
  * `nn.DataParallel` (:282-285) is not used.  It cannot train RPO (per-replica scalar losses are
    gathered into a vector and `backward()` has no `.mean()`).  Multi-GPU is one
    process per GPU: if `torch.distributed` is initialised the prompt gradients are averaged with
    one all-reduce of a flat f32 buffer before `optim.step()`.
  * `torch.autograd.set_detect_anomaly(True)` (:288) is not switched on (a debugging aid that halves
    throughput); set RPO_B200_DETECT_ANOMALY=1 to get it back.
  * RPO_B200_SHARD_TEXT=1 with torch.distributed initialised: each rank runs the text tower for its
    ceil(n_cls / world) class prompts only (`CustomCLIP.shard_text`, rpo_b200/text_shard.py); text features
    are all-gathered, their gradient reduce-scattered, results unchanged.  Every rank must then call the
    model the same number of times (training steps and evaluation batches alike).
  * PREC="amp": the reference keeps fp32 weights and autocasts the matmuls; here the fp32 engine runs
    (at least as precise as autocast) and the GradScaler is kept so the control flow is identical.
  * Fast path (default; RPO_B200_FAST=0 switches it off).  With a plain torch SGD over the two prompt tensors and
    PREC fp16 / fp32, `forward_backward` does not go through autograd and `optim.step()` at all: the batch goes
    through `input_pipeline.BatchUploader` (pinned, double-buffered, asynchronous upload instead of the blocking
    `.to(device)` of :318-323) and the whole step -- forward, CE, prompt-gradient backward, gradient exchange
    between ranks, SGD(momentum, weight decay) -- is ONE CUDA-graph replay (`runner.StepRunner`).  The learning rate
    is read from `optim.param_groups` every step (Dassl's scheduler keeps driving it) and fed to the graph through
    a device scalar; the momentum lives in the runner and is exported into `optim.state` before checkpoints are
    written.  The returned loss is read back asynchronously: step n reports the loss of step n-1 (RPO_B200_SYNC_LOSS=1
    restores the reference's blocking `loss.item()` of the current step).

Dassl is imported lazily: without it (this repository's test environment) the class is still
importable over `object` so that its methods can be exercised with a stand-in base.
"""
import os
import os.path as osp

import torch

from .model import CustomCLIP, PromptLearner  # noqa: F401  (re-exported: the reference module defines both)

try:  # pragma: no cover - Dassl is not installed in the build container
    from dassl.engine import TRAINER_REGISTRY, TrainerX
    from dassl.optim import build_lr_scheduler, build_optimizer
    from dassl.utils import load_checkpoint, load_pretrained_weights
    HAVE_DASSL = True
except Exception:  # ImportError or a broken partial install
    TRAINER_REGISTRY = None
    TrainerX = object
    build_lr_scheduler = build_optimizer = load_checkpoint = load_pretrained_weights = None
    HAVE_DASSL = False


def load_clip_to_cpu(cfg):
    """trainers/rpo.py:24-39 -- needs the host checkout's `clip` package (download + build_model are
    init-time host code and stay the reference's)."""
    from clip import clip
    backbone_name = cfg.MODEL.BACKBONE.NAME
    model_path = clip._download(clip._MODELS[backbone_name])
    try:
        jit = torch.jit.load(model_path, map_location="cpu").eval()
        state_dict = jit.state_dict()
    except RuntimeError:
        state_dict = torch.load(model_path, map_location="cpu")
    return clip.build_model(state_dict)


def allreduce_mean_(params, group=None):
    """Averages the gradients of `params` over the process group with ONE all-reduce of a flat f32
    buffer.  No-op when torch.distributed is not initialised or world size is 1.
    Returns the world size used."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return world
    flat = torch.cat([g.detach().reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g).to(g.dtype))
        off += n
    return world


def maybe_shard_text(model, group=None):
    """Class-shards the text tower over the process group when RPO_B200_SHARD_TEXT=1 and there is more than one
    rank (and at least one class per rank).  Returns the shard or None."""
    import torch.distributed as dist
    if os.environ.get("RPO_B200_SHARD_TEXT") != "1" or not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_world_size(group) == 1:
        return None
    from .text_shard import ClassShard
    n_cls, world = model.text_x.shape[0], dist.get_world_size(group)
    if not ClassShard.feasible(n_cls, world):  # same answer on every rank (depends on n_cls and world only)
        print(f"text tower stays replicated: {n_cls} classes do not split over {world} ranks without an empty part")
        return None
    return model.shard_text(group=group)


class RPO(TrainerX):
    def check_cfg(self, cfg):
        assert cfg.TRAINER.RPO.PREC in ["fp16", "fp32", "amp"]  # trainers/rpo.py:238

    def build_model(self):
        cfg = self.cfg
        classnames = self.dm.dataset.classnames
        print(f"Loading CLIP (backbone: {cfg.MODEL.BACKBONE.NAME})")
        clip_model = load_clip_to_cpu(cfg)
        if cfg.TRAINER.RPO.PREC in ("fp32", "amp"):
            clip_model.float()  # CLIP's default precision is fp16 (trainers/rpo.py:247-249)
        print("Building custom CLIP (rpo_b200: sm_100a kernels)")
        self.model = CustomCLIP(cfg, classnames, cfg.DATASET.PROMPT, clip_model)
        for name, param in self.model.named_parameters():  # :258-260
            if "prompt_learner" not in name:
                param.requires_grad_(False)
        enabled = {n for n, p in self.model.named_parameters() if p.requires_grad}
        print(f"Parameters to be updated: {enabled}")
        if cfg.MODEL.INIT_WEIGHTS:
            load_pretrained_weights(self.model.prompt_learner, cfg.MODEL.INIT_WEIGHTS)
        self.model.to(self.device)
        maybe_shard_text(self.model)
        # only the prompt learner goes to the optimizer (:274-276)
        self.optim = build_optimizer(self.model.prompt_learner, cfg.OPTIM)
        self.sched = build_lr_scheduler(self.optim, cfg.OPTIM)
        self.register_model("prompt_learner", self.model.prompt_learner, self.optim, self.sched)
        self.scaler = torch.amp.GradScaler("cuda") if cfg.TRAINER.RPO.PREC == "amp" else None
        if os.environ.get("RPO_B200_DETECT_ANOMALY") == "1":
            torch.autograd.set_detect_anomaly(True)

    # ---- fast path: uploader + one CUDA-graph replay per step (see the module text) -----------------------------
    def fast_path_available(self):
        """The fused step implements exactly torch.optim.SGD(momentum, dampening 0, L2 weight decay, no nesterov)
        over the two prompt tensors; anything else keeps the reference-shaped autograd path."""
        if os.environ.get("RPO_B200_FAST", "1") == "0" or getattr(self, "scaler", None) is not None:
            return False
        optim, model = self.optim, self.model
        if type(optim) is not torch.optim.SGD or len(optim.param_groups) != 1:
            return False
        g = optim.param_groups[0]
        pl = model.prompt_learner
        if [id(p) for p in g["params"]] != [id(pl.text_prompt), id(pl.img_prompt)]:
            return False
        if g.get("nesterov") or g.get("dampening", 0) != 0 or g.get("maximize") or not pl.text_prompt.is_cuda:
            return False
        return os.environ.get("RPO_B200_DETECT_ANOMALY") != "1"

    def _fast_state(self, batch_size, image_dtype):
        st = getattr(self, "_fast", None)
        if st is not None and st["B"] == batch_size and st["dtype"] == image_dtype:
            return st
        import torch.distributed as dist
        from .input_pipeline import BatchUploader, LossReader
        from .runner import StepRunner
        g = self.optim.param_groups[0]
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        model = self.model
        old = st["runner"] if st is not None else None
        runner = StepRunner(model, batch_size, lr=g["lr"], momentum=g.get("momentum", 0.0),
                            weight_decay=g.get("weight_decay", 0.0), world_size=world, image_dtype=image_dtype)
        runner.prepare(warmup=2)
        if old is not None:  # a different batch size mid-run: carry the optimiser state over
            runner.mom_buf.copy_(old.mom_buf)
            runner.first.copy_(old.first)
        else:
            self._import_momentum(runner)
        st = {"B": batch_size, "dtype": image_dtype, "runner": runner,
              "uploader": BatchUploader(model.w_mm.device, batch_size, model.arch.v_res, image_dtype),
              "loss": LossReader(model.w_mm.device)}
        self._fast = st
        return st

    def _momentum_views(self, runner):
        pl = self.model.prompt_learner
        nt = runner.eng.n_text
        return [(pl.text_prompt, runner.mom_buf[:nt]), (pl.img_prompt, runner.mom_buf[nt:])]

    def _import_momentum(self, runner):
        """resuming from a checkpoint whose optimiser state holds momentum buffers"""
        have = False
        for p, view in self._momentum_views(runner):
            buf = self.optim.state.get(p, {}).get("momentum_buffer")
            if buf is not None:
                view.copy_(buf.detach().reshape(-1).float())
                have = True
        if have:
            runner.first.zero_()

    def export_optimizer_state(self):
        """Writes the runner's f32 momentum into `optim.state` (torch's layout), so that Dassl's save_model /
        a later switch to the autograd path see the optimiser they expect."""
        st = getattr(self, "_fast", None)
        if st is None or int(st["runner"].first.item()) == 1:
            return
        for p, view in self._momentum_views(st["runner"]):
            self.optim.state.setdefault(p, {})["momentum_buffer"] = view.view_as(p).to(p.dtype).clone()

    def save_model(self, *args, **kwargs):  # Dassl TrainerBase.save_model: checkpoints carry optim.state_dict()
        self.export_optimizer_state()
        return super().save_model(*args, **kwargs)

    def _forward_backward_fast(self, batch):
        image, label = batch["img"], batch["label"]
        st = self._fast_state(image.shape[0], image.dtype)
        runner, up, lr_ = st["runner"], st["uploader"], st["loss"]
        if image.is_cuda:  # already on the device (a custom loader): no staging
            runner.image.copy_(image, non_blocking=True)
            runner.label.copy_(label, non_blocking=True)
        else:
            ticket = up.submit(image, label)
            img_d, lab_d = up.acquire(ticket)
            runner.image.copy_(img_d, non_blocking=True)
            runner.label.copy_(lab_d, non_blocking=True)
            up.release(ticket)
        runner.set_lr(self.optim.param_groups[0]["lr"])
        runner.step()
        lr_.push(runner.loss)
        lag = 0 if os.environ.get("RPO_B200_SYNC_LOSS") == "1" else 1
        loss_summary = {"loss": lr_.latest(lag)}
        if (self.batch_idx + 1) == self.num_batches:
            self.update_lr()
        return loss_summary

    def forward_backward(self, batch):
        if self.fast_path_available():
            return self._forward_backward_fast(batch)
        image, label = self.parse_batch_train(batch)
        model, optim, scaler = self.model, self.optim, self.scaler
        loss = model(image, label)
        optim.zero_grad()
        if scaler is not None:
            scaler.scale(loss).backward()
            allreduce_mean_(list(model.prompt_learner.parameters()))
            scaler.step(optim)
            scaler.update()
        else:
            loss.backward()
            allreduce_mean_(list(model.prompt_learner.parameters()))
            optim.step()
        loss_summary = {"loss": loss.item()}
        if (self.batch_idx + 1) == self.num_batches:
            self.update_lr()
        return loss_summary

    def parse_batch_train(self, batch):
        return batch["img"].to(self.device), batch["label"].to(self.device)

    def load_model(self, directory, epoch=None):
        if not directory:
            print("Note that load_model() is skipped as no pretrained model is given")
            return
        names = self.get_model_names()
        model_file = "model-best.pth.tar" if epoch is None else "model.pth.tar-" + str(epoch)
        loader = load_checkpoint or (lambda p: torch.load(p, map_location="cpu"))
        for name in names:
            model_path = osp.join(directory, name, model_file)
            if not osp.exists(model_path):
                raise FileNotFoundError('Model not found at "{}"'.format(model_path))
            checkpoint = loader(model_path)
            state_dict = checkpoint["state_dict"]
            epoch = checkpoint["epoch"]
            for legacy in ("token_prefix", "token_suffix"):  # CoOp leftovers (:348-352)
                state_dict.pop(legacy, None)
            print('Loading weights to {} from "{}" (epoch = {})'.format(name, model_path, epoch))
            self._models[name].load_state_dict(state_dict, strict=False)


if HAVE_DASSL:  # pragma: no cover
    RPO = TRAINER_REGISTRY.register()(RPO)
