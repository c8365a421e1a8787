"""One training step of the RPO hot path as the reference trainer performs it
(trainers/rpo.py:290-316: forward, zero_grad, backward, optimizer step) but enqueued as ONE CUDA graph:
forward + CE + prompt-gradient backward (librpo_b200), the gradient exchange between data-parallel
ranks, and the fused SGD update of the two prompt tensors.

Two shapes of the same step (both produce the same training trajectory):

* plain: rpo_forward + rpo_backward, then the gradient all-reduce + SGD.
* class-sharded text tower (`model.shard_text(...)`): the native stages -- text forward |
  image forward | logits forward+backward | text backward | image backward -- with the text stages and
  their two exchanges (all-gather of the text features, reduce-scatter of their gradient) on a side
  stream next to the vision tower.

Exchanges between ranks (world > 1), in order of preference:

* peer memory (rpo_b200/peer.py): librpo_b200's own kernels read / write the peers' buffers over
  NVLink (all-reduce fused into the SGD kernel); plain kernel launches, so the whole step is one graph.
* NCCL through torch.distributed, captured into the graph; if the capture is refused, the collectives
  stay eager between graph segments (`collectives` tells which one is in use).

Used by bench.py and by rpo_b200.trainer.RPO (fast path).  Host code is plumbing only.
"""
import torch

from . import _lib


class StepRunner:
    def __init__(self, model, batch, lr=0.01, momentum=0.9, weight_decay=5e-4, use_graph=True, process_group=None,
                 world_size=1, image_dtype=torch.float32, peer=None):
        """`peer`: None = try peer-memory exchanges when world > 1 and fall back to NCCL, False = NCCL only,
        True = peer memory or raise."""
        self.model = model
        self.B = int(batch)
        self.device = model.w_mm.device
        self.world = int(world_size)
        self.pg = process_group
        self.momentum, self.wd = float(momentum), float(weight_decay)
        self.eng = model.engine(self.B)
        res = model.arch.v_res
        # float32 = what the reference's DataLoader hands over (trainers/rpo.py:318-323); torch.uint8 = raw pixels,
        # normalised inside the patch extraction (a quarter of the host-to-device bytes)
        self.image = torch.zeros(self.B, 3, res, res, dtype=image_dtype, device=self.device)
        self.label = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        self.lr = torch.tensor(float(lr), dtype=torch.float32, device=self.device)
        self._lr_host = float(lr)
        self.first = torch.ones(1, dtype=torch.int32, device=self.device)
        n = self.eng.grad_flat.numel()
        self.mom_buf = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.use_graph = use_graph
        self.launches_per_step = 0
        self.sharded = self.eng.exchange is not None
        self.side = torch.cuda.Stream(self.device) if self.sharded else None
        self.graph = None       # the whole step
        self.segments = None    # fallback: graph segments with eager NCCL calls between them
        self.collectives = "none"
        self.peer = None
        if self.world > 1:
            self.collectives = "nccl"
            if peer is not False:
                from .peer import PeerExchange
                self.peer = PeerExchange.create(self.eng, self.world, process_group, required=bool(peer))
                if self.peer is not None:
                    self.collectives = "peer"

    # -- lr is a device scalar so that a graph replay sees the scheduler's current value ----------
    def set_lr(self, lr):
        lr = float(lr)
        if lr != self._lr_host:
            self._lr_host = lr
            self.lr.fill_(lr)

    # -- enqueue helpers (no host sync) ----------------------------------------------------------
    def _image_code(self):
        return _lib.RPO_U8 if self.image.dtype == torch.uint8 else _lib.dtype_code(self.image.dtype)

    def _update(self):
        """gradient exchange between ranks + SGD(momentum, weight decay) on both prompt tensors"""
        eng, pl, lib = self.eng, self.model.prompt_learner, self.eng.lib
        g = eng.grad_flat
        st = _lib.stream_ptr(self.device)
        code = _lib.dtype_code(self.model.dtype)
        nt = eng.n_text
        if self.peer is not None:
            # one kernel: sum of the ranks' flat gradients read over NVLink (fixed rank order: every replica
            # computes bit-identical sums) + the SGD update of both prompt tensors
            self.peer.allreduce_sgd(pl.text_prompt.data, pl.img_prompt.data, self.mom_buf, self.lr, self.momentum,
                                    self.wd, 1.0 / self.world, self.first)
        else:
            if self.world > 1:
                torch.distributed.all_reduce(g, group=self.pg)  # sum; the mean is folded into grad_scale
            scale = 1.0 / self.world
            _lib.check(lib.rpo_sgd_step(pl.text_prompt.data.data_ptr(), code, g.data_ptr(), self.mom_buf.data_ptr(), nt,
                                        self.lr.data_ptr(), self.momentum, self.wd, scale, self.first.data_ptr(), st))
            _lib.check(lib.rpo_sgd_step(pl.img_prompt.data.data_ptr(), code, g.data_ptr() + 4 * nt,
                                        self.mom_buf.data_ptr() + 4 * nt, g.numel() - nt, self.lr.data_ptr(),
                                        self.momentum, self.wd, scale, self.first.data_ptr(), st))
        self.first.zero_()

    def _stage_fns(self):
        eng, pl = self.eng, self.model.prompt_learner

        def logits():
            eng.logits_forward(self.label, None)
            eng.logits_backward()

        return {"text_fwd": lambda: eng.text_forward(pl.text_prompt.data),
                "image_fwd": lambda: eng.image_forward(self.image, self._image_code(), pl.img_prompt.data),
                "logits": logits, "text_bwd": eng.text_backward, "image_bwd": eng.image_backward}

    @staticmethod
    def _run(seg):
        seg.replay() if isinstance(seg, torch.cuda.CUDAGraph) else seg()

    def _gather(self):
        if self.peer is not None:
            self.peer.gather_text_features()
        else:
            self.eng.exchange.gather_text_features()

    def _scatter(self):
        if self.peer is not None:
            self.peer.scatter_text_grads()
        else:
            self.eng.exchange.scatter_text_grads()

    def _chain(self, segs):
        """the step up to the flat gradient; sharded: text stages and their exchanges on the side stream"""
        if not self.sharded:
            pl = self.model.prompt_learner
            self.eng.forward(self.image, pl.text_prompt.data, pl.img_prompt.data, self.label)
            self.eng.backward()
            return
        main, side = torch.cuda.current_stream(self.device), self.side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._run(segs["text_fwd"])
            self._gather()
        self._run(segs["image_fwd"])
        main.wait_stream(side)
        self._run(segs["logits"])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._scatter()
            self._run(segs["text_bwd"])
        self._run(segs["image_bwd"])
        main.wait_stream(side)

    def _enqueue(self, segs=None):
        self._chain(segs if segs is not None else (self._stage_fns() if self.sharded else None))
        self._update()

    # -- warm-up and capture ------------------------------------------------------------------------
    def _snapshot(self):
        pl = self.model.prompt_learner
        return pl.text_prompt.data.clone(), pl.img_prompt.data.clone()

    def _restore(self, snap):
        pl = self.model.prompt_learner
        pl.text_prompt.data.copy_(snap[0])
        pl.img_prompt.data.copy_(snap[1])
        self.mom_buf.zero_()
        self.first.fill_(1)

    def _try_capture(self, fn):
        g = torch.cuda.CUDAGraph()
        try:
            # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fn()
            return g
        except Exception as e:  # capture of a collective refused: the caller falls back to segments
            self.capture_error = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize(self.device)
            return None

    def prepare(self, warmup=3):
        """Warm-up (sets kernel attributes, loads modules, opens the NCCL channels) and CUDA-graph capture of the
        step.  The warm-up steps run on whatever `self.image` / `self.label` hold and DO update the prompts; the
        prompts, the momentum buffer and the first-step flag are put back afterwards, so `step()` number one starts
        from the state the caller handed over."""
        with torch.cuda.device(self.device):
            snap = self._snapshot()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(max(1, warmup)):
                    self._enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.launches_per_step = self.eng.launch_count() + (1 if self.peer is not None else 2) + \
                (2 if self.peer is not None and self.sharded else 0)
            if self.use_graph:
                own_kernels_only = self.world == 1 or self.peer is not None
                # One graph for the whole step only when every node is a kernel of librpo_b200.  Capturing NCCL's
                # collectives into the step graph hung on 2 x B200 (NCCL 2.28.9, torch 2.11: both ranks stall in
                # the first replay, gpurun_out/n2/mode_*_False.log); the NCCL fallback therefore keeps the
                # collectives eager between graph segments, as round 1 ran it on 2 / 4 / 8 GPUs.
                self.graph = self._try_capture(self._enqueue) if own_kernels_only else None
                if self.graph is None and own_kernels_only:
                    raise _lib.RpoError(f"CUDA-graph capture of the step failed: {self.capture_error}")
                if self.graph is None:
                    # eager NCCL calls between graph segments
                    self.collectives = "nccl-eager"
                    if self.sharded:
                        segs = {}
                        for name, fn in self._stage_fns().items():
                            g = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g):
                                fn()
                            segs[name] = g
                        self.segments = segs
                    else:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._chain(None)
                        self.segments = g
            self._restore(snap)
            torch.cuda.synchronize()
        return self

    def step(self):
        """Enqueues one step on the current stream.  Inputs are whatever self.image / self.label hold."""
        self.model.invalidate_text_features()  # the fused SGD kernel rewrites the prompts in place
        if self.graph is not None:
            self.graph.replay()
        elif self.segments is not None:
            if isinstance(self.segments, torch.cuda.CUDAGraph):
                self.segments.replay()
            else:
                self._chain(self.segments)
            self._update()
        else:
            self._enqueue()

    @property
    def loss(self):
        return self.eng.loss
