"""One training step of the RPO hot path as the reference trainer performs it
(trainers/rpo.py:290-316: forward, zero_grad, backward, optimizer step) but enqueued as one CUDA
graph: forward + CE + prompt-gradient backward (librpo_b200), an NCCL all-reduce of the flat
[K*Dt + K*Dv] f32 gradient when data-parallel, and the fused SGD update of the two prompt tensors.

With a class-sharded text tower (`model.shard_text(...)`) the step is five graph
segments -- text forward | image forward | logits forward+backward | text backward | image backward --
with the text segments and their two collectives (all-gather of the text features, reduce-scatter of
their gradient) on a side stream next to the vision tower.

Used by bench.py and by rpo_b200.trainer.RPO.forward_backward.  Host code is plumbing only.
"""
import torch

from . import _lib


class StepRunner:
    def __init__(self, model, batch, lr=0.01, momentum=0.9, weight_decay=5e-4, use_graph=True, process_group=None,
                 world_size=1, image_dtype=torch.float32):
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
        self.first = torch.ones(1, dtype=torch.int32, device=self.device)
        n = self.eng.grad_flat.numel()
        self.mom_buf = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = 0
        self.sharded = self.eng.exchange is not None
        self.side = torch.cuda.Stream(self.device) if self.sharded else None
        self.segments = None

    # -- enqueue helpers (no host sync) ----------------------------------------------------------
    def _fwd_bwd(self):
        pl = self.model.prompt_learner
        self.eng.forward(self.image, pl.text_prompt.data, pl.img_prompt.data, self.label)
        self.eng.backward()

    def _update(self):
        eng, pl, lib = self.eng, self.model.prompt_learner, self.eng.lib
        g = eng.grad_flat
        if self.world > 1:
            torch.distributed.all_reduce(g, group=self.pg)  # sum; the mean is folded into grad_scale
        st = _lib.stream_ptr(self.device)
        code = _lib.dtype_code(self.model.dtype)
        scale = 1.0 / self.world
        nt = eng.n_text
        _lib.check(lib.rpo_sgd_step(pl.text_prompt.data.data_ptr(), code, g.data_ptr(), self.mom_buf.data_ptr(), nt,
                                    self.lr.data_ptr(), self.momentum, self.wd, scale, self.first.data_ptr(), st))
        _lib.check(lib.rpo_sgd_step(pl.img_prompt.data.data_ptr(), code, g.data_ptr() + 4 * nt,
                                    self.mom_buf.data_ptr() + 4 * nt, g.numel() - nt, self.lr.data_ptr(), self.momentum,
                                    self.wd, scale, self.first.data_ptr(), st))
        self.first.zero_()

    def _enqueue(self):
        self._fwd_bwd()
        self._update()

    # -- class-sharded text tower: stage segments -------------------------------------------------
    def _segment_fns(self):
        eng, pl = self.eng, self.model.prompt_learner
        idt = _lib.RPO_U8 if self.image.dtype == torch.uint8 else _lib.dtype_code(self.image.dtype)

        def logits():
            eng.logits_forward(self.label, None)
            eng.logits_backward()

        return {"text_fwd": lambda: eng.text_forward(pl.text_prompt.data),
                "image_fwd": lambda: eng.image_forward(self.image, idt, pl.img_prompt.data),
                "logits": logits, "text_bwd": eng.text_backward, "image_bwd": eng.image_backward}

    def _run(self, name):
        seg = self.segments[name]
        seg.replay() if isinstance(seg, torch.cuda.CUDAGraph) else seg()

    def _sharded_fwd_bwd(self):
        """text stages + collectives on the side stream, vision stages on the current stream"""
        ex, main, side = self.eng.exchange, torch.cuda.current_stream(self.device), self.side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._run("text_fwd")
            ex.gather_text_features()
        self._run("image_fwd")
        main.wait_stream(side)
        self._run("logits")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ex.scatter_text_grads()
            self._run("text_bwd")
        self._run("image_bwd")
        main.wait_stream(side)

    def prepare(self, warmup=3):
        """Warm-up (sets kernel attributes, loads modules) and CUDA-graph capture of the step."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(max(1, warmup)):
                    self._enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.launches_per_step = self.eng.launch_count() + 2  # + two SGD kernels
            if self.sharded:
                self.segments = self._segment_fns()
                if self.use_graph:
                    for name, fn in list(self.segments.items()):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            fn()
                        self.segments[name] = g
                torch.cuda.synchronize()
            elif self.use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    if self.world > 1:
                        self._fwd_bwd()
                    else:
                        self._enqueue()
                self.graph = g
        return self

    def step(self):
        """Enqueues one step on the current stream.  Inputs are whatever self.image / self.label hold."""
        self.model.invalidate_text_features()  # the fused SGD kernel rewrites the prompts in place
        if self.sharded:
            self._sharded_fwd_bwd()
            self._update()
        elif self.graph is not None:
            self.graph.replay()
            if self.world > 1:
                self._update()
        else:
            self._enqueue()

    @property
    def loss(self):
        return self.eng.loss
