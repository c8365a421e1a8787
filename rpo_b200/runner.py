"""One training step of the RPO hot path as the reference trainer performs it
(trainers/rpo.py:290-316: forward, zero_grad, backward, optimizer step) but enqueued as CUDA graphs:
forward + CE + prompt-gradient backward (librpo_b200), an NCCL all-reduce of the flat
[K*Dt + K*Dv] f32 gradient when data-parallel, and the fused SGD update of the two prompt tensors.

Three shapes of the same step (all produce the same training trajectory):

* plain: rpo_forward + rpo_backward (+ SGD) in ONE graph.
* class-sharded text tower (`model.shard_text(...)`): the native stages -- text forward |
  image forward | logits forward+backward | text backward | image backward -- as graph segments with the
  text segments and their two collectives (all-gather of the text features, reduce-scatter of their
  gradient) on a side stream next to the vision tower.
* pipelined (`pipeline=True`): the context rows (cls + patches) of the vision tower never depend on the
  prompts (visual_mask hides the prompt columns, trainers/rpo.py:155-156), only the K prompt rows per
  image do.  `step()` therefore runs the context rows of the batch just handed over on a second stream
  into one of two activation slots, while the prompt-dependent chain of the PREVIOUS batch (text
  prompt rows, image prompt rows, logits, CE, both backwards, all-reduce, SGD) runs on the current
  stream over the other slot.  The chain is a long sequence of small latency-bound kernels, the
  context pass a short sequence of big tensor-core kernels: together they fill the GPU.  Every batch
  still sees the prompts as updated by all earlier batches, so the result is the sequential one; the
  loss `step()` leaves behind is that of the previous batch (`flush()` drains the last one).

Used by bench.py and usable from rpo_b200.trainer.  Host code is plumbing only.
"""
import os

import torch

from . import _lib


class StepRunner:
    def __init__(self, model, batch, lr=0.01, momentum=0.9, weight_decay=5e-4, use_graph=True, process_group=None,
                 world_size=1, image_dtype=torch.float32, pipeline=False, context_sms=None):
        self.model = model
        self.B = int(batch)
        self.device = model.w_mm.device
        self.world = int(world_size)
        self.pg = process_group
        self.momentum, self.wd = float(momentum), float(weight_decay)
        self.pipeline = bool(pipeline)
        if self.pipeline:
            model.pipeline_images(2)
        self.eng = model.engine(self.B)
        if self.pipeline:
            # SMs the context pass may occupy (the rest stay free for the prompt-row chain); 0 = all of them
            if context_sms is None:
                context_sms = int(os.environ.get("RPO_CTX_SMS", "0"))
            self.eng.set_context_sms(context_sms)
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
        self.staged = self.sharded or self.pipeline
        self.segments = None   # staged: per slot, {stage name: callable or CUDAGraph} or one whole-chain graph
        self.side = None
        if self.staged:
            # RPO_PIPE_PRIO=1: the prompt-row chain (many short dependent kernels) gets a higher stream priority than
            # the context pass (few long kernels), so that its CTAs are placed first whenever an SM frees up.
            # Measured slower (profiles/r01_pipeline_sweep.txt): off by default
            self.prio = self.pipeline and os.environ.get("RPO_PIPE_PRIO", "0") == "1"
            hp = -1 if self.prio else 0
            self.side = torch.cuda.Stream(self.device, priority=hp)
            self.chain_stream = torch.cuda.Stream(self.device, priority=hp) if self.prio else None
        if self.pipeline:
            self.ctx_stream = torch.cuda.Stream(self.device)
            self._img = [torch.zeros_like(self.image) for _ in range(2)]
            self._lab = [torch.zeros_like(self.label) for _ in range(2)]
            self._ctx_ready = [torch.cuda.Event() for _ in range(2)]
            self._ctx_graph = [None, None]
            self.cur = None  # slot whose context rows are ready and whose prompt chain is still to run

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

    # -- staged step: native stages, text stages (and their collectives) on the side stream --------
    def _image_code(self):
        return _lib.RPO_U8 if self.image.dtype == torch.uint8 else _lib.dtype_code(self.image.dtype)

    def _stage_fns(self, slot):
        eng, pl, idt = self.eng, self.model.prompt_learner, self._image_code()
        label = self._lab[slot] if self.pipeline else self.label

        def logits():
            eng.logits_forward(label, None)
            eng.logits_backward()

        if self.pipeline:
            def image_fwd():
                eng.image_prompts(pl.img_prompt.data, slot)
        else:
            def image_fwd():
                eng.image_forward(self.image, idt, pl.img_prompt.data)
        return {"text_fwd": lambda: eng.text_forward(pl.text_prompt.data), "image_fwd": image_fwd, "logits": logits,
                "text_bwd": eng.text_backward, "image_bwd": eng.image_backward}

    @staticmethod
    def _run(seg):
        seg.replay() if isinstance(seg, torch.cuda.CUDAGraph) else seg()

    def _chain(self, segs):
        """prompt-dependent part of a step on the current stream (+ side stream), up to the flat gradient"""
        ex, main, side = self.eng.exchange, torch.cuda.current_stream(self.device), self.side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._run(segs["text_fwd"])
            if ex is not None:
                ex.gather_text_features()
        self._run(segs["image_fwd"])
        main.wait_stream(side)
        self._run(segs["logits"])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if ex is not None:
                ex.scatter_text_grads()
            self._run(segs["text_bwd"])
        self._run(segs["image_bwd"])
        main.wait_stream(side)

    def _collectives_in_chain(self):
        return self.sharded and self.world > 1

    def _run_chain(self, slot):
        """chain of `slot` (+ all-reduce + SGD) on the current stream; graphs where captured"""
        seg = self.segments[slot]
        if isinstance(seg, torch.cuda.CUDAGraph):  # whole chain in one graph (world 1: with the SGD update)
            seg.replay()
            if self.world > 1:
                self._update()
        else:
            self._chain(seg)
            self._update()

    def _context(self, slot):
        """context rows of the batch in self._img[slot] on the context stream; returns after enqueuing"""
        main = torch.cuda.current_stream(self.device)
        self.ctx_stream.wait_stream(main)  # the slot's last reader (chain two steps ago) and the input copy are on `main`
        with torch.cuda.stream(self.ctx_stream):
            g = self._ctx_graph[slot]
            if g is not None:
                g.replay()
            else:
                self.eng.image_context(self._img[slot], self._image_code(), slot)
            self._ctx_ready[slot].record(self.ctx_stream)

    def _staged_step(self):
        if not self.pipeline:
            self._run_chain(0)
            return
        if self.cur is None:  # nothing in flight: only start the context rows of this batch
            self._img[0].copy_(self.image, non_blocking=True)
            self._lab[0].copy_(self.label, non_blocking=True)
            self._context(0)
            self.cur = 0
            return
        cur, nxt = self.cur, 1 - self.cur
        self._img[nxt].copy_(self.image, non_blocking=True)
        self._lab[nxt].copy_(self.label, non_blocking=True)
        self._context(nxt)            # batch n+1: context rows, on the context stream
        self._finish(cur)             # batch n: everything that depends on the prompts
        self.cur = nxt

    def _finish(self, slot):
        main = torch.cuda.current_stream(self.device)
        cs = self.chain_stream
        if cs is None:
            main.wait_event(self._ctx_ready[slot])
            self._run_chain(slot)
            return
        cs.wait_stream(main)
        cs.wait_event(self._ctx_ready[slot])
        with torch.cuda.stream(cs):
            self._run_chain(slot)
        main.wait_stream(cs)

    def flush(self):
        """pipelined: runs the chain of the batch still in flight (its loss is then in `self.loss`)"""
        if self.pipeline and self.cur is not None:
            self.model.invalidate_text_features()
            self._finish(self.cur)
            self.cur = None

    def prepare(self, warmup=3):
        """Warm-up (sets kernel attributes, loads modules) and CUDA-graph capture of the step."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for i in range(max(1, warmup)):
                    if self.staged:
                        slots = (0, 1) if self.pipeline else (0,)
                        for slot in slots:
                            if self.pipeline:
                                self._img[slot].copy_(self.image)
                                self._lab[slot].copy_(self.label)
                                self.eng.image_context(self._img[slot], self._image_code(), slot)
                            self._chain(self._stage_fns(slot))
                            self._update()
                    else:
                        self._enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.launches_per_step = self.eng.launch_count() + 2  # + two SGD kernels
            if self.staged:
                self._capture_staged()
            elif self.use_graph:
                g = torch.cuda.CUDAGraph()
                # RPO_MAIN_PRIO=<n>: capture on a stream of that priority (negative = higher than the text tower's
                # side stream, whose kernels then only fill the gaps the vision tower leaves).  Measured on B200: any
                # priority difference inside the graph costs 15 % (3.27 -> 3.77 ms, profiles/r01_stream_priority.txt)
                mp = int(os.environ.get("RPO_MAIN_PRIO", "0"))
                cap = torch.cuda.Stream(self.device, priority=mp) if mp else None
                with torch.cuda.graph(g, stream=cap):
                    if self.world > 1:
                        self._fwd_bwd()
                    else:
                        self._enqueue()
                self.graph = g
            if self.pipeline:  # prime: context rows of the batch now in self.image
                self.cur = None
                self._staged_step()
                torch.cuda.synchronize()
        return self

    def _capture_staged(self):
        slots = (0, 1) if self.pipeline else (0,)
        self.segments = {}
        cap_stream = self.chain_stream  # kernel nodes inherit the priority of the stream they are captured on
        for slot in slots:
            fns = self._stage_fns(slot)
            if not self.use_graph:
                self.segments[slot] = fns
            elif self._collectives_in_chain():
                segs = {}
                for name, fn in fns.items():  # NCCL calls stay outside the graphs
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap_stream):
                        fn()
                    segs[name] = g
                self.segments[slot] = segs
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream):
                    self._chain(fns)
                    if self.world == 1:
                        self._update()
                self.segments[slot] = g
            if self.pipeline and self.use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.ctx_stream):
                    self.eng.image_context(self._img[slot], self._image_code(), slot)
                self._ctx_graph[slot] = g
        torch.cuda.synchronize()

    def step(self):
        """Enqueues one step on the current stream.  Inputs are whatever self.image / self.label hold.
        Pipelined: starts the context rows of this batch and finishes the previous batch (see the module text)."""
        self.model.invalidate_text_features()  # the fused SGD kernel rewrites the prompts in place
        if self.staged:
            self._staged_step()
        elif self.graph is not None:
            self.graph.replay()
            if self.world > 1:
                self._update()
        else:
            self._enqueue()

    @property
    def loss(self):
        return self.eng.loss
