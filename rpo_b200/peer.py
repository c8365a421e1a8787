"""Peer-memory exchanges between the data-parallel ranks of one node (plumbing for csrc/peer.cu).

The step's three exchanges -- all-reduce of the flat prompt gradient (fused with the SGD update), all-gather of
the class-sharded text features, reduce-scatter of their gradient -- are librpo_b200's own
kernels reading and writing the peers' buffers over NVLink.  This module only allocates the buffers every rank
maps from every peer (torch.distributed._symmetric_memory: CUDA VMM allocations exchanged between the processes
of the group) and hands the mapped pointers to the C ABI; no arithmetic lives here.

`PeerExchange.create` returns None when the mapping is not available (no P2P between the devices, symmetric
memory refused, more than 8 ranks): every rank takes that decision together (one all-reduce of a flag), and the
caller (runner.StepRunner) then uses NCCL through torch.distributed.
"""
import ctypes as C

import torch

from . import _lib


class PeerExchange:
    @classmethod
    def create(cls, engine, world, group=None, required=False):
        import torch.distributed as dist
        err = None
        obj = None
        try:
            if world > _lib.PEER_MAX_WORLD:
                raise _lib.RpoError(f"peer exchange supports up to {_lib.PEER_MAX_WORLD} ranks")
            obj = cls(engine, world, group)
        except Exception as e:  # noqa: BLE001 - any failure means "not available here"
            err = e
        # all ranks together: peer memory only if every rank has it
        ok = torch.tensor([0 if obj is None else 1], dtype=torch.int32, device=engine.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            return obj
        if required:
            raise _lib.RpoError(f"peer-memory exchange is not available: {err}")
        return None

    def __init__(self, engine, world, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.eng, self.lib, self.world, self.group = engine, engine.lib, int(world), group
        self.device = engine.device
        self.rank = dist.get_rank(group)
        pg = group if group is not None else dist.group.WORLD
        # one symmetric allocation: [signal words | flat gradient | text features | their gradient], 256-B aligned parts
        ex = engine.exchange
        dtype = engine.dtype
        esz = torch.empty((), dtype=dtype).element_size()
        sig_bytes = int(self.lib.rpo_peer_signal_bytes())
        grad_bytes = engine.grad_flat.numel() * 4
        feat_bytes = ex.text_feat.numel() * esz if ex is not None else 0

        def up(n):
            return (n + 255) // 256 * 256

        offs = [0, up(sig_bytes), up(sig_bytes) + up(grad_bytes), up(sig_bytes) + up(grad_bytes) + up(feat_bytes)]
        total = offs[3] + up(feat_bytes)
        with torch.cuda.device(self.device):
            self.block = symm.empty(total, dtype=torch.uint8, device=self.device)
            self.block.zero_()
            torch.cuda.synchronize()
            try:  # older torch builds want the group announced first; newer ones deprecate the call
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    symm.enable_symm_mem_for_group(pg.group_name)
            except Exception:  # noqa: BLE001
                pass
            hdl = symm.rendezvous(self.block, pg)
        self._hdl = hdl
        bases = [int(p) for p in hdl.buffer_ptrs]
        if len(bases) != self.world or bases[self.rank] != self.block.data_ptr():
            raise _lib.RpoError("symmetric memory returned an unexpected mapping")
        self.epoch = torch.zeros(int(self.lib.rpo_peer_epoch_bytes()) // 4, dtype=torch.int32, device=self.device)
        comm = _lib.RpoPeerComm()
        for r in range(self.world):
            comm.signals[r] = bases[r] + offs[0]
        comm.epoch = self.epoch.data_ptr()
        comm.rank, comm.world = self.rank, self.world
        self.comm = comm
        vp = C.c_void_p * self.world
        self._grad_ptrs = vp(*[b + offs[1] for b in bases])
        self._feat_ptrs = vp(*[b + offs[2] for b in bases])
        self._dfeat_ptrs = vp(*[b + offs[3] for b in bases])
        # the engine's flat gradient and exchange buffers now live in the symmetric block
        n = engine.grad_flat.numel()
        engine.grad_flat = self.block[offs[1]:offs[1] + grad_bytes].view(torch.float32)
        assert engine.grad_flat.numel() == n
        if ex is not None:
            shape = ex.text_feat.shape
            ex.text_feat = self.block[offs[2]:offs[2] + feat_bytes].view(dtype).view(shape)
            ex.d_text_feat = self.block[offs[3]:offs[3] + feat_bytes].view(dtype).view(shape)
            _lib.check(self.lib.rpo_bind_text_exchange(engine.handle, _lib.ptr(ex.text_feat), _lib.ptr(ex.d_text_feat)))
            self.esz = esz
            self.row_elems = ex.K * ex.E
        dist.barrier(group=group)  # every rank's signal words are zeroed and mapped before the first kernel

    def allreduce_sgd(self, text_prompt, img_prompt, mom_buf, lr, momentum, wd, grad_scale, first):
        eng = self.eng
        _lib.check(self.lib.rpo_peer_allreduce_sgd(
            C.byref(self.comm), self._grad_ptrs, text_prompt.data_ptr(), img_prompt.data_ptr(),
            _lib.dtype_code(eng.dtype), eng.n_text, eng.grad_flat.numel(), mom_buf.data_ptr(), lr.data_ptr(),
            float(momentum), float(wd), float(grad_scale), first.data_ptr(), _lib.stream_ptr(self.device)))

    def gather_text_features(self):
        sh = self.eng.exchange.shard
        row = self.row_elems * self.esz
        _lib.check(self.lib.rpo_peer_all_gather(C.byref(self.comm), self._feat_ptrs, sh.first * row, sh.local * row,
                                                sh.per * row, _lib.stream_ptr(self.device)))

    def scatter_text_grads(self):
        sh = self.eng.exchange.shard
        _lib.check(self.lib.rpo_peer_reduce_scatter(
            C.byref(self.comm), self._dfeat_ptrs, _lib.dtype_code(self.eng.dtype), sh.first * self.row_elems,
            sh.local * self.row_elems, sh.per * self.row_elems, _lib.stream_ptr(self.device)))
