"""Class-sharded text tower: host side of the exchange.

The reference runs the text tower over ALL class prompts on every GPU (trainers/rpo.py:173-192); with
G data-parallel ranks that is G times the same work, and at C = 1000 classes it is most of a step.
Here rank r runs the tower for classes [r*per, (r+1)*per) only (per = ceil(C / G)) and the step gets
two collectives over NVLink (NCCL through torch.distributed; gloo on CPU for the tests):

  forward   text_feat_local [per*K, E]  --all-gather-->  text_feat [G*per*K, E]   (logits need every class)
  backward  d_text_feat [G*per*K, E] (this rank's images)  --reduce-scatter(sum)-->  d_text_feat_local [per*K, E]

followed by the usual all-reduce (sum) of the flat prompt gradient, in which the text half now adds
up the ranks' disjoint class sets and the image half the ranks' images; with the loss defined as the
mean over the global batch both halves are scaled by 1/G afterwards (backward is linear, so the 1/G
of the reduce-scatter can wait until then).

Rows past n_cls*K (padding up to G*per classes) are zero and never read by the logits.
Plumbing only: no arithmetic of the hot path lives here.
"""
import torch


class ClassShard:
    """Contiguous partition of `n_cls` classes over `world` ranks, padded to equal parts."""

    def __init__(self, n_cls, rank, world):
        n_cls, rank, world = int(n_cls), int(rank), int(world)
        if not (world >= 1 and 0 <= rank < world):
            raise ValueError(f"rank {rank} / world {world}")
        self.n_cls, self.rank, self.world = n_cls, rank, world
        self.per = -(-n_cls // world)  # classes per rank, the last rank may hold fewer
        self.n_pad = self.per * world
        self.first = rank * self.per
        self.local = min(n_cls, self.first + self.per) - self.first
        # The decision must be the same on EVERY rank (a rank that stays replicated while the others issue the
        # collectives deadlocks them), so it depends on (n_cls, world) only, never on `rank`.
        if not self.feasible(n_cls, world):
            raise ValueError(f"{n_cls} classes over {world} ranks in parts of {self.per} leave a rank without a "
                             f"class: use fewer ranks for the text tower or do not shard it")

    @staticmethod
    def feasible(n_cls, world):
        """True if every one of `world` ranks gets at least one class with parts of ceil(n_cls / world)."""
        per = -(-int(n_cls) // int(world))
        return per * (int(world) - 1) < int(n_cls)

    def __repr__(self):
        return f"ClassShard(classes [{self.first}, {self.first + self.local}) of {self.n_cls}, rank {self.rank}/{self.world})"

    @property
    def slice(self):
        return slice(self.first, self.first + self.local)


class TextExchange:
    """Owns the exchange buffers of one rank ([n_pad*K, E] text features and their gradient, bound
    into the native handle with rpo_bind_text_exchange) and issues the two collectives on the
    current stream.  Works on any device torch.distributed has a backend for."""

    def __init__(self, shard, K, E, dtype, device, group=None):
        self.shard, self.K, self.E, self.group = shard, int(K), int(E), group
        rows = shard.n_pad * self.K
        self.text_feat = torch.zeros(rows, self.E, dtype=dtype, device=device)
        self.d_text_feat = torch.zeros(rows, self.E, dtype=dtype, device=device)
        # staging rows of the collectives (out of place: the padded part of the last rank stays zero)
        self._send = torch.zeros(shard.per * self.K, self.E, dtype=dtype, device=device)
        self._recv = torch.zeros(shard.per * self.K, self.E, dtype=dtype, device=device)
        self.r0 = shard.first * self.K
        self.nl = shard.local * self.K

    def gather_text_features(self):
        """text_feat[r0 : r0+nl] holds this rank's classes -> every rank holds all of text_feat."""
        if self.shard.world == 1:
            return
        import torch.distributed as dist
        self._send[:self.nl].copy_(self.text_feat[self.r0:self.r0 + self.nl])
        dist.all_gather_into_tensor(self.text_feat, self._send, group=self.group)

    def scatter_text_grads(self):
        """d_text_feat holds d loss_r / d text_feat for this rank's images (all classes) ->
        d_text_feat[r0 : r0+nl] holds the sum over ranks for this rank's classes."""
        if self.shard.world == 1:
            return
        import torch.distributed as dist
        dist.reduce_scatter_tensor(self._recv, self.d_text_feat, group=self.group)
        self.d_text_feat[self.r0:self.r0 + self.nl].copy_(self._recv[:self.nl])
