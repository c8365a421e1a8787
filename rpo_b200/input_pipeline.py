"""Input boundary of the hot path: what replaces the reference's synchronous
`batch["img"].to(self.device)` (trainers/rpo.py:318-323).

`BatchUploader` moves the batches of a training loop to the device without stalling it: a copy stream of its own,
`depth` device-side slots, pinned host staging for sources the DataLoader did not pin, and events in both
directions (slot ready for the step / slot consumed by the step), so the upload of batch n+1 overlaps the step of
batch n.  Images may be float32 (what the reference's transform pipeline hands over: ToTensor + Normalize done on the
host, 602 KB per image) or uint8 (raw pixels after resize / crop / flip, 151 KB per image; ToTensor + Normalize,
clip/clip.py:75-78, then happen in f32 inside the patch extraction kernel -- bit-identical results, a quarter of the
PCIe bytes).  Plumbing only: no arithmetic of the path lives here.
"""
import torch


class BatchUploader:
    def __init__(self, device, batch, resolution, image_dtype=torch.float32, depth=2):
        self.device = torch.device(device)
        self.B, self.depth = int(batch), int(depth)
        shape = (self.B, 3, int(resolution), int(resolution))
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(self.device)
            self.image = [torch.empty(shape, dtype=image_dtype, device=self.device) for _ in range(self.depth)]
            self.label = [torch.empty(self.B, dtype=torch.int64, device=self.device) for _ in range(self.depth)]
            self._ready = [torch.cuda.Event() for _ in range(self.depth)]
            self._consumed = [torch.cuda.Event() for _ in range(self.depth)]
            self._staged = [torch.cuda.Event() for _ in range(self.depth)]
            cur = torch.cuda.current_stream(self.device)
            for e in self._consumed + self._staged:
                e.record(cur)
        self._pin_img = [None] * self.depth   # allocated on first use (sources that are already pinned never need it)
        self._pin_lab = [None] * self.depth
        self._n = 0
        self.image_dtype = image_dtype
        self.bytes_per_batch = self.image[0].numel() * self.image[0].element_size() + self.B * 8

    def _pinned(self, slot, image, label):
        if image.is_pinned() and label.is_pinned():
            return image, label
        if self._pin_img[slot] is None:
            self._pin_img[slot] = torch.empty(self.image[slot].shape, dtype=self.image_dtype).pin_memory()
            self._pin_lab[slot] = torch.empty(self.B, dtype=torch.int64).pin_memory()
        self._staged[slot].synchronize()  # the previous upload out of this staging buffer has left the host
        n = image.shape[0]
        self._pin_img[slot][:n].copy_(image)
        self._pin_lab[slot][:n].copy_(label)
        return self._pin_img[slot][:n], self._pin_lab[slot][:n]

    def submit(self, image, label):
        """Starts the upload of one host batch ([n <= B, 3, res, res] in the uploader's image dtype, int64 labels);
        returns the slot.  Does not block unless all `depth` slots are still in flight."""
        if image.dtype != self.image_dtype or tuple(image.shape[1:]) != tuple(self.image[0].shape[1:]) or \
                image.shape[0] > self.B:
            raise ValueError(f"batch of {tuple(image.shape)} {image.dtype} does not fit the uploader "
                             f"({tuple(self.image[0].shape)} {self.image_dtype})")
        slot = self._n % self.depth
        self._n += 1
        n = image.shape[0]
        src_i, src_l = self._pinned(slot, image.contiguous(), label.to(torch.int64).contiguous())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._consumed[slot])
            self.image[slot][:n].copy_(src_i, non_blocking=True)
            self.label[slot][:n].copy_(src_l, non_blocking=True)
            self._ready[slot].record(self.stream)
            self._staged[slot].record(self.stream)
        return slot, n

    def acquire(self, ticket):
        """Makes the current stream wait for the upload; returns device views (valid until `release`)."""
        slot, n = ticket
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        return self.image[slot][:n], self.label[slot][:n]

    def release(self, ticket):
        """The current stream is done reading the slot (call after enqueuing the consumer)."""
        self._consumed[ticket[0]].record(torch.cuda.current_stream(self.device))


class LossReader:
    """Asynchronous read-back of the step's scalar loss: a pinned host word per slot and an event, so that the host can
    report the loss of step n-1 while step n runs (the reference blocks on `loss.item()` every step,
    trainers/rpo.py:311)."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = int(depth)
        self._host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        self._evt = [torch.cuda.Event() for _ in range(self.depth)]
        self._n = 0

    def push(self, loss_dev):
        s = self._n % self.depth
        self._n += 1
        self._host[s].copy_(loss_dev.detach().reshape(1), non_blocking=True)
        self._evt[s].record(torch.cuda.current_stream(self.device))

    def latest(self, lag=0):
        """Loss of the `lag`-th most recent push (0 = the one just pushed: waits for its step to finish)."""
        if self._n - 1 - lag < 0:
            lag = self._n - 1
        s = (self._n - 1 - lag) % self.depth
        self._evt[s].synchronize()
        return float(self._host[s][0])
