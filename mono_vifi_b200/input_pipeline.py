"""GPU input pipeline (csrc/input.cu, C ABI mvf_input_pipeline): ToTensor + horizontal flip + ColorJitter of a training item's
frames on the device, from the uint8 frames the loader resized (datasets/mono_dataset.py:102-184, 206-238).  The host side draws
the per-item parameters exactly as the reference does (`random.random() > 0.5` for flip and augmentation,
`transforms.ColorJitter.get_params` ranges and random order) and ships 8-bit frames: 1.1 MB per sample at 192x640 instead of the
8.8 MB of six fp32 tensors."""
import numpy as np
import torch

from . import _lib

FRAME_IDS = (-1, 0, 1)
BRIGHTNESS, CONTRAST, SATURATION, HUE = (0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)     # mono_dataset.py:75-78


def draw_params(B, rng, train=True):
    """per item: (prm_f [B,4] fp32, prm_i [B,6] int32) as host tensors; rng: numpy RandomState / Generator"""
    prm_f = np.ones((B, 4), dtype=np.float32)
    prm_i = np.zeros((B, 6), dtype=np.int32)
    for b in range(B):
        prm_i[b, :4] = rng.permutation(4)                                  # ColorJitter.get_params: torch.randperm(4)
        prm_f[b] = [rng.uniform(*BRIGHTNESS), rng.uniform(*CONTRAST), rng.uniform(*SATURATION), rng.uniform(*HUE)]
        prm_i[b, 4] = int(train and rng.random_sample() > 0.5)            # do_color_aug (mono_dataset.py:221)
        prm_i[b, 5] = int(train and rng.random_sample() > 0.5)            # do_flip      (mono_dataset.py:222)
    return torch.from_numpy(prm_f), torch.from_numpy(prm_i)


class InputPipeline:
    """frames_u8 [B, F, H, W, 3] (device) -> {("color", f, 0), ("color_aug", f, 0)} fp32 [B,3,H,W]; `out` lets the results land
    directly in existing buffers (the static inputs of a captured training step)."""

    def __init__(self, B, H, W, device, frame_ids=FRAME_IDS):
        self.B, self.F, self.H, self.W, self.device, self.frame_ids = B, len(frame_ids), H, W, device, tuple(frame_ids)
        self.ws = torch.empty(_lib.lib().mvf_input_pipeline_workspace_floats(B, self.F), device=device, dtype=torch.float32)
        self._tables = {}

    def _table(self, out):
        key = tuple(out[(n, f, 0)].data_ptr() for n in ("color", "color_aug") for f in self.frame_ids)
        t = self._tables.get(key)
        if t is None:
            t = torch.tensor(list(key), dtype=torch.int64).to(self.device)
            self._tables[key] = t
        return t

    def __call__(self, frames_u8, prm_f, prm_i, out=None):
        B, F, H, W = self.B, self.F, self.H, self.W
        assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and tuple(frames_u8.shape) == (B, F, H, W, 3) and frames_u8.is_contiguous()
        if out is None:
            out = {(n, f, 0): torch.empty(B, 3, H, W, device=self.device, dtype=torch.float32) for n in ("color", "color_aug") for f in self.frame_ids}
        t = self._table(out)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().mvf_input_pipeline(frames_u8.data_ptr(), prm_f.data_ptr(), prm_i.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                                 t.data_ptr(), t.data_ptr() + 8 * F, B, F, H, W, st), "mvf_input_pipeline")
        return out


class U8HostFedRunner:
    """End-to-end driver for batches that arrive as uint8 frames in pinned HOST memory: the H2D copy of batch i+1 (frames + jitter
    parameters) runs on a copy stream while step i computes; the pipeline kernel writes colour / augmented colour straight into the
    captured step's static inputs, then the graph is replayed.  feed(batch) / run() as trainer.HostFedRunner."""

    def __init__(self, graphed_step, example_host_batch):
        self.g = graphed_step
        dev = graphed_step.step.device
        self.dev = dev
        B, F, H, W, _ = example_host_batch["frames_u8"].shape
        self.pipe = InputPipeline(B, H, W, dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [{k: torch.empty_like(v, device=dev) for k, v in example_host_batch.items()} for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.head = self.tail = 0
        for e in self.consumed:
            e.record(torch.cuda.current_stream(dev))

    def feed(self, host_batch):
        i = self.head % 2
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])
            for k, v in host_batch.items():
                self.stage[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head += 1

    def run(self):
        assert self.tail < self.head, "feed() a batch first"
        i = self.tail % 2
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[i])
        s = self.stage[i]
        self.pipe(s["frames_u8"], s["jitter_f"], s["jitter_i"], out=self.g.static_inputs)
        for k in s:
            if k not in ("frames_u8", "jitter_f", "jitter_i"):
                self.g.static_inputs[k].copy_(s[k], non_blocking=True)
        self.consumed[i].record(cur)
        self.tail += 1
        return self.g()
