"""TorchDataset -- the container `create_dataset` returns.

Mirror of rfi_toolbox/datasets/batched_dataset.py:10-76: `.images (N, H, W, 3) float32`,
`.labels (N, H, W) uint8`, `.metadata`, `len()`, `ds[i] -> {"image", "label"}`, the dtype and
length asserts, `save_to_disk` / `load_from_disk` (.pt with keys images / labels / metadata).

B200 differences: the tensors normally live in HBM (`share_memory_()` is a no-op for CUDA
tensors, exactly as in torch); `ds["data"]` / `ds["labels"]` / `ds["metadata"]` answer the
dict-style access the reference README documents (README.md:189-192), with `"data"` the
(N, 3, H, W) channel-first VIEW of the same storage; `.cpu()` returns a host copy that is
interchangeable with the reference's object.
"""
from __future__ import annotations

from pathlib import Path

import torch


class TorchDataset:
    def __init__(self, images, labels, metadata=None):
        assert len(images) == len(labels), "Images and labels must have same length"
        assert images.dtype == torch.float32, f"Images must be float32, got {images.dtype}"
        assert labels.dtype == torch.uint8, f"Labels must be uint8, got {labels.dtype}"
        self.images = images.share_memory_()
        self.labels = labels.share_memory_()
        self.metadata = metadata or {}

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            if idx == "data":
                return self.images.permute(0, 3, 1, 2)
            if idx == "images":
                return self.images
            if idx == "labels":
                return self.labels
            if idx == "metadata":
                return self.metadata
            raise KeyError(idx)
        return {"image": self.images[idx].contiguous(), "label": self.labels[idx].contiguous()}

    def cpu(self):
        return TorchDataset(self.images.cpu(), self.labels.cpu(), self.metadata)

    def to(self, device):
        return TorchDataset(self.images.to(device), self.labels.to(device), self.metadata)

    def to_host_async(self, images_out=None, labels_out=None, stream=None):
        """Device-resident dataset -> pinned host memory, asynchronously (the hand-over the
        reference's callers get for free: its tensors are host tensors, batched_dataset.py:25-33).
        `images_out` / `labels_out`: pinned buffers of at least this dataset's size to reuse
        (pinning costs more than the copy); `stream`: side stream to copy on (default: the current
        one).  Returns `(host_dataset, event)`; the host tensors are valid once `event` is done."""
        if not self.images.is_cuda:
            return self, None
        n = len(self)
        dev = self.images.device

        def dst(out, src):
            if out is None:
                return torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            assert out.is_pinned() and out.dtype == src.dtype and out.numel() >= src.numel()
            return out.view(-1)[:src.numel()].view(src.shape)

        hi, hl = dst(images_out, self.images), dst(labels_out, self.labels)
        cur = torch.cuda.current_stream(dev)
        st = stream or cur
        if st is not cur:
            st.wait_stream(cur)
        with torch.cuda.stream(st):
            hi.copy_(self.images, non_blocking=True)
            hl.copy_(self.labels, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        if st is not cur:
            self.images.record_stream(st)
            self.labels.record_stream(st)
        host = TorchDataset.__new__(TorchDataset)
        host.images, host.labels, host.metadata = hi, hl, self.metadata
        assert len(hi) == n
        return host, ev

    def save_to_disk(self, path):
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        torch.save({"images": self.images.cpu(), "labels": self.labels.cpu(), "metadata": self.metadata}, path)
        print(f"Saved TorchDataset to {path}")
        print(f"  {len(self)} samples, {self._size_gb():.2f} GB")

    @classmethod
    def load_from_disk(cls, path):
        blob = torch.load(path)
        return cls(blob["images"], blob["labels"], blob.get("metadata"))

    def _size_gb(self):
        return (self.images.element_size() * self.images.numel()
                + self.labels.element_size() * self.labels.numel()) / 1e9

    def __repr__(self):
        return (f"TorchDataset(samples={len(self)}, image_shape={tuple(self.images.shape[1:])}, "
                f"size={self._size_gb():.2f}GB)")


class BatchWriter:
    """Mirror of rfi_toolbox/datasets/batched_dataset.py:79-184: accumulates `TorchDataset`s and
    writes `batch_###.pt` files ({"images", "labels"} CPU tensors, `samples_per_batch` samples
    each) plus `metadata.json` -- the sink after the hot path.

    B200 differences: datasets normally arrive resident in HBM.  `add_batch` starts an
    asynchronous device->pinned-host copy on a side stream and returns at once, so the next
    `create_dataset` overlaps the download; the copy is awaited only when a file is written.
    File contents and names are the reference's.  `metadata.json` records the actual patch
    shape (the reference hard-codes 1024 x 1024, :172-173)."""

    def __init__(self, output_dir, samples_per_batch=100):
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.samples_per_batch = samples_per_batch
        self.accumulated_images = []
        self.accumulated_labels = []
        self._pending = []  # events of in-flight downloads
        self.batch_file_idx = 0
        self.total_samples = 0
        self._shape = None
        self._stream = None

    def _to_host(self, t):
        if not t.is_cuda:
            return t
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=t.device)
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        self._stream.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self._stream):
            host.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        t.record_stream(self._stream)
        self._pending.append(ev)
        return host

    def add_batch(self, dataset):
        self._shape = tuple(dataset.images.shape[1:])
        self.accumulated_images.append(self._to_host(dataset.images))
        self.accumulated_labels.append(self._to_host(dataset.labels))
        if sum(len(img) for img in self.accumulated_images) >= self.samples_per_batch:
            self._flush()

    def _flush(self):
        if not self.accumulated_images:
            return
        for ev in self._pending:
            ev.synchronize()
        self._pending = []
        images = torch.cat(self.accumulated_images)
        labels = torch.cat(self.accumulated_labels)
        self.accumulated_images, self.accumulated_labels = [], []
        for start in range(0, len(images), self.samples_per_batch):
            end = min(start + self.samples_per_batch, len(images))
            img, lab = images[start:end], labels[start:end]
            batch_file = self.output_dir / f"batch_{self.batch_file_idx:03d}.pt"
            torch.save({"images": img, "labels": lab}, batch_file)
            size_gb = (img.element_size() * img.numel() + lab.element_size() * lab.numel()) / 1e9
            print(f"    Wrote {batch_file.name}: {len(img)} patches ({size_gb:.2f} GB)")
            self.total_samples += len(img)
            self.batch_file_idx += 1

    def finalize(self):
        import json

        if self.accumulated_images:
            self._flush()
        shape = list(self._shape) if self._shape else [1024, 1024, 3]
        metadata = {
            "num_samples": self.total_samples,
            "samples_per_batch": self.samples_per_batch,
            "num_batches": self.batch_file_idx,
            "image_shape": shape,
            "mask_shape": shape[:2],
            "dtype": "float32",
        }
        metadata_path = self.output_dir / "metadata.json"
        with open(metadata_path, "w") as f:
            json.dump(metadata, f, indent=2)
        print("\nBatch writing complete:")
        print(f"  Total samples: {self.total_samples}")
        print(f"  Batch files: {self.batch_file_idx}")
        print(f"  Metadata: {metadata_path}")
