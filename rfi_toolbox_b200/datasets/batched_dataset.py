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
