"""Weight ingest from safetensors files: the host-side step between `load_model` and `initialize_model`.

The reference reads every file whole (`std::fs::read`) and materialises a `HashMap<String, Tensor>` on the device before the
adapter sees it (providers/huggingface/huggingface.rs:81-130: single `model.safetensors`, or the files named by
`model.safetensors.index.json`'s `weight_map`).  Here the files are memory-mapped and every tensor is handed to
`fl_model_put_tensor` straight from the mapping (f32 / f16 / bf16 as stored; the library rounds to bf16, keeps this rank's
shard and uploads), so the host never holds a second copy of the checkpoint.

The safetensors container (public format): 8-byte little-endian header length N, N bytes of JSON
{name: {"dtype", "shape", "data_offsets": [begin, end]}, "__metadata__": {...}}, then the raw little-endian tensor bytes.
"""
from __future__ import annotations

import json
import mmap
import os
import struct

import numpy as np

_DTYPES = {"F32": np.float32, "F16": np.float16, "BF16": np.uint16}     # bf16 travels as raw bit patterns (FL_DTYPE_BF16)


class SafetensorsFile:
    """Memory-mapped view of one .safetensors file: `names()`, `get(name)` -> numpy array over the mapping (no copy)."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 8:
            raise ValueError(f"{path}: not a safetensors file")
        (n,) = struct.unpack("<Q", self._f.read(8))
        if n > size - 8:
            raise ValueError(f"{path}: header length {n} exceeds the file")
        self.header = json.loads(self._f.read(n).decode("utf-8"))
        self.header.pop("__metadata__", None)
        self._base = 8 + n
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)

    def names(self):
        return list(self.header)

    def get(self, name: str) -> np.ndarray:
        e = self.header[name]
        if e["dtype"] not in _DTYPES:
            raise ValueError(f"{self.path}: tensor {name} has unsupported dtype {e['dtype']}")
        b, end = e["data_offsets"]
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if (end - b) != count * dt.itemsize:
            raise ValueError(f"{self.path}: tensor {name} data_offsets do not match its shape")
        return np.frombuffer(self._mm, dtype=dt, count=count, offset=self._base + b).reshape(e["shape"])

    def close(self):
        try:
            self._mm.close()
        except BufferError:
            pass          # arrays over the mapping are still alive: it is unmapped when the last of them goes away
        finally:
            self._f.close()


def checkpoint_files(model_dir: str):
    """huggingface.rs:83-121: `model.safetensors` if present, else the distinct files of the index's weight_map."""
    single = os.path.join(model_dir, "model.safetensors")
    if os.path.exists(single):
        return [single]
    index = os.path.join(model_dir, "model.safetensors.index.json")
    if not os.path.exists(index):
        raise FileNotFoundError("Failed to find either model.safetensors or model.safetensors.index.json")
    wm = json.load(open(index)).get("weight_map")
    if not isinstance(wm, dict):
        raise ValueError("Invalid index file format: missing or invalid weight_map")
    return [os.path.join(model_dir, f) for f in sorted(set(wm.values()))]


def iter_tensors(model_dir: str):
    """Yields (name, array-over-the-mapping) for every tensor of the checkpoint; the array is valid until the next yield."""
    for path in checkpoint_files(model_dir):
        sf = SafetensorsFile(path)
        try:
            for name in sf.names():
                yield name, sf.get(name)
        finally:
            sf.close()


def write_safetensors(path: str, tensors: dict):
    """Test helper: writes f32 / f16 / uint16(bf16 bits) arrays as one safetensors file."""
    rev = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.uint16): "BF16"}
    header, off, blobs = {}, 0, []
    for name, a in tensors.items():
        a = np.ascontiguousarray(a)
        raw = a.tobytes()
        header[name] = {"dtype": rev[a.dtype], "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for raw in blobs:
            f.write(raw)
