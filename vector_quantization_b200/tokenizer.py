"""The callers and data formats either side of the quantizer.

Reference: `BaseModel.quantize` / `encode_to_quant` (vq/tasks/image_tokenization/models/base.py:116-146) hand
the quantizer NCHW encoder features; `TokenizeCallback` (runners/callbacks.py:29-53) and
`tools/tokenize_llamagen.py:93-103` write the resulting token grids to disk.  The functions here take the
quantizer (not a whole tokenizer model: encoders/decoders are out of scope) and reproduce those contracts —
memo keys, shapes, dtypes, file formats — with the layout changes done by the transpose kernel of the C-ABI
and a tokenise-only path that never builds the int64 index tensor unless asked for.
"""
from __future__ import annotations

import pathlib
from typing import Any, Sequence

import torch

from . import functional as Fq
from . import ops
from .quantizers import VectorQuantizer, get_memo

__all__ = ['quantize', 'encode_to_quant', 'save_tokens', 'load_tokens', 'save_llamagen_codes', 'TokenStreamWriter']


def quantize(quantizer, x: torch.Tensor, memo: dict):
    """`BaseModel.quantize` (base.py:116-129): x [b, c, h, w] -> (z [b, c, h, w] contiguous, q_loss, memo) with
    memo['quantizer'] = the quantizer's memo (+ 'x_shape')."""
    quantizer_memo = get_memo(memo, 'quantizer')
    quantizer_memo['x_shape'] = (b, c, h, w) = tuple(x.shape)
    if isinstance(quantizer, VectorQuantizer) and quantizer.can_fuse_nchw() and x.is_cuda:
        # Layout fusion: ONE transposed copy of the latents feeds the assignment, the callbacks and the kernels' token
        # reads; z is written NCHW by the gather kernel, and the backward reads the NCHW upstream gradient and writes
        # the NCHW token gradient directly — no transpose after the quantizer, none in the backward.
        x = x.contiguous()
        rows = ops.transpose_last2(x.detach().view(b, c, h * w)).view(b * h * w, c)
        quantizer_memo['_nchw'] = x
        z, q_loss, memo['quantizer'] = quantizer(rows, quantizer_memo)
        return z, q_loss, memo
    rows = Fq.nchw_to_rows(x)
    z, q_loss, memo['quantizer'] = quantizer(rows, quantizer_memo)
    return Fq.rows_to_nchw(z, b, c, h, w), q_loss, memo


@torch.no_grad()
def encode_to_quant(quantizer, x: torch.Tensor, memo: dict, *, compact: bool = False):
    """`BaseModel.encode_to_quant` after the encoder (base.py:131-146): x [b, c, h, w] -> (quant [b, h, w], memo);
    memo['quantizer'] gets 'x_shape', 'x' (token-major, after the callbacks) and 'quant' ([b*h*w] int64).
    compact=True is the tokenise-only variant: token ids come back as uint16 (K <= 65 536) or int32 straight
    from the packed assignment keys and memo['quantizer']['quant'] holds that compact tensor."""
    quantizer_memo = get_memo(memo, 'quantizer')
    quantizer_memo['x_shape'] = (b, _, h, w) = tuple(x.shape)
    rows = Fq.nchw_to_rows(x)
    # keep the packed keys (no int64 unpack launch) unless a training callback consumes the indices in after_encode
    lazy = compact and isinstance(quantizer, VectorQuantizer) and quantizer._callbacks.packed_keys_ok()
    if lazy:
        quantizer_memo['_lazy_unpack'] = True
    rows, quant, memo['quantizer'] = quantizer.encode(rows, quantizer_memo)
    quantizer_memo = memo['quantizer']
    if lazy:
        quant = ops.compact_tokens(quant, quantizer.codebook_size)
    elif compact and quant.dtype == torch.int64:
        quant = quant.to(torch.int32)
    quantizer_memo.update(x=rows, quant=quant)
    return quant.view(b, h, w), memo


def save_tokens(path: str | pathlib.Path, id_: Sequence[Any], category: torch.Tensor, quant: torch.Tensor,
                x_shape: Sequence[int]) -> None:
    """`TokenizeCallback.after_run_iter` (runners/callbacks.py:40-53): torch.save of the `Tokens` dict
    {id_, category, tokens [b, h, w]}.  Compact ids are widened to int64 so that reference readers see the
    dtype they expect; pass an int64 tensor to skip the conversion."""
    b, _, h, w = x_shape
    tokens = quant.reshape(b, h, w)
    if tokens.dtype != torch.int64:
        tokens = tokens.to(torch.int32).to(torch.int64) if tokens.dtype == torch.uint16 else tokens.to(torch.int64)
    torch.save(dict(id_=id_, category=category, tokens=tokens.cpu()), str(path))


class TokenStreamWriter:
    """Streaming variant of `TokenizeCallback` (runners/callbacks.py:29-53): one `{iter}_{rank}.pth` `Tokens` file per
    iteration under `token_dir`, written WITHOUT stalling the GPU.  `write` enqueues an asynchronous device-to-host
    copy of the token ids into a pinned staging buffer on a side stream and returns; a background thread waits for
    the copy (CUDA event) and does the `torch.save`.  `depth` staging buffers bound the memory; `close()` (or
    leaving the `with` block) drains the queue.  The files are exactly what `save_tokens` writes."""

    def __init__(self, token_dir: str | pathlib.Path, rank: int = 0, depth: int = 4) -> None:
        import queue
        import threading
        self.token_dir = pathlib.Path(token_dir)
        self.token_dir.mkdir(parents=True, exist_ok=True)       # TokenizeCallback.bind
        self.rank = rank
        self._free: 'queue.Queue' = queue.Queue()
        for _ in range(depth):
            self._free.put(None)                                # staging slots, allocated lazily at the first size seen
        self._jobs: 'queue.Queue' = queue.Queue()
        self._stream = None
        self._error = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        while True:
            job = self._jobs.get()
            if job is None:
                return
            path, id_, category, host, event, shape = job
            try:
                if event is not None:
                    event.synchronize()
                tokens = host.reshape(shape)
                if tokens.dtype != torch.int64:
                    tokens = tokens.to(torch.int32).to(torch.int64) if tokens.dtype == torch.uint16 else tokens.to(torch.int64)
                torch.save(dict(id_=id_, category=category, tokens=tokens.clone()), str(path))
            except Exception as exc:  # noqa: BLE001 - surfaced by the next write()/close()
                self._error = exc
            finally:
                self._free.put(host)

    def write(self, iter_: int, id_: Sequence[Any], category: torch.Tensor, quant: torch.Tensor,
              x_shape: Sequence[int]) -> pathlib.Path:
        if self._error is not None:
            raise self._error
        b, _, h, w = x_shape
        host = self._free.get()                                 # blocks only when `depth` writes are in flight
        flat = quant.reshape(-1)
        if host is None or host.numel() != flat.numel() or host.dtype != flat.dtype:
            host = torch.empty(flat.numel(), dtype=flat.dtype).pin_memory() if flat.is_cuda else torch.empty_like(flat)
        event = None
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                host.copy_(flat, non_blocking=True)
                flat.record_stream(self._stream)
                event = torch.cuda.Event()
                event.record(self._stream)
        else:
            host.copy_(flat)
        path = self.token_dir / f'{iter_}_{self.rank}.pth'
        self._jobs.put((path, list(id_), category.detach().cpu(), host, event, (b, h, w)))
        return path

    def close(self) -> None:
        self._jobs.put(None)
        self._thread.join()
        if self._error is not None:
            raise self._error

    def __enter__(self) -> 'TokenStreamWriter':
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def load_tokens(path: str | pathlib.Path) -> dict:
    return torch.load(str(path), weights_only=False)


def save_llamagen_codes(code_path: str | pathlib.Path, label_path: str | pathlib.Path, quant: torch.Tensor,
                        category: torch.Tensor) -> None:
    """tools/tokenize_llamagen.py:93-103: codes as .npy of shape (1, 10, -1) (ten-crop augmentations x tokens),
    labels as a separate .npy."""
    import numpy as np
    q = quant.reshape(1, 10, -1)
    if q.dtype == torch.uint16:
        q = q.to(torch.int32)
    np.save(str(code_path), q.to(torch.int64).cpu().numpy())
    np.save(str(label_path), category.cpu().numpy())
