"""Checkpoint I/O of the reference (utils.py:121-155) for the drop-in classes - SURVEY.md "next" row N4, second half.

Same call signatures, same files (``last.ckpt`` / ``best.ckpt`` = ``torch.save`` of the caller's state dict:
``{'state_dict': ..., 'optim_dict': ..., ...}``, main.py:153-163) - a checkpoint written by either side loads on the
other (tests/test_host_cpu.py::test_checkpoint_interchanges_with_the_reference).  What changes is how the bytes move: at
the Wikidata5M shape a checkpoint is ~24 GB of parameters (+ 2x that of Adam state) sitting in HBM.

* save: ``torch.save`` of CUDA tensors copies them to pageable host memory one by one, synchronously.  Here every CUDA
  tensor of the state goes device -> pinned staging buffer -> its host copy through two staging buffers on a copy stream,
  so the PCIe copy of chunk k + 1 runs while chunk k is moved out of the staging buffer; ``best.ckpt`` is a hard link
  to ``last.ckpt`` (``shutil.copyfile`` rewrites the whole file; a link costs nothing and stays valid when the next
  ``last.ckpt`` replaces the directory entry - it is written to a temporary name and renamed, so a crash never leaves
  a torn checkpoint).
* load: ``torch.load(..., map_location='cpu', mmap=True)`` maps the file instead of reading it into a second host copy;
  ``load_state_dict`` then copies page by page into the (device) parameters.  A missing file raises
  ``FileNotFoundError`` (the reference's ``raise "<str>"`` is itself a TypeError in Python 3, utils.py:147).
"""
import os
import shutil

import torch
import torch.nn as nn

CHUNK_BYTES = 64 << 20


def get_param(shape):
    """utils.get_param (utils.py:113-118): xavier-uniform Parameter."""
    param = nn.Parameter(torch.empty(*shape))
    nn.init.xavier_uniform_(param.data)
    return param


class _Stager(object):
    """device -> host copies through two pinned staging buffers on a side stream (overlaps PCIe with the host memcpy)."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.bufs = [torch.empty(CHUNK_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]

    def to_host(self, t):
        src = t.detach().contiguous().view(-1).view(torch.uint8)
        out = torch.empty(src.numel(), dtype=torch.uint8)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        pending = []                                              # (buffer index, offset, bytes) copies in flight
        n, off, k = src.numel(), 0, 0
        while off < n or pending:
            if off < n and len(pending) < 2:
                nb = min(CHUNK_BYTES, n - off)
                with torch.cuda.stream(self.stream):
                    self.bufs[k][:nb].copy_(src[off:off + nb], non_blocking=True)
                    self.events[k].record(self.stream)
                pending.append((k, off, nb))
                off += nb
                k ^= 1
                continue
            kb, o, nb = pending.pop(0)
            self.events[kb].synchronize()
            out[o:o + nb].copy_(self.bufs[kb][:nb])
        return out.view(t.dtype).view(t.shape)


def _to_host(obj, stagers):
    if torch.is_tensor(obj):
        if not obj.is_cuda:
            return obj
        st = stagers.get(obj.device)
        if st is None:
            st = stagers[obj.device] = _Stager(obj.device)
        return st.to_host(obj)
    if isinstance(obj, dict):
        return type(obj)((k, _to_host(v, stagers)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(v, stagers) for v in obj)
    return obj


def save_checkpoint(state, is_best, checkpoint_dir):
    """utils.save_checkpoint (utils.py:121-136): ``state`` -> ``checkpoint_dir/last.ckpt``; ``is_best`` -> also
    ``best.ckpt``.  Same file format (``torch.save``); see the module docstring for how the bytes move."""
    filepath = os.path.join(checkpoint_dir, 'last.ckpt')
    if not os.path.exists(checkpoint_dir):
        print("Checkpoint Directory does not exist! Making directory {}".format(checkpoint_dir))
        os.mkdir(checkpoint_dir)
    host_state = _to_host(state, {})
    tmp = filepath + '.tmp'
    torch.save(host_state, tmp)
    os.replace(tmp, filepath)
    if is_best:
        best = os.path.join(checkpoint_dir, 'best.ckpt')
        tmp_best = best + '.tmp'
        if os.path.lexists(tmp_best):
            os.remove(tmp_best)
        try:
            os.link(filepath, tmp_best)
        except OSError:                                           # file systems without hard links
            shutil.copyfile(filepath, tmp_best)
        os.replace(tmp_best, best)


def load_checkpoint(checkpoint, model, optimizer=None):
    """utils.load_checkpoint (utils.py:139-155): loads ``state_dict`` into ``model`` (strict) and, when ``optimizer`` is
    given, ``optim_dict`` into it; returns ``checkpoint.get('measure')``."""
    if not os.path.exists(checkpoint):
        raise FileNotFoundError("File doesn't exist {}".format(checkpoint))
    try:
        ckpt = torch.load(checkpoint, map_location='cpu', mmap=True, weights_only=False)
    except (RuntimeError, ValueError, TypeError):                 # legacy (non-zip) files cannot be mapped
        ckpt = torch.load(checkpoint, map_location='cpu', weights_only=False)
    model.load_state_dict(ckpt['state_dict'])
    if optimizer:
        optimizer.load_state_dict(ckpt['optim_dict'])
    return ckpt.get('measure', None)
