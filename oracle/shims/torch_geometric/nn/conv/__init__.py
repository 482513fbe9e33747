"""Shim for torch_geometric.nn.conv.MessagePassing (PyG 1.3-1.6 era API). TEST ONLY.

Published behaviour restated: ``propagate(edge_index, size=None, **kwargs)`` with the
default ``flow='source_to_target'`` builds every ``message`` argument named ``<k>_j``
as ``kwargs[k][edge_index[0]]`` and ``<k>_i`` as ``kwargs[k][edge_index[1]]``, passes
other arguments through by name, aggregates the messages onto ``edge_index[1]`` with
``aggr`` ('add' here, model.py:50) into ``x.size(0)`` rows, then calls ``update``.
Reference call sites: model.py:4,47-50,99-101,111,120.
"""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', **kwargs):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target'
        self.aggr = aggr
        self._msg_args = [p for p in inspect.signature(self.message).parameters]
        self._upd_args = [p for p in inspect.signature(self.update).parameters][1:]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        num_rows = None
        call = {}
        for name in self._msg_args:
            if name.endswith('_j'):
                base = kwargs[name[:-2]]
                num_rows = base.size(0)
                call[name] = base.index_select(0, src)
            elif name.endswith('_i'):
                base = kwargs[name[:-2]]
                num_rows = base.size(0)
                call[name] = base.index_select(0, dst)
            elif name == 'edge_index':
                call[name] = edge_index
            else:
                call[name] = kwargs.get(name)
        if size is not None and size[1] is not None:
            num_rows = size[1]
        msg = self.message(**call)
        out = msg.new_zeros((num_rows,) + tuple(msg.shape[1:]))
        out.index_add_(0, dst, msg)
        return self.update(out, **{k: kwargs[k] for k in self._upd_args if k in kwargs})

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
