"""Shim for torch_geometric.data.Data: attribute bag with in-place ``.to``. TEST ONLY.
Reference call sites: data_loader.py:13,151-155; main.py:206 (return value of .to ignored)."""
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self
