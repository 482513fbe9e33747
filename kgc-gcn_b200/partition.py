"""Destination-range partition of the graph for one-process-per-GPU runs.

Nodes are split into ``world`` equal contiguous ranges; every directed edge (and its edge_embeddings row
and optimiser state) belongs to the rank that owns its DESTINATION node; entity_embedding rows belong to
the owner of the node.  The forward needs the source rows of the owned edges (all-gather of x), the
backward returns source-row gradients to their owners (reduce-scatter of d_x).
"""
import numpy as np
import torch

from .plan import GraphPlan


def partition_edges(edge_index, edge_type, num_nodes, world, rank):
    """Pure host integer logic (numpy): which edges rank ``rank`` owns and their local numbering.

    Returns dict(lo, hi, owned_eids [n_local] int64 (in-half edges first, ascending), n_edges_in,
    src (global ids), dst (local row ids), type, deg [2, N] int32 = GLOBAL per-half out-degree by src
    (model.py:73-75)).  ``num_nodes`` must be divisible by ``world`` (equal all-gather blocks)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    et = np.asarray(edge_type, dtype=np.int64)
    n2 = ei.shape[1]
    if n2 % 2 != 0:
        raise ValueError('edge list must hold an in half and an out half of equal size')
    if num_nodes % world != 0:
        raise ValueError('num_nodes ({}) must be divisible by the number of ranks ({})'.format(num_nodes, world))
    E = n2 // 2
    per = num_nodes // world
    lo, hi = rank * per, (rank + 1) * per
    dst = ei[1]
    owned = np.nonzero((dst >= lo) & (dst < hi))[0]          # ascending: in-half edges (< E) come first
    n_in = int(np.searchsorted(owned, E))
    deg = np.stack([np.bincount(ei[0, :E], minlength=num_nodes), np.bincount(ei[0, E:], minlength=num_nodes)]).astype(np.int32)
    return {'lo': lo, 'hi': hi, 'owned_eids': owned, 'n_edges_in': n_in, 'src': ei[0, owned], 'dst': dst[owned] - lo,
            'type': et[owned], 'deg': deg}


class GraphPartition(object):
    """This rank's share of the graph + its GraphPlan (built by K1 with the global degrees)."""

    def __init__(self, edge_index, edge_type, num_nodes, num_types, world, rank, device, group=None):
        info = partition_edges(edge_index.cpu().numpy() if torch.is_tensor(edge_index) else edge_index,
                               edge_type.cpu().numpy() if torch.is_tensor(edge_type) else edge_type, num_nodes, world, rank)
        self.world, self.rank, self.group, self.num_nodes = int(world), int(rank), group, int(num_nodes)
        self.lo, self.hi = info['lo'], info['hi']
        self.owned_eids = torch.from_numpy(info['owned_eids']).to(device)
        self.n_edges_in = info['n_edges_in']
        self.edge_index = torch.from_numpy(np.stack([info['src'], info['dst']])).to(device)
        self.edge_type = torch.from_numpy(info['type']).to(device)
        self.plan = GraphPlan(self.edge_index, self.edge_type, num_nodes, num_types, n_edges_in=self.n_edges_in,
                              n_dst_rows=self.hi - self.lo, dst_offset=self.lo, deg=torch.from_numpy(info['deg']).to(device))
