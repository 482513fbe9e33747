"""Graph plan: sorted CSRs (K1, built by the CUDA library) + the work-item tables that drive the
deterministic segmented reductions of K2/K3.

The plan is built once per graph (the reference re-derives degrees, norms and gather/scatter
indices on every step: model.py:96-101) and cached on the identity of the edge tensors.
"""
import numpy as np
import torch

from . import _lib

CHUNK0 = 32      # edge records per level-0 item (one 8-lane group walks them)
CHUNK1 = 1024    # partial rows per item on the higher levels (one 256-thread block)


def build_levels(seg_beg, seg_end, seg_row, chunk0=CHUNK0, chunk1=CHUNK1):
    """Host-side scheduling (pure integer numpy): split every segment [beg,end) into items of at most
    ``chunk`` entries.  A segment that fits one item writes its final row directly; longer segments
    write one partial row per item and become segments of the next level (over partial rows) until
    one item remains.  Returns [(items int32 [n,4] = (beg, end, out, flags), n_partial_rows), ...];
    flags = (final_row << 1) | is_final, out = final row (is_final) or partial slot."""
    levels = []
    beg = np.asarray(seg_beg, dtype=np.int64)
    end = np.asarray(seg_end, dtype=np.int64)
    row = np.asarray(seg_row, dtype=np.int64)
    chunk = chunk0
    while True:
        length = end - beg
        nch = np.maximum(1, -(-length // chunk))
        total = int(nch.sum())
        seg = np.repeat(np.arange(beg.shape[0], dtype=np.int64), nch)
        first = np.cumsum(nch) - nch
        k = np.arange(total, dtype=np.int64) - first[seg]
        ibeg = beg[seg] + k * chunk
        iend = np.minimum(ibeg + chunk, end[seg])
        final = nch[seg] == 1
        slot = np.cumsum(~final) - 1
        out = np.where(final, row[seg], slot)
        flags = (row[seg] << 1) | final.astype(np.int64)
        assert total < 2 ** 31 and (flags < 2 ** 31).all()
        items = np.stack([ibeg, iend, out, flags], 1).astype(np.int32)
        n_part = int((~final).sum())
        levels.append((items, n_part))
        if n_part == 0:
            return levels
        multi = nch > 1
        pfirst = np.cumsum(np.where(multi, nch, 0)) - np.where(multi, nch, 0)
        beg, end, row = pfirst[multi], (pfirst + nch)[multi], row[multi]
        chunk = chunk1


class ReducePlan(object):
    """Device copy of the level tables of one segmented reduction."""

    def __init__(self, levels, device):
        self.levels = [(torch.from_numpy(it).to(device), int(it.shape[0]), n_part) for it, n_part in levels]
        self.max_part = max(n_part for _, _, n_part in self.levels)

    def num_items(self):
        return [n for _, n, _ in self.levels]


class GraphPlan(object):
    """Sorted CSRs + reduction plans of one (edge_index, edge_type) pair.

    Attributes mirror include/kgc_b200.h: deg[2,N] int32, norm[2E] f32, perm_{dst,src,type}[2E] int32,
    rowptr_{dst,src}[N+1], rowptr_type[T+1], rowmid_dst[N], rec_{dst,src,type}[2E,4] int32 (bit view).
    """

    def __init__(self, edge_index, edge_type, num_nodes, num_types, n_edges_in=None, n_dst_rows=None, dst_offset=0,
                 deg=None):
        """Single GPU: edge_index [2, 2E] with the in half first (defaults).  Partitioned: only the
        edges this rank owns, ``n_edges_in`` in-half edges first, src = global ids, dst = LOCAL row ids in
        [0, n_dst_rows) (global id = dst + dst_offset) and ``deg`` = the GLOBAL per-half degrees [2, num_nodes] int32."""
        edge_index = _lib.require_cuda(edge_index, torch.int64, 'edge_index')
        edge_type = _lib.require_cuda(edge_type, torch.int64, 'edge_type')
        if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_index.size(1) != edge_type.numel():
            raise ValueError('edge_index must be [2, 2E] and edge_type [2E]')
        n2 = int(edge_type.numel())
        if n_edges_in is None:
            if n2 % 2 != 0:
                raise ValueError('the edge list must hold an in half and an out half of equal size (model.py:84-90)')
            n_edges_in = n2 // 2
        dev = edge_index.device
        N, T = int(num_nodes), int(num_types)
        Nd = N if n_dst_rows is None else int(n_dst_rows)
        self.device, self.num_nodes, self.num_types, self.num_edges2 = dev, N, T, n2
        self.num_dst_rows, self.dst_offset, self.num_edges_in = Nd, int(dst_offset), int(n_edges_in)
        i32 = dict(dtype=torch.int32, device=dev)
        if deg is None:
            self.deg = torch.empty((2, N), **i32)
        else:
            self.deg = _lib.require_cuda(deg, torch.int32, 'deg')
            if tuple(self.deg.shape) != (2, N):
                raise ValueError('deg must be [2, num_nodes]')
        self.norm = torch.empty((n2,), dtype=torch.float32, device=dev)
        self.perm_dst, self.perm_src, self.perm_type = (torch.empty((n2,), **i32) for _ in range(3))
        self.rowptr_dst, self.rowptr_src = torch.empty((Nd + 1,), **i32), torch.empty((N + 1,), **i32)
        self.rowptr_type = torch.empty((T + 1,), **i32)
        self.rowmid_dst = torch.empty((Nd,), **i32)
        self.rec_dst, self.rec_src, self.rec_type = (torch.empty((n2, 4), **i32) for _ in range(3))
        h = _lib.lib()
        ws_bytes = int(h.kgc_csr_workspace_bytes(n2, N, T))
        if ws_bytes == 0:
            raise RuntimeError('kgc_csr_workspace_bytes failed: ' + h.kgc_last_error().decode())
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
        p = _lib.ptr
        _lib.call('kgc_csr_build', p(src), p(dst), p(edge_type), n2, self.num_edges_in, N, Nd, self.dst_offset, T,
                  0 if deg is None else 1, p(self.deg), p(self.norm),
                  p(self.perm_dst), p(self.rowptr_dst), p(self.rowmid_dst), p(self.rec_dst),
                  p(self.perm_src), p(self.rowptr_src), p(self.rec_src),
                  p(self.perm_type), p(self.rowptr_type), p(self.rec_type), p(ws), ws_bytes, _lib.stream())
        del ws
        # ---- reduction plans (host integer scheduling over the row pointers)
        rp_dst = self.rowptr_dst.cpu().numpy().astype(np.int64)
        rm_dst = self.rowmid_dst.cpu().numpy().astype(np.int64)
        rp_src = self.rowptr_src.cpu().numpy().astype(np.int64)
        rp_typ = self.rowptr_type.cpu().numpy().astype(np.int64)
        rows = np.arange(N, dtype=np.int64)
        drows = np.arange(Nd, dtype=np.int64)
        # forward: row i of plane 0 (in half) = [rowptr, rowmid), row Nd+i of plane 1 (out half) = [rowmid, rowptr+1)
        self.fwd = ReducePlan(build_levels(np.concatenate([rp_dst[:-1], rm_dst]), np.concatenate([rm_dst, rp_dst[1:]]),
                                           np.concatenate([drows, drows + Nd])), dev)
        self.bwd_src = ReducePlan(build_levels(rp_src[:-1], rp_src[1:], rows), dev)
        self.bwd_rel = ReducePlan(build_levels(rp_typ[:-1], rp_typ[1:], np.arange(T, dtype=np.int64)), dev)
        self._scratch = {}

    def scratch(self, name, shape, dtype=torch.float32):
        """Reusable device workspace (caller-allocated, as the C ABI requires)."""
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def run_reduction(self, rp, level0, out_final, D, addend=None, tag=''):
        """Launch level 0 through ``level0(items, n_items, out_final, out_part)`` and the higher levels
        through kgc_rows_reduce, chaining the partial-row buffers."""
        prev = None
        for li, (items, n_items, n_part) in enumerate(rp.levels):
            part = self.scratch('{}part{}'.format(tag, li), (n_part, D)) if n_part else None
            if li == 0:
                level0(items, n_items, out_final, part)
            else:
                _lib.call('kgc_rows_reduce', _lib.ptr(prev), _lib.ptr(items), n_items, _lib.ptr(out_final),
                          _lib.ptr(part), _lib.ptr(addend), D, _lib.stream())
            prev = part


_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 8


def get_plan(edge_index, edge_type, num_nodes, num_types):
    """Plans are cached on the identity + version of the edge tensors (the graph is static across steps,
    main.py:61,121)."""
    key = (edge_index.data_ptr(), edge_type.data_ptr(), edge_index._version, edge_type._version,
           tuple(edge_index.shape), tuple(edge_index.stride()), int(num_nodes), int(num_types), str(edge_index.device))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = GraphPlan(edge_index, edge_type, num_nodes, num_types)
        plan._keepalive = (edge_index, edge_type)     # pins the addresses the key refers to
        _PLAN_CACHE[key] = plan
    return plan
