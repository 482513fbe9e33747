"""Fused 1-N scoring + filtered ranking (K6) and the reference-facing evaluation helpers.

Replaces ConvE's scoring tail (model.py:177-179) + predict's filter / double argsort (main.py:117-133)
for evaluation.  The [B, N] score matrix is never written: a tcgen05 / TMA kernel sweeps the bf16 entity
table once per pair of 128-query tiles and counts, per query, the entities whose logit beats the target's.

Tie semantics: rank = 1 + count_gt on the LOGIT (sigmoid is monotone but saturates
in fp32, which manufactures ties the reference then breaks arbitrarily); count_eq is reported so that
callers can check reference_rank in [1 + gt, 1 + gt + eq].
"""
import torch

from . import _lib


class _Score1N(torch.autograd.Function):
    """pred[B, N] = sigmoid(x @ all_ent^T + bias) with its autograd on the tensor-core kernels (K6t): the training-time
    scoring tail of model.py:177-179."""

    @staticmethod
    def forward(ctx, x, all_ent, bias):
        from .conv import gemm_nt  # noqa: F401  (same packing routine)
        x = _lib.require_cuda(x.detach(), torch.float32, 'x')
        ent = _lib.require_cuda(all_ent.detach(), torch.float32, 'all_ent')
        bias_ = _lib.require_cuda(bias.detach(), torch.float32, 'bias')
        B, D = int(x.shape[0]), int(x.shape[1])
        N = int(ent.shape[0])
        ldp = (N + 3) // 4 * 4
        p = _lib.ptr
        nbytes = int(_lib.lib().kgc_gemm_packed_b_bytes(B, D))
        packed = torch.empty((nbytes // 4,), dtype=torch.float32, device=x.device)
        buf = torch.empty((B, ldp), dtype=torch.float32, device=x.device)
        _lib.call('kgc_gemm_pack_b', p(x), x.stride(1), x.stride(0), B, D, p(packed), _lib.stream())
        _lib.call('kgc_score_1n_fwd', p(ent), N, D, ent.stride(0), p(packed), B, p(bias_), p(buf), ldp, _lib.stream())
        pred = buf[:, :N]
        ctx.save_for_backward(x, ent, pred)
        return pred

    @staticmethod
    def backward(ctx, d_pred):
        from .conv import gemm_nt, gemm_tn
        x, ent, pred = ctx.saved_tensors
        B, D = int(x.shape[0]), int(x.shape[1])
        N = int(ent.shape[0])
        if d_pred.stride(1) != 1:
            d_pred = d_pred.contiguous()
        ldt = (B + 3) // 4 * 4
        d_logit_t = torch.empty((N, ldt), dtype=torch.float32, device=x.device)
        d_bias = torch.empty((N,), dtype=torch.float32, device=x.device)
        p = _lib.ptr
        _lib.call('kgc_score_1n_bwd_logit', p(d_pred), d_pred.stride(0), p(pred), pred.stride(0), N, B, ldt,
                  p(d_logit_t), p(d_bias), _lib.stream())
        d_x = d_ent = None
        if ctx.needs_input_grad[0]:
            d_x = gemm_tn(d_logit_t, ent, torch.empty((ldt, D), dtype=torch.float32, device=x.device))[:B]
        if ctx.needs_input_grad[1]:
            d_ent = gemm_nt(d_logit_t[:, :B], x, torch.empty((N, D), dtype=torch.float32, device=x.device))
        return d_x, d_ent, (d_bias if ctx.needs_input_grad[2] else None)


B_TILE = 256      # queries per K6t launch (the packed query operand of one launch); larger batches run tile by tile


def score_1n_supported(x, all_ent):
    """Shapes the tensor-core training scorer takes (the caller reports anything else through _lib.library_path).  Any batch
    size: batches above B_TILE queries are scored tile by tile."""
    return (x.is_cuda and x.dtype == torch.float32 and all_ent.dtype == torch.float32 and x.dim() == 2
            and x.shape[0] > 0 and x.shape[1] <= 224 and x.shape[1] % 4 == 0
            and all_ent.stride(1) == 1 and all_ent.stride(0) % 4 == 0 and all_ent.data_ptr() % 16 == 0)


def score_1n(x, all_ent, bias):
    """sigmoid(x @ all_ent^T + bias) [B, N] (model.py:177-179), differentiable."""
    if x.shape[0] <= B_TILE:
        return _Score1N.apply(x, all_ent, bias)
    return torch.cat([_Score1N.apply(x[i:i + B_TILE], all_ent, bias) for i in range(0, x.shape[0], B_TILE)], 0)


class _Score1NBCE(torch.autograd.Function):
    """loss = BCELoss(sigmoid(x @ all_ent^T + bias), label) with the label given as SPARSE positives:
    model.py:177-179 + model.py:42-44 + the label build of data_loader.py:34-43 in one differentiable call.  The forward
    runs the K6t scorer, sets one bit per positive and makes ONE pass over pred that yields the mean loss, the logit
    gradient (transposed) and the bias gradient for a unit upstream gradient; the backward scales by the upstream scalar
    and runs the two gradient GEMMs.  The dense label and the BCE forward / backward tensors are never built."""

    @staticmethod
    def forward(ctx, x, all_ent, bias, qid, ptr, idx, pos, add):
        x = _lib.require_cuda(x.detach(), torch.float32, 'x')
        ent = _lib.require_cuda(all_ent.detach(), torch.float32, 'all_ent')
        bias_ = _lib.require_cuda(bias.detach(), torch.float32, 'bias')
        qid = _lib.require_cuda(qid, torch.int64, 'qid')
        ptr = _lib.require_cuda(ptr, torch.int64, 'ptr')
        idx = _lib.require_cuda(idx, torch.int32, 'idx')
        B, D = int(x.shape[0]), int(x.shape[1])
        N = int(ent.shape[0])
        if int(qid.numel()) != B:
            raise ValueError('qid has {} entries for {} queries'.format(int(qid.numel()), B))
        ldt = (B + 3) // 4 * 4
        p, dev = _lib.ptr, x.device
        packed = torch.empty((int(_lib.lib().kgc_gemm_packed_b_bytes(B, D)) // 4,), dtype=torch.float32, device=dev)
        # entity-major throughout: the scorer stores predT [N, ldt] (its natural orientation), the positives are bits per
        # entity, and the loss pass overwrites predT with the logit gradient in place (the [N, ldt] operand of both
        # gradient GEMMs) - no [B, N]-pitched buffer, no transpose
        d_logit_t = torch.empty((N, ldt), dtype=torch.float32, device=dev)
        _lib.call('kgc_gemm_pack_b', p(x), x.stride(1), x.stride(0), B, D, p(packed), _lib.stream())
        _lib.call('kgc_score_1n_fwd_t', p(ent), N, D, ent.stride(0), p(packed), B, p(bias_), p(d_logit_t), ldt, _lib.stream())
        mask_t = torch.empty((N, (B + 31) // 32), dtype=torch.int32, device=dev)          # uint32 bits; zeroed by the call
        _lib.call('kgc_label_mask_t_build', p(qid), B, p(ptr), p(idx), N, p(mask_t), _lib.stream())
        d_bias = torch.empty((N,), dtype=torch.float32, device=dev)
        partial = torch.empty((int(_lib.lib().kgc_bce_1n_t_blocks(N)),), dtype=torch.float64, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        _lib.call('kgc_bce_1n_bwd_logit_t', p(d_logit_t), p(mask_t), N, B, ldt, float(pos), float(add), p(d_bias),
                  p(partial), p(loss), _lib.stream())
        ctx.save_for_backward(x, ent, d_logit_t, d_bias)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        from .conv import gemm_nt, gemm_tn
        x, ent, d_logit_t, d_bias = ctx.saved_tensors
        B, D = int(x.shape[0]), int(x.shape[1])
        N, ldt = int(ent.shape[0]), int(d_logit_t.shape[1])
        d_x = d_ent = None
        if ctx.needs_input_grad[0]:
            d_x = gemm_tn(d_logit_t, ent, torch.empty((ldt, D), dtype=torch.float32, device=x.device))[:B] * g
        if ctx.needs_input_grad[1]:                      # the upstream scalar rides on the small operand
            d_ent = gemm_nt(d_logit_t[:, :B], (x * g).contiguous(), torch.empty((N, D), dtype=torch.float32, device=x.device))
        return d_x, d_ent, (d_bias * g if ctx.needs_input_grad[2] else None), None, None, None, None, None


def score_1n_bce(x, all_ent, bias, qid, ptr, idx, pos=1.0, add=0.0):
    """Mean BCE of sigmoid(x @ all_ent^T + bias) against the labels of queries ``qid`` (device int64 [B]) given as the
    query -> objects CSR (``ptr`` int64, ``idx`` int32, on the device): label = ``pos`` on the positives, ``add`` elsewhere
    (``KBDataset.label_values()``).  Differentiable in x, all_ent and bias; same shapes as ``score_1n_supported``."""
    B = int(x.shape[0])
    if B <= B_TILE:
        return _Score1NBCE.apply(x, all_ent, bias, qid, ptr, idx, pos, add)
    total = None
    for i in range(0, B, B_TILE):                       # mean over B * N = tile means weighted by their share of the batch
        part = _Score1NBCE.apply(x[i:i + B_TILE], all_ent, bias, qid[i:i + B_TILE], ptr, idx, pos, add)
        part = part * (float(min(B_TILE, B - i)) / B)
        total = part if total is None else total + part
    return total


def score_kpad(d):
    kpad = int(_lib.lib().kgc_score_kpad(int(d)))
    if kpad < 0:
        raise ValueError('scoring dimension {} too large (d + 3 <= 256)'.format(d))
    return kpad


class EntityTable(object):
    """bf16 copy of all_ent [N, d] with the decoder bias folded into three extra K columns (pitch kpad)."""

    def __init__(self, all_ent, bias=None):
        all_ent = _lib.require_cuda(all_ent.detach(), torch.float32, 'all_ent')
        self.n, self.d = int(all_ent.shape[0]), int(all_ent.shape[1])
        self.kpad = score_kpad(self.d)
        if bias is not None:
            bias = _lib.require_cuda(bias.detach(), torch.float32, 'bias')
        self.data = torch.empty((self.n, self.kpad), dtype=torch.bfloat16, device=all_ent.device)
        _lib.call('kgc_score_pack_entities', _lib.ptr(all_ent), _lib.ptr(bias), self.n, self.d, _lib.ptr(self.data),
                  _lib.stream())


def pack_queries(xq):
    xq = _lib.require_cuda(xq.detach(), torch.float32, 'xq')
    b, d = int(xq.shape[0]), int(xq.shape[1])
    out = torch.empty((b, score_kpad(d)), dtype=torch.bfloat16, device=xq.device)
    _lib.call('kgc_score_pack_queries', _lib.ptr(xq), b, d, _lib.ptr(out), _lib.stream())
    return out


def pair_scores(q_bf16, table, pair_q, pair_e):
    """s[p] = <Q[pair_q[p]], E[pair_e[p]]> (+ bias) through the same tcgen05 path as the sweep."""
    n_pairs = int(pair_q.numel())
    out = torch.empty((n_pairs,), dtype=torch.float32, device=q_bf16.device)
    if n_pairs == 0:
        return out
    ws_bytes = int(_lib.lib().kgc_score_pairs_workspace_bytes(n_pairs, table.kpad))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q_bf16.device)
    p = _lib.ptr
    _lib.call('kgc_score_pairs', p(q_bf16), p(table.data), p(pair_q), p(pair_e), n_pairs, table.kpad, p(out), p(ws),
              ws_bytes, _lib.stream())
    return out


def dense_logits(xq, all_ent, bias):
    """logits[B, N] = xq @ all_ent^T + bias on the fp32-grade tensor-core kernel (3xTF32, K6t without the sigmoid)."""
    xq = _lib.require_cuda(xq.detach(), torch.float32, 'xq')
    ent = _lib.require_cuda(all_ent.detach(), torch.float32, 'all_ent')
    bias = _lib.require_cuda(bias.detach(), torch.float32, 'bias')
    B, D, N = int(xq.shape[0]), int(xq.shape[1]), int(ent.shape[0])
    if not score_1n_supported(xq, ent) or B > B_TILE:
        raise ValueError('dense_logits: at most {} queries per call, Dout <= 224, Dout % 4 == 0'.format(B_TILE))
    ldp = (N + 3) // 4 * 4
    p = _lib.ptr
    packed = torch.empty((int(_lib.lib().kgc_gemm_packed_b_bytes(B, D)) // 4,), dtype=torch.float32, device=xq.device)
    buf = torch.empty((B, ldp), dtype=torch.float32, device=xq.device)
    _lib.call('kgc_gemm_pack_b', p(xq), xq.stride(1), xq.stride(0), B, D, p(packed), _lib.stream())
    _lib.call('kgc_score_1n_logits', p(ent), N, D, ent.stride(0), p(packed), B, p(bias), p(buf), ldp, _lib.stream())
    return buf, ldp


def _filtered_rank_fp32(xq, all_ent, bias, obj, filt_ptr, filt_idx, want_eq):
    """Exact-mode core: fp32-grade dense logits tile by tile (the matrix the reference materialises, main.py:121), integer
    counts over them, target / filter logits read from the SAME matrix (bit-consistent with the counts)."""
    dev, b, n = xq.device, int(xq.shape[0]), int(all_ent.shape[0])
    gt = torch.zeros((b,), dtype=torch.int32, device=dev)
    eq = torch.zeros((b,), dtype=torch.int32, device=dev) if want_eq else None
    thr = torch.empty((b,), dtype=torch.float32, device=dev)
    s_filt = torch.empty((int(filt_idx.numel()),), dtype=torch.float32, device=dev)
    lens = filt_ptr[1:] - filt_ptr[:-1]
    rows_of = torch.repeat_interleave(torch.arange(b, device=dev), lens)
    fp = filt_ptr.tolist() if b <= 4096 else None
    p = _lib.ptr
    for i in range(0, b, B_TILE):
        j = min(i + B_TILE, b)
        z, ld = dense_logits(xq[i:j], all_ent, bias)
        thr[i:j] = z[torch.arange(j - i, device=dev), obj[i:j]]
        a, e = (fp[i], fp[j]) if fp is not None else (int(filt_ptr[i]), int(filt_ptr[j]))
        if e > a:
            s_filt[a:e] = z[rows_of[a:e] - i, filt_idx[a:e].long()]
        _lib.call('kgc_rank_count_dense', p(z), ld, n, j - i, p(thr[i:j]), p(gt[i:j]), p(eq[i:j]) if eq is not None else None,
                  _lib.stream())
    return thr, s_filt, gt, eq


def filtered_rank(xq, all_ent, bias, obj, filt_ptr, filt_idx, count_eq=False, table=None, n_offset=0, group=None,
                  precision='bf16'):
    """Filtered rank of obj[q] among all entities for every query row of ``xq`` [B, d].

    filt_ptr [B+1] int64 / filt_idx [nnz] int32: the known positives of each query (CSR, the all-split filter
    of data_loader.py:108-110; may or may not contain the target).  Returns a dict with ranks [B] int32,
    count_gt [B] int32 (already filter-corrected), count_eq (if asked), thr [B] (target logits) and
    sums [13] float64 = {count, sum rank, sum 1/rank, hits@1..10} (main.py:128-133).

    ``precision='bf16'`` (default): the fused tcgen05 sweep K6 - operands rounded to bf16 (logits within 2^-7 relative of
    fp32), the [B, N] matrix never written; the throughput path.  ``precision='fp32'``: exact mode - fp32-grade (3xTF32)
    dense logits tile by tile + integer counts over them: what the reference ranks (fp32 scores, main.py:121-126), for
    evaluations whose metrics must not depend on bf16 rounding of near-tied candidates (needs ``all_ent`` / ``bias``).

    Entity-sharded use (one process per GPU, bf16 path): pass this rank's ``table`` shard (rows n_offset .. n_offset+n)
    and the process ``group``; target / filter logits are computed by the owning rank and summed (each entry
    is non-zero on exactly one rank, so the sum is exact), integer counts are all-reduced (bit-exact).
    """
    dev = xq.device
    b = int(xq.shape[0])
    obj = _lib.require_cuda(obj, torch.int64, 'obj')
    filt_ptr = _lib.require_cuda(filt_ptr, torch.int64, 'filt_ptr')
    filt_idx = _lib.require_cuda(filt_idx, torch.int32, 'filt_idx')
    p = _lib.ptr
    if precision == 'fp32':
        if group is not None or all_ent is None or bias is None:
            raise ValueError("precision='fp32' takes the unsharded fp32 all_ent / bias")
        thr, s_filt, gt, eq = _filtered_rank_fp32(xq, all_ent, bias, obj, filt_ptr, filt_idx, count_eq)
    elif precision == 'bf16':
        if table is None:
            table = EntityTable(all_ent, bias)
        q_bf16 = pack_queries(xq)
        # (query, entity) pairs whose logits are needed exactly: the targets, then every filtered positive
        lens = (filt_ptr[1:] - filt_ptr[:-1])
        rows = torch.arange(b, device=dev, dtype=torch.int32)
        pair_q = torch.cat([rows, torch.repeat_interleave(rows, lens)])
        pair_e = torch.cat([obj.to(torch.int32), filt_idx])
        if group is None:
            s_pairs = pair_scores(q_bf16, table, pair_q, pair_e)
        else:
            import torch.distributed as dist
            local = (pair_e >= n_offset) & (pair_e < n_offset + table.n)
            sel = torch.nonzero(local).squeeze(1)
            s_pairs = torch.zeros((pair_q.numel(),), dtype=torch.float32, device=dev)
            s_pairs[sel] = pair_scores(q_bf16, table, pair_q[sel].contiguous(), (pair_e[sel] - n_offset).contiguous())
            dist.all_reduce(s_pairs, group=group)
        thr = s_pairs[:b].contiguous()
        s_filt = s_pairs[b:].contiguous()
        gt = torch.zeros((b,), dtype=torch.int32, device=dev)
        eq = torch.zeros((b,), dtype=torch.int32, device=dev) if count_eq else None
        _lib.call('kgc_score_rank', p(q_bf16), p(table.data), b, table.n, table.kpad, p(thr), p(gt), p(eq), _lib.stream())
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(gt, group=group)
            if eq is not None:
                dist.all_reduce(eq, group=group)
    else:
        raise ValueError("precision must be 'bf16' or 'fp32'")
    ranks = torch.empty((b,), dtype=torch.int32, device=dev)
    eq_out = torch.empty((b,), dtype=torch.int32, device=dev) if count_eq else None
    sums = torch.empty((13,), dtype=torch.float64, device=dev)
    _lib.call('kgc_rank_finalize', p(gt), p(eq), p(thr), p(s_filt), p(filt_ptr), p(filt_idx), p(obj), b, p(ranks),
              p(eq_out), p(sums), _lib.stream())
    out = {'ranks': ranks, 'count_gt': ranks - 1, 'thr': thr, 'sums': sums}
    if count_eq:
        out['count_eq'] = eq_out
    return out


def tie_adjusted_sums(ranks, count_eq, ties):
    """sums13 under a tie policy: 'optimistic' rank = 1 + gt (what kgc_rank_finalize gives), 'pessimistic' = 1 + gt + eq,
    'mean' = 1 + gt + eq / 2 (the expectation of the reference's arbitrary tie order, main.py:126)."""
    r = ranks.to(torch.float64)
    if ties == 'pessimistic':
        r = r + count_eq.to(torch.float64)
    elif ties == 'mean':
        r = r + 0.5 * count_eq.to(torch.float64)
    elif ties != 'optimistic':
        raise ValueError("ties must be 'optimistic', 'mean' or 'pessimistic'")
    parts = [torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device), r.sum(),
             (1.0 / r.to(torch.float32)).to(torch.float64).sum()]
    parts += [(r <= k).to(torch.float64).sum() for k in range(1, 11)]
    return torch.stack(parts)


SUM_KEYS = ['count', 'mr', 'mrr'] + ['hits@{}'.format(k) for k in range(1, 11)]


def predict(model, data_iters, graph, data_type, device, mode='tail_batch', precision='bf16', ties='optimistic'):
    """Drop-in for main.py:105-135: same signature and result dict (count, mr, mrr, hits@1..10 sums).
    The encoder runs ONCE per call (in eval mode all_ent / all_rel do not depend on the batch: SURVEY.md
    "next" row N3, results identical); every batch then goes through the fused scorer.

    ``precision='fp32'`` ranks fp32-grade logits (exact mode, see filtered_rank) instead of bf16-rounded ones.  ``ties``:
    how candidates whose logit EQUALS the target's are counted - the reference's double argsort breaks ties arbitrarily
    (main.py:126); 'optimistic' (1 + #greater, default), 'mean' or 'pessimistic'.  The number of tied candidates met is
    returned under 'ties' (0 on tie-free data: then all three policies and the reference agree); a warning is issued once
    when ties occur under the optimistic policy, because a degenerate constant-score model would otherwise score MRR = 1."""
    model.eval()
    with torch.no_grad():
        all_ent, all_rel = model.encode(graph)
        table = EntityTable(all_ent, model.conv2.bias) if precision == 'bf16' else None
        total = torch.zeros((13,), dtype=torch.float64, device=all_ent.device)
        n_tied = torch.zeros((), dtype=torch.int64, device=all_ent.device)
        for trip, fptr, fidx in data_iters['{}_{}'.format(data_type, mode.split('_')[0])].sparse():
            sub, rel, obj = trip[:, 0], trip[:, 1], trip[:, 2]
            xq = model.conv2.query(all_ent.index_select(0, sub), all_rel.index_select(0, rel))
            out = filtered_rank(xq, all_ent, model.conv2.bias, obj, fptr, fidx, count_eq=True, table=table, precision=precision)
            n_tied += out['count_eq'].sum()
            total += out['sums'] if ties == 'optimistic' else tie_adjusted_sums(out['ranks'], out['count_eq'], ties)
        total = total.cpu().tolist()
        n_tied = int(n_tied)
    if n_tied and ties == 'optimistic':
        import warnings
        warnings.warn('kgc_gcn_b200.predict: {} candidates tie with their query\'s target logit; ranks are 1 + #greater '
                      '(optimistic). Pass ties="mean" or "pessimistic" for a tie-aware evaluation.'.format(n_tied), RuntimeWarning)
    res = {k: float(v) for k, v in zip(SUM_KEYS, total)}
    res['ties'] = n_tied
    return res


def evaluate(model, data_iters, graph, params, data_type, mark='Val', hits=(1, 3, 10), precision=None, ties=None):
    """Drop-in for main.py:80-102: (tail + head) / (2 * count), rounded to 5 places.  ``precision`` / ``ties`` as in predict
    (defaults: ``params.eval_precision`` / ``params.eval_ties`` when set, else 'bf16' / 'optimistic')."""
    import logging
    import numpy as np
    precision = precision or getattr(params, 'eval_precision', 'bf16')
    ties = ties or getattr(params, 'eval_ties', 'optimistic')
    tail = predict(model, data_iters, graph, data_type, params.device, mode='tail_batch', precision=precision, ties=ties)
    head = predict(model, data_iters, graph, data_type, params.device, mode='head_batch', precision=precision, ties=ties)
    count = float(tail['count'])
    results = {'mr': np.round((tail['mr'] + head['mr']) / (2 * count), 5),
               'mrr': np.round((tail['mrr'] + head['mrr']) / (2 * count), 5)}
    for k in hits:
        key = 'hits@{}'.format(k)
        results[key] = np.round((tail[key] + head[key]) / (2 * count), 5)
    logging.info('- {} metrics: {}  '.format(mark, '; '.join('{}: {:05.3f}'.format(k, v) for k, v in results.items())))
    return results
