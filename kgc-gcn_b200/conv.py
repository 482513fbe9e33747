"""MGCNConv: the relation-aware graph convolution (reference model.py:47-127), B200 path.

Same constructor, parameter names, forward signature and return values as the reference class; the
arithmetic runs in libkgc_b200.so (K1-K4) plus plain fp32 GEMMs (TF32 off, like the reference gets):

    agg_h[i]  = sum_{e in half h, dst_e = i} norm_e * x[src_e] (.) rel+[type_e] (.) ee[e]     K2
    res_h     = agg_h @ W_h ; res_loop = x @ (diag(loop_rel (.) loop_edge) W_loop)               GEMM
    all_ent   = tanh(BN((drop(res_in) + drop(res_out) + res_loop) / 3 [+ bias]))                  K4
    all_rel   = (rel+ @ W_rel)[:-1]

(aggregate-then-transform is exact algebra because the message transform is linear).
Backward is hand-derived and runs K3/K4-backward; it is deterministic.
"""
import torch
import torch.nn as nn

from . import _lib
from .plan import get_plan


def get_param(shape):
    """utils.get_param (reference utils.py:113-118): xavier-uniform Parameter."""
    param = nn.Parameter(torch.empty(*shape))
    nn.init.xavier_uniform_(param.data)
    return param


def _mm(a, b, out):
    return torch.mm(a, b, out=out)


class _ConvFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, rels, ee, w_in, w_out, w_loop, w_rel, loop_rel, loop_edge, gamma, beta, bias,
                plan, mask_in, mask_out, keep_scale, training, running_mean, running_var, eps):
        N, D = x.shape
        Dout = w_in.shape[1]
        p, st = _lib.ptr, _lib.stream
        x = _lib.require_cuda(x, torch.float32, 'x')
        ee = _lib.require_cuda(ee, torch.float32, 'edge_embs')
        if ee.shape[0] != plan.num_edges2 or N != plan.num_nodes:
            raise ValueError('edge_embs / x do not match the graph plan')
        relp = torch.cat([rels, loop_rel], 0).contiguous()          # model.py:86
        if relp.shape[0] != plan.num_types:
            raise ValueError('rels_embs rows + 1 must equal the number of edge types of the plan')

        agg = torch.empty((2, N, D), dtype=torch.float32, device=x.device)

        def level0(items, n_items, out_final, part):
            _lib.call('kgc_agg_fwd', p(x), p(relp), p(ee), p(plan.rec_dst), p(items), n_items, p(out_final), p(part),
                      D, st())
        plan.run_reduction(plan.fwd, level0, agg, D, tag='f')

        v = (loop_rel * loop_edge).reshape(D, 1)                     # self-loop: (x . lr . le) @ W = x @ (diag(v) W)
        w_loop_s = v * w_loop
        res3 = plan.scratch('res3', (3, N, Dout))
        _mm(agg[0], w_in, res3[0])
        _mm(agg[1], w_out, res3[1])
        _mm(x, w_loop_s, res3[2])

        nb = int(_lib.lib().kgc_tail_num_blocks(N))
        partials = plan.scratch('colpart', (nb, 2, Dout), torch.float64)
        pre = torch.empty((N, Dout), dtype=torch.float32, device=x.device)
        stats = torch.empty((3, Dout), dtype=torch.float32, device=x.device)
        all_ent = torch.empty((N, Dout), dtype=torch.float32, device=x.device)
        _lib.call('kgc_tail_fwd', p(res3), p(mask_in), p(mask_out), float(keep_scale), p(bias), N, Dout, p(pre),
                  p(partials), st())
        _lib.call('kgc_colstats_finalize', p(partials), nb, N, Dout, float(eps), int(training), p(running_mean),
                  p(running_var), p(stats), st())
        _lib.call('kgc_tail_apply', p(pre), p(stats), p(gamma), p(beta), N, Dout, p(all_ent), st())
        all_rel = torch.mm(relp, w_rel)[:-1]                          # model.py:107

        ctx.plan, ctx.training, ctx.keep_scale, ctx.has_bias = plan, bool(training), float(keep_scale), bias is not None
        ctx.save_for_backward(x, relp, ee, w_in, w_out, w_loop, w_rel, loop_rel, loop_edge, gamma, agg, pre, all_ent,
                              stats, mask_in, mask_out, w_loop_s)
        ctx.mark_non_differentiable(stats)
        return all_ent, all_rel, stats

    @staticmethod
    def backward(ctx, g_ent, g_rel, _g_stats):
        (x, relp, ee, w_in, w_out, w_loop, w_rel, loop_rel, loop_edge, gamma, agg, pre, all_ent, stats, mask_in,
         mask_out, w_loop_s) = ctx.saved_tensors
        plan = ctx.plan
        N, D = x.shape
        Dout = w_in.shape[1]
        T = relp.shape[0]
        p, st = _lib.ptr, _lib.stream
        dev = x.device
        if g_ent is None:
            g_ent = torch.zeros_like(all_ent)
        g_ent = g_ent.contiguous()

        # ---- K4 backward: tanh, BatchNorm, /3, dropout
        nb = int(_lib.lib().kgc_tail_num_blocks(N))
        partials = plan.scratch('colpart', (nb, 2, Dout), torch.float64)
        sums = torch.empty((2, Dout), dtype=torch.float32, device=dev)
        d_res3 = plan.scratch('d_res3', (3, N, Dout))
        _lib.call('kgc_tail_bwd_reduce', p(g_ent), p(all_ent), p(pre), p(stats), N, Dout, p(partials), st())
        _lib.call('kgc_colsum_finalize', p(partials), nb, Dout, p(sums), st())
        _lib.call('kgc_tail_bwd_apply', p(g_ent), p(all_ent), p(pre), p(stats), p(gamma), p(sums), p(mask_in),
                  p(mask_out), ctx.keep_scale, int(ctx.training), N, Dout, p(d_res3), st())
        d_beta, d_gamma = sums[0], sums[1]
        d_bias = d_res3[2].sum(0) * 3.0 if ctx.has_bias else None

        # ---- dense transforms (fp32 GEMMs)
        g3 = plan.scratch('g3', (3, N, D))
        _mm(d_res3[0], w_in.t(), g3[0])
        _mm(d_res3[1], w_out.t(), g3[1])
        _mm(d_res3[2], w_loop_s.t(), g3[2])
        d_w_in = torch.mm(agg[0].t(), d_res3[0])
        d_w_out = torch.mm(agg[1].t(), d_res3[1])
        m_loop = torch.mm(x.t(), d_res3[2])                           # [D, Dout]
        v = (loop_rel * loop_edge).reshape(D, 1)
        d_w_loop = v * m_loop
        d_v = (m_loop * w_loop).sum(1).reshape(1, D)
        d_loop_edge = d_v * loop_rel
        d_loop_rel = d_v * loop_edge

        # ---- K3: d_x (+ self-loop term) and d_ee over src-sorted rows, d_rel over type-sorted rows
        d_x = torch.empty((N, D), dtype=torch.float32, device=dev)
        d_ee = torch.empty_like(ee)
        d_relp = torch.empty((T, D), dtype=torch.float32, device=dev)

        def level0_src(items, n_items, out_final, part):
            _lib.call('kgc_agg_bwd_src', p(x), p(relp), p(ee), p(g3), p(plan.rec_src), p(items), n_items, N,
                      plan.num_edges2, p(d_ee), p(out_final), p(part), D, st())
        plan.run_reduction(plan.bwd_src, level0_src, d_x, D, addend=g3[2], tag='s')

        def level0_rel(items, n_items, out_final, part):
            _lib.call('kgc_agg_bwd_rel', p(x), p(ee), p(g3), p(plan.rec_type), p(items), n_items, N, plan.num_edges2,
                      p(out_final), p(part), D, st())
        plan.run_reduction(plan.bwd_rel, level0_rel, d_relp, D, tag='r')

        # ---- relation transform (model.py:107)
        if g_rel is not None:
            g_rel_pad = torch.cat([g_rel, g_rel.new_zeros((1, Dout))], 0)
            d_relp = d_relp + torch.mm(g_rel_pad, w_rel.t())
            d_w_rel = torch.mm(relp.t(), g_rel_pad)
        else:
            d_w_rel = torch.zeros_like(w_rel)
        d_rels = d_relp[:-1]
        d_loop_rel = d_loop_rel + d_relp[-1:]
        return (d_x, d_rels, d_ee, d_w_in, d_w_out, d_w_loop, d_w_rel, d_loop_rel, d_loop_edge, d_gamma, d_beta, d_bias,
                None, None, None, None, None, None, None, None)


class MGCNConv(nn.Module):
    """Drop-in for the reference MGCNConv (model.py:47-127): same ctor, parameters, forward."""

    def __init__(self, in_channels, out_channels, num_relations, bias=False, dropout=0.1, **kwargs):
        super(MGCNConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.aggr = 'add'

        self.ent_bn = nn.BatchNorm1d(out_channels)
        self.drop = nn.Dropout(dropout)
        self.act = torch.tanh

        self.loop_weight = get_param((in_channels, out_channels))
        self.in_weight = get_param((in_channels, out_channels))
        self.out_weight = get_param((in_channels, out_channels))
        self.rels_weight = get_param((in_channels, out_channels))
        self.loop_rel = get_param((1, in_channels))
        self.loop_edge = get_param((1, in_channels))

        if bias is True:
            self.register_parameter('bias', nn.Parameter(torch.zeros(out_channels)))
        else:
            self.register_parameter('bias', None)
        self._forced_masks = None

    def set_dropout_masks(self, mask_in, mask_out):
        """Inject the two keep masks ([N, Dout], 0/1) the next training forward uses instead of drawing
        them - the parity tests replay the reference's own Bernoulli draws this way.  None clears."""
        self._forced_masks = None if mask_in is None else (mask_in, mask_out)

    def compute_norm(self, edge_index, num_ent):
        """model.py:72-80 for one half's edge_index [2, E_h]; evaluated by K1."""
        ei = torch.cat([edge_index, edge_index], 1).contiguous()
        et = torch.zeros(ei.size(1), dtype=torch.int64, device=ei.device)
        return get_plan(ei, et, num_ent, 1).norm[:edge_index.size(1)].clone()

    def _masks(self, n_rows, device):
        p = self.drop.p
        if not self.training or p == 0.0:
            return None, None, 1.0
        if self._forced_masks is not None:
            m_in, m_out = self._forced_masks
            return (m_in.to(device=device, dtype=torch.uint8).contiguous(),
                    m_out.to(device=device, dtype=torch.uint8).contiguous(), 1.0 / (1.0 - p))
        shape = (n_rows, self.out_channels)
        m_in = torch.empty(shape, dtype=torch.uint8, device=device).bernoulli_(1.0 - p)
        m_out = torch.empty(shape, dtype=torch.uint8, device=device).bernoulli_(1.0 - p)
        return m_in, m_out, 1.0 / (1.0 - p)

    def forward(self, x, edge_index, edge_type, edge_norm, edge_embs, rels_embs, size=None):
        # edge_norm and size are accepted and ignored, exactly as the reference does (model.py:82)
        num_ent = x.size(0)
        plan = get_plan(edge_index, edge_type, num_ent, rels_embs.size(0) + 1)
        m_in, m_out, keep_scale = self._masks(num_ent, x.device)
        bn = self.ent_bn
        use_batch_stats = self.training or bn.running_mean is None
        all_ent, all_rel, stats = _ConvFn.apply(
            x, rels_embs, edge_embs, self.in_weight, self.out_weight, self.loop_weight, self.rels_weight,
            self.loop_rel, self.loop_edge, bn.weight, bn.bias, self.bias, plan, m_in, m_out, keep_scale,
            use_batch_stats, bn.running_mean, bn.running_var, bn.eps)
        if self.training and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():           # nn.BatchNorm1d bookkeeping: momentum update with the UNBIASED variance
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                unbias = float(num_ent) / float(max(num_ent - 1, 1))
                bn.running_mean.mul_(1.0 - mom).add_(stats[0], alpha=mom)
                bn.running_var.mul_(1.0 - mom).add_(stats[1], alpha=mom * unbias)
        return all_ent, all_rel

    def __repr__(self):
        return '{}({}, {}, num_relations={})'.format(self.__class__.__name__, self.in_channels, self.out_channels,
                                                     self.num_relations)
