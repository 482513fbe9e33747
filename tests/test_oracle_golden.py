"""Pin oracle/mgcn_oracle.py against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py) and against SURVEY.md Appendix B's Toy known-answers.  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

import mgcn_oracle as orc


@pytest.fixture(scope='module')
def toy(golden_dir):
    with open(os.path.join(golden_dir, 'toy_loader.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def toy_ds(toy_dir):
    return orc.load_dataset(toy_dir)


def test_vocab_and_counts(toy, toy_ds):
    assert toy_ds['entity2id'] == toy['entity2id']
    assert toy_ds['relation2id'] == toy['relation2id']
    assert (toy_ds['num_entity'], toy_ds['num_relation'], toy_ds['num_edge']) == (7, 5, 10)
    assert (toy['num_entity'], toy['num_relation'], toy['num_edge']) == (7, 5, 10)


def test_query_lists(toy, toy_ds):
    for key in ('train', 'valid_tail', 'valid_head', 'test_tail', 'test_head'):
        got = [{'triple': [int(a) for a in q['triple']], 'label': sorted(int(a) for a in q['label'])}
               for q in toy_ds['queries'][key]]
        assert got == toy['triplets'][key], key
    # Appendix B spot checks
    assert toy['triplets']['train'][0] == {'triple': [0, 0, -1], 'label': [1, 2, 3]}
    assert toy['triplets']['train'][9] == {'triple': [6, 8, -1], 'label': [0, 2]}
    assert toy['triplets']['valid_head'][4] == {'triple': [6, 6, 5], 'label': [3, 5]}


def test_graph_build(toy, toy_ds):
    g = orc.build_graph(toy_ds['triples']['train'], 7, 5)
    assert g['edge_index'].tolist() == toy['edge_index']
    assert g['edge_attr'].tolist() == toy['edge_attr']
    assert g['entity'].tolist() == toy['entity']
    np.testing.assert_array_equal(g['edge_norm'], np.asarray(toy['edge_norm'], dtype=np.float32))
    assert toy['edge_index'][0] == [0, 0, 0, 0, 0, 0, 3, 2, 2, 1, 1, 2, 3, 4, 5, 6, 6, 3, 6, 4]
    assert toy['edge_attr'][0] == [0, 0, 0, 1, 2, 3, 1, 2, 3, 3, 5, 5, 5, 6, 7, 8, 6, 7, 8, 8]


def test_labels_and_batches(golden_dir, toy_ds):
    z = np.load(os.path.join(golden_dir, 'toy_batches.npz'))
    q = toy_ds['queries']['train']
    trip, lab = orc.make_batch(q, range(len(q)), 7, lbl_smooth=0.1, training=True)
    np.testing.assert_array_equal(trip, z['train_triple'])
    np.testing.assert_array_equal(lab, z['train_label'])          # includes the 0.9 + 1/7 > 1 quirk
    assert lab.max() > 1.0
    q = toy_ds['queries']['valid_tail']
    trip, lab = orc.make_batch(q, range(len(q)), 7, lbl_smooth=0.1, training=False)
    np.testing.assert_array_equal(trip, z['valid_tail_triple'])
    np.testing.assert_array_equal(lab, z['valid_tail_label'])


def test_in_layer_norms_appendix_b(toy):
    ei = np.asarray(toy['edge_index'])
    n_in = orc.compute_norm(ei[:, :10], 7).numpy()
    n_out = orc.compute_norm(ei[:, 10:], 7).numpy()
    np.testing.assert_allclose(n_in, [.40825, .28868, .40825, 0, 0, 0, 0, .70711, 0, 0], atol=2e-5)
    np.testing.assert_allclose(n_out, [0, 0, 0, 0, 0, 0, .40825, .70711, .57735, .70711], atol=2e-5)


CONV_CASES = ['conv_toy_small', 'conv_toy_eval', 'conv_toy_full', 'conv_synth_hub', 'conv_synth_masks']


def _load_conv(golden_dir, name, dt):
    z = np.load(os.path.join(golden_dir, name + '.npz'))
    w = {k[2:]: torch.from_numpy(z[k]).to(dt) for k in z.files if k.startswith('w.')}
    return z, w


@pytest.mark.parametrize('name', CONV_CASES)
def test_conv_forward_backward_f64(golden_dir, name):
    """The restatement, run in float64, equals the reference run in float64 to round-off."""
    dt = torch.float64
    z, w = _load_conv(golden_dir, name, dt)
    ei, et = torch.from_numpy(z['edge_index']), torch.from_numpy(z['edge_type'])
    m_in = torch.from_numpy(z['mask_in']) if 'mask_in' in z.files else None
    m_out = torch.from_numpy(z['mask_out']) if 'mask_out' in z.files else None
    ent, rel, grads, aux = orc.conv_fwd_bwd(
        torch.from_numpy(z['x']).to(dt), ei, et, torch.from_numpy(z['edge_embs']).to(dt),
        torch.from_numpy(z['rels']).to(dt), w, torch.from_numpy(z['g_ent']), torch.from_numpy(z['g_rel']),
        mask_in=m_in, mask_out=m_out, training=bool(z['training']))
    np.testing.assert_array_equal(aux['norm_in'].float().numpy(), z['norm_in'])
    np.testing.assert_array_equal(aux['norm_out'].float().numpy(), z['norm_out'])
    np.testing.assert_allclose(ent.numpy(), z['all_ent.f64'], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(rel.numpy(), z['all_rel.f64'], rtol=1e-10, atol=1e-12)
    names = {'entity_embedding': 'x', 'edge_embeddings': 'edge_embs', 'relation_embedding': 'rels'}
    for k, g in grads.items():
        gk = names.get(k, 'w.' + k[len('conv1.'):])
        ref = z['grad.{}.f64'.format(gk)]
        scale = max(1e-30, float(np.abs(ref).max()))
        assert float(np.abs(g.numpy() - ref).max()) / scale < 1e-9, k


@pytest.mark.parametrize('name', ['conv_toy_full', 'conv_synth_hub'])
def test_conv_f32_error_budget(golden_dir, name):
    """two fp32 evaluations differ from fp64 truth by ~1e-6..1e-4 (max-norm relative);
    record that the restatement's fp32 error is of the same size as the reference's own."""
    z, w = _load_conv(golden_dir, name, torch.float32)
    ent, rel, _ = orc.conv_forward(torch.from_numpy(z['x']), torch.from_numpy(z['edge_index']),
                                   torch.from_numpy(z['edge_type']), torch.from_numpy(z['edge_embs']),
                                   torch.from_numpy(z['rels']), w, training=True)
    truth = z['all_ent.f64']
    err_ref = np.abs(z['all_ent.f32'] - truth).max() / np.abs(truth).max()
    err_orc = np.abs(ent.numpy() - truth).max() / np.abs(truth).max()
    assert err_orc <= max(4 * err_ref, 2e-6)


def test_rank_identity(golden_dir):
    """main.py:122-126 through the reference's own predict() == oracle dense form == 1 + count_gt (tie-free)."""
    z = np.load(os.path.join(golden_dir, 'rank_case.npz'))
    pred, label, obj = torch.from_numpy(z['pred']), torch.from_numpy(z['label']), torch.from_numpy(z['obj'])
    ranks = orc.filtered_ranks_dense(pred, label, obj).numpy()
    sums = orc.metric_sums(ranks)
    for k, v in sums.items():
        assert abs(v - float(z['res.' + k])) <= 1e-4 * max(1.0, abs(v)), k
    ptr = np.zeros(pred.size(0) + 1, dtype=np.int64)
    idx = []
    for q in range(pred.size(0)):
        idx.extend(np.nonzero(z['label'][q])[0].tolist())
        ptr[q + 1] = len(idx)
    gt, eq = orc.rank_counts(z['pred'], ptr, np.asarray(idx), z['obj'])
    assert eq.sum() == 0
    np.testing.assert_array_equal(1 + gt, ranks)


def test_toy_model_scores_and_metrics(golden_dir, toy_ds):
    """Whole-model eval forward (GCN + ConvE front end + scoring tail) restated == reference, float64."""
    z = np.load(os.path.join(golden_dir, 'toy_model.npz'))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    g = orc.build_graph(toy_ds['triples']['train'], 7, 5)
    dt = torch.float64
    w = {k[len('conv1.'):]: v.to(dt) for k, v in sd.items() if k.startswith('conv1.') and v.is_floating_point()}
    ent, rel, _ = orc.conv_forward(sd['entity_embedding'].to(dt), torch.from_numpy(g['edge_index']),
                                   torch.from_numpy(g['edge_attr'][0]), sd['edge_embeddings'].to(dt),
                                   sd['relation_embedding'].to(dt), w, training=False)
    dec = {k[len('conv2.'):]: v.to(dt) for k, v in sd.items() if k.startswith('conv2.') and v.is_floating_point()}
    tails, heads = None, None
    for mode in ('tail', 'head'):
        trip = torch.from_numpy(z['eval.{}.triple'.format(mode)])
        xq = orc.conve_front(ent[trip[:, 0]], rel[trip[:, 1]], dec, 10, 20)
        sc = orc.score_tail(xq, ent, dec['bias'])
        np.testing.assert_allclose(sc.numpy(), z['eval.{}.score.f64'.format(mode)], rtol=1e-9, atol=1e-12)
        qs = toy_ds['queries']['valid_' + mode]
        lab = torch.from_numpy(np.stack([orc.make_label(q['label'], 7) for q in qs]))
        ranks = orc.filtered_ranks_dense(sc.float(), lab, trip[:, 2]).numpy()
        if mode == 'tail':
            tails = orc.metric_sums(ranks)
        else:
            heads = orc.metric_sums(ranks)
    with open(os.path.join(golden_dir, 'toy_metrics.json')) as f:
        gold = json.load(f)
    for k, v in gold['tail'].items():
        assert abs(tails[k] - v) < 1e-4, k
    for k, v in gold['head'].items():
        assert abs(heads[k] - v) < 1e-4, k
    comb = orc.combine_metrics(tails, heads)
    for k, v in gold['evaluate'].items():
        assert abs(float(comb[k]) - v) < 1e-5, k


def test_stable_csr():
    keys = np.array([3, 1, 3, 0, 1, 3, 5])
    perm, rowptr = orc.stable_csr(keys, 7)
    assert perm.tolist() == [3, 1, 4, 0, 2, 5, 6]
    assert rowptr.tolist() == [0, 1, 3, 3, 6, 6, 7, 7]


@pytest.mark.parametrize('smooth', [0.0, 0.1])
def test_sparse_bce_matches_torch_bce_on_dense_labels(toy_ds, smooth):
    """N1 pin: the sparse-positives loss restatement == the reference's loss path (dense label of data_loader.py:34-43,
    nn.BCELoss of model.py:42-44, autograd through the sigmoid of model.py:179).  Without smoothing on the Toy training
    queries; with smoothing on 50 entities (on Toy 0.9 + 1/7 > 1 and BCELoss rejects the label).
    A saturated-logit pair exercises the clamps of torch's BCE.  float64: 1e-12, float32: 1e-6."""
    if smooth:
        n, rng = 50, np.random.default_rng(3)
        queries = [{'triple': (i, 0, -1), 'label': set(rng.choice(n, size=1 + i % 5, replace=False).tolist())} for i in range(12)]
    else:
        queries, n = toy_ds['queries']['train'], toy_ds['num_entity']
    ptr, idx = orc.queries_to_csr(queries)
    qid = list(range(len(queries)))[::-1]
    _, lab = orc.make_batch(queries, qid, n, smooth, training=True)
    pos, add = float(lab.max()), float(lab.min())               # the two float32 label values (1, 0 without smoothing)
    gen = torch.Generator().manual_seed(5)
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 1e-6)):
        z = torch.randn((len(qid), n), generator=gen, dtype=torch.float64) * 4.0
        z[0, 0], z[1, 1] = 40.0, -120.0
        z = z.to(dtype).requires_grad_(True)
        pred = torch.sigmoid(z)
        ref = torch.nn.BCELoss()(pred, torch.from_numpy(lab).to(dtype))
        ref.backward()
        loss, d_logit, d_bias = orc.bce_1n_sparse(pred.detach(), ptr, idx, qid, pos, add)
        assert abs(float(loss) - float(ref.detach())) <= tol * abs(float(ref.detach()))
        scale = float(z.grad.abs().max())
        assert float((d_logit - z.grad).abs().max()) <= tol * scale
        assert float((d_bias - z.grad.sum(0)).abs().max()) <= 10 * tol * scale
    mask = orc.label_mask(ptr, idx, qid, n)                     # bit form of the same labels
    cols = np.arange(n)
    dense = ((mask[:, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(dense, lab == pos)


def test_chunked_float64_oracle_matches_the_pinned_one():
    """conv_fwd_bwd_big (edge-chunked, hand-derived backward: the checker of the Wikidata5M-shape GPU tests) against
    conv_fwd_bwd (reference operation order + autograd, pinned to the reference's golden vectors above): every output and
    gradient to float64 rounding, with dropout masks, with and without the optional bias, chunk boundaries inside halves."""
    import torch
    N, R, E, D, Do = 300, 4, 1500, 12, 20
    tri = orc.synthetic_triples(N, R, E, 3)
    g = orc.build_graph(tri, N, R)
    p = orc.conv_params(N, R, E, D, Do, seed=1)
    gen = torch.Generator().manual_seed(2)
    g_ent, g_rel = torch.randn(N, Do, generator=gen), torch.randn(2 * R, Do, generator=gen)
    m_in = (torch.rand(N, Do, generator=gen) > 0.1).float()
    m_out = (torch.rand(N, Do, generator=gen) > 0.1).float()
    ei, et = torch.from_numpy(g['edge_index']), torch.from_numpy(g['edge_attr'][0])
    w = dict(p['w'])
    w['ent_bn.weight'] = torch.rand(Do, generator=gen) + 0.5
    w['ent_bn.bias'] = torch.randn(Do, generator=gen)
    for bias, masks in ((None, (m_in, m_out)), (torch.randn(Do, generator=gen), (m_in, m_out)), (None, (None, None))):
        w['bias'] = bias
        w64 = {k: (v.double() if v is not None else None) for k, v in w.items()}
        ent, rel, grads, _ = orc.conv_fwd_bwd(p['x'].double(), ei, et, p['edge_embs'].double(), p['rels'].double(), w64,
                                              g_ent, g_rel, mask_in=masks[0], mask_out=masks[1])
        big = orc.conv_fwd_bwd_big(p['x'], ei, et, p['edge_embs'], p['rels'], w, g_ent, g_rel, mask_in=masks[0],
                                   mask_out=masks[1], chunk=333, d_ee_check=grads['edge_embeddings'])
        scale = lambda t: float(t.abs().max()) + 1e-30                                   # noqa: E731
        assert float((big['all_ent'] - ent).abs().max()) < 1e-12 * scale(ent)
        assert float((big['all_rel'] - rel).abs().max()) < 1e-12 * scale(rel)
        assert big['d_ee_max_abs_err'] < 1e-12 * big['d_ee_max_abs']
        for k, v in grads.items():
            if k == 'edge_embeddings':
                continue
            tol = 1e-12 * scale(v) if k != 'conv1.bias' else 1e-12 * scale(grads['conv1.ent_bn.weight'])   # ~0 in training mode
            assert float((big[k] - v).abs().max()) < tol, k
