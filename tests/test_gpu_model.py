"""Whole-model parity on data/Toy (BASELINE.json configs[0], the correctness gate): the drop-in MGCN loaded
with the reference's state dict reproduces the reference's eval scores and train-step gradients."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def params(**kw):
    base = dict(gcn_in_dim=20, gcn_out_dim=200, gcn_drop=0.0, hidden_drop=0.0, feat_drop=0.0, k_w=10, k_h=20,
                num_filter=2, kernel_size=7, bias=False, lbl_smooth=0.0, batch_size=128)
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.fixture(scope='module')
def toy(golden_dir):
    import kgc_gcn_b200 as k
    cwd = os.getcwd()
    os.chdir(golden_dir)                      # the loader resolves data/<dataset> relative to the cwd (data_loader.py:57)
    try:
        dl = k.DataLoader('Toy', params())
    finally:
        os.chdir(cwd)
    dl.graph.to('cuda')
    z = np.load(os.path.join(golden_dir, 'toy_model.npz'))
    m = k.MGCN(dl.num_entity, dl.num_relation, dl.num_edge, params())
    sd = {kk[3:]: torch.from_numpy(z[kk]) for kk in z.files if kk.startswith('sd.')}
    m.load_state_dict(sd, strict=True)        # the reference's checkpoint keys load unchanged
    m.conv1.drop.p = 0.0
    return k, dl, m.cuda(), z


def test_loader_matches_reference(toy, golden_dir):
    k, dl, m, z = toy
    with open(os.path.join(golden_dir, 'toy_loader.json')) as f:
        gold = json.load(f)
    assert dl.entity2id == gold['entity2id'] and dl.relation2id == gold['relation2id']
    assert (dl.num_entity, dl.num_relation, dl.num_edge) == (7, 5, 10)
    g = dl.graph
    assert g.edge_index.cpu().tolist() == gold['edge_index']
    assert g.edge_attr.cpu().tolist() == gold['edge_attr']
    assert g.entity.cpu().tolist() == gold['entity']
    np.testing.assert_array_equal(g.edge_norm.cpu().numpy(), np.asarray(gold['edge_norm'], dtype=np.float32))
    for key in gold['triplets']:
        got = [{'triple': list(q['triple']), 'label': sorted(q['label'])} for q in dl.triplets[key]]
        assert got == gold['triplets'][key]


def test_eval_scores(toy):
    k, dl, m, z = toy
    m.eval()
    for mode in ('tail', 'head'):
        trip = torch.from_numpy(z['eval.{}.triple'.format(mode)]).cuda()
        with torch.no_grad():
            sc = m(trip[:, 0], trip[:, 1], dl.graph)
        truth = z['eval.{}.score.f64'.format(mode)]
        np.testing.assert_allclose(sc.cpu().numpy(), truth, rtol=2e-5, atol=2e-6)


def test_train_step_grads(toy):
    k, dl, m, z = toy
    m.train()
    m.zero_grad()
    trip = torch.from_numpy(z['train.triple']).cuda()
    lab = torch.from_numpy(z['train.label']).cuda()
    pred = m(trip[:, 0], trip[:, 1], dl.graph)
    loss = m.loss(pred, lab)
    loss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), z['train.pred.f64'], rtol=2e-5, atol=2e-6)
    # BCE of saturated predictions amplifies the fp32 round-off of pred (log(1 - p), p -> 1): 2e-4 relative budget
    assert abs(float(loss.detach()) - float(z['train.loss.f64'])) < 2e-4 * float(z['train.loss.f64'])
    bad = {}
    for name, prm in m.named_parameters():
        truth = z['train.grad.' + name]
        scale = max(float(np.abs(truth).max()), 1e-30)
        err = float(np.abs(prm.grad.cpu().numpy().astype(np.float64) - truth).max())
        # + absolute floor: gradients that are mathematically zero (a BN bias feeding another BN) are pure fp32 noise
        if err > 5e-5 * scale + 2e-7:
            bad[name] = (err, scale)
    assert not bad, bad


def test_iterators(toy):
    k, dl, m, z = toy
    it = dl.get_data_loaders(4, 0, params(lbl_smooth=0.0))
    assert set(it) == {'train', 'valid_head', 'valid_tail', 'test_head', 'test_tail'}
    assert len(it['train']) == 5 and len(it['valid_tail']) == 2
    seen = 0
    for trip, lab in it['train']:
        assert trip.is_cuda and lab.is_cuda and trip.dtype == torch.int64 and lab.dtype == torch.float32
        assert lab.shape == (trip.shape[0], 7)
        seen += trip.shape[0]
    assert seen == 17


def test_graphed_train_step_matches_eager(toy):
    """GraphedTrainStep (one CUDA-graph replay per step) follows the eager loop of main.py:57-71 step for step."""
    import copy
    k, dl, m0, z = toy
    prm = params(lbl_smooth=0.0)
    ds = dl._get_dataset('train', prm)
    qids = [list(range(0, 16)), list(range(1, 17)), list(range(0, 16))]
    losses = {}
    finals = {}
    for mode in ('eager', 'graph'):
        torch.manual_seed(3)
        m = copy.deepcopy(m0).train()
        m.conv1.drop.p = 0.0
        opt = torch.optim.Adam(m.parameters(), lr=1e-2, capturable=(mode == 'graph'))
        step = k.GraphedTrainStep(m, opt, dl.graph, ds, 16, warmup=0) if mode == 'graph' else None
        ls = []
        for q in qids:
            if step is not None:
                ls.append(float(step(q).item()))
            else:
                trip, lab = ds.build_batch(q, 'cuda')
                opt.zero_grad()
                loss = m.loss(m(trip[:, 0], trip[:, 1], dl.graph), lab)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                opt.step()
                ls.append(float(loss.item()))
        losses[mode] = ls
        finals[mode] = {n: p.detach().clone() for n, p in m.named_parameters()}
    for a, b in zip(losses['eager'], losses['graph']):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (losses)
    # Adam divides by sqrt(v): on parameters whose true gradient is zero (a BatchNorm scale feeding another BatchNorm)
    # it turns rounding noise into O(lr) steps, so only parameters with a real gradient signal are compared
    for n in ('entity_embedding', 'relation_embedding', 'edge_embeddings', 'conv1.in_weight', 'conv1.loop_weight',
              'conv2.fc.weight', 'conv2.bias'):
        assert torch.allclose(finals['eager'][n], finals['graph'][n], rtol=1e-3, atol=1e-4), n
