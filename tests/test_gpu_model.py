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
        # + absolute floor: gradients that are mathematically zero up to the BatchNorm eps (a BN scale / bias feeding another
        # BN through a linear map: conv2.bn0.weight is 3e-7 in the fp64 reference) are pure fp32 round-off
        if err > 5e-5 * scale + 5e-7:
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
    for mode in ('eager', 'graph', 'graph_clipadam'):
        torch.manual_seed(3)
        m = copy.deepcopy(m0).train()
        m.conv1.drop.p = 0.0
        if mode == 'graph_clipadam':                 # K9: clip + Adam inside the optimiser, device-side step / lr
            opt = k.ClipAdam(m.parameters(), lr=1e-2, max_norm=1.0)
        else:
            opt = torch.optim.Adam(m.parameters(), lr=1e-2, capturable=(mode == 'graph'))
        step = k.GraphedTrainStep(m, opt, dl.graph, ds, 16, warmup=0) if mode != 'eager' else None
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
    for other in ('graph', 'graph_clipadam'):
        for a, b in zip(losses['eager'], losses[other]):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (losses)
    # Adam divides by sqrt(v): on parameters whose true gradient is zero (a BatchNorm scale feeding another BatchNorm)
    # it turns rounding noise into O(lr) steps, so only parameters with a real gradient signal are compared
    for n in ('entity_embedding', 'relation_embedding', 'edge_embeddings', 'conv1.in_weight', 'conv1.loop_weight',
              'conv2.fc.weight', 'conv2.bias'):
        assert torch.allclose(finals['eager'][n], finals['graph'][n], rtol=1e-3, atol=1e-4), n
        assert torch.allclose(finals['eager'][n], finals['graph_clipadam'][n], rtol=1e-3, atol=1e-4), n


@pytest.mark.parametrize('B,C,H,W', [(128, 200, 14, 14), (7, 5, 2, 2), (33, 16, 6, 6)])
def test_conve_bn_relu_dropout_kernels(B, C, H, W):
    """K7: bn1 + relu + feature dropout of ConvE (model.py:168-170) against nn.BatchNorm2d / F.relu / an explicit keep mask:
    outputs, gradients and the running-statistics update (training and evaluation mode)."""
    import kgc_gcn_b200 as k
    prm = SimpleNamespace(gcn_out_dim=8, hidden_drop=0.0, feat_drop=0.0, k_w=2, k_h=4, num_filter=C, kernel_size=3, bias=False)
    g = torch.Generator().manual_seed(B + C)
    conve = k.ConvE(prm, 10).cuda()
    ref_bn = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        conve.bn1.weight.copy_(torch.rand(C, generator=g) + 0.5); conve.bn1.bias.copy_(torch.randn(C, generator=g) * 0.1)
        ref_bn.weight.copy_(conve.bn1.weight); ref_bn.bias.copy_(conve.bn1.bias)
    x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.3).cuda()
    dy = torch.randn(B, C, H, W, generator=g).cuda()
    for mode in ('train', 'eval'):
        conve.train(mode == 'train'); ref_bn.train(mode == 'train')
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya = conve._bn1_relu_drop(xa)
        yb = torch.relu(ref_bn(xb))
        ya.backward(dy); yb.backward(dy)
        assert float((ya - yb).detach().abs().max()) <= 2e-5 * max(1.0, float(yb.detach().abs().max()))
        assert float((xa.grad - xb.grad).abs().max()) <= 2e-5 * max(1.0, float(xb.grad.abs().max()))
        assert float((conve.bn1.weight.grad - ref_bn.weight.grad).abs().max()) <= 2e-5 * max(1.0, float(ref_bn.weight.grad.abs().max()))
        assert float((conve.bn1.bias.grad - ref_bn.bias.grad).abs().max()) <= 2e-5 * max(1.0, float(ref_bn.bias.grad.abs().max()))
        assert torch.allclose(conve.bn1.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(conve.bn1.running_var, ref_bn.running_var, rtol=1e-5, atol=1e-6)
        assert int(conve.bn1.num_batches_tracked) == int(ref_bn.num_batches_tracked)
        conve.bn1.weight.grad = None; conve.bn1.bias.grad = None; ref_bn.weight.grad = None; ref_bn.bias.grad = None
    # bn0: the one-channel BatchNorm2d over the input image (no ReLU, no dropout) on the same kernels
    ref0 = torch.nn.BatchNorm2d(1).cuda()
    img = (torch.randn(B, 1, 2 * prm.k_w, prm.k_h, generator=g) + 0.2).cuda()
    dimg = torch.randn(B, 1, 2 * prm.k_w, prm.k_h, generator=g).cuda()
    for mode in ('train', 'eval'):
        conve.train(mode == 'train'); ref0.train(mode == 'train')
        ia, ib = img.clone().requires_grad_(True), img.clone().requires_grad_(True)
        oa, ob = conve._bn0(ia), ref0(ib)
        oa.backward(dimg); ob.backward(dimg)
        assert float((oa - ob).detach().abs().max()) <= 2e-5 * max(1.0, float(ob.detach().abs().max()))
        assert float((ia.grad - ib.grad).abs().max()) <= 2e-5 * max(1.0, float(ib.grad.abs().max()))
        assert torch.allclose(conve.bn0.weight.grad, ref0.weight.grad, rtol=2e-4, atol=2e-4)
        assert torch.allclose(conve.bn0.bias.grad, ref0.bias.grad, rtol=2e-4, atol=2e-4)
        assert torch.allclose(conve.bn0.running_mean, ref0.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(conve.bn0.running_var, ref0.running_var, rtol=1e-5, atol=1e-6)
        conve.bn0.weight.grad = None; conve.bn0.bias.grad = None; ref0.weight.grad = None; ref0.bias.grad = None
    # dropout: every output is either 0 or the un-dropped value / (1 - p); the backward uses the same keep mask
    p = 0.3
    conve.feature_drop.p = p
    conve.train(); ref_bn.train()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = conve._bn1_relu_drop(xa)
    yb = torch.relu(ref_bn(xb))
    keep = torch.where(yb > 0, ya / yb.clamp_min(1e-30), torch.zeros_like(ya)).detach()
    pos = yb > 1e-3
    vals = keep[pos]
    assert bool(((vals.abs() < 1e-6) | ((vals - 1 / (1 - p)).abs() < 1e-3)).all())
    frac = float((vals > 0.5).float().mean())
    if vals.numel() > 10000:
        assert abs(frac - (1 - p)) < 0.01, frac
    ya.backward(dy)
    (yb * torch.where(keep > 0.5, torch.full_like(keep, 1 / (1 - p)), torch.zeros_like(keep))).backward(dy)
    assert float((xa.grad - xb.grad).abs().max()) <= 5e-5 * max(1.0, float(xb.grad.abs().max()))
    # a second step draws a different mask
    ya2 = conve._bn1_relu_drop(x.clone())
    assert not torch.equal(ya2 == 0, ya == 0)


@pytest.mark.parametrize('B,F,K,H,bias', [(128, 200, 7, 20, False), (5, 3, 3, 20, True), (37, 45, 5, 12, True), (1, 200, 7, 7, False)])
def test_conve_conv_kernels(B, F, K, H, bias):
    """K8: ConvE's one-input-channel convolution (model.py:166) against torch's conv2d evaluated in float64: output,
    input gradient, weight gradient (and bias gradient); bit-identical across two runs (fixed summation order)."""
    import kgc_gcn_b200 as k
    W = 20
    prm = SimpleNamespace(gcn_out_dim=H * W // 2, hidden_drop=0.0, feat_drop=0.0, k_w=H // 2 if H % 2 == 0 else H, k_h=W,
                          num_filter=F, kernel_size=K, bias=bias)
    g = torch.Generator().manual_seed(B * 7 + F)
    conv = torch.nn.Conv2d(1, F, (K, K), bias=bias).cuda()
    holder = SimpleNamespace(conv_e=conv)
    x = torch.randn(B, 1, H, W, generator=g).cuda()
    dy = torch.randn(B, F, H - K + 1, W - K + 1, generator=g).cuda()
    assert k._lib.lib().kgc_conv1ch_supported(F, K, H, W) == 1
    outs = []
    for _ in range(2):
        xa = x.clone().requires_grad_(True)
        conv.zero_grad()
        ya = k.ConvE._conv(holder, xa)
        assert ya.grad_fn is not None and 'Conv1ch' in type(ya.grad_fn).__name__          # the K8 path, not cuDNN
        ya.backward(dy)
        outs.append((ya.detach().clone(), xa.grad.clone(), conv.weight.grad.clone(),
                     conv.bias.grad.clone() if bias else None))
    for a, b in zip(outs[0], outs[1]):
        assert a is None or torch.equal(a, b)
    x64 = x.double().requires_grad_(True)
    w64 = conv.weight.detach().double().requires_grad_(True)
    b64 = conv.bias.detach().double().requires_grad_(True) if bias else None
    y64 = torch.nn.functional.conv2d(x64, w64, b64)
    y64.backward(dy.double())

    def close(a, b):
        return float((a.double() - b).abs().max()) <= 1e-5 * max(1e-30, float(b.abs().max()))
    assert close(outs[0][0], y64.detach())
    assert close(outs[0][1], x64.grad)
    assert close(outs[0][2], w64.grad)
    if bias:
        assert close(outs[0][3], b64.grad)
    assert k._lib.lib().kgc_conv1ch_supported(F, K, H, 19) == 0 and k._lib.lib().kgc_conv1ch_supported(F, 4, H, W) == 0


@pytest.mark.parametrize('max_norm,wd', [(1.0, 0.0), (1e6, 0.0), (0.5, 0.01), (None, 0.0)])
def test_clip_adam_matches_torch(max_norm, wd):
    """K9: ClipAdam.step() == nn.utils.clip_grad_norm_ + torch.optim.Adam.step() (main.py:68-71, 217) over several steps,
    on tensors whose sizes are not multiples of the kernel's chunk / vector width; same state_dict layout."""
    import kgc_gcn_b200 as k
    g = torch.Generator().manual_seed(5)
    shapes = [(70001, 3), (16384,), (129, 7), (1,), (333, 100)]
    pa = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = k.ClipAdam(pa, lr=3e-3, weight_decay=wd, max_norm=max_norm)
    ob = torch.optim.Adam(pb, lr=3e-3, weight_decay=wd)
    for it in range(5):
        scale = 10.0 ** (it - 2)
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * scale
            a.grad, b.grad = gr.clone(), gr.clone()
        if it == 3:
            for grp in oa.param_groups + ob.param_groups:
                grp['lr'] = 1e-3                                    # a scheduler step (main.py:219)
        if max_norm is not None:
            ref_norm = torch.nn.utils.clip_grad_norm_(pb, max_norm)
        oa.step(); ob.step()
        if max_norm is not None:
            assert abs(float(oa.last_grad_norm) - float(ref_norm)) <= 1e-5 * float(ref_norm)
        for a, b in zip(pa, pb):
            assert float((a - b).detach().abs().max()) <= 2e-6 * max(1.0, float(b.detach().abs().max())), (it, a.shape)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert set(sa['state'].keys()) == set(sb['state'].keys())
    for i in sa['state']:
        assert set(sa['state'][i].keys()) == set(sb['state'][i].keys())
        assert float(sa['state'][i]['step']) == float(sb['state'][i]['step']) == 5.0
        # max-norm relative; torch's clip coefficient comes from an fp32 norm of norms (ours: fp64 partials), the two
        # coefficients agree to ~1e-5 (checked above on the norm) and exp_avg_sq carries that difference squared
        for key, tol in (('exp_avg', 2e-5), ('exp_avg_sq', 4e-5)):
            va, vb = sa['state'][i][key], sb['state'][i][key]
            assert float((va - vb).abs().max()) <= tol * float(vb.abs().max()), (i, key)
    ob.load_state_dict(sa)                                          # a ClipAdam checkpoint loads into torch.optim.Adam ...
    oc = k.ClipAdam(pa, lr=3e-3, weight_decay=wd, max_norm=max_norm)
    oc.load_state_dict(sb)                                          # ... and vice versa, continuing at step 6
    for a in pa:
        a.grad = torch.ones_like(a)
    oc.step()
    assert float(oc.state_dict()['state'][0]['step']) == 6.0
    cpu = torch.nn.Parameter(torch.zeros(3))
    cpu.grad = torch.ones(3)
    with pytest.raises(RuntimeError):                               # no CPU fallback
        k.ClipAdam([cpu], lr=1e-3).step()


def test_checkpoint_round_trip_from_the_device(toy, tmp_path, monkeypatch):
    """utils.save_checkpoint / load_checkpoint (reference utils.py:121-155) with the model and the ClipAdam state in HBM:
    the staged device -> host copies (several chunks, both staging buffers) give the bytes torch.save would, the file
    loads back bit for bit into a fresh model + optimiser on the device, and plain torch.load reads it."""
    k, dl, m, z = toy
    from kgc_gcn_b200 import utils as ku
    monkeypatch.setattr(ku, 'CHUNK_BYTES', 4096)                       # force multi-chunk tensors through the staging ring
    opt = k.ClipAdam(m.parameters(), lr=1e-3, max_norm=1.0)
    for p_ in m.parameters():
        p_.grad = torch.randn_like(p_) * 1e-2
    opt.step()
    for p_ in m.parameters():
        p_.grad = None
    state = {'epoch': 1, 'state_dict': m.state_dict(), 'optim_dict': opt.state_dict(), 'measure': {'mrr': 0.5}}
    d = str(tmp_path / 'ckpt')
    k.save_checkpoint(state, True, d)
    plain = torch.load(os.path.join(d, 'best.ckpt'), weights_only=False)
    for name, t in m.state_dict().items():
        assert plain['state_dict'][name].device.type == 'cpu' and torch.equal(plain['state_dict'][name], t.cpu()), name
    m2 = k.MGCN(dl.num_entity, dl.num_relation, dl.num_edge, params()).cuda()
    opt2 = k.ClipAdam(m2.parameters(), lr=1e-3, max_norm=1.0)
    assert k.load_checkpoint(os.path.join(d, 'last.ckpt'), m2, opt2) == {'mrr': 0.5}
    for (n1, a), (n2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert n1 == n2 and b.is_cuda and torch.equal(a, b), n1
    s1, s2 = opt.state_dict()['state'], opt2.state_dict()['state']
    assert s1.keys() == s2.keys()
    for i in s1:
        assert torch.equal(s1[i]['exp_avg'].cpu(), s2[i]['exp_avg'].cpu()) and torch.equal(s1[i]['exp_avg_sq'].cpu(), s2[i]['exp_avg_sq'].cpu())


def test_tensor_map_encode_on_a_thread_without_a_context():
    """cuTensorMapEncodeTiled is a driver entry point: autograd's backward thread has no current context when the first
    thing it runs is one of the GEMM launchers (CUDA_ERROR_INVALID_CONTEXT, seen in a training step whose backward starts
    with the scorer's gradient GEMM).  Fresh process: forward on the main thread, backward of the tensor-core fc layer
    (two tensor-map encodes) as the backward thread's first CUDA work."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ('import sys, torch\n'
            'sys.path.insert(0, %r)\n'
            'import kgc_gcn_b200 as k\n'
            'x = torch.randn(64, 4096, device="cuda", requires_grad=True)\n'
            'w = torch.randn(200, 4096, device="cuda", requires_grad=True)\n'
            'assert k.linear_tc_supported(x, w)\n'
            'y = k.linear_tc(x, w)\n'
            'y.backward(torch.ones_like(y))\n'
            'torch.cuda.synchronize()\n'
            'ref = torch.ones(64, 200, device="cuda") @ w.detach()\n'
            'assert float((x.grad - ref).abs().max()) < 1e-3 * float(ref.abs().max())\n'
            'print("ENCODE_OK")\n') % root
    p = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and 'ENCODE_OK' in p.stdout, p.stdout[-3000:]
