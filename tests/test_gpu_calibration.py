"""ECE / reliability binning, CE loss, conf/pred and optimizer kernels against
the oracle and the reference's golden vectors.  Bin counts are bit-exact."""
import importlib

import numpy as np
import pytest
import torch

from conftest import Golden, dropin_src, load_pkg
from oracle import ece_oracle, fusion_oracle

pytestmark = pytest.mark.gpu


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _unpack(stats):
    h = stats.cpu().numpy()
    return h[0], h[1], h[2].astype(np.uint64).astype(np.float64) / 2.0 ** 32


@pytest.mark.parametrize("nb", [15, 10, 2])
def test_bins_bit_exact_vs_reference_golden(nb):
    ops = _ops()
    g = Golden("ece_seeded.npz")
    conf = g.t("conf").cuda()
    pred, label = g.t("pred").long().cuda(), g.t("label").long().cuda()
    cnt, cor, cs = _unpack(ops.ece_bin(conf, pred, label, g[f"edges_f32/{nb}"].astype(np.float64).tolist()))
    assert np.array_equal(cnt, g[f"count_f32/{nb}"])
    assert np.array_equal(cor, g[f"correct_f32/{nb}"])
    assert np.allclose(cs, g[f"confsum_f32/{nb}"], rtol=1e-6)
    cnt64, _, _ = _unpack(ops.ece_bin(conf, pred, label, np.linspace(0.0, 1.0, nb + 1).tolist()))
    assert np.array_equal(cnt64, g[f"count_f64/{nb}"])
    # unaligned views exercise the scalar path
    c2, k2, s2 = _unpack(ops.ece_bin(conf[1:], pred[1:], label[1:], g[f"edges_f32/{nb}"].astype(np.float64).tolist()))
    r = ece_oracle.bin_masks(g["conf"][1:], g["pred"][1:], g["label"][1:], g[f"edges_f32/{nb}"])
    assert np.array_equal(c2, r[0]) and np.array_equal(k2, r[1])


@pytest.mark.parametrize("nb", [15, 10, 2])
def test_dropin_metrics_match_reference_floats(nb):
    import sys
    sys.path.insert(0, dropin_src())
    import uncertainty
    g = Golden("ece_seeded.npz")
    conf, pred, label = g.t("conf"), g.t("pred").long(), g.t("label").long()
    CM = uncertainty.CalibrationMetrics
    for tensors in ((conf, pred, label), (conf.cuda(), pred.cuda(), label.cuda())):
        ece = CM.expected_calibration_error(*tensors, num_bins=nb)
        mce = CM.maximum_calibration_error(*tensors, num_bins=nb)
        assert isinstance(ece, float) and abs(ece - float(g[f"ece/{nb}"])) <= 1e-6
        assert abs(mce - float(g[f"mce/{nb}"])) <= 1e-6
    counts, avg, acc, _ = CM.reliability_bins(conf.numpy(), pred.numpy(), label.numpy(), nb)
    rc, ra, rk = ece_oracle.reliability_bins(conf.numpy(), pred.numpy(), label.numpy(), nb)
    assert np.array_equal(counts, rc) and np.allclose(avg, ra, atol=1e-6) and np.allclose(acc, rk, atol=1e-7)


def test_survey_kats():
    import sys
    sys.path.insert(0, dropin_src())
    import uncertainty
    CM = uncertainty.CalibrationMetrics
    assert CM.expected_calibration_error(torch.tensor([0.8, 0.7]), torch.tensor([0, 1]), torch.tensor([0, 1]), 2) == 0.25
    assert CM.maximum_calibration_error(torch.tensor([0.8, 0.7]), torch.tensor([0, 1]), torch.tensor([0, 1]), 2) == 0.25
    g = Golden("ece_kat.npz")
    gen = torch.Generator().manual_seed(1234)
    logits = torch.randn(100000, 25, generator=gen) * 2
    labels = torch.randint(0, 25, (100000,), generator=gen)
    conf, pred = _ops().softmax_conf_pred(logits.cuda())
    ref_conf, ref_pred = fusion_oracle.softmax_conf_pred(logits)
    assert torch.equal(pred.cpu(), ref_pred)  # argmax bit-exact on identical logits (first-max rule)
    assert float((conf.cpu() - ref_conf).abs().max()) <= 1e-6
    ece = CM.expected_calibration_error(ref_conf, ref_pred, labels, 15)
    assert abs(ece - float(g["ece15_seed1234"])) <= 1e-6
    nll = CM.negative_log_likelihood(logits, labels)
    assert abs(nll - float(torch.nn.functional.cross_entropy(logits, labels))) <= 1e-5


def test_full_size_properties_and_shard_merge():
    """N = 2^24 samples (320 MB of inputs, > L2): totals, shard-merge exactness, C oracle."""
    ops = _ops()
    n = 1 << 24
    gen = torch.Generator(device="cuda").manual_seed(3)
    conf = torch.rand(n, device="cuda", generator=gen)
    conf[::1000] = float("nan")
    conf[1::1000] = 1.0
    conf[2::1000] = 0.0
    conf[3::1000] = 1.5
    pred = torch.randint(0, 25, (n,), device="cuda", generator=gen)
    label = torch.randint(0, 25, (n,), device="cuda", generator=gen)
    edges = ece_oracle.linspace_f32(15).astype(np.float64).tolist()
    whole = ops.ece_bin(conf, pred, label, edges)
    in_range = int(((conf >= 0) & (conf <= 1)).sum())
    assert int(whole[0].sum()) == in_range  # NaN / out-of-range land in no bin
    cut = 5_000_003
    parts = ops.ece_bin(conf[:cut], pred[:cut], label[:cut], edges)
    parts = ops.ece_bin(conf[cut:], pred[cut:], label[cut:], edges, out=parts)
    assert torch.equal(parts, whole)  # integer accumulation: shards merge bit-exactly (incl. Q32 sums)
    assert torch.equal(ops.ece_bin(conf, pred, label, edges), whole)  # run-to-run deterministic
    m = 1 << 22
    cnt, cor, cs = ece_oracle.bin_masks_c(conf[:m].cpu().numpy(), pred[:m].cpu().numpy(),
                                          label[:m].cpu().numpy(), np.array(edges), threads=8)
    got = ops.ece_bin(conf[:m], pred[:m], label[:m], edges).cpu().numpy()
    assert np.array_equal(got[0], cnt) and np.array_equal(got[1], cor)
    assert np.allclose(got[2].astype(np.uint64).astype(np.float64) / 2.0 ** 32, cs, rtol=1e-7)


def test_empty_and_tiny_inputs():
    ops = _ops()
    edges = ece_oracle.linspace_f32(15).astype(np.float64).tolist()
    e = torch.empty(0, device="cuda")
    z = torch.empty(0, dtype=torch.long, device="cuda")
    assert int(ops.ece_bin(e, z, z, edges).abs().sum()) == 0
    one = ops.ece_bin(torch.tensor([1.0], device="cuda"), torch.tensor([3], device="cuda"),
                      torch.tensor([3], device="cuda"), edges).cpu()
    assert one[0].tolist() == [0] * 14 + [1] and one[1].tolist() == [0] * 14 + [1]
    assert int(one[2][14]) == 2 ** 32


def test_cross_entropy_and_optimizer_match_oracle():
    ops = _ops()
    g = Golden("fusion_pamap_small.npz")
    logits = g.t("train/logits").cuda()
    labels = g.t("labels").cuda()
    loss, grad = ops.cross_entropy(logits, labels, 0.05)
    ref = g.t("train/logits").clone().requires_grad_(True)
    ref_loss = fusion_oracle.cross_entropy_label_smoothing(ref, g.t("labels"), 0.05)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-6
    assert float((grad.cpu() - ref.grad).abs().max()) <= 1e-7
    # AdamW + clip on the flat arena vs torch.optim.AdamW / clip_grad_norm_ (golden "opt/*")
    keys = list(g.group("sd").keys())
    flat = lambda grp: torch.cat([g.t(f"{grp}/{k}").flatten() for k in keys]).cuda()
    p, gr = flat("sd"), flat("grad")
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sq = ops.grad_sq_norm(gr)
    assert abs(float(sq.sqrt()) - float(g["opt/grad_norm"])) <= 1e-5
    ops.adamw_step(p, gr, m, v, step=1, lr=1e-3, weight_decay=1e-4, max_norm=1.0, sq_norm=sq)
    assert float((p.cpu() - flat("opt").cpu()).abs().max()) <= 1e-6


def test_layout_aware_optimizer_matches_torch_adamw():
    """msf_fusion_optimizer_step (dead q/k slots = weight decay only, live slots = AdamW, global-norm clip)
    against the parameters torch.optim.AdamW + clip_grad_norm_ produced in the reference run (golden "opt/*")."""
    from helpers import module_from_golden
    ops = _ops()
    g = Golden("fusion_pamap_small.npz")
    plan = module_from_golden(g)._plan()
    keys = [k for k, _, _ in plan.slots]
    assert keys == list(g.group("sd").keys())
    flat = lambda grp: torch.cat([g.t(f"{grp}/{k}").flatten() for k in keys]).cuda()
    p, gr = flat("sd"), flat("grad")
    p2 = p.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    state = torch.tensor([0, 0, 1], dtype=torch.int64, device="cuda")
    sq = ops.fusion_optimizer_step(plan, p, gr, m, v, state, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
    assert abs(float(sq.sqrt()) - float(g["opt/grad_norm"])) <= 1e-5
    assert float((p.cpu() - flat("opt").cpu()).abs().max()) <= 1e-6
    # identical to the generic flat AdamW, and the dead slots never grow moments
    m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
    ops.adamw_step(p2, gr, m2, v2, step=1, lr=1e-3, weight_decay=1e-4, max_norm=1.0, sq_norm=ops.grad_sq_norm(gr))
    assert float((p - p2).abs().max()) <= 1e-7 and float((m - m2).abs().max()) <= 1e-9
    for key, off, shape in plan.slots:
        if ".query_proj." in key or ".key_proj." in key:
            n = int(np.prod(shape))
            assert float(m[off:off + n].abs().max()) == 0.0 and float(v[off:off + n].abs().max()) == 0.0
            assert torch.equal(p[off:off + n], (flat("sd")[off:off + n] * (1.0 - 1e-3 * 1e-4)))


def test_packed_optimizer_step_equals_step_plus_pack():
    """msf_fusion_optimizer_step_packed = msf_fusion_optimizer_step + msf_fusion_pack_bf16 +
    msf_train_state_advance in one launch, over three steps (config-2 shape).  The packed arena is exactly the
    pack of the updated parameters; the parameters agree to rounding (the gradient norm is summed in a
    different order, which can move the clip factor by an ulp)."""
    from helpers import PAMAP2, seeded_case
    ops = _ops()
    model, *_ = seeded_case(PAMAP2, 256, 4, 25, 8, seed=3, device="cuda")
    plan = model._plan()
    own = dict(model.named_parameters())
    pa = plan.gather([own[k].detach() for k, _, _ in plan.slots])
    gen = torch.Generator(device="cuda").manual_seed(5)
    gr = torch.randn(pa.shape, generator=gen, device="cuda") * 1e-2
    for key, off, shape in plan.slots:   # dead query/key slots carry exact-zero gradients
        if ".query_proj." in key or ".key_proj." in key:
            gr[off:off + int(np.prod(shape))] = 0.0
    pb, pc = pa.clone(), pa.clone()
    ma, va, mb, vb, mc, vc = (torch.zeros_like(pa) for _ in range(6))
    sa = torch.tensor([7, 3, 1], dtype=torch.int64, device="cuda")
    sb, sc = sa.clone(), sa.clone()
    a16 = plan.pack_bf16(pa)
    c16 = a16.clone()
    for step in range(3):
        gstep = gr * (1.0 + 0.25 * step)
        ops.fusion_optimizer_step(plan, pa, gstep, ma, va, sa, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
        sa[1:] += 1
        sq = ops.fusion_optimizer_step_packed(plan, pb, gstep, mb, vb, sb, a16, lr=1e-3, weight_decay=1e-4,
                                              max_norm=1.0)
        assert abs(float(sq) - float((gstep.double() ** 2).sum())) <= 1e-9 * float(sq)
        assert float((pa - pb).abs().max()) <= 1e-7 and float((ma - mb).abs().max()) <= 1e-9
        assert float((va - vb).abs().max()) <= 1e-12
        assert torch.equal(plan.pack_bf16(pb).view(torch.int16), a16.view(torch.int16))
        assert torch.equal(sa, sb)
        # MSF_OPT_NORM_GIVEN: the square norm comes in through sq_norm[1] (in the train step the pass that wrote
        # the gradients puts it there) and the launch has no norm phase
        sq2 = torch.zeros(2, dtype=torch.float64, device="cuda")
        sq2[1] = (gstep.double() ** 2).sum()
        ops.fusion_optimizer_step_packed(plan, pc, gstep, mc, vc, sc, c16, lr=1e-3, weight_decay=1e-4, max_norm=1.0,
                                         sq_norm=sq2, norm_given=True)
        assert abs(float(sq2[0]) - float(sq)) <= 1e-9 * float(sq)
        assert float((pb - pc).abs().max()) <= 1e-7 and float((mb - mc).abs().max()) <= 1e-9
        assert float((vb - vc).abs().max()) <= 1e-12
        assert torch.equal(plan.pack_bf16(pc).view(torch.int16), c16.view(torch.int16))
        assert torch.equal(sb, sc)


@pytest.mark.parametrize("n,classes", [(1000, 25), (65536, 25), (777, 40), (64, 3)])
def test_eval_epilogue_matches_host_metrics(n, classes):
    """msf_eval_accumulate (ops.EvalStats): accuracy, macro-F1 (sklearn's definition, zero_division=0), mean NLL and
    ECE / MCE from ONE pass over the logits, against the reference's host recipe (src/eval.py:89-112) on the same
    logits — integer accumulators: counts exact, NLL to fp32 rounding; two half-batches merge to the same integers."""
    import importlib
    from conftest import load_pkg
    ops = importlib.import_module(load_pkg().__name__ + ".ops")
    g = torch.Generator().manual_seed(n + classes)
    logits = torch.randn(n, classes, generator=g) * 3.0
    labels = torch.randint(0, classes, (n,), generator=g)
    if classes > 10:
        labels[labels == 7] = 8          # a class that never occurs in the labels
        logits[:, 5] -= 100.0            # ... and one that is never predicted
    st = ops.EvalStats(classes)
    conf = torch.empty(n, device="cuda")
    pred = torch.empty(n, dtype=torch.int64, device="cuda")
    st.update(logits.cuda(), labels.cuda(), conf=conf, pred=pred)
    m = st.metrics()
    probs = torch.softmax(logits.double(), 1)
    rconf, rpred = probs.max(1)
    assert torch.equal(pred.cpu(), rpred)
    assert float((conf.cpu().double() - rconf).abs().max()) <= 1e-6
    assert m["num_samples"] == n and m["out_of_range_labels"] == 0
    assert abs(m["accuracy"] - float((rpred == labels).double().mean())) < 1e-12
    nll = float(torch.nn.functional.cross_entropy(logits.double(), labels))
    assert abs(m["loss"] - nll) <= 2e-6 * max(1.0, nll)
    f1s = []
    for c in range(classes):     # sklearn f1_score(average="macro", zero_division=0) over labels seen in either vector
        tp = int(((rpred == c) & (labels == c)).sum()); fp = int(((rpred == c) & (labels != c)).sum())
        fn = int(((rpred != c) & (labels == c)).sum())
        if tp + fp + fn == 0:
            continue
        f1s.append(2 * tp / (2 * tp + fp + fn))
    assert abs(m["f1_macro"] - sum(f1s) / len(f1s)) < 1e-12
    two = ops.EvalStats(classes)
    half = n // 2
    two.update(logits[:half].cuda(), labels[:half].cuda())
    two.update(logits[half:].cuda(), labels[half:].cuda())
    assert torch.equal(two.confusion, st.confusion) and torch.equal(two.bins, st.bins)
    assert torch.equal(two.scalars, st.scalars)
    # the bins are those of the stand-alone binning kernel on the same (conf, pred, label)
    edges = torch.linspace(0, 1, 16).double().tolist()
    bins = ops.ece_bin(conf, pred, labels.cuda(), edges)
    assert torch.equal(bins, st.bins)


def test_eval_epilogue_counts_out_of_range_labels():
    import importlib
    from conftest import load_pkg
    ops = importlib.import_module(load_pkg().__name__ + ".ops")
    logits = torch.randn(32, 5, device="cuda")
    labels = torch.randint(0, 5, (32,), device="cuda")
    labels[4], labels[9] = -100, 5
    st = ops.EvalStats(5)
    st.update(logits, labels)
    m = st.metrics()
    assert m["num_samples"] == 32 and m["out_of_range_labels"] == 2 and int(st.confusion.sum()) == 30
