"""The CPU restatement (oracle/restated.py) against the fixtures produced by the
reference's OWN functions (oracle/make_golden.py -> tests/golden/*.pt).
CPU only; this is what pins the oracle (the reference itself has
no tests for this path)."""
import os

import pytest
import torch

from conftest import rel_fro
from oracle import restated as O
from oracle import synth


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), map_location="cpu", weights_only=False)


def test_covariance_hooks_match_reference(golden_dir):
    g = _load(golden_dir, "cov_toy.pt")
    net = synth.ToyDetector()
    net.load_state_dict(g["state_dict"])
    batches = synth.toy_batches(seed=g["seed"])
    assert abs(float(sum(b.double().sum() for b in batches)) - g["batch_checksum"]) < 1e-6
    fea = O.covariances_of_model(net, batches)
    assert set(fea) == set(g["fea_in"])
    for k, ref in g["fea_in"].items():
        assert fea[k].shape == ref.shape
        assert rel_fro(fea[k], ref) < 2e-6, k


def test_conv_rows_column_order_is_cin_kh_kw():
    x = torch.arange(2 * 3 * 5 * 6, dtype=torch.float32).reshape(2, 3, 5, 6)
    rows = O.conv_rows(x, (3, 3), (1, 1), (1, 1))
    assert rows.shape == (30, 27)
    m = x.mean(0)
    # output pixel (oy=2, ox=3), channel 1, tap (0,2) -> input (1, 4)
    assert rows[2 * 6 + 3, 1 * 9 + 0 * 3 + 2] == m[1, 1, 4]
    # zero padding at the border
    assert rows[0, 0] == 0


@pytest.mark.parametrize("offset", [0.0, 0.5, -0.5, 3.0])
def test_threshold_and_transform_match_reference(golden_dir, offset):
    g = _load(golden_dir, "projector.pt")
    for name, ent in g["proj"][offset].items():
        d, seed, rank = g["seeds"][name]
        cov = synth.decaying_cov(d, seed, rank=rank)
        s, v = O.eigens(cov)
        assert rel_fro(s, ent["svals"]) < 1e-5
        # the reference's own spectrum -> identical cut index
        assert O.threshold_index(ent["svals"].numpy(), offset) == ent["i_thres"]
        assert O.threshold_index(s.numpy(), offset) == ent["i_thres"]
        p = O.transform(s, v, name, offset)
        assert rel_fro(p, ent["transform"]) < 1e-4, name
        if "backbone" in name:
            assert abs(float(torch.norm(p)) - 1.0) < 1e-5


@pytest.mark.parametrize("tag", ["mom_wd", "plain", "nesterov_damp"])
def test_sgd_nscl_steps_match_reference(golden_dir, tag):
    g = _load(golden_dir, "sgd_steps.pt")
    proj = _load(golden_dir, "projector.pt")["proj"][0.0]
    transforms = {n: e["transform"] for n, e in proj.items()}
    run = g["steps"][tag]
    params = {n: v.clone() for n, v in g["init"].items()}
    states = {}
    for step, want in zip(g["grads"], run["traj"]):
        grads = {n: v.clone() for n, v in step.items()}
        O.sgd_nscl_step(params, grads, states, transforms, svd=True, **run["kw"])
        for n in params:
            assert rel_fro(params[n], want["w"][n]) < 1e-6, (tag, n)
            assert rel_fro(states[n].previous_grad, want["buf"][n]) < 1e-6 or \
                float(want["buf"][n].abs().max()) == 0.0
            # in-place weight-decay side effect on .grad (SGD_NSCL.py:399-400)
            assert rel_fro(grads[n], want["grad_after"][n]) < 1e-6, (tag, n)


def test_prototypes_match_reference(golden_dir):
    g = _load(golden_dir, "prototypes.pt")
    feats, lab = synth.proto_features()
    assert abs(float(feats.double().sum()) - g["feat_checksum"]) < 1e-3
    assert torch.equal(lab, g["labels"])
    protos, tmp_label, masks = O.build_prototypes(feats, lab, range(0, 3), max_proto=10)
    assert torch.equal(tmp_label, g["tmp_label"])
    assert [len(m) for m in masks] == [len(m) for m in g["masks"]]
    for mine, ref in zip(masks, g["masks"]):
        for a, b in zip(mine, ref):
            assert torch.equal(a, b)
    assert rel_fro(protos, g["protos"]) < 1e-6

    # next task: saved masks are replayed for old classes, class 3 is new
    feats2, lab2 = synth.proto_features(seed=1, classes=4, per_class=60, bg=20)
    feats2 = torch.cat([feats, feats2[lab2 == 3]])
    lab2 = torch.cat([lab, lab2[lab2 == 3]])
    assert torch.equal(lab2, g["labels2"])
    protos2, tmp_label2, masks2 = O.build_prototypes(
        feats2, lab2, range(0, 4), max_proto=10, saved_masks=[list(m) for m in masks])
    assert torch.equal(tmp_label2, g["tmp_label2"])
    assert [len(m) for m in masks2] == [len(m) for m in g["masks2"]]
    assert rel_fro(protos2, g["protos2"]) < 1e-6


def test_autocorrelation_identity():
    """The identity behind the B200 covariance kernel for 3x3 s1 p1 convs (csrc/geometry.h):
    the unfold Gram of the reference (nsrunner_roi_replay.py:908-930) equals spatial
    autocorrelation blocks R_(dy,dx) minus edge-row / edge-column sums plus corner terms."""
    import numpy as np
    rng = np.random.default_rng(0)
    C, H, W = 4, 5, 6
    m = rng.standard_normal((C, H, W))
    ref = O.cov_conv2d(torch.from_numpy(m)[None], (3, 3), (1, 1), (1, 1)).numpy()

    def px(u, v):
        return m[:, u, v] if 0 <= u < H and 0 <= v < W else np.zeros(C)

    def corr(points, dy, dx):
        return sum(np.outer(m[:, u, v], px(u + dy, v + dx)) for u, v in points)

    allp = [(u, v) for u in range(H) for v in range(W)]
    G = np.zeros((9 * C, 9 * C))
    for t in range(9):
        for t2 in range(t, 9):
            i, j, i2, j2 = t // 3, t % 3, t2 // 3, t2 % 3
            dy, dx = i2 - i, j2 - j
            g = corr(allp, dy, dx)
            if i == i2 and i != 1:
                g -= corr([(H - 1 if i == 0 else 0, v) for v in range(W)], dy, dx)
            if j == j2 and j != 1:
                g -= corr([(u, W - 1 if j == 0 else 0) for u in range(H)], dy, dx)
            if t == t2 and i != 1 and j != 1:
                g += corr([(H - 1 if i == 0 else 0, W - 1 if j == 0 else 0)], 0, 0)
            for c in range(C):
                for c2 in range(C):
                    G[c * 9 + t, c2 * 9 + t2] = g[c, c2]
                    G[c2 * 9 + t2, c * 9 + t] = g[c, c2]
    assert np.abs(G - ref).max() < 1e-10


def test_replay_loss_double_softmax():
    torch.manual_seed(0)
    score = torch.randn(7, 21)
    label = torch.randint(0, 10, (7,))
    want = torch.nn.functional.cross_entropy(
        torch.cat([score[:, :19], score[:, -1:]], -1).softmax(-1), label)
    assert torch.allclose(O.replay_loss(score, label, 19), want)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution
    (kat_vectors: zero counter/key and the pi-digits vector)."""
    import numpy as np
    out = O.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros(2, np.uint32))[0]
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    ctr = np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    key = np.array([0xa4093822, 0x299f31d0], np.uint32)
    out = O.philox4x32_10(ctr, key)[0]
    assert [hex(int(v)) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


# ------------------------------------------------------------------ SURVEY 8(f)-3 / 8(f)-4
def test_pseudo_label_merge_oracle_matches_reference_fixture(golden_dir):
    """oracle.restated.pseudo_label_merge against the outputs of the reference's own
    FasterRCNNRoIReplay.loss (oracle/make_golden.py section 5): bit-exact."""
    cases = torch.load(os.path.join(golden_dir, "pseudo_merge.pt"), weights_only=False)
    for case in cases:
        gt_b, gt_l, ps_b, ps_s, ps_l = synth.pseudo_label_case(case["seed"])
        got = O.pseudo_label_merge(gt_b, gt_l, ps_b, ps_s, ps_l)
        assert len(got) == len(case["out"])
        for mine, ref in zip(got, case["out"]):
            for a, b in zip(mine, ref):
                assert torch.equal(a, b)
        # the scenario is not degenerate: something was dropped for IoU, for score, and the
        # RPN / RoI target sets differ
        assert any(m[0].shape[0] != m[2].shape[0] for m in got)
        assert any(m[0].shape[0] < g.shape[0] + p.shape[0] for m, g, p in zip(got, gt_b, ps_b))


def test_ewc_oracle_matches_reference_fixture(golden_dir):
    """Importance accumulation (:978-981) and penalty (:1056-1069) restated, against the
    reference's calculate_save_importance / EWCHook run over two tasks."""
    g = torch.load(os.path.join(golden_dir, "ewc.pt"), weights_only=False)
    torch.manual_seed(0)
    net = synth.ToyBNNet()
    net.load_state_dict(g["state0"])
    reg = O.ewc_register_params(net)
    assert list(reg.keys()) == g["reg_names"]
    # task 1 importance, restated loop
    net.eval()
    imp = {n: torch.zeros_like(p) for n, p in reg.items()}
    batches = synth.ewc_batches(0)
    for b in batches:
        for p in net.parameters():
            p.grad = None
        torch.nn.functional.cross_entropy(net(b["inputs"]), b["data_samples"]).backward()
        O.ewc_accumulate(imp, {n: p.grad for n, p in reg.items()}, len(b), len(batches))
    for n in reg:
        assert torch.equal(imp[n].unsqueeze(0), g["terms"]["importance"][n][0]), n
    # penalty value and gradient at the fixture's parameters
    net.load_state_dict(g["state3"])
    reg = O.ewc_register_params(net)
    net.train()
    for p in net.parameters():
        p.grad = None
    loss = O.ewc_penalty(reg, g["terms"])
    loss.backward()
    assert abs(float(loss) - float(g["ewc_loss"])) <= 1e-5 * abs(float(g["ewc_loss"]))
    for n, ref in g["grads"].items():
        assert rel_fro(reg[n].grad, ref) < 1e-5, n
    assert reg["bn2.bias"].grad is None and "bn2.bias" not in g["grads"]


def test_roi_extract_oracle_matches_reference_fixture(golden_dir):
    """oracle.restated.roi_extract against the reference's own SingleRoIExtractor.forward /
    map_roi_levels (pooling by torchvision in both: mmcv is absent)."""
    g = torch.load(os.path.join(golden_dir, "roi_extract.pt"), weights_only=False)
    feats, rois, labels = synth.roi_case(g["seed"])
    out, lv = O.roi_extract(feats, rois)
    assert torch.equal(lv, g["levels"])
    assert torch.equal(out, g["out"])
    assert set(lv.tolist()) == {0, 1, 2, 3}
