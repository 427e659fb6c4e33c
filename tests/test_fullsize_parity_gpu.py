"""GPU parity at the BENCHMARK shape (BASELINE.json configs[1]: 400x600 reflect-padded to 420x620) against the reference's own outputs
(tests/golden/pipe_420x620.npz, written by oracle/gen_golden_fullsize.py from the unmodified reference modules; the CPU oracle reproduces
them exactly, tests/golden/PIN_REPORT.txt last line).

north_star bars: codebook indices bit-exact (an operator-level property: asserted teacher-forced on the reference's z), pixels within
1e-3 abs in fp32 (asserted teacher-forced: decoders fed the reference's z, so the VQ discontinuity is not in the way), PSNR within
0.01 dB (asserted end to end).  End to end the index map differs where the flow output z (accurate to ~1e-3 after 28 steps) lies
within rounding of a Voronoi boundary of the 8192-entry codebook; the flip fraction is asserted (<= 5e-4) and reported."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
H, W = 400, 600


@pytest.fixture(scope="module")
def case():
    from glare_b200 import synth
    g = load_golden("pipe_420x620")
    lq, gt = synth.synth_images(1, H, W, seed=0)
    lr = synth.preprocess(synth.pad_lol(lq))
    assert float(lr.double().sum()) == pytest.approx(float(g["lr_checksum"]), rel=1e-12)      # the input the golden was made from
    return {"lr": lr, "gt": gt, "z": torch.from_numpy(g["z_flow"]), "idx": g["idx"].astype(np.int64),
            "out": torch.from_numpy(g["out"]), "color_map": torch.from_numpy(g["color_map"])}


def _engine(sd_g, sd_v, name):
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    return GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense(name))


def _crop(x):
    return x[:, :, :H, 20:].clamp(0, 1)          # infer_dataset_lol.py:135


def test_fingerprints(sd_g, sd_v):
    from glare_b200 import synth
    g = load_golden("pipe_420x620")
    assert synth.state_fingerprint(sd_g) == pytest.approx(float(g["fingerprint_netG"]), rel=1e-12)
    assert synth.state_fingerprint(sd_v) == pytest.approx(float(g["fingerprint_vqgan"]), rel=1e-12)


def test_teacher_forced_fp32_bars(glare_lib, sd_g, sd_v, case):
    """reference z in: indices bit-exact (16 275 tokens), decoder pixels within 1e-3 abs of the reference output"""
    eng = _engine(sd_g, sd_v, "auto")
    with torch.no_grad():
        enc = eng.cond_encoder(case["lr"].cuda())
        assert float((enc["color_map"].cpu() - case["color_map"]).abs().max()) < 1e-3
        zq, idx = eng.vector_quantize(case["z"].cuda())
        assert np.array_equal(idx.cpu().numpy(), case["idx"])
        out = eng.aft_decoder(case["z"].cuda(), eng.vq_decoder_features(zq), enc["mid_feat"]).float().cpu()
        assert eng.dense.attention_verified()
    dmax = float((out - case["out"]).abs().max())
    print("teacher-forced 420x620: pixel max diff %.3g (unclamped, uncropped)" % dmax)
    assert dmax < 1e-3


@pytest.mark.parametrize("backend,min_agree,max_dpsnr", [("auto", 0.9995, 0.01), ("tc-bf16", 0.90, 0.05)])
def test_end_to_end_against_reference(glare_lib, sd_g, sd_v, case, backend, min_agree, max_dpsnr):
    """fp32 configuration: index agreement >= 0.9995 and |dPSNR| <= 0.01 dB against the REFERENCE output.  bf16 operands (BASELINE
    config 3): its own stated bound -- >= 90 % of the indices, |dPSNR| <= 0.05 dB (measured 94.9 % / 0.006 dB in round 1)."""
    from oracle import glare_oracle as O
    eng = _engine(sd_g, sd_v, backend)
    st = {}
    out = eng.infer(case["lr"], stages=st).float().cpu()
    agree = float((st["idx"].cpu().numpy() == case["idx"]).mean())
    dz = float((st["z_flow"].cpu() - case["z"]).abs().max())
    d = (_crop(out) - _crop(case["out"])).abs()
    dpsnr = abs(O.psnr(_crop(out), case["gt"]) - O.psnr(_crop(case["out"]), case["gt"]))
    print("%s end to end 420x620: z max diff %.3g, idx agree %.5f (%d of %d flipped), pixel max %.3g mean %.3g, |dPSNR| %.5f dB"
          % (backend, dz, agree, int(round((1 - agree) * case["idx"].size)), case["idx"].size, float(d.max()), float(d.mean()), dpsnr))
    assert agree >= min_agree and dpsnr <= max_dpsnr
    if backend == "auto":
        assert dz < 5e-3 and float(d.mean()) < 2e-4


def test_1080p_against_the_oracle(glare_lib, sd_g, sd_v):
    """BASELINE configs[4] shape: 1920x1080 -> auto_padding 1088x1936, 131 648 latent tokens, attention in bands of query rows.  Golden =
    the CPU oracle with blockwise attention (oracle/gen_golden_1080p.py; the reference itself needs 69 GB per attention matrix): indices
    bit-exact and decoder pixels within 1e-3 teacher-forced, index agreement and PSNR end to end."""
    import os
    from conftest import GOLD
    from glare_b200 import synth
    from oracle import glare_oracle as O
    if not os.path.exists(os.path.join(GOLD, "pipe_1080p.npz")):
        pytest.skip("tests/golden/pipe_1080p.npz not generated")
    g = load_golden("pipe_1080p")
    lq, gt = synth.synth_images(1, 1080, 1920, seed=200)
    xp, (h1, h2, w1, w2) = synth.auto_padding(lq)
    lr = synth.preprocess(xp)
    assert float(lr.double().sum()) == pytest.approx(float(g["lr_checksum"]), rel=1e-12)
    z_ref, idx_ref = torch.from_numpy(g["z_flow"]), g["idx"].astype(np.int64)
    out_ref_s2 = torch.from_numpy(g["out_s2"].astype(np.float32))
    eng = _engine(sd_g, sd_v, "auto")
    with torch.no_grad():
        enc = eng.cond_encoder(lr.cuda())
        zq, idx = eng.vector_quantize(z_ref.cuda())
        assert np.array_equal(idx.cpu().numpy(), idx_ref)                                   # 131 648 / 131 648 on the oracle's z
        out_tf = eng.aft_decoder(z_ref.cuda(), eng.vq_decoder_features(zq), enc["mid_feat"]).float().cpu()
        assert eng.dense.attention_verified()
    d_tf = float((out_tf[:, :, ::2, ::2] - out_ref_s2).abs().max())
    st = {}
    out = eng.infer(lr, stages=st).float().cpu()
    agree = float((st["idx"].cpu().numpy() == idx_ref).mean())
    dz = float((st["z_flow"].cpu() - z_ref).abs().max())
    crop = out[:, :, h1:out.shape[2] - h2, w1:out.shape[3] - w2].clamp(0, 1)
    dpsnr = abs(O.psnr(crop, gt) - float(g["psnr"]))
    print("1080p: teacher-forced pixel max diff %.3g (stride-2 samples, fp16-stored golden), end to end z max diff %.3g, idx agree %.5f "
          "(%d of %d flipped), |dPSNR| %.5f dB" % (d_tf, dz, agree, int(round((1 - agree) * idx_ref.size)), idx_ref.size, dpsnr))
    assert d_tf < 1e-3 + 1e-3                      # 1e-3 bar + the fp16 storage of the golden (|out| <= 2: half an ulp = 5e-4)
    assert agree >= 0.9995 and dpsnr <= 0.01
