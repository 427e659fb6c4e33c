"""GPU parity: the whole inference path (GlareEngine) vs the golden pipeline vectors (reference netG outputs)
and the oracle, stage by stage and end to end.  fp32 path; pixel bar 1e-3 abs, PSNR delta <= 0.01 dB
(north_star).  Index agreement is reported end to end and required bit-exact teacher-forced (test_vq_gpu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tc-3xtf32", "tc-tf32bf16x2", "tc-bf16x3", "torch-fp32"])
def engine(request, glare_lib, sd_g, sd_v):
    """fp32-grade configurations: the tensor-core dense path in 3xTF32 mode, and the cuDNN fp32 library baseline"""
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    if request.param == "torch-fp32":                   # comparison path of the tests, not part of the package
        from libdense import TorchDense
        return GlareEngine(sd_g, sd_v, device="cuda:0", dense=TorchDense())
    return GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense(request.param))


@pytest.mark.parametrize("name", ["pipe_32x48", "pipe_64x96"])
def test_stages_against_golden(engine, name):
    from oracle import glare_oracle as O
    g = load_golden(name)
    st = {}
    out = engine.infer(torch.from_numpy(g["lr"]), stages=st).cpu()
    c = lambda k: st[k].float().cpu()
    assert torch.allclose(c("cond_feat"), torch.from_numpy(g["cond_feat"]), atol=1e-4)
    assert torch.allclose(c("color_map"), torch.from_numpy(g["color_map"]), atol=1e-4)
    assert torch.allclose(c("mid0")[:, ::16], torch.from_numpy(g["mid0"]), atol=1e-3)
    zf = torch.from_numpy(g["z_flow"])
    assert torch.allclose(c("z_flow"), zf, atol=2e-3, rtol=1e-4), float((c("z_flow") - zf).abs().max())
    agree = float((st["idx"].cpu().numpy() == g["idx"].astype(np.int64).reshape(-1)).mean())
    assert agree >= 0.995, agree
    gt = torch.from_numpy(g["gt"])
    ref = torch.from_numpy(g["out"])
    dpsnr = abs(O.psnr(out.clamp(0, 1), gt) - O.psnr(ref.clamp(0, 1), gt))
    assert dpsnr <= 0.01, dpsnr
    if agree == 1.0:
        assert float((out - ref).abs().max()) < 1e-3


def test_teacher_forced_decoders(engine):
    """decoders fed the golden z / z_q: pixel bar 1e-3 abs without the index discontinuity in the way"""
    g = load_golden("pipe_64x96")
    lr = torch.from_numpy(g["lr"]).cuda()
    enc = engine.cond_encoder(lr)
    z = torch.from_numpy(g["z_flow"]).cuda()
    zq, idx = engine.vector_quantize(z)
    assert np.array_equal(idx.cpu().numpy(), g["idx"].astype(np.int64).reshape(-1))
    feats = engine.vq_decoder_features(zq)
    assert torch.allclose(feats[0].float().cpu()[:, ::16], torch.from_numpy(g["vq_feat1"]), atol=1e-3)
    assert torch.allclose(feats[1].float().cpu()[:, ::16], torch.from_numpy(g["vq_feat0"]), atol=1e-3)
    out = engine.aft_decoder(z, feats, enc["mid_feat"]).cpu()
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 1e-3


def test_batch_independence(engine):
    """per-sample mean ratio (deformableDecoder_arch.py:567 made per-sample): a batch equals its images run alone"""
    from glare_b200 import synth
    lq, _ = synth.synth_images(3, 32, 48, seed=5)
    lr = synth.preprocess(lq)
    both = engine.infer(lr).cpu()
    for i in range(3):
        one = engine.infer(lr[i:i + 1]).cpu()
        assert float((both[i:i + 1] - one).abs().max()) < 1e-3


@pytest.mark.parametrize("backend,min_agree", [("tc-tf32", 0.90), ("tc-bf16", 0.80)])
def test_reduced_precision_backends_psnr(glare_lib, sd_g, sd_v, backend, min_agree):
    """bf16 / single-pass TF32 operands (BASELINE config 3 is bf16): the index discontinuity rules out a pixel bar;
    the criterion is the north_star's PSNR one (|dPSNR| <= 0.01 dB is for fp32; reduced precision is reported and
    bounded at 0.1 dB) plus index agreement."""
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    from oracle import glare_oracle as O
    eng = GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense(backend))
    g = load_golden("pipe_64x96")
    st = {}
    out = eng.infer(torch.from_numpy(g["lr"]), stages=st).cpu()
    agree = float((st["idx"].cpu().numpy() == g["idx"].astype(np.int64).reshape(-1)).mean())
    gt, ref = torch.from_numpy(g["gt"]), torch.from_numpy(g["out"])
    dpsnr = abs(O.psnr(out.clamp(0, 1), gt) - O.psnr(ref.clamp(0, 1), gt))
    print("%s: idx agree %.4f dPSNR %.4f dB max pixel diff %.4g" % (backend, agree, dpsnr, float((out - ref).abs().max())))
    assert agree >= min_agree and dpsnr < 0.1


def test_stage2_forward_nll_against_reference(glare_lib):
    """BASELINE config 4 (stage-2 flow training), forward half: z and the per-sample NLL of LLFlowVQGAN2.normal_flow
    (LLFlowVQGAN2_arch.py:75-122) against the reference's own outputs (tests/golden/stage2.npz)."""
    from glare_b200 import flow, synth
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    g = load_golden("stage2")
    sd2 = synth.synth_state_dict("netG_stage2", 0)
    assert synth.state_fingerprint(sd2) == pytest.approx(float(g["fingerprint_stage2"]), rel=1e-12)
    eng = GlareEngine(sd2, {}, device="cuda:0", dense=make_dense("tc-3xtf32"), decoders=False)
    lr = torch.from_numpy(g["lr"]).cuda()
    enc = eng.cond_encoder(lr)
    z, logdet = eng.flow_encode(torch.from_numpy(g["gt_latent"]).cuda(), enc["cond_feat"])
    nll = flow.gaussian_nll(z, enc["color_map"], logdet)
    assert torch.allclose(z.cpu(), torch.from_numpy(g["z"]), atol=2e-4, rtol=2e-5)
    assert torch.allclose(nll.cpu(), torch.from_numpy(g["nll"]), atol=1e-4, rtol=1e-5)


def test_infer_rejects_unpadded_sizes(engine):
    with pytest.raises(ValueError):
        engine.infer(torch.zeros((1, 3, 30, 48)))
    with pytest.raises(ValueError):
        engine.infer(torch.zeros((1, 4, 32, 48)))
