"""GPU parity: VectorQuantizer2 lookup through the C ABI vs the golden vectors (reference outputs) and the
CPU oracle.  Bar: indices and z_q BIT-EXACT (north_star)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _run(z, cb):
    from glare_b200 import ops
    packed = ops.vq_pack_codebook(cb.cuda())
    idx, zq = ops.vq_lookup(z.cuda(), packed)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), zq.cpu().numpy()


def test_golden_small_and_large(glare_lib, sd_v):
    g = load_golden("vq")
    cb = sd_v["quantize.embedding.weight"]
    idx, zq = _run(torch.from_numpy(g["z"]), cb)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(zq.view(np.uint32), g["zq"].view(np.uint32))
    idx_l, _ = _run(torch.from_numpy(g["z_large"]), cb)
    assert np.array_equal(idx_l, g["idx_large"].astype(np.int64))


def test_duplicate_codebook_first_index(glare_lib, sd_v):
    g = load_golden("vq")
    cb = sd_v["quantize.embedding.weight"].clone()
    cb[4096:] = cb[:4096]
    idx, _ = _run(torch.from_numpy(g["z"]), cb)
    assert np.array_equal(idx, g["idx_dup"])


@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (2, 3, 7, 5), (3, 3, 33, 31), (1, 3, 105, 155), (4, 3, 104, 152)])
@pytest.mark.parametrize("K", [8192, 1000, 37])
def test_against_oracle_ragged_shapes(glare_lib, shape, K):
    from oracle import vq_lookup
    g = torch.Generator().manual_seed(shape[2] * 1000 + K)
    z = torch.randn(shape, generator=g) * 1.3
    cb = torch.randn((K, 3), generator=g)
    idx, zq = _run(z, cb)
    idx_o, zq_o = vq_lookup(z.numpy(), cb.numpy())
    assert np.array_equal(idx, idx_o)
    assert np.array_equal(zq.view(np.uint32), zq_o.view(np.uint32))


def test_empty_batch_and_special_values(glare_lib):
    from glare_b200 import ops
    from oracle import vq_lookup
    cb = torch.randn((512, 3), generator=torch.Generator().manual_seed(5))
    packed = ops.vq_pack_codebook(cb.cuda())
    idx, zq = ops.vq_lookup(torch.zeros((0, 3, 4, 4), device="cuda"), packed)
    assert idx.numel() == 0 and zq.numel() == 0
    z = torch.randn((1, 3, 4, 8), generator=torch.Generator().manual_seed(6))
    z[0, 0, 0, 0] = float("nan")
    z[0, 1, 0, 1] = float("inf")
    z[0, :, 0, 2] = 1e30
    z[0, :, 0, 3] = 0.0
    idx, _ = _run(z, cb)
    idx_o, _ = vq_lookup(z.numpy(), cb.numpy())
    assert np.array_equal(idx, idx_o)


def test_full_size_properties(glare_lib, sd_v):
    """BASELINE config 2 size (15 x 105 x 155 tokens): idempotence (quantising code vectors returns their own
    index) and agreement with a brute-force fp64 distance except at near ties."""
    from glare_b200 import ops
    cb = sd_v["quantize.embedding.weight"].cuda()
    packed = ops.vq_pack_codebook(cb)
    g = torch.Generator().manual_seed(11)
    z = torch.randn((15, 3, 105, 155), generator=g).cuda()
    idx, zq = ops.vq_lookup(z, packed)
    # z_q = z + (e - z) is within an ulp of the selected code vector
    e = cb[idx].view(15, 105, 155, 3).permute(0, 3, 1, 2)
    assert float((zq - e).abs().max()) < 1e-5
    idx2, _ = ops.vq_lookup(e.contiguous(), packed)
    d_self = ((cb[idx2] - cb[idx]) ** 2).sum(1)
    assert float(d_self.max()) == 0.0                      # duplicates aside, same code vector comes back
    zz = z.permute(0, 2, 3, 1).reshape(-1, 3)[:20000].double()
    d = ((zz[:, None, :] - cb.double()[None]) ** 2).sum(-1)
    best = d.min(1).values
    chosen = d.gather(1, idx[:20000, None])[:, 0]
    assert float((chosen - best).max()) < 1e-5
