"""GPU parity: stats / AdaIN / Welford operators (through the C ABI) vs the oracle and the golden
vectors of the real reference.  Tolerances: statistics 1e-5 relative to the fp64 evaluation of the
reference formula (BASELINE.json), AdaIN outputs 1e-5 of the output scale."""
import numpy as np
import pytest
import torch

import ccst_b200
from ccst_b200 import synth
from oracle import ccst_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda:0"


def rel_err(a, b, floor=1e-6):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs() / (b.abs() + floor)).max().item()


def test_calc_mean_std_golden(golden):
    g = golden["stats"]
    for name in sorted({k.split("/")[0] for k in g.files}):
        x = T(g[name + "/x"]).to(DEV)
        m, s = ccst_b200.calc_mean_std(x)
        assert m.shape == g[name + "/mean"].shape and s.shape == g[name + "/std"].shape
        if name == "s_hw1":
            assert torch.isnan(s).all() and rel_err(m, T(g[name + "/mean64"])) < 1e-6
            continue
        assert rel_err(m, T(g[name + "/mean64"]), 1e-4) < 1e-5, name
        assert rel_err(s, T(g[name + "/std64"])) < 1e-5, name


@pytest.mark.parametrize("shape", [(6, 512, 64, 64), (64, 512, 12, 12), (3, 64, 28, 28), (2, 16, 32, 32),
                                   (2, 7, 9, 13), (1, 3, 100, 171), (2, 4, 128, 128), (1, 2, 200, 300),
                                   (3, 5, 2, 2), (1, 1, 1, 2)])
def test_calc_mean_std_shapes(shape):
    x = synth.features(shape, 100 + shape[2])
    m64, s64 = O.calc_mean_std_f64(x)
    m, s = ccst_b200.calc_mean_std(x.to(DEV))
    assert rel_err(m, m64, 1e-3) < 1e-5
    assert rel_err(s, s64) < 1e-5
    # misaligned view of the same data (forces the streaming kernel)
    flat = torch.empty(x.numel() + 1, device=DEV)
    flat[1:] = x.to(DEV).flatten()
    m2, s2 = ccst_b200.calc_mean_std(flat[1:].view(shape))
    assert rel_err(m2, m64, 1e-3) < 1e-5 and rel_err(s2, s64) < 1e-5


def test_calc_mean_std_full_size_properties():
    """[32,512,64,64] (BASELINE config 3 batch): affine equivariance + agreement with fp64 on GPU."""
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn((32, 512, 64, 64), device=DEV, generator=g).relu_() * 2.0 + 0.25
    m, s = ccst_b200.calc_mean_std(x)
    xd = x[:4].double().flatten(2)
    assert rel_err(m[:4].flatten(), xd.mean(2).flatten(), 1e-3) < 1e-5
    assert rel_err(s[:4].flatten(), (xd.var(2) + 1e-5).sqrt().flatten()) < 1e-5
    m2, s2 = ccst_b200.calc_mean_std(x * 3.0 - 1.5, eps=9e-5)
    assert torch.allclose(m2, m * 3.0 - 1.5, rtol=1e-5, atol=1e-5)
    assert torch.allclose(s2, s * 3.0, rtol=1e-5, atol=1e-6)  # sqrt(9 var + 9 eps)


def test_calc_mean_std_vector_variant():
    x = synth.features((3, 16, 8, 8), 7)
    v = ccst_b200.calc_mean_std_vector(x.to(DEV))
    assert v.shape == (3, 32)
    assert torch.allclose(v.cpu(), O.calc_mean_std_vector(x), rtol=1e-5, atol=1e-6)


def test_adain_golden(golden):
    g = golden["adain"]
    out = ccst_b200.adaptive_instance_normalization(T(g["feat/content"]).to(DEV), T(g["feat/style"]).to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), g["feat/out"], rtol=1e-5, atol=2e-5)
    for tag in ("stat", "stat64"):
        out = ccst_b200.adaIN_StyleStat_ContentFeat(
            T(g[tag + "/content"]).to(DEV), [T(g[tag + "/mean"]).to(DEV), T(g[tag + "/std"]).to(DEV)])
        np.testing.assert_allclose(out.cpu().numpy(), g[tag + "/out"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("shape,alpha", [((2, 512, 64, 64), 1.0), ((2, 512, 64, 64), 0.5), ((8, 512, 12, 12), 1.0),
                                         ((2, 32, 28, 28), 0.3), ((1, 6, 7, 9), 1.0), ((1, 3, 130, 130), 0.8),
                                         ((2, 8, 128, 128), 1.0)])
def test_adain_blend_shapes(shape, alpha):
    x = synth.features(shape, 31 + shape[3])
    n, c = shape[:2]
    gen = torch.Generator().manual_seed(3)
    mu = torch.randn((1, c, 1, 1), generator=gen)
    sg = torch.rand((1, c, 1, 1), generator=gen) + 0.2
    ref = O.adaIN_StyleStat_ContentFeat(x.double(), [mu.double(), sg.double()]) * alpha + x.double() * (1 - alpha)
    out = ccst_b200.adain_blend(x.to(DEV), [mu.to(DEV), sg.to(DEV)], alpha)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    # per-sample style statistics [N,C,1,1]
    mu_n = torch.randn((n, c, 1, 1), generator=gen)
    sg_n = torch.rand((n, c, 1, 1), generator=gen) + 0.2
    ref = O.adaIN_StyleStat_ContentFeat(x.double(), [mu_n.double(), sg_n.double()]) * alpha + x.double() * (1 - alpha)
    out = ccst_b200.adain_blend(x.to(DEV), [mu_n.to(DEV), sg_n.to(DEV)], alpha)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


def test_adain_full_size_properties():
    """[32,512,64,64]: the output statistics equal the style statistics; input untouched."""
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn((32, 512, 64, 64), device=DEV, generator=g) * 1.7 + 0.4
    x0 = x.clone()
    mu = torch.randn((1, 512, 1, 1), device=DEV, generator=g)
    sg = torch.rand((1, 512, 1, 1), device=DEV, generator=g) + 0.5
    out = ccst_b200.adaIN_StyleStat_ContentFeat(x, [mu, sg])
    assert torch.equal(x, x0) and out.data_ptr() != x.data_ptr()
    m, s = ccst_b200.calc_mean_std(out, eps=0.0)
    assert (m - mu).abs().max().item() < 1e-4
    _, s_in = ccst_b200.calc_mean_std(x, eps=0.0)
    _, s_eps = ccst_b200.calc_mean_std(x)
    assert torch.allclose(s, sg * s_in / s_eps, rtol=1e-4)
    # alpha = 0 is the identity
    same = ccst_b200.adain_blend(x, [mu, sg], 0.0)
    assert (same - x).abs().max().item() < 1e-5


def test_adain_feat_mismatched_spatial_and_dead_channels():
    c = synth.features((2, 16, 10, 12), 41)
    s = synth.features((2, 16, 7, 5), 42)
    c[:, 3] = 0.0  # dead channel: std = sqrt(eps), normalised value 0 -> output = style mean
    ref = O.adaptive_instance_normalization(c.double(), s.double())
    out = ccst_b200.adaptive_instance_normalization(c.to(DEV), s.to(DEV))
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5
    assert torch.allclose(out[:, 3].cpu(), O.calc_mean_std(s)[0][:, 3].expand(2, 10, 12), atol=1e-6)


def test_calc_sum_and_accumulation_golden(golden):
    g = golden["overall_stats"]
    st = ccst_b200.WelfordState(12, DEV)
    for i in range(3):
        x = T(g[f"b{i}/x"]).to(DEV)
        s1, s2, cnt = ccst_b200.calc_sum(x)
        assert cnt == int(g[f"b{i}/count"]) and s1.shape == (1, 12, 1, 1)
        x64 = T(g[f"b{i}/x"]).double().transpose(0, 1).reshape(12, -1)
        assert rel_err(s1.flatten(), x64.sum(1), 1e-3) < 1e-5
        assert rel_err(s2.flatten(), (x64 ** 2).sum(1), 1e-3) < 1e-5
        st.add_features(x)
    assert st.count == int(g["final_f64/count"])
    mean, std = st.finalize()
    assert mean.shape == (1, 12, 1, 1)
    assert rel_err(mean, T(g["final_f64/mean"]), 1e-4) < 1e-5
    assert rel_err(std, T(g["final_f64/std"])) < 1e-5


def test_welford_is_cancellation_safe_and_mergeable():
    """Large mean / tiny variance over 2.1 M samples per channel: the reference's fp32
    sum-of-squares formula loses the variance entirely; the Welford state must match
    fp64 at 1e-5, and moments round-trip / shard-merge must be exact."""
    g = torch.Generator(device=DEV).manual_seed(11)
    batches = [torch.randn((8, 64, 64, 64), device=DEV, generator=g) * 0.05 + 100.0 for _ in range(8)]
    allx = torch.cat(batches).double().transpose(0, 1).reshape(64, -1)
    mean64 = allx.mean(1)
    std64 = (allx.var(1, unbiased=False) + 1e-5).sqrt()
    st = ccst_b200.WelfordState(64, DEV)
    for b in batches:
        st.add_features(b)
    mean, std = st.finalize()
    assert rel_err(mean.flatten(), mean64) < 1e-5 and rel_err(std.flatten(), std64) < 1e-5
    # two "ranks" of 4 batches each, merged through the summable moments
    a, b = ccst_b200.WelfordState(64, DEV), ccst_b200.WelfordState(64, DEV)
    for i, x in enumerate(batches):
        (a if i < 4 else b).add_features(x)
    merged = ccst_b200.WelfordState(64, DEV).load_moments(a.moments() + b.moments())
    m2, s2 = merged.finalize()
    assert merged.count == st.count
    assert rel_err(m2.flatten(), mean64) < 1e-5
    # fp64 moments of mean 100 / std 0.05 data keep ~9 digits of the variance
    assert rel_err(s2.flatten(), std64) < 1e-4
    # the oracle (reference formula, fp32) on the same data is far off -- documents H2
    m_ref, s_ref, _, _ = O.overall_style_stats([x.cpu() for x in batches])
    assert rel_err(s_ref.flatten(), std64.cpu()) > 1e-2


def test_ops_reject_cpu_and_bad_dtype():
    with pytest.raises(RuntimeError):
        ccst_b200.calc_mean_std(torch.zeros(1, 2, 4, 4))
    with pytest.raises(TypeError):
        ccst_b200.calc_mean_std(torch.zeros(1, 2, 4, 4, device=DEV, dtype=torch.float16))
    with pytest.raises(AssertionError):
        ccst_b200.adain_blend(torch.zeros(1, 2, 4, 4, device=DEV), [torch.zeros(2, device=DEV)] * 2, alpha=2.0)


def test_calc_mean_std_batch_golden(golden):
    """SURVEY 8 a10 (mean_std_computation_effcientMem.py:89-101): per channel over N*H*W, unbiased."""
    g = golden["io_u8"]
    x = torch.from_numpy(g["batchstat/x"]).to("cuda:0")
    m, s = ccst_b200.calc_mean_std_batch(x)
    assert tuple(m.shape) == tuple(s.shape) == (1, 24, 1, 1)
    assert torch.allclose(m.cpu().double(), torch.from_numpy(g["batchstat/mean64"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(s.cpu().double(), torch.from_numpy(g["batchstat/std64"]), rtol=1e-5, atol=1e-6)


def test_mixstyle_statistics_and_forward_golden(golden):
    """MixStyle (nets/layers.py:46-74): mean / sqrt(unbiased var + 1e-6) and the mixed re-normalisation,
    against the real module's output for the same random draws."""
    g = golden["f4"]
    x = T(g["mix_x"]).to(DEV)
    mu, sig = ccst_b200.mixstyle_stats(x)
    xr = T(g["mix_x"]).double()
    assert torch.allclose(mu.cpu().double(), xr.mean(dim=[2, 3], keepdim=True), rtol=1e-5, atol=1e-6)
    assert torch.allclose(sig.cpu().double(), (xr.var(dim=[2, 3], keepdim=True) + 1e-6).sqrt(), rtol=1e-5, atol=1e-6)
    y = ccst_b200.mixstyle(x, T(g["mix_lmda"]), T(g["mix_perm"]))
    ref = T(g["mix_y"])
    assert (y.cpu() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


def test_mse_loss_matches_torch_and_is_deterministic():
    g = torch.Generator().manual_seed(5)
    for n in (1, 7, 1000, 3 * 512 * 64 * 64 + 5):
        a, b = torch.randn((n,), generator=g), torch.randn((n,), generator=g)
        ref = torch.nn.functional.mse_loss(a.double(), b.double()).item()
        l1 = ccst_b200.mse_loss(a.to(DEV), b.to(DEV)).item()
        l2 = ccst_b200.mse_loss(a.to(DEV), b.to(DEV)).item()
        assert l1 == l2 and abs(l1 - ref) < 1e-6 * max(ref, 1e-6)
    with pytest.raises(AssertionError):
        ccst_b200.mse_loss(torch.zeros(3, device=DEV), torch.zeros(4, device=DEV))


def test_c_abi_allreduce_single_rank_communicator():
    """ccst_allreduce_moments binds NCCL at run time and reduces in place over the caller's ncclComm_t; with one rank
    the sum is the identity (the multi-rank comparison with torch.distributed is tools/multi_gpu_check.py)."""
    from ccst_b200 import _lib, nccl_raw
    torch.cuda.set_device(0)
    try:
        comm = nccl_raw.comm_init(1, 0, nccl_raw.unique_id())
    except RuntimeError as e:  # NCCL's own bootstrap (sockets) is the box's business, not the library's
        pytest.skip(f"no raw NCCL communicator on this box: {e}")
    try:
        x = torch.arange(1026, dtype=torch.float64, device="cuda:0") * 0.5 - 7.0
        ref = x.clone()
        st = torch.cuda.current_stream()
        _lib.check(_lib.lib().ccst_allreduce_moments(comm, x.data_ptr(), x.numel(), st.cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(x, ref)
        with pytest.raises(RuntimeError):
            _lib.check(_lib.lib().ccst_allreduce_moments(None, x.data_ptr(), x.numel(), st.cuda_stream))
    finally:
        nccl_raw.comm_destroy(comm)
