"""CPU: the C-ABI library loads and exports what include/ccst_b200.h declares, the host-side
mirror of the reference interface behaves, and nothing silently falls back to the CPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

import ccst_b200
from ccst_b200 import _lib, net, overall, synth
from oracle import ccst_oracle as O


def test_library_exports_every_header_symbol():
    names = _lib.header_functions()
    assert len(names) >= 20
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in ccst_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    assert handle.ccst_abi_version() == _lib.ABI_VERSION == 2


def test_library_has_no_torch_or_python_dependency():
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


def test_sass_contains_tcgen05_and_tma():
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path present"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    assert _lib.lib().ccst_check_device(0) < 0
    assert b"CPU fallback" in _lib.lib().ccst_last_error() or _lib.lib().ccst_last_error()
    x = synth.features((1, 4, 4, 4), 0)
    with pytest.raises(RuntimeError):
        ccst_b200.calc_mean_std(x)
    with pytest.raises(RuntimeError):
        ccst_b200.adaIN_StyleStat_ContentFeat(x, [x.mean((2, 3), keepdim=True)] * 2)
    with pytest.raises(RuntimeError):
        vgg, dec = synth.make_models(0)
        ccst_b200.style_transfer(vgg, dec, synth.images(1, 16, 16, 0), [x, x])
    # the raw C entry point also refuses instead of computing on the host
    buf = (ctypes.c_float * 64)()
    rc = _lib.lib().ccst_stats_nchw_f32(ctypes.addressof(buf), 4, 16, 1e-5, 1, ctypes.addressof(buf),
                                        ctypes.addressof(buf), None)
    assert rc < 0


def test_reference_asserts_are_kept():
    x3 = torch.zeros(2, 3, 4)
    with pytest.raises(AssertionError):
        ccst_b200.calc_mean_std(x3)  # function.py:7
    with pytest.raises(AssertionError):
        ccst_b200.adaptive_instance_normalization(torch.zeros(1, 4, 2, 2), torch.zeros(1, 5, 2, 2))  # :17
    with pytest.raises(AssertionError):
        ccst_b200.style_transfer(None, None, torch.zeros(1, 3, 8, 8), None, alpha=1.5)  # script :34


def test_net_definitions_match_reference_layout():
    vgg, dec = net.make_vgg(), net.make_decoder()
    assert len(vgg) == 53 and len(dec) == 29
    conv_keys = lambda m: sorted({int(k.split(".")[0]) for k in m.state_dict()})
    assert conv_keys(net.truncate_relu4_1(vgg)) == [0, 2, 5, 9, 12, 16, 19, 22, 25, 29]
    assert conv_keys(dec) == [1, 5, 8, 11, 14, 18, 21, 25, 28]
    n_enc = sum(p.numel() for p in net.truncate_relu4_1(vgg).parameters())
    n_dec = sum(p.numel() for p in dec.parameters())
    assert (n_enc, n_dec) == (3505740, 3505219)
    pools = [m for m in vgg if isinstance(m, torch.nn.MaxPool2d)]
    assert len(pools) == 4 and all(p.ceil_mode for p in pools)
    x = torch.zeros(1, 3, 37, 45)
    assert net.truncate_relu4_1(vgg)(x).shape == (1, 512, 5, 6)
    assert _lib.feature_hw(37, 45) == (5, 6)
    assert _lib.feature_hw(222, 222) == (28, 28) and _lib.feature_hw(512, 512) == (64, 64)
    assert dec(torch.zeros(1, 512, 5, 6)).shape == (1, 3, 40, 48)


def test_shard_range_partitions():
    for total in (0, 1, 7, 2048, 3929):
        for world in (1, 2, 3, 8):
            spans = [overall.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _moments(feat):
    x = feat.double().transpose(0, 1).reshape(feat.shape[1], -1)
    return torch.cat([torch.tensor([float(x.shape[1])], dtype=torch.float64), x.sum(1), (x * x).sum(1)])


def test_finalize_moments_matches_oracle(golden):
    g = golden["overall_stats"]
    batches = [torch.from_numpy(g[f"b{i}/x"]) for i in range(3)]
    m = sum(_moments(b) for b in batches)
    mean, std = overall.finalize_moments(m)
    np.testing.assert_allclose(mean.numpy(), g["final_f64/mean"], rtol=1e-6)
    np.testing.assert_allclose(std.numpy(), g["final_f64/std"], rtol=1e-6)


def test_style_stats_file_format(tmp_path, golden):
    g = golden["overall_stats"]
    mean, std = torch.from_numpy(g["final_f32/mean"]), torch.from_numpy(g["final_f32/std"])
    p = os.path.join(tmp_path, "art_painting_mean_std.npy")
    overall.save_style_stats(p, mean, std)
    raw = np.load(p)
    assert raw.shape == (2, 1, 12, 1, 1) and raw.dtype == np.float32
    np.testing.assert_array_equal(raw, g["npy_payload"])
    back = overall.load_style_stats(p, "cpu")
    assert torch.equal(back[0], mean) and torch.equal(back[1], std)


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 11  # images of one client, sharded contiguously
    b, e = overall.shard_range(total, rank, world)
    feats = synth.features((total, 6, 5, 5), 99)
    local = _moments(feats[b:e]) if e > b else torch.zeros(13, dtype=torch.float64)
    merged = overall.allreduce_moments(local.clone())
    mean, std = overall.finalize_moments(merged)
    torch.save((mean, std, merged), os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_allreduce_of_moments_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0[2], r1[2])  # every rank holds the same merged moments
    feats = synth.features((11, 6, 5, 5), 99)
    mean64, std64, n, _ = O.overall_style_stats([feats], dtype=torch.float64)
    assert r0[2][0].item() == n
    np.testing.assert_allclose(r0[0].numpy(), mean64.numpy(), rtol=1e-6)
    np.testing.assert_allclose(r0[1].numpy(), std64.numpy(), rtol=1e-6)
