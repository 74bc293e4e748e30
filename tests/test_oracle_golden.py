"""CPU: the oracle restatement must reproduce the vectors produced by the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib

import numpy as np
import torch

from oracle import ccst_oracle as O

T = torch.from_numpy


def test_calc_mean_std_matches_reference(golden):
    g = golden["stats"]
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 7
    for name in names:
        x = T(g[name + "/x"])
        m, s = O.calc_mean_std(x)
        np.testing.assert_array_equal(m.numpy(), g[name + "/mean"])
        np.testing.assert_array_equal(s.numpy(), g[name + "/std"])  # NaN == NaN for hw == 1
        m64, s64 = O.calc_mean_std_f64(x)
        np.testing.assert_allclose(m64.numpy(), g[name + "/mean64"], rtol=1e-12)
        np.testing.assert_allclose(s64.numpy(), g[name + "/std64"], rtol=1e-12)
    assert np.isnan(g["s_hw1/std"]).all()  # the reference's HW == 1 behaviour
    assert np.allclose(g["s_const/std"][0, 1], np.sqrt(1e-5))  # dead channel -> sqrt(eps)


def test_adain_matches_reference(golden):
    g = golden["adain"]
    out = O.adaptive_instance_normalization(T(g["feat/content"]), T(g["feat/style"]))
    np.testing.assert_array_equal(out.numpy(), g["feat/out"])
    for tag in ("stat", "stat64"):
        out = O.adaIN_StyleStat_ContentFeat(T(g[tag + "/content"]), [T(g[tag + "/mean"]), T(g[tag + "/std"])])
        np.testing.assert_array_equal(out.numpy(), g[tag + "/out"])


def test_overall_accumulation_matches_reference(golden):
    g = golden["overall_stats"]
    batches = [T(g[f"b{i}/x"]) for i in range(3)]
    for i, b in enumerate(batches):
        s1, s2, cnt = O.calc_sum(b)
        np.testing.assert_array_equal(s1.numpy(), g[f"b{i}/sum"])
        np.testing.assert_array_equal(s2.numpy(), g[f"b{i}/sqsum"])
        assert cnt == int(g[f"b{i}/count"])
    mean, std, n, imgs = O.overall_style_stats(batches)
    np.testing.assert_array_equal(mean.numpy(), g["final_f32/mean"])
    np.testing.assert_array_equal(std.numpy(), g["final_f32/std"])
    assert n == int(g["final_f32/count"]) and imgs == 8
    mean64, std64, _, _ = O.overall_style_stats(batches, dtype=torch.float64)
    np.testing.assert_allclose(mean64.numpy(), g["final_f64/mean"], rtol=1e-12)
    np.testing.assert_allclose(std64.numpy(), g["final_f64/std"], rtol=1e-12)
    payload = O.pack_style_npy(mean, std)
    assert payload.shape == (2, 1, 12, 1, 1) and payload.dtype == np.float32
    np.testing.assert_array_equal(payload, g["npy_payload"])


def test_synthetic_weights_are_reproducible(golden, models):
    vgg, dec = models
    h = hashlib.sha256()
    for m in (vgg, dec):
        for k, v in m.state_dict().items():
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes())
    assert h.hexdigest() == bytes(golden["net"]["weights_sha256"]).decode()


def test_network_and_style_transfer_match_reference(golden, models):
    g = golden["net"]
    vgg, dec = models
    torch.set_num_threads(1)
    with torch.no_grad():
        for tag in ("sq40", "odd37x45", "r96"):
            x = T(g[tag + "/x"])
            f = O.encode_relu4_1(vgg, x)
            np.testing.assert_allclose(f.numpy(), g[tag + "/feat"], rtol=0, atol=2e-5)
            ms, ss = O.single_style_stats(O.encode_relu4_1(vgg, T(g[tag + "/style_img"])))
            np.testing.assert_allclose(ms.numpy(), g[tag + "/style_mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(ss.numpy(), g[tag + "/style_std"], rtol=1e-5, atol=1e-6)
            stat = [T(g[tag + "/style_mean"]), T(g[tag + "/style_std"])]
            np.testing.assert_allclose(O.decode(dec, T(g[tag + "/feat"])).numpy(), g[tag + "/dec_of_feat"],
                                       rtol=0, atol=2e-5)
            for alpha in (1.0, 0.6):
                out = O.style_transfer(vgg, dec, x, stat, alpha)
                np.testing.assert_allclose(out.numpy(), g[f"{tag}/out_a{alpha}"], rtol=0, atol=5e-5)
            if x.shape[0] >= 2:
                out = O.style_transfer(vgg, dec, x, stat, 1.0, [0.25, 0.75])
                assert out.shape[0] == 1
                np.testing.assert_allclose(out.numpy(), g[tag + "/out_interp"], rtol=0, atol=5e-5)
    # shapes the survey lists (222 -> 28 -> 224)
    assert O.encode_relu4_1(vgg, torch.zeros(1, 3, 37, 45)).shape == (1, 512, 5, 6)


def test_image_style_form_equals_stat_form(models):
    """style given as images == per-sample calc_mean_std of the encoded style (function.py:16-24)."""
    from ccst_b200 import synth

    vgg, dec = models
    x, s = synth.images(2, 32, 32, 1), synth.images(2, 40, 32, 2)
    with torch.no_grad():
        a = O.style_transfer_image_style(vgg, dec, x, s, 0.7)
        stat = O.calc_mean_std(O.encode_relu4_1(vgg, s))
        b = O.style_transfer(vgg, dec, x, stat, 0.7)
    assert torch.allclose(a, b, atol=1e-6)


def test_image_io_matches_torchvision_and_reference(golden, models):
    """ToTensor of the loader's uint8 image and save_image's quantisation, pinned to the
    REAL torchvision transforms / save_image (PNG round trip) run on the reference's outputs."""
    g = golden["io_u8"]
    x_u8 = T(g["x_u8"])
    np.testing.assert_array_equal(O.to_tensor_u8(x_u8).numpy(), g["x_tensor"])
    assert g["x_u8"].min() == 0 and g["x_u8"].max() == 255
    for a in ("1.0", "0.5"):
        q = O.save_image_batch_u8(T(g[f"out_f32_a{a}"]))
        np.testing.assert_array_equal(q.numpy(), g[f"out_u8_a{a}"])
    # both clamps of save_image are exercised by the stretched alpha = 0.5 vector
    assert (g["out_u8_a0.5"] == 0).sum() > 100 and (g["out_u8_a0.5"] == 255).sum() > 100
    # and the float vectors are the reference's style_transfer on the ToTensor'd input
    vgg, dec = models
    with torch.no_grad():
        o = O.style_transfer(vgg, dec, O.to_tensor_u8(x_u8), [T(g["style_mean"]), T(g["style_std"])], 1.0)
    np.testing.assert_allclose(o.numpy(), g["out_f32_a1.0"], rtol=0, atol=1e-6)


def test_resize_output_matches_torchvision(golden):
    """`transforms.Resize(args.output_size)` on the stylised batch (CCST_OverallStyleTransfer.py:154-155)."""
    g = golden["io_u8"]
    o = T(g["out_f32_a1.0"])
    for tag, size in (("s24", 24), ("s17", 17), ("s64", 64), ("hw", (20, 31))):
        r = O.resize_output(o, size)
        np.testing.assert_array_equal(r.numpy(), g[f"resize_{tag}"])
        np.testing.assert_array_equal(O.save_image_batch_u8(r).numpy(), g[f"resize_{tag}_u8"])
    np.testing.assert_array_equal(O.resize_output(torch.rand((1, 3, 512, 512), generator=torch.Generator().manual_seed(77)), 96).numpy(), g["resize_512_to_96"])


def test_batch_calc_mean_std_matches_reference(golden):
    """the second calc_mean_std of mean_std_computation_effcientMem.py:89-101."""
    g = golden["io_u8"]
    m, s = O.calc_mean_std_batch(T(g["batchstat/x"]))
    np.testing.assert_array_equal(m.numpy(), g["batchstat/mean"])
    np.testing.assert_array_equal(s.numpy(), g["batchstat/std"])


def test_net_forward_losses_and_mixstyle_match_reference(golden, models):
    """SURVEY 8f rank 4 (forward only): the oracle's restatement of Net.forward (net.py:138-152) and of
    MixStyle.forward (nets/layers.py:46-74) against outputs of the real classes."""
    g = golden["f4"]
    vgg, dec = models
    content, style = torch.from_numpy(g["content"]), torch.from_numpy(g["style"])
    with torch.no_grad():
        for alpha in (1.0, 0.7):
            lc, ls = O.net_forward_losses(vgg, dec, content, style, alpha)
            assert abs(lc.item() - float(g[f"loss_c_a{alpha}"])) < 1e-6 * max(1.0, abs(lc.item()))
            assert abs(ls.item() - float(g[f"loss_s_a{alpha}"])) < 1e-6 * max(1.0, abs(ls.item()))
        feats = O.encode_with_intermediate(vgg, style)
        assert [f.shape[1] for f in feats] == [64, 128, 256, 512]
        for i, f in enumerate(feats):
            m, s = O.calc_mean_std(f)
            np.testing.assert_allclose(m.numpy(), g[f"style_level{i}_mean"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(s.numpy(), g[f"style_level{i}_std"], rtol=1e-6, atol=1e-7)
    y = O.mixstyle_forward(torch.from_numpy(g["mix_x"]), torch.from_numpy(g["mix_lmda"]), torch.from_numpy(g["mix_perm"]))
    np.testing.assert_array_equal(y.numpy(), g["mix_y"])


def test_pil_resize_matches_real_pillow(golden):
    """The loader's Resize((S, S)) (data_helper.py:45-49): the oracle's restatement of Pillow's 8-bit bilinear
    resample against the real torchvision / Pillow output, bit for bit."""
    import hashlib
    g = golden["io_u8"]
    for tag, S in (("up", 64), ("down", 48), ("pacs", 512)):
        x = T(g[f"resize_in_{tag}/x"])[None]
        y = O.pil_resize_bilinear_u8(x, S, S)[0].numpy()
        if tag == "pacs":
            assert hashlib.sha256(y.tobytes()).hexdigest().encode() == g[f"resize_in_{tag}/sha256"].tobytes()
            np.testing.assert_array_equal(y[100], g[f"resize_in_{tag}/row100"])
        else:
            np.testing.assert_array_equal(y, g[f"resize_in_{tag}/y"])


def test_image_list_reader(tmp_path):
    """`_dataset_info` (cjm_util/ImageLoader.py:31-42): "<path> <label>" per line."""
    from ccst_b200 import data
    p = tmp_path / "art_painting_train.txt"
    p.write_text("/disk1/pacs/art_painting/dog/pic_001.jpg 1\n/disk1/pacs/art_painting/house/pic_010.jpg 6\n")
    names, labels = data.dataset_info(str(p))
    assert names == ["/disk1/pacs/art_painting/dog/pic_001.jpg", "/disk1/pacs/art_painting/house/pic_010.jpg"]
    assert labels == [1, 6]
